// main.cpp -- drop-in for the reference's `main` (src/main.cpp) on top of the C ABI (ftrl_b200.h).
// C++17, no CUDA headers: everything device-side goes through libftrl_b200.so.
//
//   offline (--online false): src/task/ftrl_offline.cpp -- load the whole file, shuffle the sample
//            order every epoch (:67-71), train, print `epoch %d train time ...`, evaluate (:56-61)
//   online  (--online true, the default): src/task/ftrl_online.cpp -- stream the file in order every
//            epoch (producer/consumer of src/concurrent/pc_task.cpp -> chunked read + parser threads),
//            evaluate through the streaming Evaluator (src/eval/evaluate.cpp)
// The reference's per-sample worker loop (ftrl_offline.cpp:74-83) becomes: pack `batch_size` samples
// into a pinned CSR buffer, ftrl_train_batch (asynchronous: the copy of batch i+1 and the parsing of
// the next chunk overlap the kernels of batch i), sum the fp64 losses at the epoch barrier (ftrl_sync,
// the ThreadPool::synchronize of thread_pool.h:82-88).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <future>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "ftrl_b200.h"
#include "options.hpp"
#include "parser.hpp"

namespace {

using clk = std::chrono::steady_clock;
double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

[[noreturn]] void die(ftrl_handle *h, const char *what) {
  fprintf(stderr, "%s: %s\n", what, ftrl_last_error(h));
  exit(EXIT_FAILURE);
}

// pinned CSR staging buffers; a buffer handed to call i is reusable once call i+3 has returned
struct PinnedCsr {
  int64_t *row_ptr = nullptr;
  int32_t *field = nullptr, *feat = nullptr, *label = nullptr;
  float *val = nullptr;
  size_t cap_rows = 0, cap_nnz = 0;
  void ensure(size_t rows, size_t nnz) {
    if (rows > cap_rows) {
      ftrl_free_pinned(row_ptr);
      ftrl_free_pinned(label);
      cap_rows = rows + rows / 4 + 16;
      row_ptr = (int64_t *)ftrl_alloc_pinned(sizeof(int64_t) * (cap_rows + 1));
      label = (int32_t *)ftrl_alloc_pinned(sizeof(int32_t) * cap_rows);
    }
    if (nnz > cap_nnz) {
      ftrl_free_pinned(field);
      ftrl_free_pinned(feat);
      ftrl_free_pinned(val);
      cap_nnz = nnz + nnz / 4 + 64;
      field = (int32_t *)ftrl_alloc_pinned(sizeof(int32_t) * cap_nnz);
      feat = (int32_t *)ftrl_alloc_pinned(sizeof(int32_t) * cap_nnz);
      val = (float *)ftrl_alloc_pinned(sizeof(float) * cap_nnz);
    }
    if (!row_ptr || !label || (cap_nnz && (!field || !feat || !val))) {
      fprintf(stderr, "pinned host allocation failed\n");
      exit(EXIT_FAILURE);
    }
  }
};

// One model on one GPU, or (--n_gpus N) one model whose tables are sharded by feature id over N GPUs of this box:
// N handles in this process, peers attached in-process (ftrl_export_peer_blob / ftrl_attach_peers), every minibatch
// split into N contiguous shares -- the reference's worker fan-out (ftrl_offline.cpp:85-91) across GPUs.
class Trainer {
 public:
  // max_row_nnz: longest sample of the data (sizes the fixed buffers of a multi-GPU run; ignored for one GPU)
  Trainer(const host::Options &o, size_t max_row_nnz = 0) : opt_(o) {
    ftrl_config c;
    ftrl_config_default(&c);
    if (o.model_type == "LR") c.model_type = FTRL_LR;
    else if (o.model_type == "FM") c.model_type = FTRL_FM;
    else if (o.model_type == "FFM") c.model_type = FTRL_FFM;
    else {
      // ftrl_offline.cpp:29-32
      fprintf(stderr, "Invalid model_type: %s, expect `LR`, `FM` or `FFM`.\n", o.model_type.c_str());
      throw std::invalid_argument("invalid model_type");
    }
    c.n_feats = o.n_feats;
    c.n_fields = o.n_fields;
    c.n_factors = o.n_factors;
    c.init_mean = o.init_mean;
    c.init_stddev = o.init_stddev;
    c.w_alpha = o.w_alpha;
    c.w_beta = o.w_beta;
    c.w_l1 = o.w_l1;
    c.w_l2 = o.w_l2;
    c.mode = o.batch_size == 1 ? FTRL_MODE_SEQUENTIAL : FTRL_MODE_BATCH;
    c.seed = o.seed ? o.seed : std::random_device{}();
    // sequential mode walks the samples of a call in order on the device: hand over large blocks
    block_ = o.batch_size == 1 ? 8192 : (size_t)o.batch_size;
    const int G = std::max(1, o.n_gpus);
    ranks_.resize((size_t)G);
    for (int r = 0; r < G; r++) {
      c.device = o.device + r;
      c.rank = r;
      c.world_size = G;
      if (G > 1) {
        c.max_batch_rows = (int64_t)((block_ + G - 1) / G);
        c.max_batch_nnz = c.max_batch_rows * (int64_t)std::max<size_t>(1, max_row_nnz);
      }
      if (ftrl_create(&c, &ranks_[r].h) != FTRL_OK) die(nullptr, "ftrl_create");
    }
    if (G > 1) {
      std::vector<char> blobs((size_t)G * FTRL_PEER_BLOB_BYTES);
      for (int r = 0; r < G; r++)
        if (ftrl_export_peer_blob(ranks_[r].h, blobs.data() + (size_t)r * FTRL_PEER_BLOB_BYTES) != FTRL_OK)
          die(ranks_[r].h, "ftrl_export_peer_blob");
      for (int r = 0; r < G; r++)
        if (ftrl_attach_peers(ranks_[r].h, blobs.data()) != FTRL_OK) die(ranks_[r].h, "ftrl_attach_peers");
    }
  }
  ~Trainer() {
    for (auto &r : ranks_) ftrl_destroy(r.h);
  }

  // trains (or scores) samples idx[begin..end) of `data`: rank r takes the r-th contiguous share
  void submit(const host::Csr &data, const int32_t *order, size_t begin, size_t end, bool train) {
    const size_t G = ranks_.size(), total = end - begin, share = (total + G - 1) / G;
    // the library writes through these pointers at ftrl_sync: never reallocate with calls in flight
    if (loss_parts_.size() + G >= loss_parts_.capacity()) sync();
    const int pin = next_pin_;
    next_pin_ = (next_pin_ + 1) % kPins;
    for (size_t g = 0; g < G; g++) {
      const size_t b0 = std::min(end, begin + g * share), b1 = std::min(end, b0 + share);
      Rank &rk = ranks_[g];
      PinnedCsr &p = rk.pin[pin];
      const size_t rows = b1 - b0;
      size_t nnz = 0;
      if (order) {
        for (size_t i = b0; i < b1; i++) nnz += (size_t)(data.row_ptr[order[i] + 1] - data.row_ptr[order[i]]);
      } else {
        nnz = (size_t)(data.row_ptr[b1] - data.row_ptr[b0]);
      }
      p.ensure(rows, nnz);
      size_t w = 0;
      p.row_ptr[0] = 0;
      if (!order && rows) {  // file order: the share is one contiguous span of every array
        const size_t a = (size_t)data.row_ptr[b0];
        memcpy(p.field, data.field.data() + a, nnz * sizeof(int32_t));
        memcpy(p.feat, data.feat.data() + a, nnz * sizeof(int32_t));
        memcpy(p.val, data.val.data() + a, nnz * sizeof(float));
        memcpy(p.label, data.label.data() + b0, rows * sizeof(int32_t));
        for (size_t i = 0; i < rows; i++) p.row_ptr[i + 1] = data.row_ptr[b0 + i + 1] - (int64_t)a;
      }
      for (size_t i = b0; order && i < b1; i++) {
        const size_t r = order ? (size_t)order[i] : i;
        const size_t a = (size_t)data.row_ptr[r], n = (size_t)data.row_ptr[r + 1] - a;
        memcpy(p.field + w, data.field.data() + a, n * sizeof(int32_t));
        memcpy(p.feat + w, data.feat.data() + a, n * sizeof(int32_t));
        memcpy(p.val + w, data.val.data() + a, n * sizeof(float));
        w += n;
        p.row_ptr[i - b0 + 1] = (int64_t)w;
        p.label[i - b0] = data.label[r];
      }
      loss_parts_.push_back(0.0);
      double *slot = &loss_parts_.back();
      int rc;
      if (train) {
        // a multi-GPU step is collective: every rank is called, also with an empty share
        rc = ftrl_train_batch(rk.h, (int64_t)rows, p.row_ptr, p.field, p.feat, p.val, p.label, nullptr, slot);
      } else if (rows == 0) {
        rc = FTRL_OK;
      } else {
        if (rk.sink[pin].size() < share) rk.sink[pin].resize(std::max(share, block_));
        float *out = rk.sink[pin].data();  // predictions are not needed by the driver (only the loss) ...
        if (opt_.auc) {  // ... unless the AUC is asked for: one block per call, never moved while in flight
          scores_.emplace_back(rows);
          out = scores_.back().data();
          labels_.emplace_back(p.label, p.label + rows);
        }
        rc = ftrl_predict_batch(rk.h, (int64_t)rows, p.row_ptr, p.field, p.feat, p.val, p.label, 0, out, slot);
      }
      if (rc != FTRL_OK) die(rk.h, train ? "ftrl_train_batch" : "ftrl_predict_batch");
    }
    n_submitted_ += total;
  }

  void run_block(const host::Csr &data, const int32_t *order, size_t n, bool train) {
    for (size_t b = 0; b < n; b += block_) submit(data, order, b, std::min(n, b + block_), train);
  }

  void sync() {
    for (auto &r : ranks_)
      if (ftrl_sync(r.h) != FTRL_OK) die(r.h, "ftrl_sync");
    for (double v : loss_parts_) loss_total_ += v;
    loss_parts_.clear();
  }

  // mean loss since the last call (ftrl_offline.cpp:101-102, ftrl_online.cpp:82-94)
  double take_loss() {
    sync();
    const double r = n_submitted_ ? loss_total_ / (double)n_submitted_ : 0.0;
    loss_total_ = 0.0;
    n_submitted_ = 0;
    return r;
  }

  // ROC AUC over the scores collected since the last call (evaluation passes with --auc true); NaN if none
  double take_auc() {
    sync();
    std::vector<float> sc;
    std::vector<int32_t> la;
    for (auto &b : scores_) sc.insert(sc.end(), b.begin(), b.end());
    for (auto &b : labels_) la.insert(la.end(), b.begin(), b.end());
    scores_.clear();
    labels_.clear();
    double auc = 0.0;
    if (ftrl_eval_auc(handle(), (int64_t)sc.size(), sc.data(), la.data(), &auc) != FTRL_OK) die(handle(), "ftrl_eval_auc");
    return auc;
  }

  void begin_epoch(size_t expected_calls) {
    loss_parts_.reserve(std::max<size_t>((expected_calls + 8) * ranks_.size(), 4096));
  }
  // the handle model files are written through (a multi-GPU run: one rank reads every shard over peer memory)
  ftrl_handle *handle() { return ranks_[0].h; }
  int n_gpus() const { return (int)ranks_.size(); }

 private:
  static constexpr int kPins = 4;
  struct Rank {
    ftrl_handle *h = nullptr;
    PinnedCsr pin[kPins];
    std::vector<float> sink[kPins];
  };
  host::Options opt_;
  std::vector<Rank> ranks_;
  size_t block_ = 1024;
  int next_pin_ = 0;
  std::vector<double> loss_parts_;
  std::deque<std::vector<float>> scores_;
  std::deque<std::vector<int32_t>> labels_;
  double loss_total_ = 0.0;
  size_t n_submitted_ = 0;
};

size_t max_row_nnz(const host::Csr &c) {
  size_t m = 0;
  for (size_t r = 0; r + 1 < c.row_ptr.size(); r++) m = std::max(m, (size_t)(c.row_ptr[r + 1] - c.row_ptr[r]));
  return m;
}

// Reader::load_from_file (src/data/reader.cpp:50-91)
void load_file(const std::string &path, bool libffm, int n_threads, bool csr_cache, host::Csr &out) {
  printf("Loading data from file: %s\n", path.c_str());
  const auto t0 = clk::now();
  if (csr_cache && host::load_csr_cache(path, libffm, out)) {
    printf("Total number of samples loaded: %zu\n", out.rows());
    printf("parsing data time: %.4lfs (binary image %s.csr)\n", since(t0), path.c_str());
    return;
  }
  std::vector<char> buf;
  if (!host::read_file(path, buf)) {
    fprintf(stderr, "fail to open %s\n", path.c_str());
    exit(EXIT_FAILURE);
  }
  host::parse_buffer(buf.data(), buf.size(), libffm, n_threads, out);
  printf("Total number of samples loaded: %zu\n", out.rows());
  printf("parsing data time: %.4lfs\n", since(t0));
  if (csr_cache && !host::save_csr_cache(path, libffm, out)) fprintf(stderr, "could not write %s.csr\n", path.c_str());
}

// FtrlOffline::train / evaluate / one_epoch (src/task/ftrl_offline.cpp:44-103)
void run_offline(const host::Options &o) {
  const bool libffm = o.file_type == "libffm";
  host::Csr train, eval;
  load_file(o.train_path, libffm, o.thread_num, o.csr_cache, train);
  if (!o.eval_path.empty()) load_file(o.eval_path, libffm, o.thread_num, o.csr_cache, eval);
  Trainer tr(o, std::max(max_row_nnz(train), max_row_nnz(eval)));
  std::mt19937 gen(o.seed ? (uint32_t)o.seed : std::random_device{}());
  std::vector<int32_t> order(train.rows());
  for (int ep = 1; ep <= o.epoch; ep++) {
    const auto t0 = clk::now();
    std::iota(order.begin(), order.end(), 0);
    std::shuffle(order.begin(), order.end(), gen);  // ftrl_offline.cpp:69-71
    tr.begin_epoch(train.rows() / std::max<long>(1, o.batch_size) + 1);
    tr.run_block(train, order.data(), train.rows(), true);
    const double loss = tr.take_loss();
    printf("epoch %d train time: %.4lfs, train loss: %.4lf\n", ep, since(t0), loss);
    if (!o.eval_path.empty()) {
      const auto t1 = clk::now();
      tr.run_block(eval, nullptr, eval.rows(), false);
      const double el = tr.take_loss();
      printf("epoch %d eval time: %.4lfs, eval loss: %.4lf\n", ep, since(t1), el);
      if (o.auc) printf("epoch %d eval auc: %.6lf\n", ep, tr.take_auc());
    }
  }
  if (!o.model_path.empty() && ftrl_save_model(tr.handle(), o.model_path.c_str(), 10) != FTRL_OK)
    die(tr.handle(), "ftrl_save_model");
}

// one streaming pass over a file in file order: PcTask::run (src/concurrent/pc_task.cpp:22-80) with
// FtrlOnline::run_task (ftrl_online.cpp:70-80) / Evaluator::run_task (evaluate.cpp:23-33) as consumer.
// Like the reference's producer / consumer pair, reading + parsing of block i+1 (a producer task on the parser
// threads) overlaps the submission and the GPU work of block i (this thread).
double stream_file(Trainer &tr, const std::string &path, bool libffm, int n_threads, bool train) {
  host::TextBlockReader reader(path, libffm, n_threads);
  if (!reader.ok()) {
    fprintf(stderr, "open file <%s> error. \n", path.c_str());  // pc_task.cpp:8
    exit(EXIT_FAILURE);
  }
  // producer: next block of complete lines -> CSR; false when the file is exhausted
  auto produce = [&](host::Csr *out) -> bool { return reader.next(*out); };
  long lines = 0, next_log = 1000000;
  host::Csr blocks[2];
  tr.begin_epoch(1u << 16);
  int cur = 0;
  std::future<bool> next = std::async(std::launch::async, produce, &blocks[cur]);
  while (next.get()) {
    host::Csr &csr = blocks[cur];
    cur ^= 1;
    next = std::async(std::launch::async, produce, &blocks[cur]);  // parse the next block while this one trains
    tr.run_block(csr, nullptr, csr.rows(), train);  // copies into pinned staging
    lines += (long)csr.rows();
    while (lines >= next_log) {  // pc_task.cpp:47-49
      printf("%ld lines finished...\n", next_log);
      next_log += 1000000;
    }
  }
  return tr.take_loss();
}

// FtrlOnline::train / evaluate (src/task/ftrl_online.cpp:42-68)
void run_online(const host::Options &o) {
  const bool libffm = o.file_type == "libffm";
  if (o.cmd) return;  // `// todo: online learning` in the reference (ftrl_online.cpp:55-57)
  // --csr_cache: the files are parsed (or their binary images read) once and every epoch walks them in file
  // order from memory; otherwise each epoch streams and parses the text again like the reference
  host::Csr train, eval;
  if (o.csr_cache) {
    load_file(o.train_path, libffm, o.thread_num, true, train);
    if (!o.eval_path.empty()) load_file(o.eval_path, libffm, o.thread_num, true, eval);
  } else if (o.n_gpus > 1) {
    // the buffers peers map are sized at creation from the longest sample: the data must be known up front
    fprintf(stderr, "--n_gpus > 1 needs the data in memory: use --online false or --csr_cache true\n");
    exit(EXIT_FAILURE);
  }
  Trainer tr(o, std::max(max_row_nnz(train), max_row_nnz(eval)));
  auto one_pass = [&](const host::Csr &mem, const std::string &path, bool is_train) {
    if (!o.csr_cache) return stream_file(tr, path, libffm, o.thread_num, is_train);
    tr.begin_epoch(mem.rows() / std::max<long>(1, o.batch_size) + 1);
    tr.run_block(mem, nullptr, mem.rows(), is_train);
    return tr.take_loss();
  };
  for (int ep = 1; ep <= o.epoch; ep++) {
    const auto t0 = clk::now();
    const double loss = one_pass(train, o.train_path, true);
    printf("epoch %d train time: %.4lfs, train loss: %.4lf\n", ep, since(t0), loss);
    if (!o.eval_path.empty()) {
      const auto t1 = clk::now();
      const double el = one_pass(eval, o.eval_path, false);
      printf("epoch %d eval time: %.4lfs, eval loss: %.4lf\n", ep, since(t1), el);
      if (o.auc) printf("epoch %d eval auc: %.6lf\n", ep, tr.take_auc());
    }
  }
  if (!o.model_path.empty() && ftrl_save_model(tr.handle(), o.model_path.c_str(), 10) != FTRL_OK)
    die(tr.handle(), "ftrl_save_model");
}

}  // namespace

int main(int argc, char *argv[]) {
  host::Options opt;
  try {
    host::parse_options(argc, argv, opt);
  } catch (const std::invalid_argument &e) {  // src/main.cpp:18-24
    fprintf(stderr, "invalid argument: %s\n", e.what());
    fputs(host::kHelp, stdout);
    exit(EXIT_FAILURE);
  }
  try {
    if (opt.online) run_online(opt); else run_offline(opt);
  } catch (const std::invalid_argument &e) {
    fprintf(stderr, "%s\n", e.what());
    return EXIT_FAILURE;
  }
  return 0;
}
