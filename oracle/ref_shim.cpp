// ref_shim.cpp -- C interface around the UNMODIFIED reference classes
// ftrl::LR / ftrl::FM / ftrl::FFM (compiled from /root/reference/src/model/*.cpp
// where they lie; see oracle/Makefile, target `ref`).  TEST INFRASTRUCTURE ONLY.
//
// Purpose: (1) pin oracle/ftrl_oracle.c bit-for-bit, (2) generate tests/golden/
// fixtures, (3) serve as the CPU arm of bench.py (`cpu_baseline.kind =
// "reference"`), running the reference's own train() from n_threads workers
// with the chunking of FtrlOffline::one_epoch (src/task/ftrl_offline.cpp:63-103).
//
// The reference keeps n/z private/protected; the `#define private public`
// below gives this translation unit access without touching the sources
// (member order, hence layout, is unchanged -- SURVEY.md section 7 step 1).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <regex>
#include <fstream>
#include <iostream>
#include <sstream>
#include <shared_mutex>
#include <string>
#include <string_view>
#include <thread>
#include <tuple>
#include <vector>

#define private public
#define protected public
#include "data/parser.h"
#include "data/sample.h"
#include "eval/loss.h"
#include "model/ffm.h"
#include "model/fm.h"
#include "model/lr.h"
#undef private
#undef protected

namespace {

struct Ref {
  int model_type;
  int n_feats, n_fields, k;
  std::unique_ptr<ftrl::FtrlModel> model;
  std::vector<Sample> staged;  // samples converted from CSR for the threaded run
  float bias3[3];
};

ftrl::LR *as_lr(Ref *r) { return dynamic_cast<ftrl::LR *>(r->model.get()); }
ftrl::FM *as_fm(Ref *r) { return dynamic_cast<ftrl::FM *>(r->model.get()); }
ftrl::FFM *as_ffm(Ref *r) { return dynamic_cast<ftrl::FFM *>(r->model.get()); }

feat_vec make_sample(int nnz, const int32_t *field, const int32_t *feat, const float *val) {
  feat_vec x;
  x.reserve(nnz);
  for (int t = 0; t < nnz; t++) x.emplace_back(field[t], feat[t], val[t]);
  return x;
}

std::vector<std::vector<float>> *vec_table(Ref *r, int which) {
  if (r->model_type == 1) {
    auto *m = as_fm(r);
    return which == 0 ? &m->vec_w : which == 1 ? &m->vec_w_n : &m->vec_w_z;
  }
  if (r->model_type == 2) {
    auto *m = as_ffm(r);
    return which == 0 ? &m->vec_w : which == 1 ? &m->vec_w_n : &m->vec_w_z;
  }
  return nullptr;
}

std::vector<float> *lin_table(Ref *r, int which) {
  auto *m = r->model.get();
  return which == 0 ? &m->lin_w : which == 1 ? &m->lin_w_n : &m->lin_w_z;
}

}  // namespace

extern "C" {

// fast_init = 0: run the reference constructor at full size (Gaussian w via a
// fresh std::random_device per weight, ~7.6 us per weight).
// fast_init = 1: run the reference constructor at n_feats = 1, then grow the
// reference's own containers to n_feats with zero-filled w.  train()/predict()
// executed afterwards are the untouched reference code either way.
void *ftrl_ref_create(int model_type, int n_feats, int n_fields, int n_factors, float w_alpha,
                      float w_beta, float w_l1, float w_l2, int fast_init) {
  config_options opt;
  opt.model_type = model_type == 0 ? "LR" : model_type == 1 ? "FM" : "FFM";
  opt.n_feats = fast_init ? 1 : n_feats;
  opt.n_fields = n_fields;
  opt.n_factors = n_factors;
  opt.w_alpha = w_alpha;
  opt.w_beta = w_beta;
  opt.w_l1 = w_l1;
  opt.w_l2 = w_l2;
  auto *r = new Ref();
  r->model_type = model_type;
  r->n_feats = n_feats;
  r->n_fields = n_fields;
  r->k = n_factors;
  if (model_type == 0)
    r->model = std::make_unique<ftrl::LR>(opt);
  else if (model_type == 1)
    r->model = std::make_unique<ftrl::FM>(opt);
  else
    r->model = std::make_unique<ftrl::FFM>(opt);
  if (fast_init) {
    auto *m = r->model.get();
    m->n_feats = n_feats;
    m->lin_w.assign(n_feats, 0.0f);
    m->lin_w_n.assign(n_feats, 0.0f);
    m->lin_w_z.assign(n_feats, 0.0f);
    m->lin_w_mutex = std::vector<std::mutex>(n_feats);
    const size_t row = model_type == 1 ? (size_t)n_factors : (size_t)n_fields * n_factors;
    if (model_type == 1) {
      auto *fm = as_fm(r);
      fm->vec_w.assign(n_feats, std::vector<float>(row, 0.0f));
      fm->vec_w_n.assign(n_feats, std::vector<float>(row, 0.0f));
      fm->vec_w_z.assign(n_feats, std::vector<float>(row, 0.0f));
      fm->vec_w_mutex = std::vector<std::shared_mutex>(n_feats);
    } else if (model_type == 2) {
      auto *ffm = as_ffm(r);
      ffm->vec_w.assign(n_feats, std::vector<float>(row, 0.0f));
      ffm->vec_w_n.assign(n_feats, std::vector<float>(row, 0.0f));
      ffm->vec_w_z.assign(n_feats, std::vector<float>(row, 0.0f));
      ffm->vec_w_mutex = std::vector<std::shared_mutex>(n_feats);
    }
  }
  return r;
}

void ftrl_ref_destroy(void *h) { delete static_cast<Ref *>(h); }

int64_t ftrl_ref_row_len(void *h) {
  auto *r = static_cast<Ref *>(h);
  return r->model_type == 1 ? r->k : r->model_type == 2 ? (int64_t)r->n_fields * r->k : 0;
}

// bias triple {bias, bias_n, bias_z}
void ftrl_ref_get_bias(void *h, float *out3) {
  auto *m = static_cast<Ref *>(h)->model.get();
  out3[0] = m->bias;
  out3[1] = m->bias_n;
  out3[2] = m->bias_z;
}
void ftrl_ref_set_bias(void *h, const float *in3) {
  auto *m = static_cast<Ref *>(h)->model.get();
  m->bias = in3[0];
  m->bias_n = in3[1];
  m->bias_z = in3[2];
}
// which: 0 = w, 1 = n, 2 = z
void ftrl_ref_get_lin(void *h, int which, float *out) {
  auto *t = lin_table(static_cast<Ref *>(h), which);
  std::copy(t->begin(), t->end(), out);
}
void ftrl_ref_set_lin(void *h, int which, const float *in) {
  auto *t = lin_table(static_cast<Ref *>(h), which);
  std::copy(in, in + t->size(), t->begin());
}
void ftrl_ref_get_vec(void *h, int which, float *out) {
  auto *t = vec_table(static_cast<Ref *>(h), which);
  if (!t) return;
  for (auto &row : *t) out = std::copy(row.begin(), row.end(), out);
}
void ftrl_ref_set_vec(void *h, int which, const float *in) {
  auto *t = vec_table(static_cast<Ref *>(h), which);
  if (!t) return;
  for (auto &row : *t) {
    std::copy(in, in + row.size(), row.begin());
    in += row.size();
  }
}

float ftrl_ref_train(void *h, int nnz, const int32_t *field, const int32_t *feat,
                     const float *val, int label) {
  feat_vec x = make_sample(nnz, field, feat, val);
  return static_cast<Ref *>(h)->model->train(x, label);
}

float ftrl_ref_predict(void *h, int nnz, const int32_t *field, const int32_t *feat,
                       const float *val, int output_prob) {
  feat_vec x = make_sample(nnz, field, feat, val);
  return static_cast<Ref *>(h)->model->predict(x, output_prob != 0);
}

double ftrl_ref_train_csr(void *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field,
                          const int32_t *feat, const float *val, const int32_t *label,
                          float *logits_out) {
  auto *r = static_cast<Ref *>(h);
  double tmp_loss = 0.0;
  for (int64_t i = 0; i < n_rows; i++) {
    const int64_t b = row_ptr[i];
    feat_vec x = make_sample((int)(row_ptr[i + 1] - b), field + b, feat + b, val + b);
    const float logit = r->model->train(x, label[i]);
    if (logits_out) logits_out[i] = logit;
    tmp_loss += loss(label[i], logit);
  }
  return tmp_loss;
}

double ftrl_ref_predict_csr(void *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field,
                            const int32_t *feat, const float *val, const int32_t *label,
                            int output_prob, float *out) {
  auto *r = static_cast<Ref *>(h);
  double tmp_loss = 0.0;
  for (int64_t i = 0; i < n_rows; i++) {
    const int64_t b = row_ptr[i];
    feat_vec x = make_sample((int)(row_ptr[i + 1] - b), field + b, feat + b, val + b);
    const float logit = r->model->predict(x, false);
    if (out) out[i] = output_prob ? utils::sigmoid(logit) : logit;
    if (label) tmp_loss += loss(label[i], logit);
  }
  return tmp_loss;
}

// Convert a CSR block into the reference's in-memory form (vector<Sample>),
// outside any timed region -- the reference's Reader does this at load time.
void ftrl_ref_stage_csr(void *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field,
                        const int32_t *feat, const float *val, const int32_t *label) {
  auto *r = static_cast<Ref *>(h);
  r->staged.clear();
  r->staged.resize(n_rows);
  for (int64_t i = 0; i < n_rows; i++) {
    const int64_t b = row_ptr[i];
    r->staged[i].x = make_sample((int)(row_ptr[i + 1] - b), field + b, feat + b, val + b);
    r->staged[i].y = label[i];
  }
}

// One training pass over the staged samples with the reference's worker
// structure (ftrl_offline.cpp:63-103: contiguous chunks of ceil(N/n_threads),
// per-worker fp64 loss partials, joined at the end).  shuffle != 0 shuffles the
// index vector with the given seed first (the reference seeds from
// random_device).  Returns seconds spent inside the epoch (the span the
// reference's own `train time` timer covers); *loss_out = mean loss.
double ftrl_ref_train_staged(void *h, int n_threads, int shuffle, uint32_t seed, double *loss_out) {
  auto *r = static_cast<Ref *>(h);
  auto &samples = r->staged;
  const size_t total_num = samples.size();
  if (n_threads < 1) n_threads = 1;
  const auto t0 = std::chrono::steady_clock::now();
  const size_t unit = (size_t)std::ceil(static_cast<double>(total_num) / n_threads);
  std::vector<int> indices(total_num);
  std::iota(indices.begin(), indices.end(), 0);
  if (shuffle) std::shuffle(indices.begin(), indices.end(), std::mt19937{seed});
  std::vector<double> losses(n_threads, 0.0);
  auto one_thread = [&](size_t idx, size_t start, size_t end) {
    double tmp_loss = 0.0;
    for (auto i = start; i < end; i++) {
      Sample &sample = samples[indices[i]];
      const float logit = r->model->train(sample.x, sample.y);
      tmp_loss += loss(sample.y, logit);
    }
    losses[idx] = tmp_loss;
  };
  std::vector<std::thread> workers;
  for (size_t i = 0; i < (size_t)n_threads; i++) {
    const size_t start = std::min(i * unit, total_num);
    const size_t end = std::min(start + unit, total_num);
    workers.emplace_back([=] { one_thread(i, start, end); });
  }
  for (auto &t : workers) t.join();
  const double total = std::accumulate(losses.begin(), losses.end(), 0.0);
  const double secs =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (loss_out) *loss_out = total_num ? total / static_cast<double>(total_num) : 0.0;
  return secs;
}

// The reference's own line parsers (src/data/parser.cpp:11-41 libsvm, :62-103 libffm).
// Returns the number of kept features, or -1 if the reference throws on this line.
int ftrl_ref_parse_line(int libffm, const char *line, int cap, int32_t *field, int32_t *feat, float *val,
                        int *label) {
  Sample s;
  try {
    if (libffm) {
      ftrl::FFMParser p;
      p.parse(line, s);
    } else {
      ftrl::LibsvmParser p;
      p.parse(line, s);
    }
  } catch (...) {
    return -1;
  }
  int n = 0;
  for (auto &[f, i, v] : s.x) {
    if (n < cap) {
      field[n] = f;
      feat[n] = i;
      val[n] = v;
    }
    n++;
  }
  *label = s.y;
  return n;
}

// scalar helpers of the reference
double ftrl_ref_loss(int y, double logit) { return loss(y, logit); }
float ftrl_ref_sigmoid(float x) { return utils::sigmoid<float>(x); }
float ftrl_ref_sgn(float x) { return utils::sgn<float>(x); }
float ftrl_ref_weight(void *h, float n, float z) {
  return static_cast<Ref *>(h)->model->maybe_zero_weight<float>(n, z);
}

// model files written / read by the reference itself (ffm.cpp:138-200, lr.cpp:26-39)
int ftrl_ref_save_compressed(void *h, const char *path, int level) {
  auto *r = static_cast<Ref *>(h);
  if (r->model_type == 0) {
    as_lr(r)->save_compressed_model(path, level);
    return 0;
  }
  if (r->model_type == 2) {
    as_ffm(r)->save_compressed_model(path, level);
    return 0;
  }
  return -1;  // the reference has no FM save/load
}
int ftrl_ref_load_compressed(void *h, const char *path) {
  auto *r = static_cast<Ref *>(h);
  if (r->model_type == 0) {
    as_lr(r)->load_compressed_model(path);
    return 0;
  }
  if (r->model_type == 2) {
    as_ffm(r)->load_compressed_model(path);
    return 0;
  }
  return -1;
}
int ftrl_ref_save_text(void *h, const char *path) {
  auto *r = static_cast<Ref *>(h);
  if (r->model_type != 2) return -1;
  as_ffm(r)->save_model(path);
  return 0;
}
int ftrl_ref_load_text(void *h, const char *path) {
  auto *r = static_cast<Ref *>(h);
  if (r->model_type != 2) return -1;
  as_ffm(r)->load_model(path);
  return 0;
}

}  // extern "C"
