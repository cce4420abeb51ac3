// parser.hpp -- libsvm / libffm text -> CSR (field, feat, val, label), multi-threaded.
// Parsing rules of the reference (src/data/parser.cpp:11-41 libsvm, :62-103 libffm):
//   * tokens separated by ' '; first token label via stoi, label > 0 -> 1 else 0
//   * libsvm token `feat:val` (field forced to 0), libffm token `field:feat:val`
//   * ints via stoi, values via stof (leading blanks / '+' accepted, trailing junk ignored)
//   * tokens whose value == 0 are dropped; ids are used verbatim (no hashing, no offset)
//   * malformed token: prints `wrong input: <line>`; the reference then throws std::out_of_range out of
//     the worker (process aborts) -- here the process exits with EXIT_FAILURE
// Whole-file loading splits the byte range across n_threads at line boundaries like
// Reader::get_data_partition (src/data/reader.cpp:22-48) and keeps file order.
#pragma once
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

namespace host {

// std::vector whose resize() leaves new elements uninitialised: the merge / cache loader overwrite every element
// right away, and value-initialising 27 MB per 32 MB block of text was a fifth of the streaming front end's time
template <class T>
struct NoInit : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = NoInit<U>;
  };
  NoInit() = default;
  template <class U>
  NoInit(const NoInit<U> &) {}
  template <class U, class... A>
  void construct(U *p, A &&...a) {
    if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
    else ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
template <class T>
using uvec = std::vector<T, NoInit<T>>;

struct Csr {
  uvec<int64_t> row_ptr{0};
  uvec<int32_t> field, feat, label;
  uvec<float> val;
  size_t rows() const { return label.size(); }
  void clear() {
    row_ptr.assign(1, 0);
    field.clear(); feat.clear(); label.clear(); val.clear();
  }
  void append(const Csr &o) {
    const int64_t base = row_ptr.back();
    for (size_t i = 1; i < o.row_ptr.size(); i++) row_ptr.push_back(base + o.row_ptr[i]);
    field.insert(field.end(), o.field.begin(), o.field.end());
    feat.insert(feat.end(), o.feat.begin(), o.feat.end());
    val.insert(val.end(), o.val.begin(), o.val.end());
    label.insert(label.end(), o.label.begin(), o.label.end());
  }
};

// a few persistent worker threads: run(n, fn) calls fn(0) .. fn(n-1), one index per worker at a time, and returns
// when all are done.  The streaming front end dispatches three short parallel phases per 32 MB block of text
// (read, parse, merge): creating 48 threads per block cost more than the phases themselves.
class WorkerPool {
 public:
  explicit WorkerPool(int n_workers) {
    for (int i = 0; i < n_workers; i++) th_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  template <class F>
  void run(int n, F &&fn) {
    if (n <= 0) return;
    std::function<void(int)> f = std::forward<F>(fn);
    {
      std::lock_guard<std::mutex> g(mu_);
      fn_ = &f;
      next_ = 0;
      n_ = n;
      left_ = n;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> g(mu_);
    done_.wait(g, [this] { return left_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop() {
    std::unique_lock<std::mutex> g(mu_);
    for (;;) {
      cv_.wait(g, [this] { return stop_ || (fn_ && next_ < n_); });
      if (stop_) return;
      const int i = next_++;
      std::function<void(int)> *f = fn_;
      g.unlock();
      (*f)(i);
      g.lock();
      if (--left_ == 0) done_.notify_all();
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::function<void(int)> *fn_ = nullptr;
  int next_ = 0, n_ = 0, left_ = 0;
  bool stop_ = false;
};

[[noreturn]] inline void wrong_input(const char *b, const char *e) {
  printf("wrong input: %.*s\n", (int)(e - b), b);
  fflush(stdout);
  exit(EXIT_FAILURE);
}

// std::stoi semantics on [p, end): skips blanks, optional sign, digits; returns false if no digits
inline bool parse_int(const char *&p, const char *end, long &out) {
  while (p < end && (*p == ' ' || *p == '\t')) p++;
  const char *q = p;
  if (q < end && *q == '+') q++;
  auto r = std::from_chars(q, end, out);
  if (r.ec != std::errc()) return false;
  p = r.ptr;
  return true;
}
// std::stof semantics
inline bool parse_float(const char *&p, const char *end, float &out) {
  while (p < end && (*p == ' ' || *p == '\t')) p++;
  const char *q = p;
  if (q < end && *q == '+') q++;
  auto r = std::from_chars(q, end, out);
  if (r.ec == std::errc::result_out_of_range) {  // stof would throw; keep the saturated value
    char *e2 = nullptr;
    std::string tmp(q, end);
    out = strtof(tmp.c_str(), &e2);
    p = q + (e2 - tmp.c_str());
    return true;
  }
  if (r.ec != std::errc()) return false;
  p = r.ptr;
  return true;
}

// one token [p, tok_end), the general path: delimiters are searched first, each number may carry trailing
// junk (stoi / stof semantics)
inline void parse_token_general(const char *b, const char *e, const char *p, const char *tok_end, bool libffm, Csr &out) {
  long fld = 0, ft = 0;
  float v = 0.f;
  const char *q = p;
  if (libffm) {
    const char *c1 = (const char *)memchr(q, ':', (size_t)(tok_end - q));
    if (!c1) wrong_input(b, e);
    if (!parse_int(q, c1, fld)) wrong_input(b, e);
    q = c1 + 1;
    if (q >= tok_end) wrong_input(b, e);
  }
  const char *c2 = (const char *)memchr(q, ':', (size_t)(tok_end - q));
  if (!c2) wrong_input(b, e);
  if (!parse_int(q, c2, ft)) wrong_input(b, e);
  q = c2 + 1;
  if (q >= tok_end) wrong_input(b, e);
  if (!parse_float(q, tok_end, v)) wrong_input(b, e);
  // ids beyond the int range: std::stoi throws std::out_of_range in the reference (parser.cpp:26,33,76,83),
  // which ends the run -- never let such an id wrap into [0, n_feats)
  if (fld > 0x7fffffffL || fld < -0x80000000L || ft > 0x7fffffffL || ft < -0x80000000L) wrong_input(b, e);
  if (v != 0.0f) {
    out.field.push_back((int32_t)fld);
    out.feat.push_back((int32_t)ft);
    out.val.push_back(v);
  }
}

// unsigned decimal at p (no sign, no blanks): the common case of every id in a data file
inline bool fast_uint(const char *&p, const char *e, long &out) {
  const char *q = p;
  unsigned long v = 0;
  while (q < e && (unsigned)(*q - '0') <= 9u && q - p < 18) v = v * 10 + (unsigned)(*q++ - '0');
  if (q == p) return false;
  out = (long)v;
  p = q;
  return true;
}

// one line [b, e) without the newline.  Fast path per token: `digits:digits:number` (libffm) or
// `digits:number` (libsvm) with the delimiters exactly where the digits end; anything else (signs, blanks,
// trailing junk, ids beyond int range) goes through parse_token_general, which applies the stoi / stof rules
inline void parse_line(const char *b, const char *e, bool libffm, Csr &out) {
  const char *p = b;
  while (p < e && *p == ' ') p++;
  if (p == e) return;  // blank line
  long lab = 0;
  if (!parse_int(p, e, lab)) wrong_input(b, e);
  while (p < e && *p != ' ') p++;  // rest of the first token is ignored like stoi(substr)
  while (true) {
    while (p < e && *p == ' ') p++;
    if (p >= e) break;
    const char *q = p;
    long fld = 0, ft = 0;
    bool ok = true;
    if (libffm) ok = fast_uint(q, e, fld) && q < e && *q == ':' && ++q < e;
    ok = ok && fast_uint(q, e, ft) && q < e && *q == ':' && ++q < e && ft <= 0x7fffffffL && fld <= 0x7fffffffL;
    if (ok) {
      float v;
      long iv = 0;
      const char *r = q;
      bool have = false;
      if (fast_uint(r, e, iv) && iv < (1 << 24)) {
        if (r == e || *r == ' ') {
          v = (float)iv;  // "1", "37": exact
          have = true;
        } else if (*r == '.') {
          // "0.5489": mantissa < 2^24 and 10^digits <= 10^10 are exact in fp32, so ONE correctly rounded
          // division gives the correctly rounded value (Clinger's fast path) -- the same float stof returns
          static const float kPow10[11] = {1.f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
          const char *f = r + 1;
          unsigned long mant = (unsigned long)iv;
          int nd = 0;
          while (f < e && (unsigned)(*f - '0') <= 9u && nd < 10 && mant < (1ul << 24)) {
            mant = mant * 10 + (unsigned)(*f++ - '0');
            nd++;
          }
          if (mant < (1ul << 24) && (f == e || *f == ' ')) {
            v = (float)mant / kPow10[nd];
            r = f;
            have = true;
          }
        }
      }
      if (have) {
        q = r;
      } else {
        auto fr = std::from_chars(q, e, v);
        ok = fr.ec == std::errc() && (fr.ptr == e || *fr.ptr == ' ');
        q = fr.ptr;
      }
      if (ok) {
        if (v != 0.0f) {
          out.field.push_back((int32_t)fld);
          out.feat.push_back((int32_t)ft);
          out.val.push_back(v);
        }
        p = q;
        continue;
      }
    }
    const char *tok_end = (const char *)memchr(p, ' ', (size_t)(e - p));
    if (!tok_end) tok_end = e;
    parse_token_general(b, e, p, tok_end, libffm, out);
    p = tok_end;
  }
  out.label.push_back(lab > 0 ? 1 : 0);
  out.row_ptr.push_back((int64_t)out.feat.size());
}

inline void parse_range(const char *b, const char *e, bool libffm, Csr &out) {
  {
    // about 12 bytes of text per token and 40 tokens per line on Criteo-shaped data: one allocation, not 20
    const size_t bytes = (size_t)(e - b), tok = out.feat.size() + bytes / 10, rows = out.label.size() + bytes / 200;
    out.field.reserve(tok);
    out.feat.reserve(tok);
    out.val.reserve(tok);
    out.label.reserve(rows);
    out.row_ptr.reserve(rows + 1);
  }
  while (b < e) {
    const char *nl = (const char *)memchr(b, '\n', (size_t)(e - b));
    const char *le = nl ? nl : e;
    const char *lt = le;
    if (lt > b && lt[-1] == '\r') lt--;
    parse_line(b, lt, libffm, out);
    b = nl ? nl + 1 : e;
  }
}

// parse a text buffer with n_threads workers, file order preserved
// `scratch` (optional): the per-thread parts of the previous call; their capacity (and the pages behind it) is reused.
// `pool` (optional): persistent workers instead of n_threads new threads per phase.
inline void parse_buffer(const char *buf, size_t len, bool libffm, int n_threads, Csr &out,
                         std::vector<Csr> *scratch = nullptr, WorkerPool *pool = nullptr) {
  if (n_threads <= 1 || len < (1u << 16)) {
    parse_range(buf, buf + len, libffm, out);
    return;
  }
  auto parallel = [&](auto &&fn) {
    if (pool) {
      pool->run(n_threads, fn);
    } else {
      std::vector<std::thread> th;
      for (int i = 0; i < n_threads; i++) th.emplace_back([&fn, i] { fn(i); });
      for (auto &t : th) t.join();
    }
  };
  std::vector<size_t> cut(n_threads + 1, len);
  cut[0] = 0;
  for (int i = 1; i < n_threads; i++) {
    size_t pos = len / n_threads * i;
    const char *nl = (const char *)memchr(buf + pos, '\n', len - pos);
    cut[i] = nl ? (size_t)(nl - buf) + 1 : len;
  }
  std::vector<Csr> own;
  std::vector<Csr> &parts = scratch ? *scratch : own;
  parts.resize(n_threads);
  for (auto &pt : parts) pt.clear();
  parallel([&](int i) {
    if (cut[i] < cut[i + 1]) parse_range(buf + cut[i], buf + cut[i + 1], libffm, parts[i]);
  });
  // merge in file order: sizes first, then every worker copies its part to its offset
  std::vector<size_t> row0(n_threads + 1, out.rows()), nz0(n_threads + 1, out.feat.size());
  for (int i = 0; i < n_threads; i++) {
    row0[i + 1] = row0[i] + parts[i].rows();
    nz0[i + 1] = nz0[i] + parts[i].feat.size();
  }
  out.row_ptr.resize(row0[n_threads] + 1);
  out.label.resize(row0[n_threads]);
  out.field.resize(nz0[n_threads]);
  out.feat.resize(nz0[n_threads]);
  out.val.resize(nz0[n_threads]);
  parallel([&](int i) {
    const Csr &p = parts[i];
    const size_t nr = p.rows(), nz = p.feat.size();
    for (size_t r = 0; r < nr; r++) out.row_ptr[row0[i] + r + 1] = (int64_t)nz0[i] + p.row_ptr[r + 1];
    if (nr) memcpy(out.label.data() + row0[i], p.label.data(), nr * sizeof(int32_t));
    if (nz) {
      memcpy(out.field.data() + nz0[i], p.field.data(), nz * sizeof(int32_t));
      memcpy(out.feat.data() + nz0[i], p.feat.data(), nz * sizeof(int32_t));
      memcpy(out.val.data() + nz0[i], p.val.data(), nz * sizeof(float));
    }
  });
}

// Streams a text file as blocks of complete lines, each parsed into a CSR block (file order kept): the producer half
// of PcTask::run (src/concurrent/pc_task.cpp:22-80).  A regular file is read with one pread per parser thread, in
// parallel, into a buffer that is reused from block to block (no page faults after the first block, no serial 32 MB
// read); anything else (a pipe, a FIFO) goes through fread.  The parser threads are persistent (WorkerPool).
class TextBlockReader {
 public:
  TextBlockReader(const std::string &path, bool libffm, int n_threads, size_t block_bytes = 32u << 20)
      : libffm_(libffm), n_threads_(n_threads < 1 ? 1 : n_threads), pool_(n_threads < 1 ? 1 : n_threads),
        buf_(block_bytes + (block_bytes >> 5) + 4096) {
    f_ = fopen(path.c_str(), "rb");
    if (!f_) return;
    struct stat st;
    if (fstat(fileno(f_), &st) == 0 && S_ISREG(st.st_mode)) file_len_ = st.st_size;
  }
  ~TextBlockReader() {
    if (f_) fclose(f_);
  }
  TextBlockReader(const TextBlockReader &) = delete;
  TextBlockReader &operator=(const TextBlockReader &) = delete;
  bool ok() const { return f_ != nullptr; }

  // next block of complete lines -> `out` (cleared first); false when the file is exhausted
  bool next(Csr &out) {
    out.clear();
    while (!eof_ || have_) {
      if (!eof_) {
        size_t got = 0;
        const size_t room = buf_.size() - have_;
        if (file_len_ >= 0) {
          const size_t want = (size_t)std::min<off_t>((off_t)room, file_len_ - file_pos_);
          const int nt = pool_.size();
          const size_t per = (want + nt - 1) / nt;
          std::vector<size_t> done(nt, 0);
          const int fd = fileno(f_);
          pool_.run(nt, [&](int i) {
            const size_t o0 = std::min(want, per * (size_t)i), o1 = std::min(want, o0 + per);
            size_t d = 0;
            while (o0 + d < o1) {
              const ssize_t r = pread(fd, buf_.data() + have_ + o0 + d, o1 - o0 - d, file_pos_ + (off_t)(o0 + d));
              if (r <= 0) break;
              d += (size_t)r;
            }
            done[i] = d;
          });
          for (int i = 0; i < nt; i++) {  // a short read (file truncated meanwhile) ends the stream there
            const size_t o0 = std::min(want, per * (size_t)i), o1 = std::min(want, o0 + per);
            got += done[i];
            if (done[i] < o1 - o0) break;
          }
          file_pos_ += (off_t)got;
          if (got == 0 || file_pos_ >= file_len_) eof_ = true;
        } else {
          got = fread(buf_.data() + have_, 1, room, f_);
          if (got == 0) eof_ = true;
        }
        have_ += got;
      }
      size_t use = have_;
      if (!eof_) {  // cut at the last complete line
        while (use > 0 && buf_[use - 1] != '\n') use--;
        if (use == 0) {  // one line longer than the buffer
          if (have_ == buf_.size()) buf_.resize(buf_.size() * 2);
          continue;
        }
      }
      if (use == 0) return false;
      parse_buffer(buf_.data(), use, libffm_, n_threads_, out, &scratch_, &pool_);
      memmove(buf_.data(), buf_.data() + use, have_ - use);
      have_ -= use;
      return true;
    }
    return false;
  }

 private:
  bool libffm_;
  int n_threads_;
  WorkerPool pool_;
  std::vector<char> buf_;
  std::vector<Csr> scratch_;  // per-thread parts, reused from block to block
  FILE *f_ = nullptr;
  off_t file_len_ = -1, file_pos_ = 0;
  size_t have_ = 0;
  bool eof_ = false;
};

inline bool read_file(const std::string &path, std::vector<char> &buf) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize((size_t)n);
  const size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n;
}

// ---- binary CSR image of a parsed text file (at GPU speed the text front-end is the
// bottleneck; parse once, reuse while the text is unchanged).  `<text>.csr` = header + raw arrays; the header
// carries the size and mtime of the text file it was made from and the format it was parsed as.
struct CsrCacheHeader {
  char magic[8];  // "FTRLCSR1"
  int64_t src_size, src_mtime_ns;
  int64_t rows, nnz;
  int32_t libffm, pad;
};

inline bool file_fingerprint(const std::string &path, int64_t &size, int64_t &mtime_ns) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return false;
  size = (int64_t)st.st_size;
  mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
  return true;
}

inline bool save_csr_cache(const std::string &text_path, bool libffm, const Csr &c) {
  CsrCacheHeader h{};
  memcpy(h.magic, "FTRLCSR1", 8);
  if (!file_fingerprint(text_path, h.src_size, h.src_mtime_ns)) return false;
  h.rows = (int64_t)c.rows();
  h.nnz = (int64_t)c.feat.size();
  h.libffm = libffm ? 1 : 0;
  const std::string tmp = text_path + ".csr.tmp", dst = text_path + ".csr";
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return false;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  ok = ok && fwrite(c.row_ptr.data(), sizeof(int64_t), c.row_ptr.size(), f) == c.row_ptr.size();
  ok = ok && fwrite(c.label.data(), sizeof(int32_t), c.label.size(), f) == c.label.size();
  ok = ok && fwrite(c.field.data(), sizeof(int32_t), c.field.size(), f) == c.field.size();
  ok = ok && fwrite(c.feat.data(), sizeof(int32_t), c.feat.size(), f) == c.feat.size();
  ok = ok && fwrite(c.val.data(), sizeof(float), c.val.size(), f) == c.val.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), dst.c_str()) != 0) {
    remove(tmp.c_str());
    return false;
  }
  return true;
}

// false when there is no image, it belongs to another version of the text, or it is truncated
inline bool load_csr_cache(const std::string &text_path, bool libffm, Csr &c) {
  int64_t size = 0, mtime = 0;
  if (!file_fingerprint(text_path, size, mtime)) return false;
  FILE *f = fopen((text_path + ".csr").c_str(), "rb");
  if (!f) return false;
  CsrCacheHeader h{};
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "FTRLCSR1", 8) == 0 && h.src_size == size &&
            h.src_mtime_ns == mtime && h.libffm == (libffm ? 1 : 0) && h.rows >= 0 && h.nnz >= 0;
  if (ok) {
    // the header must agree with the length of the image before anything is sized from it
    struct stat st;
    const __int128 want = (__int128)sizeof(h) + ((__int128)h.rows + 1) * 8 + (__int128)h.rows * 4 + (__int128)h.nnz * 12;
    ok = fstat(fileno(f), &st) == 0 && (__int128)st.st_size == want;
  }
  if (ok) {
    c.row_ptr.resize((size_t)h.rows + 1);
    c.label.resize((size_t)h.rows);
    c.field.resize((size_t)h.nnz);
    c.feat.resize((size_t)h.nnz);
    c.val.resize((size_t)h.nnz);
    ok = fread(c.row_ptr.data(), sizeof(int64_t), c.row_ptr.size(), f) == c.row_ptr.size();
    ok = ok && fread(c.label.data(), sizeof(int32_t), c.label.size(), f) == c.label.size();
    ok = ok && fread(c.field.data(), sizeof(int32_t), c.field.size(), f) == c.field.size();
    ok = ok && fread(c.feat.data(), sizeof(int32_t), c.feat.size(), f) == c.feat.size();
    ok = ok && fread(c.val.data(), sizeof(float), c.val.size(), f) == c.val.size();
    ok = ok && c.row_ptr.front() == 0 && c.row_ptr.back() == h.nnz;
    for (size_t r = 0; ok && r + 1 < c.row_ptr.size(); r++) ok = c.row_ptr[r] <= c.row_ptr[r + 1];  // monotone
  }
  fclose(f);
  if (!ok) c.clear();
  return ok;
}

}  // namespace host
