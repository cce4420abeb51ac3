// model_io.h -- model files in the reference's layouts (host side, no CUDA).
//
//  * compressed: ONE zstd frame, content size recorded in the frame header (the reference's
//    loader requires it: src/compression/compress.cpp:32-36), payload = little-endian fp32
//    [bias][lin_w[0..n_feats)][vec_w rows]  (lr.cpp:26-31, ffm.cpp:138-146).
//    Written/read in streaming fashion so a 100M-row table never needs one host buffer.
//  * text (ffm.cpp:161-200): line 1 bias, n_feats lines lin_w (ostream default, 6 significant
//    digits), n_feats lines of row_len values joined by single spaces (shortest round-trip digits,
//    laid out like fmt's "{}").
//
// zstd is consumed as a library only (system libzstd.so.1, no header in this image): the few
// stable public entry points used are declared here.
#pragma once
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "engine.cuh"

extern "C" {
typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
typedef struct ZSTD_inBuffer_s { const void *src; size_t size; size_t pos; } ZSTD_inBuffer;
typedef struct ZSTD_outBuffer_s { void *dst; size_t size; size_t pos; } ZSTD_outBuffer;
ZSTD_CCtx *ZSTD_createCCtx(void);
size_t ZSTD_freeCCtx(ZSTD_CCtx *);
ZSTD_DCtx *ZSTD_createDCtx(void);
size_t ZSTD_freeDCtx(ZSTD_DCtx *);
size_t ZSTD_CCtx_setParameter(ZSTD_CCtx *, int param, int value);
size_t ZSTD_CCtx_setPledgedSrcSize(ZSTD_CCtx *, unsigned long long);
size_t ZSTD_compressStream2(ZSTD_CCtx *, ZSTD_outBuffer *, ZSTD_inBuffer *, int end_op);
size_t ZSTD_decompressStream(ZSTD_DCtx *, ZSTD_outBuffer *, ZSTD_inBuffer *);
unsigned long long ZSTD_getFrameContentSize(const void *, size_t);
unsigned ZSTD_isError(size_t);
const char *ZSTD_getErrorName(size_t);
}

namespace ftrl {

constexpr int kZstdCompressionLevel = 100;  // ZSTD_c_compressionLevel
constexpr int kZstdContentSizeFlag = 200;   // ZSTD_c_contentSizeFlag
constexpr int kZstdEnd = 2, kZstdContinue = 0;
constexpr unsigned long long kZstdSizeUnknown = 0ULL - 1, kZstdSizeError = 0ULL - 2;

class ModelWriter {
 public:
  ModelWriter(const char *path, int level, uint64_t total_bytes) : path_(path), out_(1 << 20) {
    f_ = fopen(path, "wb");
    if (!f_) throw IoFail{fmt("fopen(%s) for writing failed", path)};
    c_ = ZSTD_createCCtx();
    if (!c_) throw IoFail{"ZSTD_createCCtx failed"};
    check(ZSTD_CCtx_setParameter(c_, kZstdCompressionLevel, level));
    check(ZSTD_CCtx_setParameter(c_, kZstdContentSizeFlag, 1));
    check(ZSTD_CCtx_setPledgedSrcSize(c_, total_bytes));
  }
  ~ModelWriter() {
    if (c_) ZSTD_freeCCtx(c_);
    if (f_) fclose(f_);
  }
  void write(const void *p, size_t n) { pump(p, n, kZstdContinue); }
  uint64_t finish() {
    pump(nullptr, 0, kZstdEnd);
    if (fclose(f_) != 0) {
      f_ = nullptr;
      throw IoFail{fmt("fclose(%s) failed", path_.c_str())};
    }
    f_ = nullptr;
    return written_;
  }

 private:
  void check(size_t rc) {
    if (ZSTD_isError(rc)) throw IoFail{fmt("zstd: %s", ZSTD_getErrorName(rc))};
  }
  void pump(const void *p, size_t n, int op) {
    ZSTD_inBuffer in{p, n, 0};
    for (;;) {
      ZSTD_outBuffer out{out_.data(), out_.size(), 0};
      const size_t rem = ZSTD_compressStream2(c_, &out, &in, op);
      check(rem);
      if (out.pos && fwrite(out_.data(), 1, out.pos, f_) != out.pos) throw IoFail{fmt("fwrite(%s) failed", path_.c_str())};
      written_ += out.pos;
      if (op == kZstdEnd ? rem == 0 : in.pos == in.size) break;
    }
  }
  std::string path_;
  FILE *f_ = nullptr;
  ZSTD_CCtx *c_ = nullptr;
  std::vector<char> out_;
  uint64_t written_ = 0;
};

class ModelReader {
 public:
  explicit ModelReader(const char *path) : path_(path), in_(1 << 20) {
    f_ = fopen(path, "rb");
    if (!f_) throw IoFail{fmt("fopen(%s) for reading failed", path)};
    fseek(f_, 0, SEEK_END);
    file_size_ = (uint64_t)ftell(f_);
    fseek(f_, 0, SEEK_SET);
    fill();
    const unsigned long long cs = ZSTD_getFrameContentSize(in_.data(), have_);
    if (cs == kZstdSizeError) throw IoFail{fmt("%s: not compressed by zstd!", path)};
    if (cs == kZstdSizeUnknown) throw IoFail{fmt("%s: original size unknown!", path)};
    content_ = cs;
    d_ = ZSTD_createDCtx();
    if (!d_) throw IoFail{"ZSTD_createDCtx failed"};
  }
  ~ModelReader() {
    if (d_) ZSTD_freeDCtx(d_);
    if (f_) fclose(f_);
  }
  uint64_t content_size() const { return content_; }
  uint64_t file_size() const { return file_size_; }
  void read(void *dst, size_t n) {
    ZSTD_outBuffer out{dst, n, 0};
    while (out.pos < out.size) {
      if (pos_ == have_) {
        fill();
        if (have_ == 0) throw IoFail{fmt("%s: truncated zstd frame", path_.c_str())};
      }
      ZSTD_inBuffer in{in_.data(), have_, pos_};
      const size_t rc = ZSTD_decompressStream(d_, &out, &in);
      if (ZSTD_isError(rc)) throw IoFail{fmt("%s: zstd: %s", path_.c_str(), ZSTD_getErrorName(rc))};
      pos_ = in.pos;
    }
  }

 private:
  void fill() {
    have_ = fread(in_.data(), 1, in_.size(), f_);
    pos_ = 0;
  }
  std::string path_;
  FILE *f_ = nullptr;
  ZSTD_DCtx *d_ = nullptr;
  std::vector<char> in_;
  size_t have_ = 0, pos_ = 0;
  uint64_t content_ = 0, file_size_ = 0;
};

// shortest round-trip digits, laid out like fmt's "{}" for float: fixed notation while the decimal
// exponent is in [-4, 16), otherwise d.ddde+XX; integral values carry no ".0".
inline void append_float_fmt(std::string &s, float v) {
  if (v != v) { s += "nan"; return; }
  if (v == (float)INFINITY || v == -(float)INFINITY) { s += v < 0 ? "-inf" : "inf"; return; }
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
  std::string sci(buf, r.ptr);  // [-]d[.ddd]e[+-]XX
  size_t i = 0;
  if (sci[0] == '-') { s += '-'; i = 1; }
  const size_t epos = sci.find('e');
  std::string digits;
  for (size_t j = i; j < epos; j++)
    if (sci[j] != '.') digits += sci[j];
  const int exp10 = atoi(sci.c_str() + epos + 1);
  if (digits == "0") { s += "0"; return; }
  if (exp10 < -4 || exp10 >= 16) {
    s += digits[0];
    if (digits.size() > 1) { s += '.'; s.append(digits, 1, std::string::npos); }
    char e[16];
    snprintf(e, sizeof(e), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
    s += e;
  } else if (exp10 < 0) {
    s += "0.";
    s.append((size_t)(-exp10 - 1), '0');
    s += digits;
  } else if ((int)digits.size() <= exp10 + 1) {
    s += digits;
    s.append((size_t)(exp10 + 1 - (int)digits.size()), '0');
  } else {
    s.append(digits, 0, (size_t)exp10 + 1);
    s += '.';
    s.append(digits, (size_t)exp10 + 1, std::string::npos);
  }
}

inline void save_text_model(const char *path, float bias, const float *lin, const float *vec, int64_t n_feats,
                            int64_t row_len) {
  FILE *f = fopen(path, "w");
  if (!f) throw IoFail{fmt("fopen(%s) for writing failed", path)};
  std::string line;
  fprintf(f, "%g\n", (double)bias);  // ostream << float: %g, precision 6
  for (int64_t i = 0; i < n_feats; i++) fprintf(f, "%g\n", (double)lin[i]);
  for (int64_t i = 0; i < n_feats; i++) {
    line.clear();
    for (int64_t j = 0; j < row_len; j++) {
      if (j) line += ' ';
      append_float_fmt(line, vec[i * row_len + j]);
    }
    line += '\n';
    fwrite(line.data(), 1, line.size(), f);
  }
  if (fclose(f) != 0) throw IoFail{fmt("fclose(%s) failed", path)};
}

inline void load_text_model(const char *path, float *bias, float *lin, float *vec, int64_t n_feats, int64_t row_len) {
  FILE *f = fopen(path, "r");
  if (!f) throw IoFail{fmt("Failed to open loading file %s", path)};
  std::vector<char> buf(1 << 16);
  std::string line;
  auto getline = [&](std::string &out) -> bool {
    out.clear();
    while (fgets(buf.data(), (int)buf.size(), f)) {
      out += buf.data();
      if (!out.empty() && out.back() == '\n') { out.pop_back(); return true; }
    }
    return !out.empty();
  };
  bool ok = getline(line);
  if (!ok) { fclose(f); throw IoFail{fmt("%s: empty model file", path)}; }
  *bias = strtof(line.c_str(), nullptr);
  for (int64_t i = 0; i < n_feats; i++) {
    if (!getline(line)) { fclose(f); throw IoFail{fmt("%s: truncated (lin_w)", path)}; }
    lin[i] = strtof(line.c_str(), nullptr);
  }
  for (int64_t i = 0; i < n_feats && row_len; i++) {
    if (!getline(line)) { fclose(f); throw IoFail{fmt("%s: truncated (vec_w row %lld)", path, (long long)i)}; }
    const char *p = line.c_str();
    for (int64_t j = 0; j < row_len; j++) {
      char *end = nullptr;
      vec[i * row_len + j] = strtof(p, &end);
      if (end == p) { fclose(f); throw IoFail{fmt("%s: row %lld has fewer than %lld values", path, (long long)i, (long long)row_len)}; }
      p = end;
    }
  }
  fclose(f);
}

}  // namespace ftrl
