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
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace host {

struct Csr {
  std::vector<int64_t> row_ptr{0};
  std::vector<int32_t> field, feat, label;
  std::vector<float> val;
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
      if (fast_uint(r, e, iv) && iv < (1 << 24) && (r == e || *r == ' ')) {
        v = (float)iv;  // "1", "37": exact
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
inline void parse_buffer(const char *buf, size_t len, bool libffm, int n_threads, Csr &out) {
  if (n_threads <= 1 || len < (1u << 16)) {
    parse_range(buf, buf + len, libffm, out);
    return;
  }
  std::vector<size_t> cut(n_threads + 1, len);
  cut[0] = 0;
  for (int i = 1; i < n_threads; i++) {
    size_t pos = len / n_threads * i;
    const char *nl = (const char *)memchr(buf + pos, '\n', len - pos);
    cut[i] = nl ? (size_t)(nl - buf) + 1 : len;
  }
  std::vector<Csr> parts(n_threads);
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; i++)
    th.emplace_back([&, i] {
      if (cut[i] < cut[i + 1]) parse_range(buf + cut[i], buf + cut[i + 1], libffm, parts[i]);
    });
  for (auto &t : th) t.join();
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
  th.clear();
  for (int i = 0; i < n_threads; i++)
    th.emplace_back([&, i] {
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
  for (auto &t : th) t.join();
}

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

}  // namespace host
