// parser_capi.cpp -- C entry points around host/parser.hpp so the parser can be tested without a GPU
// (tests/test_host_parser.py compares it with the reference's Parser through oracle/_ref).
#include <cstring>

#include "parser.hpp"

extern "C" {

// parses `len` bytes of libsvm/libffm text with n_threads workers.  Returns the number of rows; the
// arrays are owned by the library until the next call (single-threaded use).
static host::Csr g_csr;

int64_t host_parse_text(const char *buf, int64_t len, int libffm, int n_threads) {
  g_csr.clear();
  host::parse_buffer(buf, (size_t)len, libffm != 0, n_threads, g_csr);
  return (int64_t)g_csr.rows();
}
// whole file, optionally through the binary CSR image next to it; *from_cache tells which path was taken
int64_t host_load_file(const char *path, int libffm, int n_threads, int use_cache, int *from_cache) {
  g_csr.clear();
  if (from_cache) *from_cache = 0;
  if (use_cache && host::load_csr_cache(path, libffm != 0, g_csr)) {
    if (from_cache) *from_cache = 1;
    return (int64_t)g_csr.rows();
  }
  std::vector<char> buf;
  if (!host::read_file(path, buf)) return -1;
  host::parse_buffer(buf.data(), buf.size(), libffm != 0, n_threads, g_csr);
  if (use_cache) host::save_csr_cache(path, libffm != 0, g_csr);
  return (int64_t)g_csr.rows();
}
// the file streamed in blocks of `block_bytes` (host::TextBlockReader: parallel pread, persistent parser threads),
// the blocks appended in order; *n_blocks tells how many there were
int64_t host_stream_file(const char *path, int libffm, int n_threads, int64_t block_bytes, int *n_blocks) {
  g_csr.clear();
  if (n_blocks) *n_blocks = 0;
  host::TextBlockReader rd(path, libffm != 0, n_threads, (size_t)block_bytes);
  if (!rd.ok()) return -1;
  host::Csr blk;
  while (rd.next(blk)) {
    g_csr.append(blk);
    if (n_blocks) ++*n_blocks;
  }
  return (int64_t)g_csr.rows();
}
int64_t host_parse_nnz() { return (int64_t)g_csr.feat.size(); }
void host_parse_fetch(int64_t *row_ptr, int32_t *field, int32_t *feat, float *val, int32_t *label) {
  memcpy(row_ptr, g_csr.row_ptr.data(), sizeof(int64_t) * g_csr.row_ptr.size());
  memcpy(field, g_csr.field.data(), sizeof(int32_t) * g_csr.field.size());
  memcpy(feat, g_csr.feat.data(), sizeof(int32_t) * g_csr.feat.size());
  memcpy(val, g_csr.val.data(), sizeof(float) * g_csr.val.size());
  memcpy(label, g_csr.label.data(), sizeof(int32_t) * g_csr.label.size());
}
}
