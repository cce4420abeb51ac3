// prep.cuh -- per-batch index work that turns "contended per-feature mutexes" of the reference
// (ftrl_model.cpp:55,68; ffm.cpp:78,99-101,125) into a sort-by-key segmented reduction:
//   1. k_prep_rows   : validity mask (remove_out_range, ftrl_model.cpp:36-42 / ffm.cpp:30-36),
//                      sort key per occurrence, owning sample per occurrence, per-sample flags,
//                      batch flag "every sample has distinct fields"
//   2. radix sort    : (key = feature id, value = occurrence index)            [cub]
//   3. k_occ_class   : per occurrence: finalised inside its sample (row occurs once in the batch)
//                      or reduced through the sorted list (its sorted position)
//   4. segment scan  : per sorted position {ordinal among segmented rows, row start}   [cub]
//   5. chunk list    : segmented rows cut into chunks of <= CH occurrences             [cub select]
// All integer / index work, HBM-light (about 30 B per occurrence).
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace ftrl {

// one minibatch in CSR form, all pointers in device memory
struct Batch {
  int64_t n_rows;
  int64_t nnz;
  const int64_t *row_ptr;
  const int32_t *field;
  const int32_t *feat;
  const float *val;
  const int32_t *label;
};

enum : uint8_t {
  SF_SIMPLE = 1,   // all valid features of the sample have distinct fields
  SF_FUSABLE = 2,  // rows of this sample that occur once in the batch are finalised per sample
};

struct SegScan {
  int32_t cnt;    // number of segmented-row heads at or before this sorted position
  int32_t start;  // sorted position of the head of this position's row
};
struct SegScanOp {
  __device__ __forceinline__ SegScan operator()(const SegScan &a, const SegScan &b) const {
    return SegScan{a.cnt + b.cnt, a.start > b.start ? a.start : b.start};
  }
};
struct HeadFunctor {
  const uint32_t *skey;
  const uint8_t *fused_sorted;  // 1: row finalised per sample, not part of the segmented list
  __device__ __forceinline__ SegScan operator()(int32_t p) const {
    const bool head = p == 0 || skey[p] != skey[p - 1];
    return SegScan{(head && !fused_sorted[p]) ? 1 : 0, head ? p : 0};
  }
};
struct ChunkHeadPred {
  const uint32_t *skey;
  const SegScan *scan;
  const uint8_t *fused_sorted;
  uint32_t sentinel;
  int32_t ch;
  __device__ __forceinline__ bool operator()(int32_t p) const {
    if (fused_sorted[p]) return false;
    const SegScan s = scan[p];
    if (s.start == p) return true;  // row head (also the head of the sentinel run)
    return skey[p] != sentinel && ((p - s.start) % ch) == 0;
  }
};

__device__ __forceinline__ bool feat_valid(const Dims &d, int32_t fld, int32_t ft) {
  bool ok = ft >= 0 && ft < d.n_feats;
  if (d.model_type == 2) ok = ok && fld >= 0 && fld < d.n_fields;
  return ok;
}

// warp per sample.  batch_flags[0] is cleared to 0 when some sample repeats a field.
// pmask[t] (FFM, n_fields <= 64): bit f set iff some OTHER valid feature of t's sample carries field f,
// i.e. the slices of t's row this sample touches (ffm.cpp:72-88 touches (feat_m, field_n) for n != m) -- also
// exact for samples that repeat a field (`dup`: the fields carried by two or more features of the sample).
__global__ void k_prep_rows(Batch b, Dims d, uint32_t *__restrict__ key, uint32_t *__restrict__ occ_idx,
                            int32_t *__restrict__ occ_row, uint8_t *__restrict__ sflags,
                            int32_t *__restrict__ batch_flags, uint64_t *__restrict__ pmask) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= b.n_rows) return;
  const int64_t r0 = b.row_ptr[warp], r1 = b.row_ptr[warp + 1];
  bool simple = true;
  uint64_t seen = 0;  // field bitmask when n_fields <= 64
  uint64_t dup = 0;   // fields carried by more than one valid feature
  for (int64_t base = r0; base < r1; base += 32) {
    const int64_t t = base + lane;
    int32_t fld = -1;
    bool ok = false;
    if (t < r1) {
      fld = b.field[t];
      const int32_t ft = b.feat[t];
      ok = feat_valid(d, fld, ft);
      key[t] = ok ? (uint32_t)ft : (uint32_t)d.n_feats;
      occ_idx[t] = (uint32_t)t;
      occ_row[t] = (int32_t)warp;
    }
    const unsigned okmask = __ballot_sync(0xffffffffu, ok);
    if (d.model_type == 2) {
      // duplicates inside this group of 32
      const unsigned same = __match_any_sync(0xffffffffu, ok ? fld : -1 - lane);
      if (ok && __popc(same & okmask) > 1) simple = false;
      if (d.n_fields <= 64) {
        const uint64_t bit = ok ? (1ull << fld) : 0ull;
        if (bit & seen) simple = false;
        const uint64_t dbit = (ok && (__popc(same & okmask) > 1 || (bit & seen))) ? bit : 0ull;
        const unsigned dlo = __reduce_or_sync(0xffffffffu, (unsigned)dbit);
        const unsigned dhi = __reduce_or_sync(0xffffffffu, (unsigned)(dbit >> 32));
        dup |= ((uint64_t)dhi << 32) | dlo;
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)bit);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(bit >> 32));
        seen |= ((uint64_t)hi << 32) | lo;
      } else if (ok) {
        for (int64_t u = r0; u < base; u++) {  // rare: wide field spaces
          const int32_t f2 = b.field[u];
          if (f2 == fld && feat_valid(d, f2, b.feat[u])) simple = false;
        }
      }
    }
  }
  simple = __all_sync(0xffffffffu, simple);
  if (lane == 0) {
    sflags[warp] = simple ? (SF_SIMPLE | SF_FUSABLE) : 0;
    if (!simple) batch_flags[0] = 0;
  }
  if (pmask && d.model_type == 2 && d.n_fields <= 64) {
    // `seen` now holds the field set of the whole sample (identical on all lanes)
    for (int64_t t = r0 + lane; t < r1; t += 32) {
      const int32_t fld = b.field[t];
      const bool ok = feat_valid(d, fld, b.feat[t]);
      // the own field counts only when another feature of the sample carries it too
      pmask[t] = ok ? ((seen & ~(1ull << fld)) | (dup & (1ull << fld))) : 0ull;
    }
  }
}

// chunk_pos[n_chunks] = nnz (terminator so chunk c of the LR/FM kernels ends at chunk_pos[c+1])
__global__ void k_terminate(int32_t *chunk_pos, const int32_t *n_chunks, int32_t nnz) {
  chunk_pos[*n_chunks] = nnz;
}

// Classifies every occurrence from the sorted list.  A row that occurs exactly once in the batch,
// in a sample with distinct fields, is "fused": its update is finalised by the per-sample kernel.
//   fused_sorted[p] = 1 for such rows (p = sorted position)
//   occ_pos[t]      = -1 when occurrence t is fused, else its sorted position p
__global__ void k_occ_class(int32_t nnz, uint32_t sentinel, int fuse, const uint32_t *__restrict__ skey,
                            const uint32_t *__restrict__ socc, const int32_t *__restrict__ occ_row,
                            const uint8_t *__restrict__ sflags, uint8_t *__restrict__ fused_sorted,
                            int32_t *__restrict__ occ_pos) {
  const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const uint32_t k = skey[p];
  const uint32_t t = socc[p];
  bool fused = false;
  if (fuse && k != sentinel) {
    const bool head = p == 0 || skey[p - 1] != k;
    const bool last = p + 1 == nnz || skey[p + 1] != k;
    fused = head && last && (sflags[occ_row[t]] & SF_FUSABLE);
  }
  fused_sorted[p] = fused ? 1 : 0;
  occ_pos[t] = fused ? -1 : p;
}

// everything a chunk-level kernel needs to know about chunk c
struct ChunkInfo {
  int32_t p0, p1;    // sorted-position range of this chunk
  uint32_t key;      // feature row
  bool valid;        // false for the sentinel run
  bool row_head;     // first chunk of its row
  bool row_last;     // last chunk of its row
  int32_t j;         // chunk ordinal inside the row
  int32_t slot;      // partial-sum slot (meaningful when the row has > 1 chunk)
};

// EXPLICIT_END: the chunk list skips fused rows, so the end of a chunk is found from the keys
// (warp-cooperative, ch <= 32); otherwise the chunk ends where the next listed chunk starts.
template <bool EXPLICIT_END>
__device__ __forceinline__ ChunkInfo chunk_info(int32_t c, int32_t nnz, uint32_t sentinel, int32_t ch,
                                                const int32_t *__restrict__ chunk_pos,
                                                const uint32_t *__restrict__ skey,
                                                const SegScan *__restrict__ scan) {
  ChunkInfo ci;
  ci.p0 = chunk_pos[c];
  ci.key = skey[ci.p0];
  ci.valid = ci.key != sentinel;
  const SegScan s = scan[ci.p0];
  ci.row_head = s.start == ci.p0;
  if (EXPLICIT_END) {
    const int lane = threadIdx.x & 31;
    const int32_t q = ci.p0 + lane;
    const unsigned same = __ballot_sync(0xffffffffu, lane < ch && q < nnz && skey[q] == ci.key);
    // keys are sorted: the run of equal keys starting at p0 is contiguous
    const int n_occ = __ffs(~same) - 1 < 0 ? 32 : __ffs(~same) - 1;
    ci.p1 = ci.p0 + n_occ;
    ci.row_last = n_occ < ch || ci.p1 >= nnz || skey[ci.p1] != ci.key;
  } else {
    ci.p1 = chunk_pos[c + 1];
    ci.row_last = ci.p1 >= nnz || skey[ci.p1] != ci.key;
  }
  ci.j = (ci.p0 - s.start) / ch;
  // extras (non-head chunks) strictly before this row's head = (c_first + 1) - row_ordinal
  const int32_t e0 = (c - ci.j + 1) - s.cnt;
  ci.slot = 2 * e0 + ci.j;
  return ci;
}

// ---- chunk descriptors ---------------------------------------------------------------------------
// chunk_info() is a chain of three dependent gathers (chunk_pos -> skey / scan -> end of the run); the FFM row
// kernels (k_row_touch, k_row_materialise, k_ffm_staged_rows, k_ffm_combine) would each walk it per work item.
// k_chunk_desc walks it ONCE per chunk, in the index phase (ids only: it overlaps the previous batch), and
// leaves one 16-byte record per chunk that the row kernels read with a single coalesced load.
//   x = p0 | row_last << 31      y = key      z = slot | (n_occ - 1) << 27      w = j
// (p0 < 2^31; slot < 2 * (nnz / ch + 2) < 2^27 for ch = 32; 1 <= n_occ <= 32)
__device__ __forceinline__ int4 chunk_pack(const ChunkInfo &ci) {
  return make_int4(ci.p0 | (ci.row_last ? (int)0x80000000u : 0), (int)ci.key, ci.slot | ((ci.p1 - ci.p0 - 1) << 27), ci.j);
}
__device__ __forceinline__ ChunkInfo chunk_unpack(const int4 &e, uint32_t sentinel) {
  ChunkInfo ci;
  ci.p0 = e.x & 0x7fffffff;
  ci.row_last = e.x < 0;
  ci.key = (uint32_t)e.y;
  ci.valid = ci.key != sentinel;
  ci.slot = e.z & 0x07ffffff;
  ci.p1 = ci.p0 + (int)((uint32_t)e.z >> 27) + 1;
  ci.j = e.w;
  ci.row_head = e.w == 0;
  return ci;
}
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_chunk_desc(int32_t nnz, uint32_t sentinel, int32_t ch, const int32_t *__restrict__ n_chunks_p,
             const int32_t *__restrict__ chunk_pos, const uint32_t *__restrict__ skey, const SegScan *__restrict__ scan,
             int4 *__restrict__ cdesc) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n_chunks = *n_chunks_p;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info<true>(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (lane == 0) cdesc[c] = chunk_pack(ci);
  }
}

constexpr int SRC_SHIFT = 28;  // source = (rank << 28) | occurrence index  (nnz per rank < 2^28)
constexpr uint32_t SRC_MASK = (1u << SRC_SHIFT) - 1;

// where the per-occurrence masks of every rank live (single GPU: p[0] only)
struct PmaskSrc {
  const uint64_t *p[MAX_SHARDS];
  __device__ __forceinline__ unsigned long long operator()(uint32_t src) const {
    return p[src >> SRC_SHIFT][src & SRC_MASK];
  }
};

// Row-level field masks of the segmented rows: rowmask[head chunk] |= OR of pmask over the chunk's
// occurrences (integer OR: order-independent, so the atomic keeps the result deterministic).
// `src` decoding: occurrence src lives on rank src >> 28 at index src & (2^28-1) (single GPU: rank 0).
template <typename PmaskOf>
__device__ __forceinline__ void row_touch_chunk(int c, const ChunkInfo &ci, int ch, const uint32_t *__restrict__ socc,
                                                PmaskOf pmask_of, unsigned long long *__restrict__ rowmask) {
  const int lane = threadIdx.x & 31;
  unsigned long long m = 0ull;
  for (int p = ci.p0 + lane; p < ci.p1; p += 32) m |= pmask_of(socc[p]);
  const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)m);
  const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(m >> 32));
  if (lane == 0) atomicOr(&rowmask[c - ci.j], ((unsigned long long)hi << 32) | lo);
}

// per-thread variant for chunks that are row heads (used to find rows spanning several chunks):
// such a row continues past its first chunk iff position p0 + ch still carries its key.
__device__ __forceinline__ ChunkInfo chunk_head_info(int32_t c, int32_t nnz, uint32_t sentinel, int32_t ch,
                                                     const int32_t *__restrict__ chunk_pos,
                                                     const uint32_t *__restrict__ skey,
                                                     const SegScan *__restrict__ scan) {
  ChunkInfo ci;
  ci.p0 = chunk_pos[c];
  ci.key = skey[ci.p0];
  ci.valid = ci.key != sentinel;
  const SegScan s = scan[ci.p0];
  ci.row_head = s.start == ci.p0;
  ci.p1 = ci.p0 + ch;
  ci.row_last = ci.p1 >= nnz || skey[ci.p1] != ci.key;
  ci.j = (ci.p0 - s.start) / ch;
  ci.slot = 2 * ((c - ci.j + 1) - s.cnt) + ci.j;
  return ci;
}

}  // namespace ftrl
