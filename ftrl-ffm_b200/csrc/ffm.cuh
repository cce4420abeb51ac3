// ffm.cuh -- FFM minibatch kernels (LDG path): forward + fused singleton finalize, segmented row
// update, partial combine, predict.
//
// Reference functions covered (src/model/ffm.cpp): update_vector_w :72-88 (materialise w from
// n,z), compute_ffm_logit :57-70, update_vector_nz :90-136; plus the linear/bias parts of
// src/model/ftrl_model.cpp:44-85.  Minibatch semantics per SURVEY.md 8(a).
//
// Work decomposition of one sample with F valid features and k factors:
//   item = (unordered pair {m<n}, factor chunk c of VEC floats) -- P*C items, P = F(F-1)/2.
//   The item owns BOTH slices the reference touches for that pair: A = (feat_m, field_n) and
//   B = (feat_n, field_m), so w_A.w_B, g_A = g w_B x, g_B = g w_A x need no transpose through
//   shared memory.  Lanes of a warp cover consecutive (n, c): slice A reads are contiguous inside
//   row feat_m (coalesced 128-bit loads), slice B reads are full 32-byte sectors of distinct rows.
#pragma once
#include "common.cuh"
#include "prep.cuh"

namespace ftrl {

constexpr int FFM_CAP = 128;               // features of a sample cached in shared memory
constexpr int PAIR_LUT_N = FFM_CAP * (FFM_CAP - 1) / 2;
constexpr int FFM_MAX_F = 32768;           // pairs are indexed with 32 bits
constexpr int FFM_SPB = 8;                 // samples walked by one CTA of the generic sample kernel

// flat pair index p (n-major: p = n(n-1)/2 + m, m < n) -> (m, n); independent of F
__device__ __forceinline__ void pair_decode(uint32_t p, const uint32_t *__restrict__ lut, int &m, int &n) {
  if (p < PAIR_LUT_N) {
    const uint32_t e = __ldg(lut + p);
    m = (int)(e & 0xffffu);
    n = (int)(e >> 16);
    return;
  }
  int64_t nn = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
  while (nn * (nn - 1) / 2 > (int64_t)p) nn--;
  while ((nn + 1) * nn / 2 <= (int64_t)p) nn++;
  n = (int)nn;
  m = (int)((int64_t)p - nn * (nn - 1) / 2);
}

__global__ void k_build_pair_lut(uint32_t *lut) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= PAIR_LUT_N) return;
  int n = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
  while (n * (n - 1) / 2 > p) n--;
  while ((n + 1) * n / 2 <= p) n++;
  lut[p] = (uint32_t)(p - n * (n - 1) / 2) | ((uint32_t)n << 16);
}

// factor-chunk decode of a flat item index: C = k / VEC chunks per pair (shift when a power of two)
struct ItemDecode {
  uint32_t C;
  int shift;  // log2(C) or -1
  __device__ __forceinline__ void operator()(uint32_t it, uint32_t &p, uint32_t &c) const {
    if (shift >= 0) {
      p = it >> shift;
      c = it & (C - 1);
    } else {
      p = it / C;
      c = it - p * C;
    }
  }
};
__host__ __device__ inline ItemDecode make_item_decode(int k, int vec) {
  ItemDecode d;
  d.C = (uint32_t)(k / vec);
  d.shift = -1;
  for (int s = 0; s < 16; s++)
    if ((1u << s) == d.C) d.shift = s;
  return d;
}

struct SampleCache {
  float *row[FFM_CAP];    // tab + feat * 3 ld   (nullptr when out of range)
  int32_t fk[FFM_CAP];    // field * k
  float x[FFM_CAP];
  uint8_t fused[FFM_CAP];
};

// ---------------------------------------------------------------------------------------------
// K1: one CTA per sample.  pass 1: gather (z,n) slices, materialise w (stored: the stale-by-one w
// the reference keeps, ffm.cpp:72-88), logit, g.  pass 2 (FUSE): rows that occur exactly once in
// the batch, in a sample with distinct fields, are finalised here: z', n' written in place (20 B
// per coordinate, the algorithmic minimum).  All other rows are left to k_ffm_rows.
// ---------------------------------------------------------------------------------------------
// SH (sharded runs, the generic fallback of shard.cuh): w is the materialised one -- the w plane of the shard row for
// the ids this rank owns, the row cache (by sorted head position) for the others; nothing is stored or finalised.
template <int VEC, bool PRECISE, int THREADS, bool SH = false>
__global__ void __launch_bounds__(THREADS)
k_ffm_sample(Batch b, Dims d, Hyper h, ItemDecode dec, float *__restrict__ tab, float4 *__restrict__ lin,
             const float4 *__restrict__ bias, const uint32_t *__restrict__ pair_lut,
             const int32_t *__restrict__ occ_pos, int fuse, const int32_t *__restrict__ batch_flags,
             int skip_if_simple, float *__restrict__ g_out, float *__restrict__ logit_out,
             const __grid_constant__ RowSpace rsp, const SegScan *__restrict__ scan) {
  if (skip_if_simple && batch_flags[0] != 0) return;  // the tile kernels (ffm_tile.cuh) took this batch
  if (SH && batch_flags[1] != 0) return;               // sharded run: the step was called off
  __shared__ SampleCache sc;
  __shared__ float red[33];
  __shared__ float s_g;
  const int tid = threadIdx.x;
  // each CTA walks FFM_SPB consecutive samples (keeps the grid, hence the cost of a skipped launch, small)
  for (int64_t s = (int64_t)blockIdx.x * FFM_SPB; s < min(b.n_rows, ((int64_t)blockIdx.x + 1) * FFM_SPB); s++) {
  __syncthreads();  // shared sample cache / reduction scratch of the previous sample are free
  const int64_t r0 = b.row_ptr[s];
  int F = (int)min((int64_t)FFM_MAX_F, b.row_ptr[s + 1] - r0);
  const bool fusable = !SH && fuse != 0;  // per occurrence: occ_pos < 0 <=> finalised here
  const int64_t ld = d.ld, rs = 3 * ld;
  // base of the row of occurrence t such that base + 2 ld is its w plane (sharded: only that plane is ever read)
  auto row_of = [&](int64_t t, int32_t ft) -> float * {
    if (!SH) return tab + (int64_t)ft * rs;
    if ((ft & rsp.Gm1) == rsp.rank) return rsp.tab + (int64_t)(ft >> rsp.log2G) * rs;
    return const_cast<float *>(rsp.rc_w) + (int64_t)scan[occ_pos[t]].start * ld - 2 * ld;
  };
  for (int t = tid; t < F && t < FFM_CAP; t += THREADS) {
    const int32_t fl = b.field[r0 + t], ft = b.feat[r0 + t];
    const bool ok = feat_valid(d, fl, ft);
    sc.row[t] = ok ? row_of(r0 + t, ft) : nullptr;
    sc.fk[t] = fl * d.k;
    sc.x[t] = b.val[r0 + t];
    sc.fused[t] = fusable && ok && occ_pos[r0 + t] < 0;
  }
  __syncthreads();
  auto get = [&](int m, float *&row, int32_t &fk, float &x, bool &fz) {
    if (m < FFM_CAP) {
      row = sc.row[m]; fk = sc.fk[m]; x = sc.x[m]; fz = sc.fused[m];
    } else {
      const int32_t fl = b.field[r0 + m], ft = b.feat[r0 + m];
      row = feat_valid(d, fl, ft) ? row_of(r0 + m, ft) : nullptr;
      fk = fl * d.k;
      x = b.val[r0 + m];
      fz = fusable && row != nullptr && occ_pos[r0 + m] < 0;
    }
  };
  const uint32_t n_items = (uint32_t)F * (uint32_t)(F - 1) / 2u * dec.C;

  // ---- pass 1 ----
  float acc = 0.f;
  for (uint32_t it = tid; it < n_items; it += THREADS) {
    uint32_t p, c;
    dec(it, p, c);
    int m, n;
    pair_decode(p, pair_lut, m, n);
    float *ra, *rb;
    int32_t fkm, fkn;
    float xm, xn;
    bool zm, zn;
    get(m, ra, fkm, xm, zm);
    get(n, rb, fkn, xn, zn);
    if (ra == nullptr || rb == nullptr) continue;
    float *pa = ra + fkn + c * VEC;
    float *pb = rb + fkm + c * VEC;
    Vec<VEC> wA, wB;
    float dot = 0.f;
    if (SH) {
      wA.load(pa + 2 * ld);
      wB.load(pb + 2 * ld);
#pragma unroll
      for (int e = 0; e < VEC; e++) dot = fmaf(wA.v[e], wB.v[e], dot);
    } else {
      Vec<VEC> zA, nA, zB, nB;
      zA.load(pa); nA.load(pa + ld); zB.load(pb); nB.load(pb + ld);
#pragma unroll
      for (int e = 0; e < VEC; e++) {
        wA.v[e] = weight_from<PRECISE>(zA.v[e], f_sqrt<PRECISE>(nA.v[e]), h);
        wB.v[e] = weight_from<PRECISE>(zB.v[e], f_sqrt<PRECISE>(nB.v[e]), h);
        dot = fmaf(wA.v[e], wB.v[e], dot);
      }
      wA.store(pa + 2 * ld);
      wB.store(pb + 2 * ld);
    }
    acc = fmaf(dot, xm * xn, acc);
  }
  // linear part (ftrl_model.cpp:44-59): thread t handles feature t
  for (int t = tid; t < F; t += THREADS) {
    const int32_t ft = b.feat[r0 + t];
    if (!feat_valid(d, b.field[r0 + t], ft)) continue;
    float w;
    if (SH) {
      w = (ft & rsp.Gm1) == rsp.rank ? rsp.lin[ft >> rsp.log2G].z : rsp.rc_lin[scan[occ_pos[r0 + t]].start];
    } else {
      const float4 e = lin[ft];
      w = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
      lin[ft].z = w;
    }
    acc = fmaf(w, b.val[r0 + t], acc);
  }
  float logit = block_sum(acc, red);
  if (tid == 0) {
    const float4 bz = *bias;
    logit += weight_from<PRECISE>(bz.x, f_sqrt<PRECISE>(bz.y), h);
    const float g = sigmoid_f(logit) - (float)b.label[s];
    s_g = g;
    g_out[s] = g;
    logit_out[s] = logit;
  }
  if (!fusable) continue;
  __syncthreads();
  const float g = s_g;

  // ---- pass 2: finalize singleton rows (z,n re-read: L1/L2 hits) ----
  for (uint32_t it = tid; it < n_items; it += THREADS) {
    uint32_t p, c;
    dec(it, p, c);
    int m, n;
    pair_decode(p, pair_lut, m, n);
    float *ra, *rb;
    int32_t fkm, fkn;
    float xm, xn;
    bool zm, zn;
    get(m, ra, fkm, xm, zm);
    get(n, rb, fkn, xn, zn);
    if (ra == nullptr || rb == nullptr || !(zm || zn)) continue;
    float *pa = ra + fkn + c * VEC;
    float *pb = rb + fkm + c * VEC;
    Vec<VEC> zA, nA, zB, nB;
    zA.load(pa); nA.load(pa + ld); zB.load(pb); nB.load(pb + ld);
    const float gx = g * (xm * xn);
    float sqA[VEC], sqB[VEC], wA[VEC], wB[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e++) {
      sqA[e] = f_sqrt<PRECISE>(nA.v[e]);
      sqB[e] = f_sqrt<PRECISE>(nB.v[e]);
      wA[e] = weight_from<PRECISE>(zA.v[e], sqA[e], h);
      wB[e] = weight_from<PRECISE>(zB.v[e], sqB[e], h);
    }
    if (zm) {
#pragma unroll
      for (int e = 0; e < VEC; e++) {
        const float gv = gx * wB[e];
        ftrl_apply_sq<PRECISE>(zA.v[e], nA.v[e], sqA[e], wA[e], gv, gv * gv, h);
      }
      zA.store(pa); nA.store(pa + ld);
    }
    if (zn) {
#pragma unroll
      for (int e = 0; e < VEC; e++) {
        const float gv = gx * wA[e];
        ftrl_apply_sq<PRECISE>(zB.v[e], nB.v[e], sqB[e], wB[e], gv, gv * gv, h);
      }
      zB.store(pb); nB.store(pb + ld);
    }
  }
  for (int t = tid; t < F; t += THREADS) {
    const int32_t ft = b.feat[r0 + t];
    if (!feat_valid(d, b.field[r0 + t], ft)) continue;
    if (!(fusable && occ_pos[r0 + t] < 0)) continue;
    float4 e = lin[ft];  // e.z = w written in pass 1 by this very thread
    const float gi = g * b.val[r0 + t];
    ftrl_apply<PRECISE>(e.x, e.y, e.z, gi, gi * gi, h);
    lin[ft] = e;
  }
  }  // sample loop
}

// ---------------------------------------------------------------------------------------------
// K2: one warp per chunk (<= CH occurrences of one feature row, from the sorted occurrence list).
// For every occurrence (sample s, position m) the partners n != m of the sample are revisited and
// g_s * w[feat_n][field_m,:] * x_m x_n is accumulated into the row's (sum g, sum g^2) image in
// shared memory; then either the closed-form update is applied (row fits one chunk) or the partial
// image is parked for k_ffm_combine.  Rows already finalised by k_ffm_sample are skipped.
// Latency structure: the 32 lanes first fetch the metadata of 32 occurrences in parallel (one
// dependent chain for the whole group instead of one per occurrence); per occurrence the partner
// gathers of up to GATHER_U warp-steps are issued together before any accumulation.
// ---------------------------------------------------------------------------------------------
constexpr int GATHER_U = 4;

template <int VEC, bool PRECISE, int WARPS, bool SH = false>
__global__ void __launch_bounds__(WARPS * 32)
k_ffm_rows(Batch b, Dims d, Hyper h, ItemDecode dec, float *__restrict__ tab, float4 *__restrict__ lin, int32_t ch,
           const int32_t *__restrict__ n_chunks_p, const int32_t *__restrict__ chunk_pos,
           const uint32_t *__restrict__ skey, const uint32_t *__restrict__ socc,
           const SegScan *__restrict__ scan, const int32_t *__restrict__ occ_row,
           const uint8_t *__restrict__ sflags, const int32_t *__restrict__ batch_flags, int skip_if_simple,
           const float *__restrict__ g_in, float *__restrict__ part, float2 *__restrict__ part_lin,
           const __grid_constant__ RowSpace rsp, const __grid_constant__ Export ex,
           const int32_t *__restrict__ occ_pos) {
  if (skip_if_simple && batch_flags[0] != 0) return;  // the tile kernels (ffm_tile.cuh) took this batch
  if (SH && batch_flags[1] != 0) return;               // sharded run: the step was called off
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  float *acc0 = smem + (int64_t)wib * 2 * ld;  // sum g
  float *acc1 = acc0 + ld;                     // sum g^2
  const int n_chunks = *n_chunks_p;
  const int32_t nnz = (int32_t)b.nnz;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  const int C = (int)dec.C;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info<true>(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid) continue;
    const bool whole_row = ci.row_head && ci.row_last;
    for (int64_t v = lane; v < 2 * ld; v += 32) acc0[v] = 0.f;
    __syncwarp();
    float sg = 0.f, sg2 = 0.f;  // linear coordinate: lane-partial sums
    for (int base = ci.p0; base < ci.p1; base += 32) {
      // lane l owns occurrence base + l
      const int my_p = base + lane;
      int64_t my_r0 = 0;
      int my_F = 0, my_m = 0, my_fmk = 0, my_simple = 1;
      float my_gx = 0.f;
      if (my_p < ci.p1) {
        const int64_t t = socc[my_p];
        const int32_t s = occ_row[t];
        my_r0 = b.row_ptr[s];
        my_F = (int)min((int64_t)FFM_MAX_F, b.row_ptr[s + 1] - my_r0);
        my_m = (int)(t - my_r0);
        my_fmk = b.field[t] * d.k;
        my_simple = (sflags[s] & SF_SIMPLE) ? 1 : 0;
        my_gx = g_in[s] * b.val[t];
        sg += my_gx;
        sg2 = fmaf(my_gx, my_gx, sg2);
      }
      const int n_here = min(32, ci.p1 - base);
      for (int j = 0; j < n_here; j++) {
        const int64_t r0 = __shfl_sync(0xffffffffu, my_r0, j);
        const int F = __shfl_sync(0xffffffffu, my_F, j);
        const int m = __shfl_sync(0xffffffffu, my_m, j);
        const int fmk = __shfl_sync(0xffffffffu, my_fmk, j);
        const int simple = __shfl_sync(0xffffffffu, my_simple, j);
        const float gxm = __shfl_sync(0xffffffffu, my_gx, j);  // g_s * x_m
        const int n_q = F * C;
        for (int q0 = 0; q0 < n_q; q0 += 32 * GATHER_U) {
          bool act[GATHER_U];
          int off[GATHER_U];
          float gx[GATHER_U];
          const float *src[GATHER_U];
#pragma unroll
          for (int u = 0; u < GATHER_U; u++) {
            const int q = q0 + u * 32 + lane;
            const int n = q / C, c4 = q - n * C;
            act[u] = q < n_q && n != m;
            off[u] = 0;
            gx[u] = 0.f;
            src[u] = nullptr;
            if (act[u]) {
              const int32_t fn = b.field[r0 + n], in = b.feat[r0 + n];
              const float xn = b.val[r0 + n];
              act[u] = feat_valid(d, fn, in);
              if (act[u]) {
                const float *wrow;  // w plane of the partner row
                if (!SH) wrow = tab + (int64_t)in * rs + 2 * ld;
                else if ((in & rsp.Gm1) == rsp.rank) wrow = rsp.tab + (int64_t)(in >> rsp.log2G) * rs + 2 * ld;
                else wrow = rsp.rc_w + (int64_t)scan[occ_pos[r0 + n]].start * ld;
                src[u] = wrow + fmk + c4 * VEC;
                off[u] = fn * d.k + c4 * VEC;
                gx[u] = gxm * xn;
              }
            }
          }
          Vec<VEC> w[GATHER_U];
#pragma unroll
          for (int u = 0; u < GATHER_U; u++)
            if (act[u]) w[u].load(src[u]);
          if (simple) {
#pragma unroll
            for (int u = 0; u < GATHER_U; u++)
              if (act[u]) {
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                  const float gv = gx[u] * w[u].v[e];
                  acc0[off[u] + e] += gv;
                  acc1[off[u] + e] = fmaf(gv, gv, acc1[off[u] + e]);
                }
              }
          } else {
            // partners may share a field: apply lane by lane, step by step (deterministic order)
#pragma unroll
            for (int u = 0; u < GATHER_U; u++)
              for (int l = 0; l < 32; l++) {
                if (l == lane && act[u]) {
#pragma unroll
                  for (int e = 0; e < VEC; e++) {
                    const float gv = gx[u] * w[u].v[e];
                    acc0[off[u] + e] += gv;
                    acc1[off[u] + e] = fmaf(gv, gv, acc1[off[u] + e]);
                  }
                }
                __syncwarp();
              }
          }
        }
        __syncwarp();
      }
    }
    sg = warp_sum(sg);
    sg2 = warp_sum(sg2);
    // sharded: a row held by one chunk goes to its owner's inbox unless this rank owns it and is its only contributor
    const int32_t dst = (SH && whole_row) ? ex.dst_at[ci.p0] : -2;
    const int64_t lrow = SH ? (int64_t)(ci.key >> ex.log2G) : (int64_t)ci.key;
    if (whole_row && dst >= 0) {
      const int q = (int)(ci.key & (uint32_t)ex.Gm1);
      float *o = ex.inbox[q] + (int64_t)dst * 2 * ld;
      for (int64_t v = lane; v < 2 * ld; v += 32) o[v] = acc0[v];
      if (lane == 0) ex.inbox_lin[q][dst] = make_float2(sg, sg2);
    } else if (whole_row) {
      float *row = tab + lrow * rs;
      for (int64_t v = lane * VEC; v < d.row_len; v += 32 * VEC) {
        Vec<VEC> a0, a1;
        a0.load(acc0 + v); a1.load(acc1 + v);
        bool any = false;
#pragma unroll
        for (int e = 0; e < VEC; e++) any = any || a1.v[e] != 0.f || a0.v[e] != 0.f;
        if (!any) continue;
        Vec<VEC> z, n, w;
        z.load(row + v); n.load(row + ld + v); w.load(row + 2 * ld + v);
#pragma unroll
        for (int e = 0; e < VEC; e++) ftrl_apply<PRECISE>(z.v[e], n.v[e], w.v[e], a0.v[e], a1.v[e], h);
        z.store(row + v); n.store(row + ld + v);
      }
      if (lane == 0) {
        float4 e = lin[lrow];
        ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
        lin[lrow] = e;
      }
    } else {
      float *pdst = part + (int64_t)ci.slot * 2 * ld;
      for (int64_t v = lane; v < 2 * ld; v += 32) pdst[v] = acc0[v];
      if (lane == 0) part_lin[ci.slot] = make_float2(sg, sg2);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// K3: rows spanning several chunks: sum the parked partials, apply once.
// A CTA scans THREADS consecutive chunks (one per thread) for multi-chunk row heads, then all of its
// warps cooperate on each such row: warp w sums the partials j = w, w+WARPS, ... (lanes over the
// row's float4 vectors, both planes), the per-warp sums are combined in warp order through shared
// memory (fixed order => deterministic), and one thread per vector applies the closed form.
// ---------------------------------------------------------------------------------------------
constexpr int COMB_VPL = 3;  // float4 vectors per lane and plane per block of the row (96 vectors = 384 floats)

template <bool PRECISE, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ffm_combine(Dims d, Hyper h, float *__restrict__ tab, float4 *__restrict__ lin,
              const int32_t *__restrict__ n_chunks_p, const int4 *__restrict__ cdesc,
              const float *__restrict__ part, const float2 *__restrict__ part_lin, const __grid_constant__ Export ex,
              const int32_t *__restrict__ batch_flags) {
  if (ex.on && batch_flags[1] != 0) return;  // sharded run: the step was called off (shard.cuh: k_check_abort)
  constexpr int WARPS = THREADS / 32;
  constexpr int VB = 32 * COMB_VPL;  // vectors per block
  __shared__ int s_list[THREADS];
  __shared__ int s_n;
  __shared__ float4 s_acc[WARPS][2][VB];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int nvec = (int)(ld >> 2);
  const int n_chunks = *n_chunks_p;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  for (int base = blockIdx.x * THREADS; base < n_chunks; base += gridDim.x * THREADS) {
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int c = base + tid;
    if (c < n_chunks) {
      const ChunkInfo ci = chunk_unpack(cdesc[c], sentinel);
      if (ci.valid && ci.row_head && !ci.row_last) s_list[atomicAdd(&s_n, 1)] = c;
    }
    __syncthreads();
    const int n_list = s_n;
    // (the order of s_list is arbitrary, but every row is handled exactly once and rows are independent)
    // Rows of up to WARPS chunks: one warp per row, chunks added in order -- the same order of additions as the
    // block-cooperative path below, which gives every warp one chunk of such a row.
    for (int li = wib; li < n_list; li += WARPS) {
      const int c0 = s_list[li];
      const ChunkInfo ci = chunk_unpack(cdesc[c0], sentinel);
      int J = 1;
      while (J <= WARPS && c0 + J < n_chunks && (uint32_t)cdesc[c0 + J].y == ci.key) J++;
      if (J > WARPS) continue;
      const int32_t dst = ex.on ? ex.dst_at[ci.p0] : -2;
      const int64_t lrow = (int64_t)(ci.key >> ex.log2G);
      float *row = tab + lrow * rs;
      float *o = dst >= 0 ? ex.inbox[ci.key & ex.Gm1] + (int64_t)dst * 2 * ld : nullptr;
      const float4 *p0 = reinterpret_cast<const float4 *>(part + (int64_t)ci.slot * 2 * ld);
      for (int v = lane; v < nvec; v += 32) {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        for (int j = 0; j < J; j++) {
          const float4 t0 = __ldcs(p0 + (int64_t)j * 2 * nvec + v), t1 = __ldcs(p0 + (int64_t)j * 2 * nvec + nvec + v);
          s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
          s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
        }
        const bool any = s0.x != 0.f || s0.y != 0.f || s0.z != 0.f || s0.w != 0.f || s1.x != 0.f || s1.y != 0.f ||
                         s1.z != 0.f || s1.w != 0.f;
        if (o) {
          reinterpret_cast<float4 *>(o)[v] = s0;
          reinterpret_cast<float4 *>(o + ld)[v] = s1;
        } else if (any) {
          float4 z = reinterpret_cast<float4 *>(row)[v], n = reinterpret_cast<float4 *>(row + ld)[v];
          const float4 w = reinterpret_cast<float4 *>(row + 2 * ld)[v];
          ftrl_apply<PRECISE>(z.x, n.x, w.x, s0.x, s1.x, h);
          ftrl_apply<PRECISE>(z.y, n.y, w.y, s0.y, s1.y, h);
          ftrl_apply<PRECISE>(z.z, n.z, w.z, s0.z, s1.z, h);
          ftrl_apply<PRECISE>(z.w, n.w, w.w, s0.w, s1.w, h);
          reinterpret_cast<float4 *>(row)[v] = z;
          reinterpret_cast<float4 *>(row + ld)[v] = n;
        }
      }
      float sg = 0.f, sg2 = 0.f;
      for (int j = lane; j < J; j += 32) {
        const float2 t = part_lin[ci.slot + j];
        sg += t.x; sg2 += t.y;
      }
      sg = warp_sum(sg);
      sg2 = warp_sum(sg2);
      if (lane == 0) {
        if (dst >= 0) {
          ex.inbox_lin[ci.key & ex.Gm1][dst] = make_float2(sg, sg2);
        } else {
          float4 e = lin[lrow];
          ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
          lin[lrow] = e;
        }
      }
    }
    // longer rows: the whole block per row, chunks split over the warps
    for (int li = 0; li < n_list; li++) {
      const int c0 = s_list[li];
      const ChunkInfo ci = chunk_unpack(cdesc[c0], sentinel);
      // number of chunks of this row: keys are sorted, binary search for the first chunk of the next row
      int lo = c0 + 1, hi = n_chunks;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)cdesc[mid].y == ci.key) lo = mid + 1; else hi = mid;
      }
      const int J = lo - c0;
      if (J <= WARPS) continue;
      const int32_t dst = ex.on ? ex.dst_at[ci.p0] : -2;  // sharded: >= 0 -> the sum goes to the owner's inbox
      const int64_t lrow = (int64_t)(ci.key >> ex.log2G);
      float *row = tab + lrow * rs;
      float *o = dst >= 0 ? ex.inbox[ci.key & ex.Gm1] + (int64_t)dst * 2 * ld : nullptr;
      const float4 *p0 = reinterpret_cast<const float4 *>(part + (int64_t)ci.slot * 2 * ld);
      for (int vb = 0; vb < nvec; vb += VB) {
        float4 a0[COMB_VPL], a1[COMB_VPL];
#pragma unroll
        for (int r = 0; r < COMB_VPL; r++) a0[r] = a1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int j = wib; j < J; j += WARPS) {
          const float4 *pj = p0 + (int64_t)j * 2 * nvec;
#pragma unroll
          for (int r = 0; r < COMB_VPL; r++) {
            const int v = vb + r * 32 + lane;
            if (v < nvec) {
              const float4 t0 = __ldcs(pj + v), t1 = __ldcs(pj + nvec + v);
              a0[r].x += t0.x; a0[r].y += t0.y; a0[r].z += t0.z; a0[r].w += t0.w;
              a1[r].x += t1.x; a1[r].y += t1.y; a1[r].z += t1.z; a1[r].w += t1.w;
            }
          }
        }
#pragma unroll
        for (int r = 0; r < COMB_VPL; r++) {
          s_acc[wib][0][r * 32 + lane] = a0[r];
          s_acc[wib][1][r * 32 + lane] = a1[r];
        }
        __syncthreads();
        if (tid < VB && vb + tid < nvec) {
          const int v = vb + tid;
          float4 s0 = s_acc[0][0][tid], s1 = s_acc[0][1][tid];
          for (int w = 1; w < WARPS; w++) {
            const float4 t0 = s_acc[w][0][tid], t1 = s_acc[w][1][tid];
            s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
            s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
          }
          const bool any = s0.x != 0.f || s0.y != 0.f || s0.z != 0.f || s0.w != 0.f || s1.x != 0.f || s1.y != 0.f ||
                           s1.z != 0.f || s1.w != 0.f;
          if (o) {
            reinterpret_cast<float4 *>(o)[v] = s0;
            reinterpret_cast<float4 *>(o + ld)[v] = s1;
          } else if (any) {
            float4 z = reinterpret_cast<float4 *>(row)[v], n = reinterpret_cast<float4 *>(row + ld)[v];
            const float4 w = reinterpret_cast<float4 *>(row + 2 * ld)[v];
            ftrl_apply<PRECISE>(z.x, n.x, w.x, s0.x, s1.x, h);
            ftrl_apply<PRECISE>(z.y, n.y, w.y, s0.y, s1.y, h);
            ftrl_apply<PRECISE>(z.z, n.z, w.z, s0.z, s1.z, h);
            ftrl_apply<PRECISE>(z.w, n.w, w.w, s0.w, s1.w, h);
            reinterpret_cast<float4 *>(row)[v] = z;
            reinterpret_cast<float4 *>(row + ld)[v] = n;
          }
        }
        __syncthreads();
      }
      if (wib == 0) {
        float sg = 0.f, sg2 = 0.f;
        for (int j = lane; j < J; j += 32) {
          const float2 t = part_lin[ci.slot + j];
          sg += t.x; sg2 += t.y;
        }
        sg = warp_sum(sg);
        sg2 = warp_sum(sg2);
        if (lane == 0) {
          if (dst >= 0) {
            ex.inbox_lin[ci.key & ex.Gm1][dst] = make_float2(sg, sg2);
          } else {
            float4 e = lin[lrow];
            ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
            lin[lrow] = e;
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// predict (ffm.cpp:51-70): stored w only, no state change.  One CTA per sample, any summation order.
// ---------------------------------------------------------------------------------------------
template <int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ffm_predict(Batch b, Dims d, ItemDecode dec, const __grid_constant__ Shards sh,
              const float4 *__restrict__ bias, const uint32_t *__restrict__ pair_lut, int output_prob,
              float *__restrict__ out, float *__restrict__ logit_out) {
  __shared__ float red[33];
  const int tid = threadIdx.x;
  const int64_t s = blockIdx.x;
  const int64_t r0 = b.row_ptr[s];
  const int F = (int)min((int64_t)FFM_MAX_F, b.row_ptr[s + 1] - r0);
  const uint32_t n_items = (uint32_t)F * (uint32_t)(F - 1) / 2u * dec.C;
  const int64_t ld = d.ld, rs = 3 * ld;
  float acc = 0.f;
  for (uint32_t it = tid; it < n_items; it += THREADS) {
    uint32_t p, c;
    dec(it, p, c);
    int m, n;
    pair_decode(p, pair_lut, m, n);
    const int32_t fm = b.field[r0 + m], im = b.feat[r0 + m], fn = b.field[r0 + n], in = b.feat[r0 + n];
    if (!feat_valid(d, fm, im) || !feat_valid(d, fn, in)) continue;
    Vec<VEC> wA, wB;
    wA.load(sh.row(im, rs) + 2 * ld + (int64_t)fn * d.k + c * VEC);
    wB.load(sh.row(in, rs) + 2 * ld + (int64_t)fm * d.k + c * VEC);
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; e++) dot = fmaf(wA.v[e], wB.v[e], dot);
    acc = fmaf(dot, b.val[r0 + m] * b.val[r0 + n], acc);
  }
  for (int t = tid; t < F; t += THREADS) {
    const int32_t fl = b.field[r0 + t], ft = b.feat[r0 + t];
    if (feat_valid(d, fl, ft)) acc = fmaf(sh.linp(ft)->z, b.val[r0 + t], acc);
  }
  float logit = block_sum(acc, red);
  if (tid == 0) {
    logit += bias->z;
    out[s] = output_prob ? sigmoid_f(logit) : logit;
    if (logit_out) logit_out[s] = logit;
  }
}

}  // namespace ftrl
