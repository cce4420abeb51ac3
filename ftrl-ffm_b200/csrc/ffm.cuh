// ffm.cuh -- FFM minibatch kernels (generic LDG path): forward + fused singleton finalize,
// segmented row update, partial combine, predict.
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

// flat pair index p (n-major: p = n(n-1)/2 + m, m < n) -> (m, n); independent of F
__device__ __forceinline__ void pair_decode(int64_t p, const uint32_t *__restrict__ lut, int &m, int &n) {
  if (p < PAIR_LUT_N) {
    const uint32_t e = __ldg(lut + p);
    m = (int)(e & 0xffffu);
    n = (int)(e >> 16);
    return;
  }
  int64_t nn = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
  while (nn * (nn - 1) / 2 > p) nn--;
  while ((nn + 1) * nn / 2 <= p) nn++;
  n = (int)nn;
  m = (int)(p - nn * (nn - 1) / 2);
}

__global__ void k_build_pair_lut(uint32_t *lut) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= PAIR_LUT_N) return;
  int n = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
  while (n * (n - 1) / 2 > p) n--;
  while ((n + 1) * n / 2 <= p) n++;
  lut[p] = (uint32_t)(p - n * (n - 1) / 2) | ((uint32_t)n << 16);
}

struct SampleCache {
  int32_t fld[FFM_CAP];
  int32_t ft[FFM_CAP];  // -1 when out of range
  float x[FFM_CAP];
  uint8_t fused[FFM_CAP];
};

// ---------------------------------------------------------------------------------------------
// K1: one CTA per sample.  pass 1: gather (z,n) slices, materialise w (stored: the stale-by-one w
// the reference keeps, ffm.cpp:72-88), logit, g, loss.  pass 2 (FUSE): rows that occur exactly once
// in the batch, in a sample with distinct fields, are finalised here: z', n' written in place
// (20 B per coordinate, the algorithmic minimum).  All other rows are left to k_ffm_rows.
// ---------------------------------------------------------------------------------------------
template <int VEC, bool PRECISE, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ffm_sample(Batch b, Dims d, Hyper h, float *__restrict__ tab, float4 *__restrict__ lin,
             const float4 *__restrict__ bias, const uint32_t *__restrict__ pair_lut,
             const uint8_t *__restrict__ occ_single, const uint8_t *__restrict__ sflags, int fuse,
             float *__restrict__ g_out, float *__restrict__ logit_out, double *__restrict__ loss_out) {
  __shared__ SampleCache sc;
  __shared__ float red[33];
  __shared__ float s_g;
  const int tid = threadIdx.x;
  const int64_t s = blockIdx.x;
  const int64_t r0 = b.row_ptr[s];
  const int F = (int)(b.row_ptr[s + 1] - r0);
  const bool fusable = fuse && (sflags[s] & SF_FUSABLE);
  for (int t = tid; t < F && t < FFM_CAP; t += THREADS) {
    const int32_t fl = b.field[r0 + t], ft = b.feat[r0 + t];
    sc.fld[t] = fl;
    sc.ft[t] = feat_valid(d, fl, ft) ? ft : -1;
    sc.x[t] = b.val[r0 + t];
    sc.fused[t] = fusable && occ_single[r0 + t];
  }
  __syncthreads();
  auto get = [&](int m, int32_t &fl, int32_t &ft, float &x, bool &fz) {
    if (m < FFM_CAP) {
      fl = sc.fld[m]; ft = sc.ft[m]; x = sc.x[m]; fz = sc.fused[m];
    } else {
      fl = b.field[r0 + m]; ft = b.feat[r0 + m]; x = b.val[r0 + m];
      if (!feat_valid(d, fl, ft)) ft = -1;
      fz = fusable && occ_single[r0 + m];
    }
  };
  const int C = (d.k + VEC - 1) / VEC;  // VEC divides k by construction of the dispatch
  const int64_t n_items = (int64_t)F * (F - 1) / 2 * C;
  const int64_t ld = d.ld, rs = 3 * ld;

  // ---- pass 1 ----
  float acc = 0.f;
  for (int64_t it = tid; it < n_items; it += THREADS) {
    const int c = (int)(it % C);
    int m, n;
    pair_decode(it / C, pair_lut, m, n);
    int32_t fm, im, fn, in;
    float xm, xn;
    bool zm, zn;
    get(m, fm, im, xm, zm);
    get(n, fn, in, xn, zn);
    if (im < 0 || in < 0) continue;
    float *pa = tab + (int64_t)im * rs + (int64_t)fn * d.k + c * VEC;
    float *pb = tab + (int64_t)in * rs + (int64_t)fm * d.k + c * VEC;
    Vec<VEC> zA, nA, zB, nB, wA, wB;
    zA.load(pa); nA.load(pa + ld); zB.load(pb); nB.load(pb + ld);
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; e++) {
      wA.v[e] = weight_from<PRECISE>(zA.v[e], f_sqrt<PRECISE>(nA.v[e]), h);
      wB.v[e] = weight_from<PRECISE>(zB.v[e], f_sqrt<PRECISE>(nB.v[e]), h);
      dot = fmaf(wA.v[e], wB.v[e], dot);
    }
    wA.store(pa + 2 * ld);
    wB.store(pb + 2 * ld);
    acc = fmaf(dot, xm * xn, acc);
  }
  // linear part (ftrl_model.cpp:44-59): thread t handles feature t
  for (int t = tid; t < F; t += THREADS) {
    int32_t fl, ft; float x; bool fz;
    get(t, fl, ft, x, fz);
    if (ft < 0) continue;
    const float4 e = lin[ft];
    const float w = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
    lin[ft].z = w;
    acc = fmaf(w, x, acc);
  }
  float logit = block_sum(acc, red);
  if (tid == 0) {
    const float4 bz = *bias;
    logit += weight_from<PRECISE>(bz.x, f_sqrt<PRECISE>(bz.y), h);
    const int y = b.label[s];
    const float g = sigmoid_f(logit) - (float)y;
    s_g = g;
    g_out[s] = g;
    if (logit_out) logit_out[s] = logit;
    loss_out[s] = logloss_d(y, logit);
  }
  if (!fusable) return;
  __syncthreads();
  const float g = s_g;

  // ---- pass 2: finalize singleton rows (z,n re-read: L1/L2 hits) ----
  for (int64_t it = tid; it < n_items; it += THREADS) {
    const int c = (int)(it % C);
    int m, n;
    pair_decode(it / C, pair_lut, m, n);
    int32_t fm, im, fn, in;
    float xm, xn;
    bool zm, zn;
    get(m, fm, im, xm, zm);
    get(n, fn, in, xn, zn);
    if (im < 0 || in < 0 || !(zm || zn)) continue;
    float *pa = tab + (int64_t)im * rs + (int64_t)fn * d.k + c * VEC;
    float *pb = tab + (int64_t)in * rs + (int64_t)fm * d.k + c * VEC;
    Vec<VEC> zA, nA, zB, nB, wA, wB;
    zA.load(pa); nA.load(pa + ld); zB.load(pb); nB.load(pb + ld);
    const float gx = g * (xm * xn);
#pragma unroll
    for (int e = 0; e < VEC; e++) {
      wA.v[e] = weight_from<PRECISE>(zA.v[e], f_sqrt<PRECISE>(nA.v[e]), h);
      wB.v[e] = weight_from<PRECISE>(zB.v[e], f_sqrt<PRECISE>(nB.v[e]), h);
    }
    if (zm) {
#pragma unroll
      for (int e = 0; e < VEC; e++) {
        const float gv = gx * wB.v[e];
        ftrl_apply<PRECISE>(zA.v[e], nA.v[e], wA.v[e], gv, gv * gv, h);
      }
      zA.store(pa); nA.store(pa + ld);
    }
    if (zn) {
#pragma unroll
      for (int e = 0; e < VEC; e++) {
        const float gv = gx * wA.v[e];
        ftrl_apply<PRECISE>(zB.v[e], nB.v[e], wB.v[e], gv, gv * gv, h);
      }
      zB.store(pb); nB.store(pb + ld);
    }
  }
  for (int t = tid; t < F; t += THREADS) {
    int32_t fl, ft; float x; bool fz;
    get(t, fl, ft, x, fz);
    if (ft < 0 || !fz) continue;
    float4 e = lin[ft];  // e.z = w written in pass 1 by this very thread
    const float gi = g * x;
    ftrl_apply<PRECISE>(e.x, e.y, e.z, gi, gi * gi, h);
    lin[ft] = e;
  }
}

// ---------------------------------------------------------------------------------------------
// K2: one warp per chunk (<= CH occurrences of one feature row, from the sorted occurrence list).
// For every occurrence (sample s, position m) the partners n != m of the sample are revisited and
// g_s * w[feat_n][field_m,:] * x_m x_n is accumulated into the row's (sum g, sum g^2) image in
// shared memory; then either the closed-form update is applied (row fits one chunk) or the partial
// image is parked for k_ffm_combine.  Rows already finalised by k_ffm_sample are skipped.
// ---------------------------------------------------------------------------------------------
template <int VEC, bool PRECISE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_ffm_rows(Batch b, Dims d, Hyper h, float *__restrict__ tab, float4 *__restrict__ lin, int32_t ch,
           const int32_t *__restrict__ n_chunks_p, const int32_t *__restrict__ chunk_pos,
           const uint32_t *__restrict__ skey, const uint32_t *__restrict__ socc,
           const SegScan *__restrict__ scan, const int32_t *__restrict__ occ_row,
           const uint8_t *__restrict__ sflags, int fuse, const float *__restrict__ g_in,
           float *__restrict__ part, float2 *__restrict__ part_lin) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  float *acc0 = smem + (int64_t)wib * 2 * ld;  // sum g
  float *acc1 = acc0 + ld;                     // sum g^2
  const int n_chunks = *n_chunks_p;
  const int32_t nnz = (int32_t)b.nnz;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  const int C = (d.k + VEC - 1) / VEC;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid) continue;
    const bool whole_row = ci.row_head && ci.row_last;
    if (fuse && whole_row && ci.p1 - ci.p0 == 1) {
      if (sflags[occ_row[socc[ci.p0]]] & SF_FUSABLE) continue;  // finalised in k_ffm_sample
    }
    for (int64_t v = lane; v < 2 * ld; v += 32) acc0[v] = 0.f;
    __syncwarp();
    float sg = 0.f, sg2 = 0.f;  // linear coordinate (lane 0)
    for (int p = ci.p0; p < ci.p1; p++) {
      const int64_t t = socc[p];
      const int32_t s = occ_row[t];
      const float g = g_in[s];
      const int64_t r0 = b.row_ptr[s];
      const int F = (int)(b.row_ptr[s + 1] - r0);
      const int m = (int)(t - r0);
      const float xm = b.val[t];
      const int32_t fm = b.field[t];
      const bool simple = sflags[s] & SF_SIMPLE;
      const float gi = g * xm;
      sg += gi;
      sg2 = fmaf(gi, gi, sg2);
      const int n_q = F * C;
      for (int q0 = 0; q0 < n_q; q0 += 32) {
        const int q = q0 + lane;
        const int n = q / C, c4 = q - n * C;
        bool act = q < n_q && n != m;
        int32_t fn = 0, in = 0;
        float xn = 0.f;
        if (act) {
          fn = b.field[r0 + n]; in = b.feat[r0 + n]; xn = b.val[r0 + n];
          act = feat_valid(d, fn, in);
        }
        Vec<VEC> gv;
        int64_t off = 0;
        if (act) {
          Vec<VEC> w;
          w.load(tab + (int64_t)in * rs + 2 * ld + (int64_t)fm * d.k + c4 * VEC);
          const float gx = g * (xm * xn);
#pragma unroll
          for (int e = 0; e < VEC; e++) gv.v[e] = gx * w.v[e];
          off = (int64_t)fn * d.k + c4 * VEC;
        }
        if (simple) {
          if (act) {
#pragma unroll
            for (int e = 0; e < VEC; e++) {
              acc0[off + e] += gv.v[e];
              acc1[off + e] = fmaf(gv.v[e], gv.v[e], acc1[off + e]);
            }
          }
        } else {
          // partners may share a field: apply lane by lane (deterministic order)
          for (int l = 0; l < 32; l++) {
            if (l == lane && act) {
#pragma unroll
              for (int e = 0; e < VEC; e++) {
                acc0[off + e] += gv.v[e];
                acc1[off + e] = fmaf(gv.v[e], gv.v[e], acc1[off + e]);
              }
            }
            __syncwarp();
          }
        }
      }
      __syncwarp();
    }
    if (whole_row) {
      float *row = tab + (int64_t)ci.key * rs;
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
        float4 e = lin[ci.key];
        ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
        lin[ci.key] = e;
      }
    } else {
      float *dst = part + (int64_t)ci.slot * 2 * ld;
      for (int64_t v = lane; v < 2 * ld; v += 32) dst[v] = acc0[v];
      if (lane == 0) part_lin[ci.slot] = make_float2(sg, sg2);
    }
    __syncwarp();
  }
}

// K3: rows spanning several chunks: sum the parked partials in chunk order, apply once.
template <int VEC, bool PRECISE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_ffm_combine(Dims d, Hyper h, int32_t nnz, float *__restrict__ tab, float4 *__restrict__ lin, int32_t ch,
              const int32_t *__restrict__ n_chunks_p, const int32_t *__restrict__ chunk_pos,
              const uint32_t *__restrict__ skey, const SegScan *__restrict__ scan,
              const float *__restrict__ part, const float2 *__restrict__ part_lin) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid || !ci.row_head || ci.row_last) continue;
    int J = 1;  // number of chunks of this row
    while (c + J < n_chunks && skey[chunk_pos[c + J]] == ci.key) J++;
    float *row = tab + (int64_t)ci.key * rs;
    const float *p0 = part + (int64_t)ci.slot * 2 * ld;
    for (int64_t v = lane * VEC; v < d.row_len; v += 32 * VEC) {
      Vec<VEC> a0, a1;
#pragma unroll
      for (int e = 0; e < VEC; e++) a0.v[e] = a1.v[e] = 0.f;
      for (int j = 0; j < J; j++) {
        Vec<VEC> t0, t1;
        t0.load(p0 + (int64_t)j * 2 * ld + v);
        t1.load(p0 + (int64_t)j * 2 * ld + ld + v);
#pragma unroll
        for (int e = 0; e < VEC; e++) { a0.v[e] += t0.v[e]; a1.v[e] += t1.v[e]; }
      }
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
      float sg = 0.f, sg2 = 0.f;
      for (int j = 0; j < J; j++) {
        const float2 t = part_lin[ci.slot + j];
        sg += t.x; sg2 += t.y;
      }
      float4 e = lin[ci.key];
      ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
      lin[ci.key] = e;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// predict (ffm.cpp:51-70): stored w only, no state change.  One CTA per sample, any summation order.
// ---------------------------------------------------------------------------------------------
template <int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ffm_predict(Batch b, Dims d, const float *__restrict__ tab, const float4 *__restrict__ lin,
              const float4 *__restrict__ bias, const uint32_t *__restrict__ pair_lut, int output_prob,
              float *__restrict__ out, double *__restrict__ loss_out) {
  __shared__ float red[33];
  const int tid = threadIdx.x;
  const int64_t s = blockIdx.x;
  const int64_t r0 = b.row_ptr[s];
  const int F = (int)(b.row_ptr[s + 1] - r0);
  const int C = (d.k + VEC - 1) / VEC;
  const int64_t n_items = (int64_t)F * (F - 1) / 2 * C;
  const int64_t ld = d.ld, rs = 3 * ld;
  float acc = 0.f;
  for (int64_t it = tid; it < n_items; it += THREADS) {
    const int c = (int)(it % C);
    int m, n;
    pair_decode(it / C, pair_lut, m, n);
    const int32_t fm = b.field[r0 + m], im = b.feat[r0 + m], fn = b.field[r0 + n], in = b.feat[r0 + n];
    if (!feat_valid(d, fm, im) || !feat_valid(d, fn, in)) continue;
    Vec<VEC> wA, wB;
    wA.load(tab + (int64_t)im * rs + 2 * ld + (int64_t)fn * d.k + c * VEC);
    wB.load(tab + (int64_t)in * rs + 2 * ld + (int64_t)fm * d.k + c * VEC);
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; e++) dot = fmaf(wA.v[e], wB.v[e], dot);
    acc = fmaf(dot, b.val[r0 + m] * b.val[r0 + n], acc);
  }
  for (int t = tid; t < F; t += THREADS) {
    const int32_t fl = b.field[r0 + t], ft = b.feat[r0 + t];
    if (feat_valid(d, fl, ft)) acc = fmaf(lin[ft].z, b.val[r0 + t], acc);
  }
  float logit = block_sum(acc, red);
  if (tid == 0) {
    logit += bias->z;
    out[s] = output_prob ? sigmoid_f(logit) : logit;
    if (loss_out) loss_out[s] = b.label ? logloss_d(b.label[s], logit) : 0.0;
  }
}

}  // namespace ftrl
