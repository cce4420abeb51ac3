// exact.cuh -- FTRL_MODE_SEQUENTIAL: the reference's per-sample trajectory reproduced on device.
//
// One CTA walks the samples of a call in row order; inside a sample the reference's fp32
// operation order is kept literally (every product/sum below is an explicit round-to-nearest
// intrinsic, so nothing is contracted into FMA or re-associated), including
//   * maybe_zero_weight's double-precision quotient (ftrl_model.h:29-33),
//   * the `z += gi - si*wi` association of the linear update (ftrl_model.cpp:73) versus the
//     `z + g - s*w` association of the latent updates (ffm.cpp:114, fm.cpp:91),
//   * the ffm.cpp:118 term sqrtf(n2 + g2*g1).
// Threads only share work where the reference's sequential order cannot be observed: distinct
// coordinates, with every fp32 sum still evaluated in source order by a single thread.
// This is the parity path (batch size 1 semantics), not the throughput path.
#pragma once
#include "common.cuh"
#include "prep.cuh"

namespace ftrl {

constexpr int EX_THREADS = 256;
constexpr int EX_CAP = 96;                        // valid features per sample on the parallel path
constexpr int EX_TERMS = EX_CAP * (EX_CAP - 1) / 2;  // pair terms kept in shared memory
constexpr int EX_KCAP = 1024;                     // FM factors kept in shared memory

__device__ __forceinline__ float ex_weight(float n, float z, const Hyper &h) {
  if (fabsf(z) <= h.l1) return 0.0f;
  const float sg = z > 0.f ? 1.0f : -1.0f;
  const float num = __fsub_rn(z, __fmul_rn(sg, h.l1));
  const float den = __fadd_rn(h.l2, __fdiv_rn(__fadd_rn(h.beta, __fsqrt_rn(n)), h.alpha));
  return __double2float_rn(__ddiv_rn(-1.0 * (double)num, (double)den));
}

// 1 / (1 + expf(-x)); expf taken as the correctly rounded value via fp64 (glibc's expf is
// correctly rounded in all but ~1e-3 ulp-midpoint cases)
__device__ __forceinline__ float ex_sigmoid(float x) {
  const float e = __double2float_rn(exp((double)(-x)));
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

// (sqrtf(n + a*b) - sqrtf(n)) / alpha
__device__ __forceinline__ float ex_sigma(float n, float a, float b, const Hyper &h) {
  return __fdiv_rn(__fsub_rn(__fsqrt_rn(__fadd_rn(n, __fmul_rn(a, b))), __fsqrt_rn(n)), h.alpha);
}

struct ExShared {
  int32_t fld[EX_CAP];
  int32_t ft[EX_CAP];
  float x[EX_CAP];
  float terms[EX_TERMS > EX_KCAP ? EX_TERMS : EX_KCAP];
  float sum_vx[EX_KCAP];
  int Fv;          // valid features (compacted, in order)
  int parallel;    // 1: coordinates of this sample are pairwise distinct
  float logit, g;
};

// lexicographic (m outer, n inner) pair ordinal -> (m, n)
__device__ __forceinline__ void lex_decode(int p, int F, int &m, int &n) {
  int mm = 0, rem = p;
  while (rem >= F - 1 - mm) { rem -= F - 1 - mm; mm++; }
  m = mm;
  n = mm + 1 + rem;
}

// one FFM pair, all factors, reference order (ffm.cpp:90-136); safe for overlapping slices
// because all reads of the pair precede its writes only per factor -- the reference reads the
// whole pair first, so buffer when slices may alias.
__device__ __forceinline__ void ex_ffm_pair_factor(float *tab, int64_t a, int64_t b, int64_t ld, float g,
                                                   float x, const Hyper &h, float &z1, float &n1,
                                                   float &z2, float &n2) {
  const float vif1 = tab[a + 2 * ld], v_nif1 = tab[a + ld], v_zif1 = tab[a];
  const float vif2 = tab[b + 2 * ld], v_nif2 = tab[b + ld], v_zif2 = tab[b];
  const float v_gif1 = __fmul_rn(__fmul_rn(g, vif2), x);
  const float v_sif1 = ex_sigma(v_nif1, v_gif1, v_gif1, h);
  z1 = __fsub_rn(__fadd_rn(v_zif1, v_gif1), __fmul_rn(v_sif1, vif1));
  n1 = __fadd_rn(v_nif1, __fmul_rn(v_gif1, v_gif1));
  const float v_gif2 = __fmul_rn(__fmul_rn(g, vif1), x);
  const float v_sif2 = ex_sigma(v_nif2, v_gif2, v_gif1, h);  // ffm.cpp:118
  z2 = __fsub_rn(__fadd_rn(v_zif2, v_gif2), __fmul_rn(v_sif2, vif2));
  n2 = __fadd_rn(v_nif2, __fmul_rn(v_gif2, v_gif2));
}

// A sample the shared-memory path cannot hold (more than EX_CAP valid features, or FM with more than EX_KCAP
// factors): one thread walks it straight from the CSR arrays in the reference's order -- the reference has no
// cap (ffm.cpp:57-70, fm.cpp:40-67).  Same operations, same order as the paths of k_exact_train below.
__device__ void ex_train_sample_serial(const Batch &b, const Dims &d, const Hyper &h, int64_t r0, int64_t r1, int y,
                                       float *tab, float4 *lin, float4 *bias, float &logit_o) {
  const int64_t ld = d.ld, rs = 3 * ld;
  const int k = d.k;
  auto valid = [&](int64_t t) { return feat_valid(d, b.field[t], b.feat[t]); };
  // ---- materialise w (idempotent: repeated coordinates give the same value) ----
  for (int64_t t = r0; t < r1; t++) {
    if (!valid(t)) continue;
    float4 e = lin[b.feat[t]];
    lin[b.feat[t]].z = ex_weight(e.y, e.x, h);
  }
  {
    float4 e = *bias;
    e.z = ex_weight(e.y, e.x, h);
    *bias = e;
  }
  if (d.model_type == 2) {
    for (int64_t m = r0; m < r1; m++) {
      if (!valid(m)) continue;
      for (int64_t n = m + 1; n < r1; n++) {
        if (!valid(n)) continue;
        const int64_t a = (int64_t)b.feat[m] * rs + (int64_t)b.field[n] * k;
        const int64_t c = (int64_t)b.feat[n] * rs + (int64_t)b.field[m] * k;
        for (int f = 0; f < k; f++) {
          tab[a + f + 2 * ld] = ex_weight(tab[a + f + ld], tab[a + f], h);
          tab[c + f + 2 * ld] = ex_weight(tab[c + f + ld], tab[c + f], h);
        }
      }
    }
  } else if (d.model_type == 1) {
    for (int64_t t = r0; t < r1; t++) {
      if (!valid(t)) continue;
      const int64_t a = (int64_t)b.feat[t] * rs;
      for (int f = 0; f < k; f++) tab[a + f + 2 * ld] = ex_weight(tab[a + f + ld], tab[a + f], h);
    }
  }
  // ---- logit: bias, linear terms, then the pair / factor terms in source order ----
  float acc = bias->z;
  for (int64_t t = r0; t < r1; t++)
    if (valid(t)) acc = __fadd_rn(acc, __fmul_rn(lin[b.feat[t]].z, b.val[t]));
  if (d.model_type == 2) {
    for (int64_t m = r0; m < r1; m++) {
      if (!valid(m)) continue;
      for (int64_t n = m + 1; n < r1; n++) {
        if (!valid(n)) continue;
        const float *wa = tab + (int64_t)b.feat[m] * rs + 2 * ld + (int64_t)b.field[n] * k;
        const float *wb = tab + (int64_t)b.feat[n] * rs + 2 * ld + (int64_t)b.field[m] * k;
        float dot = 0.0f;
        for (int f = 0; f < k; f++) dot = __fadd_rn(dot, __fmul_rn(wa[f], wb[f]));
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(dot, b.val[m]), b.val[n]));
      }
    }
  } else if (d.model_type == 1) {
    for (int f = 0; f < k; f++) {
      float s_vx = 0.0f, sum_sqr = 0.0f;
      for (int64_t t = r0; t < r1; t++) {
        if (!valid(t)) continue;
        const float vx = __fmul_rn(tab[(int64_t)b.feat[t] * rs + 2 * ld + f], b.val[t]);
        s_vx = __fadd_rn(s_vx, vx);
        sum_sqr = __fadd_rn(sum_sqr, __fmul_rn(vx, vx));
      }
      acc = __fadd_rn(acc, __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s_vx, s_vx), sum_sqr)));
    }
  }
  logit_o = acc;
  const float g = __fsub_rn(ex_sigmoid(acc), (float)y);
  // ---- update_linear_nz + update_bias_nz (ftrl_model.cpp:66-85) ----
  for (int64_t t = r0; t < r1; t++) {
    if (!valid(t)) continue;
    float4 e = lin[b.feat[t]];
    const float gi = __fmul_rn(g, b.val[t]);
    const float si = ex_sigma(e.y, gi, gi, h);
    e.x = __fadd_rn(e.x, __fsub_rn(gi, __fmul_rn(si, e.z)));
    e.y = __fadd_rn(e.y, __fmul_rn(gi, gi));
    lin[b.feat[t]] = e;
  }
  {
    float4 e = *bias;
    const float si = ex_sigma(e.y, g, g, h);
    e.x = __fadd_rn(e.x, __fsub_rn(g, __fmul_rn(si, e.z)));
    e.y = __fadd_rn(e.y, __fmul_rn(g, g));
    *bias = e;
  }
  // ---- latent n,z updates ----
  if (d.model_type == 2) {
    for (int64_t m = r0; m < r1; m++) {
      if (!valid(m)) continue;
      for (int64_t n = m + 1; n < r1; n++) {
        if (!valid(n)) continue;
        const float x = __fmul_rn(b.val[m], b.val[n]);
        const int64_t a0 = (int64_t)b.feat[m] * rs + (int64_t)b.field[n] * k;
        const int64_t c0 = (int64_t)b.feat[n] * rs + (int64_t)b.field[m] * k;
        const bool alias = a0 == c0;
        for (int f = 0; f < k; f++) {
          float z1, n1, z2, n2;
          ex_ffm_pair_factor(tab, a0 + f, c0 + f, ld, g, x, h, z1, n1, z2, n2);
          if (!alias) { tab[a0 + f] = z1; tab[a0 + f + ld] = n1; }
          tab[c0 + f] = z2; tab[c0 + f + ld] = n2;
        }
      }
    }
  } else if (d.model_type == 1) {
    // fm.cpp:80-101.  sum_vx of a factor is recomputed from the stored w (the n,z updates do not touch w)
    for (int f = 0; f < k; f++) {
      float s_vx = 0.0f;
      for (int64_t t = r0; t < r1; t++)
        if (valid(t)) s_vx = __fadd_rn(s_vx, __fmul_rn(tab[(int64_t)b.feat[t] * rs + 2 * ld + f], b.val[t]));
      for (int64_t t = r0; t < r1; t++) {
        if (!valid(t)) continue;
        const int64_t a = (int64_t)b.feat[t] * rs + f;
        const float x = b.val[t];
        const float vif = tab[a + 2 * ld], v_nif = tab[a + ld], v_zif = tab[a];
        const float v_gif = __fmul_rn(g, __fsub_rn(__fmul_rn(x, s_vx), __fmul_rn(__fmul_rn(vif, x), x)));
        const float v_sif = ex_sigma(v_nif, v_gif, v_gif, h);
        tab[a] = __fsub_rn(__fadd_rn(v_zif, v_gif), __fmul_rn(v_sif, vif));
        tab[a + ld] = __fadd_rn(v_nif, __fmul_rn(v_gif, v_gif));
      }
    }
  }
}

__global__ void __launch_bounds__(EX_THREADS)
k_exact_train(Batch b, Dims d, Hyper h, float *__restrict__ tab, float4 *__restrict__ lin,
              float4 *__restrict__ bias, float *__restrict__ logit_out, double *__restrict__ loss_sum_out,
              int32_t *__restrict__ err) {
  __shared__ ExShared sh;
  const int tid = threadIdx.x;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int k = d.k;
  double loss_sum = 0.0;  // thread 0
  for (int64_t s = 0; s < b.n_rows; s++) {
    const int64_t r0 = b.row_ptr[s];
    const int F = (int)(b.row_ptr[s + 1] - r0);
    __syncthreads();
    if (tid == 0) {
      // remove_out_range (ftrl_model.cpp:36-42, ffm.cpp:30-36): compact, keep order
      int fv = 0;
      bool overflow = false;
      for (int t = 0; t < F; t++) {
        const int32_t fl = b.field[r0 + t], ft = b.feat[r0 + t];
        if (!feat_valid(d, fl, ft)) continue;
        if (fv < EX_CAP) {
          sh.fld[fv] = fl; sh.ft[fv] = ft; sh.x[fv] = b.val[r0 + t];
        } else {
          overflow = true;
        }
        fv++;
      }
      if (overflow || (d.model_type == 1 && k > EX_KCAP)) {
        // wider than the shared-memory path: the whole sample serially, straight from the CSR arrays
        float lg;
        const int y = b.label[s];
        ex_train_sample_serial(b, d, h, r0, r0 + F, y, tab, lin, bias, lg);
        if (logit_out) logit_out[s] = lg;
        loss_sum += logloss_d(y, lg);
        fv = -1;
      }
      sh.Fv = fv;
      bool par = true;
      for (int u = 0; u < fv && par; u++)
        for (int v = u + 1; v < fv; v++)
          if (sh.ft[u] == sh.ft[v] || sh.fld[u] == sh.fld[v]) { par = false; break; }
      sh.parallel = par ? 1 : 0;
    }
    __syncthreads();
    const int Fv = sh.Fv;
    if (Fv < 0) continue;  // done by ex_train_sample_serial (uniform: every thread reads the same sh.Fv)
    const int P = Fv * (Fv - 1) / 2;

    // ---- materialise w: update_linear_w, update_bias, update_vector_w (idempotent) ----
    for (int t = tid; t < Fv; t += EX_THREADS) {
      float4 e = lin[sh.ft[t]];
      e.z = ex_weight(e.y, e.x, h);
      lin[sh.ft[t]].z = e.z;
    }
    if (tid == 0) {
      float4 e = *bias;
      e.z = ex_weight(e.y, e.x, h);
      *bias = e;
    }
    if (d.model_type == 2) {
      for (int it = tid; it < P * k; it += EX_THREADS) {
        int m, n;
        lex_decode(it / k, Fv, m, n);
        const int f = it % k;
        const int64_t a = (int64_t)sh.ft[m] * rs + (int64_t)sh.fld[n] * k + f;
        const int64_t c = (int64_t)sh.ft[n] * rs + (int64_t)sh.fld[m] * k + f;
        tab[a + 2 * ld] = ex_weight(tab[a + ld], tab[a], h);
        tab[c + 2 * ld] = ex_weight(tab[c + ld], tab[c], h);
      }
    } else if (d.model_type == 1) {
      for (int it = tid; it < Fv * k; it += EX_THREADS) {
        const int64_t a = (int64_t)sh.ft[it / k] * rs + it % k;
        tab[a + 2 * ld] = ex_weight(tab[a + ld], tab[a], h);
      }
    }
    __syncthreads();

    // ---- logit terms ----
    if (d.model_type == 2) {
      for (int p = tid; p < P; p += EX_THREADS) {
        int m, n;
        lex_decode(p, Fv, m, n);
        const float *wa = tab + (int64_t)sh.ft[m] * rs + 2 * ld + (int64_t)sh.fld[n] * k;
        const float *wb = tab + (int64_t)sh.ft[n] * rs + 2 * ld + (int64_t)sh.fld[m] * k;
        float dot = 0.0f;
        for (int f = 0; f < k; f++) dot = __fadd_rn(dot, __fmul_rn(wa[f], wb[f]));
        sh.terms[p] = __fmul_rn(__fmul_rn(dot, sh.x[m]), sh.x[n]);
      }
    } else if (d.model_type == 1) {
      for (int f = tid; f < k; f += EX_THREADS) {
        float s_vx = 0.0f, sum_sqr = 0.0f;
        for (int t = 0; t < Fv; t++) {
          const float vx = __fmul_rn(tab[(int64_t)sh.ft[t] * rs + 2 * ld + f], sh.x[t]);
          s_vx = __fadd_rn(s_vx, vx);
          sum_sqr = __fadd_rn(sum_sqr, __fmul_rn(vx, vx));
        }
        sh.sum_vx[f] = s_vx;
        sh.terms[f] = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s_vx, s_vx), sum_sqr));
      }
    }
    __syncthreads();
    if (tid == 0) {
      float acc = bias->z;
      for (int t = 0; t < Fv; t++) acc = __fadd_rn(acc, __fmul_rn(lin[sh.ft[t]].z, sh.x[t]));
      const int nt = d.model_type == 2 ? P : d.model_type == 1 ? k : 0;
      for (int p = 0; p < nt; p++) acc = __fadd_rn(acc, sh.terms[p]);
      const int y = b.label[s];
      sh.logit = acc;
      sh.g = __fsub_rn(ex_sigmoid(acc), (float)y);
      if (logit_out) logit_out[s] = acc;
      loss_sum += logloss_d(y, acc);
      // update_linear_nz + update_bias_nz (ftrl_model.cpp:66-85), sequential: repeated ids see
      // the earlier update of n,z but the same w
      const float g = sh.g;
      for (int t = 0; t < Fv; t++) {
        float4 e = lin[sh.ft[t]];
        const float gi = __fmul_rn(g, sh.x[t]);
        const float si = ex_sigma(e.y, gi, gi, h);
        e.x = __fadd_rn(e.x, __fsub_rn(gi, __fmul_rn(si, e.z)));
        e.y = __fadd_rn(e.y, __fmul_rn(gi, gi));
        lin[sh.ft[t]] = e;
      }
      float4 e = *bias;
      const float si = ex_sigma(e.y, g, g, h);
      e.x = __fadd_rn(e.x, __fsub_rn(g, __fmul_rn(si, e.z)));
      e.y = __fadd_rn(e.y, __fmul_rn(g, g));
      *bias = e;
    }
    __syncthreads();
    const float g = sh.g;

    // ---- latent n,z updates ----
    if (d.model_type == 2) {
      if (sh.parallel) {
        for (int it = tid; it < P * k; it += EX_THREADS) {
          int m, n;
          lex_decode(it / k, Fv, m, n);
          const int f = it % k;
          const int64_t a = (int64_t)sh.ft[m] * rs + (int64_t)sh.fld[n] * k + f;
          const int64_t c = (int64_t)sh.ft[n] * rs + (int64_t)sh.fld[m] * k + f;
          float z1, n1, z2, n2;
          ex_ffm_pair_factor(tab, a, c, ld, g, __fmul_rn(sh.x[m], sh.x[n]), h, z1, n1, z2, n2);
          tab[a] = z1; tab[a + ld] = n1; tab[c] = z2; tab[c + ld] = n2;
        }
      } else if (tid == 0) {
        // serial, reference order; per pair all reads precede the writes (ffm.cpp:98-133)
        for (int m = 0; m < Fv; m++)
          for (int n = m + 1; n < Fv; n++) {
            const float x = __fmul_rn(sh.x[m], sh.x[n]);
            const int64_t a0 = (int64_t)sh.ft[m] * rs + (int64_t)sh.fld[n] * k;
            const int64_t c0 = (int64_t)sh.ft[n] * rs + (int64_t)sh.fld[m] * k;
            const bool alias = a0 == c0;
            for (int f = 0; f < k; f++) {
              float z1, n1, z2, n2;
              ex_ffm_pair_factor(tab, a0 + f, c0 + f, ld, g, x, h, z1, n1, z2, n2);
              // distinct factors never alias; a fully aliased pair keeps slice 2 (written last)
              if (!alias) { tab[a0 + f] = z1; tab[a0 + f + ld] = n1; }
              tab[c0 + f] = z2; tab[c0 + f + ld] = n2;
            }
          }
      }
    } else if (d.model_type == 1) {
      // fm.cpp:80-101: per feature in order; factor f is owned by one thread
      for (int f = tid; f < k; f += EX_THREADS) {
        const float s_vx = sh.sum_vx[f];
        for (int t = 0; t < Fv; t++) {
          const int64_t a = (int64_t)sh.ft[t] * rs + f;
          const float x = sh.x[t];
          const float vif = tab[a + 2 * ld], v_nif = tab[a + ld], v_zif = tab[a];
          const float v_gif =
              __fmul_rn(g, __fsub_rn(__fmul_rn(x, s_vx), __fmul_rn(__fmul_rn(vif, x), x)));
          const float v_sif = ex_sigma(v_nif, v_gif, v_gif, h);
          tab[a] = __fsub_rn(__fadd_rn(v_zif, v_gif), __fmul_rn(v_sif, vif));
          tab[a + ld] = __fadd_rn(v_nif, __fmul_rn(v_gif, v_gif));
        }
      }
    }
  }
  if (tid == 0 && loss_sum_out) *loss_sum_out = loss_sum;
}

// predict in the reference's summation order (lr.cpp:20-24, fm.cpp:34-67, ffm.cpp:51-70):
// one thread per sample, everything sequential.
__global__ void __launch_bounds__(128)
k_exact_predict(Batch b, Dims d, const float *__restrict__ tab, const float4 *__restrict__ lin,
                const float4 *__restrict__ bias, int output_prob, float *__restrict__ out,
                float *__restrict__ logit_out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= b.n_rows) return;
  const int64_t r0 = b.row_ptr[s], r1 = b.row_ptr[s + 1];
  const int64_t ld = d.ld, rs = 3 * ld;
  const int k = d.k;
  float acc = bias->z;
  for (int64_t t = r0; t < r1; t++) {
    const int32_t fl = b.field[t], ft = b.feat[t];
    if (feat_valid(d, fl, ft)) acc = __fadd_rn(acc, __fmul_rn(lin[ft].z, b.val[t]));
  }
  if (d.model_type == 2) {
    for (int64_t m = r0; m < r1; m++) {
      const int32_t fm = b.field[m], im = b.feat[m];
      if (!feat_valid(d, fm, im)) continue;
      for (int64_t n = m + 1; n < r1; n++) {
        const int32_t fn = b.field[n], in = b.feat[n];
        if (!feat_valid(d, fn, in)) continue;
        const float *wa = tab + (int64_t)im * rs + 2 * ld + (int64_t)fn * k;
        const float *wb = tab + (int64_t)in * rs + 2 * ld + (int64_t)fm * k;
        float dot = 0.0f;
        for (int f = 0; f < k; f++) dot = __fadd_rn(dot, __fmul_rn(wa[f], wb[f]));
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(dot, b.val[m]), b.val[n]));
      }
    }
  } else if (d.model_type == 1) {
    for (int f = 0; f < k; f++) {
      float s_vx = 0.0f, sum_sqr = 0.0f;
      for (int64_t t = r0; t < r1; t++) {
        const int32_t ft = b.feat[t];
        if (ft < 0 || ft >= d.n_feats) continue;
        const float vx = __fmul_rn(tab[(int64_t)ft * rs + 2 * ld + f], b.val[t]);
        s_vx = __fadd_rn(s_vx, vx);
        sum_sqr = __fadd_rn(sum_sqr, __fmul_rn(vx, vx));
      }
      acc = __fadd_rn(acc, __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s_vx, s_vx), sum_sqr)));
    }
  }
  out[s] = output_prob ? ex_sigmoid(acc) : acc;
  if (logit_out) logit_out[s] = acc;
}

}  // namespace ftrl
