// lr_fm.cuh -- LR and FM minibatch kernels + the batch-level bias / loss reduction shared by all
// three models.
//
// Reference functions covered: ftrl_model.cpp:44-85 (linear logit, linear/bias w and n,z updates),
// lr.cpp:9-24, fm.cpp:21-101 (compute_fm_logit :40-67, update_vector_w :69-78,
// update_vector_nz :80-101).  Minibatch semantics per SURVEY.md 8(a).
#pragma once
#include "common.cuh"
#include "prep.cuh"

namespace ftrl {

// ---------------------------------------------------------------------------------------------
// sample kernel, warp per sample.  LR: logit from lin.  FM: additionally the O(F k) sum/square
// trick of fm.cpp:40-67; the per-factor sums S[s][f] are kept for the row kernel (the reference
// keeps them in the member sum_vx, fm.h:24).
// Lane mapping for FM: lane = (feature slot j, factor chunk c), C = k/VEC chunks, 32/C slots.
// ---------------------------------------------------------------------------------------------
template <int VEC, bool PRECISE, bool IS_FM>
__global__ void __launch_bounds__(256)
k_lrfm_sample(Batch b, Dims d, Hyper h, float *__restrict__ tab, float4 *__restrict__ lin,
              const float4 *__restrict__ bias, float *__restrict__ S, float *__restrict__ g_out,
              float *__restrict__ logit_out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= b.n_rows) return;
  const int64_t r0 = b.row_ptr[s];
  const int F = (int)(b.row_ptr[s + 1] - r0);
  float acc = 0.f;
  for (int t = lane; t < F; t += 32) {
    const int32_t ft = b.feat[r0 + t];
    if (ft < 0 || ft >= d.n_feats) continue;
    const float4 e = lin[ft];
    // (the stale-by-one w the reference keeps is stored once per row by the row kernels: a store per OCCURRENCE
    // serialises thousands of writes on the sectors of the hot rows)
    const float w = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
    acc = fmaf(w, b.val[r0 + t], acc);
  }
  if (IS_FM) {
    const int C = d.k / VEC;
    const int64_t ld = d.ld, rs = 3 * ld;
    // chunks of the factor axis are covered in rounds of 32 lanes when C > 32
    for (int cb = 0; cb < C; cb += 32) {
      const int Cr = min(C - cb, 32);      // chunks in this round
      const int slots = 32 / Cr;           // features processed concurrently
      const int j = lane / Cr, c = cb + lane % Cr;
      const bool lane_on = j < slots;
      Vec<VEC> sv, qv;
#pragma unroll
      for (int e = 0; e < VEC; e++) sv.v[e] = qv.v[e] = 0.f;
      for (int t0 = 0; t0 < F; t0 += slots) {
        const int t = t0 + j;
        if (!lane_on || t >= F) continue;
        const int32_t ft = b.feat[r0 + t];
        if (ft < 0 || ft >= d.n_feats) continue;
        const float x = b.val[r0 + t];
        float *row = tab + (int64_t)ft * rs + c * VEC;
        Vec<VEC> z, n, w;
        z.load(row); n.load(row + ld);
#pragma unroll
        for (int e = 0; e < VEC; e++) {
          w.v[e] = weight_from<PRECISE>(z.v[e], f_sqrt<PRECISE>(n.v[e]), h);
          const float vx = w.v[e] * x;
          sv.v[e] += vx;
          qv.v[e] = fmaf(vx, vx, qv.v[e]);
        }
      }
      // reduce over the feature slots (lanes with equal lane % Cr); Cr need not be a power of two:
      // gather through shuffles from every slot in a fixed order.
      Vec<VEC> st, qt;
#pragma unroll
      for (int e = 0; e < VEC; e++) st.v[e] = qt.v[e] = 0.f;
      for (int jj = 0; jj < slots; jj++) {
        const int src = jj * Cr + lane % Cr;
#pragma unroll
        for (int e = 0; e < VEC; e++) {
          st.v[e] += __shfl_sync(0xffffffffu, sv.v[e], src);
          qt.v[e] += __shfl_sync(0xffffffffu, qv.v[e], src);
        }
      }
      if (lane < Cr) {
#pragma unroll
        for (int e = 0; e < VEC; e++) acc += 0.5f * (st.v[e] * st.v[e] - qt.v[e]);
        st.store(S + s * (int64_t)d.k + (int64_t)(cb + lane) * VEC);
      }
    }
  }
  float logit = warp_sum(acc);
  if (lane == 0) {
    const float4 bz = *bias;
    logit += weight_from<PRECISE>(bz.x, f_sqrt<PRECISE>(bz.y), h);
    const int y = b.label[s];
    g_out[s] = sigmoid_f(logit) - (float)y;
    logit_out[s] = logit;
  }
}

// ---------------------------------------------------------------------------------------------
// sharded runs (shard.cuh): the same sample kernel over materialised w.  The owner of every row the GLOBAL batch
// touches has written w = W(n,z) into its shard (lin.z, the w plane of tab) and pushed it into the row cache of
// every remote rank that touches the row (rc_lin / rc_w, indexed by the row's sorted head position on that rank),
// so a sample only reads local memory: shard rows for the ids this rank owns, the cache for the others.
// ---------------------------------------------------------------------------------------------
template <int VEC, bool PRECISE, bool IS_FM>
__global__ void __launch_bounds__(256)
k_lrfm_sample_sh(Batch b, Dims d, Hyper h, const __grid_constant__ RowSpace rsp, const int32_t *__restrict__ occ_pos,
                 const SegScan *__restrict__ scan, const float4 *__restrict__ bias, float *__restrict__ S,
                 float *__restrict__ g_out, float *__restrict__ logit_out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= b.n_rows) return;
  const int64_t r0 = b.row_ptr[s];
  const int F = (int)(b.row_ptr[s + 1] - r0);
  const int64_t ld = d.ld, rs = 3 * ld;
  // w of occurrence t: (linear w, latent w row)
  auto locate = [&](int64_t t, int32_t ft, float &wl, const float *&wrow) {
    if ((ft & rsp.Gm1) == rsp.rank) {
      const int64_t loc = ft >> rsp.log2G;
      wl = rsp.lin[loc].z;
      wrow = rsp.tab + loc * rs + 2 * ld;
    } else {
      const int32_t head = scan[occ_pos[t]].start;
      wl = rsp.rc_lin[head];
      wrow = rsp.rc_w + (int64_t)head * ld;
    }
  };
  float acc = 0.f;
  for (int t = lane; t < F; t += 32) {
    const int32_t ft = b.feat[r0 + t];
    if (ft < 0 || ft >= d.n_feats) continue;
    float wl;
    const float *wrow;
    locate(r0 + t, ft, wl, wrow);
    acc = fmaf(wl, b.val[r0 + t], acc);
  }
  if (IS_FM) {
    const int C = d.k / VEC;
    for (int cb = 0; cb < C; cb += 32) {
      const int Cr = min(C - cb, 32);
      const int slots = 32 / Cr;
      const int j = lane / Cr, c = cb + lane % Cr;
      const bool lane_on = j < slots;
      Vec<VEC> sv, qv;
#pragma unroll
      for (int e = 0; e < VEC; e++) sv.v[e] = qv.v[e] = 0.f;
      for (int t0 = 0; t0 < F; t0 += slots) {
        const int t = t0 + j;
        if (!lane_on || t >= F) continue;
        const int32_t ft = b.feat[r0 + t];
        if (ft < 0 || ft >= d.n_feats) continue;
        const float x = b.val[r0 + t];
        float wl;
        const float *wrow;
        locate(r0 + t, ft, wl, wrow);
        Vec<VEC> w;
        w.load(wrow + c * VEC);
#pragma unroll
        for (int e = 0; e < VEC; e++) {
          const float vx = w.v[e] * x;
          sv.v[e] += vx;
          qv.v[e] = fmaf(vx, vx, qv.v[e]);
        }
      }
      Vec<VEC> st, qt;
#pragma unroll
      for (int e = 0; e < VEC; e++) st.v[e] = qt.v[e] = 0.f;
      for (int jj = 0; jj < slots; jj++) {
        const int src = jj * Cr + lane % Cr;
#pragma unroll
        for (int e = 0; e < VEC; e++) {
          st.v[e] += __shfl_sync(0xffffffffu, sv.v[e], src);
          qt.v[e] += __shfl_sync(0xffffffffu, qv.v[e], src);
        }
      }
      if (lane < Cr) {
#pragma unroll
        for (int e = 0; e < VEC; e++) acc += 0.5f * (st.v[e] * st.v[e] - qt.v[e]);
        st.store(S + s * (int64_t)d.k + (int64_t)(cb + lane) * VEC);
      }
    }
  }
  float logit = warp_sum(acc);
  if (lane == 0) {
    const float4 bz = *bias;
    logit += weight_from<PRECISE>(bz.x, f_sqrt<PRECISE>(bz.y), h);
    const int y = b.label[s];
    g_out[s] = sigmoid_f(logit) - (float)y;
    logit_out[s] = logit;
  }
}

// ---------------------------------------------------------------------------------------------
// row kernel, warp per chunk of one feature's occurrences (sorted list).  Linear coordinate:
// lanes stride over occurrences, fixed-order shuffle reduction.  FM latent row: lane owns
// coordinate(s) f = lane, lane+32, ...; gv = g (x S_f - (v x) x)  (fm.cpp:89).
// ---------------------------------------------------------------------------------------------
constexpr int FM_MAX_REGS = 8;  // latent coordinates per lane kept in registers (k <= 256)

// SH (sharded runs): w is the materialised one (shard w plane / row cache), and a row's sum goes where
// Export::dst_at says: the owner's inbox, or -- this rank owns the row and is its only contributor -- applied here.
template <bool PRECISE, bool IS_FM, int WARPS, bool SH = false>
__global__ void __launch_bounds__(WARPS * 32)
k_lrfm_rows(Batch b, Dims d, Hyper h, float *__restrict__ tab, float4 *__restrict__ lin, int32_t ch,
            const int32_t *__restrict__ n_chunks_p, const int32_t *__restrict__ chunk_pos,
            const uint32_t *__restrict__ skey, const uint32_t *__restrict__ socc,
            const SegScan *__restrict__ scan, const int32_t *__restrict__ occ_row,
            const float *__restrict__ g_in, const float *__restrict__ S, float *__restrict__ part,
            float2 *__restrict__ part_lin, const __grid_constant__ RowSpace rsp, const __grid_constant__ Export ex,
            const int32_t *__restrict__ batch_flags) {
  if (SH && batch_flags[1] != 0) return;  // sharded run: the step was called off (shard.cuh: k_check_abort)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const int32_t nnz = (int32_t)b.nnz;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info<false>(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid) continue;
    const bool whole_row = ci.row_head && ci.row_last;
    // linear coordinate
    float sg = 0.f, sg2 = 0.f;
    for (int p = ci.p0 + lane; p < ci.p1; p += 32) {
      const int64_t t = socc[p];
      const float gi = g_in[occ_row[t]] * b.val[t];
      sg += gi;
      sg2 = fmaf(gi, gi, sg2);
    }
    sg = warp_sum(sg);
    sg2 = warp_sum(sg2);
    // sharded: the row's sorted head position (its slot in the row cache / dst_at), its owner and local index
    const int32_t head = SH ? scan[ci.p0].start : 0;
    const bool mine = !SH || (int)(ci.key & (uint32_t)rsp.Gm1) == rsp.rank;
    const int64_t lrow = SH ? (int64_t)(ci.key >> rsp.log2G) : (int64_t)ci.key;
    // FM latent row
    float a0[FM_MAX_REGS], a1[FM_MAX_REGS], wv[FM_MAX_REGS];
    if (IS_FM) {
      const float *row = tab + lrow * rs;
#pragma unroll
      for (int r = 0; r < FM_MAX_REGS; r++) {
        a0[r] = a1[r] = 0.f;
        const int f = lane + 32 * r;
        if (SH) {
          wv[r] = f < d.k ? (mine ? row[2 * ld + f] : rsp.rc_w[(int64_t)head * ld + f]) : 0.f;
        } else {
          // w = W(n, z) of the pre-update state (what the sample kernel used); stored once per row, by its head chunk
          wv[r] = f < d.k ? weight_from<PRECISE>(row[f], f_sqrt<PRECISE>(row[ld + f]), h) : 0.f;
          if (f < d.k && ci.row_head) const_cast<float *>(row)[2 * ld + f] = wv[r];
        }
      }
      for (int p = ci.p0; p < ci.p1; p++) {
        const int64_t t = socc[p];
        const int32_t s = occ_row[t];
        const float g = g_in[s], x = b.val[t];
#pragma unroll
        for (int r = 0; r < FM_MAX_REGS; r++) {
          const int f = lane + 32 * r;
          if (f < d.k) {
            const float gv = g * (x * S[(int64_t)s * d.k + f] - (wv[r] * x) * x);
            a0[r] += gv;
            a1[r] = fmaf(gv, gv, a1[r]);
          }
        }
      }
    }
    const int32_t dst = (SH && whole_row) ? ex.dst_at[head] : -2;
    if (whole_row && dst >= 0) {
      // (sum g, sum g^2) into the owner's inbox
      const int q = (int)(ci.key & (uint32_t)ex.Gm1);
      if (IS_FM) {
        float *o = ex.inbox[q] + (int64_t)dst * 2 * ld;
#pragma unroll
        for (int r = 0; r < FM_MAX_REGS; r++) {
          const int f = lane + 32 * r;
          if (f < d.k) {
            o[f] = a0[r];
            o[ld + f] = a1[r];
          }
        }
      }
      if (lane == 0) ex.inbox_lin[q][dst] = make_float2(sg, sg2);
    } else if (whole_row) {
      if (IS_FM) {
        float *row = tab + lrow * rs;
#pragma unroll
        for (int r = 0; r < FM_MAX_REGS; r++) {
          const int f = lane + 32 * r;
          if (f < d.k) {
            float z = row[f], n = row[ld + f];
            ftrl_apply<PRECISE>(z, n, wv[r], a0[r], a1[r], h);
            row[f] = z;
            row[ld + f] = n;
          }
        }
      }
      if (lane == 0) {
        float4 e = lin[lrow];
        if (!SH) e.z = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);  // (sharded: materialised by the owner kernel)
        ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
        lin[lrow] = e;
      }
    } else {
      if (IS_FM) {
        float *dst = part + (int64_t)ci.slot * 2 * ld;
#pragma unroll
        for (int r = 0; r < FM_MAX_REGS; r++) {
          const int f = lane + 32 * r;
          if (f < d.k) {
            dst[f] = a0[r];
            dst[ld + f] = a1[r];
          }
        }
      }
      if (lane == 0) part_lin[ci.slot] = make_float2(sg, sg2);
    }
  }
}

template <bool PRECISE, bool IS_FM, int WARPS, bool SH = false>
__global__ void __launch_bounds__(WARPS * 32)
k_lrfm_combine(Dims d, Hyper h, int32_t nnz, float *__restrict__ tab, float4 *__restrict__ lin, int32_t ch,
               const int32_t *__restrict__ n_chunks_p, const int32_t *__restrict__ chunk_pos,
               const uint32_t *__restrict__ skey, const SegScan *__restrict__ scan,
               const float *__restrict__ part, const float2 *__restrict__ part_lin,
               const __grid_constant__ Export ex, const int32_t *__restrict__ batch_flags) {
  if (SH && batch_flags[1] != 0) return;  // sharded run: the step was called off
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info<false>(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid || !ci.row_head || ci.row_last) continue;
    int J = 1;
    while (c + J < n_chunks && skey[chunk_pos[c + J]] == ci.key) J++;
    const int32_t dst = SH ? ex.dst_at[ci.p0] : -2;  // (ci.p0 of a head chunk = the row's sorted head position)
    const int q = SH ? (int)(ci.key & (uint32_t)ex.Gm1) : 0;
    const int64_t lrow = SH ? (int64_t)(ci.key >> ex.log2G) : (int64_t)ci.key;
    if (IS_FM) {
      float *row = tab + lrow * rs;
      float *o = dst >= 0 ? ex.inbox[q] + (int64_t)dst * 2 * ld : nullptr;
      const float *p0 = part + (int64_t)ci.slot * 2 * ld;
      for (int f = lane; f < d.k; f += 32) {
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < J; j++) {
          a0 += p0[(int64_t)j * 2 * ld + f];
          a1 += p0[(int64_t)j * 2 * ld + ld + f];
        }
        if (o) {
          o[f] = a0;
          o[ld + f] = a1;
        } else {
          float z = row[f], n = row[ld + f];
          ftrl_apply<PRECISE>(z, n, row[2 * ld + f], a0, a1, h);
          row[f] = z;
          row[ld + f] = n;
        }
      }
    }
    if (lane == 0) {
      float sg = 0.f, sg2 = 0.f;
      for (int j = 0; j < J; j++) {
        const float2 t = part_lin[ci.slot + j];
        sg += t.x; sg2 += t.y;
      }
      if (dst >= 0) {
        ex.inbox_lin[q][dst] = make_float2(sg, sg2);
      } else {
        float4 e = lin[lrow];
        if (!SH) e.z = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
        ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
        lin[lrow] = e;
      }
    }
  }
}

// predict (lr.cpp:20-24, fm.cpp:34-38): warp per sample, stored w only
template <bool IS_FM>
__global__ void __launch_bounds__(256)
k_lrfm_predict(Batch b, Dims d, const __grid_constant__ Shards sh, const float4 *__restrict__ bias, int output_prob,
               float *__restrict__ out, float *__restrict__ logit_out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= b.n_rows) return;
  const int64_t r0 = b.row_ptr[s];
  const int F = (int)(b.row_ptr[s + 1] - r0);
  const int64_t ld = d.ld, rs = 3 * ld;
  float acc = 0.f;
  for (int t = lane; t < F; t += 32) {
    const int32_t ft = b.feat[r0 + t];
    if (ft >= 0 && ft < d.n_feats) acc = fmaf(sh.linp(ft)->z, b.val[r0 + t], acc);
  }
  if (IS_FM) {
    for (int f = lane; f < d.k; f += 32) {
      float sv = 0.f, qv = 0.f;
      for (int t = 0; t < F; t++) {
        const int32_t ft = b.feat[r0 + t];
        if (ft < 0 || ft >= d.n_feats) continue;
        const float vx = sh.row(ft, rs)[2 * ld + f] * b.val[r0 + t];
        sv += vx;
        qv = fmaf(vx, vx, qv);
      }
      acc += 0.5f * (sv * sv - qv);
    }
  }
  float logit = warp_sum(acc);
  if (lane == 0) {
    logit += bias->z;
    out[s] = output_prob ? sigmoid_f(logit) : logit;
    if (logit_out) logit_out[s] = logit;
  }
}

// ---------------------------------------------------------------------------------------------
// batch-level reduction, deterministic: per-CTA partial sums of g, g^2 (bias update,
// ftrl_model.cpp:61-64, 79-85 telescoped) and of the fp64 log-loss (eval/loss.h:8-12,
// ftrl_offline.cpp:72-82,101) computed here from the logits; the last CTA to finish adds the
// partials in CTA order and applies the bias update.
// ---------------------------------------------------------------------------------------------
constexpr int RED_MAX_CTAS = 256;

template <bool PRECISE>
__global__ void __launch_bounds__(256)
k_batch_reduce(int64_t n_rows, Hyper h, const float *__restrict__ g, const float *__restrict__ logit,
               const int32_t *__restrict__ label, float4 *__restrict__ bias, int update_bias,
               double *__restrict__ partials, unsigned int *__restrict__ ticket,
               double *__restrict__ loss_sum_out, double *__restrict__ publish4) {
  __shared__ double sh[3][8];
  __shared__ bool s_last;
  const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * per, i1 = min(n_rows, i0 + per);
  double a = 0.0, q = 0.0, l = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    if (g) {
      a += (double)g[i];
      q += (double)(g[i] * g[i]);
    }
    if (label && logit) l += logloss_d(label[i], logit[i]);
  }
  a = warp_sum_d(a); q = warp_sum_d(q); l = warp_sum_d(l);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sh[0][wid] = a; sh[1][wid] = q; sh[2][wid] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = q = l = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += sh[0][w]; q += sh[1][w]; l += sh[2][w]; }
    partials[3 * blockIdx.x + 0] = a;
    partials[3 * blockIdx.x + 1] = q;
    partials[3 * blockIdx.x + 2] = l;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    a = q = l = 0.0;
    for (unsigned c = 0; c < gridDim.x; c++) {
      a += __ldcg(partials + 3 * c + 0);
      q += __ldcg(partials + 3 * c + 1);
      l += __ldcg(partials + 3 * c + 2);
    }
    if (publish4) {  // sharded run: the bias update happens after the partials of all ranks are exchanged
      publish4[0] = a; publish4[1] = q; publish4[2] = l; publish4[3] = (double)n_rows;
    } else if (update_bias && n_rows > 0) {
      float4 e = *bias;
      e.z = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
      ftrl_apply<PRECISE>(e.x, e.y, e.z, (float)a, (float)q, h);
      *bias = e;
    }
    if (loss_sum_out) *loss_sum_out = l;
    *ticket = 0u;
  }
}

}  // namespace ftrl
