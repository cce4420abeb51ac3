// ffm_tile.cuh -- the FFM fast path for batches whose samples all have distinct fields
// (Criteo-shaped data): "load rows -> compute in shared memory -> store rows".
//
//  k_ffm_tile        persistent CTAs, one producer warp + consumer warps, NSTAGE-deep ring.
//                    producer : per sample, one cp.async.bulk (TMA bulk copy, 2*ld*4 bytes) per
//                               feature row brings the row's z and n planes into shared memory,
//                               completion on an mbarrier; after the consumers are done, one bulk
//                               store per row writes either the updated (z',n') row back into the
//                               table (row occurs once in the batch) or the row's per-occurrence
//                               gradient image into the staging buffer at its sorted position.
//                    consumers: pass 1 materialises w = W(n,z) for every slice the sample touches
//                               (ffm.cpp:72-88), stores w, forms the logit (ffm.cpp:57-70) and
//                               g = sigmoid(logit) - y; pass 2 applies the FTRL update in place in
//                               shared memory (ffm.cpp:90-136 telescoped) or deposits
//                               g_s w_partner x_m x_n.
//  k_ffm_staged_rows one warp per chunk of <= 32 occurrences of one row: streams the staged
//                    gradient images (coalesced 128-bit loads), accumulates sum g and sum g^2 in
//                    registers, applies the closed form or parks a partial for k_ffm_combine.
//
// HBM traffic per touched coordinate: rows that occur once: 8 B read (z,n) + 12 B written
// (z',n',w) = the algorithmic 20 B; other rows: 8 B read + 4 B (w) + 4 B (gradient image) per
// occurrence, then 4 B per occurrence + 20 B per distinct coordinate in the reduce.
#pragma once
#include "common.cuh"
#include "ffm.cuh"
#include "prep.cuh"

namespace ftrl {

// ---- PTX wrappers (sm_90+ bulk-copy / mbarrier) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared, completion (bytes) signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// ---- tile geometry ------------------------------------------------------------------------------
struct TileGeom {
  int f_cap;        // rows per stage (= n_fields: samples with distinct fields have at most that many)
  int stride;       // floats between consecutive rows of a stage (2*ld + pad), conflict-free columns
  int n_stage;
  int consumers;    // consumer threads (multiple of 32)
  size_t smem_bytes;
};

struct StageMeta {  // per stage, per row slot (arrays of f_cap entries each, laid out by tile_smem_layout)
  int32_t *feat;    // feature id
  int32_t *fk;      // field * k
  float *x;         // value
  int32_t *pos;     // -1: row finalised here, >= 0: sorted position for the staged gradient image
  float4 *lin;      // {z, n, w, -} of the linear coordinate, prefetched by the producer
  uint8_t *present; // [n_fields] 1 when some valid row of the sample carries that field
};

__host__ __device__ inline size_t tile_meta_bytes(int f_cap) {
  // feat, fk, x, pos (4 B each) + lin (16 B) per row, + header (n valid, label) 16 B, + present[f_cap]
  return (size_t)f_cap * (4 * 4 + 16) + 16 + (size_t)((f_cap + 15) / 16) * 16;
}
__host__ __device__ inline size_t tile_stage_bytes(int f_cap, int stride) {
  return (size_t)f_cap * stride * sizeof(float) + ((tile_meta_bytes(f_cap) + 15) / 16) * 16;
}

// choose stride = 2*ld + pad (floats) such that column accesses of consecutive rows by the lanes of
// one 128-bit shared-memory phase (8 lanes) fall into distinct banks
__host__ inline int tile_stride(int ld, int k) {
  const int C = k >= 4 ? k / 4 : 1;
  const int lanes_c = C >= 8 ? 8 : C;          // lanes of a phase covering one row
  for (int pad = 0; pad <= 64; pad += 4) {
    const int s4 = (2 * ld + pad) / 4;          // stride in 16-byte units
    bool ok = true;
    const int rows = 8 / lanes_c;
    // rows r = 0..rows-1 must land in distinct groups of lanes_c units modulo 8
    unsigned seen = 0;
    for (int r = 0; r < rows && ok; r++) {
      const int g = ((r * s4) % 8);
      if (g % lanes_c != 0) ok = false;
      const unsigned bit = 1u << (g / lanes_c);
      if (seen & bit) ok = false;
      seen |= bit;
    }
    if (ok) return 2 * ld + pad;
  }
  return 2 * ld;
}

// ---------------------------------------------------------------------------------------------
// k_ffm_tile
// ---------------------------------------------------------------------------------------------
template <bool PRECISE>
__global__ void __launch_bounds__(576, 1)
k_ffm_tile(Batch b, Dims d, Hyper h, ItemDecode dec, TileGeom geo, const int32_t *__restrict__ batch_flags, float *__restrict__ tab,
           float4 *__restrict__ lin, const float4 *__restrict__ bias, const uint32_t *__restrict__ pair_lut,
           const int32_t *__restrict__ occ_pos, float *__restrict__ staging, float *__restrict__ staging_lin,
           float *__restrict__ g_out, float *__restrict__ logit_out) {
  if (batch_flags[0] == 0) return;  // some sample repeats a field: the generic kernels take this batch
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int MAX_STAGE = 8;
  __shared__ uint64_t bar_full[MAX_STAGE], bar_done[MAX_STAGE];
  __shared__ float s_red[40];

  const int tid = threadIdx.x;
  const int n_cons = geo.consumers;
  const int n_cons_warps = n_cons >> 5;
  const bool is_producer = tid >= n_cons;  // last warp
  const int lane = tid & 31;
  const int ld = d.ld, k = d.k;
  const int stride = geo.stride, f_cap = geo.f_cap, NS = geo.n_stage;
  const size_t stage_bytes = tile_stage_bytes(f_cap, stride);
  const int64_t rs = 3 * (int64_t)ld;
  const uint32_t row_bytes = (uint32_t)(2 * ld * sizeof(float));

  auto stage_rows = [&](int st) -> float * { return reinterpret_cast<float *>(smem_raw + (size_t)st * stage_bytes); };
  auto stage_meta = [&](int st, StageMeta &m, int32_t *&hdr) {
    unsigned char *p = smem_raw + (size_t)st * stage_bytes + (size_t)f_cap * stride * sizeof(float);
    m.lin = reinterpret_cast<float4 *>(p);
    p += (size_t)f_cap * 16;
    m.feat = reinterpret_cast<int32_t *>(p);
    p += (size_t)f_cap * 4;
    m.fk = reinterpret_cast<int32_t *>(p);
    p += (size_t)f_cap * 4;
    m.x = reinterpret_cast<float *>(p);
    p += (size_t)f_cap * 4;
    m.pos = reinterpret_cast<int32_t *>(p);
    p += (size_t)f_cap * 4;
    hdr = reinterpret_cast<int32_t *>(p);  // [0] n valid rows, [1] label
    p += 16;
    m.present = p;
  };

  if (tid == 0) {
    for (int st = 0; st < NS; st++) {
      mbar_init(&bar_full[st], 1);
      mbar_init(&bar_done[st], n_cons_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_mine = b.n_rows > blockIdx.x ? (b.n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (is_producer) {
    // =========================== producer warp ===========================
    for (int64_t it = 0; it < n_mine + NS; it++) {
      const int st = (int)(it % NS);
      float *rows = stage_rows(st);
      StageMeta m;
      int32_t *hdr;
      stage_meta(st, m, hdr);
      if (it >= NS) {
        // retire the sample that used this stage: wait for the consumers, then store its rows
        mbar_wait(&bar_done[st], (uint32_t)(((it / NS) - 1) & 1));
        const int nv = hdr[0];
        for (int r = lane; r < nv; r += 32) {
          const int32_t pos = m.pos[r];
          if (pos < 0) {
            bulk_s2g(tab + (int64_t)m.feat[r] * rs, rows + (size_t)r * stride, row_bytes);
          } else {
            bulk_s2g(staging + (int64_t)pos * ld, rows + (size_t)r * stride, (uint32_t)(ld * sizeof(float)));
          }
        }
        bulk_commit();
        bulk_wait_read_all();
        __syncwarp();
      }
      if (it < n_mine) {
        const int64_t s = blockIdx.x + it * gridDim.x;
        const int64_t r0 = b.row_ptr[s];
        const int F = (int)min((int64_t)1 << 20, b.row_ptr[s + 1] - r0);
        int nv = 0;
        for (int f = lane; f < f_cap; f += 32) m.present[f] = 0;
        __syncwarp();
        for (int base = 0; base < F; base += 32) {
          const int t = base + lane;
          int32_t fl = 0, ft = -1;
          float x = 0.f;
          bool ok = false;
          if (t < F) {
            fl = b.field[r0 + t];
            ft = b.feat[r0 + t];
            x = b.val[r0 + t];
            ok = feat_valid(d, fl, ft);
          }
          const unsigned okm = __ballot_sync(0xffffffffu, ok);
          const int slot = nv + __popc(okm & ((1u << lane) - 1));
          if (ok && slot < f_cap) {
            m.feat[slot] = ft;
            m.fk[slot] = fl * k;
            m.x[slot] = x;
            m.pos[slot] = occ_pos[r0 + t];
            m.lin[slot] = lin[ft];
            m.present[fl] = 1;
          }
          nv += __popc(okm);
        }
        nv = min(nv, f_cap);
        if (lane == 0) {
          hdr[0] = nv;
          hdr[1] = b.label[s];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&bar_full[st], (uint32_t)nv * row_bytes);
        __syncwarp();
        for (int r = lane; r < nv; r += 32)
          bulk_g2s(rows + (size_t)r * stride, tab + (int64_t)m.feat[r] * rs, row_bytes, &bar_full[st]);
      }
    }
    bulk_wait_all();
    return;
  }

  // =========================== consumer warps ===========================
  const float bias_w = [&] {
    const float4 bz = *bias;
    return weight_from<PRECISE>(bz.x, f_sqrt<PRECISE>(bz.y), h);
  }();
  for (int64_t it = 0; it < n_mine; it++) {
    const int st = (int)(it % NS);
    const int64_t s = blockIdx.x + it * gridDim.x;
    float *rows = stage_rows(st);
    StageMeta m;
    int32_t *hdr;
    stage_meta(st, m, hdr);
    mbar_wait(&bar_full[st], (uint32_t)((it / NS) & 1));
    const int nv = hdr[0];
    const uint32_t n_items = (uint32_t)nv * (uint32_t)(nv - 1) / 2u * dec.C;

    // ---- pass 1: w, logit ----
    float acc = 0.f;
    for (uint32_t item = tid; item < n_items; item += n_cons) {
      uint32_t p, c;
      dec(item, p, c);
      int mi, ni;
      pair_decode(p, pair_lut, mi, ni);
      const int fkm = m.fk[mi], fkn = m.fk[ni];
      float *sa = rows + (size_t)mi * stride + fkn + c * 4;  // slice A = (row m, field n)
      float *sb = rows + (size_t)ni * stride + fkm + c * 4;  // slice B = (row n, field m)
      const float4 zA = *reinterpret_cast<const float4 *>(sa), nA = *reinterpret_cast<const float4 *>(sa + ld);
      const float4 zB = *reinterpret_cast<const float4 *>(sb), nB = *reinterpret_cast<const float4 *>(sb + ld);
      float4 wA, wB;
      wA.x = weight_from<PRECISE>(zA.x, f_sqrt<PRECISE>(nA.x), h);
      wA.y = weight_from<PRECISE>(zA.y, f_sqrt<PRECISE>(nA.y), h);
      wA.z = weight_from<PRECISE>(zA.z, f_sqrt<PRECISE>(nA.z), h);
      wA.w = weight_from<PRECISE>(zA.w, f_sqrt<PRECISE>(nA.w), h);
      wB.x = weight_from<PRECISE>(zB.x, f_sqrt<PRECISE>(nB.x), h);
      wB.y = weight_from<PRECISE>(zB.y, f_sqrt<PRECISE>(nB.y), h);
      wB.z = weight_from<PRECISE>(zB.z, f_sqrt<PRECISE>(nB.z), h);
      wB.w = weight_from<PRECISE>(zB.w, f_sqrt<PRECISE>(nB.w), h);
      const float dot = fmaf(wA.x, wB.x, fmaf(wA.y, wB.y, fmaf(wA.z, wB.z, wA.w * wB.w)));
      acc = fmaf(dot, m.x[mi] * m.x[ni], acc);
      // the stale-by-one w the reference keeps (ffm.cpp:72-88)
      *reinterpret_cast<float4 *>(tab + (int64_t)m.feat[mi] * rs + 2 * ld + fkn + c * 4) = wA;
      *reinterpret_cast<float4 *>(tab + (int64_t)m.feat[ni] * rs + 2 * ld + fkm + c * 4) = wB;
    }
    float lin_w = 0.f;
    if (tid < nv) {
      const float4 e = m.lin[tid];
      lin_w = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
      acc = fmaf(lin_w, m.x[tid], acc);
    }
    for (int r = tid + n_cons; r < nv; r += n_cons) {  // f_cap > consumers (not the usual case)
      const float4 e = m.lin[r];
      acc = fmaf(weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h), m.x[r], acc);
    }
    // consumer-wide sum
    acc = warp_sum(acc);
    if (lane == 0) s_red[tid >> 5] = acc;
    named_bar_sync(1, n_cons);
    if (tid < 32) {
      float t = lane < n_cons_warps ? s_red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) {
        const float logit = t + bias_w;
        const float g = sigmoid_f(logit) - (float)hdr[1];
        s_red[32] = g;
        g_out[s] = g;
        logit_out[s] = logit;
      }
    }
    named_bar_sync(1, n_cons);
    const float g = s_red[32];

    // ---- pass 2: FTRL update in place (fused rows) or gradient image into the z plane (staged rows);
    //      every (row, field) slice is read and written by exactly one item (fields are distinct) ----
    for (uint32_t item = tid; item < n_items; item += n_cons) {
      uint32_t p, c;
      dec(item, p, c);
      int mi, ni;
      pair_decode(p, pair_lut, mi, ni);
      const int fkm = m.fk[mi], fkn = m.fk[ni];
      float *sa = rows + (size_t)mi * stride + fkn + c * 4;
      float *sb = rows + (size_t)ni * stride + fkm + c * 4;
      float4 zA = *reinterpret_cast<const float4 *>(sa), nA = *reinterpret_cast<const float4 *>(sa + ld);
      float4 zB = *reinterpret_cast<const float4 *>(sb), nB = *reinterpret_cast<const float4 *>(sb + ld);
      const float gx = g * (m.x[mi] * m.x[ni]);
      float sqA[4] = {f_sqrt<PRECISE>(nA.x), f_sqrt<PRECISE>(nA.y), f_sqrt<PRECISE>(nA.z), f_sqrt<PRECISE>(nA.w)};
      float sqB[4] = {f_sqrt<PRECISE>(nB.x), f_sqrt<PRECISE>(nB.y), f_sqrt<PRECISE>(nB.z), f_sqrt<PRECISE>(nB.w)};
      float wA[4] = {weight_from<PRECISE>(zA.x, sqA[0], h), weight_from<PRECISE>(zA.y, sqA[1], h),
                     weight_from<PRECISE>(zA.z, sqA[2], h), weight_from<PRECISE>(zA.w, sqA[3], h)};
      float wB[4] = {weight_from<PRECISE>(zB.x, sqB[0], h), weight_from<PRECISE>(zB.y, sqB[1], h),
                     weight_from<PRECISE>(zB.z, sqB[2], h), weight_from<PRECISE>(zB.w, sqB[3], h)};
      if (m.pos[mi] < 0) {
        float gv;
        gv = gx * wB[0]; ftrl_apply_sq<PRECISE>(zA.x, nA.x, sqA[0], wA[0], gv, gv * gv, h);
        gv = gx * wB[1]; ftrl_apply_sq<PRECISE>(zA.y, nA.y, sqA[1], wA[1], gv, gv * gv, h);
        gv = gx * wB[2]; ftrl_apply_sq<PRECISE>(zA.z, nA.z, sqA[2], wA[2], gv, gv * gv, h);
        gv = gx * wB[3]; ftrl_apply_sq<PRECISE>(zA.w, nA.w, sqA[3], wA[3], gv, gv * gv, h);
        *reinterpret_cast<float4 *>(sa) = zA;
        *reinterpret_cast<float4 *>(sa + ld) = nA;
      } else {
        *reinterpret_cast<float4 *>(sa) = make_float4(gx * wB[0], gx * wB[1], gx * wB[2], gx * wB[3]);
      }
      if (m.pos[ni] < 0) {
        float gv;
        gv = gx * wA[0]; ftrl_apply_sq<PRECISE>(zB.x, nB.x, sqB[0], wB[0], gv, gv * gv, h);
        gv = gx * wA[1]; ftrl_apply_sq<PRECISE>(zB.y, nB.y, sqB[1], wB[1], gv, gv * gv, h);
        gv = gx * wA[2]; ftrl_apply_sq<PRECISE>(zB.z, nB.z, sqB[2], wB[2], gv, gv * gv, h);
        gv = gx * wA[3]; ftrl_apply_sq<PRECISE>(zB.w, nB.w, sqB[3], wB[3], gv, gv * gv, h);
        *reinterpret_cast<float4 *>(sb) = zB;
        *reinterpret_cast<float4 *>(sb + ld) = nB;
      } else {
        *reinterpret_cast<float4 *>(sb) = make_float4(gx * wA[0], gx * wA[1], gx * wA[2], gx * wA[3]);
      }
    }
    // linear coordinate: fused -> full update; staged -> w now, gradient to staging_lin
    for (int r = tid; r < nv; r += n_cons) {
      float4 e = m.lin[r];
      const float w = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
      const float gi = g * m.x[r];
      const int32_t pos = m.pos[r];
      if (pos < 0) {
        e.z = w;
        ftrl_apply<PRECISE>(e.x, e.y, w, gi, gi * gi, h);
        lin[m.feat[r]] = e;
      } else {
        lin[m.feat[r]].z = w;
        staging_lin[pos] = gi;
      }
    }
    // staged rows: slices no partner touches (own field, absent fields) must read as 0 in the image.
    // They are disjoint from the slices written above, so no barrier is needed.
    for (int q = tid; q < nv * d.n_fields; q += n_cons) {
      const int r = q / d.n_fields, f = q - r * d.n_fields;
      if (m.pos[r] < 0) continue;
      if (m.present[f] && f * k != m.fk[r]) continue;
      float4 *zp = reinterpret_cast<float4 *>(rows + (size_t)r * stride + f * k);
      for (int v = 0; v < (k >> 2); v++) zp[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_done[st]);
  }
}

// ---------------------------------------------------------------------------------------------
// k_ffm_staged_rows: streaming segmented reduction over the staged gradient images
// ---------------------------------------------------------------------------------------------
template <bool PRECISE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_ffm_staged_rows(Dims d, Hyper h, int32_t nnz, const int32_t *__restrict__ batch_flags, float *__restrict__ tab,
                  float4 *__restrict__ lin, int32_t ch, const int32_t *__restrict__ n_chunks_p,
                  const int32_t *__restrict__ chunk_pos, const uint32_t *__restrict__ skey,
                  const SegScan *__restrict__ scan, const float *__restrict__ staging,
                  const float *__restrict__ staging_lin, float *__restrict__ part, float2 *__restrict__ part_lin) {
  if (batch_flags[0] == 0) return;
  constexpr int R = 4;  // float4 accumulators per lane and pass: covers 512 floats of the row per pass
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  const int nvec = (int)(ld >> 2);
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info<true>(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid) continue;
    const bool whole_row = ci.row_head && ci.row_last;
    // linear coordinate
    float sg = 0.f, sg2 = 0.f;
    for (int p = ci.p0 + lane; p < ci.p1; p += 32) {
      const float gi = staging_lin[p];
      sg += gi;
      sg2 = fmaf(gi, gi, sg2);
    }
    sg = warp_sum(sg);
    sg2 = warp_sum(sg2);
    float *row = tab + (int64_t)ci.key * rs;
    float *pdst = part + (int64_t)ci.slot * 2 * ld;
    for (int vb = 0; vb < nvec; vb += 32 * R) {
      float4 a0[R], a1[R];
#pragma unroll
      for (int r = 0; r < R; r++) a0[r] = a1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int p = ci.p0; p < ci.p1; p++) {
        const float4 *src = reinterpret_cast<const float4 *>(staging + (int64_t)p * ld);
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int v = vb + r * 32 + lane;
          if (v < nvec) {
            const float4 gq = __ldcs(src + v);
            a0[r].x += gq.x; a0[r].y += gq.y; a0[r].z += gq.z; a0[r].w += gq.w;
            a1[r].x = fmaf(gq.x, gq.x, a1[r].x); a1[r].y = fmaf(gq.y, gq.y, a1[r].y);
            a1[r].z = fmaf(gq.z, gq.z, a1[r].z); a1[r].w = fmaf(gq.w, gq.w, a1[r].w);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int v = vb + r * 32 + lane;
        if (v >= nvec) continue;
        if (whole_row) {
          const bool any = a1[r].x != 0.f || a1[r].y != 0.f || a1[r].z != 0.f || a1[r].w != 0.f ||
                           a0[r].x != 0.f || a0[r].y != 0.f || a0[r].z != 0.f || a0[r].w != 0.f;
          if (!any) continue;
          float4 z = reinterpret_cast<float4 *>(row)[v], n = reinterpret_cast<float4 *>(row + ld)[v];
          const float4 w = reinterpret_cast<float4 *>(row + 2 * ld)[v];
          ftrl_apply<PRECISE>(z.x, n.x, w.x, a0[r].x, a1[r].x, h);
          ftrl_apply<PRECISE>(z.y, n.y, w.y, a0[r].y, a1[r].y, h);
          ftrl_apply<PRECISE>(z.z, n.z, w.z, a0[r].z, a1[r].z, h);
          ftrl_apply<PRECISE>(z.w, n.w, w.w, a0[r].w, a1[r].w, h);
          reinterpret_cast<float4 *>(row)[v] = z;
          reinterpret_cast<float4 *>(row + ld)[v] = n;
        } else {
          reinterpret_cast<float4 *>(pdst)[v] = a0[r];
          reinterpret_cast<float4 *>(pdst + ld)[v] = a1[r];
        }
      }
    }
    if (lane == 0) {
      if (whole_row) {
        float4 e = lin[ci.key];
        ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
        lin[ci.key] = e;
      } else {
        part_lin[ci.slot] = make_float2(sg, sg2);
      }
    }
  }
}

}  // namespace ftrl
