// ffm_tile.cuh -- the FFM fast path for batches whose samples all have distinct fields
// (Criteo-shaped data): "load rows -> compute in shared memory -> store rows".
//
//  k_ffm_tile        persistent CTAs (one per SM), warp-specialised:
//                    metadata warps : prefetch, several samples ahead, each sample's CSR row, the class of
//                                     every occurrence (fused / sorted position), row locators, linear records
//                    loader warp    : hands out sample-sized spans of the shared-memory row ring and issues
//                                     one cp.async.bulk (TMA bulk copy, mbarrier expect_tx) per feature row:
//                                     z and n planes of a fused row, the materialised w plane of a staged row
//                    consumer warps : pass 1 w = W(n,z) for fused rows (ffm.cpp:72-88, stored as the stale
//                                     w the reference keeps), logit (ffm.cpp:57-70), g = sigmoid(logit) - y;
//                                     pass 2 FTRL update in place in shared memory (ffm.cpp:90-136
//                                     telescoped) or the gradient image g_s w_partner x_m x_n
//                    storer warp    : one bulk store per row: updated (z',n') back into the table, or the
//                                     gradient image into the staging buffer at its sorted position
//  k_row_touch / k_row_materialise   owner-side pre-pass: w of the staged rows, only the touched slices
//  k_ffm_staged_rows streaming segmented reduction of the staged gradient images: work item = (chunk of
//                    <= 32 occurrences of one row, 32 float4 vectors); closed-form update, a partial for
//                    k_ffm_combine, or (sharded runs) the row's sum into its owner's inbox
//  (the row kernels read one 16-byte descriptor per chunk, built once per batch by k_chunk_desc, prep.cuh)
//
// HBM traffic per touched coordinate: rows that occur once: 8 B read (z,n) + 12 B written (z',n',w) = the
// algorithmic 20 B; other rows: 12 B once (materialise) + 4 B (w) + 4 B (gradient image) per occurrence,
// then 4 B per occurrence + 16 B per distinct coordinate in the reduce.
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// L2 prefetch of a span the row loader will bulk-copy a few samples later (hides the DRAM latency of the
// 2-stage row ring: metadata runs several samples ahead of the rows)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// position in a ring of `n` slots, advanced without division: slot index and the mbarrier phase parity
struct RingCursor {
  int slot, round;
  __device__ __forceinline__ void init(int start, int n) { slot = start % n; round = start / n; }
  __device__ __forceinline__ void advance(int by, int n) {
    slot += by;
    while (slot >= n) {
      slot -= n;
      round++;
    }
  }
  __device__ __forceinline__ uint32_t parity() const { return (uint32_t)(round & 1); }
  __device__ __forceinline__ uint32_t prev_parity() const { return (uint32_t)((round - 1) & 1); }
};

// ---- tile geometry ------------------------------------------------------------------------------
constexpr int TILE_META_WARPS = 2;   // warps prefetching sample metadata (CSR, occurrence class, linear records)
// + one row-loader warp and one row-storer warp (TMA bulk copies)
__host__ __device__ constexpr int tile_threads(int consumers) { return consumers + 32 * (2 + TILE_META_WARPS); }
constexpr int TILE_MAX_CONSUMERS = 768;
constexpr int TILE_MAX_STAGE = 4;    // samples in flight in the row ring (power of two)
constexpr int TILE_MAX_META = 8;     // metadata slots

struct TileGeom {
  int f_cap;        // rows per stage (= n_fields: samples with distinct fields have at most that many)
  int stride;       // floats a fused row (z and n planes) takes in the row ring: 2*ld + pad, conflict-free columns
  int stride1;      // floats a staged row (w plane only) takes: ld + pad1, same residue as `stride` modulo 32 floats
  int n_stage;      // the row ring holds at least n_stage * f_cap * stride floats (n_stage all-fused samples) ...
  int ring_bytes;   // ... and all the shared memory the CTA can get beyond that (multiple of 128)
  int inflight;     // samples that may share the ring (2..TILE_MAX_STAGE)
  int n_meta;       // metadata slots (> n_stage: metadata runs ahead of the row ring)
  int consumers;    // consumer threads (multiple of 32)
  int dbg;          // experiment switches (FTRL_B200_TILE_DBG): 1 skip w stores, 2 skip row/image stores, 4 local lin only
  size_t smem_bytes;
};

struct RowMeta {   // one 16-byte record per row of a sample
  int32_t fk;      // field * k | (offset of the row inside its sample's span of the row ring, floats) << 16
  float x;         // value
  int32_t pos;     // -1: row finalised here, >= 0: sorted position for the staged gradient image
  int32_t loc;     // row locator (RowSpace): >= 0 local row, < 0: -1 - head position in the remote-row cache
};
// Rows of a sample are kept sorted by class: fused rows take slots 0, 1, ... (hdr[0] of them), staged rows take
// slots f_cap-1, f_cap-2, ... (hdr[3] of them), so that the work items of a sample fall into three homogeneous
// ranges (fused x fused, fused x staged, staged x staged) and a warp rarely mixes classes: the fused path
// (w = W(n,z), FTRL update: ~10x the instructions of the staged path) is then executed only by the warps whose
// items need it instead of by every warp that happens to hold one fused row.
struct SampleMeta {
  RowMeta *row;     // [f_cap]
  float4 *lin;      // [f_cap] {z, n, w, -} of the linear coordinate, prefetched
  int32_t *hdr;     // [0] fused rows, [1] label, [2] floats of the row ring the sample needs, [3] staged rows,
                    // [4] fused x fused items, [5] fused x staged items, [6] items, [7] 1 / fused rows (float bits):
                    // computed once by the metadata warp instead of by every consumer thread
  uint8_t *present; // [n_fields] 1 when some valid row of the sample carries that field
};

__host__ __device__ inline size_t tile_meta_bytes(int f_cap) {
  // RowMeta (16 B) + lin (16 B) per row, + header 32 B, + present[f_cap] rounded to 16
  return (size_t)f_cap * 32 + 32 + 2 * (size_t)((f_cap + 15) / 16) * 16;
}
__host__ __device__ inline size_t tile_stage_bytes(int f_cap, int stride) {
  return (size_t)f_cap * stride * sizeof(float);
}
// the pair table (m | n << 8) as uint16 for f_cap rows
__host__ __device__ inline size_t tile_lut_bytes(int f_cap) {
  return (((size_t)f_cap * (f_cap - 1) / 2 * 2) + 15) / 16 * 16;
}
__host__ __device__ inline size_t tile_smem_bytes(int f_cap, int stride, int n_stage, int n_meta) {
  return tile_stage_bytes(f_cap, stride) * n_stage + tile_meta_bytes(f_cap) * n_meta + tile_lut_bytes(f_cap);
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

// span of a staged row (w plane only): >= ld floats, same residue as the fused-row span modulo 8 16-byte units,
// so consecutive rows keep landing in distinct bank groups whatever the mix of fused and staged rows
__host__ inline int tile_stride1(int ld, int stride) {
  int s1 = ld;
  while (((s1 / 4) % 8) != ((stride / 4) % 8)) s1 += 4;
  return s1;
}

// ---------------------------------------------------------------------------------------------
// k_ffm_tile.  IPT = items (pair, factor chunk) per consumer thread; w of pass 1 is kept in registers
// for pass 2 (IPT * 8 floats).
// ---------------------------------------------------------------------------------------------
template <bool PRECISE>
__device__ __forceinline__ float4 weight4(const float4 &z, const float4 &n, const Hyper &h) {
  return make_float4(weight_from<PRECISE>(z.x, f_sqrt<PRECISE>(n.x), h), weight_from<PRECISE>(z.y, f_sqrt<PRECISE>(n.y), h),
                     weight_from<PRECISE>(z.z, f_sqrt<PRECISE>(n.z), h), weight_from<PRECISE>(z.w, f_sqrt<PRECISE>(n.w), h));
}
template <bool PRECISE>
__device__ __forceinline__ void apply4(float4 &z, float4 &n, const float4 &w, const float4 &wp, float gx, const Hyper &h) {
  float gv;
  gv = gx * wp.x; ftrl_apply<PRECISE>(z.x, n.x, w.x, gv, gv * gv, h);
  gv = gx * wp.y; ftrl_apply<PRECISE>(z.y, n.y, w.y, gv, gv * gv, h);
  gv = gx * wp.z; ftrl_apply<PRECISE>(z.z, n.z, w.z, gv, gv * gv, h);
  gv = gx * wp.w; ftrl_apply<PRECISE>(z.w, n.w, w.w, gv, gv * gv, h);
}

// Thread roles: [0, consumers) compute; warp `consumers/32` is the row producer (bulk loads / bulk
// stores of the row ring); the next TILE_META_WARPS warps prefetch sample metadata into a deeper
// ring so that the row producer never waits on a dependent global-load chain.
template <bool PRECISE, int IPT, bool CACHE>
__global__ void __launch_bounds__(tile_threads(TILE_MAX_CONSUMERS), 1)
k_ffm_tile(Batch b, Dims d, Hyper h, ItemDecode dec, TileGeom geo, const int32_t *__restrict__ batch_flags,
           const __grid_constant__ RowSpace rsp, const float4 *__restrict__ bias, const uint32_t *__restrict__ pair_lut,
           const int32_t *__restrict__ occ_pos, const SegScan *__restrict__ scan, float *__restrict__ g_out,
           float *__restrict__ logit_out) {
  if (batch_flags[0] == 0) return;  // some sample repeats a field: the generic kernels take this batch
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar_full[TILE_MAX_STAGE], bar_done[TILE_MAX_STAGE], bar_free[TILE_MAX_STAGE];
  __shared__ uint64_t bar_mfull[TILE_MAX_META], bar_mfree[TILE_MAX_META];
  __shared__ float s_red[2][32];  // per-warp partial logits, double-buffered by sample parity

  const int tid = threadIdx.x;
  const int n_cons = geo.consumers;
  const int n_cons_warps = n_cons >> 5;
  const int lane = tid & 31;
  // 0 consumer, 1 row loader, 3 row storer, 2 metadata
  constexpr int NL = 1, NSW = 1;
  const int role = tid < n_cons ? 0 : (tid < n_cons + 32 * NL ? 1 : (tid < n_cons + 32 * (NL + NSW) ? 3 : 2));
  const int ld = d.ld, k = d.k;
  const int stride = geo.stride, stride1 = geo.stride1, f_cap = geo.f_cap, MD = geo.n_meta;
  // The row ring: geo.ring_bytes of shared memory (at least two all-fused samples; with one CTA per SM everything
  // the SM has left) handed out in sample-sized spans.  A fused row takes `stride` floats (z, n), a staged row
  // `stride1` (w only), so batches with many duplicated rows keep 3-4 samples in flight where all-fused samples
  // keep 2.  Up to TILE_MAX_STAGE samples share the ring.
  constexpr int NSLOT = TILE_MAX_STAGE;
  const int ring_floats = (int)((size_t)geo.ring_bytes / sizeof(float));
  float *ring = reinterpret_cast<float *>(smem_raw);
  __shared__ int s_base[NSLOT], s_need[NSLOT];
  const size_t meta_bytes = tile_meta_bytes(f_cap);
  const int64_t rs = 3 * (int64_t)ld;
  const uint32_t row_bytes = (uint32_t)(2 * ld * sizeof(float));
  unsigned char *meta_base = smem_raw + (size_t)geo.ring_bytes;
  uint16_t *s_lut = reinterpret_cast<uint16_t *>(meta_base + (size_t)MD * meta_bytes);

  auto row_off = [](const RowMeta &rm) { return rm.fk >> 16; };
  auto row_fk = [](const RowMeta &rm) { return rm.fk & 0xffff; };
  auto sample_meta = [&](int slot) {
    SampleMeta m;
    unsigned char *p = meta_base + (size_t)slot * meta_bytes;
    m.row = reinterpret_cast<RowMeta *>(p);
    p += (size_t)f_cap * 16;
    m.lin = reinterpret_cast<float4 *>(p);
    p += (size_t)f_cap * 16;
    m.hdr = reinterpret_cast<int32_t *>(p);
    p += 32;
    m.present = p;
    return m;
  };

  // r-th row of a sample with nf fused rows: fused rows from the front, staged rows from the back
  auto row_slot = [&](int r, int nf) { return r < nf ? r : f_cap - 1 - (r - nf); };

  if (tid == 0) {
    for (int st = 0; st < NSLOT; st++) {
      mbar_init(&bar_full[st], NL);
      mbar_init(&bar_done[st], n_cons_warps);
      mbar_init(&bar_free[st], NSW);
    }
    for (int sl = 0; sl < MD; sl++) {
      mbar_init(&bar_mfull[sl], 1);
      mbar_init(&bar_mfree[sl], NSW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int p = tid; p < f_cap * (f_cap - 1) / 2; p += blockDim.x) {
    const uint32_t e = pair_lut[p];
    s_lut[p] = (uint16_t)((e & 0xffu) | ((e >> 16) << 8));
  }
  __syncthreads();

  const int n_mine = b.n_rows > blockIdx.x ? (int)((b.n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  if (role == 2) {
    // =========================== metadata warps ===========================
    const int mw = (tid - n_cons - 32 * (NL + NSW)) >> 5;
    // The per-sample chain of dependent global loads is what bounds a metadata warp (and, with two of them, the
    // whole CTA: the consumers were waiting for rows a fifth of the time): row_ptr of the NEXT sample is fetched
    // one iteration ahead, and field / feat / val / occ_pos of a sample are independent loads, so a sample costs
    // two round trips (CSR entries, then the linear records) instead of four.
    RingCursor mc;
    mc.init(mw, MD);
    int64_t rp0 = 0, rp1 = 0;
    if (mw < n_mine) {
      const int64_t s0 = blockIdx.x + (int64_t)mw * gridDim.x;
      rp0 = b.row_ptr[s0];
      rp1 = b.row_ptr[s0 + 1];
    }
    for (int it = mw; it < n_mine; it += TILE_META_WARPS, mc.advance(TILE_META_WARPS, MD)) {
      const int slot = mc.slot;
      const int64_t s = blockIdx.x + (int64_t)it * gridDim.x;
      const int64_t r0 = rp0;
      const int F = (int)min((int64_t)1 << 20, rp1 - r0);
      if (it + TILE_META_WARPS < n_mine) {
        const int64_t sn = s + (int64_t)TILE_META_WARPS * gridDim.x;
        rp0 = b.row_ptr[sn];
        rp1 = b.row_ptr[sn + 1];
      }
      const int32_t label_s = b.label[s];
      if (mc.round > 0) mbar_wait(&bar_mfree[slot], mc.prev_parity());
      SampleMeta m = sample_meta(slot);
      int nf = 0, ns = 0;
      for (int f = lane; f < f_cap; f += 32) m.present[f] = 0;
      __syncwarp();
      for (int base = 0; base < F; base += 64) {
        // two blocks of 32 CSR entries at a time: all their loads are issued before the first use
        int32_t fl[2], ft[2], pos[2];
        float x[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int t = base + 32 * u + lane;
          fl[u] = 0; ft[u] = -1; pos[u] = 0; x[u] = 0.f; ok[u] = false;
          if (t < F) {
            fl[u] = b.field[r0 + t];
            ft[u] = b.feat[r0 + t];
            x[u] = b.val[r0 + t];
            pos[u] = occ_pos[r0 + t];  // (written for every occurrence, valid or not: no need to wait for `ok`)
          }
        }
        int sl[2];
        bool put[2];
        RowMeta rm[2];
        float4 le[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          ok[u] = base + 32 * u + lane < F && feat_valid(d, fl[u], ft[u]);
          const bool fz = ok[u] && pos[u] < 0, sg = ok[u] && pos[u] >= 0;
          const unsigned fm = __ballot_sync(0xffffffffu, fz), sm = __ballot_sync(0xffffffffu, sg);
          const unsigned below = (1u << lane) - 1;
          const int idx = fz ? nf + __popc(fm & below) : ns + __popc(sm & below);
          put[u] = ok[u] && idx < f_cap;
          sl[u] = fz ? idx : f_cap - 1 - idx;
          nf += __popc(fm);
          ns += __popc(sm);
          rm[u].fk = fl[u] * k;
          rm[u].x = x[u];
          rm[u].pos = pos[u];
          rm[u].loc = 0;
          le[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          if (!put[u]) continue;
          if ((ft[u] & rsp.Gm1) == rsp.rank) {
            rm[u].loc = ft[u] >> rsp.log2G;
            le[u] = rsp.lin[rm[u].loc];
          } else {  // remote rows are never fused: pos >= 0, the row's cache slot is its sorted head position
            const int32_t head = scan[pos[u]].start;
            rm[u].loc = -1 - head;
            le[u].z = rsp.rc_lin[head];
          }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          if (!put[u]) continue;
          m.lin[sl[u]] = le[u];
          m.row[sl[u]] = rm[u];
          m.present[fl[u]] = 1;
          if (geo.dbg & 64) {  // experiment: L2 prefetch of the row several samples before the loader copies it
            if (rm[u].pos < 0) bulk_prefetch_l2(rsp.tab + (int64_t)rm[u].loc * rs, row_bytes);
            else if (rm[u].loc >= 0) bulk_prefetch_l2(rsp.tab + (int64_t)rm[u].loc * rs + 2 * ld, row_bytes / 2);
            else bulk_prefetch_l2(rsp.rc_w + (int64_t)(-1 - rm[u].loc) * ld, row_bytes / 2);
          }
        }
      }
      // distinct fields: nf + ns <= f_cap (the clamps only guard malformed input)
      nf = min(nf, f_cap);
      ns = min(ns, f_cap - nf);
      const int nv = nf + ns;
      __syncwarp();
      // offsets of the rows inside the sample's span (exclusive prefix sum of the row sizes)
      int need = 0;
      for (int base = 0; base < nv; base += 32) {
        const int r = base + lane;
        const int sl = row_slot(r, nf);
        RowMeta rm;
        int sz = 0;
        if (r < nv) {
          rm = m.row[sl];
          sz = r < nf ? stride : stride1;
        }
        int inc = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (r < nv) m.row[sl].fk = rm.fk | ((need + inc - sz) << 16);
        need += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) {
        m.hdr[0] = nf;
        m.hdr[1] = label_s;
        m.hdr[2] = need;
        m.hdr[3] = ns;
        m.hdr[4] = nf * (nf - 1) / 2 * (int)dec.C;
        m.hdr[5] = nf * ns * (int)dec.C;
        m.hdr[6] = nv * (nv - 1) / 2 * (int)dec.C;
        m.hdr[7] = __float_as_int(nf > 0 ? 1.0f / (float)nf : 0.f);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_mfull[slot]);
    }
    return;
  }

  // w plane of a staged row: the local table, or the cache of remote rows
  auto w_plane = [&](int32_t loc) -> const float * {
    return loc >= 0 ? rsp.tab + (int64_t)loc * rs + 2 * ld : rsp.rc_w + (int64_t)(-1 - loc) * ld;
  };

  if (role == 1) {
    // =========================== row loader warps ===========================
    RingCursor mc;
    mc.init(0, MD);
    int head = 0;     // next free float of the ring
    int tail_it = 0;  // oldest sample whose span has not been handed back by the storer
    for (int it = 0; it < n_mine; it++, mc.advance(1, MD)) {
      const int st = it & (NSLOT - 1);
      const int slot = mc.slot;
      mbar_wait(&bar_mfull[slot], mc.parity());
      SampleMeta m = sample_meta(slot);
      const int nf = m.hdr[0], nv = nf + m.hdr[3];
      const int need = m.hdr[2];
      // a span for this sample: contiguous, after `head` or wrapped to the start of the ring; wait (in order)
      // for older samples to retire until the sample slot is free and the span overlaps no live span
      while (it - tail_it >= geo.inflight) {
        mbar_wait(&bar_free[tail_it & (NSLOT - 1)], (uint32_t)((tail_it >> 2) & 1));
        tail_it++;
      }
      const int base = head + need <= ring_floats ? head : 0;
      for (;;) {
        bool clash = false;
        for (int j = tail_it; j < it; j++) {
          const int bj = s_base[j & (NSLOT - 1)], nj = s_need[j & (NSLOT - 1)];
          clash = clash || (base < bj + nj && bj < base + need);
        }
        if (!clash) break;
        mbar_wait(&bar_free[tail_it & (NSLOT - 1)], (uint32_t)((tail_it >> 2) & 1));
        tail_it++;
      }
      head = base + need;
      __syncwarp();
      if (lane == 0) {
        s_base[st] = base;
        s_need[st] = need;
      }
      __syncwarp();
      float *rows = ring + base;
      {
        // TMA bulk copies.  Fused rows bring z and n (they are updated here); staged rows bring only the w
        // plane their owner materialised, into the z-plane slot of the stage
        if (lane == 0) mbar_expect_tx(&bar_full[st], (uint32_t)(nf * (int)row_bytes + (nv - nf) * (int)(row_bytes / 2)));
        __syncwarp();
        for (int r = lane; r < nv; r += 32) {
          const RowMeta rm = m.row[row_slot(r, nf)];
          if (rm.pos < 0) bulk_g2s(rows + row_off(rm), rsp.tab + (int64_t)rm.loc * rs, row_bytes, &bar_full[st]);
          else bulk_g2s(rows + row_off(rm), w_plane(rm.loc), row_bytes / 2, &bar_full[st]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[st]);
      // L2 prefetch of the NEXT sample's rows: its stage is still occupied, but its metadata is ready (the
      // metadata ring runs ahead); one sample ahead keeps the prefetched footprint at ~100 KB per SM
      if (it + 1 < n_mine && (geo.dbg & 32)) {  // off by default: with the variable-span ring the loads themselves run ahead
        RingCursor nx = mc;
        nx.advance(1, MD);
        mbar_wait(&bar_mfull[nx.slot], nx.parity());
        SampleMeta m2 = sample_meta(nx.slot);
        const int nf2 = m2.hdr[0], nv2 = nf2 + m2.hdr[3];
        for (int r = lane; r < nv2; r += 32) {
          const RowMeta rm = m2.row[row_slot(r, nf2)];
          if (rm.pos < 0) bulk_prefetch_l2(rsp.tab + (int64_t)rm.loc * rs, row_bytes);
          else bulk_prefetch_l2(w_plane(rm.loc), row_bytes / 2);
        }
      }
    }
    return;
  }

  if (role == 3) {
    // =========================== row storer warps ===========================
    // retire samples in order: wait for the consumers, bulk-store the rows / gradient images, then hand the
    // stage back to the loader and the metadata slot back to the metadata warps
    RingCursor mc;
    mc.init(0, MD);
    for (int it = 0; it < n_mine; it++, mc.advance(1, MD)) {
      const int st = it & (NSLOT - 1);
      const int slot = mc.slot;
      SampleMeta m = sample_meta(slot);
      mbar_wait(&bar_done[st], (uint32_t)((it >> 2) & 1));
      float *rows = ring + s_base[st];
      const int nf = m.hdr[0], nv = nf + m.hdr[3];
      for (int r = lane; r < nv && !(geo.dbg & 2); r += 32) {
        const RowMeta rm = m.row[row_slot(r, nf)];
        if (rm.pos < 0) bulk_s2g(rsp.tab + (int64_t)rm.loc * rs, rows + row_off(rm), row_bytes);
        else bulk_s2g(rsp.staging + (int64_t)rm.pos * ld, rows + row_off(rm), (uint32_t)(ld * sizeof(float)));
      }
      bulk_commit();
      bulk_wait_read_all();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar_free[st]);
        mbar_arrive(&bar_mfree[slot]);
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
  RingCursor mc;
  mc.init(0, MD);
  for (int it = 0; it < n_mine; it++, mc.advance(1, MD)) {
    const int st = it & (NSLOT - 1);
    const int slot = mc.slot;
    const int64_t s = blockIdx.x + (int64_t)it * gridDim.x;
    SampleMeta m = sample_meta(slot);
    mbar_wait(&bar_mfull[slot], mc.parity());
    mbar_wait(&bar_full[st], (uint32_t)((it >> 2) & 1));
    float *rows = ring + s_base[st];
    const int nf = m.hdr[0], ns = m.hdr[3], nv = nf + ns;
    // items of the sample in three ranges: fused x fused pairs, fused x staged, staged x staged
    const uint32_t n_ff = (uint32_t)m.hdr[4], n_fs = (uint32_t)m.hdr[5], n_items = (uint32_t)m.hdr[6];
    const float inv_nf = __int_as_float(m.hdr[7]);
    // (m, n) = row slots of item `item`
    auto item_rows = [&](uint32_t item, int &mi, int &ni, uint32_t &c) {
      uint32_t p;
      if (item < n_ff) {
        dec(item, p, c);
        const uint32_t e = s_lut[p];
        mi = e & 0xff;
        ni = e >> 8;
      } else if (item < n_ff + n_fs) {
        dec(item - n_ff, p, c);
        // p = n' * nf + m : exact for these ranges (p < 64 * 64, nf <= 64)
        const int nq = __float2int_rd(((float)p + 0.5f) * inv_nf);
        mi = (int)p - nq * nf;
        ni = f_cap - 1 - nq;
      } else {
        dec(item - n_ff - n_fs, p, c);
        const uint32_t e = s_lut[p];
        mi = f_cap - 1 - (int)(e & 0xff);
        ni = f_cap - 1 - (int)(e >> 8);
      }
    };

    // ---- pass 1: w, logit ----
    float acc = 0.f;
    float4 wAc[IPT], wBc[IPT];
    // CACHE: pass 2 reuses the item's slice offsets / classes from pass 1 instead of decoding it again
    int offA[CACHE ? IPT : 1], offB[CACHE ? IPT : 1], cls[CACHE ? IPT : 1];  // bit 0/1: row m / n fused, bit 2: valid
    float xx[CACHE ? IPT : 1];
#pragma unroll
    for (int j = 0; j < IPT; j++) {
      const uint32_t item = tid + j * n_cons;
      if (CACHE) cls[j] = 0;
      if (item < n_items) {
        uint32_t c;
        int mi, ni;
        item_rows(item, mi, ni, c);
        const RowMeta rmm = m.row[mi], rmn = m.row[ni];
        const int oA = row_off(rmm) + row_fk(rmn) + (int)c * 4;  // slice A = (row m, field n)
        const int oB = row_off(rmn) + row_fk(rmm) + (int)c * 4;  // slice B = (row n, field m)
        const float xmn = rmm.x * rmn.x;
        if (CACHE) {
          offA[j] = oA;
          offB[j] = oB;
          xx[j] = xmn;
          cls[j] = 4 | (rmm.pos < 0 ? 1 : 0) | (rmn.pos < 0 ? 2 : 0);
        }
        const float *sa = rows + oA;
        const float *sb = rows + oB;
        // staged rows hold w itself in the z-plane slot; fused rows hold (z, n): w = W(n, z), stored as the
        // stale-by-one w the reference keeps (ffm.cpp:72-88)
        float4 wA = *reinterpret_cast<const float4 *>(sa), wB = *reinterpret_cast<const float4 *>(sb);
        if (rmm.pos < 0) {
          wA = weight4<PRECISE>(wA, *reinterpret_cast<const float4 *>(sa + ld), h);
          *reinterpret_cast<float4 *>(rsp.tab + (int64_t)rmm.loc * rs + 2 * ld + row_fk(rmn) + c * 4) = wA;
        }
        if (rmn.pos < 0) {
          wB = weight4<PRECISE>(wB, *reinterpret_cast<const float4 *>(sb + ld), h);
          *reinterpret_cast<float4 *>(rsp.tab + (int64_t)rmn.loc * rs + 2 * ld + row_fk(rmm) + c * 4) = wB;
        }
        wAc[j] = wA;
        wBc[j] = wB;
        const float dot = fmaf(wA.x, wB.x, fmaf(wA.y, wB.y, fmaf(wA.z, wB.z, wA.w * wB.w)));
        acc = fmaf(dot, xmn, acc);
      }
    }
    for (int r = tid; r < nv; r += n_cons) {
      const int sl = row_slot(r, nf);
      const float4 e = m.lin[sl];
      const RowMeta rm = m.row[sl];
      const float w = rm.pos < 0 ? weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h) : e.z;
      acc = fmaf(w, rm.x, acc);
    }
    // consumer-wide sum
    // one barrier per sample: every warp adds the per-warp partials itself (same values, same order -> the
    // same logit in every warp).  The buffer of sample `it` is next written for sample it + 2, i.e. by a warp
    // that has passed the barrier of sample it + 1, which every warp reaches only after this read.
    float *red = s_red[it & 1];
    acc = warp_sum(acc);
    if (lane == 0) red[tid >> 5] = acc;
    named_bar_sync(1, n_cons);
    const float logit = warp_sum(lane < n_cons_warps ? red[lane] : 0.f) + bias_w;
    const float g = sigmoid_f(logit) - (float)m.hdr[1];
    if (tid == 0) {
      g_out[s] = g;
      logit_out[s] = logit;
    }

    // ---- pass 2: FTRL update in place (fused rows) or gradient image into the z plane (staged rows);
    //      every (row, field) slice is read and written by exactly one item (fields are distinct) ----
#pragma unroll
    for (int j = 0; j < IPT; j++) {
      int oA, oB, cl;
      float xmn;
      if (CACHE) {
        oA = offA[j]; oB = offB[j]; cl = cls[j]; xmn = xx[j];
      } else {
        const uint32_t item = tid + j * n_cons;
        cl = 0; oA = oB = 0; xmn = 0.f;
        if (item < n_items) {
          uint32_t c;
          int mi, ni;
          item_rows(item, mi, ni, c);
          const RowMeta rmm = m.row[mi], rmn = m.row[ni];
          oA = row_off(rmm) + row_fk(rmn) + (int)c * 4;
          oB = row_off(rmn) + row_fk(rmm) + (int)c * 4;
          xmn = rmm.x * rmn.x;
          cl = 4 | (rmm.pos < 0 ? 1 : 0) | (rmn.pos < 0 ? 2 : 0);
        }
      }
      if (cl & 4) {
        float *sa = rows + oA;
        float *sb = rows + oB;
        const float gx = g * xmn;
        const float4 wA = wAc[j], wB = wBc[j];
        if (cl & 1) {
          float4 zA = *reinterpret_cast<const float4 *>(sa), nA = *reinterpret_cast<const float4 *>(sa + ld);
          apply4<PRECISE>(zA, nA, wA, wB, gx, h);
          *reinterpret_cast<float4 *>(sa) = zA;
          *reinterpret_cast<float4 *>(sa + ld) = nA;
        } else {
          *reinterpret_cast<float4 *>(sa) = make_float4(gx * wB.x, gx * wB.y, gx * wB.z, gx * wB.w);
        }
        if (cl & 2) {
          float4 zB = *reinterpret_cast<const float4 *>(sb), nB = *reinterpret_cast<const float4 *>(sb + ld);
          apply4<PRECISE>(zB, nB, wB, wA, gx, h);
          *reinterpret_cast<float4 *>(sb) = zB;
          *reinterpret_cast<float4 *>(sb + ld) = nB;
        } else {
          *reinterpret_cast<float4 *>(sb) = make_float4(gx * wA.x, gx * wA.y, gx * wA.z, gx * wA.w);
        }
      }
    }
    // linear coordinate: fused -> full update; staged -> w now, gradient to staging_lin
    for (int r = tid; r < nv; r += n_cons) {
      const int sl = row_slot(r, nf);
      float4 e = m.lin[sl];
      const RowMeta rm = m.row[sl];
      const float gi = g * rm.x;
      if (rm.pos < 0) {
        const float w = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
        e.z = w;
        ftrl_apply<PRECISE>(e.x, e.y, w, gi, gi * gi, h);
        rsp.lin[rm.loc] = e;
      } else {
        rsp.staging_lin[rm.pos] = gi;  // w of staged rows: materialised by their owner
      }
    }
    // staged rows: slices no partner touches (own field, absent fields) must read as 0 in the image.
    // They are disjoint from the slices written above, so no barrier is needed.
    if (nv == d.n_fields) {
      // every field is present (fields are distinct): only the own-field slice is untouched
      for (int r = nf + tid; r < nv; r += n_cons) {   // staged rows
        const RowMeta rm = m.row[row_slot(r, nf)];
        float4 *zp = reinterpret_cast<float4 *>(rows + row_off(rm) + row_fk(rm));
        for (int v = 0; v < (k >> 2); v++) zp[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      for (int r = nf + (tid >> 5); r < nv; r += n_cons_warps) {   // staged rows
        const RowMeta rm = m.row[row_slot(r, nf)];
        for (int f = lane; f < d.n_fields; f += 32) {
          if (m.present[f] && f * k != row_fk(rm)) continue;
          float4 *zp = reinterpret_cast<float4 *>(rows + row_off(rm) + f * k);
          for (int v = 0; v < (k >> 2); v++) zp[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_done[st]);
  }
}

// ---------------------------------------------------------------------------------------------
// owner-side pre-pass for the segmented ("staged") rows: materialise w = W(n,z) (ffm.cpp:72-88,
// ftrl_model.cpp:52-59) for exactly the slices the batch touches, BEFORE the sample kernels run, so
// that those only need the row's w plane (4 B per coordinate instead of z,n = 8 B) and never store w.
//   k_row_touch       : rowmask[row] = OR over the row's occurrences of "fields of the other features
//                       of that sample" (chunk-parallel, atomicOr)
//   k_row_materialise : one warp per row head: w for the slices in rowmask, and the linear w
// ---------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_row_touch(uint32_t sentinel, int32_t ch, const int32_t *__restrict__ batch_flags, const int32_t *__restrict__ n_chunks_p,
            const int4 *__restrict__ cdesc, const uint32_t *__restrict__ socc, PmaskSrc pm,
            unsigned long long *__restrict__ rowmask) {
  if (batch_flags[0] == 0) return;
  const int wib = threadIdx.x >> 5;
  const int n_chunks = *n_chunks_p;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_unpack(cdesc[c], sentinel);
    if (!ci.valid) continue;
    row_touch_chunk(c, ci, ch, socc, pm, rowmask);
  }
}

template <bool PRECISE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_row_materialise(Dims d, Hyper h, uint32_t sentinel, const int32_t *__restrict__ batch_flags,
                  const int32_t *__restrict__ n_chunks_p, const int4 *__restrict__ cdesc,
                  const unsigned long long *__restrict__ rowmask, float *__restrict__ tab, float4 *__restrict__ lin) {
  if (batch_flags[0] == 0) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const int vpf = d.k >> 2;  // float4 vectors per field slice
  const int nv = d.n_fields * vpf;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_unpack(cdesc[c], sentinel);
    if (!ci.valid || !ci.row_head) continue;
    const unsigned long long mask = rowmask[c];
    float *row = tab + (int64_t)ci.key * rs;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane == 0) e = lin[ci.key];
    for (int v0 = 0; v0 < nv; v0 += 128) {
      // all loads of up to four vectors per lane first, then the arithmetic and the stores
      float4 z[4], n[4];
      bool act[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v = v0 + u * 32 + lane;
        act[u] = v < nv && ((mask >> (v / vpf)) & 1ull);
        if (act[u]) {
          z[u] = reinterpret_cast<const float4 *>(row)[v];
          n[u] = reinterpret_cast<const float4 *>(row + ld)[v];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (act[u]) reinterpret_cast<float4 *>(row + 2 * ld)[v0 + u * 32 + lane] = weight4<PRECISE>(z[u], n[u], h);
    }
    if (lane == 0) lin[ci.key].z = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
  }
}

// ---------------------------------------------------------------------------------------------
// k_ffm_staged_rows: streaming segmented reduction over the staged gradient images.
// Work item = (chunk of <= 32 occurrences of one row, part of 32 float4 vectors of the row): one warp,
// one 512-byte coalesced load per occurrence, all loads of the chunk independent (unrolled by 8), so a
// row of 78 vectors is reduced by three warps in parallel.  Part 0 also reduces the linear coordinate.
// ---------------------------------------------------------------------------------------------
template <bool PRECISE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_ffm_staged_rows(Dims d, Hyper h, const int32_t *__restrict__ batch_flags, float *__restrict__ tab,
                  float4 *__restrict__ lin, const int32_t *__restrict__ n_chunks_p, const int4 *__restrict__ cdesc,
                  const float *__restrict__ staging, const float *__restrict__ staging_lin, float *__restrict__ part,
                  float2 *__restrict__ part_lin, const __grid_constant__ Export ex) {
  if (batch_flags[0] == 0) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  const int nvec = (int)(ld >> 2);
  const int parts = (nvec + 31) >> 5;
  const int64_t n_items = (int64_t)n_chunks * parts;
  for (int64_t item = (int64_t)blockIdx.x * WARPS + wib; item < n_items; item += (int64_t)gridDim.x * WARPS) {
    const int c = (int)(item / parts), part_i = (int)(item - (int64_t)c * parts);
    const ChunkInfo ci = chunk_unpack(cdesc[c], sentinel);
    if (!ci.valid) continue;
    const bool whole_row = ci.row_head && ci.row_last;
    const int v = part_i * 32 + lane;
    const bool on = v < nvec;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const float4 *src = reinterpret_cast<const float4 *>(staging + (int64_t)ci.p0 * ld) + (on ? v : 0);
    const int n_occ = ci.p1 - ci.p0;
    float *row = tab + (int64_t)(ci.key >> ex.log2G) * rs;
    int p = 0;
    for (; p + 8 <= n_occ; p += 8) {
      float4 gq[8];
#pragma unroll
      for (int u = 0; u < 8; u++) gq[u] = on ? __ldcs(src + (int64_t)(p + u) * nvec) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        a0.x += gq[u].x; a0.y += gq[u].y; a0.z += gq[u].z; a0.w += gq[u].w;
        a1.x = fmaf(gq[u].x, gq[u].x, a1.x); a1.y = fmaf(gq[u].y, gq[u].y, a1.y);
        a1.z = fmaf(gq[u].z, gq[u].z, a1.z); a1.w = fmaf(gq[u].w, gq[u].w, a1.w);
      }
    }
    for (; p < n_occ; p++) {
      const float4 gt = on ? __ldcs(src + (int64_t)p * nvec) : make_float4(0.f, 0.f, 0.f, 0.f);
      a0.x += gt.x; a0.y += gt.y; a0.z += gt.z; a0.w += gt.w;
      a1.x = fmaf(gt.x, gt.x, a1.x); a1.y = fmaf(gt.y, gt.y, a1.y);
      a1.z = fmaf(gt.z, gt.z, a1.z); a1.w = fmaf(gt.w, gt.w, a1.w);
    }
    // sharded runs: the sum goes to the row's owner unless this rank owns the row and is its only contributor
    const int32_t dst = (ex.on && whole_row) ? ex.dst_at[ci.p0] : -2;
    const int64_t lrow = (int64_t)(ci.key >> ex.log2G);
    if (on) {
      if (whole_row && dst >= 0) {
        // one occurrence: sum g^2 = g^2, the owner squares it (half the bytes over NVLink)
        float *o = ex.inbox[ci.key & ex.Gm1] + (int64_t)dst * 2 * ld;
        reinterpret_cast<float4 *>(o)[v] = a0;
        if (n_occ > 1) reinterpret_cast<float4 *>(o + ld)[v] = a1;
      } else if (whole_row) {
        const bool any = a1.x != 0.f || a1.y != 0.f || a1.z != 0.f || a1.w != 0.f || a0.x != 0.f || a0.y != 0.f ||
                         a0.z != 0.f || a0.w != 0.f;
        if (any) {
          float4 z = reinterpret_cast<float4 *>(row)[v], n = reinterpret_cast<float4 *>(row + ld)[v];
          const float4 w = reinterpret_cast<float4 *>(row + 2 * ld)[v];
          ftrl_apply<PRECISE>(z.x, n.x, w.x, a0.x, a1.x, h);
          ftrl_apply<PRECISE>(z.y, n.y, w.y, a0.y, a1.y, h);
          ftrl_apply<PRECISE>(z.z, n.z, w.z, a0.z, a1.z, h);
          ftrl_apply<PRECISE>(z.w, n.w, w.w, a0.w, a1.w, h);
          reinterpret_cast<float4 *>(row)[v] = z;
          reinterpret_cast<float4 *>(row + ld)[v] = n;
        }
      } else {
        float *pdst = part + (int64_t)ci.slot * 2 * ld;
        reinterpret_cast<float4 *>(pdst)[v] = a0;
        reinterpret_cast<float4 *>(pdst + ld)[v] = a1;
      }
    }
    if (part_i == 0) {
      // linear coordinate
      float sg = 0.f, sg2 = 0.f;
      for (int q = ci.p0 + lane; q < ci.p1; q += 32) {
        const float gi = staging_lin[q];
        sg += gi;
        sg2 = fmaf(gi, gi, sg2);
      }
      sg = warp_sum(sg);
      sg2 = warp_sum(sg2);
      if (lane == 0) {
        if (whole_row && dst >= 0) {
          ex.inbox_lin[ci.key & ex.Gm1][dst] = make_float2(sg, sg2);
        } else if (whole_row) {
          float4 e = lin[lrow];
          ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
          lin[lrow] = e;
        } else {
          part_lin[ci.slot] = make_float2(sg, sg2);
        }
      }
    }
  }
}

}  // namespace ftrl
