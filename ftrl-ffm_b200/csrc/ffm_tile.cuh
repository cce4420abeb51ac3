// ffm_tile.cuh -- the FFM fast path for batches whose samples all have distinct fields
// (Criteo-shaped data): "load rows -> compute in shared memory -> store rows".
//
//  k_ffm_tile        persistent CTAs (one per SM), warp-specialised:
//                    metadata warps : prefetch, several samples ahead, each sample's CSR row, the class of
//                                     every occurrence (fused / staged), row locators, linear records; rows
//                                     are kept sorted by class so that warps rarely mix classes
//                    loader warp    : hands out sample-sized spans of the shared-memory row ring and issues
//                                     one cp.async.bulk (TMA bulk copy, mbarrier expect_tx) per feature row:
//                                     z and n planes of a fused row, the materialised w plane of a staged row
//                    consumer warps : pass 1 w = W(n,z) for fused rows (ffm.cpp:72-88, stored as the stale
//                                     w the reference keeps), logit (ffm.cpp:57-70), g = sigmoid(logit) - y;
//                                     pass 2 FTRL update of the fused rows in place in shared memory
//                                     (ffm.cpp:90-136 telescoped)
//                    storer warp    : one bulk store per fused row: updated (z',n') back into the table
//  k_row_touch / k_row_materialise   owner-side pre-pass: w of the staged rows, only the touched slices
//  k_ffm_regrad_rows row-centric update of the staged rows: the per-occurrence gradients are re-derived from
//                    the partner rows' w slices (32-byte gathers, mostly L2 hits) and reduced in registers:
//                    closed-form update, a partial for k_ffm_combine, or (sharded runs) the row's sum into
//                    its owner's inbox
//
// HBM traffic per touched coordinate: rows that occur once: 8 B read (z,n) + 12 B written (z',n',w) = the
// algorithmic 20 B; other rows: 12 B once (materialise) + 20 B once (update) per distinct coordinate; per
// occurrence only the w plane is read (twice: row-wise by the sample kernel, slice-wise by the row kernel),
// which L2 serves for the rows that make up most occurrences.
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
// the same on a precomputed shared-memory address (the generic -> shared conversion costs ~10 instructions
// per use); `hint_ns` > 0: the hardware may suspend the warp for about that long before it reports "not yet"
// -- the helper warps wait with a long hint so that their polling does not take issue slots from the consumers
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity, uint32_t hint_ns = 0) {
  uint32_t ok;
  do {
    if (hint_ns)
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok) : "r"(bar_addr), "r"(parity), "r"(hint_ns) : "memory");
    else
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
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
// + one row-loader warp (TMA bulk copies)
__host__ __device__ constexpr int tile_threads(int consumers) { return consumers + 32 * (1 + TILE_META_WARPS); }
constexpr int TILE_MAX_CONSUMERS = 768;
constexpr int TILE_MAX_STAGE = 4;    // samples in flight in the row ring (power of two)
constexpr int TILE_MAX_META = 8;     // metadata slots

struct TileGeom {
  int f_cap;        // rows per stage (= n_fields: samples with distinct fields have at most that many)
  int stride;       // floats a fused row (z and n planes) takes in the row ring: 2*ld + pad, conflict-free columns
  int stride1;      // floats a staged row (w plane only) takes: ld + pad1, same residue as `stride` modulo 32 floats
  int n_stage;      // the row ring holds n_stage * f_cap * stride floats
  int inflight;     // samples that may share the ring (2..TILE_MAX_STAGE)
  int n_meta;       // metadata slots (> n_stage: metadata runs ahead of the row ring)
  int consumers;    // consumer threads (multiple of 32)
  int dbg;          // experiment switches (FTRL_B200_TILE_DBG): 1 skip w stores, 2 skip row stores, 64 L2 cache hints
  uint32_t helper_ns;  // suspend hint (ns) of the helper warps' mbarrier waits, 0: plain polling
  size_t smem_bytes;
};

// Per-sample metadata, one slot per FIELD (samples of the tile path have distinct fields, so a field names at
// most one row of the sample): work items are fixed (field pair, factor chunk) triples, the same for every
// sample, and the only per-sample indirection left in the inner loops is the row's offset in the row ring.
struct RowEnt {    // 8 bytes
  int32_t off;     // offset of the row inside its sample's span of the row ring (floats); 0 when the field is absent
  float x;         // value; 0 when the field is absent (the aliased span start then contributes 0)
};
struct RowAux {    // 8 bytes
  int32_t loc;     // row locator (RowSpace): >= 0 local row, < 0: -1 - head position in the remote-row cache
  int32_t pos;     // -1: row finalised here (fused); staged (reduced by k_ffm_regrad_rows): base of the occurrence's
                   // image of fused-partner gradient slices (in slices of k floats)
};
struct SampleMeta {
  RowEnt *ent;      // [f_cap]
  RowAux *aux;      // [f_cap]
  float4 *lin;      // [f_cap] {z, n, w, -} of the linear coordinate, prefetched
  int32_t *hdr;     // [0] fused rows, [1] label, [2] floats of the row ring the sample needs, [3] present rows,
                    // [4],[5] present-field mask lo/hi, [6],[7] fused-field mask lo/hi
  uint8_t *flist;   // [f_cap] fields of the fused rows, ascending
};
constexpr int TILE_MAX_FR = 8;  // fused-row vectors per consumer thread (z kept in registers between the passes)

__host__ __device__ inline size_t tile_meta_bytes(int f_cap) {
  // RowEnt (8 B) + RowAux (8 B) + lin (16 B) per field, + header 32 B, + flist[f_cap] rounded to 16
  return (size_t)f_cap * 32 + 32 + (size_t)((f_cap + 15) / 16) * 16;
}
__host__ __device__ inline size_t tile_stage_bytes(int f_cap, int stride) {
  return (size_t)f_cap * stride * sizeof(float);
}
__host__ __device__ inline size_t tile_smem_bytes(int f_cap, int stride, int n_stage, int n_meta) {
  return tile_stage_bytes(f_cap, stride) * n_stage + tile_meta_bytes(f_cap) * n_meta;
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

// L2 eviction policies for the bulk copies (experiment switch 64): rows that are touched once per step stream
// through (evict_first) so that the w planes of the hot rows, read once per occurrence, stay resident
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst_gmem, const void *src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol)
               : "memory");
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

// Thread roles: [0, consumers) compute; then one row-loader warp (bulk loads into the row ring) and
// TILE_META_WARPS warps that prefetch sample metadata into a deeper ring so that the row loader never waits
// on a dependent global-load chain.
//
// Rows that occur once in the batch ("fused") are read (z, n), materialised, updated and written back here:
// the algorithmic 20 B per coordinate.  Rows that occur several times ("staged") only lend their w plane
// (materialised by k_row_materialise before this kernel) to the dot products; their gradient is re-derived
// row by row in k_ffm_regrad_rows from the w slices of the partner rows, so nothing per occurrence is
// written to HBM.
//
// Per sample the consumers run three passes:
//   pass 0  fused rows, row-centric (thread = one float4 vector of one fused row, coalesced): w = W(n, z)
//           (ffm.cpp:72-88) written over the z slot in shared memory and to the table (the stale-by-one w the
//           reference keeps); z stays in registers
//   pass 1  every (field pair, factor chunk) item, identical work for fused and staged rows: both slices are w
//           now; logit (ffm.cpp:57-70), g = sigmoid(logit) - y
//   pass 2  fused rows, same thread mapping as pass 0: FTRL update (ffm.cpp:90-136 telescoped)
//           with the partner's w slice from shared memory; z', n' stored straight to the table (512 B per warp)
template <bool PRECISE, int IPT, int FR, int MAXC>
__global__ void __launch_bounds__(tile_threads(MAXC), 1)
k_ffm_tile(Batch b, Dims d, Hyper h, ItemDecode dec, TileGeom geo, const int32_t *__restrict__ batch_flags,
           const __grid_constant__ RowSpace rsp, const float4 *__restrict__ bias, const uint32_t *__restrict__ pair_lut,
           const int32_t *__restrict__ occ_pos, const SegScan *__restrict__ scan,
           const int32_t *__restrict__ sbase, float *__restrict__ sparse, float *__restrict__ g_out,
           float *__restrict__ logit_out) {
  if (batch_flags[0] == 0) return;  // some sample repeats a field: the generic kernels take this batch
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar_full[TILE_MAX_STAGE], bar_done[TILE_MAX_STAGE];
  __shared__ uint64_t bar_mfull[TILE_MAX_META], bar_mfree[TILE_MAX_META];
  __shared__ float s_red[2][32];  // per-warp partial logits, double-buffered by sample parity

  const int tid = threadIdx.x;
  const int n_cons = geo.consumers;
  const int n_cons_warps = n_cons >> 5;
  const int lane = tid & 31;
  // 0 consumer, 1 row loader, 2 metadata
  const int role = tid < n_cons ? 0 : (tid < n_cons + 32 ? 1 : 2);
  const int ld = d.ld, k = d.k, NF = d.n_fields;
  const int stride = geo.stride, stride1 = geo.stride1, f_cap = geo.f_cap, NS = geo.n_stage, MD = geo.n_meta;
  const size_t stage_bytes = tile_stage_bytes(f_cap, stride);
  // The row ring: NS * stage_bytes of shared memory handed out in sample-sized spans.  A fused row takes
  // `stride` floats (z, n), a staged row `stride1` (w only), so batches with many duplicated rows keep 3-4
  // samples in flight where all-fused samples keep 2.  Up to TILE_MAX_STAGE samples share the ring.
  constexpr int NSLOT = TILE_MAX_STAGE;
  const int ring_floats = (int)((size_t)NS * stage_bytes / sizeof(float));
  float *ring = reinterpret_cast<float *>(smem_raw);
  __shared__ int s_need[NSLOT];
  const size_t meta_bytes = tile_meta_bytes(f_cap);
  const int64_t rs = 3 * (int64_t)ld;
  const uint32_t row_bytes = (uint32_t)(2 * ld * sizeof(float));
  unsigned char *meta_base = smem_raw + (size_t)NS * stage_bytes;

  auto sample_meta = [&](int slot) {
    SampleMeta m;
    unsigned char *p = meta_base + (size_t)slot * meta_bytes;
    m.ent = reinterpret_cast<RowEnt *>(p);
    p += (size_t)f_cap * 8;
    m.aux = reinterpret_cast<RowAux *>(p);
    p += (size_t)f_cap * 8;
    m.lin = reinterpret_cast<float4 *>(p);
    p += (size_t)f_cap * 16;
    m.hdr = reinterpret_cast<int32_t *>(p);
    p += 32;
    m.flist = p;
    return m;
  };

  if (tid == 0) {
    for (int st = 0; st < NSLOT; st++) {
      mbar_init(&bar_full[st], 1);
      mbar_init(&bar_done[st], n_cons_warps);
    }
    for (int sl = 0; sl < MD; sl++) {
      mbar_init(&bar_mfull[sl], 1);
      mbar_init(&bar_mfree[sl], n_cons_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_mine = b.n_rows > blockIdx.x ? (int)((b.n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  if (role == 2) {
    // =========================== metadata warps ===========================
    const int mw = (tid - n_cons - 32) >> 5;
    RingCursor mc;
    mc.init(mw, MD);
    for (int it = mw; it < n_mine; it += TILE_META_WARPS, mc.advance(TILE_META_WARPS, MD)) {
      const int slot = mc.slot;
      if (mc.round > 0) mbar_wait_a(smem_u32(&bar_mfree[slot]), mc.prev_parity(), geo.helper_ns);
      SampleMeta m = sample_meta(slot);
      const int64_t s = blockIdx.x + (int64_t)it * gridDim.x;
      const int64_t r0 = b.row_ptr[s];
      const int F = (int)min((int64_t)1 << 20, b.row_ptr[s + 1] - r0);
      for (int f = lane; f < NF; f += 32) m.ent[f] = RowEnt{0, 0.f};
      __syncwarp();
      unsigned long long present = 0ull, fusedm = 0ull;
      for (int base = 0; base < F; base += 32) {
        const int t = base + lane;
        unsigned long long pbit = 0ull, fbit = 0ull;
        if (t < F) {
          const int32_t fl = b.field[r0 + t], ft = b.feat[r0 + t];
          if (feat_valid(d, fl, ft)) {
            const int32_t pos = occ_pos[r0 + t];
            RowAux a;
            a.pos = pos < 0 ? -1 : sbase[pos];  // staged rows: base of the occurrence's image of fused-partner slices
            if ((ft & rsp.Gm1) == rsp.rank) {
              a.loc = ft >> rsp.log2G;
              m.lin[fl] = rsp.lin[a.loc];
            } else {  // remote rows are never fused: pos >= 0, the row's cache slot is its sorted head position
              const int32_t head = scan[pos].start;
              a.loc = -1 - head;
              m.lin[fl] = make_float4(0.f, 0.f, rsp.rc_lin[head], 0.f);
            }
            m.aux[fl] = a;
            m.ent[fl].x = b.val[r0 + t];
            pbit = 1ull << fl;
            if (pos < 0) fbit = pbit;
          }
        }
        present |= ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(pbit >> 32)) << 32) |
                   __reduce_or_sync(0xffffffffu, (unsigned)pbit);
        fusedm |= ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(fbit >> 32)) << 32) |
                  __reduce_or_sync(0xffffffffu, (unsigned)fbit);
      }
      __syncwarp();
      // the list of fused fields (the loader places the rows in the ring and fills in their offsets)
      for (int f = lane; f < NF; f += 32)
        if ((fusedm >> f) & 1ull) m.flist[__popcll(fusedm & ((1ull << f) - 1ull))] = (uint8_t)f;
      const int need = __popcll(fusedm) * stride + (__popcll(present) - __popcll(fusedm)) * stride1;
      if (lane == 0) {
        m.hdr[0] = __popcll(fusedm);
        m.hdr[1] = b.label[s];
        m.hdr[2] = need;
        m.hdr[3] = __popcll(present);
        m.hdr[4] = (int32_t)(uint32_t)present;
        m.hdr[5] = (int32_t)(uint32_t)(present >> 32);
        m.hdr[6] = (int32_t)(uint32_t)fusedm;
        m.hdr[7] = (int32_t)(uint32_t)(fusedm >> 32);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_mfull[slot]);
    }
    return;
  }

  const bool hints = (geo.dbg & 64) != 0;

  if (role == 1) {
    // =========================== row loader warp ===========================
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    RingCursor mc;
    mc.init(0, MD);
    // The row ring is a circular buffer of floats: rows are placed one after another (a row that would cross
    // the end starts over at 0), samples retire in order.  `live` = floats between the oldest live row and
    // `head`, wrap waste included; a sample is admitted when its rows plus one worst-case wrap gap fit.
    int head = 0;     // next free float of the ring
    int live = 0;
    int tail_it = 0;  // oldest sample whose rows have not been handed back by the consumers
    for (int it = 0; it < n_mine; it++, mc.advance(1, MD)) {
      const int st = it & (NSLOT - 1);
      const int slot = mc.slot;
      mbar_wait_a(smem_u32(&bar_mfull[slot]), mc.parity(), geo.helper_ns);
      SampleMeta m = sample_meta(slot);
      const int nf = m.hdr[0], np = m.hdr[3];
      const int need = m.hdr[2];
      const unsigned long long present = ((unsigned long long)(uint32_t)m.hdr[5] << 32) | (uint32_t)m.hdr[4];
      const unsigned long long fusedm = ((unsigned long long)(uint32_t)m.hdr[7] << 32) | (uint32_t)m.hdr[6];
      while (it - tail_it >= geo.inflight || (tail_it < it && live + need + stride > ring_floats)) {
        mbar_wait_a(smem_u32(&bar_done[tail_it & (NSLOT - 1)]), (uint32_t)((tail_it >> 2) & 1), geo.helper_ns);
        live -= s_need[tail_it & (NSLOT - 1)];
        tail_it++;
      }
      // place the rows in field order
      int used = 0;
      for (int base = 0; base < NF; base += 32) {
        const int f = base + lane;
        const bool here = f < NF && ((present >> f) & 1ull);
        const int sz = here ? (((fusedm >> f) & 1ull) ? stride : stride1) : 0;
        int inc = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        int pos = head + inc - sz;
        const unsigned over = __ballot_sync(0xffffffffu, here && pos + sz > ring_floats);
        if (over) {  // the first row that does not fit before the end, and every row after it, start over at 0
          const int fs = __ffs(over) - 1;
          const int wrap_pos = __shfl_sync(0xffffffffu, pos, fs);
          if (lane >= fs) pos -= wrap_pos;
          used += ring_floats - wrap_pos;
        }
        if (here) m.ent[f].off = pos;
        head = __shfl_sync(0xffffffffu, pos + sz, 31);
        used += __shfl_sync(0xffffffffu, inc, 31);
      }
      live += used;
      __syncwarp();
      if (lane == 0) s_need[st] = used;
      __syncwarp();
      // TMA bulk copies.  Fused rows bring z and n (they are updated here); staged rows bring only the w plane
      // their owner materialised
      // (experiment switches: 8 = staged rows are not loaded at all, 16 = loaded as two half copies)
      if (lane == 0)
        mbar_expect_tx(&bar_full[st], (uint32_t)(nf * (int)row_bytes + ((geo.dbg & 8) ? 0 : (np - nf) * (int)(row_bytes / 2))));
      __syncwarp();
      for (int f = lane; f < NF; f += 32) {
        if (!((present >> f) & 1ull)) continue;
        float *dst = ring + m.ent[f].off;
        const int32_t loc = m.aux[f].loc;
        if ((fusedm >> f) & 1ull) {
          if (hints) bulk_g2s_hint(dst, rsp.tab + (int64_t)loc * rs, row_bytes, &bar_full[st], pol_stream);
          else bulk_g2s(dst, rsp.tab + (int64_t)loc * rs, row_bytes, &bar_full[st]);
        } else if (geo.dbg & 16) {
          const uint32_t h1 = (row_bytes / 4) & ~15u;
          bulk_g2s(dst, rsp.w_plane(loc, ld), h1, &bar_full[st]);
          bulk_g2s(reinterpret_cast<char *>(dst) + h1, reinterpret_cast<const char *>(rsp.w_plane(loc, ld)) + h1, row_bytes / 2 - h1, &bar_full[st]);
        } else if (!(geo.dbg & 8)) {
          if (hints) bulk_g2s_hint(dst, rsp.w_plane(loc, ld), row_bytes / 2, &bar_full[st], pol_keep);
          else bulk_g2s(dst, rsp.w_plane(loc, ld), row_bytes / 2, &bar_full[st]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[st]);
    }
    return;
  }

  // =========================== consumer warps ===========================
  const float bias_w = [&] {
    const float4 bz = *bias;
    return weight_from<PRECISE>(bz.x, f_sqrt<PRECISE>(bz.y), h);
  }();
  // the thread's pair items, fixed for the whole kernel: (field m, field n, factor chunk c), m < n
  const uint32_t n_items = (uint32_t)NF * (uint32_t)(NF - 1) / 2u * dec.C;
  // (packed: registers are the scarce resource of this kernel)
  uint32_t pmn[IPT], pcol[IPT];  // m | n << 8 (0xffff: no item) ; column of slice A | column of slice B << 16
#pragma unroll
  for (int j = 0; j < IPT; j++) {
    const uint32_t item = tid + j * n_cons;
    pmn[j] = 0xffffu;
    pcol[j] = 0;
    if (item < n_items) {
      uint32_t p, c;
      dec(item, p, c);
      const uint32_t e = pair_lut[p];
      const uint32_t mi = e & 0xffffu, ni = e >> 16;
      pmn[j] = mi | (ni << 8);
      // slice A = (row of field m, field n), slice B = (row of field n, field m)
      pcol[j] = (ni * k + c * 4) | ((mi * k + c * 4) << 16);
    }
  }
  // the thread's fused-row vectors: (i-th fused row of the sample, float4 vector v of the row)
  const int nvec = ld >> 2, vpf = k >> 2;
  uint32_t fpk[FR];  // vector v | partner field of that slice << 16 | index i of the fused row << 24
#pragma unroll
  for (int r = 0; r < FR; r++) {
    const int fitem = tid + r * n_cons;
    const int i = fitem / nvec, v = fitem - i * nvec;
    fpk[r] = (uint32_t)v | ((uint32_t)(v / vpf) << 16) | ((uint32_t)min(i, 255) << 24);
  }
  auto f_v = [&](int r) { return (int)(fpk[r] & 0xffffu); };
  auto f_n = [&](int r) { return (int)((fpk[r] >> 16) & 0xffu); };
  auto f_i = [&](int r) { return (int)(fpk[r] >> 24); };

  const uint32_t a_full = smem_u32(bar_full), a_done = smem_u32(bar_done);
  const uint32_t a_mfull = smem_u32(bar_mfull), a_mfree = smem_u32(bar_mfree);
  RingCursor mc;
  mc.init(0, MD);
  for (int it = 0; it < n_mine; it++, mc.advance(1, MD)) {
    const int st = it & (NSLOT - 1);
    const int slot = mc.slot;
    const int64_t s = blockIdx.x + (int64_t)it * gridDim.x;
    SampleMeta m = sample_meta(slot);
    mbar_wait_a(a_mfull + 8 * slot, mc.parity());
    mbar_wait_a(a_full + 8 * st, (uint32_t)((it >> 2) & 1));
    const int nf = m.hdr[0];
    const unsigned long long present = ((unsigned long long)(uint32_t)m.hdr[5] << 32) | (uint32_t)m.hdr[4];
    const unsigned long long fusedm = ((unsigned long long)(uint32_t)m.hdr[7] << 32) | (uint32_t)m.hdr[6];

    // ---- pass 0: w of the fused rows, in place over z ----
    float4 zreg[FR];
    unsigned long long fm = ~0ull;  // byte r: field of the thread's r-th fused row in this sample, 0xff: nothing to do
    if (nf > 0) {
#pragma unroll
      for (int r = 0; r < FR; r++) {
        if (f_i(r) < nf) {
          const int mf = m.flist[f_i(r)];
          // slices no partner touches (own field, absent fields) keep their z, n, w (ffm.cpp:72-88 never visits them)
          if (f_n(r) != mf && ((present >> f_n(r)) & 1ull)) {
            fm = (fm & ~(0xffull << (8 * r))) | ((unsigned long long)mf << (8 * r));
            float *zp = ring + m.ent[mf].off + f_v(r) * 4;
            const float4 z = *reinterpret_cast<const float4 *>(zp);
            const float4 w = weight4<PRECISE>(z, *reinterpret_cast<const float4 *>(zp + ld), h);
            *reinterpret_cast<float4 *>(zp) = w;
            if (!(geo.dbg & 1)) *reinterpret_cast<float4 *>(rsp.tab + (int64_t)m.aux[mf].loc * rs + 2 * ld + f_v(r) * 4) = w;
            zreg[r] = z;
          }
        }
      }
      named_bar_sync(2, n_cons);  // every fused row's w is in place before any pair item reads it
    }

    // ---- pass 1: logit ----
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < IPT; j++) {
      if (pmn[j] != 0xffffu) {
        const RowEnt eA = m.ent[pmn[j] & 0xffu], eB = m.ent[pmn[j] >> 8];
        const float4 wA = *reinterpret_cast<const float4 *>(ring + eA.off + (pcol[j] & 0xffffu));
        const float4 wB = *reinterpret_cast<const float4 *>(ring + eB.off + (pcol[j] >> 16));
        const float dot = fmaf(wA.x, wB.x, fmaf(wA.y, wB.y, fmaf(wA.z, wB.z, wA.w * wB.w)));
        // an absent field has x = 0 and aliases the start of the ring: whatever is read there must not count
        const float xx = eA.x * eB.x;
        acc = xx != 0.f ? fmaf(dot, xx, acc) : acc;
      }
    }
    for (int f = tid; f < NF; f += n_cons) {
      if (!((present >> f) & 1ull)) continue;
      const float4 e = m.lin[f];
      const float w = ((fusedm >> f) & 1ull) ? weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h) : e.z;
      acc = fmaf(w, m.ent[f].x, acc);
    }
    // consumer-wide sum: every warp adds the per-warp partials itself (same values, same order -> the same
    // logit in every warp).  The buffer of sample `it` is next written for sample it + 2, i.e. by a warp that
    // has passed the barrier of sample it + 1, which every warp reaches only after this read.
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

    // ---- pass 2: FTRL update of the fused rows; every (row, field) slice belongs to exactly one thread ----
    if (nf > 0) {
#pragma unroll
      for (int r = 0; r < FR; r++) {
        const int mf = (int)((fm >> (8 * r)) & 0xffull);
        if (mf != 0xff) {
          const RowEnt eM = m.ent[mf], eP = m.ent[f_n(r)];
          const float *zp = ring + eM.off + f_v(r) * 4;
          const float4 w = *reinterpret_cast<const float4 *>(zp);
          float4 nn = *reinterpret_cast<const float4 *>(zp + ld);
          // partner slice = (row of field fn, field mf), same factor chunk
          const float4 wp = *reinterpret_cast<const float4 *>(ring + eP.off + mf * k + (f_v(r) - f_n(r) * vpf) * 4);
          float4 z = zreg[r];
          const float gx = g * (eM.x * eP.x);
          apply4<PRECISE>(z, nn, w, wp, gx, h);
          // a staged partner's row kernel would have to fetch this row's slice from HBM just for this one
          // occurrence: leave the finished gradient slice g x x w in the partner occurrence's image instead
          if (!((fusedm >> f_n(r)) & 1ull))
            __stcs(reinterpret_cast<float4 *>(sparse + ((int64_t)m.aux[f_n(r)].pos + f_i(r)) * k + (f_v(r) - f_n(r) * vpf) * 4),
                   make_float4(gx * w.x, gx * w.y, gx * w.z, gx * w.w));
          if (!(geo.dbg & 2)) {
            float *grow = rsp.tab + (int64_t)m.aux[mf].loc * rs + f_v(r) * 4;
            *reinterpret_cast<float4 *>(grow) = z;
            *reinterpret_cast<float4 *>(grow + ld) = nn;
          }
        }
      }
      // linear coordinate of the fused rows (staged rows: k_ffm_regrad_rows)
      for (int i = tid; i < nf; i += n_cons) {
        const int mf = m.flist[i];
        float4 e = m.lin[mf];
        const float gi = g * m.ent[mf].x;
        const float w = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
        e.z = w;
        ftrl_apply<PRECISE>(e.x, e.y, w, gi, gi * gi, h);
        rsp.lin[m.aux[mf].loc] = e;
      }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_a(a_done + 8 * st);     // the sample's rows go back to the loader
      mbar_arrive_a(a_mfree + 8 * slot);  // the metadata slot goes back to the metadata warps
    }
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
k_row_touch(int32_t nnz, uint32_t sentinel, int32_t ch, const int32_t *__restrict__ batch_flags,
            const int32_t *__restrict__ n_chunks_p, const int32_t *__restrict__ chunk_pos,
            const uint32_t *__restrict__ skey, const uint32_t *__restrict__ socc, const SegScan *__restrict__ scan,
            PmaskSrc pm, unsigned long long *__restrict__ rowmask) {
  if (batch_flags[0] == 0) return;
  const int wib = threadIdx.x >> 5;
  const int n_chunks = *n_chunks_p;
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_info<true>(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid) continue;
    row_touch_chunk(c, ci, ch, socc, pm, rowmask);
  }
}

template <bool PRECISE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_row_materialise(Dims d, Hyper h, int32_t nnz, uint32_t sentinel, int32_t ch, const int32_t *__restrict__ batch_flags,
                  const int32_t *__restrict__ n_chunks_p, const int32_t *__restrict__ chunk_pos,
                  const uint32_t *__restrict__ skey, const SegScan *__restrict__ scan,
                  const unsigned long long *__restrict__ rowmask, float *__restrict__ tab, float4 *__restrict__ lin) {
  if (batch_flags[0] == 0) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const int vpf = d.k >> 2;  // float4 vectors per field slice
  for (int c = blockIdx.x * WARPS + wib; c < n_chunks; c += gridDim.x * WARPS) {
    const ChunkInfo ci = chunk_head_info(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid || !ci.row_head) continue;
    const unsigned long long mask = rowmask[c];
    float *row = tab + (int64_t)ci.key * rs;
    for (int v = lane; v < d.n_fields * vpf; v += 32) {
      if (!((mask >> (v / vpf)) & 1ull)) continue;
      const float4 z = reinterpret_cast<const float4 *>(row)[v], n = reinterpret_cast<const float4 *>(row + ld)[v];
      reinterpret_cast<float4 *>(row + 2 * ld)[v] = weight4<PRECISE>(z, n, h);
    }
    if (lane == 0) {
      const float4 e = lin[ci.key];
      lin[ci.key].z = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_ffm_regrad_rows: the update of the staged rows (rows that occur more than once in the batch, or whose
// owner is another rank), row by row.  For an occurrence of row i (field f_i) in sample s the gradient of
// slice (i, field n) is g_s x_i x_n w[row of field n in s][f_i, :] (ffm.cpp:112,117); instead of having the
// sample kernel write that per-occurrence image to HBM and reading it back, it is re-derived here from the
// partner rows' w slices -- 32-byte gathers that mostly hit L2, because the partners of a duplicated row are
// themselves mostly duplicated (hot) rows.  The partner of (s, n) comes from the canonical table (prep.cuh).
// Work item = (chunk of <= 32 occurrences of one row, part of 32 float4 vectors of the row): one warp; lane v
// owns destination vector v = (partner field n, factor chunk), accumulates (sum g, sum g^2) in registers over
// the occurrences in sorted (= sample) order -> deterministic.  Then the closed-form update (row fits one
// chunk), a partial for k_ffm_combine, or (sharded runs) the row's sum into its owner's inbox.
// ---------------------------------------------------------------------------------------------
constexpr int REGRAD_U = 4;  // occurrences per pipeline stage: their canon entries / w gathers are in flight together

template <bool PRECISE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 3)
k_ffm_regrad_rows(Dims d, Hyper h, int32_t nnz, const int32_t *__restrict__ batch_flags,
                  const __grid_constant__ RowSpace rsp, int32_t ch, const int32_t *__restrict__ n_chunks_p,
                  const int32_t *__restrict__ chunk_pos, const uint32_t *__restrict__ skey,
                  const uint32_t *__restrict__ socc, const SegScan *__restrict__ scan,
                  const int4 *__restrict__ srec, const int32_t *__restrict__ sbase,
                  const float *__restrict__ sparse, const float *__restrict__ g_in,
                  const CanonEntry *__restrict__ canon, float *__restrict__ part, float2 *__restrict__ part_lin,
                  const __grid_constant__ Export ex) {
  if (batch_flags[0] == 0) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ld = d.ld, rs = 3 * ld;
  const int n_chunks = *n_chunks_p;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  const int nvec = (int)(ld >> 2);
  const int parts = (nvec + 31) >> 5;
  const int vpf = d.k >> 2;  // float4 vectors per field slice
  const int NF = d.n_fields;
  float *tab = rsp.tab;
  float4 *lin = rsp.lin;
  const uint64_t pol_keep = policy_evict_last();
  const int64_t n_items = (int64_t)n_chunks * parts;
  for (int64_t item = (int64_t)blockIdx.x * WARPS + wib; item < n_items; item += (int64_t)gridDim.x * WARPS) {
    // part-major order: rows are sorted by id, i.e. (Criteo-shaped ids) by field, so the warps in flight all
    // gather the same column of the table (slice f of every partner row); one part at a time keeps that
    // working set -- a third of the column and of the canonical table -- inside L2
    const int part_i = (int)(item / n_chunks), c = (int)(item - (int64_t)part_i * n_chunks);
    const ChunkInfo ci = chunk_info<true>(c, nnz, sentinel, ch, chunk_pos, skey, scan);
    if (!ci.valid) continue;
    const bool whole_row = ci.row_head && ci.row_last;
    const int v = part_i * 32 + lane;
    const int nfld = v / vpf;                 // destination slice = partner field
    const bool on = v < nvec && nfld < NF;
    const int col = (v - nfld * vpf) * 4;     // float offset inside a slice
    const int own = nfld * d.k;
    const int n_occ = ci.p1 - ci.p0;
    // lane l holds the metadata of occurrence p0 + l
    int my_s = 0, my_fk = -1, my_sb = 0;
    float my_gx = 0.f;
    if (lane < n_occ) {
      const int4 rec = __ldg(srec + ci.p0 + lane);  // {sample, value, field * k, -}: sequential along the sorted list
      my_sb = __ldg(sbase + ci.p0 + lane);
      my_s = rec.x;
      my_gx = g_in[my_s] * __int_as_float(rec.y);
      my_fk = rec.z;
    }
    // sharded runs: the sum goes to the row's owner unless this rank owns the row and is its only contributor
    const int32_t dst = (ex.on && whole_row) ? ex.dst_at[ci.p0] : -2;
    const int64_t lrow = (int64_t)(ci.key >> ex.log2G);
    float *row = tab + lrow * rs;
    // the row's own planes are needed only at the very end: fetch them now, under the gathers
    const bool apply_here = whole_row && dst < 0 && v < nvec;
    float4 rz = make_float4(0.f, 0.f, 0.f, 0.f), rn = rz, rw = rz;
    if (apply_here) {
      rz = __ldcs(reinterpret_cast<const float4 *>(row) + v);
      rn = __ldcs(reinterpret_cast<const float4 *>(row + ld) + v);
      rw = __ldcs(reinterpret_cast<const float4 *>(row + 2 * ld) + v);
    }
    // software pipeline over stages of REGRAD_U occurrences: the canon entries of stage b + 1 are requested
    // before the w gathers of stage b are consumed, so one L2 round trip per stage is exposed instead of two
    auto load_canon = [&](int p, CanonEntry (&e)[REGRAD_U]) {
#pragma unroll
      for (int u = 0; u < REGRAD_U; u++) {
        const int sj = __shfl_sync(0xffffffffu, my_s, (p + u) & 31);
        e[u].loc = CANON_NONE;
        e[u].x = 0.f;
        if (on && p + u < n_occ) {
          // the canonical table is re-read once per field of a sample, spread over the whole kernel: keep it in L2
          int2 t;
          asm volatile("ld.global.nc.L2::cache_hint.v2.b32 {%0, %1}, [%2], %3;"
                       : "=r"(t.x), "=r"(t.y)
                       : "l"(canon + (int64_t)sj * NF + nfld), "l"(pol_keep));
          e[u].loc = t.x;
          e[u].x = __int_as_float(t.y);
        }
      }
    };
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    CanonEntry e[REGRAD_U], en[REGRAD_U] = {};
    load_canon(0, e);
    for (int p = 0; p < n_occ; p += REGRAD_U) {
      float4 w[REGRAD_U];
      float gxx[REGRAD_U];
#pragma unroll
      for (int u = 0; u < REGRAD_U; u++) {
        const int fk = __shfl_sync(0xffffffffu, my_fk, (p + u) & 31);
        const int sb = __shfl_sync(0xffffffffu, my_sb, (p + u) & 31);
        gxx[u] = __shfl_sync(0xffffffffu, my_gx, (p + u) & 31) * e[u].x;
        // the slice of the row's own field is touched by no partner (fields are distinct)
        const bool take = e[u].loc != CANON_NONE && fk != own;
        if (take && canon_is_fused(e[u].loc)) {
          // fused partner: the sample kernel left the finished gradient slice in this occurrence's image
          w[u] = __ldcs(reinterpret_cast<const float4 *>(sparse + ((int64_t)sb + (e[u].loc & 0xff)) * d.k + col));
          gxx[u] = 1.0f;
        } else {
          w[u] = take ? __ldg(reinterpret_cast<const float4 *>(rsp.w_plane(e[u].loc, ld) + fk + col))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (p + REGRAD_U < n_occ) load_canon(p + REGRAD_U, en);
#pragma unroll
      for (int u = 0; u < REGRAD_U; u++) {
        const float4 gv = make_float4(gxx[u] * w[u].x, gxx[u] * w[u].y, gxx[u] * w[u].z, gxx[u] * w[u].w);
        a0.x += gv.x; a0.y += gv.y; a0.z += gv.z; a0.w += gv.w;
        a1.x = fmaf(gv.x, gv.x, a1.x); a1.y = fmaf(gv.y, gv.y, a1.y);
        a1.z = fmaf(gv.z, gv.z, a1.z); a1.w = fmaf(gv.w, gv.w, a1.w);
        e[u] = en[u];
      }
    }
    if (v < nvec) {
      if (whole_row && dst >= 0) {
        // one occurrence: sum g^2 = g^2, the owner squares it (half the bytes over NVLink)
        float *o = ex.inbox[ci.key & ex.Gm1] + (int64_t)dst * 2 * ld;
        reinterpret_cast<float4 *>(o)[v] = a0;
        if (n_occ > 1) reinterpret_cast<float4 *>(o + ld)[v] = a1;
      } else if (whole_row) {
        const bool any = a1.x != 0.f || a1.y != 0.f || a1.z != 0.f || a1.w != 0.f || a0.x != 0.f || a0.y != 0.f ||
                         a0.z != 0.f || a0.w != 0.f;
        if (any) {
          ftrl_apply<PRECISE>(rz.x, rn.x, rw.x, a0.x, a1.x, h);
          ftrl_apply<PRECISE>(rz.y, rn.y, rw.y, a0.y, a1.y, h);
          ftrl_apply<PRECISE>(rz.z, rn.z, rw.z, a0.z, a1.z, h);
          ftrl_apply<PRECISE>(rz.w, rn.w, rw.w, a0.w, a1.w, h);
          __stcs(reinterpret_cast<float4 *>(row) + v, rz);
          __stcs(reinterpret_cast<float4 *>(row + ld) + v, rn);
        }
      } else {
        float *pdst = part + (int64_t)ci.slot * 2 * ld;
        reinterpret_cast<float4 *>(pdst)[v] = a0;
        reinterpret_cast<float4 *>(pdst + ld)[v] = a1;
      }
    }
    if (part_i == 0) {
      // linear coordinate: g_s x_i per occurrence
      const float sg = warp_sum(my_gx), sg2 = warp_sum(my_gx * my_gx);
      if (lane == 0) {
        if (whole_row && dst >= 0) {
          ex.inbox_lin[ci.key & ex.Gm1][dst] = make_float2(sg, sg2);
        } else if (whole_row) {
          float4 le = lin[lrow];
          ftrl_apply<PRECISE>(le.x, le.y, le.z, sg, sg2, h);
          lin[lrow] = le;
        } else {
          part_lin[ci.slot] = make_float2(sg, sg2);
        }
      }
    }
  }
}

}  // namespace ftrl
