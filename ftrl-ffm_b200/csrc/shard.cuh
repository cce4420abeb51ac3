// shard.cuh -- multi-GPU: feature-sharded tables, one process per GPU, peers mapped over NVLink.
//
// Row `feat` of lin / tab lives on rank feat mod G at local row feat div G; the samples of a
// global minibatch are split across ranks.  Every rank runs the single-GPU pipeline on its own samples; what
// crosses NVLink is one w plane IN and one gradient sum OUT per DISTINCT (row, rank) pair -- not per
// occurrence -- because duplicates are reduced where the samples live.  A step has two phases on every rank:
//
// INDEX phase (ids only; with early inputs it runs on its own stream and barrier channel, under the weight phase
// of the PREVIOUS step -- train_device_sharded in ftrl_b200.cu):
//   S1  k_prep_rows, radix sort by feature id, segmented OR-scan of the field masks, list of the distinct
//       rows of the local batch, published BUCKETED BY OWNER (k_owner_keys, one radix pass, k_publish_unique:
//       feature id, sorted head position, "single occurrence", field mask; k_publish_bounds: bucket bounds)
//   --  barrier 1 (device-side, flags in peer memory, index channel)
//   S2a owner side: k_fill_owned copies the run every rank published for this owner (rank order -> deterministic),
//       radix sort by local row = the contribution list; k_contrib_class: a row touched once, by its owner only,
//       is finalised inside that sample ("fused"); a row touched by its owner only is reduced and applied there;
//       every other contribution gets a slot of the owner's inbox (written into the contributing rank's dst_at);
//       then the local chunk list and its descriptors (k_chunk_desc)
// WEIGHT phase (compute stream, weight channel):
//   S2b k_owner_materialise writes w = W(n,z) for the slices the global batch touches into the table and pushes
//       it into the row cache of every remote rank that touches the row (one NVLink store stream per distinct
//       (row, rank))
//   --  barrier 2
//   S3  k_ffm_tile over the local samples (all addresses local: shard rows, row cache, staging) -- or, when some
//       sample of some rank repeats a field, the generic k_ffm_sample / k_ffm_rows (SH variants) on every rank;
//       LR / FM: k_lrfm_sample_sh, k_lrfm_rows / k_lrfm_combine (SH);
//       k_ffm_staged_rows / k_ffm_combine reduce the local duplicates and store each row's (sum g, sum g^2)
//       into the owner's inbox (Export); the local (sum g, sum g^2, loss) of the bias goes to every peer
//   --  barrier 3
//   S4  owner side: k_owner_apply sums the <= G inbox entries of a row in rank order and applies the
//       closed-form FTRL update; k_bias_apply sums the G bias partials in rank order, so the replicated
//       bias stays bit-identical on all ranks
// The exchange is not a separate all-to-all: the kernels that need remote data load / store it in place
// through peer pointers.
#pragma once
#include "common.cuh"
#include "ffm_tile.cuh"
#include "prep.cuh"

namespace ftrl {

// lives in device memory of every rank, mapped by all peers
// Two barrier channels: the index phase of step t+1 (ids only: S1 and the id part of S2) runs on its own stream
// under the weight-dependent phase of step t, so the two phases synchronise independently.  Everything the index
// phase publishes is double-buffered by the parity of the step (`par`): a set is rewritten two steps later, after
// the barriers of the step in between have shown that every rank is done reading it (see train_device_sharded).
struct SyncArea {
  uint32_t flag[2][MAX_SHARDS];   // flag[ch][q] = last barrier epoch rank q has reached on channel ch (0: weights, 1: ids)
  int32_t boff[2][MAX_SHARDS][MAX_SHARDS + 1];  // [par] boff[q][r] .. boff[q][r+1]: the part of rank q's distinct-row list
                                                // owned by rank r (the list is published bucketed by owner; written by q)
  int32_t simple[2][MAX_SHARDS];  // [par] simple[q] != 0: every sample of rank q's batch has distinct fields
  double red[MAX_SHARDS][4];      // red[q] = {sum g, sum g^2, sum loss, n_rows} of rank q's batch
  uint32_t abort_at[2][MAX_SHARDS];  // [par] abort_at[q] = step tag at which rank q asked every rank to skip the step
};

constexpr uint32_t UINFO_SINGLE = 1u << 30;  // uinfo = sorted head position | UINFO_SINGLE
constexpr uint32_t UINFO_POS = UINFO_SINGLE - 1;

struct Peers {
  int G, log2G, rank, pad;
  SyncArea *sync[MAX_SHARDS];
  const uint32_t *ukey[MAX_SHARDS];              // [u] feature id of the u-th distinct row of that rank's batch
  const uint32_t *uinfo[MAX_SHARDS];             // [u] sorted head position | single-occurrence flag
  const unsigned long long *umask[MAX_SHARDS];   // [u] fields of the row the rank's batch touches
  int32_t *dst_at[MAX_SHARDS];                   // [sorted head position] inbox slot, written by the owner
  float *rc_w[MAX_SHARDS];                       // [sorted head position][ld] that rank's cache of remote rows
  float *rc_lin[MAX_SHARDS];                     // [sorted head position]
};

// ---- S1: distinct rows of the local batch --------------------------------------------------------
struct MaskScan {
  int32_t start;  // sorted position of the row head at or before this position (0 when none in the range)
  int32_t pad;
  unsigned long long mask;
};
struct MaskScanOp {  // segmented OR: a range that contains a row head restarts the mask
  __device__ __forceinline__ MaskScan operator()(const MaskScan &a, const MaskScan &b) const {
    MaskScan r;
    r.start = a.start > b.start ? a.start : b.start;
    r.pad = 0;
    r.mask = b.start > 0 ? b.mask : (a.mask | b.mask);
    return r;
  }
};
struct MaskIn {
  const uint32_t *skey, *socc;
  const uint64_t *pmask;
  __device__ __forceinline__ MaskScan operator()(int32_t p) const {
    const bool head = p == 0 || skey[p] != skey[p - 1];
    return MaskScan{head ? p : 0, 0, (unsigned long long)pmask[socc[p]]};
  }
};
struct RowHeadPred {
  const uint32_t *skey;
  __device__ __forceinline__ bool operator()(int32_t p) const { return p == 0 || skey[p] != skey[p - 1]; }
};

// The distinct-row list is published BUCKETED BY OWNER (each bucket still sorted by feature id): an owner then
// reads one contiguous run per peer instead of scanning every peer's whole list for its rows.
//   k_owner_keys     : bucket of the u-th distinct row = its owner; G for the sentinel run and the unused tail
//   (one stable radix-sort pass over log2(G) + 1 bits gives the bucketed order `perm`)
//   k_publish_unique : one thread per slot v of the bucketed list
//   k_publish_bounds : bucket boundaries + the "distinct fields" flag of this rank, to every peer
__global__ void k_owner_keys(int32_t nnz, int G, uint32_t sentinel, const int32_t *__restrict__ uhead,
                             const int32_t *__restrict__ n_uall_p, const uint32_t *__restrict__ skey,
                             uint32_t *__restrict__ bkey, uint32_t *__restrict__ bidx) {
  const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nnz) return;
  uint32_t bk = (uint32_t)G;
  if (u < *n_uall_p) {
    const uint32_t key = skey[uhead[u]];
    if (key != sentinel) bk = key & (uint32_t)(G - 1);
  }
  bkey[u] = bk;
  bidx[u] = (uint32_t)u;
}

__global__ void k_publish_unique(int32_t nnz, int G, const uint32_t *__restrict__ bkey_s, const uint32_t *__restrict__ perm,
                                 const int32_t *__restrict__ uhead, const int32_t *__restrict__ n_uall_p,
                                 const uint32_t *__restrict__ skey, const MaskScan *__restrict__ mscan,
                                 uint32_t *__restrict__ ukey, uint32_t *__restrict__ uinfo,
                                 unsigned long long *__restrict__ umask) {
  const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nnz || bkey_s[v] >= (uint32_t)G) return;
  const int32_t u = (int32_t)perm[v], n_uall = *n_uall_p;
  const int32_t p = uhead[u];
  const int32_t next = u + 1 < n_uall ? uhead[u + 1] : nnz;
  ukey[v] = skey[p];
  uinfo[v] = (uint32_t)p | (next - p == 1 ? UINFO_SINGLE : 0u);
  umask[v] = mscan ? mscan[next - 1].mask : ~0ull;  // LR / FM: a row has no field slices, everything is touched
}

// lane r <= G: first slot of bucket r (lower bound in the sorted bucket keys)
__global__ void k_publish_bounds(Peers pr, int par, int32_t nnz, const uint32_t *__restrict__ bkey_s,
                                 const int32_t *__restrict__ batch_flags) {
  const int r = threadIdx.x;
  if (r > pr.G) return;
  int32_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (bkey_s[mid] < (uint32_t)r) lo = mid + 1; else hi = mid;
  }
  for (int q = 0; q < pr.G; q++) {
    pr.sync[q]->boff[par][pr.rank][r] = lo;
    if (r == 0) pr.sync[q]->simple[par][pr.rank] = batch_flags[0];
  }
}

// all ranks arrive; returns when every rank has reached `epoch`.  A peer that never arrives (crashed
// process, failed call) must not hang the GPU: after `timeout_cycles` the wait gives up and raises err = 4.
__global__ void k_peer_barrier(Peers pr, int channel, uint32_t epoch, long long timeout_cycles, int32_t *err) {
  const int q = threadIdx.x;
  if (q < pr.G) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&pr.sync[q]->flag[channel][pr.rank]) = epoch;
    const volatile uint32_t *mine = reinterpret_cast<const volatile uint32_t *>(&pr.sync[pr.rank]->flag[channel][q]);
    const long long t0 = clock64();
    while ((int32_t)(*mine - epoch) < 0) {
      if (clock64() - t0 > timeout_cycles) {
        *err = 4;
        break;
      }
    }
    __threadfence_system();
  }
}

// batch_flags[0] = AND over ranks (the tile path needs every rank's batch to have distinct fields)
// (LR / FM have no such requirement: need_simple == 0).  A batch that is not simple on some rank takes the generic
// kernels on EVERY rank (k_ffm_sample / k_ffm_rows, SH variants), and the owners fuse nothing.
// batch_flags[1] = "the step was called off" (k_check_abort), cleared here.
__global__ void k_merge_flags(Peers pr, int par, int need_simple, int32_t *batch_flags) {
  int all = 1;
  for (int q = 0; q < pr.G && need_simple; q++) all = all && pr.sync[pr.rank]->simple[par][q] != 0;
  batch_flags[0] = all;
  batch_flags[1] = 0;
}

// after barrier 2: some rank called the step off (k_fill_owned) -> nothing of this step may change z / n
__global__ void k_check_abort(Peers pr, int par, uint32_t step_tag, int32_t *batch_flags, int32_t *err) {
  bool any = false;
  for (int q = 0; q < pr.G; q++)
    any = any || *reinterpret_cast<const volatile uint32_t *>(&pr.sync[pr.rank]->abort_at[par][q]) == step_tag;
  if (any) {
    batch_flags[0] = 0;  // the tile kernels skip the batch ...
    batch_flags[1] = 1;  // ... and so does everything else of the step
    if (*err == 0) *err = 3;
  }
}

// ---- S2: owner side ---------------------------------------------------------------------------------
// (local row, source) pairs of the owned contributions; the tail up to `cap` is padded with the sentinel.
// More contributions than the workspace holds (ids concentrated on one residue mod G): the whole step is
// called off on EVERY rank before any z / n is touched -- the tag goes to every peer, k_check_abort reads it
// after barrier 2 and turns the remaining kernels of the step into no-ops (batch_flags[0] = 0).
// Contribution j is the (j - start_q)-th row of the run rank q published for this owner, runs in rank order.
__global__ void k_fill_owned(Peers pr, int par, int32_t cap, uint32_t local_sentinel, uint32_t step_tag,
                             int32_t *__restrict__ n_sel, uint32_t *__restrict__ okey, uint32_t *__restrict__ osrc,
                             int32_t *__restrict__ err) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cap) return;
  const SyncArea *sa = pr.sync[pr.rank];
  int32_t n = 0, q = -1, u = 0;
  for (int r = 0; r < pr.G; r++) {
    const int32_t b0 = sa->boff[par][r][pr.rank], cnt = sa->boff[par][r][pr.rank + 1] - b0;
    if (q < 0 && j < n + cnt) {
      q = r;
      u = b0 + (j - n);
    }
    n += cnt;
  }
  if (j == 0) *n_sel = n;
  if (j == 0 && n > cap) {
    *err = 3;  // owned contributions exceed the workspace (extreme skew)
    for (int q = 0; q < pr.G; q++) *reinterpret_cast<volatile uint32_t *>(&pr.sync[q]->abort_at[par][pr.rank]) = step_tag;
  }
  if (j < n) {
    okey[j] = pr.ukey[q][u] >> pr.log2G;
    osrc[j] = ((uint32_t)q << SRC_SHIFT) | (uint32_t)u;
  } else {
    okey[j] = local_sentinel;
    osrc[j] = 0;
  }
}

enum : uint8_t {
  CF_SINGLE = 1,  // the contribution is one occurrence: only the gradient plane is in the inbox
  CF_APPLY = 2,   // head of a run of contributions that k_owner_apply reduces
  CF_MAT = 4,     // head of a run whose w must be materialised before the samples run
};

// one thread per sorted contribution c (ckey sorted, contributions of one row in rank order)
// allow_fuse == 0 (LR / FM: the sample kernel finalises no row): a row touched once, by its owner only, is reduced and
// applied by the owner's row kernels like every other owner-only row
__global__ void k_contrib_class(Peers pr, int32_t cap, const int32_t *__restrict__ n_sel, uint32_t lsent, int allow_fuse,
                                const int32_t *__restrict__ batch_flags, const uint32_t *__restrict__ ckey,
                                const uint32_t *__restrict__ csrc,
                                const uint32_t *__restrict__ socc, uint8_t *__restrict__ cflag,
                                uint8_t *__restrict__ fused_sorted, int32_t *__restrict__ occ_pos) {
  const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t n = min(*n_sel, cap);
  if (c >= n) return;
  const uint32_t k = ckey[c];
  if (k == lsent) {
    cflag[c] = 0;
    return;
  }
  const bool head = c == 0 || ckey[c - 1] != k;
  const bool last = c + 1 >= n || ckey[c + 1] != k;
  const uint32_t src = csrc[c];
  const int q = (int)(src >> SRC_SHIFT);
  const uint32_t info = pr.uinfo[q][src & SRC_MASK];
  const int32_t p_head = (int32_t)(info & UINFO_POS);
  // in a batch with repeated fields one occurrence can collect several gradients per coordinate: its sum of squares
  // is not the square of its sum, so nothing is "single" (and nothing fused) there
  const bool simple = batch_flags[0] != 0;
  const bool single = (info & UINFO_SINGLE) != 0 && simple;
  if (head && last && q == pr.rank) {
    if (single && allow_fuse) {  // finalised inside its sample by k_ffm_tile
      fused_sorted[p_head] = 1;
      occ_pos[socc[p_head]] = -1;
      cflag[c] = 0;
    } else {       // reduced and applied by this rank's own row kernels
      pr.dst_at[q][p_head] = -2;
      cflag[c] = CF_MAT;
    }
    return;
  }
  pr.dst_at[q][p_head] = c;
  cflag[c] = (uint8_t)((single ? CF_SINGLE : 0) | (head ? (CF_APPLY | CF_MAT) : 0));
}

// w = W(n,z) (ffm.cpp:72-88, ftrl_model.cpp:52-59) for exactly the slices the global batch touches, and the
// linear w: run heads are compacted per block, then one warp per row.  The fresh w goes into the table AND,
// with posted stores over NVLink, into the row cache of every remote rank that touches the row -- one
// transfer per distinct (row, rank), overlapped with the table reads of the other rows.
template <bool PRECISE, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_owner_materialise(Peers pr, Dims d, Hyper h, int all_slices, int32_t cap, const int32_t *__restrict__ n_sel,
                    const int32_t *__restrict__ batch_flags, const uint32_t *__restrict__ ckey,
                    const uint32_t *__restrict__ csrc, const uint8_t *__restrict__ cflag, float *__restrict__ tab,
                    float4 *__restrict__ lin) {
  constexpr int WARPS = THREADS / 32;
  __shared__ int s_list[THREADS];
  __shared__ int s_n;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int32_t n = min(*n_sel, cap);
  const int64_t ld = d.ld, rs = 3 * ld;
  // FFM: the vectors of the field slices the batch touches (w of the other slices keeps its stored, stale value like
  // in the reference); LR / FM: the whole latent row (none for LR)
  const int vpf = all_slices ? 1 : d.k >> 2;
  const int n_vec = all_slices ? (int)(ld >> 2) : d.n_fields * vpf;
  for (int base = blockIdx.x * THREADS; base < n; base += gridDim.x * THREADS) {
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int c = base + tid;
    if (c < n && (cflag[c] & CF_MAT)) s_list[atomicAdd(&s_n, 1)] = c;
    __syncthreads();
    const int n_list = s_n;
    for (int li = wib; li < n_list; li += WARPS) {
      const int c0 = s_list[li];
      const uint32_t k = ckey[c0];
      unsigned long long m = 0ull;
      int my_q = -1, my_head = 0;  // lane j < J: the j-th contributor of the row
      if (lane < pr.G && c0 + lane < n && ckey[c0 + lane] == k) {
        const uint32_t src = csrc[c0 + lane];
        my_q = (int)(src >> SRC_SHIFT);
        m = pr.umask[my_q][src & SRC_MASK];
        my_head = (int)(pr.uinfo[my_q][src & SRC_MASK] & UINFO_POS);
      }
      const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)m);
      const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(m >> 32));
      const unsigned long long mask = all_slices ? ~0ull : (((unsigned long long)hi << 32) | lo);
      const unsigned remote = __ballot_sync(0xffffffffu, my_q >= 0 && my_q != pr.rank);
      // the remote contributors' caches, broadcast once (warp-uniform; the loop below diverges on the mask)
      float *rcw[MAX_SHARDS];
#pragma unroll
      for (int j = 0; j < MAX_SHARDS; j++) {
        const int q = __shfl_sync(0xffffffffu, my_q, j), head = __shfl_sync(0xffffffffu, my_head, j);
        rcw[j] = nullptr;
        if ((remote >> j) & 1u) rcw[j] = pr.rc_w[q] + (int64_t)head * ld;
      }
      float *row = tab + (int64_t)k * rs;
      for (int v = lane; v < n_vec; v += 32) {
        if (!all_slices && !((mask >> (v / vpf)) & 1ull)) continue;
        const float4 z = reinterpret_cast<const float4 *>(row)[v], nn = reinterpret_cast<const float4 *>(row + ld)[v];
        const float4 w = weight4<PRECISE>(z, nn, h);
        reinterpret_cast<float4 *>(row + 2 * ld)[v] = w;
#pragma unroll
        for (int j = 0; j < MAX_SHARDS; j++)
          if (rcw[j]) reinterpret_cast<float4 *>(rcw[j])[v] = w;
      }
      __syncwarp();
      float wl = 0.f;
      if (lane == 0) {
        const float4 e = lin[k];
        wl = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
        lin[k].z = wl;
      }
      wl = __shfl_sync(0xffffffffu, wl, 0);
      if (my_q >= 0 && my_q != pr.rank) pr.rc_lin[my_q][my_head] = wl;
    }
    __syncthreads();
  }
}

// ---- S4: owner side: sum the inbox entries of a row in rank order, closed-form FTRL update -------------
// work item = (run of contributions of one row, part of 32 float4 vectors); part 0 also does the linear term
template <bool PRECISE, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_owner_apply(Peers pr, Dims d, Hyper h, int32_t cap, const int32_t *__restrict__ n_sel,
              const int32_t *__restrict__ batch_flags, const uint32_t *__restrict__ ckey,
              const uint8_t *__restrict__ cflag, const float *__restrict__ inbox, const float2 *__restrict__ inbox_lin,
              float *__restrict__ tab, float4 *__restrict__ lin) {
  if (batch_flags[1] != 0) return;  // the step was called off
  constexpr int WARPS = THREADS / 32;
  __shared__ int s_list[THREADS];
  __shared__ int s_n;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int32_t n = min(*n_sel, cap);
  const int64_t ld = d.ld, rs = 3 * ld;
  const int nvec = (int)(ld >> 2);
  const int parts = max(1, (nvec + 31) >> 5);  // (LR: no latent row, part 0 still updates the linear coordinate)
  for (int base = blockIdx.x * THREADS; base < n; base += gridDim.x * THREADS) {
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int c = base + tid;
    if (c < n && (cflag[c] & CF_APPLY)) s_list[atomicAdd(&s_n, 1)] = c;
    __syncthreads();
    const int n_items = s_n * parts;
    for (int it = wib; it < n_items; it += WARPS) {
      const int c0 = s_list[it / parts], part_i = it % parts;
      const uint32_t k = ckey[c0];
      int J = 1;
      while (J < pr.G && c0 + J < n && ckey[c0 + J] == k) J++;
      const int v = part_i * 32 + lane;
      float *row = tab + (int64_t)k * rs;
      if (v < nvec) {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        for (int j = 0; j < J; j++) {
          const float4 *e = reinterpret_cast<const float4 *>(inbox + (int64_t)(c0 + j) * 2 * ld);
          const float4 g0 = __ldcs(e + v);
          float4 g1;
          if (cflag[c0 + j] & CF_SINGLE) g1 = make_float4(g0.x * g0.x, g0.y * g0.y, g0.z * g0.z, g0.w * g0.w);
          else g1 = __ldcs(e + nvec + v);
          s0.x += g0.x; s0.y += g0.y; s0.z += g0.z; s0.w += g0.w;
          s1.x += g1.x; s1.y += g1.y; s1.z += g1.z; s1.w += g1.w;
        }
        const bool any = s0.x != 0.f || s0.y != 0.f || s0.z != 0.f || s0.w != 0.f || s1.x != 0.f || s1.y != 0.f ||
                         s1.z != 0.f || s1.w != 0.f;
        if (any) {
          float4 z = reinterpret_cast<float4 *>(row)[v], nn = reinterpret_cast<float4 *>(row + ld)[v];
          const float4 w = reinterpret_cast<float4 *>(row + 2 * ld)[v];
          ftrl_apply<PRECISE>(z.x, nn.x, w.x, s0.x, s1.x, h);
          ftrl_apply<PRECISE>(z.y, nn.y, w.y, s0.y, s1.y, h);
          ftrl_apply<PRECISE>(z.z, nn.z, w.z, s0.z, s1.z, h);
          ftrl_apply<PRECISE>(z.w, nn.w, w.w, s0.w, s1.w, h);
          reinterpret_cast<float4 *>(row)[v] = z;
          reinterpret_cast<float4 *>(row + ld)[v] = nn;
        }
      }
      if (part_i == 0 && lane == 0) {
        float sg = 0.f, sg2 = 0.f;
        for (int j = 0; j < J; j++) {
          const float2 t = inbox_lin[c0 + j];
          sg += t.x;
          sg2 += t.y;
        }
        float4 e = lin[k];
        ftrl_apply<PRECISE>(e.x, e.y, e.z, sg, sg2, h);
        lin[k] = e;
      }
    }
    __syncthreads();
  }
}

// local partial sums -> every peer's SyncArea.red[rank]   (runs after the local k_batch_reduce partials)
__global__ void k_publish_red(Peers pr, const double *__restrict__ local4) {
  const int q = threadIdx.x;
  if (q >= pr.G) return;
  for (int e = 0; e < 4; e++) pr.sync[q]->red[pr.rank][e] = local4[e];
}

// bias update from the G partials, in rank order (identical on every rank)
template <bool PRECISE>
__global__ void k_bias_apply(Peers pr, Hyper h, const int32_t *__restrict__ batch_flags, float4 *__restrict__ bias) {
  if (batch_flags[1] != 0) return;  // the step was called off
  double a = 0.0, q2 = 0.0, n = 0.0;
  for (int q = 0; q < pr.G; q++) {
    a += pr.sync[pr.rank]->red[q][0];
    q2 += pr.sync[pr.rank]->red[q][1];
    n += pr.sync[pr.rank]->red[q][3];
  }
  if (n > 0.0) {
    float4 e = *bias;
    e.z = weight_from<PRECISE>(e.x, f_sqrt<PRECISE>(e.y), h);
    ftrl_apply<PRECISE>(e.x, e.y, e.z, (float)a, (float)q2, h);
    *bias = e;
  }
}

}  // namespace ftrl
