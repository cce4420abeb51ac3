// shard.cuh -- multi-GPU: feature-sharded tables, one process per GPU, peers mapped over NVLink.
//
// Row `feat` of lin / tab lives on rank feat mod G at local row feat div G; the samples
// of a global minibatch are split across ranks.  Per step, on every rank r (all on its stream):
//   S1  k_prep_rows over the local samples; k_publish tells every peer the local nnz / batch flag
//   --  barrier 1 (device-side, flags in peer memory)
//   S2  owner side: DeviceSelect over ALL ranks' key buffers (read over NVLink) keeps the occurrences
//       whose row this rank owns, in (rank, occurrence) order -> deterministic; k_fill_owned turns them
//       into (local row, source) pairs; radix sort; k_occ_class_sharded classifies every occurrence and
//       writes its class / staging position straight into the SAMPLE-side rank's occ_pos buffer
//   --  barrier 2
//   S3  k_ffm_tile over the local samples: bulk copies pull the rows from their owners, updated rows
//       and gradient images are bulk-stored back to the owners (Shards in ffm_tile.cuh); the local
//       (sum g, sum g^2) goes to every peer (k_batch_reduce with peers)
//   --  barrier 3
//   S4  owner side: k_ffm_staged_rows + k_ffm_combine on the local shard; k_bias_apply sums the G
//       partials in rank order, so the replicated bias stays bit-identical on all ranks
// The exchange is therefore not a separate all-to-all: the kernels that need remote rows load and
// store them in place through peer pointers, tile by tile.
#pragma once
#include "common.cuh"
#include "ffm_tile.cuh"
#include "prep.cuh"

namespace ftrl {

// lives in device memory of every rank, mapped by all peers
struct SyncArea {
  uint32_t flag[MAX_SHARDS];      // flag[q] = last barrier epoch rank q has reached (written by q)
  int32_t nnz[MAX_SHARDS];        // nnz[q]  = occurrences of rank q's current batch (written by q)
  int32_t simple[MAX_SHARDS];     // simple[q] != 0: every sample of rank q's batch has distinct fields
  double red[MAX_SHARDS][4];      // red[q] = {sum g, sum g^2, sum loss, n_rows} of rank q's batch
};

struct Peers {
  int G, log2G, rank, pad;
  SyncArea *sync[MAX_SHARDS];
  const uint32_t *key[MAX_SHARDS];
  int32_t *occ_pos[MAX_SHARDS];
};

__global__ void k_publish(Peers pr, int32_t nnz, const int32_t *batch_flags) {
  const int q = threadIdx.x;
  if (q >= pr.G) return;
  pr.sync[q]->nnz[pr.rank] = nnz;
  pr.sync[q]->simple[pr.rank] = batch_flags[0];
}

// all ranks arrive; returns when every rank has reached `epoch`.  A peer that never arrives (crashed
// process, failed call) must not hang the GPU: after `timeout_cycles` the wait gives up and raises err = 4.
__global__ void k_peer_barrier(Peers pr, uint32_t epoch, long long timeout_cycles, int32_t *err) {
  const int q = threadIdx.x;
  if (q < pr.G) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&pr.sync[q]->flag[pr.rank]) = epoch;
    const volatile uint32_t *mine = reinterpret_cast<const volatile uint32_t *>(&pr.sync[pr.rank]->flag[q]);
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
__global__ void k_merge_flags(Peers pr, int32_t *batch_flags, int32_t *err) {
  int all = 1;
  for (int q = 0; q < pr.G; q++) all = all && pr.sync[pr.rank]->simple[q] != 0;
  batch_flags[0] = all;
  if (!all) *err = 2;  // sharded mode has no generic fallback yet
}

struct OwnedPred {
  Peers pr;
  int32_t nnz_max;
  uint32_t sentinel;
  __device__ __forceinline__ bool operator()(int32_t idx) const {
    const int q = idx / nnz_max, t = idx - q * nnz_max;
    if (t >= pr.sync[pr.rank]->nnz[q]) return false;
    const uint32_t key = pr.key[q][t];
    return key != sentinel && (int)(key & (uint32_t)(pr.G - 1)) == pr.rank;
  }
};

// (local row, source) pairs of the owned occurrences; the tail up to `cap` is padded with the sentinel
__global__ void k_fill_owned(Peers pr, int32_t nnz_max, int32_t cap, uint32_t local_sentinel,
                             const int32_t *__restrict__ sel, const int32_t *__restrict__ n_sel,
                             uint32_t *__restrict__ okey, uint32_t *__restrict__ osrc, int32_t *__restrict__ err) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cap) return;
  const int32_t n = *n_sel;
  if (j == 0 && n > cap) *err = 3;  // owned occurrences exceed the workspace (extreme skew)
  if (j < n) {
    const int32_t idx = sel[j];
    const int q = idx / nnz_max, t = idx - q * nnz_max;
    okey[j] = pr.key[q][t] >> pr.log2G;
    osrc[j] = ((uint32_t)q << SRC_SHIFT) | (uint32_t)t;
  } else {
    okey[j] = local_sentinel;
    osrc[j] = 0;
  }
}

// sharded twin of k_occ_class: the class / staging position goes to the rank that holds the sample
__global__ void k_occ_class_sharded(Peers pr, int32_t n, uint32_t sentinel, const uint32_t *__restrict__ skey,
                                    const uint32_t *__restrict__ socc, uint8_t *__restrict__ fused_sorted) {
  const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t k = skey[p];
  if (k == sentinel) {
    fused_sorted[p] = 0;
    return;
  }
  const uint32_t src = socc[p];
  const bool head = p == 0 || skey[p - 1] != k;
  const bool last = p + 1 == n || skey[p + 1] != k;
  // finalised inside the sample only when the sample lives on the owner: remote rows always go through
  // the owner (w in, gradient image out: 8 B per coordinate over NVLink instead of 20 B)
  const bool fused = head && last && (int)(src >> SRC_SHIFT) == pr.rank;
  fused_sorted[p] = fused ? 1 : 0;
  pr.occ_pos[src >> SRC_SHIFT][src & SRC_MASK] = fused ? -1 : p;
}

// local partial sums -> every peer's SyncArea.red[rank]   (runs after the local k_batch_reduce partials)
__global__ void k_publish_red(Peers pr, const double *__restrict__ local4) {
  const int q = threadIdx.x;
  if (q >= pr.G) return;
  for (int e = 0; e < 4; e++) pr.sync[q]->red[pr.rank][e] = local4[e];
}

// bias update from the G partials, in rank order (identical on every rank)
template <bool PRECISE>
__global__ void k_bias_apply(Peers pr, Hyper h, float4 *__restrict__ bias) {
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
