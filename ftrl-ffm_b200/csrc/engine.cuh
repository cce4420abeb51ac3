// engine.cuh -- the handle behind the C ABI: HBM tables, per-batch workspace, CSR staging slots,
// streams/events, phase timing.  Host-side C++17; the kernels live in ffm.cuh / lr_fm.cuh /
// exact.cuh / prep.cuh.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ftrl_b200.h"
#include "common.cuh"
#include "prep.cuh"
#include "shard.cuh"

namespace ftrl {

struct CudaFail {
  cudaError_t e;
  const char *what;
  const char *file;
  int line;
};

#define FTRL_CUDA(expr)                                         \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) throw CudaFail{_e, #expr, __FILE__, __LINE__}; \
  } while (0)

struct ArgFail {
  std::string msg;
};
struct IoFail {
  std::string msg;
};
struct StateFail {
  std::string msg;
};

inline std::string fmt(const char *f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return buf;
}

// Device allocations are sized in whole 2 MiB pages.  Measured on B200 (tools/exp/ipc2.cu): a cudaMalloc
// block whose size is NOT a multiple of 2 MiB is imported by cudaIpcOpenMemHandle with small pages, and
// scattered peer reads of it over NVLink drop from ~635 GB/s to ~75 GB/s (TLB misses); every buffer a
// peer rank may map therefore has to be a whole number of 2 MiB pages.
inline size_t page_round(size_t bytes) {
  constexpr size_t PAGE = size_t(2) << 20;
  return (bytes + PAGE - 1) / PAGE * PAGE;
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    release();
    if (count) FTRL_CUDA(cudaMalloc(&p, page_round(count * sizeof(T))));
    n = count;
  }
  void ensure(size_t count) {
    if (count > n) alloc(count + count / 8);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void swap(DevBuf &o) {
    T *tp = p; p = o.p; o.p = tp;
    size_t tn = n; n = o.n; o.n = tn;
  }
  ~DevBuf() { release(); }
};

template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t n = 0;
  void ensure(size_t count) {
    if (count <= n) return;
    release();
    FTRL_CUDA(cudaMallocHost(&p, (count + count / 8) * sizeof(T)));
    n = count + count / 8;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
  ~PinBuf() { release(); }
};

// one CSR staging slot of the host-pointer path (pinned/pageable host -> HBM, async)
struct Slot {
  DevBuf<int64_t> row_ptr;
  DevBuf<int32_t> field, feat, label;
  DevBuf<float> val, out;
  DevBuf<double> loss;
  PinBuf<float> h_out;
  PinBuf<double> h_loss;
  cudaEvent_t copied = nullptr, done = nullptr;
  bool busy = false;
  // deliveries performed when the slot retires (ftrl_sync or reuse)
  float *user_out = nullptr;
  double *user_loss = nullptr;
  int64_t n_out = 0;
};

struct Phase {
  const char *name;
  double ms = 0.0;
  int64_t launches = 0;
};

enum PhaseId { PH_PREP = 0, PH_SORT, PH_SEGMENT, PH_SAMPLE, PH_ROWS, PH_COMBINE, PH_REDUCE, PH_EXACT, PH_PREDICT, PH_GENERIC, PH_MATERIALISE, PH_EXCHANGE, PH_PULL, PH_APPLY, PH_COUNT };

struct PendingEvent {
  int phase;
  cudaEvent_t a, b;
};

}  // namespace ftrl

struct ftrl_handle {
  ftrl_config cfg{};
  ftrl::Dims dims{};
  ftrl::Hyper hyper{};
  std::string err;
  int n_sms = 148;

  // HBM tables
  float *tab = nullptr;     // [n_feats][3][ld]
  float4 *lin = nullptr;    // [n_feats]
  float4 *bias = nullptr;   // [1]
  uint32_t *pair_lut = nullptr;
  int32_t *d_err = nullptr;

  // streams
  cudaStream_t compute = nullptr, copy = nullptr;
  bool own_compute = true;

  // Index workspace of a batch (everything the weight-independent phase writes: sort keys, sorted list, classes,
  // chunk list, field masks).  Two sets alternate between consecutive batches: the index phase of batch i+1 runs on
  // its own stream under the forward / update kernels of batch i (train_device, FTRL_B200_PIPELINE).  The members
  // below are the CURRENT set; `alt` holds the other one (swap_idsets).
  struct IdSet {
    int64_t rows_cap = 0, nnz_cap = 0;
    ftrl::DevBuf<uint32_t> key, occ_idx, skey, socc;
    ftrl::DevBuf<int32_t> occ_row, chunk_pos, n_chunks;
    ftrl::DevBuf<uint8_t> sflags, fused_sorted;
    ftrl::DevBuf<int32_t> occ_pos, batch_flags;
    ftrl::DevBuf<uint64_t> pmask;
    ftrl::DevBuf<unsigned long long> rowmask;
    ftrl::DevBuf<ftrl::SegScan> scan;
    ftrl::DevBuf<int4> cdesc;
    // sharded runs: the sorted list of owned contributions (read by the owner-side weight kernels of the step)
    ftrl::DevBuf<uint32_t> ckey, csrc;
    ftrl::DevBuf<uint8_t> cflag;
    ftrl::DevBuf<int32_t> n_sel;
  } alt;
  int idset_cur = 0;
  cudaStream_t idstream = nullptr;
  cudaEvent_t ev_id_done[2] = {nullptr, nullptr}, ev_hot_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_id_tail = nullptr;  // end of the last index phase, whichever stream it ran on
  bool hot_recorded[2] = {false, false}, id_tail_recorded = false;
  int pipeline = 1;             // FTRL_B200_PIPELINE
  bool stable_device_inputs = false;

  // per-batch workspace (shared by consecutive batches: the compute stream is in-order)
  int64_t rows_cap = 0, nnz_cap = 0;
  ftrl::DevBuf<uint32_t> key, occ_idx, skey, socc;
  ftrl::DevBuf<int32_t> occ_row, chunk_pos, n_chunks;
  ftrl::DevBuf<uint8_t> sflags, fused_sorted, cub_tmp;
  ftrl::DevBuf<int32_t> occ_pos, batch_flags;
  ftrl::DevBuf<uint64_t> pmask;                 // per occurrence: fields of the other features of its sample
  ftrl::DevBuf<unsigned long long> rowmask;     // per segmented row: fields touched in this batch
  ftrl::PmaskSrc pmask_src{};
  ftrl::DevBuf<float> staging, staging_lin;  // per-occurrence gradient images (tile path)
  ftrl::DevBuf<ftrl::SegScan> scan;
  ftrl::DevBuf<int4> cdesc;                  // one record per chunk of the chunk list (k_chunk_desc, prep.cuh)
  ftrl::DevBuf<float> g, S, part;
  ftrl::DevBuf<float2> part_lin;
  ftrl::DevBuf<float> logit_ws;
  ftrl::DevBuf<double> red_part, loss_sum;
  ftrl::DevBuf<unsigned int> ticket;
  ftrl::DevBuf<float> xfer;  // staging for get/set rows
  size_t cub_bytes = 0;
  int32_t chunk = 32;

  static constexpr int N_SLOTS = 3;
  ftrl::Slot slots[N_SLOTS];
  int next_slot = 0;

  // measurement
  bool profiling = false;
  ftrl::Phase phases[ftrl::PH_COUNT];
  std::vector<ftrl::PendingEvent> pending;
  std::vector<cudaEvent_t> event_pool;
  ftrl_batch_stats stats{};
  int64_t launches_this_call = 0;
  int64_t last_nnz = 0;

  // feature-sharded multi-GPU (shard.cuh): this rank holds rows feat with feat % G == rank
  int G = 1, log2G = 0, rank = 0;
  int64_t n_local = 0;          // rows of lin / tab held here
  ftrl::Shards shards{};        // every shard's tables (predict)
  ftrl::RowSpace rowspace{};    // what the per-sample training kernel addresses
  ftrl::Export exportd{};       // where reduced row sums go (single GPU: applied in place)
  ftrl::Peers peers{};
  ftrl::SyncArea *sync = nullptr;
  uint32_t epoch = 0;           // barrier epochs of the weight channel (compute stream)
  uint32_t epoch_id = 0;        // ... of the index channel
  uint32_t shard_step = 0;      // sharded steps made: its parity selects the published-list / index set of a step
  long long barrier_timeout_cycles = 40000000000ll;
  bool attached = false;
  int64_t ow_cap = 0;           // owner-side capacity: (row, rank) contributions this rank may own per step
  ftrl::DevBuf<uint32_t> okey, osrc, ckey, csrc;   // owned contributions, unsorted / sorted by local row
  ftrl::DevBuf<uint8_t> cflag;
  ftrl::DevBuf<int32_t> n_sel;
  ftrl::DevBuf<uint32_t> bkey, bkey_s, bidx, perm;   // owner buckets of the distinct-row list (shard.cuh)
  ftrl::DevBuf<ftrl::MaskScan> mscan;              // segmented OR-scan of the field masks over the sorted list
  ftrl::DevBuf<int32_t> uhead, n_uall, dst_at;     // distinct rows of the local batch (sorted head positions)
  ftrl::DevBuf<uint32_t> ukey, uinfo;
  ftrl::DevBuf<unsigned long long> umask;
  ftrl::DevBuf<float> rc_w, rc_lin;                // cache of remote rows (w plane), by sorted head position
  ftrl::DevBuf<float> inbox;                       // [contribution][2][ld] (sum g, sum g^2) from the contributing ranks
  ftrl::DevBuf<float2> inbox_lin;
  ftrl::DevBuf<double> red4;
  std::vector<void *> ipc_opened;

  // tunables (env overrides for experiments)
  int fuse = 1;
  int sample_threads = 0;
  int precise = 0;  // 0: MUFU sqrt/rcp in minibatch kernels; 1: IEEE sqrt/div
  int tile = 1;     // FFM: TMA-staged per-sample kernel for batches of distinct-field samples
  bool tile_ok = false;
  int tile_ctas_per_sm = 1;
  int tile_f_cap = 0, tile_stride = 0, tile_stride1 = 0, tile_stages = 0, tile_consumers = 0, tile_ipt = 1, tile_meta = 4, tile_dbg = 0, tile_cache = 1, tile_inflight = 4;
  int tile_ring = 0;  // bytes of the row ring
  size_t tile_smem = 0;
};
