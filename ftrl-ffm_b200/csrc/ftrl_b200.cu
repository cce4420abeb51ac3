// ftrl_b200.cu -- C ABI (include/ftrl_b200.h) of the B200-native FTRL LR/FM/FFM trainer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 (see build.py).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <unistd.h>

#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "engine.cuh"
#include "exact.cuh"
#include "ffm.cuh"
#include "ffm_tile.cuh"
#include "lr_fm.cuh"
#include "model_io.h"

using namespace ftrl;

static thread_local std::string g_create_error;

// ---------------------------------------------------------------------------------------------
// small kernels: init, layout conversion, zero scan
// ---------------------------------------------------------------------------------------------
// (G, rank): this shard holds global rows rank, rank + G, ...; random streams are indexed by GLOBAL row so
// the initial model does not depend on the number of shards
__global__ void k_init_tab(float *tab, int64_t n_rows, int32_t row_len, int32_t ld, float mean, float stddev,
                           uint64_t seed, int G, int rank) {
  // one thread per 4 consecutive floats of the w plane of one row
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_row = ld / 4;
  if (q >= n_rows * per_row) return;
  const int64_t row = q / per_row;
  const int v = (int)(q % per_row) * 4;
  float *base = tab + row * 3 * (int64_t)ld;
  const float4 gz = gaussian4((uint64_t)((row * G + rank) * per_row + v / 4), seed, 1u);
  const float r[4] = {gz.x, gz.y, gz.z, gz.w};
  for (int e = 0; e < 4; e++) {
    base[v + e] = 0.f;
    base[ld + v + e] = 0.f;
    base[2 * ld + v + e] = (v + e < row_len) ? fmaf(stddev, r[e], mean) : 0.f;
  }
}

__global__ void k_init_lin(float4 *lin, int64_t n, float mean, float stddev, uint64_t seed, int G, int rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 gz = gaussian4((uint64_t)(i * G + rank), seed, 2u);
  lin[i] = make_float4(0.f, 0.f, fmaf(stddev, gz.x, mean), 0.f);
}

__global__ void k_randomize(float *tab, float4 *lin, float4 *bias, int64_t n_rows, int32_t row_len, int32_t ld,
                            uint64_t seed, float z_scale, float n_lo, float n_hi, int G, int rank) {
  // one thread per 4 consecutive coordinates; slot 0 of each row additionally handles the linear record
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_row = ld > 0 ? ld / 4 : 1;
  if (q >= n_rows * per_row) return;
  const int64_t row = q / per_row;
  const int v = (int)(q % per_row) * 4;
  const int64_t grow = row * G + rank;           // global row
  const int64_t gq = grow * per_row + v / 4;     // global counter: independent of the sharding
  const float4 gz = gaussian4((uint64_t)gq, seed, 11u);
  const uint4 u = philox4x32_10(make_uint4((uint32_t)gq, (uint32_t)(gq >> 32), 12u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float kk = 2.3283064365386963e-10f;
  const float zr[4] = {gz.x, gz.y, gz.z, gz.w};
  const float ur[4] = {u.x * kk, u.y * kk, u.z * kk, u.w * kk};
  if (ld > 0) {
    float *base = tab + row * 3 * (int64_t)ld;
    for (int e = 0; e < 4; e++)
      if (v + e < row_len) {
        base[v + e] = z_scale * zr[e];
        base[ld + v + e] = n_lo + (n_hi - n_lo) * ur[e];
      }
  }
  if (v == 0) {
    const float4 g2 = gaussian4((uint64_t)grow, seed, 13u);
    float4 e = lin[row];
    e.x = z_scale * g2.x;
    e.y = n_lo + (n_hi - n_lo) * fabsf(g2.y) * 0.25f;
    lin[row] = e;
    if (row == 0) {  // replicated bias: the same value on every shard
      const float4 g0 = gaussian4(0ull, seed, 13u);
      *bias = make_float4(z_scale * 0.01f * g0.z, 0.5f * (n_lo + n_hi), 0.f, 0.f);
    }
  }
}

// dense[n_rows][row_len] <-> plane `plane` of tab rows [row0, row0+n_rows)
__global__ void k_plane_copy(float *tab, float *dense, int64_t row0, int64_t n_rows, int32_t row_len, int32_t ld,
                             int plane, int to_dense) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_rows * row_len) return;
  const int64_t r = q / row_len;
  const int c = (int)(q % row_len);
  float *t = tab + (row0 + r) * 3 * (int64_t)ld + (int64_t)plane * ld + c;
  if (to_dense) dense[q] = *t; else *t = dense[q];
}
__global__ void k_lin_copy(float4 *lin, float *dense, int64_t row0, int64_t n_rows, int plane, int to_dense) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_rows) return;
  float *e = reinterpret_cast<float *>(lin + row0 + q) + plane;
  if (to_dense) dense[q] = *e; else *e = dense[q];
}

// the same by GLOBAL row over every shard (peer memory): a rank of an attached multi-GPU run reads the whole model
__global__ void k_plane_gather(Shards sh, float *dense, int64_t row0, int64_t n_rows, int32_t row_len, int32_t ld, int plane) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_rows * row_len) return;
  const int64_t r = q / row_len;
  const int c = (int)(q % row_len);
  dense[q] = sh.row((int32_t)(row0 + r), 3 * (int64_t)ld)[(int64_t)plane * ld + c];
}
__global__ void k_lin_gather(Shards sh, float *dense, int64_t row0, int64_t n_rows, int plane) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_rows) return;
  dense[q] = reinterpret_cast<const float *>(sh.linp((int32_t)(row0 + q)))[plane];
}

__global__ void k_has_zero(const float *tab, const float4 *lin, int64_t n_rows, int32_t row_len, int32_t ld,
                           int32_t *flag) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)row_len + 1;
  if (q >= n_rows * per) return;
  const int64_t r = q / per;
  const int c = (int)(q % per);
  const float w = c == 0 ? lin[r].z : tab[r * 3 * (int64_t)ld + 2 * (int64_t)ld + (c - 1)];
  if (w == 0.0f) *flag = 1;
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
static inline int plane_of(int which) { return which == 0 ? PLANE_W : which == 1 ? PLANE_N : PLANE_Z; }

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static cudaEvent_t get_event(ftrl_handle *h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  FTRL_CUDA(cudaEventCreate(&e));
  return e;
}

struct PhaseScope {
  ftrl_handle *h;
  int id;
  cudaEvent_t a = nullptr;
  PhaseScope(ftrl_handle *h_, int id_) : h(h_), id(id_) {
    if (h->profiling) {
      a = get_event(h);
      FTRL_CUDA(cudaEventRecord(a, h->compute));
    }
  }
  ~PhaseScope() {
    if (h->profiling && a) {
      cudaEvent_t b = get_event(h);
      cudaEventRecord(b, h->compute);
      h->pending.push_back({id, a, b});
    }
  }
};

static void launched(ftrl_handle *h, int phase, int n = 1) {
  h->phases[phase].launches += n;
  h->launches_this_call += n;
}

static void drain_profile(ftrl_handle *h) {
  for (auto &pe : h->pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(pe.b) == cudaSuccess && cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess)
      h->phases[pe.phase].ms += ms;
    h->event_pool.push_back(pe.a);
    h->event_pool.push_back(pe.b);
  }
  h->pending.clear();
}

static size_t cub_temp_bytes(int64_t nnz, int end_bit) {
  size_t a = 0, b = 0, c = 0;
  const int n = (int)nnz;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, n, 0, end_bit);
  thrust::counting_iterator<int32_t> cnt(0);
  auto it = thrust::make_transform_iterator(cnt, HeadFunctor{nullptr, nullptr});
  cub::DeviceScan::InclusiveScan(nullptr, b, it, (SegScan *)nullptr, SegScanOp(), n);
  size_t b2 = 0;
  auto mit = thrust::make_transform_iterator(cnt, MaskIn{nullptr, nullptr, nullptr});
  cub::DeviceScan::InclusiveScan(nullptr, b2, mit, (MaskScan *)nullptr, MaskScanOp(), n);
  b = std::max(b, b2);
  cub::DeviceSelect::If(nullptr, c, cnt, (int32_t *)nullptr, (int32_t *)nullptr, n, ChunkHeadPred{nullptr, nullptr, nullptr, 0, 1});
  return std::max(a, std::max(b, c)) + 256;
}

static int key_bits(int32_t n_feats) {
  int bits = 1;
  while (bits < 32 && (1ll << bits) <= (int64_t)n_feats) bits++;
  return bits;
}

// single-GPU handles: the one shard is this handle itself (workspace buffers may have been reallocated)
static void refresh_shards(ftrl_handle *h) {
  if (h->G > 1 && h->attached) return;
  Shards &sh = h->shards;
  sh.G = 1;
  sh.log2G = 0;
  sh.rank = 0;
  sh.tab[0] = h->tab;
  sh.lin[0] = h->lin;
  h->pmask_src.p[0] = h->pmask.p;
  RowSpace &rsp = h->rowspace;
  rsp = RowSpace{};
  rsp.tab = h->tab;
  rsp.lin = h->lin;
  rsp.staging = h->staging.p;
  rsp.staging_lin = h->staging_lin.p;
  h->exportd = Export{};
}

// the other index-workspace set becomes the current one (its capacity travels with it)
static void swap_idsets(ftrl_handle *h) {
  ftrl_handle::IdSet &a = h->alt;
  std::swap(h->rows_cap, a.rows_cap);
  std::swap(h->nnz_cap, a.nnz_cap);
  h->key.swap(a.key); h->occ_idx.swap(a.occ_idx); h->skey.swap(a.skey); h->socc.swap(a.socc);
  h->occ_row.swap(a.occ_row); h->chunk_pos.swap(a.chunk_pos); h->n_chunks.swap(a.n_chunks);
  h->sflags.swap(a.sflags); h->fused_sorted.swap(a.fused_sorted);
  h->occ_pos.swap(a.occ_pos); h->batch_flags.swap(a.batch_flags);
  h->pmask.swap(a.pmask); h->rowmask.swap(a.rowmask); h->scan.swap(a.scan); h->cdesc.swap(a.cdesc);
  h->ckey.swap(a.ckey); h->csrc.swap(a.csrc); h->cflag.swap(a.cflag); h->n_sel.swap(a.n_sel);
  h->idset_cur ^= 1;
}

static void ensure_workspace(ftrl_handle *h, int64_t n_rows, int64_t nnz) {
  if (n_rows <= h->rows_cap && nnz <= h->nnz_cap) return;
  FTRL_CUDA(cudaStreamSynchronize(h->compute));
  if (h->idstream) FTRL_CUDA(cudaStreamSynchronize(h->idstream));
  const int64_t rc = std::max<int64_t>(h->rows_cap, n_rows + n_rows / 8 + 16);
  const int64_t nc = std::max<int64_t>(h->nnz_cap, nnz + nnz / 8 + 64);
  if (nc >= (1ll << 31) - 64) throw ArgFail{"batch nnz must be < 2^31"};
  h->g.ensure(rc);
  h->logit_ws.ensure(rc);
  if (!h->red_part.p) {
    h->red_part.alloc(3 * RED_MAX_CTAS);
    h->ticket.alloc(1);
    FTRL_CUDA(cudaMemset(h->ticket.p, 0, sizeof(unsigned)));
  }
  h->sflags.ensure(rc);
  if (h->dims.model_type == FTRL_FM) h->S.ensure(rc * h->dims.k);
  h->key.ensure(nc);
  h->occ_idx.ensure(nc);
  h->skey.ensure(nc);
  h->socc.ensure(nc);
  h->occ_row.ensure(nc);
  // owner side of a sharded run: the (row, rank) contributions this rank may own per step -- up to 2x the
  // distinct rows of its own batch (err 3 beyond that: extreme id skew)
  const int64_t oc = h->G > 1 ? 2 * nc + 1024 : nc;
  if (h->G > 1 && h->attached) throw ArgFail{"batch exceeds max_batch_rows / max_batch_nnz of a multi-GPU handle"};
  h->ow_cap = oc;
  if (h->G > 1) {
    h->okey.ensure(oc);
    h->osrc.ensure(oc);
    h->ckey.ensure(oc);
    h->csrc.ensure(oc);
    h->cflag.ensure(oc);
    h->n_sel.ensure(4);
    h->bkey.ensure(nc);
    h->bkey_s.ensure(nc);
    h->bidx.ensure(nc);
    h->perm.ensure(nc);
    h->red4.ensure(4);
    h->mscan.ensure(nc);
    h->uhead.ensure(nc + 1);
    h->n_uall.ensure(4);
    // published to the peers by the index phase, two sets (parity of the step): [par * nnz_cap + i]
    h->dst_at.ensure(2 * nc);
    h->ukey.ensure(2 * nc);
    h->uinfo.ensure(2 * nc);
    h->umask.ensure(2 * nc);
    // (LR has no latent row: the buffers the peers map still have to exist)
    h->rc_w.ensure(std::max<size_t>(64, (size_t)nc * h->dims.ld));
    h->rc_lin.ensure(nc);
    h->inbox.ensure(std::max<size_t>(64, (size_t)oc * 2 * h->dims.ld));
    h->inbox_lin.ensure(oc);
  }
  h->fused_sorted.ensure(nc);
  h->occ_pos.ensure(nc);
  h->batch_flags.ensure(4);
  if (h->tile_ok) {
    h->pmask.ensure(nc);
    h->rowmask.ensure(nc + 2);
  }
  if (h->tile_ok) {
    h->staging.ensure((size_t)nc * h->dims.ld);
    h->staging_lin.ensure(nc);
  }
  h->scan.ensure(nc);
  h->chunk_pos.ensure(nc + 2);
  if (h->dims.model_type == FTRL_FFM) h->cdesc.ensure(nc + 2);
  h->n_chunks.ensure(4);
  const int64_t slots = 2 * (nc / h->chunk + 2);
  if (h->dims.row_len) h->part.ensure((size_t)slots * 2 * h->dims.ld);
  h->part_lin.ensure(slots);
  h->cub_bytes = cub_temp_bytes(h->G > 1 ? std::max<int64_t>(nc, oc) : nc, key_bits(h->dims.n_feats));
  h->cub_tmp.ensure(h->cub_bytes);
  h->rows_cap = rc;
  h->nnz_cap = nc;
  refresh_shards(h);
}

template <typename F>
static int guarded(ftrl_handle *h, F &&f) {
  try {
    if (h) FTRL_CUDA(cudaSetDevice(h->cfg.device));
    f();
    return FTRL_OK;
  } catch (const CudaFail &e) {
    std::string m = fmt("CUDA error %d (%s) at %s:%d: %s", (int)e.e, cudaGetErrorString(e.e), e.file, e.line, e.what);
    if (h) h->err = m; else g_create_error = m;
    cudaGetLastError();
    return FTRL_ERR_CUDA;
  } catch (const ArgFail &e) {
    if (h) h->err = e.msg; else g_create_error = e.msg;
    return FTRL_ERR_ARG;
  } catch (const IoFail &e) {
    if (h) h->err = e.msg; else g_create_error = e.msg;
    return FTRL_ERR_IO;
  } catch (const StateFail &e) {
    if (h) h->err = e.msg; else g_create_error = e.msg;
    return FTRL_ERR_STATE;
  } catch (const std::exception &e) {
    if (h) h->err = e.what(); else g_create_error = e.what();
    return FTRL_ERR_ARG;
  }
}

static int reduce_grid(int64_t n_rows) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(RED_MAX_CTAS, (n_rows + 1023) / 1024));
}

static int pick_vec(int k) { return k % 4 == 0 ? 4 : k % 2 == 0 ? 2 : 1; }

// ---------------------------------------------------------------------------------------------
// batch training, device-resident CSR
// ---------------------------------------------------------------------------------------------
template <int VEC, bool PRECISE, bool SH = false>
static void launch_ffm_sample(ftrl_handle *h, const Batch &b, float *logit_out, int skip_if_simple) {
  const Dims &d = h->dims;
  const double fbar = b.n_rows ? (double)b.nnz / (double)b.n_rows : 0.0;
  const double items = fbar * (fbar - 1) * 0.5 * (d.k / VEC);
  int threads = h->sample_threads ? h->sample_threads : items <= 64 ? 64 : items <= 256 ? 128 : items <= 1024 ? 256 : 512;
  // sharded runs: ONE instantiation, the one the dry run of ftrl_attach_peers has loaded (a first-time kernel load is
  // deferred while the spinning barrier kernels of the step are running)
  if (SH) threads = 256;
  const dim3 grid((unsigned)((b.n_rows + FFM_SPB - 1) / FFM_SPB));
  const ItemDecode dec = make_item_decode(d.k, VEC);
#define FFM_SAMPLE(T)                                                                                          \
  k_ffm_sample<VEC, PRECISE, T, SH><<<grid, T, 0, h->compute>>>(b, d, h->hyper, dec, h->tab, h->lin, h->bias, h->pair_lut, \
                                                               h->occ_pos.p, h->fuse, h->batch_flags.p, skip_if_simple, \
                                                               h->g.p, logit_out, h->rowspace, h->scan.p)
  if (threads <= 64) FFM_SAMPLE(64);
  else if (threads <= 128) FFM_SAMPLE(128);
  else if (threads <= 256) FFM_SAMPLE(256);
  else FFM_SAMPLE(512);
#undef FFM_SAMPLE
  FTRL_CUDA(cudaGetLastError());
}

// the sample kernel and the row kernel of the tile path (ffm_tile.cuh), shared by the single-GPU and the sharded step
template <bool PRECISE>
static void launch_tile(ftrl_handle *h, const Batch &b, const ItemDecode &dec, float *logit_out) {
  TileGeom geo;
  geo.f_cap = h->tile_f_cap;
  geo.stride = h->tile_stride;
  geo.stride1 = h->tile_stride1;
  geo.inflight = h->tile_inflight;
  geo.n_stage = h->tile_stages;
  geo.ring_bytes = h->tile_ring;
  geo.n_meta = h->tile_meta;
  geo.consumers = h->tile_consumers;
  geo.dbg = h->tile_dbg;
  geo.smem_bytes = h->tile_smem;
  const int tgrid = (int)std::min<int64_t>(b.n_rows, (int64_t)h->n_sms * h->tile_ctas_per_sm);
#define FFM_TILE(I)                                                                                              \
  if (h->tile_cache)                                                                                              \
    k_ffm_tile<PRECISE, I, true><<<tgrid, tile_threads(geo.consumers), geo.smem_bytes, h->compute>>>(             \
        b, h->dims, h->hyper, dec, geo, h->batch_flags.p, h->rowspace, h->bias, h->pair_lut, h->occ_pos.p, h->scan.p, h->g.p, logit_out); \
  else                                                                                                            \
    k_ffm_tile<PRECISE, I, false><<<tgrid, tile_threads(geo.consumers), geo.smem_bytes, h->compute>>>(            \
        b, h->dims, h->hyper, dec, geo, h->batch_flags.p, h->rowspace, h->bias, h->pair_lut, h->occ_pos.p, h->scan.p, h->g.p, logit_out)
  if (h->tile_ipt <= 1) FFM_TILE(1);
  else if (h->tile_ipt == 2) FFM_TILE(2);
  else if (h->tile_ipt == 3) FFM_TILE(3);
  else FFM_TILE(4);
#undef FFM_TILE
  FTRL_CUDA(cudaGetLastError());
  launched(h, PH_SAMPLE);
}

// streaming segmented reduction of the staged gradient images (local duplicates); each row's sum is applied, parked
// for k_ffm_combine, or (sharded runs) exported to its owner's inbox
template <bool PRECISE>
static void launch_staged_rows(ftrl_handle *h, const Batch &b) {
  const int grid = h->n_sms * 16;
  k_ffm_staged_rows<PRECISE, 8><<<grid, 256, 0, h->compute>>>(h->dims, h->hyper, h->batch_flags.p, h->tab, h->lin, h->n_chunks.p,
                                                             h->cdesc.p, h->staging.p, h->staging_lin.p, h->part.p,
                                                             h->part_lin.p, h->exportd);
  FTRL_CUDA(cudaGetLastError());
  launched(h, PH_ROWS);
}

// FFM minibatch: when every sample of the batch has distinct fields (device-side flag) the tile kernels
// the generic row kernel (any samples; SH: sharded runs, rows resolved through RowSpace, sums through Export)
template <int VEC, bool PRECISE, bool SH>
static void launch_ffm_rows(ftrl_handle *h, const Batch &b, const ItemDecode &dec, int skip_if_simple) {
  const Dims &d = h->dims;
  const int grid = h->n_sms * 4;
  const size_t smem8 = (size_t)8 * 2 * d.ld * sizeof(float);
  if (smem8 <= 160 * 1024) {
    auto kern = k_ffm_rows<VEC, PRECISE, 8, SH>;
    if (smem8 > 48 * 1024) FTRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
    kern<<<grid, 256, smem8, h->compute>>>(b, d, h->hyper, dec, h->tab, h->lin, h->chunk, h->n_chunks.p, h->chunk_pos.p,
                                           h->skey.p, h->socc.p, h->scan.p, h->occ_row.p, h->sflags.p,
                                           h->batch_flags.p, skip_if_simple, h->g.p, h->part.p, h->part_lin.p, h->rowspace,
                                           h->exportd, h->occ_pos.p);
  } else {
    const size_t smem1 = (size_t)2 * d.ld * sizeof(float);
    if (smem1 > 200 * 1024) throw ArgFail{"n_fields*n_factors too large for the row kernel"};
    auto kern = k_ffm_rows<VEC, PRECISE, 1, SH>;
    FTRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    kern<<<grid * 4, 32, smem1, h->compute>>>(b, d, h->hyper, dec, h->tab, h->lin, h->chunk, h->n_chunks.p,
                                              h->chunk_pos.p, h->skey.p, h->socc.p, h->scan.p, h->occ_row.p,
                                              h->sflags.p, h->batch_flags.p, skip_if_simple, h->g.p, h->part.p,
                                              h->part_lin.p, h->rowspace, h->exportd, h->occ_pos.p);
  }
  FTRL_CUDA(cudaGetLastError());
}

// (ffm_tile.cuh) process it; otherwise the generic LDG kernels do.  Both sets are enqueued, the one that
// does not apply returns at once -- no host round trip.
template <int VEC, bool PRECISE>
static void run_ffm_batch(ftrl_handle *h, const Batch &b, float *logit_out) {
  const Dims &d = h->dims;
  const bool tile = VEC == 4 && h->tile_ok && b.nnz > 0;
  const int grid = h->n_sms * 4;
  const ItemDecode dec = make_item_decode(d.k, VEC);
  if (tile) {
    {
      PhaseScope ps(h, PH_SAMPLE);
      launch_tile<PRECISE>(h, b, dec, logit_out);
    }
    {
      PhaseScope ps(h, PH_ROWS);
      launch_staged_rows<PRECISE>(h, b);
    }
  }
  {
    PhaseScope ps(h, PH_GENERIC);
    launch_ffm_sample<VEC, PRECISE>(h, b, logit_out, tile ? 1 : 0);
    launched(h, PH_GENERIC);
    if (b.nnz > 0) {
      launch_ffm_rows<VEC, PRECISE, false>(h, b, dec, tile ? 1 : 0);
      launched(h, PH_GENERIC);
    }
  }
  if (b.nnz == 0) return;
  {
    PhaseScope ps(h, PH_COMBINE);
    k_ffm_combine<PRECISE, 256><<<grid, 256, 0, h->compute>>>(d, h->hyper, h->tab, h->lin, h->n_chunks.p, h->cdesc.p, h->part.p,
                                                              h->part_lin.p, h->exportd, h->batch_flags.p);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_COMBINE);
  }
}

template <int VEC, bool PRECISE, bool IS_FM>
static void run_lrfm_batch(ftrl_handle *h, const Batch &b, float *logit_out) {
  const Dims &d = h->dims;
  {
    PhaseScope ps(h, PH_SAMPLE);
    const int64_t warps = b.n_rows;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    k_lrfm_sample<VEC, PRECISE, IS_FM><<<grid, 256, 0, h->compute>>>(b, d, h->hyper, h->tab, h->lin, h->bias, h->S.p,
                                                                     h->g.p, logit_out);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_SAMPLE);
  }
  if (b.nnz == 0) return;
  const int grid = h->n_sms * 4;
  {
    PhaseScope ps(h, PH_ROWS);
    k_lrfm_rows<PRECISE, IS_FM, 8, false><<<grid, 256, 0, h->compute>>>(b, d, h->hyper, h->tab, h->lin, h->chunk, h->n_chunks.p,
                                                                        h->chunk_pos.p, h->skey.p, h->socc.p, h->scan.p,
                                                                        h->occ_row.p, h->g.p, h->S.p, h->part.p, h->part_lin.p,
                                                                        RowSpace{}, Export{}, h->batch_flags.p);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_ROWS);
  }
  {
    PhaseScope ps(h, PH_COMBINE);
    k_lrfm_combine<PRECISE, IS_FM, 8, false><<<grid, 256, 0, h->compute>>>(d, h->hyper, (int32_t)b.nnz, h->tab, h->lin, h->chunk,
                                                                           h->n_chunks.p, h->chunk_pos.p, h->skey.p, h->scan.p,
                                                                           h->part.p, h->part_lin.p, Export{}, h->batch_flags.p);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_COMBINE);
  }
}

// sharded LR / FM: sample kernel over the materialised w (shard rows + row cache), row kernels with Export
template <int VEC, bool PRECISE, bool IS_FM>
static void run_lrfm_batch_sharded(ftrl_handle *h, const Batch &b, float *logit_out) {
  const Dims &d = h->dims;
  if (b.n_rows > 0) {
    PhaseScope ps(h, PH_SAMPLE);
    const unsigned grid = (unsigned)((b.n_rows * 32 + 255) / 256);
    k_lrfm_sample_sh<VEC, PRECISE, IS_FM><<<grid, 256, 0, h->compute>>>(b, d, h->hyper, h->rowspace, h->occ_pos.p, h->scan.p,
                                                                        h->bias, h->S.p, h->g.p, logit_out);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_SAMPLE);
  }
}
template <bool PRECISE, bool IS_FM>
static void run_lrfm_rows_sharded(ftrl_handle *h, const Batch &b) {
  const Dims &d = h->dims;
  if (b.nnz == 0) return;
  const int grid = h->n_sms * 4;
  {
    PhaseScope ps(h, PH_ROWS);
    k_lrfm_rows<PRECISE, IS_FM, 8, true><<<grid, 256, 0, h->compute>>>(b, d, h->hyper, h->tab, h->lin, h->chunk, h->n_chunks.p,
                                                                       h->chunk_pos.p, h->skey.p, h->socc.p, h->scan.p,
                                                                       h->occ_row.p, h->g.p, h->S.p, h->part.p, h->part_lin.p,
                                                                       h->rowspace, h->exportd, h->batch_flags.p);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_ROWS);
  }
  {
    PhaseScope ps(h, PH_COMBINE);
    k_lrfm_combine<PRECISE, IS_FM, 8, true><<<grid, 256, 0, h->compute>>>(d, h->hyper, (int32_t)b.nnz, h->tab, h->lin, h->chunk,
                                                                          h->n_chunks.p, h->chunk_pos.p, h->skey.p, h->scan.p,
                                                                          h->part.p, h->part_lin.p, h->exportd, h->batch_flags.p);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_COMBINE);
  }
}

// owner-side pre-pass of the tile path: materialise w of the segmented rows (ffm_tile.cuh)
// the two halves of the owner-side pre-pass: which slices of the staged rows does the batch touch (ids only) ...
static void run_row_touch(ftrl_handle *h, int32_t n_sorted, uint32_t sentinel) {
  PhaseScope ps(h, PH_MATERIALISE);
  const int grid = h->n_sms * 8;  // 64 resident warps per SM: the kernel is a chain of dependent gathers
  FTRL_CUDA(cudaMemsetAsync(h->rowmask.p, 0, sizeof(unsigned long long) * (size_t)(n_sorted + 2), h->compute));
  k_row_touch<8><<<grid, 256, 0, h->compute>>>(sentinel, h->chunk, h->batch_flags.p, h->n_chunks.p, h->cdesc.p, h->socc.p,
                                               h->pmask_src, h->rowmask.p);
  FTRL_CUDA(cudaGetLastError());
  launched(h, PH_MATERIALISE);
}
// ... and w = W(n, z) of exactly those slices (needs the weights of the previous step)
static void run_row_materialise(ftrl_handle *h, int32_t n_sorted, uint32_t sentinel) {
  PhaseScope ps(h, PH_MATERIALISE);
  const Dims &d = h->dims;
  const int grid = h->n_sms * 8;
  (void)n_sorted;
  if (h->precise)
    k_row_materialise<true, 8><<<grid, 256, 0, h->compute>>>(d, h->hyper, sentinel, h->batch_flags.p, h->n_chunks.p, h->cdesc.p,
                                                             h->rowmask.p, h->tab, h->lin);
  else
    k_row_materialise<false, 8><<<grid, 256, 0, h->compute>>>(d, h->hyper, sentinel, h->batch_flags.p, h->n_chunks.p, h->cdesc.p,
                                                              h->rowmask.p, h->tab, h->lin);
  FTRL_CUDA(cudaGetLastError());
  launched(h, PH_MATERIALISE);
}

static void run_prep(ftrl_handle *h, const Batch &b) {
  const Dims &d = h->dims;
  const int32_t nnz = (int32_t)b.nnz;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  {
    PhaseScope ps(h, PH_PREP);
    FTRL_CUDA(cudaMemsetAsync(h->batch_flags.p, 0x01, sizeof(int32_t), h->compute));  // != 0: all samples simple
    FTRL_CUDA(cudaMemsetAsync(h->batch_flags.p + 1, 0, sizeof(int32_t), h->compute));  // [1]: step called off (sharded)
    const unsigned grid = (unsigned)((b.n_rows * 32 + 255) / 256);
    k_prep_rows<<<grid, 256, 0, h->compute>>>(b, d, h->key.p, h->occ_idx.p, h->occ_row.p, h->sflags.p,
                                              h->batch_flags.p, h->tile_ok ? h->pmask.p : nullptr);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_PREP);
  }
  if (nnz == 0) {
    FTRL_CUDA(cudaMemsetAsync(h->n_chunks.p, 0, sizeof(int32_t), h->compute));
    return;
  }
  {
    PhaseScope ps(h, PH_SORT);
    size_t bytes = h->cub_bytes;
    FTRL_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, bytes, h->key.p, h->skey.p, h->occ_idx.p, h->socc.p, nnz, 0,
                                              key_bits(d.n_feats), h->compute));
  }
  {
    PhaseScope ps(h, PH_SEGMENT);
    const int fuse = (d.model_type == FTRL_FFM && h->fuse) ? 1 : 0;
    k_occ_class<<<(nnz + 255) / 256, 256, 0, h->compute>>>(nnz, sentinel, fuse, h->skey.p, h->socc.p, h->occ_row.p,
                                                           h->sflags.p, h->fused_sorted.p, h->occ_pos.p);
    launched(h, PH_SEGMENT);
    size_t bytes = h->cub_bytes;
    thrust::counting_iterator<int32_t> cnt(0);
    auto it = thrust::make_transform_iterator(cnt, HeadFunctor{h->skey.p, h->fused_sorted.p});
    FTRL_CUDA(cub::DeviceScan::InclusiveScan(h->cub_tmp.p, bytes, it, h->scan.p, SegScanOp(), nnz, h->compute));
    bytes = h->cub_bytes;
    FTRL_CUDA(cub::DeviceSelect::If(h->cub_tmp.p, bytes, cnt, h->chunk_pos.p, h->n_chunks.p, nnz,
                                    ChunkHeadPred{h->skey.p, h->scan.p, h->fused_sorted.p, sentinel, h->chunk},
                                    h->compute));
    k_terminate<<<1, 1, 0, h->compute>>>(h->chunk_pos.p, h->n_chunks.p, nnz);
    if (d.model_type == FTRL_FFM)
      k_chunk_desc<8><<<h->n_sms * 8, 256, 0, h->compute>>>(nnz, sentinel, h->chunk, h->n_chunks.p, h->chunk_pos.p, h->skey.p,
                                                           h->scan.p, h->cdesc.p);
    launched(h, PH_SEGMENT, 2);
    FTRL_CUDA(cudaGetLastError());
  }
  if (d.model_type == FTRL_FFM && h->tile_ok) run_row_touch(h, nnz, sentinel);
}

template <bool PRECISE>
static void run_model(ftrl_handle *h, const Batch &b, float *logit_out) {
  const Dims &d = h->dims;
  const int vec = pick_vec(d.k);
  if (d.model_type == FTRL_FFM) {
    if (vec == 4) run_ffm_batch<4, PRECISE>(h, b, logit_out);
    else if (vec == 2) run_ffm_batch<2, PRECISE>(h, b, logit_out);
    else run_ffm_batch<1, PRECISE>(h, b, logit_out);
  } else if (d.model_type == FTRL_FM) {
    if (vec == 4) run_lrfm_batch<4, PRECISE, true>(h, b, logit_out);
    else if (vec == 2) run_lrfm_batch<2, PRECISE, true>(h, b, logit_out);
    else run_lrfm_batch<1, PRECISE, true>(h, b, logit_out);
  } else {
    run_lrfm_batch<1, PRECISE, false>(h, b, logit_out);
  }
}

template <bool PRECISE>
static void train_device_sharded(ftrl_handle *h, const Batch &b, float *logit_out, double *loss_sum_out,
                                 cudaEvent_t inputs_ready, bool allow_pipe);

// inputs_ready: event after which the CSR arrays of `b` may be read (host path: the slot's copy), or null: the arrays
// are ready in the order of the compute stream.  With an event (or stable device inputs) the weight-independent
// index phase of this batch runs on its own stream, under the forward / update kernels of the previous batch.
static void train_device(ftrl_handle *h, const Batch &b, float *logit_out, double *loss_sum_out,
                         cudaEvent_t inputs_ready = nullptr) {
  h->launches_this_call = 0;
  h->stats = ftrl_batch_stats{};
  h->stats.n_rows = b.n_rows;
  h->last_nnz = b.nnz;
  if (b.n_rows <= 0 && h->G == 1) {
    if (loss_sum_out) FTRL_CUDA(cudaMemsetAsync(loss_sum_out, 0, sizeof(double), h->compute));
    return;
  }
  if (h->G > 1) {
    if (h->precise) train_device_sharded<true>(h, b, logit_out, loss_sum_out, inputs_ready, true);
    else train_device_sharded<false>(h, b, logit_out, loss_sum_out, inputs_ready, true);
    return;
  }
  const bool piped = h->pipeline && h->cfg.mode == FTRL_MODE_BATCH && !h->profiling && h->idstream &&
                     (inputs_ready || h->stable_device_inputs);
  if (piped) swap_idsets(h);
  ensure_workspace(h, b.n_rows, b.nnz);
  if (piped) refresh_shards(h);
  const Dims &d = h->dims;
  if (h->cfg.mode == FTRL_MODE_SEQUENTIAL) {
    if (inputs_ready) FTRL_CUDA(cudaStreamWaitEvent(h->compute, inputs_ready, 0));
    PhaseScope ps(h, PH_EXACT);
    k_exact_train<<<1, EX_THREADS, 0, h->compute>>>(b, d, h->hyper, h->tab, h->lin, h->bias, logit_out, loss_sum_out,
                                                    h->d_err);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_EXACT);
    h->stats.kernel_launches = h->launches_this_call;
    return;
  }
  if (piped) {
    const int cur = h->idset_cur;
    cudaStream_t main_stream = h->compute;
    if (inputs_ready) FTRL_CUDA(cudaStreamWaitEvent(h->idstream, inputs_ready, 0));
    // this index set was last read by the forward / update kernels two batches ago
    if (h->hot_recorded[cur]) FTRL_CUDA(cudaStreamWaitEvent(h->idstream, h->ev_hot_done[cur], 0));
    h->compute = h->idstream;  // everything run_prep enqueues goes to the index stream
    try {
      run_prep(h, b);
    } catch (...) {
      h->compute = main_stream;
      throw;
    }
    h->compute = main_stream;
    FTRL_CUDA(cudaEventRecord(h->ev_id_done[cur], h->idstream));
    FTRL_CUDA(cudaStreamWaitEvent(h->compute, h->ev_id_done[cur], 0));
  } else {
    if (inputs_ready) FTRL_CUDA(cudaStreamWaitEvent(h->compute, inputs_ready, 0));
    run_prep(h, b);
  }
  if (d.model_type == FTRL_FFM && h->tile_ok && b.nnz > 0) run_row_materialise(h, (int32_t)b.nnz, (uint32_t)d.n_feats);
  if (!logit_out) logit_out = h->logit_ws.p;
  const bool pr = h->precise != 0;
  if (pr) run_model<true>(h, b, logit_out); else run_model<false>(h, b, logit_out);
  {
    PhaseScope ps(h, PH_REDUCE);
    const int rg = reduce_grid(b.n_rows);
    if (pr) k_batch_reduce<true><<<rg, 256, 0, h->compute>>>(b.n_rows, h->hyper, h->g.p, logit_out, b.label, h->bias, 1, h->red_part.p, h->ticket.p, loss_sum_out, nullptr);
    else k_batch_reduce<false><<<rg, 256, 0, h->compute>>>(b.n_rows, h->hyper, h->g.p, logit_out, b.label, h->bias, 1, h->red_part.p, h->ticket.p, loss_sum_out, nullptr);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_REDUCE);
  }
  if (h->idstream) {  // (also after an unpipelined call: it read the current index set on the compute stream)
    FTRL_CUDA(cudaEventRecord(h->ev_hot_done[h->idset_cur], h->compute));
    h->hot_recorded[h->idset_cur] = true;
  }
  h->stats.kernel_launches = h->launches_this_call;
}

static void predict_device(ftrl_handle *h, const Batch &b, int output_prob, float *out, double *loss_sum_out) {
  if (b.n_rows <= 0) {
    if (loss_sum_out) FTRL_CUDA(cudaMemsetAsync(loss_sum_out, 0, sizeof(double), h->compute));
    return;
  }
  ensure_workspace(h, b.n_rows, 0);
  const Dims &d = h->dims;
  // the loss needs raw logits: when probabilities are requested they go to the workspace
  float *lg = loss_sum_out ? (output_prob ? h->logit_ws.p : out) : nullptr;
  float *lg_side = (loss_sum_out && output_prob) ? h->logit_ws.p : nullptr;
  PhaseScope ps(h, PH_PREDICT);
  if (h->cfg.mode == FTRL_MODE_SEQUENTIAL) {
    k_exact_predict<<<(unsigned)((b.n_rows + 127) / 128), 128, 0, h->compute>>>(b, d, h->tab, h->lin, h->bias, output_prob,
                                                                               out, lg_side);
  } else if (d.model_type == FTRL_FFM) {
    const int vec = pick_vec(d.k);
    const dim3 grid((unsigned)b.n_rows);
    if (vec == 4) k_ffm_predict<4, 256><<<grid, 256, 0, h->compute>>>(b, d, make_item_decode(d.k, 4), h->shards, h->bias, h->pair_lut, output_prob, out, lg_side);
    else if (vec == 2) k_ffm_predict<2, 256><<<grid, 256, 0, h->compute>>>(b, d, make_item_decode(d.k, 2), h->shards, h->bias, h->pair_lut, output_prob, out, lg_side);
    else k_ffm_predict<1, 256><<<grid, 256, 0, h->compute>>>(b, d, make_item_decode(d.k, 1), h->shards, h->bias, h->pair_lut, output_prob, out, lg_side);
  } else {
    const unsigned grid = (unsigned)((b.n_rows * 32 + 255) / 256);
    if (d.model_type == FTRL_FM) k_lrfm_predict<true><<<grid, 256, 0, h->compute>>>(b, d, h->shards, h->bias, output_prob, out, lg_side);
    else k_lrfm_predict<false><<<grid, 256, 0, h->compute>>>(b, d, h->shards, h->bias, output_prob, out, lg_side);
  }
  FTRL_CUDA(cudaGetLastError());
  launched(h, PH_PREDICT);
  if (loss_sum_out) {
    k_batch_reduce<true><<<reduce_grid(b.n_rows), 256, 0, h->compute>>>(b.n_rows, h->hyper, nullptr, lg, b.label, h->bias, 0, h->red_part.p, h->ticket.p, loss_sum_out, nullptr);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_PREDICT);
  }
}

// ---------------------------------------------------------------------------------------------
// ROC AUC on device (the reference has none -- only eval/loss.h:8-12; north_star asks for logloss AND AUC):
// radix sort of the scores, one scan for (head of the tie group, positives so far), then the Mann-Whitney
// rank sum with average ranks over ties, accumulated in integers (2 * rank sum), hence exact and
// independent of the order of the additions.
// ---------------------------------------------------------------------------------------------
struct AucScan {
  int32_t start;  // position of the head of the tie group at or before this position
  int32_t cum;    // positives at or before this position
};
struct AucScanOp {
  __device__ __forceinline__ AucScan operator()(const AucScan &a, const AucScan &b) const {
    return AucScan{a.start > b.start ? a.start : b.start, a.cum + b.cum};
  }
};
struct AucIn {
  const uint32_t *skey, *slab;
  __device__ __forceinline__ AucScan operator()(int32_t p) const {
    const bool head = p == 0 || skey[p] != skey[p - 1];
    return AucScan{head ? p : 0, (int32_t)slab[p]};
  }
};
__global__ void k_auc_keys(int32_t n, const float *__restrict__ score, const int32_t *__restrict__ label,
                           uint32_t *__restrict__ key, uint32_t *__restrict__ lab) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t u = __float_as_uint(score[i] + 0.0f);  // -0 and +0 tie
  key[i] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending order of the floats
  lab[i] = label[i] > 0 ? 1u : 0u;
}
// out[0] += sum over tie groups of positives(group) * (2 * start + length + 1) ; out[1] = positives
__global__ void k_auc_ranksum(int32_t n, const uint32_t *__restrict__ skey, const AucScan *__restrict__ sc,
                              unsigned long long *__restrict__ out) {
  unsigned long long acc = 0;
  for (int32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    if (p + 1 < n && skey[p + 1] == skey[p]) continue;  // not the tail of its group
    const AucScan e = sc[p];
    const long long len = p - e.start + 1;
    const long long pos = e.cum - (e.start > 0 ? sc[e.start - 1].cum : 0);
    acc += (unsigned long long)(pos * (2ll * e.start + len + 1));
    if (p == n - 1) out[1] = (unsigned long long)e.cum;
  }
  if (acc) atomicAdd(out, acc);
}

static double auc_device(ftrl_handle *h, int64_t n64, const float *d_score, const int32_t *d_label) {
  if (n64 <= 0) return std::nan("");
  if (n64 >= (1ll << 31) - 64) throw ArgFail{"ftrl_eval_auc: n must be < 2^31"};
  const int32_t n = (int32_t)n64;
  DevBuf<uint32_t> key, lab, skey, slab;
  DevBuf<AucScan> sc;
  DevBuf<unsigned long long> out;
  DevBuf<uint8_t> tmp;
  key.alloc(n); lab.alloc(n); skey.alloc(n); slab.alloc(n); sc.alloc(n); out.alloc(2);
  size_t a = 0, b = 0;
  thrust::counting_iterator<int32_t> cnt(0);
  cub::DeviceRadixSort::SortPairs(nullptr, a, key.p, skey.p, lab.p, slab.p, n, 0, 32);
  auto it = thrust::make_transform_iterator(cnt, AucIn{skey.p, slab.p});
  cub::DeviceScan::InclusiveScan(nullptr, b, it, sc.p, AucScanOp(), n);
  size_t bytes = std::max(a, b) + 256;
  tmp.alloc(bytes);
  FTRL_CUDA(cudaMemsetAsync(out.p, 0, 2 * sizeof(unsigned long long), h->compute));
  k_auc_keys<<<(n + 255) / 256, 256, 0, h->compute>>>(n, d_score, d_label, key.p, lab.p);
  size_t t = bytes;
  FTRL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, t, key.p, skey.p, lab.p, slab.p, n, 0, 32, h->compute));
  t = bytes;
  FTRL_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, t, it, sc.p, AucScanOp(), n, h->compute));
  k_auc_ranksum<<<std::min(1024, (n + 255) / 256), 256, 0, h->compute>>>(n, skey.p, sc.p, out.p);
  FTRL_CUDA(cudaGetLastError());
  unsigned long long r[2] = {0, 0};
  FTRL_CUDA(cudaMemcpyAsync(r, out.p, sizeof(r), cudaMemcpyDeviceToHost, h->compute));
  FTRL_CUDA(cudaStreamSynchronize(h->compute));
  const long double n_pos = (long double)r[1], n_neg = (long double)n - n_pos;
  if (n_pos == 0 || n_neg == 0) return std::nan("");
  return (double)(((long double)r[0] - n_pos * (n_pos + 1)) / (2 * n_pos * n_neg));
}

// ---------------------------------------------------------------------------------------------
// host-pointer path: CSR staged through slots, async copies on the copy stream
// ---------------------------------------------------------------------------------------------
static void retire_slot(ftrl_handle *h, Slot &s) {
  if (!s.busy) return;
  FTRL_CUDA(cudaEventSynchronize(s.done));
  if (s.user_out && s.n_out) memcpy(s.user_out, s.h_out.p, sizeof(float) * (size_t)s.n_out);
  if (s.user_loss) *s.user_loss = s.h_loss.p[0];
  s.busy = false;
  s.user_out = nullptr;
  s.user_loss = nullptr;
  s.n_out = 0;
}

static Slot &stage_batch(ftrl_handle *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field, const int32_t *feat,
                         const float *val, const int32_t *label, Batch &b) {
  Slot &s = h->slots[h->next_slot];
  h->next_slot = (h->next_slot + 1) % ftrl_handle::N_SLOTS;
  retire_slot(h, s);
  if (!s.copied) {
    FTRL_CUDA(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    FTRL_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  }
  if (n_rows < 0) throw ArgFail{"n_rows < 0"};
  if (n_rows > 0 && (!row_ptr)) throw ArgFail{"row_ptr is NULL"};
  const int64_t nnz = n_rows > 0 ? row_ptr[n_rows] - row_ptr[0] : 0;
  if (n_rows > 0 && row_ptr[0] != 0) throw ArgFail{"row_ptr[0] must be 0"};
  if (nnz < 0) throw ArgFail{"row_ptr not monotone"};
  if (nnz > 0 && (!field || !feat || !val)) throw ArgFail{"field/feat/val is NULL"};
  s.row_ptr.ensure(n_rows + 1);
  s.label.ensure(n_rows);
  s.out.ensure(n_rows);
  s.loss.ensure(1);
  s.h_out.ensure(n_rows);
  s.h_loss.ensure(1);
  s.field.ensure(nnz);
  s.feat.ensure(nnz);
  s.val.ensure(nnz);
  if (n_rows > 0) {
    FTRL_CUDA(cudaMemcpyAsync(s.row_ptr.p, row_ptr, sizeof(int64_t) * (n_rows + 1), cudaMemcpyHostToDevice, h->copy));
    if (label) FTRL_CUDA(cudaMemcpyAsync(s.label.p, label, sizeof(int32_t) * n_rows, cudaMemcpyHostToDevice, h->copy));
  }
  if (nnz > 0) {
    FTRL_CUDA(cudaMemcpyAsync(s.field.p, field, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, h->copy));
    FTRL_CUDA(cudaMemcpyAsync(s.feat.p, feat, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, h->copy));
    FTRL_CUDA(cudaMemcpyAsync(s.val.p, val, sizeof(float) * nnz, cudaMemcpyHostToDevice, h->copy));
  }
  FTRL_CUDA(cudaEventRecord(s.copied, h->copy));  // the caller orders its kernels behind this event
  b.n_rows = n_rows;
  b.nnz = nnz;
  b.row_ptr = s.row_ptr.p;
  b.field = s.field.p;
  b.feat = s.feat.p;
  b.val = s.val.p;
  b.label = label ? s.label.p : nullptr;
  return s;
}

static void finish_slot(ftrl_handle *h, Slot &s, int64_t n_rows, float *user_out, double *user_loss) {
  if (user_out && n_rows > 0)
    FTRL_CUDA(cudaMemcpyAsync(s.h_out.p, s.out.p, sizeof(float) * n_rows, cudaMemcpyDeviceToHost, h->compute));
  if (user_loss) FTRL_CUDA(cudaMemcpyAsync(s.h_loss.p, s.loss.p, sizeof(double), cudaMemcpyDeviceToHost, h->compute));
  FTRL_CUDA(cudaEventRecord(s.done, h->compute));
  s.busy = true;
  s.user_out = user_out;
  s.user_loss = user_loss;
  s.n_out = user_out ? n_rows : 0;
}

static void check_device_err(ftrl_handle *h) {
  int32_t e = 0;
  FTRL_CUDA(cudaMemcpy(&e, h->d_err, sizeof(e), cudaMemcpyDeviceToHost));
  if (e) {
    FTRL_CUDA(cudaMemset(h->d_err, 0, sizeof(e)));
    if (e == 3) throw ArgFail{"multi-GPU run: the rows owned by one rank exceed its workspace (extreme id skew): the step was skipped on every rank, no state was changed"};
    if (e == 4) throw StateFail{"multi-GPU run: a peer did not reach the device barrier in time"};
    throw ArgFail{"sequential mode: a sample exceeds the supported size (more than 96 valid features, or FM n_factors > 1024)"};
  }
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// feature-sharded multi-GPU step (shard.cuh)
// ---------------------------------------------------------------------------------------------
// channel 0: weight-dependent phase (compute stream); channel 1: index phase
static void peer_barrier(ftrl_handle *h, const Peers &pr, int channel) {
  uint32_t &ep = channel ? h->epoch_id : h->epoch;
  ep++;
  k_peer_barrier<<<1, 32, 0, h->compute>>>(pr, channel, ep, h->barrier_timeout_cycles, h->d_err);
  FTRL_CUDA(cudaGetLastError());
}

// the peers' published lists of the step with parity `par` (two sets in one allocation per rank)
static Peers peers_of_parity(const ftrl_handle *h, int par) {
  Peers pr = h->peers;
  const int64_t off = (int64_t)par * h->nnz_cap;
  for (int q = 0; q < pr.G; q++) {
    pr.ukey[q] += off;
    pr.uinfo[q] += off;
    pr.umask[q] += off;
    pr.dst_at[q] += off;
  }
  return pr;
}

// One sharded step = an INDEX phase (ids only: S1, barrier on the index channel, the id part of S2, the local chunk
// list) and a WEIGHT phase (k_owner_materialise, barrier 2, S3, barrier 3, S4).  With `inputs_ready` (host path) or
// stable device inputs the index phase of step t+1 runs on the index stream under the weight phase of step t.
// Everything the index phase writes is double-buffered by the parity of the step (IdSet; the published lists,
// dst_at and the SyncArea fields): a set is rewritten by the index phase of step t+2, which waits for this rank's
// weight phase of step t -- and that ends behind barrier 3 of step t, which every rank reaches only after its
// last read of a peer's set (k_owner_materialise: umask / uinfo; k_ffm_staged_rows / k_ffm_combine: dst_at;
// k_check_abort: abort_at).  boff / simple are read in the index phase of step t, which every rank has finished
// before any rank passes the index barrier of step t+1.
template <bool PRECISE>
static void train_device_sharded(ftrl_handle *h, const Batch &b, float *logit_out, double *loss_sum_out,
                                 cudaEvent_t inputs_ready, bool allow_pipe) {
  if (!h->attached) throw StateFail{"multi-GPU handle: call ftrl_attach_peers before training"};
  const Dims &d = h->dims;
  if (b.n_rows > h->rows_cap || b.nnz > h->nnz_cap) throw ArgFail{"batch exceeds max_batch_rows / max_batch_nnz of a multi-GPU handle"};
  const int32_t nnz = (int32_t)b.nnz;
  const uint32_t sentinel = (uint32_t)d.n_feats;
  const uint32_t lsent = (uint32_t)h->n_local;  // sentinel in local-row space (n_local may differ by 1 across ranks)
  const int32_t oc = (int32_t)h->ow_cap;
  const int grid = h->n_sms * 4;
  thrust::counting_iterator<int32_t> cnt(0);
  if (!logit_out) logit_out = h->logit_ws.p;
  const int par = (int)(h->shard_step & 1u);
  const uint32_t step_tag = ++h->shard_step;  // the same on every rank, never 0
  if (h->idstream && h->idset_cur != par) swap_idsets(h);
  const Peers pr = peers_of_parity(h, par);
  h->exportd.dst_at = h->dst_at.p + (int64_t)par * h->nnz_cap;
  uint32_t *const ukey = h->ukey.p + (int64_t)par * h->nnz_cap, *const uinfo = h->uinfo.p + (int64_t)par * h->nnz_cap;
  unsigned long long *const umask = h->umask.p + (int64_t)par * h->nnz_cap;
  const bool piped = allow_pipe && h->pipeline && h->idstream && !h->profiling && (inputs_ready || h->stable_device_inputs);
  cudaStream_t main_stream = h->compute;
  const bool is_ffm = d.model_type == FTRL_FFM;

  // ------------------------------- index phase -------------------------------
  if (piped) {
    if (inputs_ready) FTRL_CUDA(cudaStreamWaitEvent(h->idstream, inputs_ready, 0));
    // this parity's sets were last read by the weight phase two steps ago
    if (h->hot_recorded[par]) FTRL_CUDA(cudaStreamWaitEvent(h->idstream, h->ev_hot_done[par], 0));
    h->compute = h->idstream;  // everything below goes to the index stream
  } else if (inputs_ready) {
    FTRL_CUDA(cudaStreamWaitEvent(h->compute, inputs_ready, 0));
  }
  // index phases run in order, whichever stream the previous one used
  if (h->id_tail_recorded) FTRL_CUDA(cudaStreamWaitEvent(h->compute, h->ev_id_tail, 0));
  try {
    {
      PhaseScope ps(h, PH_PREP);
      FTRL_CUDA(cudaMemsetAsync(h->batch_flags.p, 0x01, sizeof(int32_t), h->compute));
      FTRL_CUDA(cudaMemsetAsync(h->batch_flags.p + 1, 0, sizeof(int32_t), h->compute));
      if (b.n_rows > 0) {
        const unsigned pg = (unsigned)((b.n_rows * 32 + 255) / 256);
        k_prep_rows<<<pg, 256, 0, h->compute>>>(b, d, h->key.p, h->occ_idx.p, h->occ_row.p, h->sflags.p, h->batch_flags.p,
                                                h->pmask.p);
        launched(h, PH_PREP);
      }
      FTRL_CUDA(cudaGetLastError());
    }
    {
      PhaseScope ps(h, PH_SORT);
      if (nnz > 0) {
        size_t bytes = h->cub_bytes;
        FTRL_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, bytes, h->key.p, h->skey.p, h->occ_idx.p, h->socc.p, nnz, 0,
                                                  key_bits(d.n_feats), h->compute));
      }
    }
    {
      // distinct rows of the local batch, their field masks; tell the peers
      PhaseScope ps(h, PH_SEGMENT);
      if (nnz > 0) {
        k_occ_class<<<(nnz + 255) / 256, 256, 0, h->compute>>>(nnz, sentinel, 0, h->skey.p, h->socc.p, h->occ_row.p, h->sflags.p,
                                                               h->fused_sorted.p, h->occ_pos.p);
        size_t bytes = h->cub_bytes;
        if (is_ffm) {  // which field slices of a row does this rank's batch touch (LR / FM rows have no slices)
          auto mit = thrust::make_transform_iterator(cnt, MaskIn{h->skey.p, h->socc.p, h->pmask.p});
          FTRL_CUDA(cub::DeviceScan::InclusiveScan(h->cub_tmp.p, bytes, mit, h->mscan.p, MaskScanOp(), nnz, h->compute));
        }
        bytes = h->cub_bytes;
        FTRL_CUDA(cub::DeviceSelect::If(h->cub_tmp.p, bytes, cnt, h->uhead.p, h->n_uall.p, nnz, RowHeadPred{h->skey.p}, h->compute));
      }
      if (nnz > 0) {
        // the list goes out bucketed by owner: one stable radix pass over the owner bits
        k_owner_keys<<<(nnz + 255) / 256, 256, 0, h->compute>>>(nnz, h->G, sentinel, h->uhead.p, h->n_uall.p, h->skey.p, h->bkey.p,
                                                               h->bidx.p);
        size_t bytes = h->cub_bytes;
        FTRL_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, bytes, h->bkey.p, h->bkey_s.p, h->bidx.p, h->perm.p, nnz, 0,
                                                  h->log2G + 1, h->compute));
        k_publish_unique<<<(nnz + 255) / 256, 256, 0, h->compute>>>(nnz, h->G, h->bkey_s.p, h->perm.p, h->uhead.p, h->n_uall.p, h->skey.p,
                                                                   is_ffm ? h->mscan.p : nullptr, ukey, uinfo, umask);
      }
      k_publish_bounds<<<1, 32, 0, h->compute>>>(pr, par, nnz, h->bkey_s.p, h->batch_flags.p);
      FTRL_CUDA(cudaGetLastError());
      launched(h, PH_SEGMENT, 4);
      peer_barrier(h, pr, 1);  // 1: every rank's distinct-row list is published
    }
    {
      // owner side: contributions (row, rank) of the rows this rank owns
      PhaseScope ps(h, PH_EXCHANGE);
      k_merge_flags<<<1, 1, 0, h->compute>>>(pr, par, is_ffm ? 1 : 0, h->batch_flags.p);
      k_fill_owned<<<(oc + 255) / 256, 256, 0, h->compute>>>(pr, par, oc, lsent, step_tag, h->n_sel.p, h->okey.p, h->osrc.p, h->d_err);
      size_t bytes = h->cub_bytes;
      FTRL_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, bytes, h->okey.p, h->ckey.p, h->osrc.p, h->csrc.p, oc, 0,
                                                key_bits((int32_t)h->n_local), h->compute));
      k_contrib_class<<<(oc + 255) / 256, 256, 0, h->compute>>>(pr, oc, h->n_sel.p, lsent, is_ffm ? 1 : 0, h->batch_flags.p,
                                                               h->ckey.p, h->csrc.p, h->socc.p, h->cflag.p,
                                                               h->fused_sorted.p, h->occ_pos.p);
      FTRL_CUDA(cudaGetLastError());
      launched(h, PH_EXCHANGE, 3);
    }
    {
      // local chunk list of the rows that are not finalised inside a sample
      PhaseScope ps(h, PH_SEGMENT);
      if (nnz > 0) {
        size_t bytes = h->cub_bytes;
        auto it = thrust::make_transform_iterator(cnt, HeadFunctor{h->skey.p, h->fused_sorted.p});
        FTRL_CUDA(cub::DeviceScan::InclusiveScan(h->cub_tmp.p, bytes, it, h->scan.p, SegScanOp(), nnz, h->compute));
        bytes = h->cub_bytes;
        FTRL_CUDA(cub::DeviceSelect::If(h->cub_tmp.p, bytes, cnt, h->chunk_pos.p, h->n_chunks.p, nnz,
                                        ChunkHeadPred{h->skey.p, h->scan.p, h->fused_sorted.p, sentinel, h->chunk}, h->compute));
        k_terminate<<<1, 1, 0, h->compute>>>(h->chunk_pos.p, h->n_chunks.p, nnz);
        if (is_ffm)
          k_chunk_desc<8><<<h->n_sms * 8, 256, 0, h->compute>>>(nnz, sentinel, h->chunk, h->n_chunks.p, h->chunk_pos.p, h->skey.p,
                                                               h->scan.p, h->cdesc.p);
      } else {
        FTRL_CUDA(cudaMemsetAsync(h->n_chunks.p, 0, sizeof(int32_t), h->compute));
      }
      FTRL_CUDA(cudaGetLastError());
      launched(h, PH_SEGMENT);
    }
    if (h->ev_id_tail) {
      FTRL_CUDA(cudaEventRecord(h->ev_id_tail, h->compute));
      h->id_tail_recorded = true;
    }
    if (piped) FTRL_CUDA(cudaEventRecord(h->ev_id_done[par], h->idstream));
  } catch (...) {
    h->compute = main_stream;
    throw;
  }
  h->compute = main_stream;
  if (piped) FTRL_CUDA(cudaStreamWaitEvent(h->compute, h->ev_id_done[par], 0));

  // ------------------------------- weight phase -------------------------------
  {
    PhaseScope ps(h, PH_EXCHANGE);
    // the cub scratch is shared by the two phases' sorts: this phase has none
    k_owner_materialise<PRECISE, 256><<<grid, 256, 0, h->compute>>>(pr, d, h->hyper, is_ffm ? 0 : 1, oc, h->n_sel.p,
                                                                    h->batch_flags.p, h->ckey.p, h->csrc.p, h->cflag.p, h->tab,
                                                                    h->lin);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_EXCHANGE);
    peer_barrier(h, pr, 0);  // 2: classes / inbox slots are known everywhere, w of the touched slices is materialised
    k_check_abort<<<1, 1, 0, h->compute>>>(pr, par, step_tag, h->batch_flags.p, h->d_err);
    FTRL_CUDA(cudaGetLastError());
  }
  const ItemDecode dec4 = make_item_decode(d.k, 4);
  if (is_ffm) {
    {
      PhaseScope ps(h, PH_SAMPLE);
      if (b.n_rows > 0) launch_tile<PRECISE>(h, b, dec4, logit_out);
    }
    // a batch in which some sample (of any rank) repeats a field: the generic kernels take it (device-side flag)
    PhaseScope ps(h, PH_GENERIC);
    if (b.n_rows > 0) {
      launch_ffm_sample<4, PRECISE, true>(h, b, logit_out, 1);
      launched(h, PH_GENERIC);
    }
  } else if (d.model_type == FTRL_FM) {
    const int vec = pick_vec(d.k);
    if (vec == 4) run_lrfm_batch_sharded<4, PRECISE, true>(h, b, logit_out);
    else if (vec == 2) run_lrfm_batch_sharded<2, PRECISE, true>(h, b, logit_out);
    else run_lrfm_batch_sharded<1, PRECISE, true>(h, b, logit_out);
  } else {
    run_lrfm_batch_sharded<1, PRECISE, false>(h, b, logit_out);
  }
  {
    PhaseScope ps(h, PH_REDUCE);
    k_batch_reduce<PRECISE><<<reduce_grid(std::max<int64_t>(1, b.n_rows)), 256, 0, h->compute>>>(
        b.n_rows, h->hyper, h->g.p, logit_out, b.label, h->bias, 0, h->red_part.p, h->ticket.p, loss_sum_out, h->red4.p);
    k_publish_red<<<1, 32, 0, h->compute>>>(pr, h->red4.p);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_REDUCE, 2);
  }
  if (is_ffm) {
    {
      // local duplicates are reduced here; each row's sum goes to its owner's inbox (or is applied here)
      PhaseScope ps(h, PH_ROWS);
      launch_staged_rows<PRECISE>(h, b);
    }
    if (nnz > 0) {
      PhaseScope ps(h, PH_GENERIC);
      launch_ffm_rows<4, PRECISE, true>(h, b, dec4, 1);
      launched(h, PH_GENERIC);
    }
    PhaseScope ps(h, PH_COMBINE);
    k_ffm_combine<PRECISE, 256><<<grid, 256, 0, h->compute>>>(d, h->hyper, h->tab, h->lin, h->n_chunks.p, h->cdesc.p, h->part.p,
                                                              h->part_lin.p, h->exportd, h->batch_flags.p);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_COMBINE);
  } else if (d.model_type == FTRL_FM) {
    run_lrfm_rows_sharded<PRECISE, true>(h, b);
  } else {
    run_lrfm_rows_sharded<PRECISE, false>(h, b);
  }
  peer_barrier(h, pr, 0);  // 3: every contribution is in its owner's inbox, the bias partials are exchanged
  {
    PhaseScope ps(h, PH_APPLY);
    k_owner_apply<PRECISE, 256><<<grid, 256, 0, h->compute>>>(pr, d, h->hyper, oc, h->n_sel.p, h->batch_flags.p, h->ckey.p,
                                                              h->cflag.p, h->inbox.p, h->inbox_lin.p, h->tab, h->lin);
    k_bias_apply<PRECISE><<<1, 1, 0, h->compute>>>(pr, h->hyper, h->batch_flags.p, h->bias);
    FTRL_CUDA(cudaGetLastError());
    launched(h, PH_APPLY, 2);
  }
  if (h->ev_hot_done[par]) {
    FTRL_CUDA(cudaEventRecord(h->ev_hot_done[par], h->compute));
    h->hot_recorded[par] = true;
  }
  h->stats.kernel_launches = h->launches_this_call;
}

constexpr int PEER_BUFS = 11;
struct PeerBlob {  // FTRL_PEER_BLOB_BYTES
  uint32_t magic;
  int32_t rank, world, device;
  int64_t pid;
  int64_t nnz_cap, ow_cap;
  void *raw[PEER_BUFS];                // same-process attach
  cudaIpcMemHandle_t ipc[PEER_BUFS];   // tab, lin, ukey, uinfo, umask, dst_at, inbox, inbox_lin, sync, rc_w, rc_lin
};
static_assert(sizeof(PeerBlob) <= FTRL_PEER_BLOB_BYTES, "peer blob too large");


extern "C" {

int ftrl_abi_version(void) { return FTRL_B200_ABI_VERSION; }

void ftrl_config_default(ftrl_config *cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->model_type = FTRL_FFM;
  cfg->n_feats = 10000;
  cfg->n_fields = 8;
  cfg->n_factors = 16;
  cfg->init_mean = 0.0f;
  cfg->init_stddev = 0.02f;
  cfg->w_alpha = 1e-4f;
  cfg->w_beta = 1.0f;
  cfg->w_l1 = 0.1f;
  cfg->w_l2 = 5.0f;
  cfg->mode = FTRL_MODE_BATCH;
  cfg->device = 0;
  cfg->seed = 42;
  cfg->world_size = 1;
}

const char *ftrl_last_error(const ftrl_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ftrl_create(const ftrl_config *cfg, ftrl_handle **out) {
  if (!cfg || !out) {
    g_create_error = "cfg/out is NULL";
    return FTRL_ERR_ARG;
  }
  *out = nullptr;
  ftrl_handle *h = nullptr;
  int rc = guarded(nullptr, [&] {
    if (cfg->model_type < 0 || cfg->model_type > 2) throw ArgFail{fmt("Invalid model_type: %d, expect LR(0), FM(1) or FFM(2).", cfg->model_type)};
    if (cfg->n_feats <= 0) throw ArgFail{"n_feats must be > 0"};
    if (cfg->model_type != FTRL_LR && cfg->n_factors <= 0) throw ArgFail{"n_factors must be > 0"};
    if (cfg->model_type == FTRL_FFM && cfg->n_fields <= 0) throw ArgFail{"n_fields must be > 0"};
    if (cfg->model_type == FTRL_FM && cfg->mode == FTRL_MODE_BATCH && cfg->n_factors > 32 * FM_MAX_REGS)
      throw ArgFail{"FM n_factors > 256 is unsupported in minibatch mode (sequential mode has no cap)"};
    if (cfg->mode != FTRL_MODE_BATCH && cfg->mode != FTRL_MODE_SEQUENTIAL) throw ArgFail{"bad mode"};
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
      throw CudaFail{e == cudaSuccess ? cudaErrorNoDevice : e, "no usable CUDA device (this library has no CPU fallback)", __FILE__, __LINE__};
    if (cfg->device < 0 || cfg->device >= n_dev) throw ArgFail{fmt("device %d out of range (%d devices)", cfg->device, n_dev)};
    FTRL_CUDA(cudaSetDevice(cfg->device));
    h = new ftrl_handle();
    h->cfg = *cfg;
    if (h->cfg.world_size < 1) h->cfg.world_size = 1;
    Dims &d = h->dims;
    d.model_type = cfg->model_type;
    d.n_feats = cfg->n_feats;
    d.n_fields = cfg->model_type == FTRL_FFM ? cfg->n_fields : 1;
    d.k = cfg->model_type == FTRL_LR ? 0 : cfg->n_factors;
    const int64_t rl = cfg->model_type == FTRL_FFM ? (int64_t)cfg->n_fields * cfg->n_factors : cfg->model_type == FTRL_FM ? cfg->n_factors : 0;
    if (rl > (1 << 24)) throw ArgFail{"n_fields*n_factors too large"};
    d.row_len = (int32_t)rl;
    d.ld = (int32_t)((rl + 3) / 4 * 4);
    h->hyper = Hyper{cfg->w_alpha, cfg->w_beta, cfg->w_l1, cfg->w_l2, 1.0f / cfg->w_alpha};
    h->G = h->cfg.world_size;
    h->rank = h->cfg.rank;
    if (h->G > 1) {
      if (h->G > MAX_SHARDS || (h->G & (h->G - 1))) throw ArgFail{"world_size must be 1, 2, 4 or 8"};
      if (h->rank < 0 || h->rank >= h->G) throw ArgFail{"rank out of range"};
      if (cfg->mode != FTRL_MODE_BATCH) throw ArgFail{"feature-sharded multi-GPU runs are minibatch mode only"};
      if (cfg->max_batch_rows <= 0 || cfg->max_batch_nnz <= 0)
        throw ArgFail{"multi-GPU runs need max_batch_rows / max_batch_nnz (buffers are mapped by peers, they cannot grow)"};
      if (cfg->max_batch_nnz >= (1ll << SRC_SHIFT)) throw ArgFail{"max_batch_nnz must be < 2^28 in multi-GPU runs"};
      while ((1 << h->log2G) < h->G) h->log2G++;
    }
    h->n_local = h->G > 1 ? ((int64_t)cfg->n_feats - h->rank + h->G - 1) / h->G : cfg->n_feats;
    static const char *names[PH_COUNT] = {"prep_rows", "sort", "segment", "sample", "rows", "combine", "reduce", "exact", "predict", "generic", "materialise", "exchange", "pull", "apply"};
    for (int i = 0; i < PH_COUNT; i++) h->phases[i].name = names[i];
    cudaDeviceProp prop;
    FTRL_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    h->n_sms = prop.multiProcessorCount;
    h->fuse = env_int("FTRL_B200_FUSE", 1);
    h->sample_threads = env_int("FTRL_B200_SAMPLE_THREADS", 0);
    h->precise = env_int("FTRL_B200_PRECISE", 0);
    h->barrier_timeout_cycles = (long long)env_int("FTRL_B200_BARRIER_TIMEOUT_S", 20) * 2000000000ll;
    h->chunk = env_int("FTRL_B200_CHUNK", cfg->model_type == FTRL_FFM ? 32 : cfg->model_type == FTRL_FM ? 256 : 2048);
    if (h->chunk < 1) h->chunk = 1;
    if (cfg->model_type == FTRL_FFM && h->chunk > 32) h->chunk = 32;  // chunk ends are found with one ballot
    h->tile = env_int("FTRL_B200_TILE", 1);
    h->tile_dbg = env_int("FTRL_B200_TILE_DBG", 0);
    if (cfg->model_type == FTRL_FFM && h->tile && h->fuse && d.k % 4 == 0 && d.n_fields <= 64) {
      const int stride = tile_stride(d.ld, d.k);
      const size_t stage = tile_stage_bytes(d.n_fields, stride);
      const size_t budget = (size_t)prop.sharedMemPerBlockOptin - 2048;  // static mbarriers / reduction scratch
      const int64_t items = (int64_t)d.n_fields * (d.n_fields - 1) / 2 * (d.k / 4);
      int cons = 64;
      while (cons < 512 && cons * 3 < items) cons *= 2;
      if (cons == 512 && items > 2 * 512) cons = TILE_MAX_CONSUMERS;  // 24 consumer warps, 2 items per thread at cfg4
      cons = env_int("FTRL_B200_TILE_CONSUMERS", cons);
      const int ipt = (int)((items + cons - 1) / cons);
      const size_t lut = tile_lut_bytes(d.n_fields) + 4 * tile_meta_bytes(d.n_fields);  // + minimal meta ring
      if (2 * stage + lut <= budget && ipt <= 4) {
        int ctas = (int)std::min<size_t>(8, (size_t)prop.sharedMemPerMultiprocessor / (2 * stage + lut + 2048));
        ctas = std::max(1, std::min(ctas, 2048 / tile_threads(cons)));
        ctas = env_int("FTRL_B200_TILE_CTAS", ctas);
        int stages = (int)std::min<size_t>(TILE_MAX_STAGE, ((size_t)prop.sharedMemPerMultiprocessor / ctas - 2048 - lut) / stage);
        stages = std::max(2, std::min(stages, (int)((budget - lut) / stage)));
        stages = env_int("FTRL_B200_TILE_STAGES", stages);
        int metas = std::min(TILE_MAX_META, stages + 4);
        while (metas > stages + 1 && tile_smem_bytes(d.n_fields, stride, stages, metas) > budget) metas--;
        metas = env_int("FTRL_B200_TILE_META", metas);
        h->tile_meta = metas;
        h->tile_ok = true;
        h->tile_f_cap = d.n_fields;
        h->tile_stride = stride;
        h->tile_stride1 = tile_stride1(d.ld, stride);
        h->tile_stages = stages;
        h->tile_consumers = cons;
        h->tile_ipt = std::max(1, ipt);
        h->tile_ctas_per_sm = ctas;
        h->tile_smem = tile_smem_bytes(d.n_fields, stride, stages, metas);
        h->tile_ring = (int)(stage * stages);
        if (ctas == 1 && env_int("FTRL_B200_TILE_BIGRING", 1)) {
          // one CTA per SM: the variable-span row ring takes all the shared memory that is left (Zipf batches keep
          // 3.8 instead of 3.4 samples in flight at cfg4)
          const size_t extra = (budget - h->tile_smem) / 128 * 128;
          h->tile_ring += (int)extra;
          h->tile_smem += extra;
        }
        const int sm = (int)h->tile_smem;
#define TILE_ATTR(P, I)                                                                                              \
  FTRL_CUDA(cudaFuncSetAttribute(k_ffm_tile<P, I, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); \
  FTRL_CUDA(cudaFuncSetAttribute(k_ffm_tile<P, I, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm))
        TILE_ATTR(false, 1); TILE_ATTR(false, 2); TILE_ATTR(false, 3); TILE_ATTR(false, 4);
        TILE_ATTR(true, 1); TILE_ATTR(true, 2); TILE_ATTR(true, 3); TILE_ATTR(true, 4);
        h->tile_cache = env_int("FTRL_B200_TILE_CACHE", 1);
        h->tile_inflight = std::max(2, std::min(TILE_MAX_STAGE, env_int("FTRL_B200_TILE_INFLIGHT", TILE_MAX_STAGE)));
#undef TILE_ATTR
      }
    }
    FTRL_CUDA(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
    FTRL_CUDA(cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
    h->pipeline = env_int("FTRL_B200_PIPELINE", 1);
    h->stable_device_inputs = (cfg->reserved[0] & 1) != 0;
    if (h->pipeline && cfg->mode == FTRL_MODE_BATCH) {
      // same priority as the compute stream: with a lower one the index phase only runs once the weight-dependent
      // kernels have no pending CTA left, i.e. not under them at all (measured: the overlap gain vanished)
      FTRL_CUDA(cudaStreamCreateWithFlags(&h->idstream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) {
        FTRL_CUDA(cudaEventCreateWithFlags(&h->ev_id_done[i], cudaEventDisableTiming));
        FTRL_CUDA(cudaEventCreateWithFlags(&h->ev_hot_done[i], cudaEventDisableTiming));
      }
      FTRL_CUDA(cudaEventCreateWithFlags(&h->ev_id_tail, cudaEventDisableTiming));
    }
    if (h->G > 1 && d.model_type == FTRL_FFM && !h->tile_ok)
      throw ArgFail{"multi-GPU FFM runs need the tile path (n_factors % 4 == 0, n_fields <= 64, sample tile must fit shared memory)"};
    const int64_t n = std::max<int64_t>(1, h->n_local);
    FTRL_CUDA(cudaMalloc(&h->lin, page_round(sizeof(float4) * n)));
    FTRL_CUDA(cudaMalloc(&h->bias, sizeof(float4)));
    FTRL_CUDA(cudaMemsetAsync(h->bias, 0, sizeof(float4), h->compute));
    FTRL_CUDA(cudaMalloc(&h->d_err, sizeof(int32_t)));
    FTRL_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), h->compute));
    FTRL_CUDA(cudaMalloc(&h->pair_lut, sizeof(uint32_t) * PAIR_LUT_N));
    k_build_pair_lut<<<(PAIR_LUT_N + 255) / 256, 256, 0, h->compute>>>(h->pair_lut);
    k_init_lin<<<(unsigned)((n + 255) / 256), 256, 0, h->compute>>>(h->lin, n, cfg->init_mean, cfg->init_stddev, cfg->seed, h->G, h->rank);
    if (!d.row_len && h->G > 1) {  // sharded LR: a placeholder the peers can map (no latent table)
      FTRL_CUDA(cudaMalloc(&h->tab, page_round(64)));
      FTRL_CUDA(cudaMemsetAsync(h->tab, 0, 64, h->compute));
    }
    if (d.row_len) {
      FTRL_CUDA(cudaMalloc(&h->tab, page_round(sizeof(float) * 3 * (size_t)d.ld * (size_t)n)));
      const int64_t q = n * (d.ld / 4);
      k_init_tab<<<(unsigned)((q + 255) / 256), 256, 0, h->compute>>>(h->tab, n, d.row_len, d.ld, cfg->init_mean, cfg->init_stddev, cfg->seed, h->G, h->rank);
    }
    FTRL_CUDA(cudaGetLastError());
    if (cfg->max_batch_rows > 0) {
      // no allocation on the hot path: workspace and the CSR staging slots are sized up front
      const int64_t mr = cfg->max_batch_rows, mn = std::max<int64_t>(cfg->max_batch_nnz, 0);
      ensure_workspace(h, mr, mn);
      if (h->idstream) {  // both index sets
        swap_idsets(h);
        ensure_workspace(h, mr, mn);
        swap_idsets(h);
      }
      for (auto &sl : h->slots) {
        sl.row_ptr.ensure(mr + 1);
        sl.label.ensure(mr);
        sl.out.ensure(mr);
        sl.loss.ensure(1);
        sl.h_out.ensure(mr);
        sl.h_loss.ensure(1);
        sl.field.ensure(mn);
        sl.feat.ensure(mn);
        sl.val.ensure(mn);
        FTRL_CUDA(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
        FTRL_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
      }
    }
    if (h->G > 1) {
      FTRL_CUDA(cudaMalloc(&h->sync, page_round(sizeof(SyncArea))));
      FTRL_CUDA(cudaMemsetAsync(h->sync, 0, sizeof(SyncArea), h->compute));
    }
    refresh_shards(h);
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
  });
  if (rc != FTRL_OK) {
    if (h) ftrl_destroy(h);
    return rc;
  }
  *out = h;
  return FTRL_OK;
}

void ftrl_destroy(ftrl_handle *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (auto &s : h->slots) {
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.done) cudaEventDestroy(s.done);
  }
  for (auto &pe : h->pending) {
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  if (h->sync) cudaFree(h->sync);
  if (h->tab) cudaFree(h->tab);
  if (h->lin) cudaFree(h->lin);
  if (h->bias) cudaFree(h->bias);
  if (h->pair_lut) cudaFree(h->pair_lut);
  if (h->d_err) cudaFree(h->d_err);
  for (int i = 0; i < 2; i++) {
    if (h->ev_id_done[i]) cudaEventDestroy(h->ev_id_done[i]);
    if (h->ev_hot_done[i]) cudaEventDestroy(h->ev_hot_done[i]);
  }
  if (h->ev_id_tail) cudaEventDestroy(h->ev_id_tail);
  if (h->idstream) cudaStreamDestroy(h->idstream);
  if (h->compute && h->own_compute) cudaStreamDestroy(h->compute);
  if (h->copy) cudaStreamDestroy(h->copy);
  delete h;
}

int64_t ftrl_row_len(const ftrl_handle *h) { return h ? h->dims.row_len : 0; }

int ftrl_train_batch_device(ftrl_handle *h, int64_t n_rows, int64_t nnz, const int64_t *row_ptr, const int32_t *field,
                            const int32_t *feat, const float *val, const int32_t *label, float *logits_out,
                            double *loss_sum_out) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (n_rows < 0 || nnz < 0) throw ArgFail{"negative size"};
    if (n_rows > 0 && (!row_ptr || !label)) throw ArgFail{"row_ptr/label is NULL"};
    if (nnz > 0 && (!field || !feat || !val)) throw ArgFail{"field/feat/val is NULL"};
    Batch b{n_rows, nnz, row_ptr, field, feat, val, label};
    train_device(h, b, logits_out, loss_sum_out);
  });
}

int ftrl_train_batch(ftrl_handle *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field, const int32_t *feat,
                     const float *val, const int32_t *label, float *logits_out, double *loss_sum_out) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (n_rows > 0 && !label) throw ArgFail{"label is NULL"};
    Batch b{};
    Slot &s = stage_batch(h, n_rows, row_ptr, field, feat, val, label, b);
    train_device(h, b, logits_out ? s.out.p : nullptr, s.loss.p, s.copied);
    finish_slot(h, s, n_rows, logits_out, loss_sum_out);
  });
}

int ftrl_predict_batch_device(ftrl_handle *h, int64_t n_rows, int64_t nnz, const int64_t *row_ptr, const int32_t *field,
                              const int32_t *feat, const float *val, const int32_t *label, int output_prob, float *out,
                              double *loss_sum_out) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (n_rows < 0 || nnz < 0) throw ArgFail{"negative size"};
    if (n_rows > 0 && (!row_ptr || !out)) throw ArgFail{"row_ptr/out is NULL"};
    if (loss_sum_out && !label && n_rows > 0) throw ArgFail{"loss requested without labels"};
    Batch b{n_rows, nnz, row_ptr, field, feat, val, label};
    predict_device(h, b, output_prob, out, loss_sum_out);
  });
}

int ftrl_predict_batch(ftrl_handle *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field, const int32_t *feat,
                       const float *val, const int32_t *label, int output_prob, float *out, double *loss_sum_out) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (n_rows > 0 && !out) throw ArgFail{"out is NULL"};
    if (loss_sum_out && !label && n_rows > 0) throw ArgFail{"loss requested without labels"};
    Batch b{};
    Slot &s = stage_batch(h, n_rows, row_ptr, field, feat, val, label, b);
    FTRL_CUDA(cudaStreamWaitEvent(h->compute, s.copied, 0));
    predict_device(h, b, output_prob, s.out.p, loss_sum_out ? s.loss.p : nullptr);
    finish_slot(h, s, n_rows, out, loss_sum_out);
  });
}

int ftrl_sync(ftrl_handle *h) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    FTRL_CUDA(cudaStreamSynchronize(h->copy));
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
    for (auto &s : h->slots) retire_slot(h, s);
    if (h->cfg.mode == FTRL_MODE_SEQUENTIAL || h->G > 1) check_device_err(h);
  });
}

// ---- state access ---------------------------------------------------------------------------
static void xfer_rows(ftrl_handle *h, int which, int64_t row0, int64_t n_rows, float *lin, float *vec, bool get) {
  const Dims &d = h->dims;
  if (which < 0 || which > 2) throw ArgFail{"which must be 0 (w), 1 (n) or 2 (z)"};
  if (row0 < 0 || n_rows < 0 || row0 + n_rows > h->n_local) throw ArgFail{"row range out of bounds"};
  const int plane = plane_of(which);
  FTRL_CUDA(cudaStreamSynchronize(h->compute));
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(16 << 20) / std::max<int32_t>(1, d.row_len));
  h->xfer.ensure((size_t)std::min<int64_t>(n_rows, chunk_rows) * std::max<int32_t>(1, d.row_len));
  for (int64_t r = 0; r < n_rows; r += chunk_rows) {
    const int64_t nr = std::min(chunk_rows, n_rows - r);
    if (lin) {
      const unsigned grid = (unsigned)((nr + 255) / 256);
      if (get) {
        k_lin_copy<<<grid, 256, 0, h->compute>>>(h->lin, h->xfer.p, row0 + r, nr, plane, 1);
        FTRL_CUDA(cudaMemcpyAsync(lin + r, h->xfer.p, sizeof(float) * nr, cudaMemcpyDeviceToHost, h->compute));
      } else {
        FTRL_CUDA(cudaMemcpyAsync(h->xfer.p, lin + r, sizeof(float) * nr, cudaMemcpyHostToDevice, h->compute));
        k_lin_copy<<<grid, 256, 0, h->compute>>>(h->lin, h->xfer.p, row0 + r, nr, plane, 0);
      }
      FTRL_CUDA(cudaStreamSynchronize(h->compute));
    }
    if (vec && d.row_len) {
      const int64_t q = nr * d.row_len;
      const unsigned grid = (unsigned)((q + 255) / 256);
      if (get) {
        k_plane_copy<<<grid, 256, 0, h->compute>>>(h->tab, h->xfer.p, row0 + r, nr, d.row_len, d.ld, plane, 1);
        FTRL_CUDA(cudaMemcpyAsync(vec + r * d.row_len, h->xfer.p, sizeof(float) * q, cudaMemcpyDeviceToHost, h->compute));
      } else {
        FTRL_CUDA(cudaMemcpyAsync(h->xfer.p, vec + r * d.row_len, sizeof(float) * q, cudaMemcpyHostToDevice, h->compute));
        k_plane_copy<<<grid, 256, 0, h->compute>>>(h->tab, h->xfer.p, row0 + r, nr, d.row_len, d.ld, plane, 0);
      }
      FTRL_CUDA(cudaStreamSynchronize(h->compute));
    }
  }
  FTRL_CUDA(cudaGetLastError());
}

// rows [row0, row0 + n_rows) of the WHOLE model (global feature ids), read through the mapped peer shards
static void get_rows_global(ftrl_handle *h, int which, int64_t row0, int64_t n_rows, float *lin, float *vec) {
  const Dims &d = h->dims;
  if (row0 < 0 || n_rows < 0 || row0 + n_rows > d.n_feats) throw ArgFail{"row range out of bounds"};
  if (h->G > 1 && !h->attached) throw StateFail{"multi-GPU handle: call ftrl_attach_peers first"};
  const int plane = plane_of(which);
  FTRL_CUDA(cudaStreamSynchronize(h->compute));
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(16 << 20) / std::max<int32_t>(1, d.row_len));
  h->xfer.ensure((size_t)std::min<int64_t>(n_rows, chunk_rows) * std::max<int32_t>(1, d.row_len));
  for (int64_t r = 0; r < n_rows; r += chunk_rows) {
    const int64_t nr = std::min(chunk_rows, n_rows - r);
    if (lin) {
      k_lin_gather<<<(unsigned)((nr + 255) / 256), 256, 0, h->compute>>>(h->shards, h->xfer.p, row0 + r, nr, plane);
      FTRL_CUDA(cudaMemcpyAsync(lin + r, h->xfer.p, sizeof(float) * nr, cudaMemcpyDeviceToHost, h->compute));
      FTRL_CUDA(cudaStreamSynchronize(h->compute));
    }
    if (vec && d.row_len) {
      const int64_t q = nr * d.row_len;
      k_plane_gather<<<(unsigned)((q + 255) / 256), 256, 0, h->compute>>>(h->shards, h->xfer.p, row0 + r, nr, d.row_len, d.ld, plane);
      FTRL_CUDA(cudaMemcpyAsync(vec + r * d.row_len, h->xfer.p, sizeof(float) * q, cudaMemcpyDeviceToHost, h->compute));
      FTRL_CUDA(cudaStreamSynchronize(h->compute));
    }
  }
  FTRL_CUDA(cudaGetLastError());
}

// rows [row0, row0 + n_rows) of the whole model arrive in `lin` / `vec`; this rank keeps the rows it owns
static void set_rows_owned(ftrl_handle *h, int which, int64_t row0, int64_t n_rows, const float *lin, const float *vec) {
  if (h->G == 1) {
    xfer_rows(h, which, row0, n_rows, const_cast<float *>(lin), const_cast<float *>(vec), false);
    return;
  }
  const int64_t G = h->G, rl = h->dims.row_len;
  int64_t g0 = row0 + ((h->rank - row0 % G) % G + G) % G;  // first global row >= row0 owned by this rank
  if (g0 >= row0 + n_rows) return;
  const int64_t n_mine = (row0 + n_rows - g0 + G - 1) / G;
  std::vector<float> l2, v2;
  if (lin) {
    l2.resize((size_t)n_mine);
    for (int64_t i = 0; i < n_mine; i++) l2[(size_t)i] = lin[g0 - row0 + i * G];
  }
  if (vec && rl) {
    v2.resize((size_t)(n_mine * rl));
    for (int64_t i = 0; i < n_mine; i++) memcpy(&v2[(size_t)(i * rl)], vec + (g0 - row0 + i * G) * rl, sizeof(float) * (size_t)rl);
  }
  xfer_rows(h, which, g0 / G, n_mine, lin ? l2.data() : nullptr, (vec && rl) ? v2.data() : nullptr, false);
}

static void xfer_bias(ftrl_handle *h, int which, float *v, bool get) {
  if (!v) return;
  FTRL_CUDA(cudaStreamSynchronize(h->compute));
  float *p = reinterpret_cast<float *>(h->bias) + plane_of(which);
  if (get) FTRL_CUDA(cudaMemcpy(v, p, sizeof(float), cudaMemcpyDeviceToHost));
  else FTRL_CUDA(cudaMemcpy(p, v, sizeof(float), cudaMemcpyHostToDevice));
}

int ftrl_get_rows(ftrl_handle *h, int which, int64_t row0, int64_t n_rows, float *lin, float *vec) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] { xfer_rows(h, which, row0, n_rows, lin, vec, true); });
}
int ftrl_set_rows(ftrl_handle *h, int which, int64_t row0, int64_t n_rows, const float *lin, const float *vec) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] { xfer_rows(h, which, row0, n_rows, const_cast<float *>(lin), const_cast<float *>(vec), false); });
}
int ftrl_get_weights(ftrl_handle *h, float *bias, float *lin_w, float *vec_w) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    xfer_bias(h, 0, bias, true);
    xfer_rows(h, 0, 0, h->n_local, lin_w, vec_w, true);
  });
}
int ftrl_set_weights(ftrl_handle *h, const float *bias, const float *lin_w, const float *vec_w) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    xfer_bias(h, 0, const_cast<float *>(bias), false);
    xfer_rows(h, 0, 0, h->n_local, const_cast<float *>(lin_w), const_cast<float *>(vec_w), false);
  });
}
int ftrl_get_state(ftrl_handle *h, int which, float *bias_s, float *lin_s, float *vec_s) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (which != 1 && which != 2) throw ArgFail{"which must be 1 (n) or 2 (z)"};
    xfer_bias(h, which, bias_s, true);
    xfer_rows(h, which, 0, h->n_local, lin_s, vec_s, true);
  });
}
int ftrl_set_state(ftrl_handle *h, int which, const float *bias_s, const float *lin_s, const float *vec_s) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (which != 1 && which != 2) throw ArgFail{"which must be 1 (n) or 2 (z)"};
    xfer_bias(h, which, const_cast<float *>(bias_s), false);
    xfer_rows(h, which, 0, h->n_local, const_cast<float *>(lin_s), const_cast<float *>(vec_s), false);
  });
}

int ftrl_has_zero_weights(ftrl_handle *h, int *out) {
  if (!h || !out) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    const Dims &d = h->dims;
    FTRL_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), h->compute));
    const int64_t q = h->n_local * (d.row_len + 1);
    k_has_zero<<<(unsigned)((q + 255) / 256), 256, 0, h->compute>>>(h->tab, h->lin, h->n_local, d.row_len, d.ld, h->d_err);
    FTRL_CUDA(cudaGetLastError());
    int32_t f = 0;
    FTRL_CUDA(cudaMemcpyAsync(&f, h->d_err, sizeof(f), cudaMemcpyDeviceToHost, h->compute));
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
    FTRL_CUDA(cudaMemset(h->d_err, 0, sizeof(int32_t)));
    float b = 0.f;
    xfer_bias(h, 0, &b, true);
    *out = f ? 1 : 0;
  });
}

int ftrl_randomize_state(ftrl_handle *h, uint64_t seed, float z_scale, float n_lo, float n_hi) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (!(n_lo >= 0.f) || !(n_hi >= n_lo)) throw ArgFail{"need 0 <= n_lo <= n_hi"};
    const Dims &d = h->dims;
    const int64_t per_row = d.ld > 0 ? d.ld / 4 : 1;
    const int64_t q = h->n_local * per_row;
    k_randomize<<<(unsigned)((q + 255) / 256), 256, 0, h->compute>>>(h->tab, h->lin, h->bias, h->n_local, d.row_len, d.ld,
                                                                    seed, z_scale, n_lo, n_hi, h->G, h->rank);
    FTRL_CUDA(cudaGetLastError());
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
  });
}

// ---- evaluation metric -----------------------------------------------------------------------
int ftrl_eval_auc_device(ftrl_handle *h, int64_t n, const float *scores, const int32_t *label, double *auc_out) {
  if (!h || !auc_out) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (n > 0 && (!scores || !label)) throw ArgFail{"scores/label is NULL"};
    *auc_out = auc_device(h, n, scores, label);
  });
}
int ftrl_eval_auc(ftrl_handle *h, int64_t n, const float *scores, const int32_t *label, double *auc_out) {
  if (!h || !auc_out) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (n > 0 && (!scores || !label)) throw ArgFail{"scores/label is NULL"};
    if (n < 0) throw ArgFail{"negative size"};
    DevBuf<float> ds;
    DevBuf<int32_t> dl;
    ds.alloc((size_t)n);
    dl.alloc((size_t)n);
    if (n > 0) {
      FTRL_CUDA(cudaMemcpyAsync(ds.p, scores, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->compute));
      FTRL_CUDA(cudaMemcpyAsync(dl.p, label, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->compute));
    }
    *auc_out = auc_device(h, n, ds.p, dl.p);
  });
}

// ---- model files ----------------------------------------------------------------------------
int ftrl_save_model(ftrl_handle *h, const char *path, int compress_level) {
  if (!h || !path) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    // multi-GPU: ONE attached rank writes the file, reading every shard through peer memory
    const Dims &d = h->dims;
    const uint64_t total = sizeof(float) * (1ull + (uint64_t)d.n_feats + (uint64_t)d.n_feats * d.row_len);
    ModelWriter w(path, compress_level, total);
    float b = 0.f;
    xfer_bias(h, 0, &b, true);
    w.write(&b, sizeof(float));
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(16 << 20) / std::max<int32_t>(1, d.row_len));
    std::vector<float> buf((size_t)chunk_rows * std::max<int32_t>(1, d.row_len));
    for (int64_t r = 0; r < d.n_feats; r += chunk_rows) {
      const int64_t nr = std::min<int64_t>(chunk_rows, d.n_feats - r);
      get_rows_global(h, 0, r, nr, buf.data(), nullptr);
      w.write(buf.data(), sizeof(float) * nr);
    }
    if (d.row_len)
      for (int64_t r = 0; r < d.n_feats; r += chunk_rows) {
        const int64_t nr = std::min<int64_t>(chunk_rows, d.n_feats - r);
        get_rows_global(h, 0, r, nr, nullptr, buf.data());
        w.write(buf.data(), sizeof(float) * nr * d.row_len);
      }
    const uint64_t csize = w.finish();
    printf("saving to %s, before: %zu -> after: %zu\n", path, (size_t)total, (size_t)csize);
  });
}

int ftrl_load_model(ftrl_handle *h, const char *path) {
  if (!h || !path) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    // multi-GPU: EVERY rank reads the file and keeps the rows it owns (and the replicated bias)
    const Dims &d = h->dims;
    const uint64_t total = sizeof(float) * (1ull + (uint64_t)d.n_feats + (uint64_t)d.n_feats * d.row_len);
    ModelReader rd(path);
    if (rd.content_size() != total)
      throw IoFail{fmt("%s: payload is %llu bytes, model expects %llu", path, (unsigned long long)rd.content_size(), (unsigned long long)total)};
    float b = 0.f;
    rd.read(&b, sizeof(float));
    xfer_bias(h, 0, &b, false);
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(16 << 20) / std::max<int32_t>(1, d.row_len));
    std::vector<float> buf((size_t)chunk_rows * std::max<int32_t>(1, d.row_len));
    for (int64_t r = 0; r < d.n_feats; r += chunk_rows) {
      const int64_t nr = std::min<int64_t>(chunk_rows, d.n_feats - r);
      rd.read(buf.data(), sizeof(float) * nr);
      set_rows_owned(h, 0, r, nr, buf.data(), nullptr);
    }
    if (d.row_len)
      for (int64_t r = 0; r < d.n_feats; r += chunk_rows) {
        const int64_t nr = std::min<int64_t>(chunk_rows, d.n_feats - r);
        rd.read(buf.data(), sizeof(float) * nr * d.row_len);
        set_rows_owned(h, 0, r, nr, nullptr, buf.data());
      }
    printf("loading from %s, before: %zu -> after: %zu \n", path, (size_t)rd.file_size(), (size_t)total);
  });
}

int ftrl_save_model_text(ftrl_handle *h, const char *path) {
  if (!h || !path) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    const Dims &d = h->dims;
    std::vector<float> lin((size_t)d.n_feats), vec((size_t)d.n_feats * d.row_len);
    float b = 0.f;
    xfer_bias(h, 0, &b, true);
    get_rows_global(h, 0, 0, d.n_feats, lin.data(), vec.empty() ? nullptr : vec.data());
    save_text_model(path, b, lin.data(), vec.data(), d.n_feats, d.row_len);
  });
}

int ftrl_load_model_text(ftrl_handle *h, const char *path) {
  if (!h || !path) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    const Dims &d = h->dims;
    std::vector<float> lin((size_t)d.n_feats), vec((size_t)d.n_feats * d.row_len);
    float b = 0.f;
    load_text_model(path, &b, lin.data(), vec.data(), d.n_feats, d.row_len);
    xfer_bias(h, 0, &b, false);
    set_rows_owned(h, 0, 0, d.n_feats, lin.data(), vec.empty() ? nullptr : vec.data());
  });
}

void *ftrl_alloc_pinned(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void ftrl_free_pinned(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- measurement hooks ----------------------------------------------------------------------
int ftrl_set_stream(ftrl_handle *h, void *cuda_stream) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
    if (h->own_compute && h->compute) FTRL_CUDA(cudaStreamDestroy(h->compute));
    h->compute = static_cast<cudaStream_t>(cuda_stream);
    h->own_compute = false;
  });
}

int ftrl_profile_enable(ftrl_handle *h, int on) {
  if (!h) return FTRL_ERR_ARG;
  h->profiling = on != 0;
  return FTRL_OK;
}
int ftrl_profile_reset(ftrl_handle *h) {
  if (!h) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
    drain_profile(h);
    for (auto &p : h->phases) {
      p.ms = 0.0;
      p.launches = 0;
    }
  });
}
int ftrl_profile_read(ftrl_handle *h, int i, const char **name, double *ms, int64_t *launches) {
  if (!h || i < 0 || i >= PH_COUNT) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (!h->pending.empty()) {
      FTRL_CUDA(cudaStreamSynchronize(h->compute));
      drain_profile(h);
    }
    if (name) *name = h->phases[i].name;
    if (ms) *ms = h->phases[i].ms;
    if (launches) *launches = h->phases[i].launches;
  });
}

__global__ void k_batch_stats(int32_t nnz, uint32_t sentinel, const uint32_t *skey, const SegScan *scan,
                              const uint8_t *occ_single, const int32_t *n_chunks, int64_t *out) {
  // out: [0] nnz_valid, [1] n_unique, [2] n_single, [3] n_chunks
  __shared__ unsigned long long sh[3];
  if (threadIdx.x < 3) sh[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long v = 0, u = 0, s1 = 0;
  for (int32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x) {
    if (skey[p] != sentinel) {
      v++;
      if (scan[p].start == p) u++;
    }
    if (occ_single && occ_single[p]) s1++;
  }
  atomicAdd(&sh[0], v);
  atomicAdd(&sh[1], u);
  atomicAdd(&sh[2], s1);
  __syncthreads();
  if (threadIdx.x < 3) atomicAdd(reinterpret_cast<unsigned long long *>(out) + threadIdx.x, sh[threadIdx.x]);
  if (blockIdx.x == 0 && threadIdx.x == 0) out[3] = *n_chunks;
}

// sharded runs: distinct rows this rank OWNS among the rows the global batch touches = runs of the sorted
// contribution list
__global__ void k_count_runs(int32_t cap, const int32_t *n_sel, uint32_t lsent, const uint32_t *ckey, int64_t *out) {
  const int32_t n = min(*n_sel, cap);
  unsigned long long u = 0;
  for (int32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
    if (ckey[c] != lsent && (c == 0 || ckey[c - 1] != ckey[c])) u++;
  if (u) atomicAdd(reinterpret_cast<unsigned long long *>(out), u);
}

int ftrl_last_batch_stats(ftrl_handle *h, ftrl_batch_stats *out) {
  if (!h || !out) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    *out = h->stats;
    if (h->cfg.mode != FTRL_MODE_BATCH || h->nnz_cap == 0 || h->stats.n_rows == 0) return;
    // recomputed from the workspace of the last batch (diagnostics only, not on the hot path)
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
    DevBuf<int64_t> tmp;
    tmp.alloc(4);
    FTRL_CUDA(cudaMemset(tmp.p, 0, sizeof(int64_t) * 4));
    int32_t nnz = (int32_t)h->last_nnz;
    if (nnz > 0) {
      const bool have_single = h->dims.model_type == FTRL_FFM && h->fuse;
      k_batch_stats<<<256, 256, 0, h->compute>>>(nnz, (uint32_t)h->dims.n_feats, h->skey.p, h->scan.p,
                                                 have_single ? h->fused_sorted.p : nullptr, h->n_chunks.p, tmp.p);
      FTRL_CUDA(cudaGetLastError());
    }
    if (h->G > 1) {
      // n_unique: rows this rank owns (each distinct row of the global batch is counted on exactly one rank)
      FTRL_CUDA(cudaMemsetAsync(tmp.p + 1, 0, sizeof(int64_t), h->compute));
      k_count_runs<<<256, 256, 0, h->compute>>>((int32_t)h->ow_cap, h->n_sel.p, (uint32_t)h->n_local, h->ckey.p, tmp.p + 1);
      FTRL_CUDA(cudaGetLastError());
    }
    int64_t r[4] = {0, 0, 0, 0};
    FTRL_CUDA(cudaMemcpyAsync(r, tmp.p, sizeof(r), cudaMemcpyDeviceToHost, h->compute));
    FTRL_CUDA(cudaStreamSynchronize(h->compute));
    out->nnz_valid = r[0];
    out->n_unique = r[1];
    out->n_fused_rows = r[2];
    out->n_segmented_rows = r[1] - r[2];
    out->n_chunks = r[3];
  });
}

// ---- multi-GPU ------------------------------------------------------------------------------
static void peer_buffers(ftrl_handle *h, void **ptrs) {
  void *mine[PEER_BUFS] = {h->tab, h->lin, h->ukey.p, h->uinfo.p, h->umask.p, h->dst_at.p, h->inbox.p, h->inbox_lin.p, h->sync,
                           h->rc_w.p, h->rc_lin.p};
  memcpy(ptrs, mine, sizeof(mine));
}

int ftrl_export_peer_blob(ftrl_handle *h, void *blob) {
  if (!h || !blob) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (h->G <= 1) throw StateFail{"not a multi-GPU handle (world_size <= 1)"};
    PeerBlob pb;
    memset(&pb, 0, sizeof(pb));
    pb.magic = 0xF7B20003u;
    pb.rank = h->rank;
    pb.world = h->G;
    pb.device = h->cfg.device;
    pb.pid = (int64_t)getpid();
    pb.nnz_cap = h->nnz_cap;
    pb.ow_cap = h->ow_cap;
    void *ptrs[PEER_BUFS];
    peer_buffers(h, ptrs);
    for (int i = 0; i < PEER_BUFS; i++) {
      pb.raw[i] = ptrs[i];
      if (!ptrs[i]) throw StateFail{"multi-GPU buffers are not allocated"};
      FTRL_CUDA(cudaIpcGetMemHandle(&pb.ipc[i], ptrs[i]));
    }
    memset(blob, 0, FTRL_PEER_BLOB_BYTES);
    memcpy(blob, &pb, sizeof(pb));
  });
}

static void wire_peer(ftrl_handle *, Shards &sh, Peers &pr, Export &ex, int q, void *const *ptr) {
  sh.tab[q] = static_cast<float *>(ptr[0]);
  sh.lin[q] = static_cast<float4 *>(ptr[1]);
  pr.ukey[q] = static_cast<const uint32_t *>(ptr[2]);
  pr.uinfo[q] = static_cast<const uint32_t *>(ptr[3]);
  pr.umask[q] = static_cast<const unsigned long long *>(ptr[4]);
  pr.dst_at[q] = static_cast<int32_t *>(ptr[5]);
  ex.inbox[q] = static_cast<float *>(ptr[6]);
  ex.inbox_lin[q] = static_cast<float2 *>(ptr[7]);
  pr.sync[q] = static_cast<SyncArea *>(ptr[8]);
  pr.rc_w[q] = static_cast<float *>(ptr[9]);
  pr.rc_lin[q] = static_cast<float *>(ptr[10]);
}

static void set_rowspace(ftrl_handle *h, int log2G, int rank, int G) {
  RowSpace &rsp = h->rowspace;
  rsp = RowSpace{};
  rsp.tab = h->tab;
  rsp.lin = h->lin;
  rsp.staging = h->staging.p;
  rsp.staging_lin = h->staging_lin.p;
  rsp.rc_w = h->rc_w.p;
  rsp.rc_lin = h->rc_lin.p;
  rsp.log2G = log2G;
  rsp.rank = rank;
  rsp.Gm1 = G - 1;
}

int ftrl_attach_peers(ftrl_handle *h, const void *blobs) {
  if (!h || !blobs) return FTRL_ERR_ARG;
  return guarded(h, [&] {
    if (h->G <= 1) throw StateFail{"not a multi-GPU handle (world_size <= 1)"};
    if (h->attached) throw StateFail{"peers already attached"};
    Shards sh{};
    Peers pr{};
    Export ex{};
    sh.G = pr.G = h->G;
    sh.log2G = pr.log2G = h->log2G;
    sh.rank = pr.rank = h->rank;
    ex.on = 1;
    ex.log2G = h->log2G;
    ex.Gm1 = h->G - 1;
    ex.dst_at = h->dst_at.p;
    for (int q = 0; q < h->G; q++) {
      PeerBlob pb;
      memcpy(&pb, static_cast<const char *>(blobs) + (size_t)q * FTRL_PEER_BLOB_BYTES, sizeof(pb));
      if (pb.magic != 0xF7B20003u || pb.rank != q || pb.world != h->G) throw ArgFail{fmt("peer blob %d is not from rank %d of %d", q, q, h->G)};
      if (pb.nnz_cap != h->nnz_cap) throw ArgFail{"all ranks must use the same max_batch_nnz"};
      void *ptr[PEER_BUFS];
      if (q == h->rank) {
        peer_buffers(h, ptr);
      } else if (pb.pid == (int64_t)getpid()) {
        // same process (several handles in one process): plain pointers, peer access if another device
        if (pb.device != h->cfg.device) {
          int can = 0;
          FTRL_CUDA(cudaDeviceCanAccessPeer(&can, h->cfg.device, pb.device));
          if (!can) throw StateFail{fmt("device %d cannot access device %d", h->cfg.device, pb.device)};
          cudaError_t e = cudaDeviceEnablePeerAccess(pb.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FTRL_CUDA(e);
          cudaGetLastError();
        }
        memcpy(ptr, pb.raw, sizeof(ptr));
      } else {
        for (int i = 0; i < PEER_BUFS; i++) {
          FTRL_CUDA(cudaIpcOpenMemHandle(&ptr[i], pb.ipc[i], cudaIpcMemLazyEnablePeerAccess));
          h->ipc_opened.push_back(ptr[i]);
        }
      }
      wire_peer(h, sh, pr, ex, q, ptr);
    }
    // Dry run of one complete sharded step against this rank alone (one sample whose features are all out
    // of range: no row is touched; the bias is restored afterwards).  CUDA loads kernels lazily and defers
    // a first-time load while another kernel is running -- with spinning device barriers that turns into a
    // stall, so every kernel of the step is loaded here, before the first real step.
    {
      Shards self_sh{};
      Peers me{};
      Export self_ex{};
      self_sh.G = me.G = 1;
      self_ex.on = 1;
      self_ex.dst_at = h->dst_at.p;
      void *mine[PEER_BUFS];
      peer_buffers(h, mine);
      // the dry run synchronises on a PRIVATE SyncArea: a peer that finished attaching earlier may already be
      // writing its first step's flag / row counts into this rank's real one, which must not be reset here
      DevBuf<SyncArea> scratch_sync;
      scratch_sync.alloc(1);
      FTRL_CUDA(cudaMemset(scratch_sync.p, 0, sizeof(SyncArea)));
      mine[8] = scratch_sync.p;
      wire_peer(h, self_sh, me, self_ex, 0, mine);
      const int G = h->G, log2G = h->log2G, rank = h->rank;
      h->shards = self_sh;
      h->peers = me;
      h->exportd = self_ex;
      set_rowspace(h, 0, 0, 1);
      h->attached = true;
      Slot &ws = h->slots[0];
      const int64_t rp[2] = {0, 2};
      const int32_t bad[2] = {-1, -1}, lab = 0;
      const float one[2] = {1.f, 1.f};
      float4 bias_save;
      FTRL_CUDA(cudaMemcpy(&bias_save, h->bias, sizeof(float4), cudaMemcpyDeviceToHost));
      FTRL_CUDA(cudaMemcpy(ws.row_ptr.p, rp, sizeof(rp), cudaMemcpyHostToDevice));
      FTRL_CUDA(cudaMemcpy(ws.field.p, bad, sizeof(bad), cudaMemcpyHostToDevice));
      FTRL_CUDA(cudaMemcpy(ws.feat.p, bad, sizeof(bad), cudaMemcpyHostToDevice));
      FTRL_CUDA(cudaMemcpy(ws.val.p, one, sizeof(one), cudaMemcpyHostToDevice));
      FTRL_CUDA(cudaMemcpy(ws.label.p, &lab, sizeof(lab), cudaMemcpyHostToDevice));
      Batch wb{1, 2, ws.row_ptr.p, ws.field.p, ws.feat.p, ws.val.p, ws.label.p};
      h->G = 1;  // one shard: this rank against itself
      // both parities: the dry run loads every kernel and leaves the step counter even
      for (int rep = 0; rep < 2; rep++) {
        if (h->precise) train_device_sharded<true>(h, wb, nullptr, ws.loss.p, nullptr, false);
        else train_device_sharded<false>(h, wb, nullptr, ws.loss.p, nullptr, false);
      }
      FTRL_CUDA(cudaStreamSynchronize(h->compute));
      h->G = G; h->log2G = log2G; h->rank = rank;
      FTRL_CUDA(cudaMemcpy(h->bias, &bias_save, sizeof(float4), cudaMemcpyHostToDevice));
      FTRL_CUDA(cudaMemset(h->d_err, 0, sizeof(int32_t)));
      h->epoch = h->epoch_id = 0;  // the real SyncArea (zeroed by ftrl_create) has not been touched
      h->shard_step = 0;
    }
    h->shards = sh;
    h->peers = pr;
    h->exportd = ex;
    set_rowspace(h, h->log2G, h->rank, h->G);
    h->attached = true;
  });
}

}  // extern "C"

// ---- debug probe (tools/ only, not part of the ABI): scattered row traffic against one shard ----
namespace ftrl {
__global__ void k_dbg_peer_traffic(Shards sh, Export ex, int q, int64_t n_local, int64_t ld, int32_t n_rows, int flags, float *sink) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31, nvec = (int)(ld >> 2);
  float acc = 0.f;
  for (int64_t i = warp; i < n_rows; i += nw) {
    const int64_t row = (int64_t)(((uint64_t)i * 2654435761ull) % (uint64_t)n_local);
    if (flags & 1) {
      const float4 *p = reinterpret_cast<const float4 *>(sh.tab[q] + row * 3 * ld + 2 * ld);
      for (int v = lane; v < nvec; v += 32) acc += __ldcs(p + v).x;
    }
    if (flags & 2) {
      float4 *p = reinterpret_cast<float4 *>(ex.inbox[q] + i * ld);
      for (int v = lane; v < nvec; v += 32) __stcs(p + v, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (flags & 4) {
      const float4 *p = reinterpret_cast<const float4 *>(ex.inbox[q] + (row % n_rows) * ld);
      for (int v = lane; v < nvec; v += 32) acc += __ldcs(p + v).x;
    }
    if (flags & 8) {  // destroys the w plane: probe only
      float4 *p = reinterpret_cast<float4 *>(sh.tab[q] + row * 3 * ld + 2 * ld);
      for (int v = lane; v < nvec; v += 32) __stcs(p + v, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (flags & 16) {  // sequential rows instead of hashed
      const float4 *p = reinterpret_cast<const float4 *>(sh.tab[q] + (i % n_local) * 3 * ld + 2 * ld);
      for (int v = lane; v < nvec; v += 32) acc += __ldcs(p + v).x;
    }
  }
  if (acc == 12345.f) sink[0] = acc;
}
}  // namespace ftrl
extern "C" int ftrl_dbg_peer_traffic(ftrl_handle *h, int q, int flags, int n_rows, int grid, int reps, float *ms_out) {
  return guarded(h, [&] {
    cudaEvent_t a, b;
    FTRL_CUDA(cudaEventCreate(&a));
    FTRL_CUDA(cudaEventCreate(&b));
    FTRL_CUDA(cudaEventRecord(a, h->compute));
    for (int r = 0; r < reps; r++)
      ftrl::k_dbg_peer_traffic<<<grid, 256, 0, h->compute>>>(h->shards, h->exportd, q, h->n_local, h->dims.ld, n_rows, flags, h->logit_ws.p);
    FTRL_CUDA(cudaEventRecord(b, h->compute));
    FTRL_CUDA(cudaEventSynchronize(b));
    FTRL_CUDA(cudaEventElapsedTime(ms_out, a, b));
    *ms_out /= reps;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  });
}
