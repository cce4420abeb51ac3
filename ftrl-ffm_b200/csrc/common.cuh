// common.cuh -- device helpers shared by every kernel of the FTRL LR/FM/FFM path (sm_100a).
//
// HBM layout (chosen for this chip, not the reference's vector<vector<float>>):
//   lin  : float4[n_feats]        {z, n, w, 0}   one 16-byte sector-friendly record per feature
//   tab  : float [n_feats][3][ld] planes {z, n, w}, ld = round_up(row_len, 4) so every plane of
//          every row is 16-byte aligned (128-bit loads / cp.async.bulk); a row is one contiguous
//          3*ld*4-byte span (FFM F=39,k=8: 3744 B), z and n adjacent because training reads z,n
//          and writes z,n,w.
//   bias : float4 {z, n, w, 0}
// Reference semantics restated here: maybe_zero_weight (src/include/model/ftrl_model.h:29-33),
// sgn/sigmoid (src/include/utils/utils.h:16-23), loss (src/include/eval/loss.h:8-12).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ftrl {

struct Hyper {
  float alpha, beta, l1, l2, inv_alpha;
};

struct Dims {
  int32_t model_type;  // 0 LR, 1 FM, 2 FFM
  int32_t n_feats, n_fields, k;
  int32_t row_len;  // 0 | k | n_fields*k
  int32_t ld;       // padded row length (floats)
};

enum : int { PLANE_Z = 0, PLANE_N = 1, PLANE_W = 2 };

// ---- feature-sharded tables -------------------------------------------------------------------
// Row `feat` lives on shard feat mod G at local row feat div G (G a power of two, G = 1: one GPU).
// Peer shards are mapped into this process (CUDA IPC) and reached over NVLink with ordinary loads / stores.
constexpr int MAX_SHARDS = 8;
struct Shards {   // every shard's tables (predict reads rows where they live)
  int G, log2G, rank, pad;
  float *tab[MAX_SHARDS];
  float4 *lin[MAX_SHARDS];
  __device__ __forceinline__ float *row(int32_t feat, int64_t rs) const {
    return tab[feat & (G - 1)] + (int64_t)(feat >> log2G) * rs;
  }
  __device__ __forceinline__ float4 *linp(int32_t feat) const { return lin[feat & (G - 1)] + (feat >> log2G); }
};

// What the per-sample training kernel sees: the local shard, the per-step cache of remote rows (their w
// plane, pushed by its owner once per distinct row: shard.cuh) and the local staging area of gradient images.
// A row is named by a locator: >= 0 local row index; < 0: -1 - (sorted head position of the remote row).
struct RowSpace {
  float *tab;            // [n_local][3][ld]
  float4 *lin;           // [n_local]
  float *staging;        // [sorted position][ld]
  float *staging_lin;    // [sorted position]
  const float *rc_w;     // [sorted head position][ld]   (null when G == 1)
  const float *rc_lin;   // [sorted head position]
  int log2G, rank, Gm1, pad;
};

// Where the reduced (sum g, sum g^2) of a row goes (k_ffm_staged_rows / k_ffm_combine).  Single GPU: applied
// in place.  Sharded: dst_at[sorted head position] = -2 apply here (this rank owns the row and is its only
// contributor), >= 0: slot in the owner's inbox.
struct Export {
  int on, log2G, Gm1, pad;
  const int32_t *dst_at;
  float *inbox[MAX_SHARDS];        // [slot][2][ld]
  float2 *inbox_lin[MAX_SHARDS];   // [slot]
};

constexpr int32_t KEY_INVALID_BITS = 0;  // invalid occurrences get key == n_feats (sorts last)

// ---------------------------------------------------------------------------------------------
// math: two flavours.  PRECISE = IEEE sqrt/div (round-to-nearest) but free association/FMA;
// fast = MUFU approximations (sqrt.approx, rcp.approx), ~2 ulp.
// The reference-exact (sequential) kernels do not use these; see exact.cuh.
// ---------------------------------------------------------------------------------------------
template <bool PRECISE>
__device__ __forceinline__ float f_sqrt(float x) {
  if (PRECISE) return __fsqrt_rn(x);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <bool PRECISE>
__device__ __forceinline__ float f_div(float a, float b) {
  if (PRECISE) return __fdiv_rn(a, b);
  return __fdividef(a, b);
}

// w = W(n, z) given sq = sqrt(n)                                 (ftrl_model.h:29-33)
template <bool PRECISE>
__device__ __forceinline__ float weight_from(float z, float sq, const Hyper &h) {
  const float den = PRECISE ? (h.l2 + f_div<true>(h.beta + sq, h.alpha)) : fmaf(h.beta + sq, h.inv_alpha, h.l2);
  const float s = z > 0.f ? h.l1 : -h.l1;  // sgn(0) = -1 (utils.h:16-18)
  const float w = f_div<PRECISE>(s - z, den);
  return fabsf(z) <= h.l1 ? 0.f : w;
}

// telescoped per-coordinate FTRL update:  n' = n + sg2 ; z' = (z + sg) - w (sqrt n' - sqrt n)/alpha
template <bool PRECISE>
__device__ __forceinline__ void ftrl_apply(float &z, float &n, float w, float sg, float sg2, const Hyper &h) {
  const float n_new = n + sg2;
  const float d = f_sqrt<PRECISE>(n_new) - f_sqrt<PRECISE>(n);
  const float sigma = PRECISE ? f_div<true>(d, h.alpha) : d * h.inv_alpha;
  z = (z + sg) - sigma * w;
  n = n_new;
}

// same with sq = sqrt(n) already at hand
template <bool PRECISE>
__device__ __forceinline__ void ftrl_apply_sq(float &z, float &n, float sq, float w, float sg, float sg2, const Hyper &h) {
  const float n_new = n + sg2;
  const float d = f_sqrt<PRECISE>(n_new) - sq;
  const float sigma = PRECISE ? f_div<true>(d, h.alpha) : d * h.inv_alpha;
  z = (z + sg) - sigma * w;
  n = n_new;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// eval/loss.h:8-12 in fp64, no clipping
__device__ __forceinline__ double logloss_d(int y, float logit) {
  const double s = 1.0 / (1.0 + exp(-(double)logit));
  return -(double)y * log(s) - (double)(1 - y) * log(1.0 - s);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; every thread gets the result.  scratch: >= 33 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? scratch[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// ---------------------------------------------------------------------------------------------
// vector access helpers (VEC = 1, 2 or 4 floats)
// ---------------------------------------------------------------------------------------------
template <int VEC>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float *p) { v[0] = *p; }
  __device__ __forceinline__ void store(float *p) const { *p = v[0]; }
};
template <>
struct Vec<2> {
  float v[2];
  __device__ __forceinline__ void load(const float *p) {
    const float2 t = *reinterpret_cast<const float2 *>(p);
    v[0] = t.x; v[1] = t.y;
  }
  __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float *p) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float *p) const {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: Gaussian init of w on device (the reference draws every weight
// from a fresh std::random_device-seeded mt19937, utils.h:31-36 -- there is no stream to match,
// only the distribution N(init_mean, init_stddev)).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// four N(0,1) samples for counter idx
__device__ __forceinline__ float4 gaussian4(uint64_t idx, uint64_t seed, uint32_t stream) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), stream, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
  const float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
  const float ra = sqrtf(-2.0f * logf(fminf(fmaxf(u0, 1e-12f), 1.0f)));
  const float rb = sqrtf(-2.0f * logf(fminf(fmaxf(u2, 1e-12f), 1.0f)));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

}  // namespace ftrl
