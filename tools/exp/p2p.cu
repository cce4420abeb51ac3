// peer-memory bandwidth probe: GPU0 kernels reading / writing GPU1 memory, contiguous and in scattered 1248-byte chunks
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__global__ void rd(const float4 *src, float4 *sink, size_t n_chunks, int vpc, int scatter, int unroll_dummy) {
  // each warp handles chunks of vpc float4 (vpc*16 bytes); chunk order contiguous or pseudo-random
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t c = warp; c < n_chunks; c += nw) {
    const size_t cc = scatter ? (c * 2654435761ull) % n_chunks : c;
    const float4 *p = src + cc * vpc;
    for (int v = lane; v < vpc; v += 32) { const float4 t = __ldcs(p + v); acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
  }
  if (acc.x == 12345.f) sink[0] = acc;
}

// limited-concurrency probe: 148 CTAs x 128 threads, every lane keeps LU scattered 16-byte loads in flight (like the
// loader warps of k_ffm_tile); effective latency = bytes in flight / bandwidth
__global__ void rd_lowconc(const float4 *src, float4 *sink, size_t n_chunks, int vpc, int iters) {
  constexpr int LU = 14;
  const int t = threadIdx.x;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; it++) {
    float4 buf[LU];
#pragma unroll
    for (int u = 0; u < LU; u++) {
      const size_t item = ((size_t)it * gridDim.x + blockIdx.x) * (LU * 128) + u * 128 + t;
      const size_t c = ((item / vpc) * 2654435761ull) % n_chunks;
      buf[u] = __ldcs(src + c * vpc + item % vpc);
    }
#pragma unroll
    for (int u = 0; u < LU; u++) { acc.x += buf[u].x; acc.y += buf[u].y; }
  }
  if (acc.x == 12345.f) sink[0] = acc;
}
__global__ void wr(float4 *dst, size_t n_chunks, int vpc, int scatter) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t c = warp; c < n_chunks; c += nw) {
    const size_t cc = scatter ? (c * 2654435761ull) % n_chunks : c;
    float4 *p = dst + cc * vpc;
    for (int v = lane; v < vpc; v += 32) __stcs(p + v, make_float4(1, 2, 3, 4));
  }
}
int main() {
  int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("can access peer: %d\n", can);
  const size_t bytes = 4ull << 30; const int vpc = 78; const size_t n_chunks = bytes / (vpc * 16);
  float4 *remote, *local, *sink;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, bytes)); CK(cudaMemset(remote, 0, bytes));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&local, bytes)); CK(cudaMalloc(&sink, 64)); CK(cudaMemset(local, 0, bytes));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int tgt = 0; tgt < 2; tgt++) for (int scatter = 0; scatter < 2; scatter++) for (int w = 0; w < 2; w++) for (int blocks = 148 * 4; blocks <= 148 * 16; blocks *= 4) {
    float4 *p = tgt ? remote : local;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      if (w) wr<<<blocks, 256>>>(p, n_chunks, vpc, scatter); else rd<<<blocks, 256>>>(p, sink, n_chunks, vpc, scatter, 0);
      cudaEventRecord(b); CK(cudaEventSynchronize(b));
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%s %s %s blocks=%d: %.1f GB/s\n", tgt ? "REMOTE" : "local ", scatter ? "scattered-1248B" : "contiguous     ", w ? "write" : "read ", blocks, bytes / ms / 1e6);
  }
  for (int tgt = 0; tgt < 2; tgt++) for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a); rd_lowconc<<<148, 128>>>(tgt ? remote : local, sink, n_chunks, vpc, 400); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double by = 148.0 * 128 * 14 * 16 * 400;
    printf("%s same-process low-concurrency scattered read: %.1f GB/s -> effective latency %.2f us\n", tgt ? "REMOTE" : "local ", by / ms / 1e6, 148.0 * 128 * 14 * 16 / (by / ms / 1e3) * 1e-3 * 1e3);
  }
  return 0;
}
