// how does remote (NVLink peer) scattered-read throughput scale with warps per SM vs loads in flight per lane?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
template <int LU>
__global__ void rd(const float4 *src, float4 *sink, size_t n_chunks, int vpc, int iters) {
  const int t = threadIdx.x, nt = blockDim.x;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; it++) {
    float4 buf[LU];
#pragma unroll
    for (int u = 0; u < LU; u++) {
      const size_t item = ((size_t)it * gridDim.x + blockIdx.x) * (size_t)(LU * nt) + u * nt + t;
      const size_t c = ((item / vpc) * 2654435761ull) % n_chunks;
      buf[u] = __ldcs(src + c * vpc + item % vpc);
    }
#pragma unroll
    for (int u = 0; u < LU; u++) { acc.x += buf[u].x; acc.y += buf[u].y; }
  }
  if (acc.x == 12345.f) sink[0] = acc;
}
template <int LU>
void run(const float4 *p, float4 *sink, size_t n_chunks, int threads, const char *what) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4000 / LU * 128 / threads + 1;
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a); rd<LU><<<148, threads>>>(p, sink, n_chunks, 78, iters); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&ms, a, b);
  }
  const double by = 148.0 * threads * LU * 16 * iters;
  printf("%s threads/SM=%4d loads/lane=%2d in-flight/SM=%5.1f KB: %7.1f GB/s\n", what, threads, LU, threads * LU * 16 / 1024.0, by / ms / 1e6);
}
int main() {
  const size_t bytes = 8ull << 30; const size_t n_chunks = bytes / (78 * 16);
  float4 *remote, *local, *sink;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, bytes)); CK(cudaMemset(remote, 0, bytes));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&local, bytes)); CK(cudaMalloc(&sink, 64)); CK(cudaMemset(local, 0, bytes));
  for (int tgt = 1; tgt >= 0; tgt--) {
    const float4 *p = tgt ? remote : local; const char *w = tgt ? "REMOTE" : "local ";
    run<14>(p, sink, n_chunks, 128, w); run<4>(p, sink, n_chunks, 128, w); run<1>(p, sink, n_chunks, 128, w);
    run<8>(p, sink, n_chunks, 256, w); run<4>(p, sink, n_chunks, 512, w); run<2>(p, sink, n_chunks, 1024, w);
    run<1>(p, sink, n_chunks, 1024, w); run<4>(p, sink, n_chunks, 1024, w); run<8>(p, sink, n_chunks, 1024, w);
  }
  return 0;
}
