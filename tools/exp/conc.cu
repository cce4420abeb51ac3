// do two kernels in two streams overlap on this box?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <chrono>
__global__ void busy(long long cycles, int *sink) { const long long t0 = clock64(); while (clock64() - t0 < cycles) {} if (sink) *sink = 1; }
__global__ void spin(volatile unsigned *flag, long long timeout, int *err) {
  const long long t0 = clock64();
  while (*flag == 0) { if (clock64() - t0 > timeout) { *err = 1; break; } }
}
__global__ void setflag(volatile unsigned *flag) { *flag = 1; __threadfence_system(); }
int main(int argc, char **argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  cudaStream_t a, b;
  cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  unsigned *flag; int *err; cudaMalloc(&flag, 4); cudaMalloc(&err, 4); cudaMemset(flag, 0, 4); cudaMemset(err, 0, 4);
  // preload every kernel (lazy module loading) and warm up
  busy<<<1, 32, 0, a>>>(1000, nullptr); setflag<<<1, 1, 0, b>>>(flag); spin<<<1, 32, 0, a>>>(flag, 1000, err);
  cudaDeviceSynchronize(); cudaMemset(flag, 0, 4); cudaMemset(err, 0, 4); cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  if (variant == 0) { busy<<<1, 32, 0, a>>>(2000000000ll, nullptr); busy<<<1, 32, 0, b>>>(2000000000ll, nullptr); }
  if (variant == 1) { spin<<<1, 32, 0, a>>>(flag, 4000000000ll, err); setflag<<<1, 1, 0, b>>>(flag); }
  cudaDeviceSynchronize();
  double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  int herr = 0; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
  printf("variant %d: total %.4fs timeout=%d\n", variant, t, herr);
  return 0;
}
