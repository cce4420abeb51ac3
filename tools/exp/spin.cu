// concurrency probe: can a kernel spinning on a flag in stream A be released by work enqueued later in stream B?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <chrono>
__global__ void spin(volatile unsigned *flag, long long timeout, int *err) {
  const long long t0 = clock64();
  while (*flag == 0) { if (clock64() - t0 > timeout) { *err = 1; break; } }
}
__global__ void setflag(volatile unsigned *flag) { *flag = 1; }
__global__ void dummy(int *x) { if (x) *x = 1; }
int main(int argc, char **argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  unsigned *flag; int *err, *x;
  cudaMalloc(&flag, 4); cudaMalloc(&err, 4); cudaMalloc(&x, 4 << 20);
  cudaMemset(flag, 0, 4); cudaMemset(err, 0, 4);
  cudaStream_t a, b, c;
  cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking);
  cudaDeviceSynchronize();
  void *hp = malloc(4 << 20); void *pin; cudaMallocHost(&pin, 4 << 20);
  auto t0 = std::chrono::steady_clock::now();
  spin<<<1, 32, 0, a>>>(flag, 4000000000ll, err);
  if (variant == 1) cudaMemcpyAsync(x, hp, 4 << 20, cudaMemcpyHostToDevice, c);     // pageable H2D on a third stream
  if (variant == 2) cudaMemcpyAsync(x, pin, 4 << 20, cudaMemcpyHostToDevice, c);    // pinned H2D
  if (variant == 3) cudaMemsetAsync(x, 1, 4, b);
  if (variant == 4) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); cudaMemcpyAsync(x, hp, 1 << 20, cudaMemcpyHostToDevice, c); cudaEventRecord(e, c); cudaStreamWaitEvent(b, e, 0); }
  if (variant == 5) { void *p2; cudaMalloc(&p2, 64 << 20); }
  if (variant == 6) { void *p2; cudaMallocHost(&p2, 1 << 20); }
  if (variant == 7) { cudaMemcpyAsync(pin, x, 8, cudaMemcpyDeviceToHost, a); }  // D2H queued behind the spinner in stream a
  setflag<<<1, 1, 0, b>>>(flag);
  double tenq = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  cudaDeviceSynchronize();
  double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  int herr = 0; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
  printf("variant %d: enqueue %.4fs total %.4fs timeout=%d (%s)\n", variant, tenq, t, herr, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
