// cross-PROCESS peer memory probe: legacy cudaIpc mapping vs VMM (cuMemCreate + POSIX fd via pidfd_getfd)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/syscall.h>
#include <sys/wait.h>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char *s; cuGetErrorString(e, &s); printf("%s: %s\n", #x, s); exit(1); } } while (0)
__global__ void rd(const float4 *src, float4 *sink, size_t n_chunks, int vpc) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t c = warp; c < n_chunks; c += nw) {
    const size_t cc = (c * 2654435761ull) % n_chunks;
    const float4 *p = src + cc * vpc;
    for (int v = lane; v < vpc; v += 32) { const float4 t = __ldcs(p + v); acc.x += t.x; acc.y += t.y; }
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
__global__ void wr(float4 *dst, size_t n_chunks, int vpc) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t c = warp; c < n_chunks; c += nw) {
    const size_t cc = (c * 2654435761ull) % n_chunks;
    float4 *p = dst + cc * vpc;
    for (int v = lane; v < vpc; v += 32) __stcs(p + v, make_float4(1, 2, 3, 4));
  }
}
struct Msg { int mode; long pid; int fd; size_t size; cudaIpcMemHandle_t h; };
int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;  // 0 legacy ipc, 1 vmm
  const size_t bytes = 16ull << 30; const int vpc = 78; const size_t n_chunks = bytes / (vpc * 16);
  int p2c[2], c2p[2]; pipe(p2c); pipe(c2p);
  pid_t child = fork();
  if (child == 0) {  // owner of the memory: GPU 1
    CK(cudaSetDevice(1)); CK(cudaFree(0));
    Msg m; memset(&m, 0, sizeof(m)); m.mode = mode; m.pid = getpid(); m.size = bytes;
    if (mode == 0) { void *p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 0, bytes)); CK(cudaIpcGetMemHandle(&m.h, p)); }
    else {
      CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 1;
      prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t gran = 0; CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
      printf("vmm granularity %zu\n", gran);
      CUmemGenericAllocationHandle h; CU(cuMemCreate(&h, bytes, &prop, 0));
      CUdeviceptr va; CU(cuMemAddressReserve(&va, bytes, gran, 0, 0)); CU(cuMemMap(va, bytes, 0, h, 0));
      CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; CU(cuMemSetAccess(va, bytes, &acc, 1));
      CK(cudaMemset((void *)va, 0, bytes));
      CU(cuMemExportToShareableHandle(&m.fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    }
    CK(cudaDeviceSynchronize());
    write(c2p[1], &m, sizeof(m));
    char c; read(p2c[0], &c, 1);
    return 0;
  }
  Msg m; read(c2p[0], &m, sizeof(m));
  CK(cudaSetDevice(0)); CK(cudaFree(0));
  float4 *remote = nullptr;
  if (mode == 0) { CK(cudaIpcOpenMemHandle((void **)&remote, m.h, cudaIpcMemLazyEnablePeerAccess)); }
  else {
    int pidfd = (int)syscall(SYS_pidfd_open, (pid_t)m.pid, 0);
    int fd = (int)syscall(SYS_pidfd_getfd, pidfd, m.fd, 0);
    printf("pidfd %d fd %d\n", pidfd, fd);
    if (fd < 0) { perror("pidfd_getfd"); return 1; }
    CUmemGenericAllocationHandle h; CU(cuMemImportFromShareableHandle(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va; CU(cuMemAddressReserve(&va, m.size, 2 << 20, 0, 0)); CU(cuMemMap(va, m.size, 0, h, 0));
    CUmemAccessDesc acc = {}; acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(va, m.size, &acc, 1));
    remote = (float4 *)va;
  }
  float4 *sink; CK(cudaMalloc(&sink, 64));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 2; w++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      if (w) wr<<<148 * 8, 256>>>(remote, n_chunks, vpc); else rd<<<148 * 8, 256>>>(remote, sink, n_chunks, vpc);
      cudaEventRecord(b); CK(cudaEventSynchronize(b));
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("mode %s cross-process REMOTE scattered-1248B over %zu GB %s: %.1f GB/s\n", mode ? "VMM" : "legacy-IPC", bytes >> 30, w ? "write" : "read ", bytes / ms / 1e6);
  }
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a); rd_lowconc<<<148, 128>>>(remote, sink, n_chunks, vpc, 400); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double by = 148.0 * 128 * 14 * 16 * 400;
    printf("mode %s low-concurrency scattered read: %.1f GB/s -> effective latency %.2f us\n", mode ? "VMM" : "legacy-IPC", by / ms / 1e6, 148.0 * 128 * 14 * 16 / (by / ms / 1e3) * 1e-3 * 1e3);
  }
  char c = 1; write(p2c[1], &c, 1); waitpid(child, nullptr, 0);
  return 0;
}
