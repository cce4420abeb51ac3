// symmetric two-GPU peer-memory probe: BOTH GPUs run kernels at the same time (like the sharded tile kernel),
// mode 0 = two processes + legacy cudaIpc, mode 1 = one process, two devices (cudaDeviceEnablePeerAccess).
// tests: A one side reads remote, B both read remote, C both write remote, D both read+write remote,
//        E rank 0 reads remote while rank 1 streams its own HBM, F = D plus a local stream on both
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>
#include <sys/mman.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// flags: 1 read src, 2 write dst, 4 local stream (read lsrc, write ldst)
__global__ void traffic(const float4 *src, float4 *dst, const float4 *lsrc, float4 *ldst, float4 *sink, size_t n_chunks,
                        int vpc, int flags) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t c = warp; c < n_chunks; c += nw) {
    const size_t cc = (c * 2654435761ull) % n_chunks;
    if (flags & 1) {
      const float4 *p = src + cc * vpc;
      for (int v = lane; v < vpc; v += 32) { const float4 t = __ldcs(p + v); acc.x += t.x; acc.y += t.y; }
    }
    if (flags & 2) {
      float4 *p = dst + ((cc + 7) % n_chunks) * vpc;
      for (int v = lane; v < vpc; v += 32) __stcs(p + v, make_float4(1, 2, 3, 4));
    }
    if (flags & 8) {
      const size_t n_rows = n_chunks / 3;
      const float4 *p = src + ((c * 2654435761ull) % n_rows) * 3 * vpc + 2 * vpc;
      for (int v = lane; v < vpc; v += 32) { const float4 t = __ldcs(p + v); acc.x += t.x; acc.y += t.y; }
    }
    if (flags & 4) {
      const float4 *p = lsrc + cc * vpc;
      float4 *q = ldst + cc * vpc;
      for (int v = lane; v < vpc; v += 32) { const float4 t = __ldcs(p + v); __stcs(q + v, t); }
    }
  }
  if (acc.x == 12345.f) sink[0] = acc;
}

struct Shared { volatile int arrive[32][2]; cudaIpcMemHandle_t h[2][2]; volatile float ms[16][2]; };
static size_t BYTES = 8ull << 30;
static const int VPC = 78;

static void barrier(Shared *sh, int idx, int r) {
  __sync_synchronize();
  sh->arrive[idx][r] = 1;
  while (!sh->arrive[idx][1 - r]) {}
  __sync_synchronize();
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  if (argc > 2) BYTES = (size_t)(atof(argv[2]) * (1ull << 30));
  const int small_first = argc > 3 ? atoi(argv[3]) : 0;
  const size_t n_chunks = BYTES / (VPC * 16);
  struct T { const char *name; int f0, f1; } tests[] = {
      {"A r0 reads remote, r1 idle          ", 1, 0}, {"B both read remote                  ", 1, 1},
      {"C both write remote                 ", 2, 2}, {"D both read+write remote            ", 3, 3},
      {"E r0 reads remote, r1 local stream  ", 1, 4}, {"F both read+write remote + local    ", 7, 7},
      {"G r0 read+write remote, r1 idle     ", 3, 0}, {"H both read remote rows stride 3744 ", 8, 8}};
  const int n_tests = sizeof(tests) / sizeof(tests[0]);
  if (mode == 0) {
    Shared *sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset((void *)sh, 0, sizeof(Shared));
    pid_t child = fork();
    const int r = child == 0 ? 1 : 0;
    CK(cudaSetDevice(r)); CK(cudaFree(0));
    if (small_first) { void *t; CK(cudaMalloc(&t, 80 << 20)); CK(cudaMalloc(&t, 16)); CK(cudaMalloc(&t, 4)); CK(cudaMalloc(&t, 5000)); }
    float4 *a, *b, *sink; CK(cudaMalloc(&a, BYTES)); CK(cudaMalloc(&b, BYTES)); CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(a, 0, BYTES)); CK(cudaMemset(b, 0, BYTES));
    CK(cudaIpcGetMemHandle(&sh->h[r][0], a)); CK(cudaIpcGetMemHandle(&sh->h[r][1], b));
    CK(cudaDeviceSynchronize());
    barrier(sh, 0, r);
    float4 *ra, *rb;
    CK(cudaIpcOpenMemHandle((void **)&ra, sh->h[1 - r][0], cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle((void **)&rb, sh->h[1 - r][1], cudaIpcMemLazyEnablePeerAccess));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // warm every variant (lazy module load)
    traffic<<<8, 256>>>(ra, rb, a, b, sink, 1024, VPC, 7); CK(cudaDeviceSynchronize());
    for (int t = 0; t < n_tests; t++) {
      const int f = r ? tests[t].f1 : tests[t].f0;
      barrier(sh, 1 + t, r);
      float ms = 0;
      if (f) {
        cudaEventRecord(e0);
        traffic<<<148 * 8, 256>>>(ra, rb, a, b, sink, n_chunks, VPC, f);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
      }
      sh->ms[t][r] = ms;
      barrier(sh, 16 + t, r);
      if (r == 0) {
        printf("2-proc legacy-IPC  %s:", tests[t].name);
        for (int q = 0; q < 2; q++) {
          const int fq = q ? tests[t].f1 : tests[t].f0;
          if (fq) printf("  r%d %.1f ms (%.0f GB/s per stream)", q, sh->ms[t][q], BYTES / sh->ms[t][q] / 1e6);
        }
        printf("\n");
      }
    }
    if (r == 0) waitpid(child, nullptr, 0);
    return 0;
  }
  // mode 1: one process
  float4 *a[2], *b[2], *sink[2]; cudaEvent_t e0[2], e1[2];
  for (int r = 0; r < 2; r++) {
    CK(cudaSetDevice(r)); CK(cudaDeviceEnablePeerAccess(1 - r, 0));
    CK(cudaMalloc(&a[r], BYTES)); CK(cudaMalloc(&b[r], BYTES)); CK(cudaMalloc(&sink[r], 64));
    CK(cudaMemset(a[r], 0, BYTES)); CK(cudaMemset(b[r], 0, BYTES));
    cudaEventCreate(&e0[r]); cudaEventCreate(&e1[r]);
  }
  for (int r = 0; r < 2; r++) { CK(cudaSetDevice(r)); traffic<<<8, 256>>>(a[1 - r], b[1 - r], a[r], b[r], sink[r], 1024, VPC, 7); CK(cudaDeviceSynchronize()); }
  for (int t = 0; t < n_tests; t++) {
    for (int r = 0; r < 2; r++) {
      const int f = r ? tests[t].f1 : tests[t].f0;
      if (!f) continue;
      CK(cudaSetDevice(r));
      cudaEventRecord(e0[r]);
      traffic<<<148 * 8, 256>>>(a[1 - r], b[1 - r], a[r], b[r], sink[r], n_chunks, VPC, f);
      cudaEventRecord(e1[r]);
    }
    printf("1-proc peer-access %s:", tests[t].name);
    for (int r = 0; r < 2; r++) {
      const int f = r ? tests[t].f1 : tests[t].f0;
      if (!f) continue;
      CK(cudaSetDevice(r)); CK(cudaEventSynchronize(e1[r]));
      float ms; cudaEventElapsedTime(&ms, e0[r], e1[r]);
      printf("  r%d %.1f ms (%.0f GB/s per stream)", r, ms, BYTES / ms / 1e6);
    }
    printf("\n");
  }
  return 0;
}
