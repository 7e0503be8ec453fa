// grid barrier with 40 words: effect of (a) delaying the first poll, (b) a pause between polls,
// (c) polling one word first.  148 CTAs x 352 threads, words 1 KB apart.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_bar(unsigned long long *ctr, long long *cyc, int rounds, int NWORDS, int delay, int pause, int onefirst) {
  const int stride = 128;
  long long t0 = clock64();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < rounds; ++r) {
    unsigned long long *base = ctr + (size_t)(r & 1) * NWORDS * stride;
    if (threadIdx.x < NWORDS)
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(base + (size_t)threadIdx.x * stride), "l"(1ull) : "memory");
    if (warp == 0) {
      const unsigned long long target = (unsigned long long)(r / 2 + 1) * gridDim.x;
      long long ts = clock64();
      while (clock64() - ts < delay) {}
      if (onefirst) {
        if (lane == 0) {
          unsigned long long v;
          do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(base + (size_t)(NWORDS - 1) * stride) : "memory");
               if (v < target && pause) { long long tp = clock64(); while (clock64() - tp < pause) {} } } while (v < target);
        }
        __syncwarp();
      }
      for (int w = lane; w < NWORDS; w += 32) {
        unsigned long long v;
        do {
          asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(base + (size_t)w * stride) : "memory");
          if (v < target && pause) { long long tp = clock64(); while (clock64() - tp < pause) {} }
        } while (v < target);
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / rounds;
}
int main() {
  long long *cyc; unsigned long long *g; size_t bytes = 16 << 20;
  cudaMalloc(&cyc, 64); cudaMalloc(&g, bytes);
  int rounds = 2000, nw = 40;
  for (int onefirst : {0, 1})
    for (int pause : {0, 100, 300})
      for (int delay : {0, 300, 600, 900, 1200}) {
        cudaMemset(g, 0, bytes);
        void *args[] = {&g, &cyc, &rounds, &nw, &delay, &pause, &onefirst};
        cudaLaunchCooperativeKernel((void *)k_bar, dim3(148), dim3(352), args, 0, 0);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("onefirst %d pause %3d delay %4d: %lld cycles/round (%s)\n", onefirst, pause, delay, h, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
