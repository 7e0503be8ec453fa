// grid-barrier variants on B200: S sub-counters per barrier (CTA c adds to sub-counter c % S,
// pollers read all S and sum), NW independent words handled by NW lanes of the polling warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_bar(unsigned long long *ctr, long long *cyc, int rounds, int S, int NWORDS, int stride) {
  // ctr layout: [parity][word][sub] each `stride` u64 apart
  long long t0 = clock64();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < rounds; ++r) {
    unsigned long long *base = ctr + (size_t)(r & 1) * NWORDS * S * stride;
    // publish: thread w < NWORDS adds its word
    if (threadIdx.x < NWORDS)
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(base + ((size_t)threadIdx.x * S + blockIdx.x % S) * stride), "l"(1ull) : "memory");
    if (warp == 0) {
      for (int w = lane; w < NWORDS; w += 32) {
        unsigned long long target = (unsigned long long)(r / 2 + 1) * gridDim.x, v;
        do {
          v = 0;
          for (int s = 0; s < S; ++s) {
            unsigned long long t;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(t) : "l"(base + ((size_t)w * S + s) * stride) : "memory");
            v += t;
          }
        } while (v < target);
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / rounds;
}
int main() {
  long long *cyc; unsigned long long *g; size_t bytes = 64 << 20;
  cudaMalloc(&cyc, 64); cudaMalloc(&g, bytes);
  int rounds = 2000;
  for (int stride : {1, 16, 128})
    for (int nw : {1, 40})
      for (int S : {1, 2, 4, 8, 16}) {
        cudaMemset(g, 0, bytes);
        void *args[] = {&g, &cyc, &rounds, &S, &nw, &stride};
        cudaLaunchCooperativeKernel((void *)k_bar, dim3(148), dim3(352), args, 0, 0);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("stride %4d B words %2d sub-counters %2d: %lld cycles/round (%s)\n", stride * 8, nw, S, h, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
