// grid barrier with replicated words: CTA c adds to replica c % S of each of its 40 words; the
// polling lane reads all S replicas of its word IN PARALLEL and sums them.  148 CTAs x 352 threads.
#include <cstdio>
#include <cuda_runtime.h>
template <int S>
__global__ void k_bar(unsigned long long *ctr, long long *cyc, int rounds, int NWORDS, int stride) {
  long long t0 = clock64();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < rounds; ++r) {
    unsigned long long *base = ctr + (size_t)(r & 1) * NWORDS * S * stride;
    if (threadIdx.x < NWORDS)
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(base + ((size_t)threadIdx.x * S + blockIdx.x % S) * stride), "l"(1ull) : "memory");
    if (warp == 0) {
      const unsigned long long target = (unsigned long long)(r / 2 + 1) * gridDim.x;
      for (int w = lane; w < NWORDS; w += 32) {
        unsigned long long v;
        do {
          unsigned long long t[S];
#pragma unroll
          for (int s = 0; s < S; ++s) t[s] = *(volatile unsigned long long *)(base + ((size_t)w * S + s) * stride);
          v = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) v += t[s];
        } while (v < target);
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / rounds;
}
template <int S> void run(unsigned long long *g, long long *cyc, size_t bytes, int nw, int stride) {
  int rounds = 2000;
  cudaMemset(g, 0, bytes);
  void *args[] = {&g, &cyc, &rounds, &nw, &stride};
  cudaLaunchCooperativeKernel((void *)k_bar<S>, dim3(148), dim3(352), args, 0, 0);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("words %2d replicas %2d stride %4d B: %lld cycles/round (%s)\n", nw, S, stride * 8, h, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long *cyc; unsigned long long *g; size_t bytes = 64 << 20;
  cudaMalloc(&cyc, 64); cudaMalloc(&g, bytes);
  for (int nw : {20, 40}) for (int stride : {128, 160, 544, 1056, 2080, 4128, 16416, 65568})
    run<1>(g, cyc, bytes, nw, stride);
  return 0;
}
