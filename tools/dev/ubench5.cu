// Round-2 experiment: two-level grid barrier for the 40 fixed-point words of a round.
// Level 1: the CTAs of a group (gs consecutive CTAs) add into the group's words; the group's first
// CTA waits for them and adds the group total (arrival count 1) into the top-level words.
// Level 2: every CTA polls the top-level words until all groups have arrived.
// Fewer arrivals per address and in total at the cost of one more dependent L2 round trip.
// gs = 148 is the flat barrier of the product.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ubench5.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int STRIDE = 128;  // u64 words between statistics (1 KB)
__device__ __forceinline__ unsigned long long ldr(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red(unsigned long long *p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// mem: [parity 2][level-1 groups (<=148) + 1 top][64 words][STRIDE]
__global__ void k_bar2(unsigned long long *mem, long long *cyc, unsigned long long *sum, int rounds, int gs, int NW) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = gridDim.x, ngroups = (G + gs - 1) / gs, grp = blockIdx.x / gs;
  const int gsize = min(gs, G - grp * gs);
  const bool leader = (blockIdx.x % gs) == 0, flat = gs >= G;
  unsigned long long prevg[2][2] = {{0, 0}, {0, 0}}, prevt[2][2] = {{0, 0}, {0, 0}};
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const int par = r & 1;
    unsigned long long *lvl1 = mem + ((size_t)par * 149 + grp) * 64 * STRIDE;
    unsigned long long *top = mem + ((size_t)par * 149 + 148) * 64 * STRIDE;
    if (threadIdx.x < NW) red((flat ? top : lvl1) + (size_t)threadIdx.x * STRIDE, (1ull << 54) + blockIdx.x + 1);
    if (warp == 0) {
      for (int q = 0; q < 2; ++q) {
        const int w = lane + 32 * q;
        if (w >= NW) continue;
        if (!flat && leader) {
          unsigned long long d;
          do { d = ldr(lvl1 + (size_t)w * STRIDE) - prevg[par][q]; } while ((d >> 54) != (unsigned long long)gsize);
          prevg[par][q] += d;
          red(top + (size_t)w * STRIDE, (1ull << 54) + (d & ((1ull << 54) - 1)));
        }
        unsigned long long d;
        const unsigned long long want = flat ? G : ngroups;
        do { d = ldr(top + (size_t)w * STRIDE) - prevt[par][q]; } while ((d >> 54) != want);
        prevt[par][q] += d;
        if (r == rounds - 1 && blockIdx.x == 0 && w == 0) *sum = d & ((1ull << 54) - 1);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (clock64() - t0) / rounds;
}
int main() {
  long long *cyc; unsigned long long *g, *sum; const size_t bytes = (size_t)2 * 149 * 64 * STRIDE * 8;
  cudaMalloc(&cyc, 8); cudaMalloc(&sum, 8); cudaMalloc(&g, bytes);
  int rounds = 2000;
  for (int nw : {1, 40})
    for (int gs : {148, 37, 16, 12, 8, 4}) {
      cudaMemset(g, 0, bytes);
      void *args[] = {&g, &cyc, &sum, &rounds, &gs, &nw};
      cudaLaunchCooperativeKernel((void *)k_bar2, dim3(148), dim3(256), args, 0, 0);
      long long h; unsigned long long s;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&s, sum, 8, cudaMemcpyDeviceToHost);
      printf("words %2d group size %3d: %lld cycles/round, sum %llu (expected %d) %s\n", nw, gs, h, s, 148 * 149 / 2, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
