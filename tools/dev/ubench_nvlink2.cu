// NVLink exchange as used by the persistent kernel, in isolation: on each GPU CTA 0 / warp 0 stores
// 20 tagged 16-byte pairs into the peer, all CTAs x 20 lanes poll the local copy.  Round trip per
// exchange for different numbers of polling CTAs.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void xch(unsigned long long *local, unsigned long long *remote, int iters, long long *cyc, int pollers) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (warp == 0 && lane < 20) {
      unsigned long long *slotL = local + (i & 1) * 64, *slotR = remote + (i & 1) * 64;
      if (blockIdx.x == 0)
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slotR + 2 * lane), "l"((unsigned long long)i), "l"((unsigned long long)i) : "memory");
      if ((int)blockIdx.x < pollers) {
        unsigned long long a, b;
        do { asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(slotL + 2 * lane) : "memory"); } while (a < (unsigned long long)i || b < (unsigned long long)i);
      }
    }
    __syncthreads();
    // CTAs that do not poll still have to stay in step with CTA 0: a cheap grid-free way is to let
    // them read CTA 0's progress word
    if ((int)blockIdx.x >= pollers) {
      if (threadIdx.x == 0) { unsigned long long a; do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(local + 200) : "memory"); } while (a < (unsigned long long)i); }
      __syncthreads();
    } else if (blockIdx.x == 0 && threadIdx.x == 0) {
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(local + 200), "l"((unsigned long long)i) : "memory");
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = (clock64() - t0) / iters;
}
int main() {
  unsigned long long *b[2]; long long *c[2];
  for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaDeviceEnablePeerAccess(1 - d, 0); cudaMalloc(&b[d], 4096); cudaMalloc(&c[d], 8); }
  for (int pollers : {1, 8, 37, 148}) {
    int iters = 2000;
    for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaMemset(b[d], 0, 4096); cudaDeviceSynchronize(); }
    for (int d = 0; d < 2; ++d) {
      cudaSetDevice(d);
      void *args[] = {&b[d], &b[1 - d], &iters, &c[d], &pollers};
      cudaLaunchCooperativeKernel((void *)xch, dim3(148), dim3(288), args, 0, 0);
    }
    long long h[2];
    for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); cudaMemcpy(&h[d], c[d], 8, cudaMemcpyDeviceToHost); }
    printf("polling CTAs %3d: %lld / %lld cycles per exchange round (%s)\n", pollers, h[0], h[1], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
