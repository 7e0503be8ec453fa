// NVLink ping-pong between two B200s: store a counter into the peer's memory, the peer polls its
// local copy and answers.  Reports the round-trip time per exchange for several store/load flavours.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void pingpong(volatile unsigned long long *local, unsigned long long *remote, int iters, int me, long long *cyc) {
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (me == 0) {
      if (MODE == 0) *(volatile unsigned long long *)remote = i;
      else if (MODE == 1) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote), "l"((unsigned long long)i) : "memory"); }
      else { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote), "l"((unsigned long long)i) : "memory"); __threadfence_system(); }
      unsigned long long v;
      do { if (MODE == 0) v = *local; else asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(local) : "memory"); } while (v < (unsigned long long)i);
    } else {
      unsigned long long v;
      do { if (MODE == 0) v = *local; else asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(local) : "memory"); } while (v < (unsigned long long)i);
      if (MODE == 0) *(volatile unsigned long long *)remote = i;
      else if (MODE == 1) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote), "l"((unsigned long long)i) : "memory"); }
      else { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote), "l"((unsigned long long)i) : "memory"); __threadfence_system(); }
    }
  }
  *cyc = clock64() - t0;
}
template <int MODE> void run(unsigned long long *b0, unsigned long long *b1, long long *c0, long long *c1) {
  int iters = 2000;
  cudaSetDevice(0); cudaMemset(b0, 0, 64); cudaSetDevice(1); cudaMemset(b1, 0, 64); cudaDeviceSynchronize(); cudaSetDevice(0); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  pingpong<MODE><<<1, 1>>>(b0, b1, iters, 0, c0);
  cudaEventRecord(e1);
  cudaSetDevice(1); pingpong<MODE><<<1, 1>>>(b1, b0, iters, 1, c1);
  cudaDeviceSynchronize(); cudaSetDevice(0); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, c0, 8, cudaMemcpyDeviceToHost);
  printf("mode %d: round trip %.2f us (%lld cycles) -> one way ~%.2f us  [%s]\n", MODE, 1e3 * ms / iters, h / iters, 0.5e3 * ms / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  unsigned long long *b0, *b1; long long *c0, *c1;
  cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0); cudaMalloc(&b0, 64); cudaMalloc(&c0, 8);
  cudaSetDevice(1); cudaDeviceEnablePeerAccess(0, 0); cudaMalloc(&b1, 64); cudaMalloc(&c1, 8);
  run<0>(b0, b1, c0, c1); run<1>(b0, b1, c0, c1); run<2>(b0, b1, c0, c1);
  return 0;
}
