// B200 micro-benchmarks that ground the kernel design: FP64 dependent-issue latency and
// throughput, shuffle / shared-memory latency, L2 atomic + relaxed-load round trips.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int CH>
__global__ void k_dfma(double *out, long long *cyc, double a, double b) {
  double x[CH];
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-9 + c;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(double *out, long long *cyc) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_rcp(double *out, long long *cyc) {
  double x = 1.5 + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.25; }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_l2(unsigned long long *g, long long *cyc) {
  // one thread: red.add then spin until visible (round trip), then plain relaxed load latency
  unsigned long long v = 0;
  long long t0 = clock64();
  for (int i = 0; i < 256; ++i) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(g), "l"(1ull) : "memory");
    do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(g) : "memory"); } while (v != (unsigned long long)(i + 1));
  }
  long long t1 = clock64();
  for (int i = 0; i < 256; ++i) { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(g + (v & 1)) : "memory"); }
  long long t2 = clock64();
  cyc[0] = (t1 - t0) / 256; cyc[1] = (t2 - t1) / 256; g[8] = v;
}
// grid barrier latency: all CTAs red.add a counter, spin until count reached; repeated R rounds
__global__ void k_gridbar(unsigned long long *ctr, long long *cyc, int rounds) {
  __shared__ unsigned long long now;
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) {
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(ctr + (r & 1) * 128), "l"(1ull) : "memory");
      unsigned long long target = (unsigned long long)(r / 2 + 1) * gridDim.x, v;
      do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr + (r & 1) * 128) : "memory"); } while (v < target);
      now = v;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / rounds;
}
int main() {
  double *out; long long *cyc; unsigned long long *g;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 64); cudaMalloc(&g, 4096); cudaMemset(g, 0, 4096);
  long long h[2];
#define RUN(CH, BLK, THR) k_dfma<CH><<<BLK, THR>>>(out, cyc, 1.0000001, 1e-9); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("DFMA chains=%d blocks=%d threads=%d: %.2f cycles per chain step (per warp-instr issue: %.2f)\n", CH, BLK, THR, (double)h[0] / N, (double)h[0] / N / CH);
  RUN(1, 1, 32) RUN(2, 1, 32) RUN(4, 1, 32) RUN(8, 1, 32) RUN(16, 1, 32)
  RUN(1, 1, 128) RUN(4, 1, 128) RUN(8, 1, 128) RUN(4, 1, 352) RUN(8, 1, 352) RUN(8, 1, 704) RUN(8, 148, 352)
  k_shfl<<<1, 32>>>(out, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("64-bit SHFL+DADD dependent: %.1f cycles\n", (double)h[0] / N);
  k_rcp<<<1, 32>>>(out, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("rcp.approx.f64 + DADD dependent: %.1f cycles\n", (double)h[0] / N);
  k_l2<<<1, 1>>>(g, cyc); cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost); printf("L2: red+poll round trip %lld cycles, relaxed load %lld cycles\n", h[0], h[1]);
  for (int thr : {32, 352}) {
    cudaMemset(g, 0, 4096);
    void *args[] = {&g, &cyc, nullptr}; int rounds = 2000; args[2] = &rounds;
    cudaLaunchCooperativeKernel((void *)k_gridbar, dim3(148), dim3(thr), args, 0, 0);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("grid barrier (148 CTAs x %d thr, 1 word): %lld cycles per round (%s)\n", thr, h[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
