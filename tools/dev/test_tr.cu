// dev harness: checks tsp::tr_reduce + the lane->statistic mapping on the GPU for several V
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tsgpu.h"
struct Params; struct PState;
namespace tsm {}
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// minimal stand-ins so that ts_persist.cuh's helpers compile
#define TS_TR_ONLY 1
namespace tsp {
template <int N, int V>
__device__ __forceinline__ void tr_level(double (&v)[V], const bool up, const int bit) {
  constexpr int LO = (N + 1) / 2, HI = N / 2;
#pragma unroll
  for (int i = 0; i < LO; ++i) {
    const double lo = v[i];
    const double hi = (i < HI) ? v[LO + i] : 0.0;
    const double send = up ? lo : hi;
    const double recv = __shfl_xor_sync(0xffffffffu, send, bit);
    v[i] = (up ? hi : lo) + recv;
  }
}
template <int V>
__device__ __forceinline__ void tr_reduce(double (&v)[V], const int lane) {
  constexpr int N1 = (V + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2;
  tr_level<V, V>(v, lane & 16, 16);
  tr_level<N1, V>(v, lane & 8, 8);
  tr_level<N2, V>(v, lane & 4, 4);
  tr_level<N3, V>(v, lane & 2, 2);
  tr_level<N4, V>(v, lane & 1, 1);
}
}
template <int V>
__global__ void k(double *out) {
  const int lane = threadIdx.x;
  double vv[V];
  for (int v = 0; v < V; ++v) vv[v] = (double)((lane * 131 + v * 7) % 97 + 1);
  tsp::tr_reduce<V>(vv, lane);
  int tr_start = 0, tr_len = V, nominal = V;
  for (int bit = 16; bit > 0; bit >>= 1) {
    const int lo = (nominal + 1) / 2;
    if (lane & bit) { tr_start += lo; tr_len = max(tr_len - lo, 0); } else tr_len = min(tr_len, lo);
    nominal = lo;
  }
  constexpr int VPL = (V + 31) / 32;
  for (int q = 0; q < VPL; ++q) if (q < tr_len) out[tr_start + q] = vv[q];
}
template <int V> int run() {
  double *d; cudaMalloc(&d, V * 8); cudaMemset(d, 0, V * 8);
  k<V><<<1, 32>>>(d);
  double h[V]; cudaMemcpy(h, d, V * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int v = 0; v < V; ++v) {
    double want = 0; for (int lane = 0; lane < 32; ++lane) want += (double)((lane * 131 + v * 7) % 97 + 1);
    if (h[v] != want) { bad++; printf("V=%d v=%d got %g want %g\n", V, v, h[v], want); }
  }
  printf("V=%d bad=%d (%s)\n", V, bad, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); return bad;
}
int main() { int b = 0; b += run<2>(); b += run<6>(); b += run<10>(); b += run<14>(); b += run<20>(); b += run<26>(); b += run<40>(); b += run<64>(); return b != 0; }
