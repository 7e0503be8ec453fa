// Round-2 experiment (not part of the product): the per-round statistic exchange of the persistent
// kernel done as an in-switch reduction -- every CTA of every GPU issues multimem.red.add.u64 on the
// words of a multicast object (NVLS), every GPU's copy receives all contributions, and the arrival
// count rides in the top bits of the word exactly as in the single-GPU barrier.  Compared with the
// round-1 scheme (local barrier, THEN CTA 0 forwards totals over NVLink: 2 557 cycles on top of the
// local barrier) the NVLink hop overlaps the local reduction instead of following it.
//
// One process, GPUs 0..n-1 (default 2), driver-API multicast objects.  Prints cycles per round for
// 1 and 40 words (1 KB apart), each CTA arriving once per word and round, all CTAs polling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_multimem ubench_multimem.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("%s -> %s\n", #x, s_); exit(1); } } while (0)
#define CR(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); exit(1); } } while (0)

constexpr int CNT_SHIFT = 52;  // 12 count bits: up to 8 x 148 arrivals per round
constexpr int STRIDE = 128;    // u64 words between two statistics (1 KB)

__global__ void rounds(unsigned long long *mc, const unsigned long long *uc, int words, int iters, unsigned long long expect,
                       long long *cyc, unsigned long long *sums) {
  const int lane = threadIdx.x;
  unsigned long long prev[2] = {0, 0};
  // start together: the first round doubles as the start-up barrier (peers launch at different times)
  long long t0 = 0;
  for (int i = 0; i <= iters; ++i) {
    const int par = i & 1;
    if (i == 1 && blockIdx.x == 0 && lane == 0) t0 = clock64();
    if (lane < words) {
      const size_t w = (size_t)(par * 64 + lane) * STRIDE;
      const unsigned long long v = (1ull << CNT_SHIFT) + (unsigned long long)(blockIdx.x + 1);
      asm volatile("multimem.red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(mc + w), "l"(v) : "memory");
      unsigned long long d;
      do {
        unsigned long long a;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(uc + w) : "memory");
        d = a - prev[par];
      } while ((d >> CNT_SHIFT) != expect);
      prev[par] += d;
      if (i == iters && blockIdx.x == 0) sums[lane] = d & ((1ull << CNT_SHIFT) - 1);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && lane == 0) *cyc = (clock64() - t0) / iters;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2;
  CK(cuInit(0));
  std::vector<CUdevice> dev(n);
  for (int d = 0; d < n; ++d) {
    CR(cudaSetDevice(d));
    CR(cudaFree(0));
    CK(cuDeviceGet(&dev[d], d));
    int mcs = 0;
    CK(cuDeviceGetAttribute(&mcs, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[d]));
    if (!mcs) { printf("device %d: multicast not supported\n", d); return 0; }
  }
  size_t bytes = (size_t)2 * 64 * STRIDE * 8;  // two parity sets of up to 64 words
  CUmulticastObjectProp mp = {};
  mp.numDevices = n;
  mp.handleTypes = 0;
  mp.size = bytes;
  size_t gran = 0, agran = 0;
  CK(cuMulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  CUmemAllocationProp ap = {};
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = 0;
  CK(cuMemGetAllocationGranularity(&agran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (agran > gran) gran = agran;
  bytes = (bytes + gran - 1) / gran * gran;
  mp.size = bytes;
  CUmemGenericAllocationHandle mch;
  CK(cuMulticastCreate(&mch, &mp));
  for (int d = 0; d < n; ++d) CK(cuMulticastAddDevice(mch, dev[d]));
  std::vector<CUmemGenericAllocationHandle> mem(n);
  std::vector<CUdeviceptr> uc(n), mc(n);
  for (int d = 0; d < n; ++d) {
    CR(cudaSetDevice(d));
    ap.location.id = d;
    CK(cuMemCreate(&mem[d], bytes, &ap, 0));
    CK(cuMulticastBindMem(mch, 0, mem[d], 0, bytes, 0));
    CUmemAccessDesc ad = {};
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = d;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemAddressReserve(&uc[d], bytes, gran, 0, 0));
    CK(cuMemMap(uc[d], bytes, 0, mem[d], 0));
    CK(cuMemSetAccess(uc[d], bytes, &ad, 1));
    CK(cuMemAddressReserve(&mc[d], bytes, gran, 0, 0));
    CK(cuMemMap(mc[d], bytes, 0, mch, 0));
    CK(cuMemSetAccess(mc[d], bytes, &ad, 1));
  }
  std::vector<long long *> cyc(n);
  std::vector<unsigned long long *> sums(n);
  for (int d = 0; d < n; ++d) { CR(cudaSetDevice(d)); CR(cudaMalloc(&cyc[d], 8)); CR(cudaMalloc(&sums[d], 64 * 8)); }
  const int G = 148;
  for (int words : {1, 20, 40}) {
    int iters = 2000;
    for (int d = 0; d < n; ++d) { CR(cudaSetDevice(d)); CR(cudaMemset((void *)uc[d], 0, bytes)); CR(cudaDeviceSynchronize()); }
    unsigned long long expect = (unsigned long long)G * n;
    for (int d = 0; d < n; ++d) {
      CR(cudaSetDevice(d));
      unsigned long long *m = (unsigned long long *)mc[d];
      const unsigned long long *u = (const unsigned long long *)uc[d];
      void *args[] = {&m, &u, &words, &iters, &expect, &cyc[d], &sums[d]};
      CR(cudaLaunchCooperativeKernel((void *)rounds, dim3(G), dim3(64), args, 0, 0));
    }
    printf("%d GPUs, %2d words:", n, words);
    for (int d = 0; d < n; ++d) {
      long long h; unsigned long long s;
      CR(cudaSetDevice(d)); CR(cudaDeviceSynchronize());
      CR(cudaMemcpy(&h, cyc[d], 8, cudaMemcpyDeviceToHost)); CR(cudaMemcpy(&s, sums[d], 8, cudaMemcpyDeviceToHost));
      printf("  gpu%d %lld cycles/round (sum %llu, expected %llu)", d, h, s, (unsigned long long)n * G * (G + 1) / 2);
    }
    printf("\n");
  }
  return 0;
}
