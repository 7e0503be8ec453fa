// Round-2 experiment (not part of the product): one round of the persistent kernel's statistic
// exchange in isolation, on n GPUs of one box, for every candidate scheme.  Each CTA contributes
// `words` fixed-point words per round (the product: 40 at K = 10); a round ends when every CTA of
// every GPU holds the grand total.  Prints cycles per round as seen by CTA 0 of every GPU.
//
//   A  product r1: local red.add barrier -> CTA 0 stores the GPU total (tagged) into its slot on
//      every peer (n unicast NVLink stores) -> every CTA polls n slots
//   B  as A, the n unicast stores replaced by ONE multimem.st (NVLS multicast store)
//   C  local barrier -> CTA 0 does ONE multimem.red.add into a global accumulator replicated on all
//      GPUs (count bits = ranks arrived) -> every CTA polls 1 word per statistic
//   D  no local barrier: every CTA of every GPU does multimem.red.add (count = CTAs x ranks)
//   E  two-level: groups of `gs` CTAs add into a group word, the group's leader CTA forwards the group
//      total with multimem.red.add (count = groups x ranks): the hop overlaps the rest of the local
//      reduction, and no word sees more than max(gs, groups x ranks) arrivals
//   F  as C without NVLS: CTA 0 does n unicast red.add.sys into the peers' global accumulators
//   L  single GPU reference: flat local barrier only (what one GPU pays per round)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_xchg ubench_xchg.cu -lcuda
//   ./ubench_xchg <ngpus> [rounds]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("%s -> %s\n", #x, s_); exit(1); } } while (0)
#define CR(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); exit(1); } } while (0)

constexpr int CNT_SHIFT = 52;            // 12 count bits
constexpr unsigned long long DMASK = (1ull << CNT_SHIFT) - 1;
constexpr int STRIDE = 128;              // u64 words between two statistics (1 KB)
constexpr int MAXW = 64, MAXR = 8, MAXG = 40;
// layout of one GPU's buffer, in u64 words
constexpr size_t OFF_LOCAL = 0;                                           // [2][MAXW][STRIDE]
constexpr size_t OFF_GLOBAL = OFF_LOCAL + (size_t)2 * MAXW * STRIDE;      // [2][MAXW][STRIDE]
constexpr size_t OFF_GROUP = OFF_GLOBAL + (size_t)2 * MAXW * STRIDE;      // [2][MAXG][MAXW][STRIDE]
constexpr size_t OFF_SLOT = OFF_GROUP + (size_t)2 * MAXG * MAXW * STRIDE; // [MAXR][2][MAXW] (adjacent words)
constexpr size_t TOTAL_WORDS = OFF_SLOT + (size_t)MAXR * 2 * MAXW;

struct Ptrs {
  unsigned long long *uc;         // this GPU's buffer
  unsigned long long *mc;         // multicast alias of the same offsets on all GPUs (or null)
  unsigned long long *peer[MAXR]; // every GPU's buffer (unicast, over NVLink)
};

__device__ __forceinline__ unsigned long long ld_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void mm_red(unsigned long long *p, unsigned long long v) {
  asm volatile("multimem.red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void mm_st(unsigned long long *p, unsigned long long v) {
  asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// variant: 'A'..'F', 'L'
__global__ void rounds(Ptrs P, int variant, int rank, int nranks, int words, int iters, int gs, long long *cyc,
                       unsigned long long *sums) {
  const int lane = threadIdx.x;
  const int G = gridDim.x;
  const int ngroups = (G + gs - 1) / gs, grp = blockIdx.x / gs, gsize = min(gs, G - grp * gs);
  const bool leader = blockIdx.x % gs == 0;
  unsigned long long prevL[2] = {0, 0}, prevG[2] = {0, 0}, prevGr[2] = {0, 0};
  long long t0 = 0;
  for (int i = 0; i <= iters; ++i) {  // round 0 doubles as the start-up barrier between the GPUs
    const int par = i & 1;
    if (i == 1 && blockIdx.x == 0 && lane == 0) t0 = clock64();
    if (lane < words) {
      const size_t w = (size_t)(par * MAXW + lane) * STRIDE;
      const unsigned long long one = 1ull << CNT_SHIFT;
      const unsigned long long val = one + (unsigned long long)(blockIdx.x + 1);
      const unsigned long long tag = (unsigned long long)((i + 1) & 1023) << CNT_SHIFT;
      unsigned long long total = 0;
      if (variant == 'D') {
        mm_red(P.mc + OFF_GLOBAL + w, val);
      } else if (variant == 'E') {
        unsigned long long *gw = P.uc + OFF_GROUP + ((size_t)(par * MAXG + grp) * MAXW + lane) * STRIDE;
        red_gpu(gw, val);
        if (leader) {
          unsigned long long d;
          do { d = ld_gpu(gw) - prevGr[par]; } while ((d >> CNT_SHIFT) != (unsigned long long)gsize);
          prevGr[par] += d;
          if (P.mc) mm_red(P.mc + OFF_GLOBAL + w, one + (d & DMASK));
          else for (int r = 0; r < nranks; ++r) red_sys(P.peer[r] + OFF_GLOBAL + w, one + (d & DMASK));
        }
      } else {
        red_gpu(P.uc + OFF_LOCAL + w, val);
        if (blockIdx.x == 0 || variant == 'L') {
          unsigned long long d;
          do { d = ld_gpu(P.uc + OFF_LOCAL + w) - prevL[par]; } while ((d >> CNT_SHIFT) != (unsigned long long)G);
          prevL[par] += d;
          d &= DMASK;
          total = d;
          if (variant == 'A') for (int r = 0; r < nranks; ++r) st_sys(P.peer[r] + OFF_SLOT + ((size_t)rank * 2 + par) * MAXW + lane, tag | d);
          if (variant == 'B') mm_st(P.mc + OFF_SLOT + ((size_t)rank * 2 + par) * MAXW + lane, tag | d);
          if (variant == 'C') mm_red(P.mc + OFF_GLOBAL + w, one + d);
          if (variant == 'F') for (int r = 0; r < nranks; ++r) red_sys(P.peer[r] + OFF_GLOBAL + w, one + d);
        }
      }
      // every CTA waits for the grand total
      if (variant == 'A' || variant == 'B') {
        total = 0;
        for (int r = 0; r < nranks; ++r) {
          unsigned long long s;
          do { s = ld_sys(P.uc + OFF_SLOT + ((size_t)r * 2 + par) * MAXW + lane); } while ((s & ~DMASK) != tag);
          total += s & DMASK;
        }
      } else if (variant != 'L') {
        const unsigned long long want = variant == 'D' ? (unsigned long long)G * nranks
                                      : variant == 'E' ? (unsigned long long)ngroups * nranks : (unsigned long long)nranks;
        unsigned long long d;
        do { d = ld_sys(P.uc + OFF_GLOBAL + w) - prevG[par]; } while ((d >> CNT_SHIFT) != want);
        prevG[par] += d;
        total = d & DMASK;
      }
      if (i == iters && blockIdx.x == G - 1) sums[lane] = total;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && lane == 0) *cyc = (clock64() - t0) / iters;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2;
  const int iters = argc > 2 ? atoi(argv[2]) : 2000;
  CK(cuInit(0));
  std::vector<CUdevice> dev(n);
  bool mc_ok = n > 1;
  for (int d = 0; d < n; ++d) {
    CR(cudaSetDevice(d));
    CR(cudaFree(0));
    CK(cuDeviceGet(&dev[d], d));
    int mcs = 0;
    CK(cuDeviceGetAttribute(&mcs, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[d]));
    if (!mcs) mc_ok = false;
  }
  printf("%d GPUs, multicast %s\n", n, mc_ok ? "supported" : "NOT available (variants B, C, D skipped; E uses unicast red)");
  size_t bytes = TOTAL_WORDS * 8;
  CUmemAllocationProp ap = {};
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = 0;
  size_t gran = 0;
  CK(cuMemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  CUmulticastObjectProp mp = {};
  CUmemGenericAllocationHandle mch = 0;
  if (mc_ok) {
    mp.numDevices = n;
    mp.handleTypes = 0;
    mp.size = bytes;
    size_t mgran = 0;
    CK(cuMulticastGetGranularity(&mgran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    if (mgran > gran) gran = mgran;
  }
  bytes = (bytes + gran - 1) / gran * gran;
  if (mc_ok) {
    mp.size = bytes;
    CK(cuMulticastCreate(&mch, &mp));
    for (int d = 0; d < n; ++d) CK(cuMulticastAddDevice(mch, dev[d]));
  }
  std::vector<CUmemGenericAllocationHandle> mem(n);
  std::vector<CUdeviceptr> uc(n), mc(n, 0);
  std::vector<CUmemAccessDesc> all(n);
  for (int d = 0; d < n; ++d) {
    all[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    all[d].location.id = d;
    all[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  for (int d = 0; d < n; ++d) {
    CR(cudaSetDevice(d));
    ap.location.id = d;
    CK(cuMemCreate(&mem[d], bytes, &ap, 0));
    CK(cuMemAddressReserve(&uc[d], bytes, gran, 0, 0));
    CK(cuMemMap(uc[d], bytes, 0, mem[d], 0));
    CK(cuMemSetAccess(uc[d], bytes, all.data(), n));  // every GPU may store into every buffer
    if (mc_ok) {
      CK(cuMulticastBindMem(mch, 0, mem[d], 0, bytes, 0));
      CK(cuMemAddressReserve(&mc[d], bytes, gran, 0, 0));
      CK(cuMemMap(mc[d], bytes, 0, mch, 0));
      CK(cuMemSetAccess(mc[d], bytes, &all[d], 1));
    }
  }
  std::vector<long long *> cyc(n);
  std::vector<unsigned long long *> sums(n);
  for (int d = 0; d < n; ++d) { CR(cudaSetDevice(d)); CR(cudaMalloc(&cyc[d], 8)); CR(cudaMalloc(&sums[d], MAXW * 8)); }
  const int G = 148;
  struct Case { char v; int gs; };
  std::vector<Case> cases = {{'L', G}, {'A', G}, {'F', G}, {'E', 37}, {'E', 19}, {'E', 10}, {'E', 5}};
  if (mc_ok) { cases.push_back({'B', G}); cases.push_back({'C', G}); cases.push_back({'D', G}); }
  for (int words : {1, 40})
    for (const Case &c : cases) {
      for (int d = 0; d < n; ++d) { CR(cudaSetDevice(d)); CR(cudaMemset((void *)uc[d], 0, bytes)); CR(cudaDeviceSynchronize()); }
      for (int d = 0; d < n; ++d) {
        CR(cudaSetDevice(d));
        Ptrs P;
        P.uc = (unsigned long long *)uc[d];
        P.mc = (unsigned long long *)mc[d];
        for (int r = 0; r < MAXR; ++r) P.peer[r] = (unsigned long long *)uc[r < n ? r : 0];
        int variant = c.v, rank = d, nr = n, w = words, it = iters, gs = c.gs;
        void *args[] = {&P, &variant, &rank, &nr, &w, &it, &gs, &cyc[d], &sums[d]};
        CR(cudaLaunchCooperativeKernel((void *)rounds, dim3(G), dim3(64), args, 0, 0));
      }
      const unsigned long long per_gpu = (unsigned long long)G * (G + 1) / 2;
      const unsigned long long expect = c.v == 'L' ? 0 : per_gpu * n;
      printf("variant %c gs %3d words %2d:", c.v, c.gs, words);
      for (int d = 0; d < n; ++d) {
        long long h; unsigned long long s;
        CR(cudaSetDevice(d)); CR(cudaDeviceSynchronize());
        CR(cudaMemcpy(&h, cyc[d], 8, cudaMemcpyDeviceToHost)); CR(cudaMemcpy(&s, sums[d], 8, cudaMemcpyDeviceToHost));
        printf("  gpu%d %lld cyc/round%s", d, h, (c.v == 'L' || s == expect) ? "" : " SUM-MISMATCH");
      }
      printf("\n");
      fflush(stdout);
    }
  return 0;
}
