// Instantiations of the persistent SVI kernel for K in [TS_KLO, TS_KHI] (compile-time range, one
// translation unit per range: see Makefile).  Exactly one unit (TS_PERSIST_MAIN) also provides the
// dispatcher over ranges.
#include "ts_device.cuh"
#include "ts_persist.cuh"

#ifndef TS_KMIN  // developer builds (tools/dev/ab.sh) instantiate a sub-range of K only: smaller, faster to build and ship
#define TS_KMIN 1
#endif
#ifndef TS_KMAX
#define TS_KMAX 32
#endif
#define TS_CAT2(a, b) a##b
#define TS_CAT(a, b) TS_CAT2(a, b)
#define TS_RANGE_FN TS_CAT(ts_launch_persist_k, TS_KLO)

template <int K, int I, bool TIER, bool MG>
static cudaError_t launch_kim(const Params &prm, uint32_t n_items, int grid, int block, cudaStream_t stream) {
  if constexpr (I > tsp::persist_imax(K) || (TIER && I != tsp::persist_itier(K))) {
    return cudaErrorInvalidValue;
  } else {
    const size_t smem = tsp::persist_smem_bytes(K, TIER ? tsp::TIER_THREADS : tsp::persist_tmax(K, I)) +
                        (TIER ? (size_t)prm.tier_j * tsp::persist_tier_slot_bytes(K) : 0);
    // Set once per device: cudaFuncSetAttribute can serialise behind a running kernel, and a
    // peer's kernel may be spinning on the one this call is about to launch.
    static size_t attr_set[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && attr_set[dev & 63] < smem) {
      cudaError_t er = cudaFuncSetAttribute(tsp::k_persist<K, I, TIER, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (er != cudaSuccess) return er;
      attr_set[dev & 63] = smem;
    }
    Params p = prm;
    void *args[] = {(void *)&p, (void *)&n_items};
    return cudaLaunchCooperativeKernel((const void *)tsp::k_persist<K, I, TIER, MG>, dim3(grid), dim3(block), args, smem, stream);
  }
}

template <int K, int I, bool TIER>
static cudaError_t launch_ki(const Params &prm, uint32_t n_items, int grid, int block, cudaStream_t stream) {
  return prm.nranks > 1 ? launch_kim<K, I, TIER, true>(prm, n_items, grid, block, stream)
                        : launch_kim<K, I, TIER, false>(prm, n_items, grid, block, stream);
}

template <int K>
static cudaError_t launch_k(int I, bool tier, const Params &prm, uint32_t n_items, int grid, int block, cudaStream_t stream) {
  if (tier) return I == tsp::persist_itier(K) ? launch_ki<K, tsp::persist_itier(K), true>(prm, n_items, grid, block, stream) : cudaErrorInvalidValue;
  switch (I) {
    case 1: return launch_ki<K, 1, false>(prm, n_items, grid, block, stream);
    case 2: return launch_ki<K, 2, false>(prm, n_items, grid, block, stream);
    case 3: return launch_ki<K, 3, false>(prm, n_items, grid, block, stream);
    case 4: return launch_ki<K, 4, false>(prm, n_items, grid, block, stream);
  }
  return cudaErrorInvalidValue;
}

template <int K>
static cudaError_t launch_from(int k, int I, bool tier, const Params &prm, uint32_t n_items, int grid, int block, cudaStream_t stream) {
  if constexpr (K > TS_KHI) {
    return cudaErrorInvalidValue;
  } else {
    if (k == K) {
      if constexpr (K >= TS_KMIN && K <= TS_KMAX) return launch_k<K>(I, tier, prm, n_items, grid, block, stream);
      else return cudaErrorInvalidValue;  // a developer build restricted to some K (make KMIN=.. KMAX=..)
    }
    return launch_from<K + 1>(k, I, tier, prm, n_items, grid, block, stream);
  }
}

cudaError_t TS_RANGE_FN(int K, int I, bool tier, const Params &prm, uint32_t n_items, int grid, int block, cudaStream_t stream) {
  return launch_from<TS_KLO>(K, I, tier, prm, n_items, grid, block, stream);
}

#ifdef TS_PERSIST_MAIN
cudaError_t ts_launch_persist_k5(int, int, bool, const Params &, uint32_t, int, int, cudaStream_t);
cudaError_t ts_launch_persist_k9(int, int, bool, const Params &, uint32_t, int, int, cudaStream_t);
cudaError_t ts_launch_persist_k13(int, int, bool, const Params &, uint32_t, int, int, cudaStream_t);
cudaError_t ts_launch_persist_k17(int, int, bool, const Params &, uint32_t, int, int, cudaStream_t);
cudaError_t ts_launch_persist_k21(int, int, bool, const Params &, uint32_t, int, int, cudaStream_t);

cudaError_t ts_launch_persist(int K, int I, bool tier, const Params &prm, uint32_t n_items, int grid, int block, cudaStream_t stream) {
  if (K <= 4) return ts_launch_persist_k1(K, I, tier, prm, n_items, grid, block, stream);
  if (K <= 8) return ts_launch_persist_k5(K, I, tier, prm, n_items, grid, block, stream);
  if (K <= 12) return ts_launch_persist_k9(K, I, tier, prm, n_items, grid, block, stream);
  if (K <= 16) return ts_launch_persist_k13(K, I, tier, prm, n_items, grid, block, stream);
  if (K <= 20) return ts_launch_persist_k17(K, I, tier, prm, n_items, grid, block, stream);
  return ts_launch_persist_k21(K, I, tier, prm, n_items, grid, block, stream);
}
int ts_persist_imax(int K) { return tsp::persist_imax(K); }
int ts_persist_itier(int K) { return tsp::persist_itier(K); }
int ts_persist_tmax(int K, int I) { return tsp::persist_tmax(K, I); }
int ts_persist_tier_threads(void) { return tsp::TIER_THREADS; }
size_t ts_persist_smem_base(int K, int threads) { return tsp::persist_smem_bytes(K, threads); }
size_t ts_persist_tier_slot_bytes(int K) { return tsp::persist_tier_slot_bytes(K); }
#endif
