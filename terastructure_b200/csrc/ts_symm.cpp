// See ts_symm.hpp.  Host code only (compiled by nvcc for the include paths).
#include "ts_symm.hpp"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

namespace {

struct Drv {
  decltype(&cuDeviceGet) DeviceGet = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
  bool ok = false, mc_ok = false;
};

template <typename F>
bool load(const char *name, F *fn) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

const Drv &drv() {
  static Drv d = [] {
    Drv x;
    x.ok = load("cuDeviceGet", &x.DeviceGet) && load("cuDeviceGetAttribute", &x.DeviceGetAttribute) &&
           load("cuMemGetAllocationGranularity", &x.MemGetAllocationGranularity) && load("cuMemCreate", &x.MemCreate) &&
           load("cuMemRelease", &x.MemRelease) && load("cuMemAddressReserve", &x.MemAddressReserve) &&
           load("cuMemAddressFree", &x.MemAddressFree) && load("cuMemMap", &x.MemMap) && load("cuMemUnmap", &x.MemUnmap) &&
           load("cuMemSetAccess", &x.MemSetAccess) && load("cuGetErrorString", &x.GetErrorString);
    x.mc_ok = x.ok && load("cuMulticastCreate", &x.MulticastCreate) && load("cuMulticastAddDevice", &x.MulticastAddDevice) &&
              load("cuMulticastBindMem", &x.MulticastBindMem) && load("cuMulticastGetGranularity", &x.MulticastGetGranularity);
    return x;
  }();
  return d;
}

std::string cu_msg(const char *what, CUresult r) {
  const char *s = nullptr;
  if (drv().GetErrorString) drv().GetErrorString(r, &s);
  return std::string(what) + ": " + (s ? s : "unknown driver error");
}

}  // namespace

#define CU_TRY(call)                                   \
  do {                                                 \
    CUresult r_ = (call);                              \
    if (r_ != CUDA_SUCCESS) {                          \
      if (err) *err = cu_msg(#call, r_);               \
      symm_free(g);                                    \
      return nullptr;                                  \
    }                                                  \
  } while (0)

SymmGroup *symm_alloc_local(const int *devices, int n, size_t bytes, bool want_multicast, std::string *err) {
  const Drv &D = drv();
  if (!D.ok) {
    if (err) *err = "CUDA virtual-memory driver entry points are not available";
    return nullptr;
  }
  SymmGroup *g = new SymmGroup;
  g->n = n;
  g->devices.assign(devices, devices + n);
  std::vector<CUdevice> dev(n);
  bool mc = want_multicast && D.mc_ok && n > 1;
  for (int d = 0; d < n; ++d) {
    if (cudaSetDevice(devices[d]) != cudaSuccess || cudaFree(0) != cudaSuccess) {  // make sure the primary context exists
      if (err) *err = "cudaSetDevice failed";
      symm_free(g);
      return nullptr;
    }
    CU_TRY(D.DeviceGet(&dev[d], devices[d]));
    int sup = 0;
    if (mc && (D.DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[d]) != CUDA_SUCCESS || !sup)) mc = false;
  }
  CUmemAllocationProp ap = {};
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = dev[0];
  size_t gran = 0;
  CU_TRY(D.MemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  CUmulticastObjectProp mp = {};
  if (mc) {
    mp.numDevices = (unsigned)n;
    mp.handleTypes = 0;
    mp.size = bytes;
    size_t mgran = 0;
    if (D.MulticastGetGranularity(&mgran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) mc = false;
    else if (mgran > gran) gran = mgran;
  }
  g->bytes = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle mch = 0;
  if (mc) {
    mp.size = g->bytes;
    if (D.MulticastCreate(&mch, &mp) != CUDA_SUCCESS) mc = false;  // e.g. no fabric manager: fall back to unicast
  }
  if (mc) {
    g->mc_handle = mch;
    for (int d = 0; d < n; ++d) CU_TRY(D.MulticastAddDevice(mch, dev[d]));
  }
  std::vector<CUmemAccessDesc> all(n);
  for (int d = 0; d < n; ++d) {
    all[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    all[d].location.id = dev[d];
    all[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  for (int d = 0; d < n; ++d) {
    ap.location.id = dev[d];
    CUmemGenericAllocationHandle h = 0;
    CU_TRY(D.MemCreate(&h, g->bytes, &ap, 0));
    g->mem_handles.push_back(h);
    CUdeviceptr va = 0;
    CU_TRY(D.MemAddressReserve(&va, g->bytes, gran, 0, 0));
    g->uc.push_back((void *)va);
    CU_TRY(D.MemMap(va, g->bytes, 0, h, 0));
    CU_TRY(D.MemSetAccess(va, g->bytes, all.data(), (size_t)n));
    if (cudaSetDevice(devices[d]) != cudaSuccess || cudaMemset((void *)va, 0, g->bytes) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {
      if (err) *err = "clearing the symmetric buffer failed";
      symm_free(g);
      return nullptr;
    }
  }
  if (mc) {
    for (int d = 0; d < n; ++d) CU_TRY(D.MulticastBindMem(mch, 0, g->mem_handles[d], 0, g->bytes, 0));
    for (int d = 0; d < n; ++d) {
      CUdeviceptr va = 0;
      CU_TRY(D.MemAddressReserve(&va, g->bytes, gran, 0, 0));
      g->mc.push_back((void *)va);
      CU_TRY(D.MemMap(va, g->bytes, 0, mch, 0));
      CU_TRY(D.MemSetAccess(va, g->bytes, &all[d], 1));
    }
  }
  return g;
}

void symm_free(SymmGroup *g) {
  if (!g) return;
  const Drv &D = drv();
  if (D.ok) {
    for (void *p : g->mc) { D.MemUnmap((CUdeviceptr)p, g->bytes); D.MemAddressFree((CUdeviceptr)p, g->bytes); }
    for (void *p : g->uc) { D.MemUnmap((CUdeviceptr)p, g->bytes); D.MemAddressFree((CUdeviceptr)p, g->bytes); }
    for (unsigned long long h : g->mem_handles) D.MemRelease((CUmemGenericAllocationHandle)h);
    if (g->mc_handle) D.MemRelease((CUmemGenericAllocationHandle)g->mc_handle);
  }
  delete g;
}
