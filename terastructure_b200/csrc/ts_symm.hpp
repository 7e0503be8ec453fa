// Symmetric exchange buffers for several GPUs driven by ONE process (the CLI's -gpus N): one
// physical allocation per device, every device mapped into every other (NVLink peer access through
// the virtual-memory API) and, where the fabric supports it, an NVLS multicast object bound to all of
// them, so that a multimem.st / multimem.red issued by any GPU lands in every GPU's copy.
// The driver API is reached through cudaGetDriverEntryPoint: libtsgpu.so does not link libcuda, so it
// still loads (and its host-side entry points work) on a machine without a driver.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

struct SymmGroup {
  int n = 0;
  size_t bytes = 0;               // mapped size (rounded up to the allocation granularity)
  std::vector<int> devices;
  std::vector<void *> uc;         // uc[d]: device d's buffer, accessible from every device of the group
  std::vector<void *> mc;         // mc[d]: multicast alias as mapped for device d (empty: no NVLS)
  std::vector<unsigned long long> mem_handles;
  unsigned long long mc_handle = 0;
  bool has_multicast() const { return !mc.empty(); }
};

// Allocates zero-filled buffers of at least `bytes` on each device.  `want_multicast` = false skips the
// NVLS part.  Returns nullptr and fills `err` when the virtual-memory API itself is unusable; a missing
// multicast capability is not an error (the group then has no mc pointers).
SymmGroup *symm_alloc_local(const int *devices, int n, size_t bytes, bool want_multicast, std::string *err);
void symm_free(SymmGroup *g);
