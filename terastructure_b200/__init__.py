"""terastructure_b200 -- B200-native implementation of TeraStructure's SVI hot path.

Only what the path needs lives here:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/tsgpu.h) + the CLI
  capi.py          ctypes binding of the C ABI (plumbing)
  snpsamplinge.py  host-side mirror of the reference's SNPSamplingE driver (same names,
                   same RNG stream, same report/stop rules), calling the C ABI
  plink.py         .bed/.bim/.fam reader that keeps genotypes 2-bit packed
  synth.py         PSD/Balding-Nichols synthetic genotype generator
"""
from .capi import Engine, Rng, TsError, lib, plan_shard, plan_tiers, LIB_PATH  # noqa: F401
from .snpsamplinge import Env, SNPSamplingE  # noqa: F401
