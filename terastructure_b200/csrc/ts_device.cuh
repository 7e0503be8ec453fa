// Device-side structures shared by the kernels of libtsgpu.so (ts_engine.cu, ts_persist_inst.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ts_math.cuh"
#include "tsgpu.h"

// ------------------------------------------------------------------------------------------
// device-side structures
// ------------------------------------------------------------------------------------------
enum : uint32_t { ITEM_HOL = 1u, ITEM_FIRST = 2u };

// Exchange of the per-round statistics between the ranks:
//   SLOTS   CTA 0 stores the GPU's totals into its slot on every peer (unicast NVLink stores)
//   MCSLOT  the same with ONE multicast store (NVLS): the switch replicates it to every GPU
//   MCRED   no local stage: every CTA adds its words with multimem.red to every GPU's accumulators
//   GACC    CTA 0 adds the GPU's totals into an accumulator on every rank (one NVLink red.add per peer and
//           word), every CTA polls ONE local word pair whose count bits count ranks        (the default)
//   MCACC   the same with ONE multimem.red per word (the switch applies it to every rank's copy)
enum : int { XMODE_SLOTS = 0, XMODE_MCSLOT = 1, XMODE_MCRED = 2, XMODE_GACC = 3, XMODE_MCACC = 4 };

struct WorkItem {
  const unsigned char *col;  // packed column the E-step/gamma step read (bed row or vcol row)
  uint32_t loc;
  int32_t vslot;  // index among validation loci, or -1
  uint32_t flags;
  uint32_t pad;
};

constexpr int MAXK = TS_MAX_K;
constexpr int MAXR = 16;  // ranks in one exchange group

struct Ctl {
  long long cursor;  // index of the work item being processed
  uint32_t x;        // rounds completed for the current item
  uint32_t done;     // round loop finished
  uint32_t ticket;   // last-CTA-done counter of k_estep
  uint32_t epoch;    // exchange epoch (multi-GPU)
  uint32_t fault;    // set when a peer wait timed out
  uint32_t pad;
  double bcur[2 * MAXK];   // exp(Elogbeta[loc]) for the next E-step, [k*2+t]
  double bprev[2 * MAXK];  // the values the last executed E-step used (phi of the gamma step)
};

// Exchange buffer written by peers over NVLink: two epochs' worth of slots.
struct Xbuf {
  double val[2][MAXR][2 * MAXK];
  unsigned long long flag[2][MAXR];
};

// State of the persistent kernel's fence-free grid barrier (ts_persist.cuh).
constexpr int MG_SUB_MAX = 8;

struct PState {
  // [round parity][statistic][0 = high word, 1 = low word]: monotonic fixed-point accumulators, one 16-byte
  // pair per KB so that the 2K pairs of a round spread over the L2 slices (the upper half of the middle index
  // is unused)
  unsigned long long acc[2][4 * MAXK][128];
  unsigned long long prev[2][4 * MAXK];        // [set][hi: v, lo: 2K + v] word values at the end of the previous launch
  unsigned long long round_ctr;                // rounds run so far (the parity of the word set in use; same on every rank)
  uint32_t fault;
  uint32_t pad;
  // several ranks (XMODE_GACC / XMODE_MCACC): accumulators of the GPUs' totals, replicated on every rank (same
  // layout as acc; the top 6 bits count ranks, ts_persist.cuh GX_CNT_SHIFT), and their values at the end of the
  // previous launch
  unsigned long long gacc[2][4 * MAXK][128];
  unsigned long long gprev[2][4 * MAXK];
  // several ranks: the local words come in MG_SUB copies 128 bytes apart (acc[par][word][16 s]); the CTAs whose
  // index is s modulo MG_SUB arrive on copy s, and each copy's last arrival forwards the copy's total.  Values of the
  // copies at the end of the previous launch (copy 0 lives in prev):
  unsigned long long lprev[MG_SUB_MAX][2][4 * MAXK];
};

// Everything a peer GPU writes into lives in one allocation (one IPC handle).
struct Xchg {
  Xbuf x;
  PState ps;
};

struct Params {
  const unsigned char *bed;
  size_t pitch;
  double *gamma;
  double *E;
  uint32_t *cnt;
  size_t npad;
  uint32_t n_local;
  double *lambda;
  Ctl *ctl;
  const WorkItem *items;
  double *partial;  // [grid][2K]
  uint32_t *rounds; // per item
  // heldout
  const unsigned long long *voff;  // CSR over validation loci, local ids
  const uint32_t *vind;
  double *ll;  // per validation locus
  // hyper-parameters
  double alpha, eta0, eta1, nodetau0, nodekappa, thresh, lscale;
  uint32_t max_rounds;
  // exchange
  int rank, nranks;
  Xbuf *xlocal;
  Xbuf *xpeer[MAXR];
  // persistent kernel
  PState *pst;
  PState *pst_peer[MAXR];
  PState *pst_mc;                  // NVLS multicast alias of the symmetric PState (same offsets on all ranks), or null
  int xmode;                       // XMODE_*: how the ranks' totals of a round are exchanged
  unsigned long long mc_arrivals;  // XMODE_MCRED: CTAs per GPU x ranks = arrivals per word and round
  uint32_t tier_j;          // TIER kernels: individuals per thread kept in shared memory
  double fx_scale, fx_inv;  // 2^sh and 2^-sh of the fixed-point statistics
  int xflush;               // fence.sys after the peer stores (TSGPU_XFLUSH, default off)
  unsigned long long timeout_ns;  // wall-clock limit of one grid/peer wait (TSGPU_TIMEOUT_S, default 60 s)
  long long *trace;         // optional phase trace of CTA 0 (TSGPU_TRACE=1), 64 items x 128 slots
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// Persistent-kernel launcher, instantiated per K range in ts_persist_inst.cu (one translation unit
// per range so that the build parallelises).  Returns cudaErrorInvalidValue for a (K, I) pair that
// is not instantiated.
cudaError_t ts_launch_persist(int K, int I, bool tier, const Params &prm, uint32_t n_items, int grid, int block,
                              cudaStream_t stream);
int ts_persist_imax(int K);
int ts_persist_itier(int K);
int ts_persist_tmax(int K, int I);
int ts_persist_tier_threads(void);
size_t ts_persist_smem_base(int K, int threads);   // shared memory of a kernel compiled for `threads`, without the E tier
size_t ts_persist_tier_slot_bytes(int K);          // shared memory per individual-per-thread of the E tier
