// Persistent SVI kernel: a whole batch of SVI iterations in ONE cooperative launch.
//
// Replaces, for n consecutive SNPs, the reference's optimize_lambda <-> PhiRunnerE::do_work
// hand-off (snpsamplinge.cc:320-366 <-> :649-686: two queue hops and a condvar per round) and
// the lazy update_gamma/estimate_theta (cc:695-740) with a grid-resident loop:
//
//   per SNP:  warp 0 of every CTA: b[k][t] = f(lambda[loc][k][t]) / f(lambda[k][0]+lambda[k][1]),
//                                  f = exp o digamma                         (estimate_beta)
//             -- f and 1/f from piecewise polynomials in shared memory (ts_ftab.cuh): this path is
//             one warp's dependent chain, six FP64 instructions from lambda to b
//     per round (<= online_iterations):
//       every thread: its individuals' E-step from REGISTERS (E = f(gamma) of up to 4 individuals
//             per thread stays in registers for the whole launch), 4K FMA + 2 reciprocals each
//       warp: transposed shuffle reduction of the 2K partial sums (each lane ends with one sum)
//       CTA:  every warp converts its 2K sums to 98-bit FIXED-POINT integers (two u64 words);
//             2K threads add the warps' words -- integer addition is associative, so the total
//             does not depend on the order of anything
//       grid: the CTA totals are added into 2K global 16-byte word pairs with relaxed red.add; each
//             word also counts arrivals in its top 10 bits, so the grid barrier needs no fence and
//             no separate flag: one L2 trip to publish, one round trip to observe.  The words are
//             monotonic (never reset); two sets alternate by round parity and every CTA remembers
//             the previous total of each set.  Lane v of the control warp waits 400 cycles (early
//             polls only queue in front of the arrivals), then spins on the pair of statistic v
//             with one 16-byte load.  Pairs sit 1 KB apart so that they spread over the L2 slices.
//       several GPUs (template parameter MG; the single-GPU kernel has none of this code): the
//             arrival is an atomic WITH return value on one of four copies of the local word; the
//             CTA whose arrival completes a copy forwards the copy's total into an accumulator that
//             is replicated on every rank (one multimem.red through NVLS, or one red.add per peer),
//             and every CTA of every GPU polls one local pair of that accumulator.
//       warp 0 of every CTA (redundantly, bit-identically): lambda, new b, convergence test -- the
//             two chains interleaved in one basic block
//     gamma step + E = f(gamma) refresh for the CTA's individuals (skipped in hol mode); when all
//     permitted rounds run it starts right after the last round's arrivals, in the shadow of that
//     round's barrier, and the last warp first prepares b for the next SNP's first round
//
// No host round trip, no kernel launch and no fence on the critical path of a round.
#pragma once
#include <type_traits>

#include "ts_expsi.cuh"
#include "ts_fixed.cuh"
#include "ts_ftab.cuh"

// Build-time switch for measurements (make lib SUFFIX=_nofence OPTS=8): bit 3 drops the fences of the
// row hand-off (see "row hand-off" below) to price them.  Variants measured and rejected in round 2:
// profiles/r2_summary.md sections 1 and 6.
#ifndef TS_OPTS
#define TS_OPTS 0
#endif
#define TS_OPT_NOFENCE ((TS_OPTS) & 8)
// One GPU: cycles between a CTA's arrival and its first poll.  Polls that come back incomplete are not free: the
// word's L2 slice serves the 148 CTAs' loads and arrivals one after the other.  Measured on two B200s, us per SVI
// iteration at 100 000 individuals: first poll at once 27.75 / 30.07, after 200 cycles - / 28.71, 300: 27.16 / 28.32,
// 400: 26.90 / 28.03, 500: 26.96, 600: 27.41; a pause between polls (fixed, or in proportion to the arrivals still
// missing) and a second poll in flight half a round trip later all lose; with the delay in place the pairs still beat
// low words a slice apart from the high words (26.90 against 27.04) (profiles/r2_summary.md).
#ifndef TS_POLL_DELAY
#define TS_POLL_DELAY 400
#endif

// Several GPUs: copies of the local words, each receiving the arrivals of the CTAs with index s modulo TS_MG_SUB.  The
// arrival is an atomic with return value, and a word's atomics are served one after the other by its L2 slice: with
// 148 CTAs on one word the last arrival's return value -- the moment the GPU's total can be forwarded -- is ~1 200
// cycles away, half of it queueing.
#ifndef TS_MG_SUB
#define TS_MG_SUB 4
#endif
#ifndef TS_CODE_AHEAD   // 1: the register tier's genotype bytes are loaded one SNP ahead
#define TS_CODE_AHEAD 1
#endif
#ifndef TS_TIER_UNROLL  // individuals of the shared-memory / streaming tiers in flight per thread (K <= 12)
#define TS_TIER_UNROLL 2
#endif
#ifndef TS_TIER_I       // individuals per thread in registers in the TIER kernels (K <= 12)
#define TS_TIER_I 2
#endif

// The two words of statistic v in an accumulator array a[2][4 MAXK][128] are adjacent (one 16-byte pair per KB, the
// pairs of a round spread over the L2 slices): the high word is a[par][v][0], the low word a[par][v][1]; one 16-byte
// load polls both.
#define TS_LO(a, par, v) (&(a)[par][v][1])
// Several GPUs: nobody polls the LOCAL words (arrivals are atomics with return value), and a word's 148 atomics are
// served one after the other by its L2 slice -- there the low words live a slice apart from the high words (the
// upper half of the middle index), which halves the queue in front of the last arrival's return value.
#define TS_LO_MG(a, par, v) (&(a)[par][V + (v)][0])

namespace tsp {

constexpr int FX_CNT_SHIFT = tsfx::CNT_SHIFT;
constexpr unsigned long long FX_MASK = tsfx::MASK;
// The ranks' accumulators (PState::gacc) count forwarded totals in their top 6 bits (ranks x copies of the local
// words, at most 63) and carry the totals unfolded: the low words of a GPU's CTAs add up to 148 x 2^44 and those of
// eight GPUs to 2^54.3 -- 58 data bits.
constexpr int GX_CNT_SHIFT = 58;
constexpr unsigned long long GX_MASK = (1ull << GX_CNT_SHIFT) - 1;

// one level of the transposed reduction: lanes whose `bit` is clear keep the low half of the
// N live values, the others the high half; each lane adds what its partner held of its half.
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

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// 16-byte (hi, lo) pair in one access; each word carries its own arrival count, so a torn pair is harmless
__device__ __forceinline__ void ld_pair(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void ld_pair_sys(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void red_add(unsigned long long *p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_add(unsigned long long *p, unsigned long long v) {  // returns the old value
  unsigned long long o;
  asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(v) : "memory");
  return o;
}
__device__ __forceinline__ void red_add_sys(unsigned long long *p, unsigned long long v) {  // also into a peer's memory
  asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// NVLS (NVSwitch multicast): the operation is applied to every GPU's copy of the symmetric buffer.
__device__ __forceinline__ void mm_red_add(unsigned long long *p, unsigned long long v) {
  asm volatile("multimem.red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// A lambda row travels between CTAs inside a launch (CTA 0 publishes the converged row; every CTA
// reads rows of loci it revisits): strong (L2-coherent) accesses, ordered by fence_gpu() on both
// sides of the grid barrier that separates the store from the loads (see "row hand-off" below).
__device__ __forceinline__ double ld_row(const double *p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_row(double *p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void fence_gpu() {
#if !TS_OPT_NOFENCE
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Poll loops give up after Params::timeout_ns of WALL CLOCK (a lost CTA or rank).  The timer is read
// once per POLL_BURST polls, OUTSIDE the burst loop: the burst itself must stay as tight as the
// round-1 loop (two loads, a compare, a branch; ptxas unrolls it) -- a guard inside it cost 500
// cycles per round (gpurun_out/r2c1_trace.txt).
constexpr int POLL_BURST = 4096;
struct SpinGuard {
  unsigned long long t0 = 0;
  __device__ __forceinline__ bool expired(unsigned long long limit_ns) {
    const unsigned long long now = global_ns();
    if (t0 == 0) { t0 = now; return false; }
    return now - t0 > limit_ns;
  }
};

// Phase trace for developers: CTA 0 / thread 0 stamps clock64() at phase boundaries.  Compiled in only with
// -DTS_TRACE_BUILD (make trace -> lib/libtsgpu_trace.so, used by tools/dev/trace_*.py with TSGPU_TRACE=1): in
// the product every stamp would put a constant load, a compare and a branch on the round's dependent chain
// (ncu showed the control warp waiting on exactly those between its FP64 instructions).
#ifdef TS_TRACE_BUILD
#define TS_TRACE(slot)                                                                        \
  do {                                                                                        \
    if (p.trace && blockIdx.x == 0 && tid == 0 && i < 64 && (unsigned)(slot) < 128u)          \
      p.trace[(size_t)i * 128 + (slot)] = clock64();                                          \
  } while (0)
#else
#define TS_TRACE(slot) do { } while (0)
#endif

// Row hand-off inside a launch.  The converged lambda row of a locus is needed again when the locus
// is revisited (round-0 b, the convergence test's old lambda).  Every CTA forms that row itself, so
// each keeps the rows of the last RING finished SNPs in shared memory and reads revisits within that
// window from there.  Older rows come from global memory, where CTA 0 publishes every row with a
// strong store and fences once every RING / 2 SNPs before a barrier arrival: a row that has left the
// ring was stored more than RING SNPs ago, so a fence and at least RING / 2 grid barriers lie
// between its store and any global read of it (release on the writer side; the readers fence before
// the strong loads).  Price of the fence on CTA 0's path: ~950 cycles, once per RING / 2 SNPs.
constexpr int RING = 8;

// Where a shard's E = exp(psi(gamma)) rows live during a launch (all tiers read them every round):
//   registers      I individuals per thread (k_persist<K, I, false>): shards up to 148 x T x I
//   shared memory  TIER kernels (k_persist<K, I, true>): J more individuals per thread in the CTA's
//                  shared memory (layout [j][k][thread]: conflict-free), J chosen at launch
//   global memory  TIER kernels: what is left streams from L2/HBM every round
// Individual n = m * (grid x T) + global thread id; m < I: registers, m < I + J: shared, else global.
// Threads per CTA are capped so that the register file holds the register tier without spills.
__host__ __device__ constexpr int persist_imax(int K) { return K <= 12 ? 4 : (K <= 20 ? 3 : 1); }
__host__ __device__ constexpr int persist_itier(int K) { return K <= 12 ? TS_TIER_I : 1; }  // register tier of the TIER kernels
constexpr int TIER_THREADS = 256;
constexpr int TIER_JMAX = 16;  // codes of the shared-memory tier travel as 2 bits each in one register
// 256 threads per CTA for every instantiation: the shard-sizing rule never gives a CTA more than eight
// warps (two per scheduler), and at 256 threads ptxas may use 255 registers -- the 384- and 512-thread
// caps of round 1 (168 / 128 registers) spilled 50-190 bytes in the I = 1 and I = 2 kernels.
__host__ __device__ constexpr int persist_tmax(int, int) { return 256; }
// shared memory without the E tier, for a kernel compiled for at most T threads per CTA (multiple of 16 bytes):
// the round's working set, then the control path's coefficient table (ts_ftab.cuh)
__host__ __device__ constexpr size_t persist_smem_work_bytes(int K, int T) {
  return (sizeof(double) * ((12 + 2 * RING) * K) + sizeof(long long) * (4 * K * (T / 32 + 1)) + 16 + sizeof(uint32_t) * RING + 15) / 16 * 16;
}
__host__ __device__ constexpr size_t persist_smem_bytes(int K, int T) { return persist_smem_work_bytes(K, T) + FTAB_BYTES; }
// bytes of shared-memory E tier per individual-per-thread slot
__host__ __device__ constexpr size_t persist_tier_slot_bytes(int K) { return sizeof(double) * K * TIER_THREADS; }

// s_t = sum_k E[k] b[k][t], t = 0, 1 (the softmax denominators of the E-step).  SPLIT = false: one
// dependent FMA chain per t (register tier: three individuals per thread give the FP64 pipe enough
// independent chains).  SPLIT = true: two half-length chains per t for the shared-memory and streaming
// tiers, whose individuals pass one or two at a time (dependent DFMA latency on B200: ~23 cycles).
// b = f(own) / f(s) without the table (an argument outside its domain): out of line, so that the rarely taken
// ~150 instructions do not sit between the table path and the end of the control block
static __device__ __noinline__ double beta_analytic(double own, double s) { return f_expsi(own) * fast_rcp(f_expsi(s)); }

template <int K, bool SPLIT>
__device__ __forceinline__ void dot_b(const double (&en)[K], const double *b, double &s0, double &s1) {
  if constexpr (!SPLIT || K < 4) {
    s0 = 0.0;
    s1 = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 bk = *reinterpret_cast<const double2 *>(b + 2 * k);
      s0 = fma(en[k], bk.x, s0);
      s1 = fma(en[k], bk.y, s1);
    }
  } else {
    double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 bk = *reinterpret_cast<const double2 *>(b + 2 * k);
      if (k & 1) { c0 = fma(en[k], bk.x, c0); c1 = fma(en[k], bk.y, c1); }
      else { a0 = fma(en[k], bk.x, a0); a1 = fma(en[k], bk.y, a1); }
    }
    s0 = a0 + c0;
    s1 = a1 + c1;
  }
}

// MG = false: the single-GPU kernel, with no trace of the exchange code (a third of the round loop's instructions
// and half of its branches; the control warp's path through a round is short enough to stay in the
// instruction cache).  MG = true: several ranks, exchange chosen at run time by Params::xmode.
template <int K, int I, bool TIER, bool MG>
__global__ void __launch_bounds__(TIER ? TIER_THREADS : persist_tmax(K, I), 1) k_persist(Params p, uint32_t n_items) {
  static_assert(I >= 1, "the register tier holds at least one individual per thread");
  constexpr int V = 2 * K;
  constexpr int NW = 2 * V;           // fixed-point words per round: word = hl * V + v
  constexpr int VPL = (V + 31) / 32;  // statistics per lane of the control warp
  constexpr int TM = TIER ? TIER_THREADS : persist_tmax(K, I);
  constexpr int TIER_UNROLL = K <= 12 ? TS_TIER_UNROLL : 1;  // tier individuals in flight per thread, while registers allow
  constexpr int WS = TM / 32 + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_b = reinterpret_cast<double *>(smem_raw);                // [2][V]: b of round x >= 1 at [x&1]
  double *s_b0 = s_b + 2 * V;                                        // [2][V]: b of round 0 of SNP i at [i&1]
  double *s_row0 = s_b0 + 2 * V;                                     // [2][V]: the lambda row that b came from
  double *s_ring = s_row0 + 2 * V;                                   // [RING][V]: rows of the loci finished last
  long long *s_fix = reinterpret_cast<long long *>(s_ring + RING * V);  // [NW][WS] per-warp fixed-point words
  int *s_flag = reinterpret_cast<int *>(s_fix + NW * WS);  // bit 0: round loop done, bit 1: abort
  uint32_t *s_ring_loc = reinterpret_cast<uint32_t *>(s_flag + 4);  // [RING]: locus of a ring slot, ~0 = empty
  double *s_tab = reinterpret_cast<double *>(smem_raw + persist_smem_work_bytes(K, TM));  // [FTAB_NI][FTAB_STRIDE]: f and 1/f
  double *s_E = reinterpret_cast<double *>(smem_raw + persist_smem_bytes(K, TM));  // TIER: [J][K][blockDim.x]

  PState *st = p.pst;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
  const uint32_t GT = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + tid;
  const unsigned long long G = gridDim.x;
  // Totals -> lambda (tsfx::to_double_plus): the four constants stay in registers for the whole launch.  Left to
  // itself the compiler rebuilds them from the kernel parameters in every round (three constant-bank loads and an
  // FP64 instruction in front of the round's first dependent DFMA); the empty asm makes them opaque.
  const tsfx::Unscale fxu = tsfx::unscale(p.fx_inv);
  double fx_hi_inv = fxu.hi_inv, fx_lo_inv = fxu.lo_inv, fx_lo_off = fxu.lo_off;
  double fx_hi_off_eta = fxu.hi_off + ((threadIdx.x & 1) ? p.eta1 : p.eta0);  // statistic v = lane + 32 q has the parity of the lane
  asm volatile("" : "+d"(fx_hi_inv), "+d"(fx_lo_inv), "+d"(fx_lo_off), "+d"(fx_hi_off_eta));
  // chg / V < thresh for certain below conv_lo, false for certain above conv_hi (see the convergence test)
  const double conv_lo = (p.thresh * (double)V) * (1.0 - 0x1p-48), conv_hi = (p.thresh * (double)V) * (1.0 + 0x1p-48);

  // which statistics this lane ends up holding after tr_reduce
  int tr_start, tr_len;
  tsfx::tr_slot<V>(lane, tr_start, tr_len);

  // control-warp state: previous totals of the two word sets it polls (the local words on one GPU, the ranks'
  // accumulators on several), current lambda row; on several GPUs also the local words' values before the round's
  // first arrival (lh/ll: what an arrival compares the atomic's return value with to learn that it is the last)
  unsigned long long ph0[VPL], pl0[VPL], ph1[VPL], pl1[VPL];
  unsigned long long lh0[VPL], ll0[VPL], lh1[VPL], ll1[VPL];
  double lam[VPL];
  unsigned long long rc = st->round_ctr;
  const int nranks = MG ? p.nranks : 1, xmode = MG ? p.xmode : XMODE_GACC;
  constexpr bool gacc_mode = MG;  // several ranks: the replicated accumulator (XMODE_GACC / XMODE_MCACC) is the only exchange
  // this CTA's copy of the local words, the CTAs that share it, the copies in use (all ranks run the same grid size)
  static_assert(TS_MG_SUB >= 1 && TS_MG_SUB <= MG_SUB_MAX, "copies of the local words");
  // Every rank adds NSUB_ALL to the count of a ranks' accumulator word per round, whatever its grid: a rank with fewer
  // CTAs than copies lets copy 0's forward stand for the copies it does not use.
  const unsigned NSUB_ALL = MG ? max(1u, min((unsigned)TS_MG_SUB, 63u / (unsigned)nranks)) : 1u;
  const unsigned nsub = min(NSUB_ALL, gridDim.x), sub = blockIdx.x % nsub;
  const unsigned long long G_sub = (gridDim.x - sub + nsub - 1) / nsub;
  const unsigned long long fwd_count = (unsigned long long)(sub == 0 ? NSUB_ALL - nsub + 1 : 1) << GX_CNT_SHIFT;
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int v = lane + 32 * q;
      const bool act = v < V;
      const unsigned long long(*pv)[4 * MAXK] = gacc_mode ? st->gprev : st->prev;
      ph0[q] = act ? pv[0][v] : 0;
      pl0[q] = act ? pv[0][V + v] : 0;
      ph1[q] = act ? pv[1][v] : 0;
      pl1[q] = act ? pv[1][V + v] : 0;
      lh0[q] = (gacc_mode && act) ? st->lprev[sub][0][v] : 0;
      ll0[q] = (gacc_mode && act) ? st->lprev[sub][0][V + v] : 0;
      lh1[q] = (gacc_mode && act) ? st->lprev[sub][1][v] : 0;
      ll1[q] = (gacc_mode && act) ? st->lprev[sub][1][V + v] : 0;
      lam[q] = 1024.0;  // idle lanes hold a large dummy so they never take f's small-argument path
    }
  }
  uint32_t prev_loc = 0xffffffffu;
  if (tid < RING) s_ring_loc[tid] = 0xffffffffu;
  // slots of warps this CTA does not have stay zero, so the CTA sum below adds all WS - 1 slots
  // unconditionally (straight-line vector loads instead of a predicated chain)
  for (int idx = tid; idx < NW * WS; idx += blockDim.x) s_fix[idx] = 0;
  for (int idx = tid; idx < FTAB_DOUBLES / 2; idx += blockDim.x)
    reinterpret_cast<double2 *>(s_tab)[idx] = reinterpret_cast<const double2 *>(d_ftab)[idx];
  // b = f(lambda_t) / f(lambda_0 + lambda_1) (estimate_beta, cc:279-296) for the statistic of each lane of a whole
  // warp: table-driven while every lane's arguments lie in the table's domain (ts_ftab.cuh)
  const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(s_tab);  // the table's address in the shared window
  // whole warp: lane's statistic(s) own[q] of a lambda row -> b; the table serves the row only if ALL its 2K
  // statistics (and their pair sums) lie in the table's domain, so that every path that turns a given row into b
  // (helper warp, control warp at a launch's first SNP, control warp between rounds) takes the same branch
  auto beta_row = [&](const double (&own)[VPL], double (&b)[VPL]) {
    double s2[VPL];
    bool in_tab = true;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int v = lane + 32 * q;
      const double other = __shfl_xor_sync(0xffffffffu, own[q], 1);
      const double l0 = (v & 1) ? other : own[q], l1 = (v & 1) ? own[q] : other;
      s2[q] = l0 + l1;  // the reference adds 0 + l0 + l1 (cc:283-286): 0 + l0 is exact
      in_tab = in_tab && ftab_covers(own[q], s2[q]);
    }
    in_tab = __all_sync(0xffffffffu, in_tab);
#pragma unroll
    for (int q = 0; q < VPL; ++q)
      b[q] = in_tab ? ftab_f_sh(tab_sa, ftab_index(own[q]), own[q]) * ftab_g_sh(tab_sa, ftab_index(s2[q]), s2[q])
                    : beta_analytic(own[q], s2[q]);
  };

  // this thread's individuals and their E = exp(psi(gamma)) rows: register tier
  constexpr int IR = I;
  uint32_t nj[IR];
  bool valid[IR];
  double e[IR][K];
#pragma unroll
  for (int j = 0; j < I; ++j) {
    nj[j] = gtid + (uint32_t)j * GT;
    valid[j] = nj[j] < p.n_local;
#pragma unroll
    for (int k = 0; k < K; ++k) e[j][k] = valid[j] ? p.E[(size_t)k * p.npad + nj[j]] : 0.0;
  }
  // shared-memory tier (J individuals per thread) and the first individual of the streaming tier
  const int J = TIER ? (int)p.tier_j : 0;
  const uint32_t T = blockDim.x;
  const uint32_t stream_begin = (uint32_t)(I + J) * GT;
  auto tier_n = [&](int j) { return (uint32_t)(I + j) * GT + gtid; };
  if constexpr (TIER) {
    for (int j = 0; j < J; ++j) {
      const uint32_t n = tier_n(j);
#pragma unroll
      for (int k = 0; k < K; ++k) s_E[(size_t)(j * K + k) * T + tid] = n < p.n_local ? p.E[(size_t)k * p.npad + n] : 0.0;
    }
  }
  __syncthreads();

  // Genotype bytes of the register tier's individuals travel one SNP ahead: the loads for SNP i + 1 are issued at
  // the start of SNP i and first touched at the start of SNP i + 1, so the dependent pair of loads (work item ->
  // column pointer -> byte, an L2 round trip each) is off the path between two SNPs.
#if TS_CODE_AHEAD
  unsigned raw[IR];
#pragma unroll
  for (int j = 0; j < I; ++j) raw[j] = (valid[j] && n_items > 0) ? p.items[0].col[nj[j] >> 2] : 0u;
#endif

  for (uint32_t i = 0; i < n_items; ++i) {
    const WorkItem it = p.items[i];
    const unsigned char *col = it.col;
    int code[IR];
#if TS_CODE_AHEAD
#pragma unroll
    for (int j = 0; j < I; ++j) code[j] = valid[j] ? (int)((raw[j] >> (2 * (nj[j] & 3))) & 3u) : 1;
    if (i + 1 < n_items) {
      const unsigned char *ncol = p.items[i + 1].col;
#pragma unroll
      for (int j = 0; j < I; ++j)
        if (valid[j]) raw[j] = ncol[nj[j] >> 2];
    }
#else
#pragma unroll
    for (int j = 0; j < I; ++j) code[j] = valid[j] ? tsm::plink_code(col, nj[j]) : 1;
#endif
    unsigned scode = 0;  // TIER: codes of the shared-memory tier, 2 bits per individual
    if constexpr (TIER) {
      for (int j = 0; j < J; ++j) {
        const uint32_t n = tier_n(j);
        scode |= (unsigned)(n < p.n_local ? tsm::plink_code(col, n) : 1) << (2 * j);
      }
    }
    if (i + 1 < n_items) {  // next SNP's genotype column and lambda row -> L2
      const WorkItem nx = p.items[i + 1];
      if constexpr (!TIER) {
#if !TS_CODE_AHEAD
#pragma unroll
        for (int j = 0; j < I; ++j)
          if (valid[j] && (nj[j] & 511) == 0) prefetch_l2(nx.col + (nj[j] >> 2));  // one per 128-byte line
#endif
      } else {
        for (uint32_t n = gtid * 512u; n < p.n_local; n += GT * 512u) prefetch_l2(nx.col + (n >> 2));
      }
      if (warp == 0 && lane < V) prefetch_l2(p.lambda + (size_t)nx.loc * V + lane);
    }
    // b of round 0 comes from the stored lambda row (estimate_beta, cc:279-296).  A helper warp
    // prepared it while the previous SNP's gamma step ran, unless this is the launch's first SNP or
    // the same locus again (its row was not final then): the control warp does it here in that case.
    //
    // With a single permitted round the helper would run BEFORE this SNP's only barrier and before
    // the previous SNP's row exists anywhere (locus sequence A, B, A): no preparation then.
    // whole warp: ring first, global memory otherwise (`fenced`: the caller ran the readers' fence already)
    auto load_row = [&](uint32_t loc, double (&own)[VPL], bool fenced) {
      const unsigned hit = __ballot_sync(0xffffffffu, lane < RING && s_ring_loc[lane] == loc);
      if (hit) {
        const double *row = s_ring + (__ffs(hit) - 1) * V;
#pragma unroll
        for (int q = 0; q < VPL; ++q) own[q] = (lane + 32 * q < V) ? row[lane + 32 * q] : 1024.0;
      } else {
        if (!fenced) fence_gpu();
#pragma unroll
        for (int q = 0; q < VPL; ++q) own[q] = (lane + 32 * q < V) ? ld_row(p.lambda + (size_t)loc * V + lane + 32 * q) : 1024.0;
      }
    };
    auto b_from_row = [&](uint32_t loc, double *dst, double *row_dst) {
      double own[VPL], b[VPL];
      load_row(loc, own, true);
      beta_row(own, b);
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int v = lane + 32 * q;
        if (v < V) { dst[v] = b[q]; row_dst[v] = own[q]; }
      }
    };
    const bool can_prepare = p.max_rounds >= 2;
    const bool prepared = can_prepare && i > 0 && it.loc != prev_loc;
    double *b_first = s_b0 + (i & 1) * V;
    if (warp == 0) {
      if (i > 0) {
        // the previous SNP's finished row (still in `lam`) enters this CTA's ring here, not where it
        // was formed: the helper warp may have been reading the ring until that SNP's last barrier
        const int slot = (int)((i - 1) % RING);
        if (lane < RING && (lane == slot || s_ring_loc[lane] == prev_loc)) s_ring_loc[lane] = lane == slot ? prev_loc : 0xffffffffu;
#pragma unroll
        for (int q = 0; q < VPL; ++q)
          if (lane + 32 * q < V) s_ring[slot * V + lane + 32 * q] = lam[q];
        __syncwarp();
      }
      if (prepared) {  // the helper left the row next to b
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int v = lane + 32 * q;
          lam[q] = (v < V) ? s_row0[(i & 1) * V + v] : 1024.0;
        }
      } else if (it.loc != prev_loc) {
        load_row(it.loc, lam, false);
      }
      if (!prepared) {
        double b[VPL];
        beta_row(lam, b);
#pragma unroll
        for (int q = 0; q < VPL; ++q)
          if (lane + 32 * q < V) b_first[lane + 32 * q] = b[q];
      }
      if (lane == 0) *s_flag = 0;
    }
    // the helper's job for the NEXT SNP; called once per SNP by warp W-1
    auto prepare_next = [&]() {
      if (can_prepare && warp == W - 1 && i + 1 < n_items) {
        const uint32_t nloc = p.items[i + 1].loc;
        if (nloc != it.loc) b_from_row(nloc, s_b0 + ((i + 1) & 1) * V, s_row0 + ((i + 1) & 1) * V);
        __syncwarp();
      }
    };
    prev_loc = it.loc;
    TS_TRACE(0);
    __syncthreads();
    TS_TRACE(1);

    uint32_t x = 0;
    double r0[IR], r1[IR];
    bool gamma_done = false, next_prepared = false;
    // ---- gamma natural-gradient step + E refresh (update_gamma/estimate_theta, cc:695-740) ----
    // phi of the LAST E-step: r0/r1 are still in registers, `bl` is the b that E-step used.
    auto gamma_one = [&](const double *bl, uint32_t n, int y, const double (&en)[K], double q0, double q1, double (&enew)[K]) {
      const uint32_t cn = p.cnt[n];
      double g[K];
#pragma unroll
      for (int k = 0; k < K; ++k) g[k] = p.gamma[(size_t)k * p.npad + n];
      const double base = p.nodetau0 + (double)cn;
      const double rho = (p.nodekappa == 0.5) ? rsqrt(base) : pow(base, -p.nodekappa);
      p.cnt[n] = cn + 1;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double2 bk = *reinterpret_cast<const double2 *>(bl + 2 * k);
        // y*phimom + (2-y)*phidad = E[k] * (b0[k]*y/s0 + b1[k]*(2-y)/s1)
        const double w = en[k] * fma(bk.x, q0, bk.y * q1);
        const double gn = g[k] + rho * (p.alpha + p.lscale * w - g[k]);
        p.gamma[(size_t)k * p.npad + n] = gn;
        enew[k] = f_expsi(gn);
      }
      (void)y;
    };
    // q of an individual outside the register tier, recomputed with the instruction sequence of the
    // E-step (same bits as the phi that entered the round's statistics)
    auto requantify = [&](const double *bl, int c, const double (&en)[K], double &q0, double &q1) {
      const int y = tsm::code_to_y(c);
      double s0, s1;
      dot_b<K, true>(en, bl, s0, s1);
      q0 = tsm::weight_of(y) * fast_rcp1(s0);
      q1 = tsm::weight_of(2 - y) * fast_rcp1(s1);
    };
    auto gamma_step = [&](const double *bl) {
#pragma unroll
      for (int j = 0; j < I; ++j) {
        if (code[j] == 1) continue;
        gamma_one(bl, nj[j], tsm::code_to_y(code[j]), e[j], r0[j], r1[j], e[j]);
      }
      if constexpr (TIER) {
#pragma unroll 1
        for (int j = 0; j < J; ++j) {  // shared-memory tier: E stays on chip, gamma streams through
          const int c = (scode >> (2 * j)) & 3;
          if (c == 1) continue;
          double en[K], q0, q1;
#pragma unroll
          for (int k = 0; k < K; ++k) en[k] = s_E[(size_t)(j * K + k) * T + tid];
          requantify(bl, c, en, q0, q1);
          gamma_one(bl, tier_n(j), tsm::code_to_y(c), en, q0, q1, en);
#pragma unroll
          for (int k = 0; k < K; ++k) s_E[(size_t)(j * K + k) * T + tid] = en[k];
        }
#pragma unroll 1
        for (uint32_t n = stream_begin + gtid; n < p.n_local; n += GT) {  // streaming tier
          const int c = tsm::plink_code(col, n);
          if (c == 1) continue;
          double en[K], q0, q1;
#pragma unroll
          for (int k = 0; k < K; ++k) en[k] = p.E[(size_t)k * p.npad + n];
          requantify(bl, c, en, q0, q1);
          gamma_one(bl, n, tsm::code_to_y(c), en, q0, q1, en);
#pragma unroll
          for (int k = 0; k < K; ++k) p.E[(size_t)k * p.npad + n] = en[k];
        }
      }
      if (i + 1 == n_items) {  // E leaves the chip only at the end of the launch
#pragma unroll
        for (int j = 0; j < I; ++j)
          if (valid[j]) {
#pragma unroll
            for (int k = 0; k < K; ++k) p.E[(size_t)k * p.npad + nj[j]] = e[j][k];
          }
        if constexpr (TIER) {
          for (int j = 0; j < J; ++j) {
            const uint32_t n = tier_n(j);
            if (n < p.n_local) {
#pragma unroll
              for (int k = 0; k < K; ++k) p.E[(size_t)k * p.npad + n] = s_E[(size_t)(j * K + k) * T + tid];
            }
          }
        }
      }
    };
    while (true) {
      const double *bx = x == 0 ? b_first : s_b + (x & 1) * V;
      const int par = (int)(rc & 1);
      // ---- E-step over this thread's individuals: registers + broadcast shared-memory b --------
      double vv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) vv[v] = 0.0;
      auto estep_one = [&](auto split, int c, bool ok, const double (&en)[K], double &q0, double &q1) {
        // missing or held out (kv_ok, hh:389-408) -> weight 0; branch-free, the warp stays converged
        // weights y and 2 - y built from the code's bits (no I2F.F64 in the round loop): codes 0, 2, 3 = y 0, 1, 2
        const double w0 = tsm::weight_of(c == 1 ? 0 : tsm::code_to_y(c)), w1 = tsm::weight_of(c == 1 ? 0 : 2 - tsm::code_to_y(c));
        double s0, s1;
        dot_b<K, decltype(split)::value>(en, bx, s0, s1);
        // padding threads (no individual) have e = 0, s = 0: keep the reciprocal finite
        q0 = w0 * fast_rcp1(ok ? s0 : 1.0);
        q1 = w1 * fast_rcp1(ok ? s1 : 1.0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          vv[2 * k] = fma(en[k], q0, vv[2 * k]);
          vv[2 * k + 1] = fma(en[k], q1, vv[2 * k + 1]);
        }
      };
#pragma unroll
      for (int j = 0; j < I; ++j) estep_one(std::false_type{}, code[j], valid[j], e[j], r0[j], r1[j]);
      if constexpr (TIER) {
#pragma unroll(TIER_UNROLL)
        for (int j = 0; j < J; ++j) {  // shared-memory tier, two individuals in flight while registers allow
          double en[K], q0, q1;
#pragma unroll
          for (int k = 0; k < K; ++k) en[k] = s_E[(size_t)(j * K + k) * T + tid];
          estep_one(std::true_type{}, (scode >> (2 * j)) & 3, tier_n(j) < p.n_local, en, q0, q1);
        }
#pragma unroll(TIER_UNROLL)
        for (uint32_t n = stream_begin + gtid; n < p.n_local; n += GT) {  // streaming tier: E from L2/HBM
          double en[K], q0, q1;
#pragma unroll
          for (int k = 0; k < K; ++k) en[k] = __ldcg(p.E + (size_t)k * p.npad + n);
          estep_one(std::true_type{}, tsm::plink_code(col, n), true, en, q0, q1);
        }
        __syncwarp();  // lane-dependent trip count: reconverge before the shuffles
      }
      TS_TRACE(2 + 8 * x + 0);
      // ---- warp: transposed reduction; every warp hands its sums over as fixed point -----------
      tr_reduce<V>(vv, lane);
#pragma unroll
      for (int q = 0; q < VPL; ++q)
        if (q < tr_len) {
          const int v = tr_start + q;
          // S_t[k] contribution of this warp = hi * 2^-sh + lo * 2^-(sh+44), hi = rint(sc) >= 0,
          // |lo| <= 2^43; both conversions are "add 2^52 and read the mantissa" (sc < 2^52)
          const double sc = (bx[v] * vv[q]) * p.fx_scale;
          long long whi, wlo;
          tsfx::split(sc, whi, wlo);
          s_fix[v * WS + warp] = whi;
          s_fix[(V + v) * WS + warp] = wlo;
        }
      TS_TRACE(2 + 8 * x + 1);
      __syncthreads();
      TS_TRACE(2 + 8 * x + 2);
      if constexpr (!MG) {
        for (int v = tid; v < V; v += blockDim.x) {  // add the warps' words, publish with arrival count 1
          const long long *sh = s_fix + v * WS, *sl = s_fix + (V + v) * WS;
          long long hi = 0, lo = 0;
#pragma unroll
          for (int ww = 0; ww < WS - 1; ++ww) { hi += sh[ww]; lo += sl[ww]; }
          tsfx::normalize(hi, lo);  // the low word becomes [0, 2^44)
          red_add(&st->acc[par][v][0], (unsigned long long)hi + (1ull << FX_CNT_SHIFT));
          red_add(TS_LO(st->acc, par, v), (unsigned long long)lo + (1ull << FX_CNT_SHIFT));
        }
      } else if (warp == 0) {
        // Several GPUs: the arrival is an atomic WITH return value.  The CTA whose addition completes a word (G - 1
        // arrivals before it) holds the GPU's total of that word and forwards it at once into every rank's
        // accumulator -- with one multimem.red where an NVLS alias exists, one NVLink red.add per peer otherwise.
        // No CTA waits for the local words: the detection of "all local CTAs arrived" by a polling CTA 0 (half to
        // one and a half L2 round trips) is replaced by the return trip of the last arrival's own atomic.
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int v = lane + 32 * q;
          if (v < V) {
            const long long *sh = s_fix + v * WS, *sl = s_fix + (V + v) * WS;
            long long hi = 0, lo = 0;
#pragma unroll
            for (int ww = 0; ww < WS - 1; ++ww) { hi += sh[ww]; lo += sl[ww]; }
            tsfx::normalize(hi, lo);
            const unsigned long long one = 1ull << FX_CNT_SHIFT, gone = fwd_count;
            const unsigned long long oh = atom_add(&st->acc[par][v][16 * sub], (unsigned long long)hi + one) - (par ? lh1[q] : lh0[q]);
            const unsigned long long ol = atom_add(TS_LO_MG(st->acc, par, v) + 16 * sub, (unsigned long long)lo + one) - (par ? ll1[q] : ll0[q]);
            const bool last_h = (oh >> FX_CNT_SHIFT) == G_sub - 1, last_l = (ol >> FX_CNT_SHIFT) == G_sub - 1;
            const unsigned long long th = gone + (oh & FX_MASK) + (unsigned long long)hi, tl = gone + (ol & FX_MASK) + (unsigned long long)lo;
            if (xmode == XMODE_MCACC) {
              if (last_h) mm_red_add(&p.pst_mc->gacc[par][v][0], th);
              if (last_l) mm_red_add(TS_LO(p.pst_mc->gacc, par, v), tl);
            } else if (last_h || last_l) {
              for (int r = 0; r < nranks; ++r) {
                if (last_h) red_add_sys(&p.pst_peer[r]->gacc[par][v][0], th);
                if (last_l) red_add_sys(TS_LO(p.pst_peer[r]->gacc, par, v), tl);
              }
            }
          }
        }
        __syncwarp();
      }
      TS_TRACE(2 + 8 * x + 3);
      // Row hand-off, both sides off the critical path.  Writer: the lanes of CTA 0 that stored the rows of
      // the SNPs before this one fence AFTER this arrival (the fence overlaps the barrier wait that follows
      // and precedes CTA 0's next arrival).  Reader: the helper warp fences in round 0's wait and prepares
      // the next SNP's round-0 b in round 1's wait, while the other warps sit in front of sync2.
      if (x == 0) {
        if (blockIdx.x == 0 && tid < V && (i % (RING / 2)) == 0 && i > 0) fence_gpu();
        if (warp == W - 1 && can_prepare && i + 1 < n_items) fence_gpu();
      } else if (x == 1) {
        prepare_next();
        next_prepared = true;
      }
      // The last round's totals only feed lambda[loc] (the gamma step uses the phi of THIS E-step),
      // so when this round is known to be the last one the gamma step runs now, in the shadow of
      // the grid barrier, and the control warp collects the totals afterwards.
      if (x + 1 >= p.max_rounds && !(it.flags & ITEM_HOL)) {
        gamma_step(bx);
        gamma_done = true;
      }
      // ---- control warp: grid barrier + totals, lambda update, convergence, new b -------------
      if (warp == 0) {
        __syncwarp();  // the gamma step above has lane-dependent control flow
        double tot[VPL];
        bool abort = false;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int v = lane + 32 * q;
          unsigned long long dh = 0, dl = 0;
          if (v < V) {
            SpinGuard guard;
            // single GPU: every CTA waits for the local words.  Several GPUs: only CTA 0 does (it forwards
            // the GPU's totals into every rank's accumulator); the other CTAs wait for the accumulator alone,
            // which keeps the pollers off the words the arrivals are being added to.
            if (gacc_mode) {
              // every CTA: one word pair per statistic, complete when every rank's last arrival has added its GPU's totals
              const unsigned long long bh = par ? ph1[q] : ph0[q], bl = par ? pl1[q] : pl0[q], want = (unsigned long long)nranks * NSUB_ALL;
              while (true) {
                bool complete = false;
                for (int t = 0; t < POLL_BURST; ++t) {
                  ld_pair_sys(&st->gacc[par][v][0], dh, dl);
                  dh -= bh;
                  dl -= bl;
                  if ((dh >> GX_CNT_SHIFT) == want && (dl >> GX_CNT_SHIFT) == want) { complete = true; break; }
                }
                if (complete) break;
                if (guard.expired(p.timeout_ns)) { abort = true; break; }
              }
              if (par) { ph1[q] = bh + dh; pl1[q] = bl + dl; } else { ph0[q] = bh + dh; pl0[q] = bl + dl; }
              dh &= GX_MASK;
              dl &= GX_MASK;
              tsfx::fold(dh, dl);  // the low words of all CTAs of all ranks: carry first (to_double wants both below 2^52)
              // The local words of this set are final now (every rank forwarded, so this GPU's last arrival happened):
              // their values are the base of the set's next use, two rounds from now.  The loads go straight into the
              // registers of that use; nothing touches them before, so the round does not wait for them.  (They are
              // issued a full round -- a grid barrier and an NVLink round trip -- before any CTA can arrive on the set again.)
              if (par) { lh1[q] = ld_relaxed(&st->acc[par][v][16 * sub]); ll1[q] = ld_relaxed(TS_LO_MG(st->acc, par, v) + 16 * sub); }
              else { lh0[q] = ld_relaxed(&st->acc[par][v][16 * sub]); ll0[q] = ld_relaxed(TS_LO_MG(st->acc, par, v) + 16 * sub); }
            } else {
              const unsigned long long bh = par ? ph1[q] : ph0[q], bl = par ? pl1[q] : pl0[q];
#if TS_POLL_DELAY > 0
              if (q == 0 && G > 32) { const long long t0 = clock64(); while (clock64() - t0 < TS_POLL_DELAY) {} }  // small grids complete at once
#endif
              while (true) {
                bool complete = false;
                for (int t = 0; t < POLL_BURST; ++t) {
                  ld_pair(&st->acc[par][v][0], dh, dl);
                  dh -= bh;
                  dl -= bl;
                  if ((dh >> FX_CNT_SHIFT) == G && (dl >> FX_CNT_SHIFT) == G) { complete = true; break; }
                }
                if (complete) break;
                if (guard.expired(p.timeout_ns)) { abort = true; break; }
              }
              if (par) { ph1[q] = bh + dh; pl1[q] = bl + dl; } else { ph0[q] = bh + dh; pl0[q] = bl + dl; }
              dh &= FX_MASK;
              dl &= FX_MASK;
              if (q == 0) TS_TRACE(82 + 2 * x);  // local words complete
            }
          }
          // update_lambda (cc:267-277): lambda = eta + S, u64 -> double through the mantissa (both words < 2^52)
          tot[q] = fma(tsfx::double_of((long long)(dh | 0x4330000000000000ull)), fx_hi_inv, fx_hi_off_eta) +
                   fma(tsfx::double_of((long long)(dl | 0x4330000000000000ull)), fx_lo_inv, fx_lo_off);
        }
        TS_TRACE(2 + 8 * x + 4);
        // New lambda -> new b AND the convergence sum in ONE basic block: both are chains of dependent FP64
        // operations (b: argument reduction, four Estrin levels, a product; the sum: five shuffle + DADD levels),
        // a warp issues in order, and only instructions of one block can be interleaved by the compiler --
        // written one after the other, behind the table-or-fallback branch, the two chains cost their sum
        // (~650 cycles) instead of the longer one.
        // Convergence (cc:359-365): mean |delta lambda| < thresh.  chg / V < thresh is decided without the IEEE
        // division (eight more dependent instructions) whenever chg is further than 2^-48 (relative) from
        // thresh * V -- the division's and the product's roundings move the boundary by at most 2^-52.
        double own[VPL], oldlam[VPL], s2[VPL], dlt[VPL];
        double *bn = s_b + ((x + 1) & 1) * V;
        const bool last = x + 1 >= p.max_rounds;
        bool in_tab = true;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int v = lane + 32 * q;
          oldlam[q] = lam[q];
          own[q] = (v < V) ? tot[q] : 1024.0;  // idle lanes: any large value
          lam[q] = own[q];
          const double other = __shfl_xor_sync(0xffffffffu, own[q], 1);
          const double l0 = (v & 1) ? other : own[q], l1 = (v & 1) ? own[q] : other;
          s2[q] = l0 + l1;  // the reference adds 0 + l0 + l1 (cc:283-286): 0 + l0 is exact
          dlt[q] = (v < V) ? fabs(own[q] - oldlam[q]) : 0.0;
          in_tab = in_tab && ftab_covers(own[q], s2[q]);
        }
        in_tab = __all_sync(0xffffffffu, in_tab);
        abort = __any_sync(0xffffffffu, abort);
        double chg = 0.0;
        if (in_tab) {
#pragma unroll
          for (int q = 0; q < VPL; ++q) {
            const double b = ftab_f_sh(tab_sa, ftab_index(own[q]), own[q]) * ftab_g_sh(tab_sa, ftab_index(s2[q]), s2[q]);  // estimate_beta (cc:279-296)
            if (lane + 32 * q < V && !last) bn[lane + 32 * q] = b;  // the last permitted round has no successor that would read b
            chg += dlt[q];
          }
          chg = warp_sum(chg);
        } else {
#pragma unroll
          for (int q = 0; q < VPL; ++q) {
            const double b = beta_analytic(own[q], s2[q]);
            if (lane + 32 * q < V && !last) bn[lane + 32 * q] = b;
            chg += dlt[q];
          }
          chg = warp_sum(chg);
        }
        TS_TRACE(2 + 8 * x + 7);
        // (the finished row and the round count go to global memory after the loop: nothing but the decision and
        // the flag stands between the totals and the barrier that releases the other warps)
        bool done = last || chg < conv_lo;
        if (!done && chg <= conv_hi) done = chg / (double)V < p.thresh;
        if (lane == 0) {
          if (abort) { st->fault = 1; *s_flag = 2; }
          else if (done) *s_flag = 1;
        }
        TS_TRACE(2 + 8 * x + 5);
      }
      __syncthreads();
      TS_TRACE(2 + 8 * x + 6);
      ++x;
      ++rc;
      const int flag = *s_flag;
      if (flag & 2) return;
      if (flag & 1) break;
    }

    // early-converged SNPs take the gamma step here; the usual case (all rounds run) took it
    // inside the last round, in the shadow of that round's grid barrier
    if (blockIdx.x == 0 && warp == 0) {  // the finished row (in `lam`; every CTA keeps it for its ring) to global memory
#pragma unroll
      for (int q = 0; q < VPL; ++q)
        if (lane + 32 * q < V) st_row(p.lambda + (size_t)it.loc * V + lane + 32 * q, lam[q]);
      if (lane == 0) p.rounds[i] = x;
    }
    if (!next_prepared) prepare_next();  // the SNP ended after its first round
    if (!gamma_done && !(it.flags & ITEM_HOL)) gamma_step(x == 1 ? b_first : s_b + ((x - 1) & 1) * V);
    TS_TRACE(100);
    __syncthreads();  // s_b is rewritten for the next SNP
    TS_TRACE(101);
  }

  if (gacc_mode && blockIdx.x < nsub && warp == 0) {  // the first CTA of every copy of the local words keeps the copy's values
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int v = lane + 32 * q;
      if (v < V) {
        st->lprev[sub][0][v] = lh0[q];
        st->lprev[sub][0][V + v] = ll0[q];
        st->lprev[sub][1][v] = lh1[q];
        st->lprev[sub][1][V + v] = ll1[q];
      }
    }
  }
  if (blockIdx.x == 0 && warp == 0) {
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int v = lane + 32 * q;
      if (v < V) {
        unsigned long long(*pv)[4 * MAXK] = gacc_mode ? st->gprev : st->prev;
        pv[0][v] = ph0[q];
        pv[0][V + v] = pl0[q];
        pv[1][v] = ph1[q];
        pv[1][V + v] = pl1[q];
      }
    }
    if (lane == 0) st->round_ctr = rc;
  }
}

}  // namespace tsp
