// Table-driven f = exp(digamma(x)) and 1/f for the CONTROL path of the persistent kernel (ts_persist.cuh).
//
// estimate_beta (snpsamplinge.cc:279-296) turns a lambda row into b[k][t] = f(lambda_kt) / f(lambda_k0 + lambda_k1)
// once per round on ONE warp of every CTA, in the middle of the round's dependent chain (totals -> lambda -> b ->
// next E-step).  Latency is all that counts there: the table-free f_expsi of ts_expsi.cuh (built for the gamma
// step, where 100 000 x K evaluations per SVI iteration make instruction count and memory traffic count) is a chain
// of ~21 dependent FP64 instructions, followed by a reciprocal (5 more) -- ~650 cycles at B200's 23-cycle DFMA
// latency.  Here: one piecewise polynomial for f and one for 1/f (tools/gen_ftab.py -> ts_ftab.inc), argument
// reduction by integer operations on the exponent and mantissa bits, Estrin evaluation: 6 dependent FP64
// instructions from lambda to b.  The 20 KB of coefficients sit in the CTA's shared memory.
//
// Domain [1, 2^24) (lambda >= eta = 1 with the reference's prior; 2^24 = two alleles of 8M individuals); a warp
// with any argument outside falls back to f_expsi / fast_rcp (ts_expsi.cuh), bit-identically in every CTA and
// rank since all of them hold the same row.
// Host-compilable like ts_expsi.cuh (tests/expsi_check.cpp, tests/kernel_model.cpp); define TS_FTAB_HOST_TABLE
// before including to get the coefficient table in a host translation unit.
#pragma once
#include "ts_expsi.cuh"

namespace tsp {

constexpr int FTAB_OCTAVES = 24, FTAB_SUB = 4, FTAB_NI = FTAB_OCTAVES * FTAB_SUB;
constexpr int FTAB_NF = 12, FTAB_NG = 14, FTAB_STRIDE = FTAB_NF + FTAB_NG;  // doubles per interval
constexpr int FTAB_DOUBLES = FTAB_NI * FTAB_STRIDE;
constexpr size_t FTAB_BYTES = sizeof(double) * FTAB_DOUBLES;

#if defined(__CUDACC__)
static __device__ const double d_ftab[FTAB_DOUBLES] = {
#include "ts_ftab.inc"
};
#endif
#if defined(TS_FTAB_HOST_TABLE)
static const double h_ftab[FTAB_DOUBLES] = {
#include "ts_ftab.inc"
};
#endif

TSM_HD int hi_int(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t b;
  std::memcpy(&b, &x, 8);
  return (int)(uint32_t)(b >> 32);
#endif
}
TSM_HD double from_hilo(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x;
  std::memcpy(&x, &b, 8);
  return x;
#endif
}

// interval of x: four per octave, counted from x = 1; anything outside [1, 2^24) (NaN and negatives included)
// gives an index >= FTAB_NI
TSM_HD unsigned ftab_index(double x) { return (unsigned)((hi_int(x) >> 18) - (1023 << 2)); }

// t = (x - centre) / halfwidth in [-1, 1), exact: x with its exponent replaced by 3 lies in [8, 16), the
// interval [8 + 2j, 10 + 2j) has centre 9 + 2j = 8 (1 + j/4 + 1/8): mantissa bits j (two) and a one below
TSM_HD double ftab_t(double x) {
  const int h = hi_int(x);
  const double xs = from_hilo((h & 0x000fffff) | 0x40200000, lo_int(x));
  const double xc = from_hilo((h & 0x000c0000) | 0x40220000, 0);
  return xs - xc;
}

struct FtabPow {
  double t, t2, t4, t8;
};
TSM_HD FtabPow ftab_pow(double x) {
  FtabPow p;
  p.t = ftab_t(x);
  p.t2 = p.t * p.t;
  p.t4 = p.t2 * p.t2;
  p.t8 = p.t4 * p.t4;
  return p;
}

// sum_{i < N} c[i] t^i, Estrin: four dependent FMAs after t whatever N <= 16 (c is 16-byte aligned, N even)
template <int N>
TSM_HD double ftab_poly(const double *c, const FtabPow &p) {
  static_assert(N % 2 == 0 && N > 8 && N <= 16, "pairs of coefficients, three Estrin levels above them");
  double pr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (2 * i < N) {
#if defined(__CUDA_ARCH__)
      const double2 cc = *reinterpret_cast<const double2 *>(c + 2 * i);
      pr[i] = fma(cc.y, p.t, cc.x);
#else
      pr[i] = std::fma(c[2 * i + 1], p.t, c[2 * i]);
#endif
    } else {
      pr[i] = 0.0;
    }
  }
  constexpr int NP = N / 2;  // live pairs: 5..8
  const double q0 = fma(pr[1], p.t2, pr[0]);
  const double q1 = fma(pr[3], p.t2, pr[2]);
  const double q2 = NP > 5 ? fma(pr[5], p.t2, pr[4]) : pr[4];
  const double q3 = NP > 7 ? fma(pr[7], p.t2, pr[6]) : pr[6];
  const double h0 = fma(q1, p.t4, q0);
  const double h1 = NP > 6 ? fma(q3, p.t4, q2) : q2;
  return fma(h1, p.t8, h0);
}

#if defined(__CUDACC__)
// The same with the table addressed in the shared window (32-bit address, ld.shared): a generic pointer to shared
// memory costs an S2R SR_CgaCtaId + address arithmetic in front of every group of loads -- on the round's chain.
template <int N>
__device__ __forceinline__ double ftab_poly_sh(uint32_t sa, const FtabPow &p) {
  static_assert(N % 2 == 0 && N > 8 && N <= 16, "pairs of coefficients, three Estrin levels above them");
  double pr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (2 * i < N) {
      double c0, c1;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(sa + 16u * i));
      pr[i] = fma(c1, p.t, c0);
    } else {
      pr[i] = 0.0;
    }
  }
  constexpr int NP = N / 2;
  const double q0 = fma(pr[1], p.t2, pr[0]);
  const double q1 = fma(pr[3], p.t2, pr[2]);
  const double q2 = NP > 5 ? fma(pr[5], p.t2, pr[4]) : pr[4];
  const double q3 = NP > 7 ? fma(pr[7], p.t2, pr[6]) : pr[6];
  const double h0 = fma(q1, p.t4, q0);
  const double h1 = NP > 6 ? fma(q3, p.t4, q2) : q2;
  return fma(h1, p.t8, h0);
}
__device__ __forceinline__ double ftab_f_sh(uint32_t sa, unsigned idx, double x) {
  return ftab_poly_sh<FTAB_NF>(sa + idx * (unsigned)(FTAB_STRIDE * sizeof(double)), ftab_pow(x));
}
__device__ __forceinline__ double ftab_g_sh(uint32_t sa, unsigned idx, double x) {
  return ftab_poly_sh<FTAB_NG>(sa + (idx * FTAB_STRIDE + FTAB_NF) * (unsigned)sizeof(double), ftab_pow(x));
}
#endif

// f(x) and 1/f(x) for x in interval idx < FTAB_NI of the table at `tab` (shared memory in the kernel)
TSM_HD double ftab_f(const double *tab, unsigned idx, double x) { return ftab_poly<FTAB_NF>(tab + idx * FTAB_STRIDE, ftab_pow(x)); }
TSM_HD double ftab_g(const double *tab, unsigned idx, double x) { return ftab_poly<FTAB_NG>(tab + idx * FTAB_STRIDE + FTAB_NF, ftab_pow(x)); }

// b = f(own) / f(s) for one statistic; `all_in` = every lane's two arguments lie inside the table's domain
// (one vote per warp in the kernel, per group of 32 statistics in the host model)
TSM_HD bool ftab_covers(double own, double s) { return ftab_index(own) < (unsigned)FTAB_NI && ftab_index(s) < (unsigned)FTAB_NI; }
TSM_HD double beta_ratio(const double *tab, bool all_in, double own, double s) {
  if (all_in) return ftab_f(tab, ftab_index(own), own) * ftab_g(tab, ftab_index(s), s);
  return f_expsi(own) * fast_rcp(f_expsi(s));
}

}  // namespace tsp
