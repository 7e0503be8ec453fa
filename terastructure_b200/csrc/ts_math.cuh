// Device-side fp64 special functions for the TeraStructure hot path (sm_100a).
//
// The reference calls gsl_sf_psi (snpsamplinge.cc:292-294, :734-737; lib.hh:29-32) and libm
// exp/log (matrix.hh:271-293).  The contract is theta/beta within 1e-6 relative of the
// reference, so these are full fp64 implementations accurate to a few ulp.
#pragma once
#include <cuda_runtime.h>

namespace tsm {

// digamma(x), x > 0.  Upward recurrence psi(x) = psi(x+1) - 1/x until x >= 10, then the
// asymptotic series ln x - 1/(2x) - sum_n B_2n/(2n x^2n) through B_14 (truncation < 5e-17).
// The recurrence terms for up to ten steps are folded two at a time so that the common
// large-x case costs one log, one division and a degree-7 polynomial.
__device__ __forceinline__ double digamma(double x) {
  double acc = 0.0;
  while (x < 10.0) {
    // 1/x + 1/(x+1) = (2x+1) / (x(x+1))
    acc -= (2.0 * x + 1.0) / (x * (x + 1.0));
    x += 2.0;
  }
  const double r = 1.0 / x;
  const double r2 = r * r;
  double p = 1.0 / 12.0;  // coefficient of r2^7 : B_14/14 = (7/6)/14
  p = fma(-p, r2, 691.0 / 32760.0);
  p = fma(-p, r2, 1.0 / 132.0);
  p = fma(-p, r2, 1.0 / 240.0);
  p = fma(-p, r2, 1.0 / 252.0);
  p = fma(-p, r2, 1.0 / 120.0);
  p = fma(-p, r2, 1.0 / 12.0);
  // p = 1/12 - r2/120 + r2^2/252 - r2^3/240 + r2^4/132 - r2^5*691/32760 + r2^6/12
  return acc + (log(x) - 0.5 * r - p * r2);
}

// exp(digamma(x)): the only function of gamma[n][k] the E-step needs, because
// softmax_k(Elogtheta[n][k] + Elogbeta[k]) is invariant to the per-individual constant
// psi(sum_k gamma[n][k]) that Elogtheta subtracts (snpsamplinge.hh:276-300).
__device__ __forceinline__ double exp_digamma(double x) { return exp(digamma(x)); }

// PLINK 2-bit code -> genotype count; code 1 (binary 01) is "missing" (snp.cc:203-216).
__device__ __forceinline__ int plink_code(const unsigned char *col, unsigned long long n) {
  return (col[n >> 2] >> (2 * (unsigned)(n & 3))) & 3;
}
__device__ __forceinline__ int code_to_y(int code) { return code - (code >> 1); }

}  // namespace tsm

// ---- table-driven exp(digamma(x)) ---------------------------------------------------------------
// f(x) = exp(psi(x)) has no cheap closed form (essential singularity at 0, ~x - 1/2 at infinity);
// it is evaluated from a piecewise degree-12 polynomial table (tools/gen_expsi_table.py: 16
// sub-intervals per binade on [2^-5, 2^25), max relative error 2.5e-16, one 128-byte row per
// evaluation) at the cost of 12 FMAs instead of a log, a division, a 7-term series and an exp.
#define TS_EXPSI_DECL static __device__ __align__(128) const double g_expsi[TS_EXPSI_ROWS * TS_EXPSI_STRIDE]
#include "ts_expsi_table.inc"

namespace tsm {

__device__ __forceinline__ double exp_digamma_tab(double x) {
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) - 1023;  // x > 0 and finite on every call site
  if (e < TS_EXPSI_EMIN || e > TS_EXPSI_EMAX) return exp(digamma(x));  // out of table: exact slow path
  const int j = (hi >> 16) & 15;
  const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(x));  // x * 2^-e
  const double t = fma(m, 32.0, -(double)(33 + 2 * j));  // 32 * (m - (1 + (2j+1)/32)) in [-1, 1)
  const double2 *c = reinterpret_cast<const double2 *>(g_expsi + ((e - TS_EXPSI_EMIN) * TS_EXPSI_SUB + j) * TS_EXPSI_STRIDE);
  const double2 c01 = __ldg(c + 0), c23 = __ldg(c + 1), c45 = __ldg(c + 2), c67 = __ldg(c + 3),
                c89 = __ldg(c + 4), cab = __ldg(c + 5), ccd = __ldg(c + 6);
  double p = ccd.x;
  p = fma(p, t, cab.y);
  p = fma(p, t, cab.x);
  p = fma(p, t, c89.y);
  p = fma(p, t, c89.x);
  p = fma(p, t, c67.y);
  p = fma(p, t, c67.x);
  p = fma(p, t, c45.y);
  p = fma(p, t, c45.x);
  p = fma(p, t, c23.y);
  p = fma(p, t, c23.x);
  p = fma(p, t, c01.y);
  p = fma(p, t, c01.x);
  return p;
}

}  // namespace tsm
