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
// (double)w for w in {0, 1, 2} through the exponent field: 1.0 = 0x3FF00000'00000000, 2.0 = 0x40000000'00000000
__device__ __forceinline__ double weight_of(int w) { return __hiloint2double(w ? 0x3FE00000 + (w << 20) : 0, 0); }

}  // namespace tsm
