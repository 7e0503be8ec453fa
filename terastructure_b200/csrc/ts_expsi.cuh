// exp(digamma(x)) and its building blocks for the persistent kernel (ts_persist.cuh), fp64.
//
// The reference evaluates gsl_sf_psi and libm exp/log per individual and round
// (snpsamplinge.cc:292-294, :734-737; matrix.hh:271-293); the kernel needs only f = exp o psi
// (see DESIGN.md section 3).  Host-compilable: outside device code the MUFU.RCP64H seed is modelled
// by a reciprocal truncated to 20 mantissa bits, so tests/ can check accuracy without a GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define TSM_HD __host__ __device__ __forceinline__
#else
#define TSM_HD inline
#endif

namespace tsp {

// ~20-bit reciprocal seed: rcp.approx.ftz.f64 (MUFU.RCP64H) on the device
TSM_HD double rcp_seed(double s) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  return r;
#else
  double r = 1.0 / s;
  uint64_t b;
  std::memcpy(&b, &r, 8);
  b &= 0xffffffff00000000ull;  // the hardware result has zero low mantissa bits
  std::memcpy(&r, &b, 8);
  return r;
#endif
}
TSM_HD int lo_int(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t b;
  std::memcpy(&b, &x, 8);
  return (int)(uint32_t)b;
#endif
}
TSM_HD double add_exponent(double p, int k) {  // p * 2^k through the exponent field (no overflow checks)
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
  uint64_t b;
  std::memcpy(&b, &p, 8);
  b += (uint64_t)(int64_t)k << 52;
  std::memcpy(&p, &b, 8);
  return p;
#endif
}

// 1/s for s > 0, normal: MUFU.RCP64H seed (~20 bits) + two Newton steps; within ~1 ulp, which is
// all the 1e-6 contract (and the 1e-9 test tolerance) can see.  The IEEE division it replaces
// costs ~4x as many FP64-pipe slots per individual and round.
TSM_HD double fast_rcp(double s) {
  double r = rcp_seed(s);
  double e = fma(-s, r, 1.0);
  r = fma(r, e, r);
  e = fma(-s, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// 1/s with ONE Newton step (relative error < 1e-11): enough where the result only weights a sum
// (E-step: the trajectory of the host model moves 3e-13 away from the oracle instead of 5e-14).
TSM_HD double fast_rcp1(double s) {
  const double r = rcp_seed(s);
  return fma(r, fma(-s, r, 1.0), r);
}

// f(x) = exp(digamma(x)) without any table: every coefficient is an immediate, so a warp whose
// lanes hold unrelated arguments still executes one instruction stream with no memory traffic
// (a polynomial-table variant needed 104 bytes of coefficients PER LANE and was LSU-bound in
// shared as well as in global memory).
//   x >= 8 : asymptotic series f(x) = x - 1/2 + sum_{j>=2} g_j x^(1-j), through j = 16 (error < 3e-16)
//   x <  8 : f(x) = f(x+8) * exp(-(1/x + ... + 1/(x+7))); the harmonic sum is P'(x)/P(x) with
//            P(x) = x(x+1)...(x+7) (all coefficients positive: no cancellation)
// Max relative error against mpmath (tests/expsi_check.cpp + tests/test_host.py, host model of the
// reciprocal seed): 2.2e-16 for x >= 8, 7.4e-15 on [0.03, 8) (the exp's truncation), 1.1e-13 down to
// x = 1.5e-3 (the exp's argument reaches 700 there and carries its own rounding).
// exp(-r) for r >= 0 (the only case needed): k = rint(-r*log2(e)), Cody-Waite reduction, degree-11
// Taylor polynomial on |t| <= ln2/2 (truncation t^12/12! <= 6.3e-15, the largest error term of f),
// 2^k applied through the exponent field.
// Arguments beyond 700 flush to zero (exp(psi(x)) for x < 1.4e-3 is below 1e-300).
TSM_HD double exp_neg(double r) {
  if (r > 700.0) return 0.0;
  const double kd = fma(-r, 1.4426950408889634, 6755399441055744.0);  // 2^52+2^51: rint in the mantissa
  const int k = lo_int(kd);
  const double kf = kd - 6755399441055744.0;
  double t = fma(kf, -0.693147180369123816490, -r);   // ln2 split: high part has 11 trailing zero bits
  t = fma(kf, -1.90821492927058770002e-10, t);
  const double t2 = t * t;
  // Estrin: p = sum_{j=0}^{11} t^j / j!
  const double p01 = 1.0 + t, p23 = fma(t, 1.0 / 6, 0.5), p45 = fma(t, 1.0 / 120, 1.0 / 24),
               p67 = fma(t, 1.0 / 5040, 1.0 / 720), p89 = fma(t, 1.0 / 362880, 1.0 / 40320),
               pab = fma(t, 1.0 / 39916800, 1.0 / 3628800);
  const double t4 = t2 * t2;
  const double q0 = fma(p23, t2, p01), q1 = fma(p67, t2, p45), q2 = fma(pab, t2, p89);
  const double p = fma(fma(q2, t4, q1), t4, q0);
  return add_exponent(p, k);
}

TSM_HD double f_expsi(double x) {
  const bool small = x < 8.0;
  const double xs = small ? x + 8.0 : x;
  // u = 1/xs enters f only through the correction u*q(u) <= f/1500, so ONE Newton step on the
  // 20-bit seed (relative error < 1e-11) moves f by less than 1e-14 relative
  double u = rcp_seed(xs);
  u = fma(u, fma(-xs, u, 1.0), u);
  // q(u) = g2 + g3 u + ... + g16 u^14, Estrin (dependent DFMA latency on B200 is ~23 cycles)
  const double u2 = u * u;
  const double a0 = fma(0x1.5555555555555p-6, u, 0x1.5555555555555p-5);
  const double a1 = fma(-0x1.2222222222222p-8, u, 0x1.05b05b05b05b0p-8);
  const double a2 = fma(0x1.1a4cc13ddafa2p-9, u, -0x1.c7f80db9bf2a3p-9);
  const double a3 = fma(-0x1.1e6ee98a17aecp-9, u, 0x1.05f536517fa45p-8);
  const double a4 = fma(0x1.fe414efb9852ap-9, u, -0x1.e5f884ccda9f9p-8);
  const double a5 = fma(-0x1.5f836e8779d89p-7, u, 0x1.54c7f9f55e0ebp-6);
  const double a6 = fma(0x1.59488e35cad4dp-5, u, -0x1.51ea52a4cdfabp-4);
  const double a7 = 0x1.c276c25d1fbddp-2;
  const double u4 = u2 * u2;
  const double b0 = fma(a1, u2, a0), b1 = fma(a3, u2, a2), b2 = fma(a5, u2, a4), b3 = fma(a7, u2, a6);
  const double u8 = u4 * u4;
  const double c0 = fma(b1, u4, b0), c1 = fma(b3, u4, b2);
  const double q = fma(c1, u8, c0);
  double f = fma(u, q, xs - 0.5);
  if (small) {
    // P = x(x+1)...(x+7), D = P'; even/odd split halves the dependency chains
    const double x2 = x * x;
    const double pe = fma(fma(fma(x2 + 322.0, x2, 6769.0), x2, 13068.0), x2, 0.0);          // x^8+322x^6+6769x^4+13068x^2
    const double po = fma(fma(fma(28.0, x2, 1960.0), x2, 13132.0), x2, 5040.0);              // 28x^6+1960x^4+13132x^2+5040 (times x)
    const double de = fma(fma(fma(196.0, x2, 9800.0), x2, 39396.0), x2, 5040.0);             // 196x^6+9800x^4+39396x^2+5040
    const double dod = fma(fma(fma(8.0, x2, 1932.0), x2, 27076.0), x2, 26136.0);             // 8x^6+1932x^4+27076x^2+26136 (times x)
    const double P = fma(po, x, pe), D = fma(dod, x, de);
    f *= exp_neg(D * fast_rcp(P));
  }
  return f;
}

}  // namespace tsp
