// 98-bit fixed point for the 2K sufficient statistics of a round (persistent kernel, ts_persist.cuh).
//
// A statistic S (0 <= S <= 2N, the sum over individuals of phi-weighted allele counts,
// snpsamplinge.cc:660-680) is accumulated as two 64-bit integer words
//     S * 2^sh = hi + lo * 2^-44,      sh = 52 - bits(2 N_total + 2)
// so that the grid-wide (and multi-GPU) sum is an INTEGER sum: associative, hence independent of the
// order in which warps, CTAs and GPUs arrive -- bit-identical results run to run and rank to rank.
// The top 10 bits of a global word count arriving CTAs (grid barrier in the same word as the data).
//
// All conversions are "add 2^52 and read the mantissa": exact, two FP64 instructions each.
// Host-compilable (plain C++) so that tests/ can check the arithmetic without a GPU.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define TSFX_HD __host__ __device__ __forceinline__
#else
#define TSFX_HD inline
#endif

namespace tsfx {

constexpr int CNT_SHIFT = 54;                                // arrival count lives in bits 54..63
constexpr unsigned long long MASK = (1ull << CNT_SHIFT) - 1; // data bits of a word
constexpr int LO_BITS = 44;
constexpr double LO_SCALE = 17592186044416.0;                // 2^44
constexpr double TWO52 = 4503599627370496.0;                 // 2^52
constexpr double TWO52_51 = 6755399441055744.0;              // 2^52 + 2^51: signed mantissa trick

TSFX_HD long long bits_of(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  long long b;
  std::memcpy(&b, &x, 8);
  return b;
#endif
}
TSFX_HD double double_of(long long b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(b);
#else
  double x;
  std::memcpy(&x, &b, 8);
  return x;
#endif
}

// One warp's contribution sc = S_warp * 2^sh (0 <= sc < 2^52) -> hi = rint(sc) >= 0 and the signed
// remainder lo = rint((sc - hi) * 2^44), |lo| <= 2^43.
TSFX_HD void split(double sc, long long &hi, long long &lo) {
  const double th = sc + TWO52;
  const double rem = sc - (th - TWO52);
#if defined(__CUDA_ARCH__)
  const double tl = fma(rem, LO_SCALE, TWO52_51);
#else
  const double tl = rem * LO_SCALE + TWO52_51;  // rem * 2^44 is exact: same value as the fma
#endif
  hi = bits_of(th) - 0x4330000000000000ll;
  lo = bits_of(tl) - 0x4338000000000000ll;
}

// CTA level: after adding the warps' words, move the low word's carry so that 0 <= lo < 2^44
// (floor division: lo may be negative before).
TSFX_HD void normalize(long long &hi, long long &lo) {
  const long long carry = lo >> LO_BITS;
  hi += carry;
  lo -= carry << LO_BITS;
}

// Sum over ranks of per-GPU totals: every GPU's low word is below 148 * 2^44, but the sum over
// four or more ranks can pass 2^52, the limit of to_double below: move the carry up first (exact).
TSFX_HD void fold(unsigned long long &hi, unsigned long long &lo) {
  hi += lo >> LO_BITS;
  lo &= (1ull << LO_BITS) - 1;
}

// In-switch (NVLS) exchange, XMODE_MCRED: every CTA of every GPU arrives on every word, up to
// 8 x 148 = 1184 arrivals, so the count takes 11 bits (MC_CNT_SHIFT = 53) and the low word travels
// as lo >> 2 (< 2^42 per CTA, < 2^52.3 in total: fits the 53 data bits; the two dropped bits are
// 2^-42 of the high word's unit).  mc_unpack restores the (hi, lo) convention from the totals.
constexpr int MC_CNT_SHIFT = 53;
constexpr unsigned long long MC_MASK = (1ull << MC_CNT_SHIFT) - 1;
constexpr int MC_LO_DROP = 2;
TSFX_HD void mc_unpack(unsigned long long &hi, unsigned long long &lo) {
  hi += lo >> (LO_BITS - MC_LO_DROP);
  lo = (lo & ((1ull << (LO_BITS - MC_LO_DROP)) - 1)) << MC_LO_DROP;
}

// Totals (both words < 2^52) back to the statistic: (2^52 + d) * 2^-s - 2^(52-s) is exact, so each
// word costs one integer OR and one DFMA; the sum of the two rounds once.
struct Unscale {
  double hi_inv, hi_off, lo_inv, lo_off;
};
TSFX_HD Unscale unscale(double fx_inv /* 2^-sh */) {
  Unscale u;
  u.hi_inv = fx_inv;
  u.hi_off = -TWO52 * fx_inv;
  u.lo_inv = fx_inv * (1.0 / LO_SCALE);
  u.lo_off = -TWO52 * u.lo_inv;
  return u;
}
TSFX_HD double to_double(unsigned long long hi, unsigned long long lo, const Unscale &u) {
#if defined(__CUDA_ARCH__)
  const double dh = fma(double_of((long long)(hi | 0x4330000000000000ull)), u.hi_inv, u.hi_off);
  const double dl = fma(double_of((long long)(lo | 0x4330000000000000ull)), u.lo_inv, u.lo_off);
#else
  // products with a power of two are exact, so mul + add equals the fma
  const double dh = double_of((long long)(hi | 0x4330000000000000ull)) * u.hi_inv + u.hi_off;
  const double dl = double_of((long long)(lo | 0x4330000000000000ull)) * u.lo_inv + u.lo_off;
#endif
  return dh + dl;
}

// lambda = eta + S in one rounding: eta (a small integer-valued prior count, a multiple of 2^-sh) is
// folded into the high word's offset -- hi * 2^-sh + eta is exactly representable (< 2^53 units of 2^-sh),
// so the only rounding is the final addition of the low word's part.  One dependent FP64 instruction
// less on the round's critical path than eta + to_double(...), and one rounding less.
TSFX_HD double to_double_plus(unsigned long long hi, unsigned long long lo, const Unscale &u, double eta) {
#if defined(__CUDA_ARCH__)
  const double dh = fma(double_of((long long)(hi | 0x4330000000000000ull)), u.hi_inv, u.hi_off + eta);
  const double dl = fma(double_of((long long)(lo | 0x4330000000000000ull)), u.lo_inv, u.lo_off);
#else
  const double dh = double_of((long long)(hi | 0x4330000000000000ull)) * u.hi_inv + (u.hi_off + eta);
  const double dl = double_of((long long)(lo | 0x4330000000000000ull)) * u.lo_inv + u.lo_off;
#endif
  return dh + dl;
}

// 2^sh for a data set of n_total individuals: every statistic is at most 2 n_total.
TSFX_HD int shift_for(unsigned long long n_total) {
  int bits = 1;
  while ((2 * n_total + 2) >> bits) bits++;
  return 52 - bits;
}

// Which of the V per-thread partial sums a lane holds after the transposed warp reduction
// (tr_reduce in ts_persist.cuh): level by level (lane bits 16, 8, 4, 2, 1) a lane keeps the lower
// ceil(n/2) of the n live values if its bit is clear, the upper floor(n/2) otherwise.  The split uses
// the NOMINAL count ceil(n/2), the same for every lane; a lane that took a short upper half carries
// zero padding, so its live count `len` can be smaller than the nominal one (0 for V < 32 on some lanes).
template <int V>
TSFX_HD void tr_slot(int lane, int &start, int &len) {
  start = 0;
  len = V;
  int nominal = V;
#pragma unroll
  for (int bit = 16; bit > 0; bit >>= 1) {
    const int lo = (nominal + 1) / 2;
#if defined(__CUDA_ARCH__)
    if (lane & bit) { start += lo; len = max(len - lo, 0); } else len = min(len, lo);
#else
    if (lane & bit) { start += lo; len = len - lo > 0 ? len - lo : 0; } else len = len < lo ? len : lo;
#endif
    nominal = lo;
  }
}

}  // namespace tsfx
