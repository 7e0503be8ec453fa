// Host check of terastructure_b200/csrc/ts_fixed.cuh (the fixed-point words the persistent kernel
// reduces with integer atomics).  Built and run by tests/test_host.py; prints "ok" or the failure.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "ts_fixed.cuh"

typedef __int128 i128;
#define REQUIRE(c)                                                       \
  do {                                                                   \
    if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; }  \
  } while (0)

// one round of the kernel's reduction: warps -> CTA words -> global words with arrival counts ->
// per-GPU totals -> sum over ranks -> double
static double reduce(const std::vector<double> &sc, int ranks, int ctas, int warps, int sh, bool fold_ranks,
                     unsigned long long word0, i128 *exact_out) {
  const tsfx::Unscale u = tsfx::unscale(std::ldexp(1.0, -sh));
  i128 exact = 0;
  unsigned long long th = 0, tl = 0;
  size_t i = 0;
  for (int r = 0; r < ranks; ++r) {
    unsigned long long acc_hi = word0, acc_lo = word0 * 3 + 12345;  // monotonic words, arbitrary history
    const unsigned long long prev_hi = acc_hi, prev_lo = acc_lo;
    for (int c = 0; c < ctas; ++c) {
      long long hi = 0, lo = 0;
      for (int w = 0; w < warps; ++w, ++i) {
        long long h, l;
        tsfx::split(sc[i], h, l);
        if (h < 0 || l > (1ll << 43) || l < -(1ll << 43)) return -1.0;
        exact += ((i128)h << 44) + l;
        hi += h;
        lo += l;
      }
      tsfx::normalize(hi, lo);
      if (lo < 0 || lo >= (1ll << 44) || hi < 0) return -2.0;
      acc_hi += (unsigned long long)hi + (1ull << tsfx::CNT_SHIFT);  // red.add: data + one arrival
      acc_lo += (unsigned long long)lo + (1ull << tsfx::CNT_SHIFT);
    }
    unsigned long long dh = acc_hi - prev_hi, dl = acc_lo - prev_lo;  // what the poller sees
    if ((dh >> tsfx::CNT_SHIFT) != (unsigned long long)ctas || (dl >> tsfx::CNT_SHIFT) != (unsigned long long)ctas) return -3.0;
    th += dh & tsfx::MASK;
    tl += dl & tsfx::MASK;
  }
  if (fold_ranks) tsfx::fold(th, tl);
  *exact_out = exact;
  return tsfx::to_double(th, tl, u);
}

int main() {
  std::mt19937_64 g(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);

  // shift_for: the largest possible statistic (2N) stays below 2^52 after scaling, with < 2x slack
  for (unsigned long long n : {1ull, 3ull, 200ull, 100000ull, 1000000ull, 1ull << 30, (1ull << 40) + 7}) {
    const int sh = tsfx::shift_for(n);
    REQUIRE(std::ldexp((double)(2 * n + 2), sh) < 4503599627370496.0);
    REQUIRE(std::ldexp((double)(2 * n + 2), sh + 1) >= 4503599627370496.0);
  }

  // split: hi + lo * 2^-44 reproduces sc to 2^-45, exactly when sc has no bits below 2^-44
  for (int t = 0; t < 200000; ++t) {
    const double sc = std::ldexp(U(g), (int)(U(g) * 70) - 18);  // 2^-18 .. 2^52
    long long h, l;
    tsfx::split(sc, h, l);
    REQUIRE(h >= 0 && l >= -(1ll << 43) && l <= (1ll << 43));
    const long double back = (long double)h + std::ldexp((long double)l, -44);
    REQUIRE(fabsl(back - (long double)sc) <= std::ldexp(1.0L, -45));
    if (sc >= 512.0) REQUIRE(back == (long double)sc);
  }
  { long long h, l; tsfx::split(0.0, h, l); REQUIRE(h == 0 && l == 0); }

  // whole reduction: exact integer total, one rounding at the end, independent of arrival order,
  // count bits intact across 64-bit wrap-around of the monotonic words
  for (int ranks : {1, 2, 4, 8}) {
    const int ctas = 148, warps = 8, sh = tsfx::shift_for(1000000ull * ranks / 8 + 1000);
    const size_t m = (size_t)ranks * ctas * warps;
    for (int rep = 0; rep < 20; ++rep) {
      std::vector<double> sc(m);
      const double top = std::ldexp(1.0, 52) / (double)m;  // totals up to ~2^51
      for (auto &x : sc) x = top * U(g) * (rep % 3 == 0 ? 1e-6 : 1.0);
      i128 exact, exact2;
      const unsigned long long word0 = rep % 2 ? 0xfffffffffff00000ull : 0x0123456789abcdefull;
      const double a = reduce(sc, ranks, ctas, warps, sh, true, word0, &exact);
      REQUIRE(a >= 0.0);
      // reference value: the exact 96-bit total, rounded once
      const long double ref = (std::ldexp((long double)(long long)(exact >> 44), 0) +
                               std::ldexp((long double)(long long)(exact & (((i128)1 << 44) - 1)), -44)) *
                              std::ldexp(1.0L, -sh);
      REQUIRE(fabsl((long double)a - ref) <= fabsl(ref) * 1.2e-16L);
      std::shuffle(sc.begin(), sc.end(), g);  // other warps/CTAs/GPUs hold the values: same bits
      const double b = reduce(sc, ranks, ctas, warps, sh, true, word0 ^ 0x5555, &exact2);
      REQUIRE(exact == exact2);
      REQUIRE(a == b);
    }
  }

  // regression: with 4+ GPUs the ranks' low words pass 2^52; without fold() the mantissa conversion
  // drops those bits (absolute error 2^(8-sh) per lost unit), with fold() the result is exact
  {
    const int ranks = 8, ctas = 148, warps = 8, sh = tsfx::shift_for(1000000);
    std::vector<double> sc((size_t)ranks * ctas * warps);
    for (auto &x : sc) x = 1000.0 + 0.999 * U(g) + 0.0004;  // low words near full scale
    i128 exact;
    const double good = reduce(sc, ranks, ctas, warps, sh, true, 7, &exact);
    const double bad = reduce(sc, ranks, ctas, warps, sh, false, 7, &exact);
    const long double ref = ((long double)(long long)(exact >> 44) + std::ldexp((long double)(long long)(exact & (((i128)1 << 44) - 1)), -44)) *
                            std::ldexp(1.0L, -sh);
    REQUIRE(fabsl((long double)good - ref) <= fabsl(ref) * 1.2e-16L);
    REQUIRE(fabsl((long double)bad - ref) > fabsl(ref) * 1e-13L);
  }
  // in-switch (NVLS) packing, XMODE_MCRED: 8 x 148 arrivals per word, 11 count bits, low word sent as lo >> 2
  {
    const int ranks = 8, ctas = 148, warps = 8, sh = tsfx::shift_for(1000000);
    const tsfx::Unscale u = tsfx::unscale(std::ldexp(1.0, -sh));
    unsigned long long acc_hi = 0xfff0000000000123ull, acc_lo = 0x8000000000000000ull;  // arbitrary history, wraps
    const unsigned long long prev_hi = acc_hi, prev_lo = acc_lo;
    i128 exact = 0;
    for (int c = 0; c < ranks * ctas; ++c) {
      long long hi = 0, lo = 0;
      for (int w = 0; w < warps; ++w) {
        long long h, l;
        tsfx::split(1000.0 + 0.999 * U(g) + 0.0004, h, l);
        exact += ((i128)h << 44) + l;
        hi += h;
        lo += l;
      }
      tsfx::normalize(hi, lo);
      acc_hi += (unsigned long long)hi + (1ull << tsfx::MC_CNT_SHIFT);
      acc_lo += ((unsigned long long)lo >> tsfx::MC_LO_DROP) + (1ull << tsfx::MC_CNT_SHIFT);
    }
    unsigned long long dh = acc_hi - prev_hi, dl = acc_lo - prev_lo;
    REQUIRE((dh >> tsfx::MC_CNT_SHIFT) == (unsigned long long)(ranks * ctas) && (dl >> tsfx::MC_CNT_SHIFT) == (unsigned long long)(ranks * ctas));
    dh &= tsfx::MC_MASK;
    dl &= tsfx::MC_MASK;
    tsfx::mc_unpack(dh, dl);
    REQUIRE(dh < (1ull << 52) && dl < (1ull << 44));
    const double got = tsfx::to_double(dh, dl, u);
    const long double ref = ((long double)(long long)(exact >> 44) + std::ldexp((long double)(long long)(exact & (((i128)1 << 44) - 1)), -44)) *
                            std::ldexp(1.0L, -sh);
    // at most 3 units of 2^-44 dropped per CTA
    REQUIRE(fabsl((long double)got - ref) <= std::ldexp((long double)(4 * ranks * ctas), -44 - sh) + fabsl(ref) * 1.2e-16L);
  }
  // lambda = eta + S with eta folded into the high word's offset: equals the exact value rounded once
  {
    const int sh = tsfx::shift_for(1000000);
    const tsfx::Unscale u = tsfx::unscale(std::ldexp(1.0, -sh));
    for (int it = 0; it < 100000; ++it) {
      const unsigned long long hi = (unsigned long long)(U(g) * 4.4e15) & ((1ull << 52) - 1), lo = (unsigned long long)(U(g) * 1.7e13) & ((1ull << 44) - 1);
      const double eta = (it & 1) ? 1.0 : 3.0;
      const long double exact = (long double)eta + std::ldexp((long double)hi, -sh) + std::ldexp((long double)lo, -sh - 44);
      const double got = tsfx::to_double_plus(hi, lo, u, eta);
      REQUIRE((long double)got == (long double)(double)exact || fabsl((long double)got - exact) <= fabsl(exact) * 1.2e-16L);
      REQUIRE(fabsl((long double)got - exact) <= fabsl((long double)(eta + tsfx::to_double(hi, lo, u)) - exact) + fabsl(exact) * 1e-19L);
    }
  }
  printf("ok\n");
  return 0;
}
