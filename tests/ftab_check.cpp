// Host evaluation of terastructure_b200/csrc/ts_ftab.cuh (the control path's table-driven f = exp(digamma) and 1/f;
// the same source the persistent kernel inlines, same coefficient table ts_ftab.inc).  Prints, for a log grid of
// arguments from 0.5 to 3e7 with the interval boundaries among them: x, inside the table's domain?, f, 1/f, t
// as hex floats; tests/test_host.py compares them with mpmath.
#define TS_FTAB_HOST_TABLE
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "ts_ftab.cuh"

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 3000;
  const double lo = std::log(0.5), hi = std::log(3e7);
  for (int i = 0; i <= n; ++i) {
    double x = std::exp(lo + (hi - lo) * i / n);
    if (i % 7 == 3) x = std::ldexp(std::floor(std::ldexp(x, 2 - std::ilogb(x))), std::ilogb(x) - 2);  // an interval's first point
    if (i % 7 == 5) x = std::nextafter(std::ldexp(std::floor(std::ldexp(x, 2 - std::ilogb(x))), std::ilogb(x) - 2), 0.0);  // the last point of the one before
    const bool in = tsp::ftab_covers(x, x);
    const double f = in ? tsp::ftab_f(tsp::h_ftab, tsp::ftab_index(x), x) : -1.0;
    const double g = in ? tsp::ftab_g(tsp::h_ftab, tsp::ftab_index(x), x) : -1.0;
    printf("%a %d %a %a %a %a\n", x, (int)in, f, g, in ? tsp::ftab_t(x) : 0.0, tsp::beta_ratio(tsp::h_ftab, tsp::ftab_covers(x, 2.0 * x + 0.25), x, 2.0 * x + 0.25));
  }
  return 0;
}
