// Host evaluation of terastructure_b200/csrc/ts_expsi.cuh (the same source the persistent kernel
// inlines; the MUFU reciprocal seed is modelled by a 20-bit reciprocal).  Prints, for a log grid of
// arguments, x, f_expsi(x), fast_rcp(x), exp_neg(x) as hex floats; tests/test_host.py compares them
// with mpmath.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "ts_expsi.cuh"

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2000;
  const double lo = std::log(1e-3), hi = std::log(1e9);
  for (int i = 0; i <= n; ++i) {
    double x = std::exp(lo + (hi - lo) * i / n);
    if (i % 7 == 3) x = std::floor(x * 4.0 + 1.0) / 4.0;  // some exactly representable points, 8.0 among them
    printf("%a %a %a %a\n", x, tsp::f_expsi(x), tsp::fast_rcp(x), tsp::exp_neg(x < 750.0 ? x : 750.0 + 1e-7 * x));
  }
  return 0;
}
