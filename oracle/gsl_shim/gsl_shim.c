/* TEST INFRASTRUCTURE ONLY -- part of oracle/, never linked into the product.
 *
 * Restatement of the GSL routines the reference's hot path calls
 * (call sites: snpsamplinge.cc:59-63,203,212,233,292-294,423,734-737;
 * lib.hh:29-32; snpsamplinge.hh:348-349).  GSL is absent from this image and
 * from /root/reference (third-party, unpinned), so its published algorithms
 * are restated:
 *   - gsl_rng_mt19937: Matsumoto & Nishimura MT19937, Knuth-style seeding
 *     (mt[i] = 1812433253*(mt[i-1]^(mt[i-1]>>30))+i), seed 0 -> 4357.
 *   - gsl_rng_uniform_int: scale = range/n; k = get()/scale; retry while k>=n.
 *   - gsl_ran_gamma: Marsaglia & Tsang (2000), using the ziggurat Gaussian.
 *   - gsl_ran_gaussian_ziggurat: Voss' 128-level ziggurat; tables regenerated
 *     by tools/gen_zig_tables.py.
 *   - gsl_sf_psi: digamma by upward recurrence + asymptotic series (GSL uses
 *     Chebyshev fits; both are accurate to ~1e-15, which is all the reference
 *     fixture can distinguish).
 * Pinned by: `make -C oracle kat` => md5(theta.txt) == md5(data/output_theta.txt).
 */
#include <math.h>
#include <stdlib.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_sf.h>
#include "zig_tables.h"

static const gsl_rng_type mt_type = {"mt19937", 0xffffffffUL, 0};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;
const gsl_rng_type *gsl_rng_default = &mt_type;
unsigned long gsl_rng_default_seed = 0;

const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }

void gsl_rng_set(gsl_rng *r, unsigned long s) {
  if (s == 0) s = 4357;
  r->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < 624; ++i)
    r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  r->mti = 624;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  r->type = T;
  gsl_rng_set(r, gsl_rng_default_seed);
  return r;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

unsigned long gsl_rng_get(gsl_rng *r) {
  unsigned long *mt = r->mt;
  if (r->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; ++kk) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
    }
    for (; kk < 623; ++kk) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
    }
    unsigned long y = (mt[623] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
    r->mti = 0;
  }
  unsigned long k = mt[r->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  return k & 0xffffffffUL;
}

double gsl_rng_uniform(gsl_rng *r) { return gsl_rng_get(r) / 4294967296.0; }

double gsl_rng_uniform_pos(gsl_rng *r) {
  double x;
  do x = gsl_rng_uniform(r); while (x == 0);
  return x;
}

unsigned long gsl_rng_uniform_int(gsl_rng *r, unsigned long n) {
  unsigned long range = r->type->max - r->type->min;
  unsigned long scale = range / n, k;
  do k = (gsl_rng_get(r) - r->type->min) / scale; while (k >= n);
  return k;
}

double gsl_ran_gaussian_ziggurat(gsl_rng *r, double sigma) {
  unsigned long i, j;
  int sign;
  double x, y;
  for (;;) {
    unsigned long k = gsl_rng_get(r);
    i = k & 0xFF;
    j = (k >> 8) & 0xFFFFFF;
    sign = (i & 0x80) ? +1 : -1;
    i &= 0x7f;
    x = j * zig_wtab[i];
    if (j < zig_ktab[i]) break;
    if (i < 127) {
      double y0 = zig_ytab[i], y1 = zig_ytab[i + 1];
      double U1 = gsl_rng_uniform(r);
      y = y1 + (y0 - y1) * U1;
    } else {
      double U1 = 1.0 - gsl_rng_uniform(r);
      double U2 = gsl_rng_uniform(r);
      x = ZIG_PARAM_R - log(U1) / ZIG_PARAM_R;
      y = exp(-ZIG_PARAM_R * (x - 0.5 * ZIG_PARAM_R)) * U2;
    }
    if (y < exp(-0.5 * x * x)) break;
  }
  return sign * sigma * x;
}

double gsl_ran_gamma(gsl_rng *r, double a, double b) {
  if (a < 1) {
    double u = gsl_rng_uniform_pos(r);
    return gsl_ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
  }
  double x, v, u;
  double d = a - 1.0 / 3.0;
  double c = (1.0 / 3.0) / sqrt(d);
  for (;;) {
    do {
      x = gsl_ran_gaussian_ziggurat(r, 1.0);
      v = 1.0 + c * x;
    } while (v <= 0);
    v = v * v * v;
    u = gsl_rng_uniform_pos(r);
    if (u < 1 - 0.0331 * x * x * x * x) break;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
  }
  return b * d * v;
}

double gsl_sf_psi(double x) {
  /* x > 0 on every reference call site. */
  double acc = 0.0;
  while (x < 12.0) {
    acc -= 1.0 / x;
    x += 1.0;
  }
  double r = 1.0 / x, r2 = r * r;
  /* sum_{n>=1} B_2n / (2n x^2n), through B_14 */
  double s = r2 * (1.0 / 12 - r2 * (1.0 / 120 - r2 * (1.0 / 252 - r2 * (1.0 / 240 - r2 * (1.0 / 132
             - r2 * (691.0 / 32760 - r2 * (1.0 / 12)))))));
  return acc + log(x) - 0.5 * r - s;
}

double gsl_sf_fact(unsigned int n) {
  double f = 1.0;
  for (unsigned int i = 2; i <= n; ++i) f *= i;
  return f;
}
