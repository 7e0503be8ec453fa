/* TEST INFRASTRUCTURE ONLY -- part of oracle/, never linked into the product.
 *
 * Minimal stand-in for <gsl/gsl_rng.h> so that the reference sources under
 * /root/reference/src compile in an image without GSL.  GSL is a third-party
 * dependency of the reference (configure.ac:17-19, version unpinned, source not
 * vendored); the functions below restate GSL's published algorithms
 * (MT19937 with init_genrand seeding; gsl_rng_uniform_int's scale/reject rule).
 * Pinned end to end by the reference's own fixture data/output_theta.txt
 * (see oracle/Makefile target `kat`). */
#ifndef TS_SHIM_GSL_RNG_H
#define TS_SHIM_GSL_RNG_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsl_rng_type_s {
  const char *name;
  unsigned long max, min;
} gsl_rng_type;

typedef struct gsl_rng_s {
  const gsl_rng_type *type;
  unsigned long mt[624];
  int mti;
} gsl_rng;

extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_default;
extern unsigned long gsl_rng_default_seed;

const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
unsigned long gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
double gsl_rng_uniform_pos(gsl_rng *r);
unsigned long gsl_rng_uniform_int(gsl_rng *r, unsigned long n);

#ifdef __cplusplus
}
#endif
#endif
