/* TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <gsl/gsl_randist.h>. */
#ifndef TS_SHIM_GSL_RANDIST_H
#define TS_SHIM_GSL_RANDIST_H
#include <gsl/gsl_rng.h>
#ifdef __cplusplus
extern "C" {
#endif
double gsl_ran_gaussian_ziggurat(gsl_rng *r, double sigma);
double gsl_ran_gamma(gsl_rng *r, double a, double b);
#ifdef __cplusplus
}
#endif
#endif
