/* TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <gsl/gsl_sf.h>. */
#ifndef TS_SHIM_GSL_SF_H
#define TS_SHIM_GSL_SF_H
#include <gsl/gsl_sf_psi.h>
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_fact(unsigned int n);
#ifdef __cplusplus
}
#endif
#endif
