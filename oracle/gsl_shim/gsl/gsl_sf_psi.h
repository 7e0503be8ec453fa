/* TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <gsl/gsl_sf_psi.h>. */
#ifndef TS_SHIM_GSL_SF_PSI_H
#define TS_SHIM_GSL_SF_PSI_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_psi(double x);
#ifdef __cplusplus
}
#endif
#endif
