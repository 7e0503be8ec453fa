/* TEST INFRASTRUCTURE ONLY (oracle/).  The reference includes this header
 * (snpsamplinge.cc:4) but uses nothing from it. */
#ifndef TS_SHIM_GSL_HISTOGRAM_H
#define TS_SHIM_GSL_HISTOGRAM_H
#endif
