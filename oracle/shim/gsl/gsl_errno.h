/* TEST INFRASTRUCTURE — stand-in for <gsl/gsl_errno.h>; the reference only includes it
 * (SLICER/utilities.h:14) and uses no symbol from it. */
#ifndef SLICER_SHIM_GSL_ERRNO_H
#define SLICER_SHIM_GSL_ERRNO_H
#define GSL_SUCCESS 0
#define GSL_EDOM 1
#endif
