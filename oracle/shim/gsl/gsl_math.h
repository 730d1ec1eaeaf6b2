// TEST INFRASTRUCTURE (oracle build only). Empty stand-in for <gsl/gsl_math.h>; the reference only needs <cmath>.
#ifndef WF_SHIM_GSL_MATH_H
#define WF_SHIM_GSL_MATH_H
#include <cmath>
#include <cstddef>
#endif
