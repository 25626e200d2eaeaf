/* TEST INFRASTRUCTURE ONLY - minimal stand-in for <gsl/gsl_deriv.h>. */
#ifndef AIRICE_GSL_STANDIN_DERIV_H
#define AIRICE_GSL_STANDIN_DERIV_H
#include <gsl/gsl_math.h>
#ifdef __cplusplus
extern "C" {
#endif
int gsl_deriv_central(const gsl_function *f, double x, double h, double *result, double *abserr);
#ifdef __cplusplus
}
#endif
#endif
