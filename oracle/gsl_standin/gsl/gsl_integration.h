/* TEST INFRASTRUCTURE ONLY - stand-in for <gsl/gsl_integration.h>.  Only the in-ice
 * attenuation code (IceRayTracing.cc:179-200, out of scope) uses it; qags here is an
 * adaptive Gauss-Kronrod(21) bisection scheme without the epsilon extrapolation. */
#ifndef AIRICE_GSL_STANDIN_INTEGRATION_H
#define AIRICE_GSL_STANDIN_INTEGRATION_H
#include <stdlib.h>
#include <gsl/gsl_math.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t limit; } gsl_integration_workspace;
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qags(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                         size_t limit, gsl_integration_workspace *workspace, double *result,
                         double *abserr);
#ifdef __cplusplus
}
#endif
#endif
