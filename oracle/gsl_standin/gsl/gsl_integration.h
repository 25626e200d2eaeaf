/* TEST INFRASTRUCTURE ONLY - stand-in for <gsl/gsl_integration.h>: gsl_integration_qags as the in-ice attenuation
 * integrals call it (IceRayTracing.cc:179-200), restated from GSL 2.x / QUADPACK in gsl_standin.c. */
#ifndef AIRICE_GSL_STANDIN_INTEGRATION_H
#define AIRICE_GSL_STANDIN_INTEGRATION_H
#include <stdlib.h>
#include <gsl/gsl_math.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  size_t limit, size, nrmax, i, maximum_level;
  double *alist, *blist, *rlist, *elist;
  size_t *order, *level;
} gsl_integration_workspace;
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qags(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                         size_t limit, gsl_integration_workspace *workspace, double *result,
                         double *abserr);
extern size_t gsl_standin_qags_last_size;   /* stand-in only: intervals the last call used */
#ifdef __cplusplus
}
#endif
#endif
