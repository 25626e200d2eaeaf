/* TEST INFRASTRUCTURE ONLY - minimal stand-in for <gsl/gsl_spline.h>:
 * natural cubic spline only (gsl_interp_cspline). */
#ifndef AIRICE_GSL_STANDIN_SPLINE_H
#define AIRICE_GSL_STANDIN_SPLINE_H
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t cache; size_t miss_count; size_t hit_count; } gsl_interp_accel;
typedef struct { const char *name; unsigned int min_size; } gsl_interp_type;
typedef struct {
  const gsl_interp_type *type;
  double *x;
  double *y;
  double *c; /* second-derivative/2 coefficients */
  size_t size;
} gsl_spline;
extern const gsl_interp_type *gsl_interp_cspline;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *spline, const double xa[], const double ya[], size_t size);
double gsl_spline_eval(const gsl_spline *spline, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *spline);
#ifdef __cplusplus
}
#endif
#endif
