/* TEST INFRASTRUCTURE ONLY - minimal stand-in for <gsl/gsl_math.h>. */
#ifndef AIRICE_GSL_STANDIN_MATH_H
#define AIRICE_GSL_STANDIN_MATH_H
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_NAN (NAN)
#define GSL_POSINF (INFINITY)
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))
struct gsl_function_struct {
  double (*function)(double x, void *params);
  void *params;
};
typedef struct gsl_function_struct gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
struct gsl_function_fdf_struct {
  double (*f)(double x, void *params);
  double (*df)(double x, void *params);
  void (*fdf)(double x, void *params, double *f, double *df);
  void *params;
};
typedef struct gsl_function_fdf_struct gsl_function_fdf;
#define GSL_FN_FDF_EVAL_F(FDF, x) (*((FDF)->f))(x, (FDF)->params)
#define GSL_FN_FDF_EVAL_DF(FDF, x) (*((FDF)->df))(x, (FDF)->params)
#define GSL_FN_FDF_EVAL_F_DF(FDF, x, y, dy) (*((FDF)->fdf))(x, (FDF)->params, (y), (dy))
#ifdef __cplusplus
}
#endif
#endif
