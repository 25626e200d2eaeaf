/* TEST INFRASTRUCTURE ONLY -- not part of the product, never linked into it.
 *
 * Minimal stand-in for the parts of GNU GSL that the reference sources under
 * /root/reference call (GSL is an un-vendored dependency of the reference,
 * README.md:37 "GSL 2.4"; prebuilt objects link libgsl.so.23/.27; it is absent from
 * this image and there is no network).  It exists so that the UNMODIFIED reference
 * sources compile into oracle/_ref/ and can serve as the parity oracle.
 *
 * Every routine restates the algorithm GSL publishes for it:
 *   roots/fsolver.c, roots/bisection.c, roots/brent.c, roots/falsepos.c,
 *   roots/newton.c, roots/convergence.c, deriv/deriv.c,
 *   interpolation/cspline.c + linalg/tridiag.c (natural spline).
 * Bisection and gsl_root_test_interval are pure IEEE add/halve/compare, so they are
 * exact restatements.  Brent / falsepos / Newton / deriv_central / cspline follow the
 * published recurrences term by term but cannot be cross-checked against a real
 * libgsl here; call sites that depend on them say so in their tests.
 *
 * One deliberate difference: solver state is calloc'ed (GSL mallocs).  It only matters
 * when gsl_root_fsolver_set() fails early (non-finite endpoint, reference ignores the
 * return code, MultiRayAirIceRefraction.cc:351) -- real GSL then iterates on
 * uninitialised memory; here the state is zero, which makes oracle runs reproducible.
 * Those cases are excluded from numeric parity (DESIGN.md "undefined-behaviour cases").
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <gsl/gsl_deriv.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_roots.h>
#include <gsl/gsl_spline.h>

/* ---------------------------------------------------------------- errno */
gsl_error_handler_t *gsl_set_error_handler_off(void) { return NULL; }
const char *gsl_strerror(const int e) {
  switch (e) {
    case GSL_SUCCESS: return "success";
    case GSL_CONTINUE: return "the iteration has not converged yet";
    case GSL_EINVAL: return "invalid argument supplied by user";
    case GSL_EBADFUNC: return "problem with user-supplied function";
    case GSL_EZERODIV: return "tried to divide by zero";
    default: return "error";
  }
}

/* ---------------------------------------------------------------- fsolver driver */
gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *T) {
  gsl_root_fsolver *s = (gsl_root_fsolver *)calloc(1, sizeof(gsl_root_fsolver));
  if (!s) return NULL;
  s->state = calloc(1, T->size);
  s->type = T;
  s->function = NULL;
  return s;
}
void gsl_root_fsolver_free(gsl_root_fsolver *s) {
  if (!s) return;
  free(s->state);
  free(s);
}
int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double x_lower, double x_upper) {
  if (x_lower > x_upper) return GSL_EINVAL;
  s->function = f;
  s->root = 0.5 * (x_lower + x_upper);
  s->x_lower = x_lower;
  s->x_upper = x_upper;
  return (s->type->set)(s->state, s->function, &(s->root), x_lower, x_upper);
}
int gsl_root_fsolver_iterate(gsl_root_fsolver *s) {
  return (s->type->iterate)(s->state, s->function, &(s->root), &(s->x_lower), &(s->x_upper));
}
const char *gsl_root_fsolver_name(const gsl_root_fsolver *s) { return s->type->name; }
double gsl_root_fsolver_root(const gsl_root_fsolver *s) { return s->root; }
double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s) { return s->x_lower; }
double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s) { return s->x_upper; }

#define SAFE_CALL(f, x, yp)                 \
  do {                                      \
    *(yp) = GSL_FN_EVAL(f, x);              \
    if (!isfinite(*(yp))) return GSL_EBADFUNC; \
  } while (0)

static int straddle_check(double fl, double fu) {
  if ((fl < 0.0 && fu < 0.0) || (fl > 0.0 && fu > 0.0)) return GSL_EINVAL;
  return GSL_SUCCESS;
}

/* ---------------------------------------------------------------- bisection */
typedef struct { double f_lower, f_upper; } bracket_state_t;

static int bracket_init(void *vstate, gsl_function *f, double *root, double x_lower, double x_upper) {
  bracket_state_t *st = (bracket_state_t *)vstate;
  double fl, fu;
  *root = 0.5 * (x_lower + x_upper);
  SAFE_CALL(f, x_lower, &fl);
  SAFE_CALL(f, x_upper, &fu);
  st->f_lower = fl;
  st->f_upper = fu;
  return straddle_check(fl, fu);
}

static int bisection_iterate(void *vstate, gsl_function *f, double *root, double *x_lower,
                             double *x_upper) {
  bracket_state_t *st = (bracket_state_t *)vstate;
  const double xl = *x_lower, xr = *x_upper;
  const double fl = st->f_lower, fu = st->f_upper;
  double xb, fb;
  if (fl == 0.0) { *root = xl; *x_upper = xl; return GSL_SUCCESS; }
  if (fu == 0.0) { *root = xr; *x_lower = xr; return GSL_SUCCESS; }
  xb = (xl + xr) / 2.0;
  SAFE_CALL(f, xb, &fb);
  if (fb == 0.0) { *root = xb; *x_lower = xb; *x_upper = xb; return GSL_SUCCESS; }
  if ((fl > 0.0 && fb < 0.0) || (fl < 0.0 && fb > 0.0)) {
    *root = 0.5 * (xl + xb);
    *x_upper = xb;
    st->f_upper = fb;
  } else {
    *root = 0.5 * (xb + xr);
    *x_lower = xb;
    st->f_lower = fb;
  }
  return GSL_SUCCESS;
}
static const gsl_root_fsolver_type bisection_type = {"bisection", sizeof(bracket_state_t),
                                                     &bracket_init, &bisection_iterate};
const gsl_root_fsolver_type *gsl_root_fsolver_bisection = &bisection_type;

/* ---------------------------------------------------------------- false position */
static int falsepos_iterate(void *vstate, gsl_function *f, double *root, double *x_lower,
                            double *x_upper) {
  bracket_state_t *st = (bracket_state_t *)vstate;
  const double xl = *x_lower, xr = *x_upper;
  const double fl = st->f_lower, fu = st->f_upper;
  double x_lin, f_lin, xb, fb, w;
  if (fl == 0.0) { *root = xl; *x_upper = xl; return GSL_SUCCESS; }
  if (fu == 0.0) { *root = xr; *x_lower = xr; return GSL_SUCCESS; }
  x_lin = xr - (fu * (xl - xr) / (fl - fu));
  SAFE_CALL(f, x_lin, &f_lin);
  if (f_lin == 0.0) { *root = x_lin; *x_lower = x_lin; *x_upper = x_lin; return GSL_SUCCESS; }
  if ((fl > 0.0 && f_lin < 0.0) || (fl < 0.0 && f_lin > 0.0)) {
    *root = x_lin; *x_upper = x_lin; st->f_upper = f_lin; w = x_lin - xl;
  } else {
    *root = x_lin; *x_lower = x_lin; st->f_lower = f_lin; w = xr - x_lin;
  }
  if (w < 0.5 * (xr - xl)) return GSL_SUCCESS;
  xb = 0.5 * (xl + xr);
  SAFE_CALL(f, xb, &fb);
  if ((fl > 0.0 && fb < 0.0) || (fl < 0.0 && fb > 0.0)) {
    *x_upper = xb; st->f_upper = fb;
    if (*root > xb) *root = 0.5 * (xl + xb);
  } else {
    *x_lower = xb; st->f_lower = fb;
    if (*root < xb) *root = 0.5 * (xb + xr);
  }
  return GSL_SUCCESS;
}
static const gsl_root_fsolver_type falsepos_type = {"falsepos", sizeof(bracket_state_t),
                                                    &bracket_init, &falsepos_iterate};
const gsl_root_fsolver_type *gsl_root_fsolver_falsepos = &falsepos_type;

/* ---------------------------------------------------------------- Brent-Dekker */
typedef struct { double a, b, c, d, e; double fa, fb, fc; } brent_state_t;

static int brent_init(void *vstate, gsl_function *f, double *root, double x_lower, double x_upper) {
  brent_state_t *st = (brent_state_t *)vstate;
  double fl, fu;
  *root = 0.5 * (x_lower + x_upper);
  SAFE_CALL(f, x_lower, &fl);
  SAFE_CALL(f, x_upper, &fu);
  st->a = x_lower; st->fa = fl;
  st->b = x_upper; st->fb = fu;
  st->c = x_upper; st->fc = fu;
  st->d = x_upper - x_lower;
  st->e = x_upper - x_lower;
  return straddle_check(fl, fu);
}

static int brent_iterate(void *vstate, gsl_function *f, double *root, double *x_lower,
                         double *x_upper) {
  brent_state_t *st = (brent_state_t *)vstate;
  double tol, m;
  int ac_equal = 0;
  double a = st->a, b = st->b, c = st->c;
  double fa = st->fa, fb = st->fb, fc = st->fc;
  double d = st->d, e = st->e;

  if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) {
    ac_equal = 1; c = a; fc = fa; d = b - a; e = b - a;
  }
  if (fabs(fc) < fabs(fb)) {
    ac_equal = 1; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa;
  }
  tol = 0.5 * GSL_DBL_EPSILON * fabs(b);
  m = 0.5 * (c - b);
  if (fb == 0) { *root = b; *x_lower = b; *x_upper = b; return GSL_SUCCESS; }
  if (fabs(m) <= tol) {
    *root = b;
    if (b < c) { *x_lower = b; *x_upper = c; } else { *x_lower = c; *x_upper = b; }
    return GSL_SUCCESS;
  }
  if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
    d = m; e = m; /* bisection */
  } else {
    double p, q, r;
    double s = fb / fa;
    if (ac_equal) {
      p = 2 * m * s; q = 1 - s;
    } else {
      q = fa / fc; r = fb / fc;
      p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
      q = (q - 1) * (r - 1) * (s - 1);
    }
    if (p > 0) q = -q; else p = -p;
    if (2 * p < GSL_MIN(3 * m * q - fabs(tol * q), fabs(e * q))) {
      e = d; d = p / q;
    } else {
      d = m; e = m;
    }
  }
  a = b; fa = fb;
  if (fabs(d) > tol) b += d; else b += (m > 0 ? +tol : -tol);
  SAFE_CALL(f, b, &fb);
  st->a = a; st->b = b; st->c = c; st->d = d; st->e = e;
  st->fa = fa; st->fb = fb; st->fc = fc;
  *root = b;
  if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
  if (b < c) { *x_lower = b; *x_upper = c; } else { *x_lower = c; *x_upper = b; }
  return GSL_SUCCESS;
}
static const gsl_root_fsolver_type brent_type = {"brent", sizeof(brent_state_t), &brent_init,
                                                 &brent_iterate};
const gsl_root_fsolver_type *gsl_root_fsolver_brent = &brent_type;

/* ---------------------------------------------------------------- Newton (fdf) */
typedef struct { double f, df; } newton_state_t;
static int newton_init(void *vstate, gsl_function_fdf *fdf, double *root) {
  newton_state_t *st = (newton_state_t *)vstate;
  const double x = *root;
  st->f = GSL_FN_FDF_EVAL_F(fdf, x);
  st->df = GSL_FN_FDF_EVAL_DF(fdf, x);
  return GSL_SUCCESS;
}
static int newton_iterate(void *vstate, gsl_function_fdf *fdf, double *root) {
  newton_state_t *st = (newton_state_t *)vstate;
  double root_new, f_new, df_new;
  if (st->df == 0.0) return GSL_EZERODIV;
  root_new = *root - (st->f / st->df);
  *root = root_new;
  GSL_FN_FDF_EVAL_F_DF(fdf, root_new, &f_new, &df_new);
  st->f = f_new;
  st->df = df_new;
  if (!isfinite(f_new)) return GSL_EBADFUNC;
  if (!isfinite(df_new)) return GSL_EBADFUNC;
  return GSL_SUCCESS;
}
static const gsl_root_fdfsolver_type newton_type = {"newton", sizeof(newton_state_t), &newton_init,
                                                    &newton_iterate};
const gsl_root_fdfsolver_type *gsl_root_fdfsolver_newton = &newton_type;

gsl_root_fdfsolver *gsl_root_fdfsolver_alloc(const gsl_root_fdfsolver_type *T) {
  gsl_root_fdfsolver *s = (gsl_root_fdfsolver *)calloc(1, sizeof(gsl_root_fdfsolver));
  if (!s) return NULL;
  s->state = calloc(1, T->size);
  s->type = T;
  s->fdf = NULL;
  return s;
}
int gsl_root_fdfsolver_set(gsl_root_fdfsolver *s, gsl_function_fdf *fdf, double root) {
  s->fdf = fdf;
  s->root = root;
  return (s->type->set)(s->state, s->fdf, &(s->root));
}
int gsl_root_fdfsolver_iterate(gsl_root_fdfsolver *s) {
  return (s->type->iterate)(s->state, s->fdf, &(s->root));
}
void gsl_root_fdfsolver_free(gsl_root_fdfsolver *s) {
  if (!s) return;
  free(s->state);
  free(s);
}
const char *gsl_root_fdfsolver_name(const gsl_root_fdfsolver *s) { return s->type->name; }
double gsl_root_fdfsolver_root(const gsl_root_fdfsolver *s) { return s->root; }

/* ---------------------------------------------------------------- convergence tests */
int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel) {
  const double abs_lower = fabs(x_lower);
  const double abs_upper = fabs(x_upper);
  double min_abs, tolerance;
  if (epsrel < 0.0) return GSL_EBADTOL;
  if (epsabs < 0.0) return GSL_EBADTOL;
  if (x_lower > x_upper) return GSL_EINVAL;
  if ((x_lower > 0.0 && x_upper > 0.0) || (x_lower < 0.0 && x_upper < 0.0))
    min_abs = GSL_MIN(abs_lower, abs_upper);
  else
    min_abs = 0;
  tolerance = epsabs + epsrel * min_abs;
  if (fabs(x_upper - x_lower) < tolerance) return GSL_SUCCESS;
  return GSL_CONTINUE;
}
int gsl_root_test_residual(double f, double epsabs) {
  if (epsabs < 0.0) return GSL_EBADTOL;
  if (fabs(f) < epsabs) return GSL_SUCCESS;
  return GSL_CONTINUE;
}
int gsl_root_test_delta(double x1, double x0, double epsabs, double epsrel) {
  const double tolerance = epsabs + epsrel * fabs(x1);
  if (epsabs < 0.0) return GSL_EBADTOL;
  if (epsrel < 0.0) return GSL_EBADTOL;
  if (fabs(x1 - x0) < tolerance || x1 == x0) return GSL_SUCCESS;
  return GSL_CONTINUE;
}

/* ---------------------------------------------------------------- deriv_central */
static void central_deriv(const gsl_function *f, double x, double h, double *result,
                          double *abserr_round, double *abserr_trunc) {
  double fm1 = GSL_FN_EVAL(f, x - h);
  double fp1 = GSL_FN_EVAL(f, x + h);
  double fmh = GSL_FN_EVAL(f, x - h / 2);
  double fph = GSL_FN_EVAL(f, x + h / 2);
  double r3 = 0.5 * (fp1 - fm1);
  double r5 = (4.0 / 3.0) * (fph - fmh) - (1.0 / 3.0) * r3;
  double e3 = (fabs(fp1) + fabs(fm1)) * GSL_DBL_EPSILON;
  double e5 = 2.0 * (fabs(fph) + fabs(fmh)) * GSL_DBL_EPSILON + e3;
  double dy = GSL_MAX(fabs(r3 / h), fabs(r5 / h)) * (fabs(x) / h) * GSL_DBL_EPSILON;
  *result = r5 / h;
  *abserr_trunc = fabs((r5 - r3) / h);
  *abserr_round = fabs(e5 / h) + dy;
}
int gsl_deriv_central(const gsl_function *f, double x, double h, double *result, double *abserr) {
  double r_0, round, trunc, error;
  central_deriv(f, x, h, &r_0, &round, &trunc);
  error = round + trunc;
  if (round < trunc && (round > 0 && trunc > 0)) {
    double r_opt, round_opt, trunc_opt, error_opt;
    double h_opt = h * pow(round / (2.0 * trunc), 1.0 / 3.0);
    central_deriv(f, x, h_opt, &r_opt, &round_opt, &trunc_opt);
    error_opt = round_opt + trunc_opt;
    if (error_opt < error && fabs(r_opt - r_0) < 4.0 * error) {
      r_0 = r_opt;
      error = error_opt;
    }
  }
  *result = r_0;
  *abserr = error;
  return GSL_SUCCESS;
}

/* ---------------------------------------------------------------- natural cubic spline */
static const gsl_interp_type cspline_type = {"cspline", 3};
const gsl_interp_type *gsl_interp_cspline = &cspline_type;

gsl_interp_accel *gsl_interp_accel_alloc(void) {
  return (gsl_interp_accel *)calloc(1, sizeof(gsl_interp_accel));
}
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size) {
  gsl_spline *s = (gsl_spline *)calloc(1, sizeof(gsl_spline));
  if (!s) return NULL;
  s->type = T;
  s->size = size;
  s->x = (double *)malloc(size * sizeof(double));
  s->y = (double *)malloc(size * sizeof(double));
  s->c = (double *)calloc(size, sizeof(double));
  return s;
}
void gsl_spline_free(gsl_spline *s) {
  if (!s) return;
  free(s->x); free(s->y); free(s->c); free(s);
}

/* Symmetric positive-definite tridiagonal solve by L.D.L^T (linalg/tridiag.c solve_tridiag). */
static void solve_symm_tridiag(const double *diag, const double *offdiag, const double *b, double *x,
                               size_t N) {
  double *gamma = (double *)malloc(N * sizeof(double));
  double *alpha = (double *)malloc(N * sizeof(double));
  double *c = (double *)malloc(N * sizeof(double));
  double *z = (double *)malloc(N * sizeof(double));
  size_t i, j;
  alpha[0] = diag[0];
  gamma[0] = offdiag[0] / alpha[0];
  for (i = 1; i < N - 1; i++) {
    alpha[i] = diag[i] - offdiag[i - 1] * gamma[i - 1];
    gamma[i] = offdiag[i] / alpha[i];
  }
  if (N > 1) alpha[N - 1] = diag[N - 1] - offdiag[N - 2] * gamma[N - 2];
  z[0] = b[0];
  for (i = 1; i < N; i++) z[i] = b[i] - gamma[i - 1] * z[i - 1];
  for (i = 0; i < N; i++) c[i] = z[i] / alpha[i];
  x[N - 1] = c[N - 1];
  if (N >= 2) {
    for (i = N - 2, j = 0; j <= N - 2; j++, i--) x[i] = c[i] - gamma[i] * x[i + 1];
  }
  free(gamma); free(alpha); free(c); free(z);
}

int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size) {
  size_t i;
  const size_t max_index = size - 1;
  const size_t sys_size = max_index - 1;
  double *g, *diag, *offdiag;
  if (size != s->size) return GSL_EINVAL;
  memcpy(s->x, xa, size * sizeof(double));
  memcpy(s->y, ya, size * sizeof(double));
  s->c[0] = 0.0;
  s->c[max_index] = 0.0;
  g = (double *)malloc(size * sizeof(double));
  diag = (double *)malloc(size * sizeof(double));
  offdiag = (double *)malloc(size * sizeof(double));
  for (i = 0; i < sys_size; i++) {
    const double h_i = xa[i + 1] - xa[i];
    const double h_ip1 = xa[i + 2] - xa[i + 1];
    const double ydiff_i = ya[i + 1] - ya[i];
    const double ydiff_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
    const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    offdiag[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
  }
  if (sys_size == 1) {
    s->c[1] = g[0] / diag[0];
  } else {
    solve_symm_tridiag(diag, offdiag, g, s->c + 1, sys_size);
  }
  free(g); free(diag); free(offdiag);
  return GSL_SUCCESS;
}

double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a) {
  size_t lo = 0, hi = s->size - 1;
  (void)a;
  if (x < s->x[0] || x > s->x[s->size - 1]) return GSL_NAN;
  while (hi > lo + 1) { /* gsl_interp_bsearch: x[lo] <= x < x[hi] */
    size_t mid = (hi + lo) / 2;
    if (s->x[mid] > x) hi = mid; else lo = mid;
  }
  {
    const double x_lo = s->x[lo], x_hi = s->x[lo + 1];
    const double dx = x_hi - x_lo;
    if (dx > 0.0) {
      const double y_lo = s->y[lo], y_hi = s->y[lo + 1];
      const double dy = y_hi - y_lo;
      const double delx = x - x_lo;
      const double c_i = s->c[lo], c_ip1 = s->c[lo + 1];
      const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
      const double d_i = (c_ip1 - c_i) / (3.0 * dx);
      return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
    }
    return 0.0;
  }
}

/* ---------------------------------------------------------------- integration: gsl_integration_qags
 * (integration/qags.c, qk.c, qk21.c, qpsrt.c, qelg.c, util.c of GSL 2.x = QUADPACK's dqagse / dqk21 / dqpsrt / dqelg):
 * 21-point Gauss-Kronrod rule, bisection of the interval with the largest error estimate, Wynn's epsilon algorithm on
 * the sequence of totals once the smallest intervals are reached.  The in-ice attenuation integrals
 * (IceRayTracing.cc:179-219) call it with epsabs 0, epsrel 1e-7, limit 1000; direct and reflected rays return from the
 * first rule, refracted rays (1/sqrt end-point singularity at the turning depth) need the extrapolation. */
#define GSL_DBL_MIN_ 2.2250738585072014e-308
#define GSL_DBL_MAX_ 1.7976931348623157e+308
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n) {
  gsl_integration_workspace *w = (gsl_integration_workspace *)calloc(1, sizeof(*w));
  if (!w || n == 0) { free(w); return 0; }
  w->alist = (double *)calloc(n, sizeof(double));
  w->blist = (double *)calloc(n, sizeof(double));
  w->rlist = (double *)calloc(n, sizeof(double));
  w->elist = (double *)calloc(n, sizeof(double));
  w->order = (size_t *)calloc(n, sizeof(size_t));
  w->level = (size_t *)calloc(n, sizeof(size_t));
  w->limit = n;
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) {
  if (!w) return;
  free(w->alist); free(w->blist); free(w->rlist); free(w->elist); free(w->order); free(w->level);
  free(w);
}

static double qk_rescale_error(double err, const double result_abs, const double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    double scale = pow((200 * err / result_asc), 1.5);
    if (scale < 1) err = result_asc * scale;
    else err = result_asc;
  }
  if (result_abs > GSL_DBL_MIN_ / (50 * GSL_DBL_EPSILON)) {
    double min_err = 50 * GSL_DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

/* gsl_integration_qk21 through the generic gsl_integration_qk with n = 11 */
static void qk21(const gsl_function *f, double a, double b, double *result, double *abserr, double *resabs, double *resasc) {
  static const double xgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
                                 0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
                                 0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                                 0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
                                 0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
                                 0.000000000000000000000000000000000};
  static const double wg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
                               0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
                               0.295524224714752870173815619188769};
  static const double wgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
                                 0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
                                 0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                                 0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
                                 0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
                                 0.149445554002916905664936468389821};
  const int n = 11;
  double fv1[11], fv2[11];
  const double center = 0.5 * (a + b);
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  const double f_center = GSL_FN_EVAL(f, center);
  double result_gauss = 0;
  double result_kronrod = f_center * wgk[n - 1];
  double result_abs = fabs(result_kronrod);
  double result_asc = 0;
  double mean = 0, err = 0;
  int j;
  if (n % 2 == 0) result_gauss = f_center * wg[n / 2 - 1];
  for (j = 0; j < (n - 1) / 2; j++) {
    const int jtw = j * 2 + 1;
    const double abscissa = half_length * xgk[jtw];
    const double fval1 = GSL_FN_EVAL(f, center - abscissa);
    const double fval2 = GSL_FN_EVAL(f, center + abscissa);
    const double fsum = fval1 + fval2;
    fv1[jtw] = fval1;
    fv2[jtw] = fval2;
    result_gauss += wg[j] * fsum;
    result_kronrod += wgk[jtw] * fsum;
    result_abs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
  }
  for (j = 0; j < n / 2; j++) {
    int jtwm1 = j * 2;
    const double abscissa = half_length * xgk[jtwm1];
    const double fval1 = GSL_FN_EVAL(f, center - abscissa);
    const double fval2 = GSL_FN_EVAL(f, center + abscissa);
    fv1[jtwm1] = fval1;
    fv2[jtwm1] = fval2;
    result_kronrod += wgk[jtwm1] * (fval1 + fval2);
    result_abs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
  }
  mean = result_kronrod * 0.5;
  result_asc = wgk[n - 1] * fabs(f_center - mean);
  for (j = 0; j < n - 1; j++) result_asc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
  err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  *result = result_kronrod;
  *resabs = result_abs;
  *resasc = result_asc;
  *abserr = qk_rescale_error(err, result_abs, result_asc);
}

/* qpsrt.c: keep `order` sorted by decreasing error estimate */
static void qpsrt(gsl_integration_workspace *workspace) {
  const size_t last = workspace->size - 1;
  const size_t limit = workspace->limit;
  double *elist = workspace->elist;
  size_t *order = workspace->order;
  double errmax, errmin;
  int i, k, top;
  size_t i_nrmax = workspace->nrmax;
  size_t i_maxerr = order[i_nrmax];
  if (last < 2) {
    order[0] = 0;
    order[1] = 1;
    workspace->i = i_maxerr;
    return;
  }
  errmax = elist[i_maxerr];
  while (i_nrmax > 0 && errmax > elist[order[i_nrmax - 1]]) {
    order[i_nrmax] = order[i_nrmax - 1];
    i_nrmax--;
  }
  if (last < (limit / 2 + 2)) top = (int)last;
  else top = (int)(limit - last + 1);
  i = (int)i_nrmax + 1;
  while (i < top && errmax < elist[order[i]]) {
    order[i - 1] = order[i];
    i++;
  }
  order[i - 1] = i_maxerr;
  errmin = elist[last];
  k = top - 1;
  while (k > i - 2 && errmin >= elist[order[k]]) {
    order[k + 1] = order[k];
    k--;
  }
  order[k + 1] = last;
  i_maxerr = order[i_nrmax];
  workspace->i = i_maxerr;
  workspace->nrmax = i_nrmax;
}

static void ws_update(gsl_integration_workspace *workspace, double a1, double b1, double area1, double error1,
                      double a2, double b2, double area2, double error2) {
  double *alist = workspace->alist, *blist = workspace->blist, *rlist = workspace->rlist, *elist = workspace->elist;
  size_t *level = workspace->level;
  const size_t i_max = workspace->i;
  const size_t i_new = workspace->size;
  const size_t new_level = workspace->level[i_max] + 1;
  if (error2 > error1) {
    alist[i_max] = a2;
    rlist[i_max] = area2;
    elist[i_max] = error2;
    level[i_max] = new_level;
    alist[i_new] = a1;
    blist[i_new] = b1;
    rlist[i_new] = area1;
    elist[i_new] = error1;
    level[i_new] = new_level;
  } else {
    blist[i_max] = b1;
    rlist[i_max] = area1;
    elist[i_max] = error1;
    level[i_max] = new_level;
    alist[i_new] = a2;
    blist[i_new] = b2;
    rlist[i_new] = area2;
    elist[i_new] = error2;
    level[i_new] = new_level;
  }
  workspace->size++;
  if (new_level > workspace->maximum_level) workspace->maximum_level = new_level;
  qpsrt(workspace);
}

static int ws_increase_nrmax(gsl_integration_workspace *workspace) {
  int k;
  int id = (int)workspace->nrmax;
  int jupbnd;
  const size_t *level = workspace->level;
  const size_t *order = workspace->order;
  size_t limit = workspace->limit;
  size_t last = workspace->size - 1;
  if (last > (1 + limit / 2)) jupbnd = (int)(limit + 1 - last);
  else jupbnd = (int)last;
  for (k = id; k <= jupbnd; k++) {
    size_t i_max = order[workspace->nrmax];
    workspace->i = i_max;
    if (level[i_max] < workspace->maximum_level) return 1;
    workspace->nrmax++;
  }
  return 0;
}

struct extrapolation_table { size_t n; double rlist2[52]; size_t nres; double res3la[3]; };

/* qelg.c: Wynn's epsilon algorithm */
static void qelg(struct extrapolation_table *table, double *result, double *abserr) {
  double *epstab = table->rlist2;
  double *res3la = table->res3la;
  const size_t n = table->n - 1;
  const double current = epstab[n];
  double absolute = GSL_DBL_MAX_;
  double relative = 5 * GSL_DBL_EPSILON * fabs(current);
  const size_t newelm = n / 2;
  const size_t n_orig = n;
  size_t n_final = n;
  size_t i;
  const size_t nres_orig = table->nres;
  *result = current;
  *abserr = GSL_DBL_MAX_;
  if (n < 2) {
    *result = current;
    *abserr = GSL_MAX(absolute, relative);
    return;
  }
  epstab[n + 2] = epstab[n];
  epstab[n] = GSL_DBL_MAX_;
  for (i = 0; i < newelm; i++) {
    double res = epstab[n - 2 * i + 2];
    double e0 = epstab[n - 2 * i - 2];
    double e1 = epstab[n - 2 * i - 1];
    double e2 = res;
    double e1abs = fabs(e1);
    double delta2 = e2 - e1;
    double err2 = fabs(delta2);
    double tol2 = GSL_MAX(fabs(e2), e1abs) * GSL_DBL_EPSILON;
    double delta3 = e1 - e0;
    double err3 = fabs(delta3);
    double tol3 = GSL_MAX(e1abs, fabs(e0)) * GSL_DBL_EPSILON;
    double e3, delta1, err1, tol1, ss;
    if (err2 <= tol2 && err3 <= tol3) {
      *result = res;
      absolute = err2 + err3;
      relative = 5 * GSL_DBL_EPSILON * fabs(res);
      *abserr = GSL_MAX(absolute, relative);
      return;
    }
    e3 = epstab[n - 2 * i];
    epstab[n - 2 * i] = e1;
    delta1 = e1 - e3;
    err1 = fabs(delta1);
    tol1 = GSL_MAX(e1abs, fabs(e3)) * GSL_DBL_EPSILON;
    if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
      n_final = 2 * i;
      break;
    }
    ss = (1 / delta1 + 1 / delta2) - 1 / delta3;
    if (fabs(ss * e1) <= 0.0001) {
      n_final = 2 * i;
      break;
    }
    res = e1 + 1 / ss;
    epstab[n - 2 * i] = res;
    {
      const double error = err2 + fabs(res - e2) + err3;
      if (error <= *abserr) {
        *abserr = error;
        *result = res;
      }
    }
  }
  {
    const size_t limexp = 50 - 1;
    if (n_final == limexp) n_final = 2 * (limexp / 2);
  }
  if (n_orig % 2 == 1) {
    for (i = 0; i <= newelm; i++) epstab[1 + i * 2] = epstab[i * 2 + 3];
  } else {
    for (i = 0; i <= newelm; i++) epstab[i * 2] = epstab[i * 2 + 2];
  }
  if (n_orig != n_final) {
    for (i = 0; i <= n_final; i++) epstab[i] = epstab[n_orig - n_final + i];
  }
  table->n = n_final + 1;
  if (nres_orig < 3) {
    res3la[nres_orig] = *result;
    *abserr = GSL_DBL_MAX_;
  } else {
    *abserr = (fabs(*result - res3la[2]) + fabs(*result - res3la[1]) + fabs(*result - res3la[0]));
    res3la[0] = res3la[1];
    res3la[1] = res3la[2];
    res3la[2] = *result;
  }
  table->nres = nres_orig + 1;
  *abserr = GSL_MAX(*abserr, 5 * GSL_DBL_EPSILON * fabs(*result));
}

size_t gsl_standin_qags_last_size = 0;     /* diagnostics for the tests: intervals used by the last call */

int gsl_integration_qags(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                         size_t limit, gsl_integration_workspace *workspace, double *result, double *abserr) {
  double area, errsum;
  double res_ext, err_ext;
  double result0, abserr0, resabs0, resasc0;
  double tolerance;
  double ertest = 0;
  double error_over_large_intervals = 0;
  double reseps = 0, abseps = 0, correc = 0;
  size_t ktmin = 0;
  int roundoff_type1 = 0, roundoff_type2 = 0, roundoff_type3 = 0;
  int error_type = 0, error_type2 = 0;
  size_t iteration = 0;
  int positive_integrand = 0;
  int extrapolate = 0;
  int disallow_extrapolation = 0;
  struct extrapolation_table table;
  size_t k;

  /* initialise (workspace, a, b) */
  workspace->size = 0; workspace->nrmax = 0; workspace->i = 0;
  workspace->alist[0] = a; workspace->blist[0] = b; workspace->rlist[0] = 0.0; workspace->elist[0] = 0.0;
  workspace->order[0] = 0; workspace->level[0] = 0; workspace->maximum_level = 0;
  *result = 0;
  *abserr = 0;
  gsl_standin_qags_last_size = 0;
  if (limit > workspace->limit) return GSL_EINVAL;
  if (epsabs <= 0 && (epsrel < 50 * GSL_DBL_EPSILON || epsrel < 0.5e-28)) return GSL_EBADTOL;

  qk21(f, a, b, &result0, &abserr0, &resabs0, &resasc0);
  workspace->size = 1; workspace->rlist[0] = result0; workspace->elist[0] = abserr0;   /* set_initial_result */
  gsl_standin_qags_last_size = 1;
  tolerance = GSL_MAX(epsabs, epsrel * fabs(result0));
  if (abserr0 <= 100 * GSL_DBL_EPSILON * resabs0 && abserr0 > tolerance) {
    *result = result0; *abserr = abserr0;
    return GSL_EROUND;
  } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
    *result = result0; *abserr = abserr0;
    return GSL_SUCCESS;
  } else if (limit == 1) {
    *result = result0; *abserr = abserr0;
    return GSL_EMAXITER;
  }

  table.n = 0; table.nres = 0;
  table.rlist2[table.n] = result0; table.n++;        /* append_table */
  area = result0;
  errsum = abserr0;
  res_ext = result0;
  err_ext = GSL_DBL_MAX_;
  positive_integrand = (fabs(result0) >= (1 - 50 * GSL_DBL_EPSILON) * resabs0);
  iteration = 1;

  do {
    size_t current_level;
    double a1, b1, a2, b2;
    double a_i, b_i, r_i, e_i;
    double area1 = 0, area2 = 0, area12 = 0;
    double error1 = 0, error2 = 0, error12 = 0;
    double resasc1, resasc2;
    double resabs1, resabs2;
    double last_e_i;

    a_i = workspace->alist[workspace->i]; b_i = workspace->blist[workspace->i];       /* retrieve */
    r_i = workspace->rlist[workspace->i]; e_i = workspace->elist[workspace->i];
    current_level = workspace->level[workspace->i] + 1;
    a1 = a_i;
    b1 = 0.5 * (a_i + b_i);
    a2 = b1;
    b2 = b_i;
    iteration++;
    qk21(f, a1, b1, &area1, &error1, &resabs1, &resasc1);
    qk21(f, a2, b2, &area2, &error2, &resabs2, &resasc2);
    area12 = area1 + area2;
    error12 = error1 + error2;
    last_e_i = e_i;
    errsum = errsum + error12 - e_i;
    area = area + area12 - r_i;
    tolerance = GSL_MAX(epsabs, epsrel * fabs(area));
    if (resasc1 != error1 && resasc2 != error2) {
      double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) {
        if (!extrapolate) roundoff_type1++;
        else roundoff_type2++;
      }
      if (iteration > 10 && error12 > e_i) roundoff_type3++;
    }
    if (roundoff_type1 + roundoff_type2 >= 10 || roundoff_type3 >= 20) error_type = 2;
    if (roundoff_type2 >= 5) error_type2 = 1;
    {                                                                  /* subinterval_too_small (a1, a2, b2) */
      const double e = GSL_DBL_EPSILON;
      const double u = GSL_DBL_MIN_;
      double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
      if (fabs(a1) <= tmp && fabs(b2) <= tmp) error_type = 4;
    }
    ws_update(workspace, a1, b1, area1, error1, a2, b2, area2, error2);
    gsl_standin_qags_last_size = workspace->size;
    if (errsum <= tolerance) goto compute_result;
    if (error_type) break;
    if (iteration >= limit - 1) {
      error_type = 1;
      break;
    }
    if (iteration == 2) {
      error_over_large_intervals = errsum;
      ertest = tolerance;
      table.rlist2[table.n] = area; table.n++;
      continue;
    }
    if (disallow_extrapolation) continue;
    error_over_large_intervals += -last_e_i;
    if (current_level < workspace->maximum_level) error_over_large_intervals += error12;
    if (!extrapolate) {
      if (workspace->level[workspace->i] < workspace->maximum_level) continue;     /* large_interval */
      extrapolate = 1;
      workspace->nrmax = 1;
    }
    if (!error_type2 && error_over_large_intervals > ertest) {
      if (ws_increase_nrmax(workspace)) continue;
    }
    table.rlist2[table.n] = area; table.n++;
    qelg(&table, &reseps, &abseps);
    ktmin++;
    if (ktmin > 5 && err_ext < 0.001 * errsum) error_type = 5;
    if (abseps < err_ext) {
      ktmin = 0;
      err_ext = abseps;
      res_ext = reseps;
      correc = error_over_large_intervals;
      ertest = GSL_MAX(epsabs, epsrel * fabs(reseps));
      if (err_ext <= ertest) break;
    }
    if (table.n == 1) disallow_extrapolation = 1;
    if (error_type == 5) break;
    workspace->nrmax = 0; workspace->i = workspace->order[0];           /* reset_nrmax */
    extrapolate = 0;
    error_over_large_intervals = errsum;
  } while (iteration < limit);

  *result = res_ext;
  *abserr = err_ext;
  if (err_ext == GSL_DBL_MAX_) goto compute_result;
  if (error_type || error_type2) {
    if (error_type2) err_ext += correc;
    if (error_type == 0) error_type = 3;
    if (res_ext != 0.0 && area != 0.0) {
      if (err_ext / fabs(res_ext) > errsum / fabs(area)) goto compute_result;
    } else if (err_ext > errsum) {
      goto compute_result;
    } else if (area == 0.0) {
      goto return_error;
    }
  }
  {
    double max_area = GSL_MAX(fabs(res_ext), fabs(area));
    if (!positive_integrand && max_area < 0.01 * resabs0) goto return_error;
  }
  {
    double ratio = res_ext / area;
    if (ratio < 0.01 || ratio > 100.0 || errsum > fabs(area)) error_type = 6;
  }
  goto return_error;

compute_result:
  {
    double result_sum = 0;
    for (k = 0; k < workspace->size; k++) result_sum += workspace->rlist[k];
    *result = result_sum;
  }
  *abserr = errsum;

return_error:
  if (error_type > 2) error_type--;
  if (error_type == 0) return GSL_SUCCESS;
  else if (error_type == 1) return GSL_EMAXITER;
  else if (error_type == 2) return GSL_EROUND;
  else if (error_type == 3) return GSL_ESING;
  else if (error_type == 4) return GSL_EROUND;
  else if (error_type == 5) return GSL_EDIVERGE;
  return GSL_EFAILED;
}
