/* inice_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle; never linked into or called by the product).
 *
 * Plain-C restatement of the in-ice ray solver of uzairlatif90/AirIceRayTracing, IceRayTracing.cc, for the
 * configuration the reference ships (TransitionBoundary == 0, IceRayTracing.hh:49): direct, reflected and up to two
 * refracted rays between a transmitter (0, z0) and a receiver (x1, z1), both in the ice.  Each function cites the
 * reference lines it follows.  The GSL pieces (falsepos, newton, deriv_central, the stopping tests) come from
 * gsl_standin/, which restates GSL 2.x's published algorithms; the closed forms are written out literally (pow(n,2),
 * the long fpathD expression, exp() per evaluation) and NOT through the identities the CUDA code uses, so this file is
 * an independent check of those too.
 *
 * Pinned against the unmodified reference (oracle/_ref/libiceray_ref.so) by tests/test_oracle.py and against
 * tests/golden/inice.npz: all 29 outputs bit-equal.
 */
#include <math.h>
#include <stdlib.h>
#include <gsl/gsl_deriv.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_roots.h>

/* IceRayTracing.hh:41-56 (pi as written there) and the SetA/SetB/SetC setters */
static double g_A = 1.78, g_B = -0.43, g_C = 0.0132;
static const double k_pi = 3.14159265359, k_c = 299792458.0;

void inice_oracle_set_model(double A, double B, double C) { g_A = A; g_B = B; g_C = C; }

/* Getnz, IceRayTracing.cc:56-59 */
static double nz(double z) { z = fabs(z); return g_A + g_B * exp(-g_C * z); }

typedef struct { double a, b, c, l; } depth_par;      /* fDnfR_params / ftimeD_params (speed of light is global here) */
typedef struct { double a, b, c, z; } lpar;            /* fDnfR_L_params */
typedef struct { double a, z0, x1, z1; } pair_par;     /* fDanfRa_params */
typedef struct { double a, l; } minnz_par;             /* Minnz_params */

/* fDnfR, IceRayTracing.cc:355-365: ray path x(depth) for a given L */
static double f_depth(double x, void *pv) {
  const depth_par *p = (const depth_par *)pv;
  const double A = p->a, C = p->c, L = p->l;
  return (L / C) * (1.0 / sqrt(A * A - L * L)) * (C * x - log(A * nz(x) - L * L + sqrt(A * A - L * L) * sqrt(pow(nz(x), 2) - L * L)));
}
/* fDnfR_L, IceRayTracing.cc:368-379: the same closed form as a function of L at fixed depth Z */
static double f_L(double x, const lpar *p) {
  const double A = p->a, C = p->c, Z = p->z;
  return (x / C) * (1.0 / sqrt(A * A - x * x)) * (C * Z - log(A * nz(Z) - x * x + sqrt(A * A - x * x) * sqrt(pow(nz(Z), 2) - x * x)));
}
/* ftimeD, IceRayTracing.cc:382-393 */
static double f_time(double x, const depth_par *p) {
  const double A = p->a, C = p->c, L = p->l;
  return (1.0 / (k_c * C * sqrt(pow(nz(x), 2) - L * L))) *
         (pow(nz(x), 2) - L * L +
          (C * x - log(A * nz(x) - L * L + sqrt(A * A - L * L) * sqrt(pow(nz(x), 2) - L * L))) * (A * A * sqrt(pow(nz(x), 2) - L * L)) / sqrt(A * A - L * L) +
          A * sqrt(pow(nz(x), 2) - L * L) * log(nz(x) + sqrt(pow(nz(x), 2) - L * L)));
}
/* fpathD, IceRayTracing.cc:396-408 (the antiderivative of sec(asin(L/n)) written with exp(C x)) */
static double f_path(double x, const depth_par *p) {
  const double A = p->a, B = p->b, C = p->c, L = p->l;
  const double e = exp(C * x), e2 = exp(2 * C * x);
  const double q = sqrt((A * A + 2 * A * B * e + B * B * e2 - L * L) / ((A + B * e) * (A + B * e)));
  return (log((A + B * e) * (q + 1)) -
          (A * log(A * sqrt(A * A - L * L) * q + B * sqrt(A * A - L * L) * e * q + A * A + A * B * e - L * L)) / sqrt(A * A - L * L) +
          (A * C * x) / sqrt(A * A - L * L)) / C;
}

/* GetMinnz, IceRayTracing.cc:338-343 (raw x, not |x|) */
static double f_minnz(double x, void *pv) {
  const minnz_par *p = (const minnz_par *)pv;
  return p->a + g_B * exp(-g_C * x) - p->l;
}

/* FindFunctionRoot, IceRayTracing.cc:261-300: GSL falsepos, stop on |f(root)| < 1e-6, at most 100 iterations */
static double find_root(gsl_function F, double x_lo, double x_hi) {
  int status, iter = 0;
  const int max_iter = 100;
  gsl_root_fsolver *s = gsl_root_fsolver_alloc(gsl_root_fsolver_falsepos);
  double r = 0;
  gsl_set_error_handler_off();
  status = gsl_root_fsolver_set(s, &F, x_lo, x_hi);
  do {
    iter++;
    status = gsl_root_fsolver_iterate(s);
    r = gsl_root_fsolver_root(s);
    status = gsl_root_test_residual(F.function(r, F.params), 1e-6);
  } while (status == GSL_CONTINUE && iter < max_iter);
  gsl_root_fsolver_free(s);
  return r;
}
/* FindFunctionRootZmax, IceRayTracing.cc:303-335: the same, stopped on the bracket width (1e-6 abs, 1e-6 rel) */
static double find_root_zmax(gsl_function F, double x_lo, double x_hi) {
  int status, iter = 0;
  const int max_iter = 100;
  gsl_root_fsolver *s = gsl_root_fsolver_alloc(gsl_root_fsolver_falsepos);
  double r = 0;
  gsl_set_error_handler_off();
  status = gsl_root_fsolver_set(s, &F, x_lo, x_hi);
  do {
    iter++;
    status = gsl_root_fsolver_iterate(s);
    r = gsl_root_fsolver_root(s);
    x_lo = gsl_root_fsolver_x_lower(s);
    x_hi = gsl_root_fsolver_x_upper(s);
    status = gsl_root_test_interval(x_lo, x_hi, 1e-6, 1e-6);
  } while (status == GSL_CONTINUE && iter < max_iter);
  gsl_root_fsolver_free(s);
  return r;
}
/* FindFunctionRootFDF, IceRayTracing.cc:222-258: GSL newton from the bracket midpoint, stop on |dx| < 1e-6 |x| */
static double find_root_fdf(gsl_function_fdf FDF, double x_lo, double x_hi) {
  int status, iter = 0;
  const int max_iter = 100;
  gsl_root_fdfsolver *s = gsl_root_fdfsolver_alloc(gsl_root_fdfsolver_newton);
  double x0, x = (x_lo + x_hi) / 2;
  gsl_root_fdfsolver_set(s, &FDF, x);
  do {
    iter++;
    status = gsl_root_fdfsolver_iterate(s);
    x0 = x;
    x = gsl_root_fdfsolver_root(s);
    status = gsl_root_test_delta(x, x0, 0, 1e-6);
  } while (status == GSL_CONTINUE && iter < max_iter);
  gsl_root_fdfsolver_free(s);
  return x;
}

/* GetZmax, IceRayTracing.cc:346-353 */
static double get_zmax(double A, double L) {
  gsl_function F;
  minnz_par p = {A, L};
  F.function = &f_minnz;
  F.params = &p;
  return find_root_zmax(F, 0.0, 5000);
}

/* fDa, IceRayTracing.cc:411-449 (TransitionBoundary == 0 branch, :445) */
static double f_Da(double x, void *pv) {
  const pair_par *p = (const pair_par *)pv;
  lpar a = {p->a, g_B, g_C, p->z1}, b = {p->a, g_B, g_C, p->z0};
  return (f_L(x, &a) - f_L(x, &b)) - p->x1;
}
/* fRa, IceRayTracing.cc:466-520 (:515-516) */
static double f_Ra(double x, void *pv) {
  const pair_par *p = (const pair_par *)pv;
  lpar a = {p->a, g_B, -g_C, -p->z1}, b = {p->a, g_B, -g_C, -p->z0}, c = {p->a, g_B, -g_C, 1e-7};
  const double d01 = f_L(x, &a) - f_L(x, &b);
  const double d0s = f_L(x, &c) - f_L(x, &b);
  return d01 - 2 * (d0s) - p->x1;
}
/* fRaa, IceRayTracing.cc:537-607 (:592-604) */
static double f_Raa(double x, void *pv) {
  const pair_par *p = (const pair_par *)pv;
  const double zmax = get_zmax(p->a, x) + 1e-7;
  double out = 0;
  if (zmax > 0) {
    lpar a = {p->a, g_B, -g_C, -p->z1}, b = {p->a, g_B, -g_C, -p->z0}, c = {p->a, g_B, -g_C, zmax};
    double d01 = f_L(x, &a) - f_L(x, &b);
    double d0s = f_L(x, &c) - f_L(x, &b);
    if (isnan(d01)) d01 = 1e9;
    if (isnan(d0s)) d0s = 1e9;
    out = d01 - 2 * (d0s) - p->x1;
  } else {
    out = 1e9;
  }
  return out;
}
/* fRaa_df / fRaa_fdf, IceRayTracing.cc:609-621 */
static double f_Raa_df(double x, void *pv) {
  gsl_function F;
  double result, abserr;
  F.function = &f_Raa;
  F.params = pv;
  gsl_deriv_central(&F, x, 1e-8, &result, &abserr);
  return result;
}
static void f_Raa_fdf(double x, void *pv, double *y, double *dy) {
  *y = f_Raa(x, pv);
  *dy = f_Raa_df(x, pv);
}

static double min2(double a, double b) { return b < a ? b : a; }   /* std::min_element keeps the first of equals */

/* GetDirectRayPar, IceRayTracing.cc:626-742 -> out[6] = RangD, LangD, timeD, lvalueD, checkzeroD, pathD */
static void direct_ray(double z0, double x1, double z1, double *out) {
  int flip = 0;
  double dsw = z0;
  if (z0 > z1) { z0 = z1; z1 = dsw; flip = 1; }
  gsl_function F1;
  pair_par p1 = {g_A, z0, x1, z1};
  F1.function = &f_Da;
  F1.params = &p1;
  const double up = min2(nz(z1), nz(z0));
  const double lvalueD = find_root(F1, 1e-7, up);
  const double LangD = asin(lvalueD / nz(z0)) * (180.0 / k_pi);
  double checkD = f_Da(lvalueD, &p1);
  depth_par ta = {g_A, g_B, -g_C, lvalueD}, tb = {g_A, g_B, -g_C, lvalueD};
  const double timeD = f_time(-z0, &ta) - f_time(-z1, &tb);
  const double pathD = f_path(-z0, &ta) - f_path(-z1, &tb);
  gsl_function F5;
  depth_par p5 = {g_A, g_B, -g_C, lvalueD};
  double result, abserr;
  F5.function = &f_depth;
  F5.params = &p5;
  gsl_deriv_central(&F5, -z1, 1e-8, &result, &abserr);
  double RangD = atan(result) * (180.0 / k_pi);
  if (z1 == z0 && isnan(RangD)) RangD = 180 - LangD;
  if (isnan(checkD)) checkD = -1000;
  out[0] = RangD; out[1] = LangD; out[2] = timeD; out[3] = lvalueD; out[4] = checkD; out[5] = pathD;
  if (flip) { out[0] = 180 - LangD; out[1] = 180 - RangD; }
}

/* GetReflectedRayPar, IceRayTracing.cc:745-920 -> out[11] */
static void reflected_ray(double z0, double x1, double z1, double *out) {
  int flip = 0;
  double dsw = z0;
  if (z0 > z1) { z0 = z1; z1 = dsw; flip = 1; }
  gsl_function F3;
  pair_par p3 = {g_A, z0, x1, z1};
  F3.function = &f_Ra;
  F3.params = &p3;
  const double up = min2(min2(nz(z1), nz(z0)), nz(1e-7));
  const double lvalueR = find_root(F3, 1e-7, up);
  const double LangR = asin(lvalueR / nz(z0)) * (180.0 / k_pi);
  double checkR = f_Ra(lvalueR, &p3);
  depth_par pa = {g_A, g_B, g_C, lvalueR}, pb = {g_A, g_B, g_C, lvalueR}, pc = {g_A, g_B, g_C, lvalueR};
  double timeR1 = f_time(-1e-7, &pc) - f_time(z0, &pa);
  double timeR2 = f_time(-1e-7, &pc) - f_time(z1, &pb);
  double pathR1 = f_path(-1e-7, &pc) - f_path(z0, &pa);
  double pathR2 = f_path(-1e-7, &pc) - f_path(z1, &pb);
  const double timeR = timeR1 + timeR2, pathR = pathR1 + pathR2;
  if (flip) {
    double d = timeR2; timeR2 = timeR1; timeR1 = d;
    d = pathR2; pathR2 = pathR1; pathR1 = d;
  }
  gsl_function F5;
  depth_par p5 = {g_A, g_B, g_C, lvalueR};
  double result, abserr;
  F5.function = &f_depth;
  F5.params = &p5;
  gsl_deriv_central(&F5, z1, 1e-8, &result, &abserr);
  double RangR = 180 - atan(result) * (180.0 / k_pi);
  if (z1 == z0 && isnan(RangR)) RangR = 180 - LangR;
  if (z1 != z0 && isnan(RangR)) RangR = 90;
  if (isnan(checkR)) checkR = -1000;
  gsl_deriv_central(&F5, -1e-7, 1e-8, &result, &abserr);     /* incidence angle on the surface, :889-894 */
  const double inc = atan(result) * (180.0 / k_pi);
  out[0] = RangR; out[1] = LangR; out[2] = timeR; out[3] = lvalueR; out[4] = checkR; out[5] = timeR1; out[6] = timeR2;
  out[7] = inc; out[8] = pathR; out[9] = pathR1; out[10] = pathR2;
  if (flip) { out[0] = 180 - LangR; out[1] = 180 - RangR; }
}

/* GetRefractedRayPar, IceRayTracing.cc:923-1253 -> out[22] */
static void refracted_ray(double z0, double x1, double z1, double LangR, double RangR, double checkD, double checkR, double *out) {
  int flip = 0, i;
  double dsw = z0;
  if (z0 > z1) { z0 = z1; z1 = dsw; flip = 1; }
  if (flip) { dsw = 180 - LangR; LangR = 180 - RangR; RangR = dsw; }
  double lv[2] = {0, 0}, La[2] = {0, 0}, cz[2] = {-1000, -1000}, tRa[2] = {0, 0}, tRa1[2] = {0, 0}, tRa2[2] = {0, 0};
  double pRa[2] = {0, 0}, pRa1[2] = {0, 0}, pRa2[2] = {0, 0}, raytime[2] = {0, 0}, Ra[2] = {0, 0}, zm[2] = {10, 10};
  const double up = min2(nz(z0), nz(z1));
  gsl_function F4;
  pair_par p4 = {g_A, z0, x1, z1};
  F4.function = &f_Raa;
  F4.params = &p4;
  gsl_function_fdf F4b;
  F4b.f = &f_Raa; F4b.df = &f_Raa_df; F4b.fdf = &f_Raa_fdf; F4b.params = &p4;
  double lower = nz(z0) * sin((64.0 * (k_pi / 180.0)));
  if (lower > up) lower = nz(z0) * sin((LangR * (k_pi / 180.0)));
#define ORACLE_SECOND(root_expr)                          \
  do {                                                    \
    lv[1] = (root_expr);                                  \
    La[1] = asin(lv[1] / nz(z0)) * (180.0 / k_pi);        \
    cz[1] = f_Raa(lv[1], &p4);                            \
    zm[1] = get_zmax(g_A, lv[1]) + 1e-7;                  \
  } while (0)
#define ORACLE_BAD (fabs(cz[1]) > 0.5 || isnan(cz[1]) || fabs(lv[1] - lv[0]) < 1e-4)
  lv[0] = find_root(F4, lower, up);
  La[0] = asin(lv[0] / nz(z0)) * (180.0 / k_pi);
  cz[0] = f_Raa(lv[0], &p4);
  zm[0] = get_zmax(g_A, lv[0]) + 1e-7;
  if (fabs(cz[0]) > 0.5) {
    lv[0] = find_root_fdf(F4b, lower, up);
    La[0] = asin(lv[0] / nz(z0)) * (180.0 / k_pi);
    cz[0] = f_Raa(lv[0], &p4);
    zm[0] = get_zmax(g_A, lv[0]) + 1e-7;
  }
  if (lv[0] < 0) cz[0] = -1000;
  if (fabs(cz[0]) < 0.5 && fabs(checkD) > 0.5 && fabs(checkR) > 0.5) {
    ORACLE_SECOND(find_root(F4, lv[0] - 0.23, lv[0] - 0.023));
    if (ORACLE_BAD) ORACLE_SECOND(find_root(F4, lv[0] - 0.15, lv[0] - 0.023));
    if (ORACLE_BAD) {
      if (lv[0] + 0.005 < up) ORACLE_SECOND(find_root(F4, lv[0] + 0.005, up));
      else ORACLE_SECOND(find_root(F4, lv[0] - 0.1, lv[0] - 0.01));
    }
    if (ORACLE_BAD) {
      const double tmp = find_root_fdf(F4b, lv[0] - 0.23, lv[0] - 0.023);
      if (fabs(tmp) < g_A) ORACLE_SECOND(find_root_fdf(F4b, lv[0] - 0.23, lv[0] - 0.023));   /* solved twice, :1029-1031 */
    }
    if (ORACLE_BAD) {
      const double tmp = find_root_fdf(F4b, lv[0] - 0.1, lv[0] - 0.023);
      if (fabs(tmp) < g_A) ORACLE_SECOND(tmp);
    }
    if (lv[1] < 0) cz[1] = -1000;
    if (fabs(cz[1]) < 0.5 && fabs(cz[0]) < 0.5 && fabs(lv[1] - lv[0]) < 1e-4) cz[1] = -1000;
    if (isnan(La[0])) La[0] = 0;
    if (isnan(La[1])) La[1] = 0;
    if (La[1] < La[0] && fabs(cz[0]) < 0.5 && fabs(cz[1]) < 0.5) {
      double t;
      t = lv[1]; lv[1] = lv[0]; lv[0] = t;
      t = La[1]; La[1] = La[0]; La[0] = t;
      t = cz[1]; cz[1] = cz[0]; cz[0] = t;
      t = zm[1]; zm[1] = zm[0]; zm[0] = t;
    }
  } else {
    lv[1] = 0; La[1] = 0; cz[1] = -1000; zm[1] = -1000;
  }
#undef ORACLE_SECOND
#undef ORACLE_BAD
  for (i = 0; i < 2; i++) {
    if (isnan(cz[i])) cz[i] = -1000;
    if (zm[i] == 1e-7 || zm[i] <= 0) cz[i] = -1000;
    depth_par pa = {g_A, g_B, g_C, lv[i]}, pb = {g_A, g_B, g_C, lv[i]}, pc = {g_A, g_B, g_C, lv[i]};
    if ((z0 < -zm[i] || zm[i] < -z1)) {
      tRa1[i] = f_time(-zm[i], &pc) - f_time(z0, &pa);
      tRa2[i] = f_time(-zm[i], &pc) - f_time(z1, &pb);
      pRa1[i] = f_path(-zm[i], &pc) - f_path(z0, &pa);
      pRa2[i] = f_path(-zm[i], &pc) - f_path(z1, &pb);
      raytime[i] = tRa1[i] + tRa2[i];
      pRa[i] = pRa1[i] + pRa2[i];
      if (flip) {
        double d = tRa2[i]; tRa2[i] = tRa1[i]; tRa1[i] = d;
        d = pRa2[i]; pRa2[i] = pRa1[i]; pRa1[i] = d;
      }
    }
    tRa[i] = raytime[i];
    gsl_function F5;
    depth_par p5 = {g_A, g_B, g_C, lv[i]};
    double result, abserr;
    F5.function = &f_depth;
    F5.params = &p5;
    gsl_deriv_central(&F5, z1, 1e-8, &result, &abserr);
    Ra[i] = 180 - atan(result) * (180.0 / k_pi);
    if (z1 == z0 && isnan(Ra[i])) Ra[i] = 180 - La[i];
    if (z1 != z0 && isnan(Ra[i])) Ra[i] = 90;
  }
  if (isnan(cz[0])) cz[0] = -1000;
  if (isnan(cz[1])) cz[1] = -1000;
  out[0] = Ra[0]; out[1] = La[0]; out[2] = tRa[0]; out[3] = lv[0]; out[4] = cz[0]; out[5] = tRa1[0]; out[6] = tRa2[0]; out[7] = zm[0];
  if (flip) { out[0] = 180 - La[0]; out[1] = 180 - Ra[0]; }
  out[8] = Ra[1]; out[9] = La[1]; out[10] = tRa[1]; out[11] = lv[1]; out[12] = cz[1]; out[13] = tRa1[1]; out[14] = tRa2[1]; out[15] = zm[1];
  out[16] = pRa[0]; out[17] = pRa1[0]; out[18] = pRa2[0]; out[19] = pRa[1]; out[20] = pRa1[1]; out[21] = pRa2[1];
  if (flip) { out[8] = 180 - La[1]; out[9] = 180 - Ra[1]; }
}

/* IceRayTracing::IceRayTracing(x0, z0, x1, z1), IceRayTracing.cc:1745-1919 -> out[29].
 * Slots 12..17 are assigned only when their branch exists (:1872-1883; the reference leaves them uninitialised
 * otherwise); here they are 0 then. */
void inice_oracle_solve(double z0, double x1, double z1, double *out) {
  double d[6], r[11], a[22];
  int k;
  direct_ray(z0, x1, z1, d);
  reflected_ray(z0, x1, z1, r);
  const double checkD = d[4], checkR = r[4];
  double RangRa[2] = {0, 0}, LangRa[2] = {0, 0}, timeRa[2] = {0, 0}, lvalueRa[2] = {0, 0}, checkRa[2] = {-1000, -1000};
  double timeRa1[2] = {0, 0}, timeRa2[2] = {0, 0}, zmax[2] = {0, 0}, pathRa[2] = {0, 0};
  if (fabs(checkR) > 0.5 || fabs(checkD) > 0.5) {
    refracted_ray(z0, x1, z1, r[1], r[0], checkD, checkR, a);
    RangRa[0] = a[0]; LangRa[0] = a[1]; timeRa[0] = a[2]; lvalueRa[0] = a[3]; checkRa[0] = a[4];
    timeRa1[0] = a[5]; timeRa2[0] = a[6]; zmax[0] = a[7];
    if (fabs(checkR) > 0.5 && fabs(checkD) > 0.5) {
      RangRa[1] = a[8]; LangRa[1] = a[9]; timeRa[1] = a[10]; lvalueRa[1] = a[11]; checkRa[1] = a[12];
      timeRa1[1] = a[13]; timeRa2[1] = a[14]; zmax[1] = a[15];
    }
    pathRa[0] = a[16]; pathRa[1] = a[19];
  }
  for (k = 0; k < 29; k++) out[k] = 0;
  out[0] = d[1]; out[1] = r[1]; out[2] = LangRa[0]; out[3] = LangRa[1];
  out[4] = d[2]; out[5] = r[2]; out[6] = timeRa[0]; out[7] = timeRa[1];
  out[8] = d[0]; out[9] = r[0]; out[10] = RangRa[0]; out[11] = RangRa[1];
  if (fabs(checkR) < 0.5) { out[12] = r[5]; out[13] = r[6]; }
  if (fabs(checkRa[0]) < 0.5) { out[14] = timeRa1[0]; out[15] = timeRa2[0]; }
  if (fabs(checkRa[1]) < 0.5) { out[16] = timeRa1[1]; out[17] = timeRa2[1]; }
  out[18] = r[7];
  out[19] = d[3]; out[20] = r[3]; out[21] = lvalueRa[0]; out[22] = lvalueRa[1];
  out[23] = zmax[0]; out[24] = zmax[1];
  out[25] = d[5]; out[26] = r[8]; out[27] = pathRa[0]; out[28] = pathRa[1];
  if (fabs(checkD) > 0.5) out[8] = -1000;
  if (fabs(checkR) > 0.5) out[9] = -1000;
  if (fabs(checkRa[0]) > 0.5) out[10] = -1000;
  if (fabs(checkRa[1]) > 0.5) out[11] = -1000;
}

void inice_oracle_solve_batch(long n, const double *z0, const double *x1, const double *z1, double *out) {
  long i;
  for (i = 0; i < n; i++) inice_oracle_solve(z0[i], x1[i], z1[i], out + 29 * i);
}

/* GetRayTracingSolutions, IceRayTracing.cc:2907-3210, without the attenuation integrals (:2952-2966; AttRay is not
 * produced): the two physical rays of a pair.  out10 = TimeRay[2], PathRay[2], LaunchAngle[2], RecieveAngle[2],
 * IncidenceAngleInIce[2]; ignore2 = IgnoreCh[2]. */
void inice_oracle_two_rays(double RxDepth, double Distance, double TxDepth, double *out10, int *ignore2) {
  double r[29];
  inice_oracle_solve(TxDepth, Distance, RxDepth, r);
  const double timeD = r[4], timeR = r[5], timeRa[2] = {r[6], r[7]};
  const double pathD = r[25], pathR = r[26], pathRa[2] = {r[27], r[28]};
  const double RangD = r[8], RangR = r[9], RangRa[2] = {r[10], r[11]};
  const double LangD = r[0], LangR = r[1], LangRa[2] = {r[2], r[3]};
  double TimeRay[2], PathRay[2], RecieveAngle[2], LaunchAngle[2], Inc[2];
  int IgnoreCh[2];
  TimeRay[0] = timeD; TimeRay[1] = timeR; PathRay[0] = pathD; PathRay[1] = pathR;
  RecieveAngle[0] = RangD; RecieveAngle[1] = RangR; LaunchAngle[0] = LangD; LaunchAngle[1] = LangR;
  Inc[0] = 100; Inc[1] = r[18];
  if (RangR == -1000) { Inc[0] = 100; Inc[1] = 100; }
#define USE_D(s)     do { TimeRay[s] = timeD; PathRay[s] = pathD; RecieveAngle[s] = RangD; LaunchAngle[s] = LangD; } while (0)
#define USE_R(s)     do { TimeRay[s] = timeR; PathRay[s] = pathR; RecieveAngle[s] = RangR; LaunchAngle[s] = LangR; } while (0)
#define USE_RA(s, i) do { TimeRay[s] = timeRa[i]; PathRay[s] = pathRa[i]; RecieveAngle[s] = RangRa[i]; LaunchAngle[s] = LangRa[i]; } while (0)
  if (RangD != -1000) USE_D(0);                                            /* :2984-2991 */
  if (RangR != -1000) USE_R(1);                                            /* :2993-3000 */
  if (RangRa[0] != -1000 && RangD != -1000) { USE_D(0); USE_RA(1, 0); }    /* :3002-3016 */
  if (RangRa[0] != -1000 && RangR != -1000) { USE_R(1); USE_RA(0, 0); }    /* :3018-3032 */
  if (RangRa[1] != -1000 && RangD != -1000) { USE_D(0); USE_RA(1, 1); }    /* :3034-3048 */
  if (RangRa[1] != -1000 && RangR != -1000) { USE_R(1); USE_RA(0, 1); }    /* :3050-3064 */
  if (RangRa[1] != -1000 && RangRa[0] != -1000) { USE_RA(1, 1); USE_RA(0, 0); }   /* :3066-3080 */
  if (RecieveAngle[1] == -1000 && RecieveAngle[0] == -1000 && RangRa[0] != -1000) USE_RA(0, 0);
  if (RecieveAngle[1] == -1000 && RecieveAngle[0] == -1000 && RangRa[1] != -1000) USE_RA(1, 1);
#undef USE_D
#undef USE_R
#undef USE_RA
  IgnoreCh[0] = 1; IgnoreCh[1] = 1;
  if (RecieveAngle[0] == -1000) IgnoreCh[0] = 0;
  if (RecieveAngle[1] == -1000) IgnoreCh[1] = 0;
  if (TimeRay[0] > TimeRay[1] && RecieveAngle[0] != -1000 && RecieveAngle[1] != -1000) {   /* :3141-3148 */
    double t;
    t = LaunchAngle[0]; LaunchAngle[0] = LaunchAngle[1]; LaunchAngle[1] = t;
    t = RecieveAngle[0]; RecieveAngle[0] = RecieveAngle[1]; RecieveAngle[1] = t;
    t = TimeRay[0]; TimeRay[0] = TimeRay[1]; TimeRay[1] = t;
    t = PathRay[0]; PathRay[0] = PathRay[1]; PathRay[1] = t;
  }
  if (RxDepth == TxDepth && TimeRay[0] == 0 && PathRay[0] == 0) {          /* :3190-3200 */
    if (Distance == 0) { IgnoreCh[0] = 0; IgnoreCh[1] = 0; }
    PathRay[0] = Distance;
    TimeRay[0] = Distance / (k_c / nz(TxDepth));
    LaunchAngle[0] = 90.; RecieveAngle[0] = 90.;
    IgnoreCh[0] = 1;
  }
  out10[0] = TimeRay[0]; out10[1] = TimeRay[1]; out10[2] = PathRay[0]; out10[3] = PathRay[1];
  out10[4] = LaunchAngle[0]; out10[5] = LaunchAngle[1]; out10[6] = RecieveAngle[0]; out10[7] = RecieveAngle[1];
  out10[8] = Inc[0]; out10[9] = Inc[1];
  ignore2[0] = IgnoreCh[0]; ignore2[1] = IgnoreCh[1];
}
void inice_oracle_two_rays_batch(long n, const double *rx, const double *dist, const double *tx, double *out10, int *ignore2) {
  long i;
  for (i = 0; i < n; i++) inice_oracle_two_rays(rx[i], dist[i], tx[i], out10 + 10 * i, ignore2 + 2 * i);
}
