/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's air->ice hot path.
 * See airice_oracle.h for the rules on who may use this file and for the citation abbreviations.
 *
 * The aim is bit-for-bit agreement with the reference's own build (x86-64, no FMA contraction,
 * glibc libm), so every floating-point expression below keeps the operand order and association of
 * the reference expression it cites; only the program structure (no heap traffic, no per-call layer
 * scans repeated three times, no std::vector) is ours.  Compile with -ffp-contract=off.
 */
#include "airice_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ tiny istream emulation
 * The reference parses Atmosphere.dat with the "while(getline){ file >> a >> b; }" idiom
 * (M.cc:36-57, 97-121) whose side effects (the last row is seen twice, M.cc:137) decide how many
 * knots the spline gets, so the stream semantics are emulated rather than re-imagined. */
typedef struct { const char *buf; long pos, len; int fail; } istream_t;

static int is_getline(istream_t *s) {
  if (s->fail) return 0;
  if (s->pos >= s->len) { s->fail = 1; return 0; }
  while (s->pos < s->len && s->buf[s->pos] != '\n') s->pos++;
  if (s->pos < s->len) s->pos++; /* consume '\n' */
  return 1;
}
static void is_ignore_line(istream_t *s, long maxn) {
  long k = 0;
  if (s->fail) return;
  while (s->pos < s->len && k < maxn) {
    char ch = s->buf[s->pos++];
    k++;
    if (ch == '\n') return;
  }
}
static int is_read_double(istream_t *s, double *v) {
  char *end;
  double x;
  if (s->fail) return 0;
  while (s->pos < s->len && (s->buf[s->pos] == ' ' || s->buf[s->pos] == '\t' || s->buf[s->pos] == '\n' ||
                             s->buf[s->pos] == '\r' || s->buf[s->pos] == '\v' || s->buf[s->pos] == '\f'))
    s->pos++;
  if (s->pos >= s->len) { s->fail = 1; return 0; }
  x = strtod(s->buf + s->pos, &end);
  if (end == s->buf + s->pos) { s->fail = 1; return 0; }
  s->pos = end - s->buf;
  *v = x;
  return 1;
}

/* natural cubic spline through (x,y)[0..n), value at xq: GSL interpolation/cspline.c +
 * linalg/tridiag.c (L.D.L^T), the algorithm behind gsl_spline_eval at M.cc:203. */
static double natural_spline_at(const double *x, const double *y, long n, double xq) {
  const long max_index = n - 1, sys = max_index - 1;
  double *c = (double *)calloc((size_t)n, sizeof(double));
  double *g = (double *)malloc((size_t)n * sizeof(double));
  double *diag = (double *)malloc((size_t)n * sizeof(double));
  double *off = (double *)malloc((size_t)n * sizeof(double));
  double *gam = (double *)malloc((size_t)n * sizeof(double));
  double *alp = (double *)malloc((size_t)n * sizeof(double));
  double *z = (double *)malloc((size_t)n * sizeof(double));
  double result;
  long i, lo, hi;
  for (i = 0; i < sys; i++) {
    const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    const double yd_i = y[i + 1] - y[i], yd_ip1 = y[i + 2] - y[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  if (sys == 1) {
    c[1] = g[0] / diag[0];
  } else if (sys > 1) {
    double *sol = c + 1;
    alp[0] = diag[0];
    gam[0] = off[0] / alp[0];
    for (i = 1; i < sys - 1; i++) {
      alp[i] = diag[i] - off[i - 1] * gam[i - 1];
      gam[i] = off[i] / alp[i];
    }
    alp[sys - 1] = diag[sys - 1] - off[sys - 2] * gam[sys - 2];
    z[0] = g[0];
    for (i = 1; i < sys; i++) z[i] = g[i] - gam[i - 1] * z[i - 1];
    for (i = 0; i < sys; i++) z[i] = z[i] / alp[i];
    sol[sys - 1] = z[sys - 1];
    for (i = sys - 2; i >= 0; i--) sol[i] = z[i] - gam[i] * sol[i + 1];
  }
  lo = 0; hi = n - 1;
  while (hi > lo + 1) {
    long mid = (hi + lo) / 2;
    if (x[mid] > xq) hi = mid; else lo = mid;
  }
  {
    const double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo], delx = xq - x[lo];
    const double b_i = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
    const double d_i = (c[lo + 1] - c[lo]) / (3.0 * dx);
    result = y[lo] + delx * (b_i + delx * (c[lo] + delx * d_i));
  }
  free(c); free(g); free(diag); free(off); free(gam); free(alp); free(z);
  return result;
}

/* readATMpar (M.cc:24-71) + readnhFromFile (M.cc:73-147) + MakeAtmosphere (M.cc:920-942) +
 * FillInAirRefractiveIndex (M.cc:193-213).  P.cc:4-118,148-165,860-882 is the same code. */
int oracle_atm_load(const char *path, int variant, oracle_atm *atm) {
  FILE *fp = fopen(path, "rb");
  char *buf;
  long len, cap, i;
  istream_t s;
  double d5[5] = {0, 0, 0, 0, 0};
  int n1 = 0, layer = 0, nvec = 0, pending = 0;
  double *hx, *ny;
  long np = 0;
  double dummy1 = 0, dummy2 = 0;
  if (!fp) return -1;
  fseek(fp, 0, SEEK_END);
  len = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  buf = (char *)malloc((size_t)len + 1);
  if (fread(buf, 1, (size_t)len, fp) != (size_t)len) { fclose(fp); free(buf); return -2; }
  buf[len] = 0;
  fclose(fp);

  memset(atm, 0, sizeof(*atm));
  atm->variant = variant;
  atm->pi = (variant == ORACLE_VARIANT_PYWRAP) ? 4.0 * atan(1.0) : 3.1415927;
  atm->c = 299792458.0;
  atm->A_air = 1.00;
  atm->A_ice = 1.78; atm->B_ice = -0.43; atm->C_ice = 0.0132;

  /* pass 1: ATMLAY and a,b,c (M.cc:36-57) */
  s.buf = buf; s.pos = 0; s.len = len; s.fail = 0;
  while (is_getline(&s)) {
    if (n1 < 4) { for (i = 0; i < 5; i++) is_read_double(&s, &d5[i]); }
    if (n1 == 0) for (i = 0; i < 5; i++) atm->atmlay_cm[i] = d5[i];
    if (n1 == 1) for (i = 0; i < 5; i++) atm->abc[i][0] = d5[i];
    if (n1 == 2) for (i = 0; i < 5; i++) atm->abc[i][1] = d5[i];
    if (n1 == 3) for (i = 0; i < 5; i++) atm->abc[i][2] = d5[i];
    n1++;
  }
  for (i = 0; i < 3; i++) atm->abc[4][i] = atm->abc[3][i]; /* M.cc:62-64 */
  atm->atmlay_cm[4] = 150000 * 100;                        /* M.cc:66 */

  /* pass 2: tabulated n(h) (M.cc:88-140).  Only the flattened knot list and the number of
   * per-layer vectors matter downstream. */
  cap = 1024; hx = (double *)malloc((size_t)cap * sizeof(double)); ny = (double *)malloc((size_t)cap * sizeof(double));
  s.pos = 0; s.fail = 0;
  for (i = 0; i < 5; i++) is_ignore_line(&s, 256);
  while (is_getline(&s)) {
    is_read_double(&s, &dummy1);
    is_read_double(&s, &dummy2);
    if (dummy1 > -1) {
      if (np == cap) { cap *= 2; hx = (double *)realloc(hx, (size_t)cap * sizeof(double)); ny = (double *)realloc(ny, (size_t)cap * sizeof(double)); }
      hx[np] = dummy1; ny[np] = dummy2; np++; pending++;
      if (layer < 5 && dummy1 * 100 >= atm->atmlay_cm[layer]) {
        if (layer > 0) { nvec++; pending = 0; }
        layer++;
      }
    }
  }
  if (layer > 0) { nvec++; }
  (void)pending;
  np -= 1; /* M.cc:137-140: the idiom reads the last row twice; the duplicate is erased */
  atm->max_layers = nvec + 1; /* M.cc:142 */
  atm->npoints = (int)np;
  atm->n0 = natural_spline_at(hx, ny, np, 0.0);
  free(hx); free(ny); free(buf);

  { /* M.cc:193-213 */
    double N0 = 0;
    int il;
    for (il = 0; il < 5; il++) {
      double hlow = atm->atmlay_cm[il] / 100;
      atm->C_air[il] = 1.0 / (atm->abc[il][2] / 100);
      if (il > 0) N0 = atm->A_air + atm->B_air[il - 1] * exp(-hlow * atm->C_air[il - 1]);
      if (il == 0) N0 = atm->n0;
      atm->B_air[il] = ((N0 - 1) / exp(-hlow * atm->C_air[il]));
    }
  }
  return 0;
}

/* GetB_air/GetC_air layer scan (M.cc:216-256) */
int oracle_layer_of(const oracle_atm *a, double z) {
  double zabs = fabs(z);
  int which = 0, il;
  for (il = 0; il < a->max_layers - 1; il++) {
    if (zabs < a->atmlay_cm[il + 1] / 100 && zabs >= a->atmlay_cm[il] / 100) { which = il; break; }
  }
  if (zabs >= a->atmlay_cm[a->max_layers - 1] / 100) which = a->max_layers - 1;
  return which;
}
/* Getnz_air (M.cc:259-263) */
double oracle_nz_air(const oracle_atm *a, double z) {
  double zabs = fabs(z);
  int k = oracle_layer_of(a, zabs);
  return a->A_air + a->B_air[k] * exp(-a->C_air[k] * zabs);
}
/* Getnz_ice (M.cc:188-191), TransitionBoundary==0 (M.h:67) */
double oracle_nz_ice(const oracle_atm *a, double z) {
  z = fabs(z);
  return a->A_ice + a->B_ice * exp(-a->C_ice * z);
}

typedef struct { double A, B, C, L; int air; } medium_t; /* C already negated as at M.cc:455-461 */

static medium_t medium_at(const oracle_atm *a, double A, double z, double L, int air) {
  medium_t m;
  m.A = A; m.L = L; m.air = air;
  if (air) { int k = oracle_layer_of(a, z); m.B = a->B_air[k]; m.C = -a->C_air[k]; }
  else { m.B = a->B_ice; m.C = -a->C_ice; }
  return m;
}
static double nz_medium(const oracle_atm *a, double x, int air) { return air ? oracle_nz_air(a, x) : oracle_nz_ice(a, x); }

/* fDnfR (M.cc:377-386) */
static double fDnfR(double x, const medium_t *p) {
  double A = p->A, B = p->B, C = p->C, L = p->L;
  return (L / C) * (1.0 / sqrt(A * A - L * L)) *
         (C * x - log(A * (A + B * exp(C * x)) - L * L + sqrt(A * A - L * L) * sqrt(pow(A + B * exp(C * x), 2) - L * L)));
}
/* ftimeD (M.cc:412-431) */
static double ftimeD(const oracle_atm *a, double x, const medium_t *p) {
  double A = p->A, C = p->C, L = p->L, Speedc = a->c;
  double n = nz_medium(a, x, p->air);
  return (1.0 / (Speedc * C * sqrt(pow(n, 2) - L * L))) *
         (pow(n, 2) - L * L +
          (C * x - log(A * n - L * L + sqrt(A * A - L * L) * sqrt(pow(n, 2) - L * L))) * (A * A * sqrt(pow(n, 2) - L * L)) /
              sqrt(A * A - L * L) +
          A * sqrt(pow(n, 2) - L * L) * log(n + sqrt(pow(n, 2) - L * L)));
}
/* fpathD (M.cc:434-447) */
static double fpathD(double x, const medium_t *p) {
  double A = p->A, B = p->B, C = p->C, L = p->L;
  return (log((A + B * exp(C * x)) *
              (sqrt((A * A + 2 * A * B * exp(C * x) + B * B * exp(2 * C * x) - L * L) / ((A + B * exp(C * x)) * (A + B * exp(C * x)))) + 1)) -
          (A * log(A * sqrt(A * A - L * L) *
                       sqrt((A * A + 2 * A * B * exp(C * x) + B * B * exp(2 * C * x) - L * L) / ((A + B * exp(C * x)) * (A + B * exp(C * x)))) +
                   B * sqrt(A * A - L * L) * exp(C * x) *
                       sqrt((A * A + 2 * A * B * exp(C * x) + B * B * exp(2 * C * x) - L * L) / ((A + B * exp(C * x)) * (A + B * exp(C * x)))) +
                   A * A + A * B * exp(C * x) - L * L)) /
              sqrt(A * A - L * L) +
          (A * C * x) / sqrt(A * A - L * L)) /
         C;
}

/* GetRayHorizontalPath / GetRayPropagationTime / GetRayGeometricPath (M.cc:449-513) */
static double seg_x(const oracle_atm *a, double A, double rx, double tx, double L, int air) {
  medium_t pa = medium_at(a, A, rx, L, air), pb = medium_at(a, A, tx, L, air);
  double v = +fDnfR(rx, &pa) - fDnfR(tx, &pb);
  if (air) v *= -1;
  return v;
}
static double seg_t(const oracle_atm *a, double A, double rx, double tx, double L, int air) {
  medium_t pa = medium_at(a, A, rx, L, air), pb = medium_at(a, A, tx, L, air);
  double v = +ftimeD(a, rx, &pa) - ftimeD(a, tx, &pb);
  if (air) v *= -1;
  return v;
}
static double seg_p(const oracle_atm *a, double A, double rx, double tx, double L, int air) {
  medium_t pa = medium_at(a, A, rx, L, air), pb = medium_at(a, A, tx, L, air);
  double v = fpathD(rx, &pa) - fpathD(tx, &pb);
  if (air) v *= -1;
  return v;
}

/* GetLayerHitPointPar (M.cc:521-646): out = {X, recv deg, L, t, path} */
static void layer_hit(const oracle_atm *a, double n_layer1, double rx, double tx, double inc_deg, int air, double *out) {
  double inc = inc_deg * (a->pi / 180.0);
  double A = air ? a->A_air : a->A_ice;
  double nzRx = nz_medium(a, rx, air), nzTx = nz_medium(a, tx, air);
  double Lang = asin((n_layer1 / nzTx) * sin(inc));
  double recv = asin((nz_medium(a, tx, air) * sin(Lang)) / nz_medium(a, rx, air));
  double L = nzRx * sin(recv);
  out[0] = seg_x(a, A, rx, tx, L, air);
  out[1] = recv * (180 / a->pi);
  out[2] = L;
  out[3] = seg_t(a, A, rx, tx, L, air);
  out[4] = seg_p(a, A, rx, tx, L, air);
}

/* SkipLayersAbove / SkipLayersBelow scans (M.cc:666-690, 1799-1825).  The reference indexes
 * ATMLAY[max_layers] and ATMLAY[-1]; with max_layers==4 the former is the 150 km cap, the latter is
 * only reached when the height is in no layer at all (h<0 or h>=150 km), outside the domain. */
static double atmlay_m(const oracle_atm *a, int i) {
  if (i < 0) return -1e300;
  if (i > 4) return -1e300; /* ATMLAY[5] aliases unrelated storage in the reference (UB) */
  return a->atmlay_cm[i] / 100;
}
static void skip_layers(const oracle_atm *a, double h, double ice, int *above, int *below) {
  int skip = 0, il;
  for (il = a->max_layers; il > -1; il--) {
    if (h < atmlay_m(a, il) && h >= atmlay_m(a, il - 1)) il = -100;
    if (il > -1) skip++;
  }
  *above = skip;
  skip = 0;
  for (il = 0; il < a->max_layers; il++) {
    if (ice >= atmlay_m(a, il) && ice < atmlay_m(a, il + 1)) il = 100;
    if (il < a->max_layers) skip++;
  }
  *below = skip;
}

/* GetAirPropagationPar (M.cc:661-804) */
void oracle_air_walk(const oracle_atm *a, double theta, double h, double ice, double *out) {
  int above, below, il, cnt = 0, top;
  double start_angle = 0, start_h = 0, stop_h = 0, start_n = 0, L0 = 0;
  skip_layers(a, h, ice, &above, &below);
  top = a->max_layers - above - 1;
  for (il = top; il > below - 1; il--) {
    start_h = (il == top) ? h : a->atmlay_cm[il + 1] / 100 - 0.00001;
    start_n = oracle_nz_air(a, start_h);
    stop_h = (il == (below - 1) + 1) ? ice : a->atmlay_cm[il] / 100;
    if (il == top) {
      double hp[5];
      start_angle = 180 - theta;
      layer_hit(a, start_n, stop_h, start_h, start_angle, 1, hp);
      memcpy(out + 5 * cnt, hp, sizeof(hp));
      L0 = hp[2];
      start_angle = hp[1];
    } else {
      double nstop = oracle_nz_air(a, stop_h);
      double rec = asin(L0 / nstop);
      rec = rec * (180 / a->pi);
      out[5 * cnt + 0] = seg_x(a, a->A_air, stop_h, start_h, L0, 1);
      out[5 * cnt + 1] = rec;
      out[5 * cnt + 2] = L0;
      out[5 * cnt + 3] = seg_t(a, a->A_air, stop_h, start_h, L0, 1);
      out[5 * cnt + 4] = seg_p(a, a->A_air, stop_h, start_h, L0, 1);
      start_angle = rec;
    }
    cnt++;
  }
  (void)start_angle;
  out[5 * a->max_layers + 1] = cnt;
}

/* GetIcePropagationPar (M.cc:807-869), TransitionBoundary==0 branch: out = {X, recv deg, L, t, path} */
static void ice_leg(const oracle_atm *a, double depth_pos, double L, double *out) {
  double nstop = oracle_nz_ice(a, depth_pos);
  out[0] = seg_x(a, a->A_ice, depth_pos, 0.0, L, 0);
  out[1] = asin(L / nstop) * (180 / a->pi);
  out[2] = L;
  out[3] = seg_t(a, a->A_ice, depth_pos, 0.0, L, 0);
  out[4] = seg_p(a, a->A_ice, depth_pos, 0.0, L, 0);
}

/* Only the horizontal distance of the walk: what MinimizeforLaunchAngle actually consumes.  The
 * reference computes time and path too (M.cc:763-768) and discards them; skipping them changes no
 * bit of the result. */
static double air_x_only(const oracle_atm *a, double theta, double h, double ice, double *L_out) {
  int above, below, il, top;
  double X = 0, L0 = 0;
  skip_layers(a, h, ice, &above, &below);
  top = a->max_layers - above - 1;
  for (il = top; il > below - 1; il--) {
    double start_h = (il == top) ? h : a->atmlay_cm[il + 1] / 100 - 0.00001;
    double stop_h = (il == (below - 1) + 1) ? ice : a->atmlay_cm[il] / 100;
    if (il == top) {
      double start_n = oracle_nz_air(a, start_h);
      double inc = (180 - theta) * (a->pi / 180.0);
      double nzRx = oracle_nz_air(a, stop_h), nzTx = oracle_nz_air(a, start_h);
      double Lang = asin((start_n / nzTx) * sin(inc));
      double recv = asin((nzTx * sin(Lang)) / nzRx);
      L0 = nzRx * sin(recv);
    }
    X += seg_x(a, a->A_air, stop_h, start_h, L0, 1);
  }
  *L_out = L0;
  return X;
}

/* MinimizeforLaunchAngle (M.cc:873-917) */
double oracle_rootfn(const oracle_atm *a, double theta, double h, double ice, double depth_pos, double d) {
  double L, Xair, Xice = 0;
  Xair = air_x_only(a, theta, h, ice, &L);
  if (depth_pos != 0) Xice += seg_x(a, a->A_ice, depth_pos, 0.0, L, 0);
  return d - (Xice + Xair);
}

/* Fresnel transmission coefficients (M.cc:285-301, 321-337) */
static double trans_s(const oracle_atm *a, double thetai, double ice) {
  double n1 = oracle_nz_air(a, ice), n2 = oracle_nz_ice(a, 0);
  double sq = sqrt(1 - pow((n1 / n2) * (sin(thetai)), 2));
  double num = n1 * cos(thetai) - n2 * sq, den = n1 * cos(thetai) + n2 * sq;
  double t = 1 + (num / den);
  if (isnan(t)) t = 0;
  return t;
}
static double trans_p(const oracle_atm *a, double thetai, double ice) {
  double n1 = oracle_nz_air(a, ice), n2 = oracle_nz_ice(a, 0);
  double sq = sqrt(1 - pow((n1 / n2) * (sin(thetai)), 2));
  double num = n1 * sq - n2 * cos(thetai), den = n1 * sq + n2 * cos(thetai);
  double t = (1 - (num / den)) * (n1 / n2);
  if (isnan(t)) t = 0;
  return t;
}

/* gsl_root_fsolver_bisection driven as in FindFunctionRoot (M.cc:340-374): GSL roots/bisection.c +
 * gsl_root_test_interval(lo,hi,0,tol), at most 40 iterations.  Solver state starts zeroed (see
 * oracle/gsl_standin/gsl_standin.c header for why). */
typedef struct { const oracle_atm *a; double h, ice, depth_pos, d; int nevals; } rootctx_t;
static double F(rootctx_t *c, double x) { c->nevals++; return oracle_rootfn(c->a, x, c->h, c->ice, c->depth_pos, c->d); }

static double bisect_root(rootctx_t *c, double x_lo, double x_hi, double tol, int max_iter) {
  double f_lower = 0, f_upper = 0, root = 0, r = 0;
  int iter = 0, cont, set_ok = 0, have_fn = 0;
  if (!(x_lo > x_hi)) {
    have_fn = 1;
    root = 0.5 * (x_lo + x_hi);
    {
      double fl = F(c, x_lo);
      if (isfinite(fl)) {
        double fu = F(c, x_hi);
        if (isfinite(fu)) { f_lower = fl; f_upper = fu; set_ok = 1; }
      }
    }
  } else {
    /* gsl_root_fsolver_set refuses lo>hi before storing anything; the reference then iterates on a
     * solver that holds no function and no bracket (undefined behaviour in real GSL).  With the
     * zero-initialised stand-in state the 40 iterations all report root 0; the oracle does the same. */
    return 0.0;
  }
  (void)set_ok; (void)have_fn;
  do {
    double xl = x_lo, xr = x_hi;
    iter++;
    if (f_lower == 0.0) { root = xl; x_hi = xl; }
    else if (f_upper == 0.0) { root = xr; x_lo = xr; }
    else {
      double xb = (xl + xr) / 2.0;
      double fb = F(c, xb);
      if (!isfinite(fb)) { /* SAFE_FUNC_CALL returns before touching root/bracket */ }
      else if (fb == 0.0) { root = xb; x_lo = xb; x_hi = xb; }
      else if ((f_lower > 0.0 && fb < 0.0) || (f_lower < 0.0 && fb > 0.0)) { root = 0.5 * (xl + xb); x_hi = xb; f_upper = fb; }
      else { root = 0.5 * (xb + xr); x_lo = xb; f_lower = fb; }
    }
    r = root;
    {
      double al = fabs(x_lo), au = fabs(x_hi), mn;
      if ((x_lo > 0.0 && x_hi > 0.0) || (x_lo < 0.0 && x_hi < 0.0)) mn = al < au ? al : au; else mn = 0;
      cont = !(fabs(x_hi - x_lo) < 0 + tol * mn);
      if (x_lo > x_hi) cont = 0; /* GSL_EINVAL != GSL_CONTINUE */
    }
  } while (cont && iter < max_iter);
  return r;
}

/* Air2IceRayTracing (M.cc:1464-1616; P.cc:929-1086 for variant 1) */
int oracle_air2ice(const oracle_atm *a, double h, double d, double ice, double depth, double thR, double *out) {
  rootctx_t c;
  double lo, hi, theta, Xair = 0, tair = 0, pair_ = 0, L, inc_ice, Xice = 0, recv = 0, tice = 0, pice = 0;
  double walk[5 * 5 + 2];
  int filled, i;
  if (depth >= 0) { ice = depth + ice; depth = 0; c.depth_pos = depth; }
  else { c.depth_pos = -depth; }
  c.a = a; c.h = h; c.ice = ice; c.d = d; c.nevals = 0;
  lo = thR - 16;
  hi = thR;
  if (lo < 90.001) {
    int checknan = 0;
    lo = 90.001;
    while (checknan == 0 && lo > 89.9) {
      double Xt = 0;
      oracle_air_walk(a, lo, h, ice, walk);
      filled = (int)walk[5 * a->max_layers + 1];
      for (i = 0; i < filled; i++) Xt += walk[0 + i * 5];
      if ((isnan(Xt) == 0 && Xt > 0) || lo > hi - 0.1) checknan = 1; else lo = lo + 0.05;
    }
  }
  if (hi < 90.001 && hi > 90.00) hi = 90.05;
  theta = bisect_root(&c, lo, hi, 0.000000001, 40);

  oracle_air_walk(a, theta, h, ice, walk);
  filled = (int)walk[5 * a->max_layers + 1];
  for (i = 0; i < filled; i++) { Xair += walk[0 + i * 5]; tair += walk[3 + i * 5]; pair_ += walk[4 + i * 5]; }
  /* with no layer traversed (Tx outside every layer) the reference reads unset heap slots here
   * (M.cc:1537-1538); the oracle substitutes NaN so the case is at least reproducible */
  L = filled > 0 ? walk[2] : NAN;
  inc_ice = filled > 0 ? walk[1 + (filled - 1) * 5] : NAN;
  if (depth < 0) {
    double il[5];
    ice_leg(a, -depth, L, il);
    Xice = il[0]; recv = il[1]; tice = il[3]; pice = il[4];
  }
  {
    double X = Xice + Xair, t = tice + tair;
    out[0] = h; out[1] = X; out[2] = Xair; out[3] = Xice;
    out[4] = t * a->c; out[5] = tice * a->c; out[6] = tair * a->c;
    out[7] = t; out[8] = tice; out[9] = tair;
    out[10] = theta;
    if (a->variant == ORACLE_VARIANT_PYWRAP) { /* P.cc:1081-1084 */
      out[11] = asin((oracle_nz_air(a, ice) / oracle_nz_ice(a, 0)) * sin(inc_ice * (a->pi / 180))) * (180. / a->pi);
      out[12] = recv; out[13] = pair_; out[14] = pice; out[15] = 0; out[16] = inc_ice;
    } else { /* M.cc:1608-1614 */
      out[11] = recv;
      out[12] = trans_s(a, inc_ice * (a->pi / 180.0), ice);
      out[13] = trans_p(a, inc_ice * (a->pi / 180.0), ice);
      out[14] = pair_; out[15] = pice; out[16] = inc_ice;
    }
  }
  return c.nevals;
}

static double straight_angle(const oracle_atm *a, double h, double d, double ice, double depth) {
  double thR = 0; /* M.cc:952-958 */
  if (depth < 0) thR = 180 - (atan(d / (h - ice - depth)) * (180.0 / a->pi));
  if (depth >= 0) thR = 180 - (atan(d / (h - (ice + depth))) * (180.0 / a->pi));
  return thR;
}
static int check_solution(double thd, double d) { /* M.cc:974-983 */
  int ok = 0;
  if ((fabs(thd - d) / d < 0.01 && d <= 100) || (fabs(thd - d) < 1 && d > 100)) ok = 1;
  if (thd < 0) ok = 0;
  return ok;
}

/* GetHorizontalDistanceToIntersectionPoint (M.cc:945-989) */
int oracle_solve_cm(const oracle_atm *a, double h_cm, double d_cm, double depth_cm, double ice_cm, double *out) {
  double h = h_cm / 100, d = d_cm / 100, ice = ice_cm / 100, depth = depth_cm / 100;
  double thR = straight_angle(a, h, d, ice, depth);
  double r[17];
  oracle_air2ice(a, h, d, ice, depth, thR, r);
  out[0] = r[5] * 100;  /* opticalPathLengthInIce */
  out[1] = r[6] * 100;  /* opticalPathLengthInAir */
  out[2] = r[15] * 100; /* geometricalPathLengthInIce */
  out[3] = r[14] * 100; /* geometricalPathLengthInAir */
  out[4] = r[10] * (a->pi / 180); /* launchAngle */
  out[5] = r[2] * 100;  /* horizontalDistanceToIntersectionPoint */
  out[6] = r[12]; out[7] = r[13];
  out[8] = r[11] * (a->pi / 180); /* RecievedAngleInIce */
  return check_solution(r[1], d);
}
void oracle_solve_cm_batch(const oracle_atm *a, long n, const double *h_cm, const double *d_cm, double depth_cm,
                           double ice_cm, double *out, unsigned char *ok) {
  long i;
  for (i = 0; i < n; i++) ok[i] = (unsigned char)oracle_solve_cm(a, h_cm[i], d_cm[i], depth_cm, ice_cm, out + 9 * i);
}

/* AirIceRayTracing::GetRayTracingSolution (P.cc:884-927) */
int oracle_pywrap_solution(const oracle_atm *a, double h, double d, double depth, double ice, double *out) {
  double thR = straight_angle(a, h, d, ice, depth);
  double r[17];
  oracle_air2ice(a, h, d, ice, depth, thR, r);
  out[0] = r[5]; out[1] = r[6]; out[2] = r[14]; out[3] = r[13];
  out[4] = r[10]; out[5] = r[2]; out[6] = r[11]; out[7] = r[12];
  return check_solution(r[1], d);
}
/* TraceIceToAir / Py_TraceIceToAir (T.C:5-79) */
void oracle_py_trace(const oracle_atm *a, double depth, double ice, double h, double d, double *out) {
  double s[8];
  int ok = oracle_pywrap_solution(a, h, d, depth, ice, s);
  double launch = s[4], received = s[7], tmp;
  int i;
  tmp = launch; launch = received; received = tmp; /* std::swap, T.C:33 */
  received = 180 - received;                        /* T.C:34 */
  if (ok) {
    out[0] = h; out[1] = d; out[2] = s[2]; out[3] = s[3]; out[4] = launch; out[5] = received;
    out[6] = s[5]; out[7] = s[6]; out[8] = 0; out[9] = 0;
  } else {
    for (i = 0; i < 10; i++) out[i] = -1000;
  }
}

/* GetRayTracingSolutions (M.cc:1796-2017) */
void oracle_forward(const oracle_atm *a, double theta, double h, double ice, double depth, int inice, double *out) {
  int above, below, il, top, i;
  double start_angle = 0, Xair = 0, pair_ = 0, tair = 0, inc_ice, Xice = 0, pice = 0, tice = 0, recv_ice = 0;
  skip_layers(a, h, ice, &above, &below);
  top = a->max_layers - above - 1;
  for (il = top; il > below - 1; il--) {
    double hp[5];
    double start_h = (il == top) ? h : a->atmlay_cm[il + 1] / 100 - 0.00001;
    double start_n = oracle_nz_air(a, start_h);
    double stop_h = (il == (below - 1) + 1) ? ice : a->atmlay_cm[il] / 100;
    if (il == top) start_angle = 180 - theta;
    layer_hit(a, start_n, stop_h, start_h, start_angle, 1, hp);
    Xair += hp[0]; start_angle = hp[1]; tair += hp[3]; pair_ += hp[4];
  }
  inc_ice = start_angle;
  if (inice) {
    double hp[5];
    double start_n = oracle_nz_air(a, ice);
    layer_hit(a, start_n, -depth, 0.0, inc_ice, 0, hp);
    Xice += hp[0]; tice += hp[3]; pice += hp[4]; recv_ice = hp[1];
  }
  for (i = 0; i < 18; i++) out[i] = 0;
  out[1] = h;
  out[2] = Xair + Xice; out[3] = Xair; out[4] = Xice;
  out[5] = (tice + tair) * a->c; out[6] = tair * a->c; out[7] = tice * a->c;
  out[8] = (tice + tair) * pow(10, 9); out[9] = tair * pow(10, 9); out[10] = tice * pow(10, 9);
  out[11] = theta; out[12] = inc_ice; out[13] = recv_ice;
  out[14] = trans_s(a, inc_ice * (a->pi / 180.0), ice);
  out[15] = trans_p(a, inc_ice * (a->pi / 180.0), ice);
  out[16] = pair_; out[17] = pice;
}

/* MakeRayTracingTable (M.cc:2019-2158) */
oracle_table *oracle_table_build(const oracle_atm *a, double depth_cm, double ice_cm, double angle_step,
                                 double angle_start, double angle_stop, double height_step) {
  oracle_table *t = (oracle_table *)calloc(1, sizeof(oracle_table));
  int inice = depth_cm < 0 ? 1 : 0, ihei, iang, k;
  double depth = depth_cm / 100, ice = ice_cm / 100, h = 100000;
  long cell = 0;
  t->angle_step = angle_step; t->angle_start = angle_start; t->angle_stop = angle_stop; t->height_step = height_step;
  t->n_th = floor((angle_stop - angle_start) / angle_step) + 1; /* M.cc:15 */
  t->loop_start_h = h;
  t->loop_stop_h = inice ? ice : ice + depth;
  t->n_h = floor((t->loop_start_h - t->loop_stop_h) / height_step) + 1; /* M.cc:2061 */
  t->depth_m = depth; t->ice_m = ice;
  for (k = 0; k < 11; k++) t->col[k] = (float *)malloc(sizeof(float) * (size_t)t->n_h * (size_t)t->n_th);
  for (ihei = 0; ihei < t->n_h; ihei++) {
    h = t->loop_start_h - height_step * ihei;
    if (h > 0) {
      double r[18];
      for (iang = 0; iang < t->n_th; iang++) {
        double th = angle_start + angle_step * iang;
        if (h != t->loop_stop_h && ihei == t->n_h - 1) h = t->loop_stop_h;
        if (iang == t->n_th - 1) th = angle_stop;
        oracle_forward(a, th, h, t->loop_stop_h, depth, inice, r);
        t->col[0][cell] = r[1]; t->col[1][cell] = r[2]; t->col[2][cell] = r[7]; t->col[3][cell] = r[6];
        t->col[4][cell] = r[11]; t->col[5][cell] = r[3]; t->col[6][cell] = r[14]; t->col[7][cell] = r[15];
        t->col[8][cell] = r[16]; t->col[9][cell] = r[17]; t->col[10][cell] = r[13];
        cell++;
      }
    }
  }
  t->cells = cell;
  return t;
}
void oracle_table_free(oracle_table *t) {
  int k;
  if (!t) return;
  for (k = 0; k < 11; k++) free(t->col[k]);
  free(t);
}

/* FindClosestAirTxHeight (M.cc:1033-1126) */
void oracle_find_rows(const oracle_table *t, double h, int *idx, double *cv) {
  int cur = floor((h - t->loop_stop_h) / t->height_step);
  int Index = t->n_h - cur - 1;
  int MaxBin = Index * t->n_th + t->n_th - 1, MinBin = Index * t->n_th + 0;
  double val = -0.001;
  int StartBin = MaxBin, EndBin = MinBin, went = 0;
  while ((val != 0 && val < 0.01) || isnan(val)) { val = t->col[1][StartBin]; StartBin--; went = 1; }
  if (went) StartBin = StartBin + 1;
  val = -0.001; went = 0;
  while ((val != 0 && val < 0.01) || isnan(val)) { val = t->col[1][EndBin]; EndBin++; went = 1; }
  if (went) EndBin = EndBin - 1;
  idx[0] = EndBin; idx[1] = StartBin;
  cv[0] = fabs(t->col[0][Index] - h);
  idx[2] = idx[0] - t->n_th; idx[3] = idx[1] - t->n_th;
  if (idx[2] < 0) idx[2] = idx[0] + t->n_th;
  if (idx[3] < 0) idx[3] = idx[1] + t->n_th;
  cv[1] = fabs(t->col[0][Index] - h);
}
/* FindClosestTHD (M.cc:1128-1169) */
void oracle_find_thd(const oracle_table *t, double d, int StartIndex, int EndIndex, int *idx, double *cv) {
  int Mid, i, ipnt, index2 = 0;
  double minimum = 100000000000, minval, index1;
  for (i = 0; i < 8; i++) {
    if (EndIndex - StartIndex >= 3) {
      Mid = floor((StartIndex + EndIndex) / 2);
      if (t->col[1][Mid] - d > 0) StartIndex = Mid;
      if (t->col[1][Mid] - d < 0) EndIndex = Mid;
    }
  }
  for (ipnt = StartIndex; ipnt < EndIndex + 1; ipnt++) {
    minval = fabs(t->col[1][ipnt] - d);
    if (minval < minimum && t->col[1][ipnt] > d) minimum = minval;
    else { index2 = ipnt; break; }
  }
  index1 = index2 - 1;
  minimum = fabs(d - t->col[1][index2]);
  if (minimum > fabs(d - t->col[1][(int)index1])) minimum = fabs(d - t->col[1][(int)index1]);
  idx[0] = index1; idx[1] = index2; *cv = minimum;
}
static double lerp1(double x, double xa, double ya, double xb, double yb) { /* M.cc:992-995 */
  return ya + (yb - ya) * ((x - xa) / (xb - xa));
}
/* one row of GetParValues (M.cc:1199-1240 / 1250-1289) */
static void row_values(const oracle_table *t, double d, int s0, int e0, double *par) {
  int ip;
  double maxthd = t->col[1][s0];
  if (d <= maxthd) {
    int idx[2]; double cv;
    oracle_find_thd(t, d, s0, e0, idx, &cv);
    if (cv != 0) {
      double x1 = t->col[1][idx[0]], x2 = t->col[1][idx[1]];
      for (ip = 0; ip < 10; ip++) par[ip] = lerp1(d, x1, t->col[1 + ip][idx[0]], x2, t->col[1 + ip][idx[1]]);
    }
    if (cv == 0) {
      int s = idx[0] + 1;
      for (ip = 0; ip < 10; ip++) par[ip] = t->col[1 + ip][s];
    }
  } else {
    for (ip = 0; ip < 10; ip++) par[ip] = -pow(10, 9);
  }
}
/* GetHorizontalDistanceToIntersectionPoint_Table (M.cc:1305-1462) incl. GetParValues (M.cc:1172-1302).
 * ParInterpolatedValues is uninitialised in the reference when the height is out of range
 * (M.cc:1366,1405); the oracle zero-fills it (flag is false either way). */
int oracle_lookup_cm(const oracle_atm *a, const oracle_table *t, double h_cm, double d_cm, double depth_cm,
                     double ice_cm, double *out) {
  double h = h_cm / 100, d = d_cm / 100, ice = ice_cm / 100;
  int ok = 1, total = (int)t->cells - 1, ip;
  double maxh = t->col[0][0], minh = t->col[0][total];
  double y1 = 0, y2 = 0, x1 = 0, x2 = 0, P1[15], P2[15], PI[15];
  int solb = 0;
  (void)ice;
  for (ip = 0; ip < 15; ip++) { P1[ip] = 0; P2[ip] = 0; PI[ip] = 0; }
  if (h <= maxh && h >= minh && h > 0) {
    int idx[4]; double cv[2]; double h1, h2;
    oracle_find_rows(t, h, idx, cv);
    h1 = t->col[0][idx[0]];
    row_values(t, d, idx[0], idx[1], P1);
    if (cv[0] != 0 && h > minh && idx[2] < total) {
      h2 = t->col[0][idx[2]];
      row_values(t, d, idx[2], idx[3], P2);
    } else {
      h2 = h1;
      for (ip = 0; ip < 10; ip++) P2[ip] = P1[ip];
    }
    x1 = h1; x2 = h2;
    for (ip = 0; ip < 10; ip++) {
      double v = 0; int chk = 0;
      y1 = P1[ip]; y2 = P2[ip];
      if (y1 == -pow(10, 9) || y2 == -pow(10, 9)) chk = 1;
      if (x1 != x2 && chk == 0) v = lerp1(h, x1, y1, x2, y2);
      else {
        if (x1 == x2 && y1 == y2) v = P1[ip];
        if (y2 == -pow(10, 9) && y1 == -pow(10, 9)) ip = 9;
      }
      PI[ip] = v;
    }
  }
  {
    double THD = PI[0];
    out[0] = PI[1] * 100; out[1] = PI[2] * 100; out[2] = PI[8] * 100; out[3] = PI[7] * 100;
    out[4] = PI[3] * (a->pi / 180); out[5] = PI[4] * 100; out[6] = PI[5]; out[7] = PI[6];
    out[8] = PI[9] * (a->pi / 180);
    if ((y1 == -pow(10, 9) && y2 != -pow(10, 9)) || (y2 == -pow(10, 9) && y1 != -pow(10, 9))) {
      /* M.cc:1419: arguments scaled by 100 a second time and optical/geometric slots swapped */
      double r[9];
      solb = oracle_solve_cm(a, h_cm * 100, d_cm * 100, depth_cm * 100, ice * 100, r);
      out[2] = r[0]; out[3] = r[1]; out[0] = r[2]; out[1] = r[3];
      out[4] = r[4]; out[5] = r[5]; out[6] = r[6]; out[7] = r[7]; out[8] = r[8];
    }
    if (y2 == -pow(10, 9) && y1 == -pow(10, 9)) ok = 0;
    if (((y1 == -pow(10, 9) && y2 != -pow(10, 9)) || (y2 == -pow(10, 9) && y1 != -pow(10, 9))) && solb == 0) ok = 0;
    if (h > maxh) ok = 0;
    if (h < minh) ok = 0;
    if (h < 0) ok = 0;
    if (out[4] < 0) ok = 0;
    if ((fabs(THD - d) / d > 0.01 && d <= 100) || (fabs(THD - d) > 1 && d > 100)) ok = 0;
    if (!ok) { out[0] = 0; out[1] = 0; out[4] = 0; out[5] = 0; }
  }
  return ok;
}
void oracle_lookup_cm_batch(const oracle_atm *a, const oracle_table *t, long n, const double *h_cm, const double *d_cm,
                            double depth_cm, double ice_cm, double *out, unsigned char *ok) {
  long i;
  for (i = 0; i < n; i++) ok[i] = (unsigned char)oracle_lookup_cm(a, t, h_cm[i], d_cm[i], depth_cm, ice_cm, out + 9 * i);
}

/* Ray-path dump of the reference's CLI, SingleRayAirIceRefraction.C:133-152 (layer walk carrying the first layer's L)
 * and :226-299 (1 m polyline in air, then in ice).  depth_pos > 0 as the CLI takes it; x/z receive at most max_points
 * entries; returns the number of points of the full path.  The file the CLI writes holds "index x z" with 6 digits. */
static double path_F(double x, double A, double B, double C, double L, double n) { /* R.cc fDnfR with n supplied */
  return (L / C) * (1.0 / sqrt(A * A - L * L)) * (C * x - log(A * n - L * L + sqrt(A * A - L * L) * sqrt(pow(n, 2) - L * L)));
}
long oracle_ray_path(const oracle_atm *a, double theta, double h, double ice, double depth_pos, long max_points, double *x, double *z) {
  int above, below, il, top, first = 1;
  double walk[5 * 5 + 2];
  double L, last_x = 0, last_h = 0, rx = 0, i;
  long ip = 0;
  skip_layers(a, h, ice, &above, &below);
  top = a->max_layers - above - 1;
  oracle_air_walk(a, theta, h, ice, walk);
  L = walk[2];                                  /* Lvalue = GetHitPar[2] of the first layer, reused below (:140,:148) */
  for (il = top; il > below - 1; il--) {
    double start = first ? h : last_h - 0.00001;
    double stop = (il == (below - 1) + 1) ? ice : a->atmlay_cm[il] / 100;
    for (i = start; i > stop - 1; i = i - 1) {
      int k;
      double B, C;
      if (i < stop) i = stop;
      k = oracle_layer_of(a, i);                /* GetB_air(-i), GetC_air(-i): parameters looked up at the point itself */
      B = a->B_air[k]; C = a->C_air[k];
      rx = path_F(-i, a->A_air, B, C, L, oracle_nz_air(a, -i)) - path_F(-start, a->A_air, B, C, L, oracle_nz_air(a, -start)) + last_x;
      if (ip < max_points) { x[ip] = rx; z[ip] = i; }
      ip++;
      last_h = i;
    }
    last_x = rx;
    first = 0;
  }
  {
    int ii;
    for (ii = 0; ii > -(depth_pos + 1); ii--) {
      double px = last_x - path_F((double)ii, a->A_ice, a->B_ice, a->C_ice, L, oracle_nz_ice(a, (double)ii)) +
                  path_F(0, a->A_ice, a->B_ice, a->C_ice, L, oracle_nz_ice(a, 0));
      if (ip < max_points) { x[ip] = px; z[ip] = (double)ii + ice; }
      ip++;
    }
  }
  return ip;
}
