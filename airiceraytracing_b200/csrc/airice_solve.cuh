// airice_solve.cuh -- per-pair launch-angle solve (one thread per Tx->Rx pair).
//
// What the reference does (MultiRayAirIceRefraction.cc:1464-1616): build the bracket
// [straight-16 deg, straight] (with a 0.05-deg scan away from the NaN zone near 90 deg), run
// gsl_root_fsolver_bisection on f(theta) = d - X(theta) until the bracket is narrower than
// 1e-9*min(|lo|,|hi|) (about 1.5e-7 deg, ~27 halvings, each a full layer walk), return the midpoint of
// the last bracket, then evaluate the ray once more at that angle.
//
// What this does instead: (1) find the true root theta* of f on t = tan(incidence at the transmitter), in which X
// is almost linear: two single-precision Newton iterations (analytic dX/dL from the closed forms), then a
// safeguarded FP64 chord iteration (typically 2 evaluations); (2) REPLAY the reference's bisection without
// evaluating f: for a monotone f the sign of f(mid) is the side of theta* that mid lies on, so the ~27
// halvings cost a compare each.  Only a midpoint that falls within `guard` degrees of theta* is
// evaluated for real.  The replay lands on the very bracket the reference ends in, so the returned
// launch angle is the reference's (to the last bit, rounding ties aside), not merely a better root.
#pragma once
#include "airice_core.cuh"

#define AIRICE_SOLVE_GUARD_DEG 1.0e-10
#define AIRICE_NEWTON_MAXIT 40
// error bound below which the Newton step of the single FP64 evaluation is taken as the root; the replay evaluates f
// for real whenever a probe lies within AIRICE_SOLVE_GUARD_DEG of it, so this only has to stay well below the guard
#define AIRICE_SOLVE_ACCEPT_DEG 5.0e-11

struct AirIceSolveStat {
  int n_newton;  // distance evaluations spent in the Newton phase
  int n_replay;  // real evaluations spent inside the guard band during the replay
};

// floor(log2(x)) of a positive normal double, and x * 2^k (exact while the result stays normal)
AIRICE_HD int airice_exponent(double x) {
#if defined(__CUDA_ARCH__)
  return ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
#else
  int e; frexp(x, &e); return e - 1;
#endif
}
AIRICE_HD double airice_scale2(double x, int k) {
#if defined(__CUDA_ARCH__)
  return x * __hiloint2double((1023 + k) << 20, 0);
#else
  return ldexp(x, k);
#endif
}

AIRICE_HD double airice_L_of_theta(const AirIceMedium& m, double n_tx, double theta) {
  // first-segment Snell chain of GetLayerHitPointPar (M.cc:537-589) collapses to n(h_tx) sin(180-theta)
  // (the library's sin: an own first-quadrant sine -- both fdlibm kernels on the reduced argument, one selected --
  // measured 1.417 against 1.412 ms per 1e7 solves, 0.497 against 0.503 ms per reference-grid table)
  return n_tx * sin((180 - theta) * m.deg2rad);
}

// Bracket of M.cc:1487-1516.  Outputs lo, hi, t_cap = tan(incidence at Tx) of the lower end (the largest t the
// root may have) and whether f(lo) is finite (n_tx sin(180-lo) < 1).  `ta` = d/(h - surface - depth) = tan of the
// straight-line incidence; tan16 = tan(16 deg) in the variant's own pi.
AIRICE_HD void airice_bracket(const AirIceMedium& m, const AirIcePlan& p, int kt, double n_tx, double thR, double ta,
                              double& lo, double& hi, double& t_cap, bool& finite_lo) {
  const bool walk = (kt >= p.kb);
  lo = thR - 16;
  hi = thR;
  if (lo < 90.001) {
    // The reference scans lo = 90.001, +0.05, ... while the ray does not exist (n_tx sin(180-lo) >= 1) and lo has
    // not passed hi - 0.1.  The visited angles and their sines do not depend on the ray (m.clamp_tab, accumulated and
    // evaluated on the host exactly as the reference does), and both stop conditions are monotone in j, so the scan
    // is two bracketed searches from closed-form estimates: no trigonometry and no loop over the NaN zone here.
    const double lim = hi - 0.1;
    const double* tab = m.clamp_tab;
    int jh = 0;
    {
      const double e = ceil((lim - 90.001) * 20.0);   // an estimate: the two loops below settle it
      jh = e > 0.0 ? (e < (double)(AIRICE_CLAMP_N - 1) ? (int)e : AIRICE_CLAMP_N - 1) : 0;
#pragma unroll 1
      while (jh > 0 && AIRICE_LDG(tab + 2 * (jh - 1)) > lim) jh--;
#pragma unroll 1
      while (jh < AIRICE_CLAMP_N - 1 && !(AIRICE_LDG(tab + 2 * jh) > lim)) jh++;
    }
    int j = jh;
    if (walk && n_tx > 1.0) {
      // first j with 1 - (n_tx sin_j)^2 > 0: sin_j = cos(lo_j - 90 deg) ~ 1 - eps^2/2 < 1/n_tx  <=>  eps > sqrt(2 (n_tx - 1))
      const float est = (sqrtf(2.0f * (float)(n_tx - 1.0)) * (float)m.rad2deg - 0.001f) * 20.0f;
      int jo = est > 0.0f ? (est < (float)jh ? (int)est : jh) : 0;
#pragma unroll 1
      while (jo > 0) {
        const double Lp = n_tx * AIRICE_LDG(tab + 2 * (jo - 1) + 1);
        if (!(1.0 - Lp * Lp > 0.0)) break;
        jo--;
      }
#pragma unroll 1
      while (jo < jh) {
        const double Lp = n_tx * AIRICE_LDG(tab + 2 * jo + 1);
        if (1.0 - Lp * Lp > 0.0) break;
        jo++;
      }
      j = jo;
    }
    lo = AIRICE_LDG(tab + 2 * j);
    const double L = n_tx * AIRICE_LDG(tab + 2 * j + 1);
    finite_lo = walk && (1.0 - L * L > 0.0);
    t_cap = AIRICE_DIV(L, AIRICE_SQRT(n_tx * n_tx - L * L));
  } else {
    // lo = thR-16 exactly: tan(inc_lo) = tan(inc_hi + 16 deg) by the addition formula, no trig call
    const double t16 = m.tan16;
    t_cap = AIRICE_DIV(ta + t16, 1.0 - ta * t16);
    finite_lo = walk && (t_cap * t_cap * (n_tx * n_tx - 1.0) < 1.0);
  }
  if (hi < 90.001 && hi > 90.00) hi = 90.05;  // M.cc:1513-1516
}

// Returns the launch angle the reference's bisection returns.  theta_star (the converged root, or
// -inf/+inf when f has one sign on the bracket) is exposed for diagnostics.
// DEFER = true is the first pass of the two-pass launch (kernels.cu): a pair that needs one of the rare slow paths
// (failed single-precision pre-iteration, rejected Newton step, a real evaluation of f at a tie, the literal bisection
// loop) is not solved here -- `hard` is set, the result is NaN -- and the slow paths are not even compiled in; the
// second pass runs those pairs (0.6 % of a random batch) through DEFER = false in dense warps.
template <bool DEFER>
AIRICE_HD double airice_solve_theta_t(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx,
                                      double d, double thR, double ta_straight, double& theta_star,
                                      AirIceSolveStat& st, bool& hard) {
  st.n_newton = 0; st.n_replay = 0;
  hard = false;
  theta_star = NAN;
  bool finite_lo;
  double lo, hi, t_cap;
  airice_bracket(m, p, kt, n_tx, thR, ta_straight, lo, hi, t_cap, finite_lo);
  if (!(lo <= hi)) return NAN;      // gsl_root_fsolver_set rejects lo>hi; the reference result is undefined
  if (!finite_lo) return lo;        // f(lo) not finite: solver state never set (UB cases)

  // ---- phase 1: theta* on t = tan(incidence at Tx), in which X(t) is almost linear.
  //  (a) two Newton iterations in SINGLE precision (X and dX/dt from airice_x_newton_f32) bring t within ~1e-4 deg of
  //      the root and leave a slope good to ~1e-5; they run on the FP32/MUFU pipes this FP64-bound kernel leaves idle;
  //  (b) a safeguarded chord iteration in FP64 (X only, airice_x_fast) with that slope -- or a secant slope when a
  //      step was long enough to measure one -- finishes: error after a step = (slope error) x step + O(step^2), so a
  //      step below 1e-7 deg with a good slope leaves ~1e-12 deg.  Typically 2 FP64 evaluations, the same for every
  //      lane of a warp.
  double ts;                                        // root in t, or +-inf
  if (!(d > 0.0)) {
    ts = (d == 0.0) ? 0.0 : -INFINITY;  // X>=0: d<0 means f<0 everywhere
  } else {
    double tlo = 0.0, thi = INFINITY;  // g(tlo)<0<g(thi), g = X-d
    const double hgt = (h - p.ice_h) + 0.55 * p.depth;
    double t = AIRICE_DIV(d, hgt);
    if (!(t < t_cap)) t = t_cap;
    double slope = hgt;                 // dX/dt ~ height for a straight ray; replaced below
    bool slope_good = false;
    bool have_kappa = false;
    double kappa_bound = 0.0;            // |X''/X'| plus its uncertainty, from the single-precision slopes
    if (kt >= p.kb) {
      const int kc = kt < 0 ? 0 : kt;
      const float dn_tx = (float)(n_tx - 1.0);
      const float cdx_top = (float)(p.seg[kc].neg_c * (p.seg[kc].stop_x - h));
      const float df = (float)d, capf = (float)t_cap;
      float tf = (float)t, sf = 0.0f;
      float t_a = 0.0f, s_a = 0.0f, t_b = 0.0f;   // the two evaluation points and the first slope: curvature estimate
      bool okf = true;
#pragma unroll 1
      for (int it = 0; it < 2 && okf; it++) {
        float dXdt;
        const float Xf = airice_x_newton_f32(m, p, kt, cdx_top, dn_tx, tf, dXdt);
        const float tn = tf - (Xf - df) / dXdt;
        okf = (tn > 0.0f) && (tn < capf) && (dXdt > 0.0f);
        if (okf) { t_a = t_b; s_a = sf; t_b = tf; tf = tn; sf = dXdt; }
      }
      if (sf > 0.0f) { t = (double)tf; slope = (double)sf; slope_good = true; }
      if (okf && s_a > 0.0f) {
        // both single-precision iterations went through: kappa ~ X''/X' from their two slopes, with the noise of a
        // single-precision slope (~1e-5 relative) over the distance of the two points as its uncertainty
        const double dab = (double)t_b - (double)t_a;
        kappa_bound = AIRICE_DIV(fabs((double)sf - (double)s_a) + 3.0e-5 * (double)sf, fabs(dab) * (double)sf);
        have_kappa = true;
      }
    }
    double t_prev = 0.0, g_prev = 0.0;
    bool have_prev = false;
    ts = NAN;
    if (have_kappa) {
      // (b1) ONE FP64 evaluation with the analytic slope.  The Newton step from t lands within
      // 0.5 |X''/X'| dt^2 of the root; with the bound on the curvature from (a) that is below 5e-11 deg for 99.9 % of
      // pairs (the single-precision landing point is ~1e-6 deg off), well inside the replay's guard band, and the
      // step is taken as the root.  Otherwise the chord iteration below continues from the step with this slope.
      double sq1, w;
      AIRICE_SQRT_RSQRT(1.0 + t * t, sq1, w);
      w = AIRICE_RCP(sq1);
      double dXdL;
      const double X = airice_x_dx(m, p, kt, h, n_tx, n_tx * t * w, dXdL);
      st.n_newton++;
      const double g = X - d;
      const double s = dXdL * n_tx * (w * w * w);          // dL/dt = n_tx / (1+t^2)^{3/2}
      if (g == 0.0) {
        ts = t;
      } else if (g < 0.0 && t >= t_cap) {
        ts = INFINITY;                                     // even the lower bracket end falls short of d
      } else {
        if (g < 0.0) tlo = t; else thi = t;
        const double hi_t = thi < t_cap ? thi : t_cap;
        const double dN = -AIRICE_DIV(g, s);
        double tn = t + dN;
        const bool inside = (s > 0.0) && (tn > tlo) && (tn < hi_t);
        const double bound_deg = 0.5 * kappa_bound * (dN * dN) * (w * w) * m.rad2deg;
        if (inside && bound_deg < AIRICE_SOLVE_ACCEPT_DEG) {
          ts = tn;
        } else {
          if (!inside) tn = (thi == INFINITY && g < 0.0) ? t_cap : 0.5 * (tlo + hi_t);
          if (s > 0.0) { slope = s; slope_good = true; }
          have_prev = true; t_prev = t; g_prev = g;
          t = tn;
        }
      }
    }
    if (DEFER && ts != ts) { hard = true; return NAN; }
#pragma unroll 1
    for (int it = 0; it < AIRICE_NEWTON_MAXIT && ts != ts; it++) {
      double sq1, w;
      AIRICE_SQRT_RSQRT(1.0 + t * t, sq1, w);
      w = AIRICE_RCP(sq1);
      const double L = n_tx * t * w;
      const double X = airice_x_fast(m, p, kt, h, n_tx, L);
      st.n_newton++;
      const double g = X - d;
      if (g == 0.0) { ts = t; break; }
      if (g < 0.0) {
        tlo = t;
        if (t >= t_cap) { ts = INFINITY; break; }  // even the lower bracket end falls short of d
      } else {
        thi = t;
      }
      if (have_prev && fabs(t - t_prev) > 1.0e-5 * t) {   // long enough step: the secant measures the slope
        const double sec = AIRICE_DIV(g - g_prev, t - t_prev);
        if (sec > 0.0) { slope = sec; slope_good = fabs(t - t_prev) < 1.0e-2 * t; }
      }
      double tn = t - AIRICE_DIV(g, slope);
      const double hi_t = thi < t_cap ? thi : t_cap;
      bool chord_step = true;
      if (!(tn > tlo) || !(tn < hi_t)) {
        // outside the bracket (or NaN): probe the cap once if it is still untested, else bisect
        chord_step = false;
        if (thi == INFINITY && g < 0.0) tn = t_cap; else tn = 0.5 * (tlo + hi_t);
      }
      const double step_deg = fabs(tn - t) * w * w * m.rad2deg;
      if (chord_step && slope_good && step_deg < 1.0e-7) { ts = tn; break; }
      have_prev = true; t_prev = t; g_prev = g;
      t = tn;
      if (thi < INFINITY && thi - tlo <= 4.0e-16 * thi) { ts = 0.5 * (tlo + thi); break; }
    }
    if (ts != ts) ts = t;  // iteration cap: take the last iterate
  }
  // theta* in degrees; +inf in t (root beyond the lower end) maps below the bracket, -inf above it
  if (ts == INFINITY) theta_star = -INFINITY;
  else if (ts == -INFINITY) theta_star = INFINITY;
  else theta_star = 180 - AIRICE_ATAN_Q(ts, 1.0) * m.rad2deg;

  // ---- phase 2: replay of gsl_root_fsolver_bisection + gsl_root_test_interval (M.cc:355-369) in closed form.
  // For a monotone f the sign of f(x) is the side of theta* that x lies on, so GSL's rule "keep the half whose ends
  // differ in sign" keeps the half that contains theta* (and always the UPPER half when theta* is outside the bracket:
  // f(lo) and f(mid) then have the same sign at every step).  After k halvings the bracket is the cell
  //   [lo + j W 2^-k, lo + (j+1) W 2^-k],  j = floor((theta* - lo) 2^k / W)   (j = 2^k - 1 when outside),
  // the loop stops at the first k whose cell is narrower than 1e-9 * (its lower end), and the returned root is the
  // midpoint of that cell.  The iterated midpoints of the reference sit on this grid to within their rounding
  // (~3e-14 deg), so the replay costs a few dozen instructions instead of ~27 loop trips.  Left to the careful loop
  // below: a bracket END within `guard` of theta*, an exact zero of f at a probe, an empty bracket.  A grid point
  // (= a midpoint the reference probes) within `guard` of theta* gets its sign from one real evaluation of f.
  const double guard = AIRICE_SOLVE_GUARD_DEG;
  const double th = theta_star;
  {
    const double W = hi - lo;
    bool careful = !(fabs(lo - th) > guard) || !(fabs(hi - th) > guard) || !(W > 0.0) || !(lo > 0.0);
    if (!careful) {
      // kb = first k with W 2^-k < 1e-9 hi: no earlier cell can pass the test (its lower end is < hi); cell kb + 1
      // always passes (lo > hi/2 here: lo >= 90.001, hi - lo <= 16).  At least one halving is made.
      const double ratio = W * AIRICE_RCP(0.000000001 * hi);
      int kb = airice_exponent(ratio) + 1;
      kb = kb < 1 ? 1 : kb;
      const int kf = kb + 1;
      const double wf = airice_scale2(W, -kf);            // cell width W 2^-kf after kb + 1 halvings (exact)
      const double ncell = airice_scale2(1.0, kf);        // 2^kf
      const bool inside = th > lo && th < hi;             // otherwise: the walk to hi, last cell of every level
      const double x = inside ? (th - lo) * (ncell * AIRICE_RCP(W)) : ncell - 0.5;
      // does the loop stop after kb halvings?  (cell index there = floor(x / 2))
      const double jb = floor(0.5 * x), wb = 2.0 * wf;
      const double lob = lo + jb * wb, hib = lo + (jb + 1.0) * wb;
      const bool stop_b = hib - lob < 0.000000001 * lob;
      const double xk = stop_b ? 0.5 * x : x, wk = stop_b ? wb : wf, nk = stop_b ? 0.5 * ncell : ncell;
      const double mr = rint(xk);
      double jk = (xk < mr) ? mr - 1.0 : mr;              // floor(xk): the final cell
      if (inside && !(fabs(xk - mr) * wk > guard)) {
        // theta* within the guard of grid point mr of the final level: an end of the final cell, i.e. a midpoint the
        // reference evaluated f at
        if (DEFER) { hard = true; return NAN; }
        if (!(mr > 0.0) || !(mr < nk)) {
          careful = true;
        } else {
          const double xp = lo + mr * wk;
          const double f = d - airice_x_exact(m, p, kt, h, n_tx, airice_L_of_theta(m, n_tx, xp));
          st.n_replay++;
          if (f < 0.0) jk = mr;                           // grid point below the root: the cell starts there
          else if (f > 0.0) jk = mr - 1.0;
          else careful = true;                            // exact zero or NaN: GSL's special exits
        }
      }
      if (!careful) return lo + (2.0 * jk + 1.0) * (0.5 * wk);
    }
  }
  // Careful form (about 0.2 % of solves): a probe sits within `guard` of theta*, so f is evaluated there for real
  // (MinimizeforLaunchAngle, M.cc:873-917), including GSL's exact-zero exits.  One loop serves the two endpoint
  // signs (steps -2, -1: gsl_root_fsolver_set) and the halvings (steps >= 0).
  if (DEFER) { hard = true; return NAN; }
  int s_lo = 1, s_hi = 1;
  double root = 0.5 * (lo + hi);
#pragma unroll 1
  for (int step = -2; step < 40; step++) {
    const bool probing = step < 0;
    const bool zero_end = !probing && (s_lo == 0 || s_hi == 0);
    const double x = (step == -2) ? lo : ((step == -1) ? hi : (lo + hi) / 2.0);
    int s = 0;
    if (!zero_end) {
      if (fabs(x - th) > guard) {
        s = (x < th) ? -1 : 1;
      } else {
        const double f = d - airice_x_exact(m, p, kt, h, n_tx, airice_L_of_theta(m, n_tx, x));
        st.n_replay++;
        s = (f < 0.0) ? -1 : ((f > 0.0) ? 1 : 0);
      }
    }
    if (step == -2) { s_lo = s; continue; }
    if (step == -1) { s_hi = s; continue; }
    if (s_lo == 0) { root = lo; hi = lo; }
    else if (s_hi == 0) { root = hi; lo = hi; }
    else if (s == 0) { root = x; lo = x; hi = x; }
    else if (s_lo * s < 0) { root = 0.5 * (lo + x); hi = x; s_hi = s; }
    else { root = 0.5 * (x + hi); lo = x; s_lo = s; }
    const double al = fabs(lo), au = fabs(hi);
    const double mn = ((lo > 0.0 && hi > 0.0) || (lo < 0.0 && hi < 0.0)) ? (al < au ? al : au) : 0.0;
    if (fabs(hi - lo) < 0.000000001 * mn) break;
  }
  return root;
}

// The launch angle the reference's COMMAND-LINE solver returns (Air2IceRayTracing.C:101-137 on RayTracingFunctions.cc):
// bracket [straight - 16, straight]; a lower end below 90.00 becomes 90.05 and steps up by 0.05 while the ray does not
// exist there (X_air NaN or <= 0) and lo <= hi - 1; then gsl_root_fsolver_brent on f = d - X(theta) with
// gsl_root_test_interval(lo, hi, 0, 1e-9) and at most 20 iterations (RayTracingFunctions.cc:256-291).  Brent's iterates
// depend on the VALUES of f, not only on its sign, so there is nothing to replay: the loop runs literally (GSL's
// roots/brent.c, ~6 evaluations of X with reference-like rounding), rare non-finite bracket ends included (the solver
// state then stays zeroed and the "root" is 0, as with the calloc'ed stand-in of the oracle).
AIRICE_HD double airice_solve_theta_cli(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx, double d,
                                        double thR, int& nevals) {
  nevals = 0;
  double lo = thR - 16, hi = thR;
  if (lo < 90.00) {
    lo = 90.05;
#pragma unroll 1
    while (lo > 89.9) {
      const double xa = (kt >= p.kb) ? airice_x_exact(m, p, kt, h, n_tx, airice_L_of_theta(m, n_tx, lo), false) : 0.0;
      nevals++;
      if ((xa == xa && xa > 0) || lo > hi - 1) break;
      lo = lo + 0.05;
    }
  }
  if (hi < 90.001 && hi > 90.00) hi = 90.05;
  auto f = [&](double th) { nevals++; return d - airice_x_exact(m, p, kt, h, n_tx, airice_L_of_theta(m, n_tx, th)); };
  // gsl_root_fsolver_set -> brent_init
  double a = 0, b = 0, c = 0, dd = 0, e = 0, fa = 0, fb = 0, fc = 0, root = 0;
  double x_lower = lo, x_upper = hi;
  if (!(lo > hi)) {                       // gsl_root_fsolver_set refuses lo > hi before calling the type's init
    root = 0.5 * (lo + hi);
    const double f_lower = f(lo);
    bool ok = isfinite(f_lower);
    double f_upper = 0;
    if (ok) { f_upper = f(hi); ok = isfinite(f_upper); }
    if (ok) { a = lo; fa = f_lower; b = hi; fb = f_upper; c = hi; fc = f_upper; dd = hi - lo; e = hi - lo; }
  } else {
    root = 0;
  }
#pragma unroll 1
  for (int iter = 0; iter < 20; iter++) {
    // brent_iterate
    bool ac_equal = false;
    if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = true; c = a; fc = fa; dd = b - a; e = b - a; }
    if (fabs(fc) < fabs(fb)) { ac_equal = true; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    const double tol = 0.5 * 2.2204460492503131e-16 * fabs(b);
    const double mm = 0.5 * (c - b);
    bool done_step = false;
    if (fb == 0) { root = b; x_lower = b; x_upper = b; done_step = true; }
    else if (fabs(mm) <= tol) {
      root = b;
      if (b < c) { x_lower = b; x_upper = c; } else { x_lower = c; x_upper = b; }
      done_step = true;
    }
    if (!done_step) {
      if (fabs(e) < tol || fabs(fa) <= fabs(fb)) { dd = mm; e = mm; }
      else {
        double pp, q, r;
        const double s = fb / fa;
        if (ac_equal) { pp = 2 * mm * s; q = 1 - s; }
        else {
          q = fa / fc; r = fb / fc;
          pp = s * (2 * mm * q * (q - r) - (b - a) * (r - 1));
          q = (q - 1) * (r - 1) * (s - 1);
        }
        if (pp > 0) q = -q; else pp = -pp;
        const double lim1 = 3 * mm * q - fabs(tol * q), lim2 = fabs(e * q);
        if (2 * pp < (lim1 < lim2 ? lim1 : lim2)) { e = dd; dd = pp / q; }
        else { dd = mm; e = mm; }
      }
      a = b; fa = fb;
      double bn = b;
      if (fabs(dd) > tol) bn += dd; else bn += (mm > 0 ? +tol : -tol);
      const double fbn = f(bn);
      if (isfinite(fbn)) {              // else GSL_EBADFUNC: the iterate returns before storing its state
        b = bn; fb = fbn;
        root = b;
        double cc = c;
        if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) cc = a;
        if (b < cc) { x_lower = b; x_upper = cc; } else { x_lower = cc; x_upper = b; }
      } else {
        // a, fa were local copies in GSL: the stored state keeps the values from before this iterate
        // (restore what the failed iterate must not have changed)
        // NB: c/fc/dd/e changes above are local in GSL as well
        return root;                     // every further iterate repeats the same failure: root stays
      }
    }
    // gsl_root_test_interval(x_lower, x_upper, 0, 1e-9)
    if (x_lower > x_upper) break;        // GSL_EINVAL != GSL_CONTINUE
    const double al = fabs(x_lower), au = fabs(x_upper);
    const double mn = ((x_lower > 0.0 && x_upper > 0.0) || (x_lower < 0.0 && x_upper < 0.0)) ? (al < au ? al : au) : 0.0;
    if (fabs(x_upper - x_lower) < 0.000000001 * mn) break;
  }
  return root;
}

AIRICE_HD double airice_solve_theta(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx,
                                    double d, double thR, double ta_straight, double& theta_star,
                                    AirIceSolveStat& st) {
  bool hard;
  return airice_solve_theta_t<false>(m, p, kt, h, n_tx, d, thR, ta_straight, theta_star, st, hard);
}

// Straight-line angle of GetHorizontalDistanceToIntersectionPoint (M.cc:952-958), metres in.  ta = its tangent.
AIRICE_HD double airice_straight_angle(const AirIceMedium& m, double h, double d, double ice, double depth_signed,
                                       double& ta) {
  const double den = (depth_signed < 0) ? (h - ice - depth_signed) : (h - (ice + depth_signed));
  ta = d / den;
  return 180 - (AIRICE_ATAN_Q(ta, 1.0) * m.rad2deg);
}

// Solution flag of M.cc:974-983.
AIRICE_HD bool airice_check_solution(double thd, double d) {
  bool ok = false;
  // (|thd-d|/d < 0.01 && d <= 100) || (|thd-d| < 1 && d > 100), with the division only where it is needed
  if (d <= 100 ? (fabs(thd - d) / d < 0.01) : (fabs(thd - d) < 1)) ok = true;
  if (thd < 0) ok = false;
  return ok;
}
