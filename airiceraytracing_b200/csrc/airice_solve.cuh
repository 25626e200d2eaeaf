// airice_solve.cuh -- per-pair launch-angle solve (one thread per Tx->Rx pair).
//
// What the reference does (MultiRayAirIceRefraction.cc:1464-1616): build the bracket
// [straight-16 deg, straight] (with a 0.05-deg scan away from the NaN zone near 90 deg), run
// gsl_root_fsolver_bisection on f(theta) = d - X(theta) until the bracket is narrower than
// 1e-9*min(|lo|,|hi|) (about 1.5e-7 deg, ~27 halvings, each a full layer walk), return the midpoint of
// the last bracket, then evaluate the ray once more at that angle.
//
// What this does instead: (1) find the true root theta* of f with a safeguarded Newton iteration on
// t = tan(incidence at the transmitter), in which X is almost linear, using the analytic dX/dL that
// falls out of the closed forms (3-5 evaluations); (2) REPLAY the reference's bisection without
// evaluating f: for a monotone f the sign of f(mid) is the side of theta* that mid lies on, so the ~27
// halvings cost a compare each.  Only a midpoint that falls within `guard` degrees of theta* is
// evaluated for real.  The replay lands on the very bracket the reference ends in, so the returned
// launch angle is the reference's (to the last bit, rounding ties aside), not merely a better root.
#pragma once
#include "airice_core.cuh"

#define AIRICE_SOLVE_GUARD_DEG 1.0e-10
#define AIRICE_NEWTON_MAXIT 40
#ifndef AIRICE_HERMITE_ACCEPT
#define AIRICE_HERMITE_ACCEPT 1.0e-6   // (pending step [deg]) x (previous step [deg]) below which the Hermite root is taken
#endif

struct AirIceSolveStat {
  int n_newton;  // distance evaluations spent in the Newton phase
  int n_replay;  // real evaluations spent inside the guard band during the replay
};

AIRICE_HD double airice_L_of_theta(const AirIceMedium& m, double n_tx, double theta) {
  // first-segment Snell chain of GetLayerHitPointPar (M.cc:537-589) collapses to n(h_tx) sin(180-theta)
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
    lo = 90.001;
    // closed-form jump over the NaN zone: X(theta) is NaN exactly while n_tx sin(180-theta) >= 1
    int k = 0;
    if (walk && n_tx > 1.0) {
      const double th_c = 180 - asin(1.0 / n_tx) * m.rad2deg;
      const double kk = ceil((th_c - 90.001) / 0.05) - 1.0;
      k = kk > 0.0 ? (int)kk : 0;
    }
#pragma unroll 1
    for (int i = 0; i < k && !(lo > hi - 0.1); i++) lo = lo + 0.05;
    // finish with the literal loop (0-2 iterations) so that rounding at the zone edge matches
    double L = 0.0;
#pragma unroll 1
    for (int it = 0; it < 4000; it++) {
      L = airice_L_of_theta(m, n_tx, lo);
      const bool okx = walk && (1.0 - L * L > 0.0);
      if (okx || lo > hi - 0.1) break;
      lo = lo + 0.05;
    }
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
AIRICE_HD double airice_solve_theta(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx,
                                    double d, double thR, double ta_straight, double& theta_star,
                                    AirIceSolveStat& st) {
  st.n_newton = 0; st.n_replay = 0;
  theta_star = NAN;
  bool finite_lo;
  double lo, hi, t_cap;
  airice_bracket(m, p, kt, n_tx, thR, ta_straight, lo, hi, t_cap, finite_lo);
  if (!(lo <= hi)) return NAN;      // gsl_root_fsolver_set rejects lo>hi; the reference result is undefined
  if (!finite_lo) return lo;        // f(lo) not finite: solver state never set (UB cases)

  // ---- phase 1: theta* by safeguarded Newton on t = tan(incidence at Tx), finished by two-point inverse Hermite
  // interpolation: with (t,g,g') at the last two iterates, the cubic t(g) through both points and slopes is
  // evaluated at g=0.  Its error is ~ e_prev^2 e_cur^2, so once the pending Newton step is small the Hermite root is
  // already converged and the confirming evaluation (a third of the Newton work, and the cause of most intra-warp
  // iteration-count divergence) is skipped.
  double ts;                                        // root in t, or +-inf
  if (!(d > 0.0)) {
    ts = (d == 0.0) ? 0.0 : -INFINITY;  // X>=0: d<0 means f<0 everywhere
  } else {
    double tlo = 0.0, thi = INFINITY;  // g(tlo)<0<g(thi), g = X-d
    double t = AIRICE_DIV(d, (h - p.ice_h) + 0.55 * p.depth);
    if (!(t < t_cap)) t = t_cap;
    double t_prev = 0.0, g_prev = 0.0, dg_prev = 0.0, step_prev_deg = INFINITY;
    bool have_prev = false;
    ts = NAN;
#pragma unroll 1
    for (int it = 0; it < AIRICE_NEWTON_MAXIT; it++) {
      double sq1, w;
      AIRICE_SQRT_RSQRT(1.0 + t * t, sq1, w);
      w = AIRICE_RCP(sq1);
      const double L = n_tx * t * w;
      double dXdL;
      const double X = airice_x_newton(m, p, kt, h, n_tx, L, dXdL);
      st.n_newton++;
      const double g = X - d;
      if (g == 0.0) { ts = t; break; }
      if (g < 0.0) {
        tlo = t;
        if (t >= t_cap) { ts = INFINITY; break; }  // even the lower bracket end falls short of d
      } else {
        thi = t;
      }
      const double dgdt = dXdL * n_tx * w * w * w;
      const double inv_dg = AIRICE_RCP(dgdt);
      double tn = t - g * inv_dg;
      const double hi_t = thi < t_cap ? thi : t_cap;
      bool newton_step = true;
      if (!(tn > tlo) || !(tn < hi_t)) {
        // outside the bracket (or NaN): probe the cap once if it is still untested, else bisect
        newton_step = false;
        if (thi == INFINITY && g < 0.0) tn = t_cap; else tn = 0.5 * (tlo + hi_t);
      }
      const double step_deg = fabs(tn - t) * w * w * m.rad2deg;
      if (newton_step && step_deg < 1.0e-7) { ts = tn; break; }
      if (newton_step && have_prev && step_deg * step_prev_deg < AIRICE_HERMITE_ACCEPT) {
        // inverse Hermite through (g_prev, t_prev, 1/g'_prev) and (g, t, 1/g'), evaluated at g = 0
        const double dy = g - g_prev;
        const double sfrac = -g_prev * AIRICE_RCP(dy);
        const double s2 = sfrac * sfrac, s3 = s2 * sfrac;
        const double h00 = 2.0 * s3 - 3.0 * s2 + 1.0, h10 = s3 - 2.0 * s2 + sfrac, h01 = 3.0 * s2 - 2.0 * s3, h11 = s3 - s2;
        const double th_ = h00 * t_prev + h10 * dy * AIRICE_RCP(dg_prev) + h01 * t + h11 * dy * inv_dg;
        if (th_ > tlo && th_ < hi_t) { ts = th_; break; }
      }
      have_prev = newton_step;
      t_prev = t; g_prev = g; dg_prev = dgdt; step_prev_deg = step_deg;
      t = tn;
      if (thi < INFINITY && thi - tlo <= 4.0e-16 * thi) { ts = 0.5 * (tlo + thi); break; }
    }
    if (ts != ts) ts = t;  // iteration cap: take the last iterate
  }
  // theta* in degrees; +inf in t (root beyond the lower end) maps below the bracket, -inf above it
  if (ts == INFINITY) theta_star = -INFINITY;
  else if (ts == -INFINITY) theta_star = INFINITY;
  else theta_star = 180 - atan(ts) * m.rad2deg;

  // ---- phase 2: replay of gsl_root_fsolver_bisection + gsl_root_test_interval (M.cc:355-369).
  // Fast form first: as long as no probe (lo, hi or a midpoint) comes within `guard` of theta*, sign f(x) is just the
  // side of theta* that x lies on, and GSL's update rule "keep the half whose ends differ in sign" becomes
  //   below = mid < theta*;  take_hi = (lo_below != below);  hi = take_hi ? mid : hi;  lo = take_hi ? lo : mid.
  // The returned root is the midpoint of the final bracket (GSL: root = 0.5*(lo+mid) or 0.5*(mid+hi)).
  const double guard = AIRICE_SOLVE_GUARD_DEG;
  const double th = theta_star;
  {
    double flo = lo, fhi = hi;
    bool near = !(fabs(flo - th) > guard) || !(fabs(fhi - th) > guard);
    bool lo_below = flo < th;
#pragma unroll 1
    for (int iter = 0; iter < 40; iter++) {
      const double mid = (flo + fhi) / 2.0;
      near = near || !(fabs(mid - th) > guard);
      const bool below = mid < th;
      const bool take_hi = (lo_below != below);
      fhi = take_hi ? mid : fhi;
      flo = take_hi ? flo : mid;
      lo_below = take_hi ? lo_below : below;
      if (fhi - flo < 0.000000001 * flo) break;   // gsl_root_test_interval with 0 < lo < hi
    }
    if (!near && flo > 0.0) return 0.5 * (flo + fhi);
  }
  // Careful form (about 0.2 % of solves): a probe sits within `guard` of theta*, so f is evaluated there for real
  // (MinimizeforLaunchAngle, M.cc:873-917), including GSL's exact-zero exits.  One loop serves the two endpoint
  // signs (steps -2, -1: gsl_root_fsolver_set) and the halvings (steps >= 0).
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

// Straight-line angle of GetHorizontalDistanceToIntersectionPoint (M.cc:952-958), metres in.  ta = its tangent.
AIRICE_HD double airice_straight_angle(const AirIceMedium& m, double h, double d, double ice, double depth_signed,
                                       double& ta) {
  const double den = (depth_signed < 0) ? (h - ice - depth_signed) : (h - (ice + depth_signed));
  ta = d / den;
  return 180 - (atan(ta) * (180.0 / m.pi));
}

// Solution flag of M.cc:974-983.
AIRICE_HD bool airice_check_solution(double thd, double d) {
  bool ok = false;
  if ((fabs(thd - d) / d < 0.01 && d <= 100) || (fabs(thd - d) < 1 && d > 100)) ok = true;
  if (thd < 0) ok = false;
  return ok;
}
