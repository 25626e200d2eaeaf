// airice_inice.cuh -- in-ice ray solver (Tx and Rx both in the ice): direct, reflected and up to two refracted
// rays per Tx->Rx pair, one thread per pair.
//
// Reference: /root/reference/IceRayTracing.cc -- IceRayTracing() :1745-1919, GetDirectRayPar :626-742,
// GetReflectedRayPar :745-920, GetRefractedRayPar :923-1253, root functions fDa/fRa/fRaa :411-607, fDnfR_L :368-379,
// GetZmax :346-353, FindFunctionRoot (GSL falsepos, |f|<1e-6, <=100 it) :261-300, FindFunctionRootZmax :303-335,
// FindFunctionRootFDF (GSL newton + gsl_deriv_central) :222-258.  TransitionBoundary == 0 (IceRayTracing.hh:49).
//
// Unlike the air->ice solve, nothing here can be short-cut: the reference stops its false-position iteration as
// soon as |f| < 1e-6 m, so the L it returns (and with it every launch angle to ~1e-6 deg, and the accept/reject
// decision |f|<0.5 that defines the solution-branch count) is a property of the ITERATION, not of the root.  The
// kernel therefore runs the same recurrences -- regula falsi with GSL's bisection safeguard, the nested falsepos for
// the turning depth z_max(L), Newton with the 5-point numerical derivative, and the retry ladder for the second
// refracted ray -- on registers.  What changes is the cost of each function evaluation: n(z0), n(z1), n(1e-7) are
// computed once per pair instead of 2-3 exp() per evaluation, and heap/indirection is gone.
#pragma once
#include "airice_core.cuh"
#include "airice_glibc_math.cuh"

// exp / log / pow of every quantity the reference ITERATES on round like the glibc the reference links (see
// airice_glibc_math.cuh): the stopping iterate, hence L, hence the branch flags, then equal the x86 build's bit for bit.
#define INICE_EXP(x) airice_glibc_exp(x)
#define INICE_LOG(x) airice_glibc_log(x)
#define INICE_POW(x, y) airice_glibc_pow((x), (y))

// Building blocks that pass 1 / pass 3 use many times over (fL ~30x, the 8-point derivative 3x, time+path 5x): real calls
// on the device, so that the kernel (15k instructions when everything is inlined, more than the instruction cache
// holds: measured 12 "no instruction" stall cycles per issue) is emitted once per block.
#if defined(__CUDACC__)
#define AIRICE_INICE_CALL __host__ __device__ __noinline__
#else
#define AIRICE_INICE_CALL inline
#endif

struct AirIceInIce {      // ice model + constants of the IceRayTracing namespace
  double A, B, C;         // IceRayTracing.hh:45-56
  double pi;              // 3.14159265359 (IceRayTracing.hh:41, sic)
  double c;               // 299792458
  double sin64;           // sin(64.0 * (pi / 180.0)) of IceRayTracing.cc:979 -- a constant; kept as one so that the device
                          // does not round it with its own sin()
};
AIRICE_HD AirIceInIce inice_make_model(double A, double B, double C) {
  AirIceInIce m;
  m.A = A; m.B = B; m.C = C;
  m.pi = 3.14159265359; m.c = 299792458.0;      // IceRayTracing.hh:41-43
  m.sin64 = 0x1.cc2ebbb5639ecp-1;               // glibc's sin() and GCC's constant folder agree on it (tests/test_glibc_math.py)
  return m;
}

struct InIcePair {
  double A, B, C, z0, z1, x1;   // z0 <= z1 after the flip (Tx deeper), both negative
  double n0, n1, ns;            // n(z0), n(z1), n(1e-7)
};

AIRICE_HD double inice_nz(const AirIceInIce& m, double z) { z = fabs(z); return m.A + m.B * INICE_EXP(-m.C * z); }

// fDnfR_L (IceRayTracing.cc:368-379) with n(Z) supplied
AIRICE_INICE_CALL double inice_fL(double A, double L, double Cp, double Z, double nZ) {
  return (L / Cp) * (1.0 / sqrt(A * A - L * L)) * (Cp * Z - INICE_LOG(A * nZ - L * L + sqrt(A * A - L * L) * sqrt(nZ * nZ - L * L)));
}
// The same in two parts for the root functions, which evaluate it at two or three depths for ONE L: the factor
// (L/C)(1/sqrt(A^2-L^2)) and sqrt(A^2-L^2) depend on L only (two divisions and a square root of the two divisions,
// two square roots and one log a call costs).  The product is formed in the order of the expression above
// (left to right), so the value of every term is bit-identical to inice_fL's.
struct InIceFLPre { double t, sA, L2; };
AIRICE_HD InIceFLPre inice_fL_pre(double A, double L, double Cp) {
  InIceFLPre q;
  q.sA = sqrt(A * A - L * L);
  q.t = (L / Cp) * (1.0 / q.sA);
  q.L2 = L * L;
  return q;
}
AIRICE_INICE_CALL double inice_fL_at(const InIceFLPre& q, double A, double Cp, double Z, double nZ) {
  return q.t * (Cp * Z - INICE_LOG(A * nZ - q.L2 + q.sA * sqrt(nZ * nZ - q.L2)));
}

// ---- GSL pieces, restated (see oracle/gsl_standin/gsl_standin.c for the same algorithms on the test side)
struct InIceBracket {
  double f_lower, f_upper, root, x_lower, x_upper;
  double f_root;     // f(root) when the iterate has just evaluated it (root == x_linear), else NaN
};

template <class F>
AIRICE_HD void inice_falsepos_set(const F& f, InIceBracket& s, double lo, double hi) {
  s.f_lower = 0.0; s.f_upper = 0.0;            // zero-initialised solver state
  s.root = 0.5 * (lo + hi); s.x_lower = lo; s.x_upper = hi;
  const double fl = f(lo);
  if (!isfinite(fl)) return;                    // GSL_EBADFUNC: state left as it was
  const double fu = f(hi);
  if (!isfinite(fu)) return;
  s.f_lower = fl; s.f_upper = fu;
}

template <class F>
AIRICE_HD void inice_falsepos_iterate(const F& f, InIceBracket& s) {
  const double xl = s.x_lower, xr = s.x_upper, fl = s.f_lower, fu = s.f_upper;
  s.f_root = NAN;
  if (fl == 0.0) { s.root = xl; s.x_upper = xl; return; }
  if (fu == 0.0) { s.root = xr; s.x_lower = xr; return; }
  const double x_lin = xr - (fu * (xl - xr) / (fl - fu));
  const double f_lin = f(x_lin);
  if (!isfinite(f_lin)) return;
  s.f_root = f_lin;
  if (f_lin == 0.0) { s.root = x_lin; s.x_lower = x_lin; s.x_upper = x_lin; return; }
  double w;
  if ((fl > 0.0 && f_lin < 0.0) || (fl < 0.0 && f_lin > 0.0)) { s.root = x_lin; s.x_upper = x_lin; s.f_upper = f_lin; w = x_lin - xl; }
  else { s.root = x_lin; s.x_lower = x_lin; s.f_lower = f_lin; w = xr - x_lin; }
  if (w < 0.5 * (xr - xl)) return;
  const double xb = 0.5 * (xl + xr);
  const double fb = f(xb);
  if (!isfinite(fb)) return;
  if ((fl > 0.0 && fb < 0.0) || (fl < 0.0 && fb > 0.0)) {
    s.x_upper = xb; s.f_upper = fb;
    if (s.root > xb) { s.root = 0.5 * (xl + xb); s.f_root = NAN; }
  } else {
    s.x_lower = xb; s.f_lower = fb;
    if (s.root < xb) { s.root = 0.5 * (xb + xr); s.f_root = NAN; }
  }
}

// FindFunctionRoot (IceRayTracing.cc:261-300)
template <class F>
AIRICE_HD double inice_find_root(const F& f, double lo, double hi) {
  InIceBracket s;
  if (lo > hi) {            // gsl_root_fsolver_set refuses; solver has no bracket (zeroed): every iterate reports 0
    s.f_lower = 0; s.f_upper = 0; s.root = 0; s.x_lower = 0; s.x_upper = 0;
  } else {
    inice_falsepos_set(f, s, lo, hi);
  }
  double r = 0;
  s.f_root = NAN;
#pragma unroll 1
  for (int iter = 0; iter < 100; iter++) {
    const InIceBracket before = s;
    inice_falsepos_iterate(f, s);
    r = s.root;
    // the reference re-evaluates f(root) for its residual test (IceRayTracing.cc:285); when root is the regula-falsi
    // point the iterate has that very value already (f is deterministic)
    const double check = (s.f_root == s.f_root) ? s.f_root : f(r);
    if (fabs(check) < 1e-6) break;
    // a step that changed nothing (e.g. the regula-falsi point fell outside the domain, f not finite: GSL returns
    // before touching its state) will change nothing the next 99 times either
    if (s.root == before.root && s.x_lower == before.x_lower && s.x_upper == before.x_upper &&
        s.f_lower == before.f_lower && s.f_upper == before.f_upper)
      break;
  }
  return r;
}

// GetZmax (IceRayTracing.cc:346-353) = FindFunctionRootZmax(GetMinnz, 0, 5000) (IceRayTracing.cc:303-335)
struct InIceMinnz {
  double A, B, C, L;
  AIRICE_HD double operator()(double x) const { return A + B * INICE_EXP(-C * x) - L; }   // raw x, not |x| (IceRayTracing.cc:342)
};
AIRICE_HD double inice_zmax_literal(double A, double B, double C, double L) {
  InIceMinnz f = {A, B, C, L};
  InIceBracket s;
  inice_falsepos_set(f, s, 0.0, 5000.0);
  double r = 0;
#pragma unroll 1
  for (int iter = 0; iter < 100; iter++) {
    inice_falsepos_iterate(f, s);
    r = s.root;
    const double lo = s.x_lower, hi = s.x_upper;
    if (lo > hi) break;  // GSL_EINVAL != GSL_CONTINUE
    const double al = fabs(lo), au = fabs(hi);
    const double mn = ((lo > 0.0 && hi > 0.0) || (lo < 0.0 && hi < 0.0)) ? (al < au ? al : au) : 0.0;
    if (fabs(hi - lo) < 1e-6 + 1e-6 * mn) break;
  }
  return r;
}
// The same iteration written for the machine it runs on: this function is ~3/4 of all the work of the refracted-ray
// search (it sits inside every evaluation of fRaa) and a pure dependency chain, ~13 falsepos steps of two exp() each
// (or a single step when L is below n(0): the first regula-falsi point is negative and ends the search at once).
// The regula-falsi point and the bisection point of a step depend only on the bracket the step starts from, so both
// function values are computed together (the second one is needed in >90% of the steps and simply dropped otherwise),
// which halves the chain.  The iteration is a struct so that the GPU kernel can run it one step at a time with the
// lanes of a warp taking the next pending evaluation as soon as their own has converged.
// Same points, same values, same updates as inice_zmax_literal; tests compare the two.
struct InIceZmaxIter {
  double xl, xr, fl, fu, root, L;
  int iter;
  // e5000 = exp(-C * 5000.0)
  AIRICE_HD void init(double A, double B, double e5000, double L_) {
    L = L_; xl = 0.0; xr = 5000.0; root = 0.5 * (0.0 + 5000.0); fl = 0.0; fu = 0.0; iter = 0;
    const double f0 = A + B * 1.0 - L;          // exp(-C * 0.0) == 1
    if (isfinite(f0)) {
      const double f1 = A + B * e5000 - L;
      if (isfinite(f1)) { fl = f0; fu = f1; }
    }
  }
  // one falsepos iteration and the stopping tests of the loop around it; true = finished (root is the answer).
  // Written with selects on the five state values: as nested if / else the compiler copied the whole state into a second
  // register set on every path (ncu: 230 instructions per step, 36 of them FP64 arithmetic).  fl and fu are finite and,
  // on the main path, non-zero, so "opposite signs" is one comparison of two predicates.
  AIRICE_HD bool step(double A, double B, double C) {
    // Every rare case ends the search in this very step (a zero at a bracket end or at the regula-falsi point collapses the
    // bracket, which passes the interval test below; a non-finite f leaves the state as it is, and nothing would ever
    // change it): they return at once, so the loop around this function carries no second copy of the state for them.
    if (fl == 0.0 || fu == 0.0) {                    // GSL's early exits
      root = (fl == 0.0) ? xl : xr;
      xl = root; xr = root;
      return true;
    }
    const double oxl = xl, oxr = xr, ofl = fl, ofu = fu;
    const double x_lin = oxr - (ofu * (oxl - oxr) / (ofl - ofu));
    const double xb = 0.5 * (oxl + oxr);
    const double f_lin = A + B * INICE_EXP(-C * x_lin) - L;
    const double fb = A + B * INICE_EXP(-C * xb) - L;
    // f not finite at the regula-falsi point (exp overflow for L far below the physical range): GSL returns before
    // touching its state, so this and all the remaining iterations up to the 100th change nothing
    if (!isfinite(f_lin)) return true;
    if (f_lin == 0.0) { root = x_lin; xl = x_lin; xr = x_lin; return true; }
    const bool pos = ofl > 0.0;
    const bool opp = pos != (f_lin > 0.0);
    const double w = opp ? x_lin - oxl : oxr - x_lin;
    double nxl = opp ? oxl : x_lin, nfl = opp ? ofl : f_lin;
    double nxr = opp ? x_lin : oxr, nfu = opp ? f_lin : ofu;
    double nroot = x_lin;
    if (!(w < 0.5 * (oxr - oxl)) && isfinite(fb)) {
      const bool oppb = pos ? (fb < 0.0) : (fb > 0.0);
      const double ra = 0.5 * (oxl + xb), rb = 0.5 * (xb + oxr);
      nroot = oppb ? (x_lin > xb ? ra : x_lin) : (x_lin < xb ? rb : x_lin);
      nxr = oppb ? xb : nxr; nfu = oppb ? fb : nfu;
      nxl = oppb ? nxl : xb; nfl = oppb ? nfl : fb;
    }
    xl = nxl; xr = nxr; fl = nfl; fu = nfu; root = nroot;
    if (xl > xr) return true;  // GSL_EINVAL != GSL_CONTINUE
    // gsl_root_test_interval: min(|xl|, |xr|) when both have one sign, else 0 -- with xl <= xr from here on that is xl when
    // xl > 0, -xr when xr < 0, else 0; and |xr - xl| = xr - xl
    const double mn = xl > 0.0 ? xl : (xr < 0.0 ? -xr : 0.0);
    if (xr - xl < 1e-6 + 1e-6 * mn) return true;
    return ++iter >= 100;
  }
};
AIRICE_HD double inice_zmax(double A, double B, double C, double L) {
  InIceZmaxIter z;
  z.init(A, B, INICE_EXP(-C * 5000.0), L);
#pragma unroll 1
  while (!z.step(A, B, C)) {}
  return z.root;
}

// root functions of L (IceRayTracing.cc:411-607), TransitionBoundary == 0 branches
struct InIceFDa {
  InIcePair g;
  AIRICE_HD double operator()(double L) const {
    const InIceFLPre q = inice_fL_pre(g.A, L, g.C);
    return (inice_fL_at(q, g.A, g.C, g.z1, g.n1) - inice_fL_at(q, g.A, g.C, g.z0, g.n0)) - g.x1;
  }
};
struct InIceFRa {
  InIcePair g;
  AIRICE_HD double operator()(double L) const {
    const InIceFLPre q = inice_fL_pre(g.A, L, -g.C);
    const double fb = inice_fL_at(q, g.A, -g.C, -g.z0, g.n0);
    const double d01 = inice_fL_at(q, g.A, -g.C, -g.z1, g.n1) - fb;
    const double d0s = inice_fL_at(q, g.A, -g.C, 1e-7, g.ns) - fb;
    return d01 - 2 * (d0s) - g.x1;
  }
};
// zm: the turning depth (+1e-7) this evaluation used -- the same number GetRefractedRayPar computes again right after
// every root search (IceRayTracing.cc:957,985...), so callers that just evaluated f at the root can keep it
AIRICE_HD double inice_fraa_given_zmax(const InIcePair& g, double L, double zmax) {   // zmax = inice_zmax(L) + 1e-7
  if (!(zmax > 0)) return 1e9;
  const double nzm = g.A + g.B * INICE_EXP(-g.C * fabs(zmax));
  const InIceFLPre q = inice_fL_pre(g.A, L, -g.C);
  const double fb = inice_fL_at(q, g.A, -g.C, -g.z0, g.n0);
  double d01 = inice_fL_at(q, g.A, -g.C, -g.z1, g.n1) - fb;
  double d0s = inice_fL_at(q, g.A, -g.C, zmax, nzm) - fb;
  if (d01 != d01) d01 = 1e9;
  if (d0s != d0s) d0s = 1e9;
  return d01 - 2 * (d0s) - g.x1;
}
AIRICE_HD double inice_fraa_eval(const InIcePair& g, double L, double& zm) {
  zm = inice_zmax(g.A, g.B, g.C, L) + 1e-7;
  return inice_fraa_given_zmax(g, L, zm);
}
// fRaa outside the physical range of L without the iteration: for L = NaN (a Newton search that has left the domain
// keeps asking for it) the turning-depth search stops at its first step with root 0, for finite L > A at the first
// regula-falsi point, beyond 5000 m, where f < 0; in both cases every fL term is NaN and fRaa is its NaN penalty
// 1e9 - 2e9 - x1.  Returns false when L is not of that kind (nothing written).  Tests compare with inice_fraa_eval.
AIRICE_HD bool inice_fraa_shortcut(double A, double B, double e5000, double x1, double L, double& y, double& zm) {
  if (L != L) { y = 1e9 - 2 * (1e9) - x1; zm = 0.0 + 1e-7; return true; }
  if (L > A && L < 1e300) {
    const double f0 = A + B * 1.0 - L, f1 = A + B * e5000 - L;          // InIceZmaxIter::init; both < 0
    const double x_lin = 5000.0 - (f1 * (0.0 - 5000.0) / (f0 - f1));   // first step
    if (x_lin > 5000.0 && x_lin < 1e300) { y = 1e9 - 2 * (1e9) - x1; zm = x_lin + 1e-7; return true; }
  }
  return false;
}
struct InIceFRaa {
  InIcePair g;
  AIRICE_HD double operator()(double L) const { double zm; return inice_fraa_eval(g, L, zm); }
};

// Horizontal reach of a refracted ray with parameter L in exact form: X_Ra(L) = F(-z1) + F(-z0) - 2 F(z_max), with the
// turning depth z_max = -ln((L-A)/B)/C where n(z_max) = L (there R = 0 and T = L (A - L)).
AIRICE_HD double inice_xra_exact(const InIcePair& g, double L) {
  const double zmax = -log((L - g.A) / g.B) / g.C;
  const InIceFLPre q = inice_fL_pre(g.A, L, -g.C);
  const double fm = q.t * (-g.C * zmax - log(L * (g.A - L)));
  return inice_fL_at(q, g.A, -g.C, -g.z1, g.n1) + inice_fL_at(q, g.A, -g.C, -g.z0, g.n0) - 2.0 * fm;
}

// Certificate that NO refracted ray can be accepted: fRaa(L) = X_Ra(L) - x1 wherever it is not a 1e9 penalty, the
// reference's version differs from the exact one only through its ~1e-6 m error in z_max (worth < 1e-2 m in X), and a
// branch is accepted only if |fRaa| < 0.5 at the returned L (IceRayTracing.cc:1905-1916).  So if the exact X_Ra stays
// more than `margin` below x1 on the whole admissible range (A+B, min(n0,n1)], every path through the reference's
// falsepos / Newton retry ladder ends with the branch rejected, and the ladder (hundreds of evaluations, each with a
// nested falsepos for z_max: two thirds of the reference's total run time) need not be walked.
AIRICE_HD bool inice_no_refracted_possible(const InIcePair& g, double margin) {
  const double lo = g.A + g.B, hi = (g.n0 < g.n1 ? g.n0 : g.n1);
  if (!(hi > lo)) return false;
  // X_Ra is smooth with one interior maximum on (lo, hi); sample it, then polish around the best sample, and add a
  // curvature-based slack so that the bound holds between samples.
  const int N = 24;
  double best = -INFINITY, bestL = lo;
  double prev = -INFINITY, prev2 = -INFINITY, slack = 0.0;
#pragma unroll 1
  for (int i = 1; i <= N; i++) {
    const double L = lo + (hi - lo) * ((double)i - 0.5) / (double)N;
    const double x = inice_xra_exact(g, L);
    if (!(x == x)) return false;
    if (x > best) { best = x; bestL = L; }
    if (i >= 3) { const double c = fabs(x - 2.0 * prev + prev2); if (c > slack) slack = c; }
    prev2 = prev; prev = x;
  }
  // the maximum lies within one sample spacing of the best sample; |second difference| bounds how far X can rise there
  return best + 2.0 * slack + margin < g.x1 && bestL == bestL;
}

// gsl_deriv_central (deriv/deriv.c)
template <class F>
AIRICE_HD void inice_central(const F& f, double x, double h, double& result, double& round, double& trunc) {
  const double fm1 = f(x - h), fp1 = f(x + h), fmh = f(x - h / 2), fph = f(x + h / 2);
  const double r3 = 0.5 * (fp1 - fm1);
  const double r5 = (4.0 / 3.0) * (fph - fmh) - (1.0 / 3.0) * r3;
  const double e3 = (fabs(fp1) + fabs(fm1)) * 2.2204460492503131e-16;
  const double e5 = 2.0 * (fabs(fph) + fabs(fmh)) * 2.2204460492503131e-16 + e3;
  const double a = fabs(r3 / h), b = fabs(r5 / h);
  const double dy = (a > b ? a : b) * (fabs(x) / h) * 2.2204460492503131e-16;
  result = r5 / h;
  trunc = fabs((r5 - r3) / h);
  round = fabs(e5 / h) + dy;
}
template <class F>
AIRICE_INICE_CALL double inice_deriv_central(const F& f, double x, double h) {
  double r0, round, trunc;
  inice_central(f, x, h, r0, round, trunc);
  double error = round + trunc;
  if (round < trunc && (round > 0 && trunc > 0)) {
    double ro, round_o, trunc_o;
    const double h_opt = h * INICE_POW(round / (2.0 * trunc), 1.0 / 3.0);
    inice_central(f, x, h_opt, ro, round_o, trunc_o);
    const double error_o = round_o + trunc_o;
    if (error_o < error && fabs(ro - r0) < 4.0 * error) { r0 = ro; error = error_o; }
  }
  return r0;
}

// FindFunctionRootFDF (IceRayTracing.cc:222-258): GSL newton with df from gsl_deriv_central(h=1e-8)
template <class F>
AIRICE_HD double inice_newton_root(const F& f, double lo, double hi) {
  double x = (lo + hi) / 2;
  double fv = f(x), df = inice_deriv_central(f, x, 1e-8);
  double root = x;
#pragma unroll 1
  for (int iter = 0; iter < 100; iter++) {
    if (df != 0.0) {                       // else GSL_EZERODIV: root unchanged
      const double rn = root - (fv / df);
      root = rn;
      fv = f(rn);
      df = inice_deriv_central(f, rn, 1e-8);
    }
    const double x0 = x;
    x = root;
    if (fabs(x - x0) < 1e-6 * fabs(x) || x == x0) break;
  }
  return x;
}

// fDnfR as a function of depth (IceRayTracing.cc:355-365), n(x) evaluated per call like the reference
struct InIceFDepth {
  double A, B, C, Cp, L;
  AIRICE_HD double operator()(double x) const {
    const double n = A + B * INICE_EXP(-C * fabs(x));
    return (L / Cp) * (1.0 / sqrt(A * A - L * L)) * (Cp * x - INICE_LOG(A * n - L * L + sqrt(A * A - L * L) * sqrt(n * n - L * L)));
  }
};

// ftimeD / fpathD (IceRayTracing.cc:382-408) through the identities of airice_core.cuh
AIRICE_INICE_CALL void inice_time_path(const AirIceInIce& m, double x, double Cp, double L, double& t, double& p) {
  const double A = m.A;
  const double n = inice_nz(m, x);
  const double D = n * n - L * L, R = sqrt(D), sA = sqrt(A * A - L * L);
  const double G = Cp * x - INICE_LOG(A * n - L * L + sA * R), H = INICE_LOG(n + R);
  t = (1.0 / ((m.c * Cp) * R)) * ((D + (G * (A * A * R)) / sA) + (A * R) * H);
  p = (H + (A / sA) * G) / Cp;
}

// The solver is split in two so that the kernel can run the cheap, uniform part (direct + reflected ray) for every pair
// and the long, irregular part (the refracted-ray ladder) only for the pairs that need it, packed densely into warps.

AIRICE_HD InIcePair inice_make_pair(const AirIceInIce& m, double z0_in, double x1, double z1_in, bool& flip) {
  // the tracer wants the transmitter deeper than the receiver (IceRayTracing.cc:631-637)
  double z0 = z0_in, z1 = z1_in;
  flip = z0 > z1;
  if (flip) { z0 = z1_in; z1 = z0_in; }
  InIcePair g;
  g.A = m.A; g.B = m.B; g.C = m.C; g.z0 = z0; g.z1 = z1; g.x1 = x1;
  g.n0 = inice_nz(m, z0); g.n1 = inice_nz(m, z1); g.ns = inice_nz(m, 1e-7);
  return g;
}

// Direct + reflected ray (GetDirectRayPar :626-742, GetReflectedRayPar :745-920) and the IceRayTracing() bookkeeping for
// them; all 29 slots are written, the refracted ones as "absent".  needs_ra: the refracted ladder has to run
// (IceRayTracing.cc:1806) and is not ruled out by inice_no_refracted_possible.  Returns mask bits 0 (D) and 1 (R).
AIRICE_HD int inice_solve_dr(const AirIceInIce& m, double z0_in, double x1, double z1_in, double* out, bool& needs_ra) {
  const double k180pi = 180.0 / m.pi;
  bool flip;
  const InIcePair g = inice_make_pair(m, z0_in, x1, z1_in, flip);
  const double z0 = g.z0, z1 = g.z1;
  // ---------------- direct ray (IceRayTracing.cc:626-742)
  double RangD, LangD, timeD, pathD, lvalueD, checkD;
  {
    InIceFDa f = {g};
    const double up = g.n1 < g.n0 ? g.n1 : g.n0;   // min_element keeps the first of equals; values equal then
    lvalueD = inice_find_root(f, 1e-7, up);
    LangD = asin(lvalueD / g.n0) * k180pi;
    checkD = f(lvalueD);
    double ta, pa, tb, pb;
    inice_time_path(m, -z0, -m.C, lvalueD, ta, pa);
    inice_time_path(m, -z1, -m.C, lvalueD, tb, pb);
    timeD = ta - tb; pathD = pa - pb;
    InIceFDepth fd = {m.A, m.B, m.C, -m.C, lvalueD};
    RangD = atan(inice_deriv_central(fd, -z1, 1e-8)) * k180pi;
    if (z1 == z0 && RangD != RangD) RangD = 180 - LangD;
    if (checkD != checkD) checkD = -1000;
  }
  double outD0 = RangD, outD1 = LangD;
  if (flip) { outD0 = 180 - LangD; outD1 = 180 - RangD; }

  // ---------------- reflected ray (IceRayTracing.cc:745-920)
  double RangR, LangR, timeR, timeR1, timeR2, pathR, lvalueR, checkR, incAng;
  {
    InIceFRa f = {g};
    double up = g.n1;
    if (g.n0 < up) up = g.n0;
    if (g.ns < up) up = g.ns;
    lvalueR = inice_find_root(f, 1e-7, up);
    LangR = asin(lvalueR / g.n0) * k180pi;
    checkR = f(lvalueR);
    double ts, ps, ta, pa, tb, pb;
    inice_time_path(m, -1e-7, m.C, lvalueR, ts, ps);
    inice_time_path(m, z0, m.C, lvalueR, ta, pa);
    inice_time_path(m, z1, m.C, lvalueR, tb, pb);
    timeR1 = ts - ta; timeR2 = ts - tb;
    double pathR1 = ps - pa, pathR2 = ps - pb;
    timeR = timeR1 + timeR2; pathR = pathR1 + pathR2;
    if (flip) { const double d = timeR2; timeR2 = timeR1; timeR1 = d; }
    InIceFDepth fd = {m.A, m.B, m.C, m.C, lvalueR};
    RangR = 180 - atan(inice_deriv_central(fd, z1, 1e-8)) * k180pi;
    if (z1 == z0 && RangR != RangR) RangR = 180 - LangR;
    if (z1 != z0 && RangR != RangR) RangR = 90;
    if (checkR != checkR) checkR = -1000;
    incAng = atan(inice_deriv_central(fd, -1e-7, 1e-8)) * k180pi;
  }
  double outR0 = RangR, outR1 = LangR;
  if (flip) { outR0 = 180 - LangR; outR1 = 180 - RangR; }

  (void)LangR; (void)RangR;
  out[0] = outD1; out[1] = outR1; out[2] = 0; out[3] = 0;
  out[4] = timeD; out[5] = timeR; out[6] = 0; out[7] = 0;
  out[8] = outD0; out[9] = outR0; out[10] = -1000; out[11] = -1000;
  out[12] = 0; out[13] = 0; out[14] = 0; out[15] = 0; out[16] = 0; out[17] = 0;
  if (fabs(checkR) < 0.5) { out[12] = timeR1; out[13] = timeR2; }
  out[18] = incAng;
  out[19] = lvalueD; out[20] = lvalueR; out[21] = 0; out[22] = 0;
  out[23] = 0; out[24] = 0;
  out[25] = pathD; out[26] = pathR; out[27] = 0; out[28] = 0;
  int mask = 3;
  if (fabs(checkD) > 0.5) { out[8] = -1000; mask &= ~1; }
  if (fabs(checkR) > 0.5) { out[9] = -1000; mask &= ~2; }
  needs_ra = (mask != 3) && !inice_no_refracted_possible(g, 2.0);
  return mask;
}

// Refracted rays (GetRefractedRayPar :923-1253) for a pair whose direct and/or reflected ray is missing, in two parts:
// the root-search LADDER (falsepos, Newton retry, up to five more searches for a second root) and the FINISH (times,
// paths, angles, bookkeeping).  The ladder is the long irregular part; it exists twice: literally below, and as a
// resumable state machine in airice_inice_machine.cuh that the GPU kernel steps (same evaluations, same order).
struct InIceRaLadder { double lv[2], cz[2], zm[2]; };   // L, f(L) and z_max(L)+1e-7 of the two candidate roots

// bracket of the first search (IceRayTracing.cc:937-950); LangR_in: the reflected ray's launch angle as the callee sees it
AIRICE_HD void inice_ra_first_bracket(const AirIceInIce& m, const InIcePair& g, bool flip, double lvalueR, double& lower,
                                      double& up) {
  const double k180pi = 180.0 / m.pi, kpi180 = m.pi / 180.0;
  // the callee gets the reflected ray's launch angle back in the flipped frame (IceRayTracing.cc:937-941): that is
  // the internal LangR = asin(L_R / n(z0))
  double LangR_in = asin(lvalueR / g.n0) * k180pi;
  if (flip) LangR_in = 180 - (180 - LangR_in);   // it travels out as 180-LangR and is flipped back, with both roundings
  up = g.n0 < g.n1 ? g.n0 : g.n1;
  lower = g.n0 * m.sin64;
  if (lower > up) lower = g.n0 * sin((LangR_in * kpi180));
}

AIRICE_HD InIceRaLadder inice_ra_ladder(const AirIceInIce& m, const InIcePair& g, bool flip, bool d_absent, bool r_absent,
                                        double lvalueR) {
  InIceFRaa f = {g};
  double lv[2] = {0, 0}, cz[2] = {-1000, -1000}, zm[2] = {10, 10};
  double lower, up;
  inice_ra_first_bracket(m, g, flip, lvalueR, lower, up);
  lv[0] = inice_find_root(f, lower, up);
  cz[0] = f(lv[0]);
  zm[0] = inice_zmax(m.A, m.B, m.C, lv[0]) + 1e-7;
  if (fabs(cz[0]) > 0.5) {
    lv[0] = inice_newton_root(f, lower, up);
    cz[0] = f(lv[0]);
    zm[0] = inice_zmax(m.A, m.B, m.C, lv[0]) + 1e-7;
  }
  if (lv[0] < 0) cz[0] = -1000;
#define INICE_RETRY(expr_root)                                   \
  do {                                                           \
    lv[1] = (expr_root);                                         \
    cz[1] = f(lv[1]);                                            \
    zm[1] = inice_zmax(m.A, m.B, m.C, lv[1]) + 1e-7;             \
  } while (0)
#define INICE_BAD1 (fabs(cz[1]) > 0.5 || cz[1] != cz[1] || fabs(lv[1] - lv[0]) < 1e-4)
  if (fabs(cz[0]) < 0.5 && d_absent && r_absent) {
    INICE_RETRY(inice_find_root(f, lv[0] - 0.23, lv[0] - 0.023));
    if (INICE_BAD1) INICE_RETRY(inice_find_root(f, lv[0] - 0.15, lv[0] - 0.023));
    if (INICE_BAD1) {
      if (lv[0] + 0.005 < up) INICE_RETRY(inice_find_root(f, lv[0] + 0.005, up));
      else INICE_RETRY(inice_find_root(f, lv[0] - 0.1, lv[0] - 0.01));
    }
    if (INICE_BAD1) {
      const double tmp = inice_newton_root(f, lv[0] - 0.23, lv[0] - 0.023);
      if (fabs(tmp) < m.A) INICE_RETRY(tmp);   // the reference solves the same problem twice (IceRayTracing.cc:1029-1031)
    }
    if (INICE_BAD1) {
      const double tmp = inice_newton_root(f, lv[0] - 0.1, lv[0] - 0.023);
      if (fabs(tmp) < m.A) INICE_RETRY(tmp);
    }
    if (lv[1] < 0) cz[1] = -1000;
    if (fabs(cz[1]) < 0.5 && fabs(cz[0]) < 0.5 && fabs(lv[1] - lv[0]) < 1e-4) cz[1] = -1000;
  } else {
    lv[1] = 0; cz[1] = -1000; zm[1] = -1000;
  }
#undef INICE_RETRY
#undef INICE_BAD1
  InIceRaLadder r;
  r.lv[0] = lv[0]; r.lv[1] = lv[1]; r.cz[0] = cz[0]; r.cz[1] = cz[1]; r.zm[0] = zm[0]; r.zm[1] = zm[1];
  return r;
}

// fills the refracted slots of out[] (2,3,6,7,10,11,14-17,21-24,27,28); returns mask bits 2 (Ra1) and 3 (Ra2)
AIRICE_HD int inice_ra_finish(const AirIceInIce& m, const InIcePair& g, bool flip, bool d_absent, bool r_absent,
                              const InIceRaLadder& lad, double* out) {
  const double k180pi = 180.0 / m.pi;
  const double z0 = g.z0, z1 = g.z1;
  const bool both = d_absent && r_absent;
  double lv[2] = {lad.lv[0], lad.lv[1]}, cz[2] = {lad.cz[0], lad.cz[1]}, zm[2] = {lad.zm[0], lad.zm[1]};
  double La[2] = {asin(lv[0] / g.n0) * k180pi, asin(lv[1] / g.n0) * k180pi};
  if (fabs(cz[0]) < 0.5 && both) {     // the second root was searched for (cz[0] does not change in that branch)
    if (La[0] != La[0]) La[0] = 0;
    if (La[1] != La[1]) La[1] = 0;
    if (La[1] < La[0] && fabs(cz[0]) < 0.5 && fabs(cz[1]) < 0.5) {
      double t;
      t = lv[1]; lv[1] = lv[0]; lv[0] = t;
      t = La[1]; La[1] = La[0]; La[0] = t;
      t = cz[1]; cz[1] = cz[0]; cz[0] = t;
      t = zm[1]; zm[1] = zm[0]; zm[0] = t;
    }
  } else {
    La[1] = 0;
  }
  double RangRa[2] = {0, 0}, LangRa[2] = {0, 0}, timeRa[2] = {0, 0}, lvalueRa[2] = {0, 0}, checkRa[2] = {-1000, -1000};
  double timeRa1[2] = {0, 0}, timeRa2[2] = {0, 0}, zmaxv[2] = {0, 0}, pathRa[2] = {0, 0};
#pragma unroll 1
  for (int i = 0; i < 2; i++) {
    double tRa = 0, tRa1 = 0, tRa2 = 0, pRa = 0;
    if (cz[i] != cz[i]) cz[i] = -1000;
    if (zm[i] == 1e-7 || zm[i] <= 0) cz[i] = -1000;
    if ((z0 < -zm[i] || zm[i] < -z1)) {
      double tm, pm, ta, pa, tb, pb;
      inice_time_path(m, -zm[i], m.C, lv[i], tm, pm);
      inice_time_path(m, z0, m.C, lv[i], ta, pa);
      inice_time_path(m, z1, m.C, lv[i], tb, pb);
      tRa1 = tm - ta; tRa2 = tm - tb;
      tRa = tRa1 + tRa2;
      pRa = (pm - pa) + (pm - pb);
      if (flip) { const double d = tRa2; tRa2 = tRa1; tRa1 = d; }
    }
    InIceFDepth fd = {m.A, m.B, m.C, m.C, lv[i]};
    double Ra = 180 - atan(inice_deriv_central(fd, z1, 1e-8)) * k180pi;
    if (z1 == z0 && Ra != Ra) Ra = 180 - La[i];
    if (z1 != z0 && Ra != Ra) Ra = 90;
    // callee outputs (IceRayTracing.cc:1215-1250): angles un-flipped
    double o0 = Ra, o1 = La[i];
    if (flip) { o0 = 180 - La[i]; o1 = 180 - Ra; }
    const bool take = (i == 0) || both;  // IceRayTracing.cc:1816
    if (take) {
      RangRa[i] = o0; LangRa[i] = o1; timeRa[i] = tRa; lvalueRa[i] = lv[i]; checkRa[i] = cz[i];
      timeRa1[i] = tRa1; timeRa2[i] = tRa2; zmaxv[i] = zm[i];
    }
    pathRa[i] = pRa;
  }
  out[2] = LangRa[0]; out[3] = LangRa[1];
  out[6] = timeRa[0]; out[7] = timeRa[1];
  out[10] = RangRa[0]; out[11] = RangRa[1];
  out[14] = 0; out[15] = 0; out[16] = 0; out[17] = 0;
  if (fabs(checkRa[0]) < 0.5) { out[14] = timeRa1[0]; out[15] = timeRa2[0]; }
  if (fabs(checkRa[1]) < 0.5) { out[16] = timeRa1[1]; out[17] = timeRa2[1]; }
  out[21] = lvalueRa[0]; out[22] = lvalueRa[1];
  out[23] = zmaxv[0]; out[24] = zmaxv[1];
  out[27] = pathRa[0]; out[28] = pathRa[1];
  int mask = 12;
  if (fabs(checkRa[0]) > 0.5) { out[10] = -1000; mask &= ~4; }
  if (fabs(checkRa[1]) > 0.5) { out[11] = -1000; mask &= ~8; }
  return mask;
}

// d_absent / r_absent: |checkzero| > 0.5 of the direct / reflected ray; lvalueR: the reflected ray's L (slot 20), from
// which the callee's LangR is re-derived exactly.
AIRICE_HD int inice_solve_ra(const AirIceInIce& m, double z0_in, double x1, double z1_in, bool d_absent, bool r_absent,
                             double lvalueR, double* out) {
  bool flip;
  const InIcePair g = inice_make_pair(m, z0_in, x1, z1_in, flip);
  const InIceRaLadder lad = inice_ra_ladder(m, g, flip, d_absent, r_absent, lvalueR);
  return inice_ra_finish(m, g, flip, d_absent, r_absent, lad, out);
}

// IceRayTracing::IceRayTracing(0, z0, x1, z1) -> out[29] (IceRayTracing.cc:1745-1919).  Slots 12..17 are written only
// when the branch exists in the reference; here absent ones are 0.  Returns the 4-bit branch mask (D,R,Ra1,Ra2).
AIRICE_HD int inice_solve(const AirIceInIce& m, double z0_in, double x1, double z1_in, double* out) {
  bool needs_ra;
  int mask = inice_solve_dr(m, z0_in, x1, z1_in, out, needs_ra);
  if (needs_ra) mask |= inice_solve_ra(m, z0_in, x1, z1_in, (mask & 1) == 0, (mask & 2) == 0, out[20], out);
  return mask;
}

// IceRayTracing::GetRayTracingSolutions (IceRayTracing.cc:2907-3210) without its attenuation integrals: picks the two
// physical rays of a pair out of the four candidates D, R, Ra1, Ra2 of IceRayTracing() (o = its 29 slots), orders them
// by arrival time and patches the same-depth straight-line case.  res[10] = TimeRay[2], PathRay[2], LaunchAngle[2],
// RecieveAngle[2], IncidenceAngleInIce[2]; ignore[2] = IgnoreCh (1 = ray present); type[2] = the reference's internal
// RayType (1 D, 2 R, 3 Ra1, 4 Ra2).  The assignments are a cascade in which later cases overwrite earlier ones; the
// order below is the reference's.
// att4 / att_out: optional AttD, AttR, AttRa[0..1] (1 - attenuation of the four candidates) and the AttRay[2] they become;
// they ride through the same cascade and swap (IceRayTracing.cc:3014-3148)
AIRICE_HD void inice_pick_two_rays(const AirIceInIce& m, const double* o, double rx_depth, double distance, double tx_depth,
                                   double* res, int* ignore, int* type, const double* att4 = nullptr, double* att_out = nullptr) {
  const double timeD = o[4], timeR = o[5], timeRa0 = o[6], timeRa1 = o[7];
  const double pathD = o[25], pathR = o[26], pathRa0 = o[27], pathRa1 = o[28];
  const double RangD = o[8], RangR = o[9], RangRa0 = o[10], RangRa1 = o[11];
  const double LangD = o[0], LangR = o[1], LangRa0 = o[2], LangRa1 = o[3];
  double T[2] = {timeD, timeR}, P[2] = {pathD, pathR}, Rv[2] = {RangD, RangR}, La[2] = {LangD, LangR};
  const double AttD = att4 ? att4[0] : 0, AttR = att4 ? att4[1] : 0, AttRa0 = att4 ? att4[2] : 0, AttRa1 = att4 ? att4[3] : 0;
  double At[2] = {AttD, AttR};
  int ty[2] = {1, 2};
  double inc[2] = {100, o[18]};
  if (RangR == -1000) { inc[0] = 100; inc[1] = 100; }
#define INICE_SET(slot, t, p, r, l, k, at) do { T[slot] = t; P[slot] = p; Rv[slot] = r; La[slot] = l; ty[slot] = k; At[slot] = at; } while (0)
  if (RangD != -1000) INICE_SET(0, timeD, pathD, RangD, LangD, 1, AttD);
  if (RangR != -1000) INICE_SET(1, timeR, pathR, RangR, LangR, 2, AttR);
  if (RangRa0 != -1000 && RangD != -1000) { INICE_SET(0, timeD, pathD, RangD, LangD, 1, AttD); INICE_SET(1, timeRa0, pathRa0, RangRa0, LangRa0, 3, AttRa0); }
  if (RangRa0 != -1000 && RangR != -1000) { INICE_SET(1, timeR, pathR, RangR, LangR, 2, AttR); INICE_SET(0, timeRa0, pathRa0, RangRa0, LangRa0, 3, AttRa0); }
  if (RangRa1 != -1000 && RangD != -1000) { INICE_SET(0, timeD, pathD, RangD, LangD, 1, AttD); INICE_SET(1, timeRa1, pathRa1, RangRa1, LangRa1, 4, AttRa1); }
  if (RangRa1 != -1000 && RangR != -1000) { INICE_SET(1, timeR, pathR, RangR, LangR, 2, AttR); INICE_SET(0, timeRa1, pathRa1, RangRa1, LangRa1, 4, AttRa1); }
  if (RangRa1 != -1000 && RangRa0 != -1000) { INICE_SET(1, timeRa1, pathRa1, RangRa1, LangRa1, 4, AttRa1); INICE_SET(0, timeRa0, pathRa0, RangRa0, LangRa0, 3, AttRa0); }
  if (Rv[1] == -1000 && Rv[0] == -1000 && RangRa0 != -1000) INICE_SET(0, timeRa0, pathRa0, RangRa0, LangRa0, 3, AttRa0);
  if (Rv[1] == -1000 && Rv[0] == -1000 && RangRa1 != -1000) INICE_SET(1, timeRa1, pathRa1, RangRa1, LangRa1, 4, AttRa1);
#undef INICE_SET
  int ig[2] = {1, 1};
  if (Rv[0] == -1000) ig[0] = 0;
  if (Rv[1] == -1000) ig[1] = 0;
  if (T[0] > T[1] && Rv[0] != -1000 && Rv[1] != -1000) {
    double d;
    d = La[0]; La[0] = La[1]; La[1] = d;
    d = Rv[0]; Rv[0] = Rv[1]; Rv[1] = d;
    d = T[0]; T[0] = T[1]; T[1] = d;
    d = P[0]; P[0] = P[1]; P[1] = d;
    d = At[0]; At[0] = At[1]; At[1] = d;
    const int k = ty[0]; ty[0] = ty[1]; ty[1] = k;
  }
  if (rx_depth == tx_depth && T[0] == 0 && P[0] == 0) {     // IceRayTracing.cc:3190-3200
    if (distance == 0) { ig[0] = 0; ig[1] = 0; }
    P[0] = distance;
    T[0] = distance / (m.c / inice_nz(m, tx_depth));
    La[0] = 90.; Rv[0] = 90.;
    ig[0] = 1;
  }
  res[0] = T[0]; res[1] = T[1]; res[2] = P[0]; res[3] = P[1]; res[4] = La[0]; res[5] = La[1]; res[6] = Rv[0]; res[7] = Rv[1];
  res[8] = inc[0]; res[9] = inc[1];
  ignore[0] = ig[0]; ignore[1] = ig[1];
  type[0] = ty[0]; type[1] = ty[1];
  if (att_out) { att_out[0] = At[0]; att_out[1] = At[1]; }
}
