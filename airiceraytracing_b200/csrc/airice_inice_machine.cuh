// airice_inice_machine.cuh -- the refracted-ray root-search ladder of airice_inice.cuh as a resumable state machine.
//
// inice_ra_ladder() is a ladder of up to seven root searches (GSL falsepos / Newton with a numerical derivative), each
// a loop of a data-dependent number of evaluations of ONE function, fRaa(L) (which itself hides a ~13-step falsepos for
// the turning depth).  A pair needs between 3 and several thousand evaluations (median ~40, mean ~90).  Run as nested
// loops, the 32 lanes of a warp sit in different searches at different iterations and execute one after the other
// (measured: 5.6 of 32 lanes active).  Here the nesting is turned inside out: a lane holds the whole search state of
// its pair in this struct and REQUESTS evaluations; advance() consumes the values and moves to the next request.
// The requests of one step are the evaluations the literal code makes next whose arguments do not depend on each
// other: both ends of a new bracket, the regula-falsi point together with the bisection point, a Newton iterate
// together with the four samples of its central-difference derivative.  The GPU kernel pools the requests of the 32
// lanes of a warp and spreads them over the lanes, so the common loop body is a single evaluation of fRaa -- identical
// code whatever search the owning pair is in -- and a pair in a long Newton search borrows the lanes of pairs that
// are finished.
//
// Every evaluation point, comparison and update is the one inice_find_root / inice_newton_root / inice_deriv_central
// make (an evaluation requested early and then not needed is dropped; f is a pure function, so evaluating it early
// changes nothing); tests/test_inice.py runs both forms on the host over the same pairs and requires identical bits.
// Two evaluations of the literal form are not repeated because their values are already at hand: f(root) after a
// search = the last residual check of the falsepos loop, or the Newton iterate's own f value.
#pragma once
#include "airice_inice.cuh"

namespace airice {

struct InIceRaMachine {
  static constexpr int kMaxReq = 5;
  enum Phase : int {
    // states that wait for evaluations of fRaa
    FP_SET, FP_LINBIS, FP_CHK, NW_FD, NW_D2,
    // internal states (no evaluation pending)
    FP_ITER, FP_POST, FP_TEST, NW_AFTER_D, NW_STEP, NW_TEST, LADDER, AFTER_FIRST, END_SECOND, DONE
  };
  double xq;                                   // first request of the step (see query())
  double A, lower, up;                         // ice A, first bracket
  double lv0, cz0, zm0, lv1, cz1, zm1;         // ladder results
  union {
    struct { double xl, xr, fl, fu, root, oxl, oxr, ofl, ofu, oroot, froot, zroot; } fp;   // falsepos search
    struct { double x, root, fv, zfv, df, h, r0, err0; } nw;                               // Newton search
  };
  int ph, iter, stage, nq;
  bool both_absent;
  int n_eval;                                  // evaluations requested (diagnostics)

  AIRICE_HD bool done() const { return ph == DONE; }

  // k-th evaluation point of the pending step, 0 <= k < nq
  AIRICE_HD double query(int k) const {
    if (ph == FP_SET) return k == 0 ? xq : fp.xr;
    if (ph == FP_LINBIS) return k == 0 ? xq : 0.5 * (fp.oxl + fp.oxr);
    if (ph == FP_CHK) return xq;
    // Newton: the iterate itself (NW_FD only), then the samples of inice_central around it
    const int d = (ph == NW_FD) ? k - 1 : k;
    const double c = nw.root, h = nw.h;
    switch (d) {
      case -1: return c;
      case 0: return c - h;
      case 1: return c + h;
      case 2: return c - h / 2;
      default: return c + h / 2;
    }
  }

  AIRICE_HD void fp_begin(double lo, double hi) {
    iter = 0;
    if (lo > hi) {      // gsl_root_fsolver_set refuses: zeroed solver state
      fp.fl = 0; fp.fu = 0; fp.root = 0; fp.xl = 0; fp.xr = 0;
      ph = FP_ITER; nq = 0;
      return;
    }
    fp.fl = 0; fp.fu = 0; fp.root = 0.5 * (lo + hi); fp.xl = lo; fp.xr = hi;
    xq = lo; ph = FP_SET; nq = 2;
  }
  AIRICE_HD void nw_begin(double lo, double hi) {
    nw.x = (lo + hi) / 2;
    nw.root = nw.x; nw.h = 1e-8; iter = -1;    // iter -1: the evaluation before the loop
    xq = nw.x; ph = NW_FD; nq = 5;
  }

  AIRICE_HD void init(const AirIceInIce& m, const InIcePair& g, bool flip, bool d_absent, bool r_absent, double lvalueR) {
    A = m.A;
    inice_ra_first_bracket(m, g, flip, lvalueR, lower, up);
    both_absent = d_absent && r_absent;
    lv0 = 0; cz0 = -1000; zm0 = 10; lv1 = 0; cz1 = -1000; zm1 = 10;
    stage = 0; n_eval = 0; nq = 0;
    fp_begin(lower, up);
    if (ph >= FP_ITER) { const double none[kMaxReq] = {0, 0, 0, 0, 0}; run(none, none); }
    n_eval += nq;
  }

  // consume y[k] = fRaa(query(k)) and zm[k] = the turning depth that evaluation used, k < nq; leaves the next requests
  // (nq, query()) or done()
  AIRICE_HD void advance(const double* y, const double* zm) { run(y, zm); n_eval += nq; }

  AIRICE_HD bool bad1() const { return fabs(cz1) > 0.5 || cz1 != cz1 || fabs(lv1 - lv0) < 1e-4; }

  // pow(round / (2 trunc), 1/3) of inice_deriv_central: a real call (pow is ~300 instructions)
  AIRICE_INICE_CALL static double cube_root_ratio(double round, double trunc) { return INICE_POW(round / (2.0 * trunc), 1.0 / 3.0); }

  // inice_central on samples f(c-h), f(c+h), f(c-h/2), f(c+h/2)
  AIRICE_INICE_CALL static void central(double x, double h, double fm1, double fp1, double fmh, double fph, double& result,
                                        double& round, double& trunc) {
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

  AIRICE_HD void run(const double* y, const double* zm) {
    double r_root = 0, r_cz = 0, r_zm = 0;     // result of a finished search, handed to LADDER
    double chk = 0, chk_zm = 0;
    nq = 0;
    for (;;) {
      switch (ph) {
        // ---------------- falsepos: inice_falsepos_set
        case FP_SET:
          if (isfinite(y[0]) && isfinite(y[1])) { fp.fl = y[0]; fp.fu = y[1]; }
          ph = FP_ITER;
          break;
        // ---------------- inice_falsepos_iterate
        case FP_ITER: {
          fp.oxl = fp.xl; fp.oxr = fp.xr; fp.ofl = fp.fl; fp.ofu = fp.fu; fp.oroot = fp.root;
          fp.froot = NAN;
          if (fp.fl == 0.0) { fp.root = fp.xl; fp.xr = fp.xl; ph = FP_POST; break; }
          if (fp.fu == 0.0) { fp.root = fp.xr; fp.xl = fp.xr; ph = FP_POST; break; }
          xq = fp.xr - (fp.fu * (fp.xl - fp.xr) / (fp.fl - fp.fu));
          ph = FP_LINBIS; nq = 2;
          return;
        }
        case FP_LINBIS: {
          const double xl = fp.oxl, xr = fp.oxr, fl = fp.ofl, x_lin = xq, f_lin = y[0];
          ph = FP_POST;
          if (!isfinite(f_lin)) break;
          fp.froot = f_lin; fp.zroot = zm[0];
          if (f_lin == 0.0) { fp.root = x_lin; fp.xl = x_lin; fp.xr = x_lin; break; }
          double w;
          if ((fl > 0.0 && f_lin < 0.0) || (fl < 0.0 && f_lin > 0.0)) { fp.root = x_lin; fp.xr = x_lin; fp.fu = f_lin; w = x_lin - xl; }
          else { fp.root = x_lin; fp.xl = x_lin; fp.fl = f_lin; w = xr - x_lin; }
          if (w < 0.5 * (xr - xl)) break;
          const double xb = 0.5 * (xl + xr), fb = y[1];
          if (!isfinite(fb)) break;
          if ((fl > 0.0 && fb < 0.0) || (fl < 0.0 && fb > 0.0)) {
            fp.xr = xb; fp.fu = fb;
            if (fp.root > xb) { fp.root = 0.5 * (xl + xb); fp.froot = NAN; }
          } else {
            fp.xl = xb; fp.fl = fb;
            if (fp.root < xb) { fp.root = 0.5 * (xb + xr); fp.froot = NAN; }
          }
          break;
        }
        // ---------------- the loop body of inice_find_root after the iterate
        case FP_POST:
          if (fp.froot == fp.froot) { chk = fp.froot; chk_zm = fp.zroot; ph = FP_TEST; break; }
          xq = fp.root; ph = FP_CHK; nq = 1;
          return;
        case FP_CHK:
          chk = y[0]; chk_zm = zm[0]; ph = FP_TEST;
          break;
        case FP_TEST: {
          const bool same = fp.root == fp.oroot && fp.xl == fp.oxl && fp.xr == fp.oxr && fp.fl == fp.ofl && fp.fu == fp.ofu;
          if (fabs(chk) < 1e-6 || same || ++iter >= 100) { r_root = fp.root; r_cz = chk; r_zm = chk_zm; ph = LADDER; }
          else ph = FP_ITER;
          break;
        }
        // ---------------- Newton: inice_newton_root; f at the iterate and the first pass of inice_deriv_central
        case NW_FD: {
          nw.fv = y[0]; nw.zfv = zm[0];
          double result, round, trunc;
          central(nw.root, nw.h, y[1], y[2], y[3], y[4], result, round, trunc);
          nw.r0 = result; nw.err0 = round + trunc;
          if (round < trunc && (round > 0 && trunc > 0)) {
            nw.h = nw.h * cube_root_ratio(round, trunc);
            ph = NW_D2; nq = 4;
            return;
          }
          ph = NW_AFTER_D;
          break;
        }
        case NW_D2: {                           // second pass at the optimised step
          double result, round, trunc;
          central(nw.root, nw.h, y[0], y[1], y[2], y[3], result, round, trunc);
          const double error_o = round + trunc;
          if (error_o < nw.err0 && fabs(result - nw.r0) < 4.0 * nw.err0) nw.r0 = result;
          ph = NW_AFTER_D;
          break;
        }
        case NW_AFTER_D:
          nw.df = nw.r0;
          if (iter < 0) { iter = 0; ph = NW_STEP; } else ph = NW_TEST;
          break;
        case NW_STEP:
          if (nw.df != 0.0) {                   // else GSL_EZERODIV: root unchanged
            nw.root = nw.root - (nw.fv / nw.df);
            nw.h = 1e-8;
            xq = nw.root; ph = NW_FD; nq = 5;
            return;
          }
          ph = NW_TEST;
          break;
        case NW_TEST: {
          const double x0 = nw.x;
          nw.x = nw.root;
          if (fabs(nw.x - x0) < 1e-6 * fabs(nw.x) || nw.x == x0 || ++iter >= 100) { r_root = nw.x; r_cz = nw.fv; r_zm = nw.zfv; ph = LADDER; }
          else ph = NW_STEP;
          break;
        }
        // ---------------- inice_ra_ladder
        case LADDER:
          switch (stage) {
            case 0:
              lv0 = r_root; cz0 = r_cz; zm0 = r_zm;
              if (fabs(cz0) > 0.5) { stage = 1; nw_begin(lower, up); return; }
              ph = AFTER_FIRST;
              break;
            case 1:
              lv0 = r_root; cz0 = r_cz; zm0 = r_zm;
              ph = AFTER_FIRST;
              break;
            case 2:
              lv1 = r_root; cz1 = r_cz; zm1 = r_zm;
              if (bad1()) { stage = 3; fp_begin(lv0 - 0.15, lv0 - 0.023); if (ph < FP_ITER) return; }
              else ph = END_SECOND;
              break;
            case 3:
              lv1 = r_root; cz1 = r_cz; zm1 = r_zm;
              if (bad1()) {
                stage = 4;
                if (lv0 + 0.005 < up) fp_begin(lv0 + 0.005, up); else fp_begin(lv0 - 0.1, lv0 - 0.01);
                if (ph < FP_ITER) return;
              } else ph = END_SECOND;
              break;
            case 4:
              lv1 = r_root; cz1 = r_cz; zm1 = r_zm;
              if (bad1()) { stage = 5; nw_begin(lv0 - 0.23, lv0 - 0.023); return; }
              ph = END_SECOND;
              break;
            case 5:
              if (fabs(r_root) < A) { lv1 = r_root; cz1 = r_cz; zm1 = r_zm; }
              if (bad1()) { stage = 6; nw_begin(lv0 - 0.1, lv0 - 0.023); return; }
              ph = END_SECOND;
              break;
            default:
              if (fabs(r_root) < A) { lv1 = r_root; cz1 = r_cz; zm1 = r_zm; }
              ph = END_SECOND;
              break;
          }
          break;
        case AFTER_FIRST:
          if (lv0 < 0) cz0 = -1000;
          if (fabs(cz0) < 0.5 && both_absent) { stage = 2; fp_begin(lv0 - 0.23, lv0 - 0.023); if (ph < FP_ITER) return; }
          else { lv1 = 0; cz1 = -1000; zm1 = -1000; ph = DONE; return; }
          break;
        case END_SECOND:
          if (lv1 < 0) cz1 = -1000;
          if (fabs(cz1) < 0.5 && fabs(cz0) < 0.5 && fabs(lv1 - lv0) < 1e-4) cz1 = -1000;
          ph = DONE;
          return;
        default:   // DONE
          return;
      }
    }
  }

  AIRICE_HD InIceRaLadder result() const {
    InIceRaLadder r;
    r.lv[0] = lv0; r.lv[1] = lv1; r.cz[0] = cz0; r.cz[1] = cz1; r.zm[0] = zm0; r.zm[1] = zm1;
    return r;
  }
};

// the ladder through the machine, one step at a time (what the GPU kernel does with the requests pooled per warp);
// n_steps: steps taken = the length of the pair's critical path in evaluations
AIRICE_HD InIceRaLadder inice_ra_ladder_stepped(const AirIceInIce& m, const InIcePair& g, bool flip, bool d_absent,
                                                bool r_absent, double lvalueR, int* n_eval, int* n_steps) {
  InIceRaMachine M;
  M.init(m, g, flip, d_absent, r_absent, lvalueR);
  int steps = 0;
  while (!M.done()) {
    double y[InIceRaMachine::kMaxReq] = {0, 0, 0, 0, 0}, zm[InIceRaMachine::kMaxReq] = {0, 0, 0, 0, 0};
    for (int k = 0; k < M.nq; k++) y[k] = inice_fraa_eval(g, M.query(k), zm[k]);
    M.advance(y, zm);
    steps++;
  }
  if (n_eval) *n_eval = M.n_eval;
  if (n_steps) *n_steps = steps;
  return M.result();
}

}  // namespace airice
