// airice_inice_machine.cuh -- the refracted-ray root-search ladder of airice_inice.cuh as a resumable state machine.
//
// inice_ra_ladder() is a ladder of up to seven root searches (GSL falsepos / Newton with a numerical derivative), each
// a loop of a data-dependent number of evaluations of ONE function, fRaa(L) (which itself hides a ~30-step falsepos for
// the turning depth).  Run as nested loops, the 32 lanes of a warp sit in different searches at different iterations
// and execute one after the other (measured: 5.6 of 32 lanes active).  Here the nesting is turned inside out: a lane
// holds its whole search state in this struct, asks for ONE evaluation at `xq`, and advance() consumes the value and
// moves to the next request.  The warp then loops over "evaluate fRaa(xq)" -- the expensive part, identical code for
// every lane whatever search it is in -- and only the short bookkeeping in advance() diverges.  A lane whose ladder is
// complete takes the next pair from the work list in the same loop.
//
// Every evaluation, comparison and update is the one inice_find_root / inice_newton_root / inice_deriv_central make, in
// the same order; tests/test_inice.py runs both forms on the host over the same pairs and requires identical bits.
// Two evaluations of the literal form are not repeated because their values are already at hand (f is deterministic):
// f(root) after a search = the last residual check of the falsepos loop, or the Newton iterate's own f value.
#pragma once
#include "airice_inice.cuh"

namespace airice {

struct InIceRaMachine {
  enum Phase : int {
    // states that wait for an evaluation of fRaa at xq
    FP_LO, FP_HI, FP_LIN, FP_BIS, FP_CHK, NW_F0, NW_F, DV_1, DV_2, DV_3, DV_4,
    // internal states (no evaluation pending)
    FP_ITER, FP_POST, FP_TEST, NW_DERIV, NW_AFTER_D, NW_STEP, NW_TEST, LADDER, AFTER_FIRST, END_SECOND, DONE
  };
  double xq;                                   // where fRaa is wanted next
  double A, lower, up;                         // ice A, first bracket
  double lv0, cz0, zm0, lv1, cz1, zm1;         // ladder results
  union {
    struct { double xl, xr, fl, fu, root, oxl, oxr, ofl, ofu, oroot, froot, zroot; } fp;   // falsepos search
    struct { double x, root, fv, zfv, df, h, fm1, fp1, fmh, r0, err0; } nw;                // Newton search
  };
  int ph, iter, stage;
  bool both_absent, dpass;
  int n_eval;                                  // evaluations consumed (diagnostics)

  AIRICE_HD bool done() const { return ph == DONE; }

  AIRICE_HD void fp_begin(double lo, double hi) {
    iter = 0;
    if (lo > hi) {      // gsl_root_fsolver_set refuses: zeroed solver state
      fp.fl = 0; fp.fu = 0; fp.root = 0; fp.xl = 0; fp.xr = 0;
      ph = FP_ITER;
      return;
    }
    fp.fl = 0; fp.fu = 0; fp.root = 0.5 * (lo + hi); fp.xl = lo; fp.xr = hi;
    xq = lo; ph = FP_LO;
  }
  AIRICE_HD void nw_begin(double lo, double hi) {
    nw.x = (lo + hi) / 2;
    xq = nw.x; ph = NW_F0;
  }

  AIRICE_HD void init(const AirIceInIce& m, const InIcePair& g, bool flip, bool d_absent, bool r_absent, double lvalueR) {
    A = m.A;
    inice_ra_first_bracket(m, g, flip, lvalueR, lower, up);
    both_absent = d_absent && r_absent;
    lv0 = 0; cz0 = -1000; zm0 = 10; lv1 = 0; cz1 = -1000; zm1 = 10;
    stage = 0; n_eval = 0; dpass = false;
    fp_begin(lower, up);
    if (ph >= FP_ITER) run(0.0, 0.0);
  }

  // consume y = fRaa(xq) and zm = the turning depth that evaluation used; leaves the next request in xq, or done()
  AIRICE_HD void advance(double y, double zm) { n_eval++; run(y, zm); }

  AIRICE_HD bool bad1() const { return fabs(cz1) > 0.5 || cz1 != cz1 || fabs(lv1 - lv0) < 1e-4; }

  AIRICE_HD void run(double y, double zm) {
    double r_root = 0, r_cz = 0, r_zm = 0;     // result of a finished search, handed to LADDER
    double chk = 0, chk_zm = 0;
    for (;;) {
      switch (ph) {
        // ---------------- falsepos: inice_falsepos_set
        case FP_LO:
          if (!isfinite(y)) { ph = FP_ITER; break; }
          fp.ofl = y;                           // parked until f(hi) is known to be finite too
          xq = fp.xr; ph = FP_HI;
          return;
        case FP_HI:
          if (isfinite(y)) { fp.fl = fp.ofl; fp.fu = y; }
          ph = FP_ITER;
          break;
        // ---------------- inice_falsepos_iterate
        case FP_ITER: {
          fp.oxl = fp.xl; fp.oxr = fp.xr; fp.ofl = fp.fl; fp.ofu = fp.fu; fp.oroot = fp.root;
          fp.froot = NAN;
          if (fp.fl == 0.0) { fp.root = fp.xl; fp.xr = fp.xl; ph = FP_POST; break; }
          if (fp.fu == 0.0) { fp.root = fp.xr; fp.xl = fp.xr; ph = FP_POST; break; }
          xq = fp.xr - (fp.fu * (fp.xl - fp.xr) / (fp.fl - fp.fu));
          ph = FP_LIN;
          return;
        }
        case FP_LIN: {
          const double xl = fp.oxl, xr = fp.oxr, fl = fp.ofl, x_lin = xq, f_lin = y;
          if (!isfinite(f_lin)) { ph = FP_POST; break; }
          fp.froot = f_lin; fp.zroot = zm;
          if (f_lin == 0.0) { fp.root = x_lin; fp.xl = x_lin; fp.xr = x_lin; ph = FP_POST; break; }
          double w;
          if ((fl > 0.0 && f_lin < 0.0) || (fl < 0.0 && f_lin > 0.0)) { fp.root = x_lin; fp.xr = x_lin; fp.fu = f_lin; w = x_lin - xl; }
          else { fp.root = x_lin; fp.xl = x_lin; fp.fl = f_lin; w = xr - x_lin; }
          if (w < 0.5 * (xr - xl)) { ph = FP_POST; break; }
          xq = 0.5 * (xl + xr);
          ph = FP_BIS;
          return;
        }
        case FP_BIS: {
          const double xl = fp.oxl, xr = fp.oxr, fl = fp.ofl, xb = xq, fb = y;
          ph = FP_POST;
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
          xq = fp.root; ph = FP_CHK;
          return;
        case FP_CHK:
          chk = y; chk_zm = zm; ph = FP_TEST;
          break;
        case FP_TEST: {
          const bool same = fp.root == fp.oroot && fp.xl == fp.oxl && fp.xr == fp.oxr && fp.fl == fp.ofl && fp.fu == fp.ofu;
          if (fabs(chk) < 1e-6 || same || ++iter >= 100) { r_root = fp.root; r_cz = chk; r_zm = chk_zm; ph = LADDER; }
          else ph = FP_ITER;
          break;
        }
        // ---------------- Newton: inice_newton_root
        case NW_F0:
          nw.fv = y; nw.zfv = zm; nw.root = nw.x; iter = -1;
          ph = NW_DERIV;
          break;
        case NW_F:
          nw.fv = y; nw.zfv = zm;
          ph = NW_DERIV;
          break;
        case NW_DERIV:                          // inice_deriv_central(f, root, 1e-8)
          nw.h = 1e-8; dpass = false;
          xq = nw.root - nw.h; ph = DV_1;
          return;
        case DV_1: nw.fm1 = y; xq = nw.root + nw.h; ph = DV_2; return;
        case DV_2: nw.fp1 = y; xq = nw.root - nw.h / 2; ph = DV_3; return;
        case DV_3: nw.fmh = y; xq = nw.root + nw.h / 2; ph = DV_4; return;
        case DV_4: {
          const double x = nw.root, h = nw.h, fm1 = nw.fm1, fp1 = nw.fp1, fmh = nw.fmh, fph = y;
          const double r3 = 0.5 * (fp1 - fm1);
          const double r5 = (4.0 / 3.0) * (fph - fmh) - (1.0 / 3.0) * r3;
          const double e3 = (fabs(fp1) + fabs(fm1)) * 2.2204460492503131e-16;
          const double e5 = 2.0 * (fabs(fph) + fabs(fmh)) * 2.2204460492503131e-16 + e3;
          const double a = fabs(r3 / h), b = fabs(r5 / h);
          const double dy = (a > b ? a : b) * (fabs(x) / h) * 2.2204460492503131e-16;
          const double result = r5 / h, trunc = fabs((r5 - r3) / h), round = fabs(e5 / h) + dy;
          if (!dpass) {
            nw.r0 = result; nw.err0 = round + trunc;
            if (round < trunc && (round > 0 && trunc > 0)) {
              nw.h = h * pow(round / (2.0 * trunc), 1.0 / 3.0);
              dpass = true;
              xq = nw.root - nw.h; ph = DV_1;
              return;
            }
          } else {
            const double error_o = round + trunc;
            if (error_o < nw.err0 && fabs(result - nw.r0) < 4.0 * nw.err0) nw.r0 = result;
          }
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
            xq = nw.root; ph = NW_F;
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

// the ladder through the machine, one evaluation at a time (what a GPU lane does)
AIRICE_HD InIceRaLadder inice_ra_ladder_stepped(const AirIceInIce& m, const InIcePair& g, bool flip, bool d_absent,
                                                bool r_absent, double lvalueR, int* n_eval) {
  InIceRaMachine M;
  M.init(m, g, flip, d_absent, r_absent, lvalueR);
  while (!M.done()) {
    double zm;
    const double y = inice_fraa_eval(g, M.xq, zm);
    M.advance(y, zm);
  }
  if (n_eval) *n_eval = M.n_eval;
  return M.result();
}

}  // namespace airice
