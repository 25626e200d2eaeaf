// airice_qags.cuh -- adaptive quadrature of the in-ice attenuation integrals, one integral per thread.
//
// Reference: IceRayTracing::IntegrateOverLAttn (IceRayTracing.cc:179-200) calls gsl_integration_qags(f, a, b, epsabs = 0,
// epsrel = 1e-7, limit = 1000): QUADPACK's QAGS -- 21-point Gauss-Kronrod rule, bisection of the interval with the
// largest error estimate, Wynn's epsilon algorithm on the sequence of totals once the smallest intervals are reached
// (GSL integration/qags.c, qk.c, qk21.c, qpsrt.c, qelg.c).  Like the root finders of the in-ice solver, WHAT it returns is
// a property of the procedure (a tolerance of 1e-7 leaves ~1e-9 .. 1e-8 of slack in the attenuation, more than the
// north-star tolerance), so the same procedure runs here: same rule, same bisection order, same extrapolation table, same
// tests in the same order.  Direct and reflected rays return from the first rule (21 evaluations); refracted rays end at
// the turning depth, where the integrand has a 1/sqrt end-point singularity, and take 6-20 bisections plus extrapolation.
//
// The interval list lives in the thread's local memory with room for kCap intervals (the reference allows 1000; the
// largest count seen over 1e5 random refracted rays is ~20).  A thread that would need more stops with status -1.
#pragma once
#include "airice_glibc_math.cuh"

#define AIRICE_QAGS_DBL_EPS 2.2204460492503131e-16
#define AIRICE_QAGS_DBL_MIN 2.2250738585072014e-308
#define AIRICE_QAGS_DBL_MAX 1.7976931348623157e+308

#if defined(__CUDACC__)
#define AIRICE_Q_FN __host__ __device__ __forceinline__
#define AIRICE_Q_NOINLINE __host__ __device__ __noinline__
#else
#define AIRICE_Q_FN inline
#define AIRICE_Q_NOINLINE inline
#endif

struct AirIceQk { double result, abserr, resabs, resasc; };

// gsl_integration_qk21 (generic qk with n = 11; rescale_error with pow(., 1.5) from the same libm as the reference's)
template <class F>
AIRICE_Q_NOINLINE AirIceQk airice_qk21(const F& f, double a, double b) {
  const double xgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
                          0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
                          0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                          0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
                          0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
                          0.000000000000000000000000000000000};
  const double wg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
                        0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
                        0.295524224714752870173815619188769};
  const double wgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
                          0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
                          0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                          0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
                          0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
                          0.149445554002916905664936468389821};
  double fv1[11], fv2[11];
  const double center = 0.5 * (a + b);
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  const double f_center = f(center);
  double result_gauss = 0;
  double result_kronrod = f_center * wgk[10];
  double result_abs = fabs(result_kronrod);
#pragma unroll 1
  for (int j = 0; j < 5; j++) {
    const int jtw = j * 2 + 1;
    const double abscissa = half_length * xgk[jtw];
    const double fval1 = f(center - abscissa);
    const double fval2 = f(center + abscissa);
    const double fsum = fval1 + fval2;
    fv1[jtw] = fval1;
    fv2[jtw] = fval2;
    result_gauss += wg[j] * fsum;
    result_kronrod += wgk[jtw] * fsum;
    result_abs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
  }
#pragma unroll 1
  for (int j = 0; j < 5; j++) {
    const int jtwm1 = j * 2;
    const double abscissa = half_length * xgk[jtwm1];
    const double fval1 = f(center - abscissa);
    const double fval2 = f(center + abscissa);
    fv1[jtwm1] = fval1;
    fv2[jtwm1] = fval2;
    result_kronrod += wgk[jtwm1] * (fval1 + fval2);
    result_abs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
  }
  const double mean = result_kronrod * 0.5;
  double result_asc = wgk[10] * fabs(f_center - mean);
#pragma unroll 1
  for (int j = 0; j < 10; j++) result_asc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
  double err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  // rescale_error
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    const double scale = airice_glibc_pow((200 * err / result_asc), 1.5);
    if (scale < 1) err = result_asc * scale;
    else err = result_asc;
  }
  if (result_abs > AIRICE_QAGS_DBL_MIN / (50 * AIRICE_QAGS_DBL_EPS)) {
    const double min_err = 50 * AIRICE_QAGS_DBL_EPS * result_abs;
    if (min_err > err) err = min_err;
  }
  AirIceQk q;
  q.result = result_kronrod; q.abserr = err; q.resabs = result_abs; q.resasc = result_asc;
  return q;
}

template <int kCap>
struct AirIceQagsWork {
  static constexpr int kLimit = 1000;      // the reference's `limit`: it shapes qpsrt's window, not the storage
  double alist[kCap], blist[kCap], rlist[kCap], elist[kCap];
  short order[kCap + 1], level[kCap];
  int size, nrmax, i, maximum_level;

  // qpsrt.c: keep `order` sorted by decreasing error estimate
  AIRICE_Q_FN void qpsrt() {
    const int last = size - 1;
    int i_nrmax = nrmax;
    int i_maxerr = order[i_nrmax];
    if (last < 2) {
      order[0] = 0;
      order[1] = 1;
      i = i_maxerr;
      return;
    }
    const double errmax = elist[i_maxerr];
    while (i_nrmax > 0 && errmax > elist[order[i_nrmax - 1]]) {
      order[i_nrmax] = order[i_nrmax - 1];
      i_nrmax--;
    }
    const int top = (last < (kLimit / 2 + 2)) ? last : kLimit - last + 1;
    int ii = i_nrmax + 1;
    while (ii < top && errmax < elist[order[ii]]) {
      order[ii - 1] = order[ii];
      ii++;
    }
    order[ii - 1] = (short)i_maxerr;
    const double errmin = elist[last];
    int k = top - 1;
    while (k > ii - 2 && errmin >= elist[order[k]]) {
      order[k + 1] = order[k];
      k--;
    }
    order[k + 1] = (short)last;
    i_maxerr = order[i_nrmax];
    i = i_maxerr;
    nrmax = i_nrmax;
  }

  AIRICE_Q_FN void update(double a1, double b1, double area1, double error1, double a2, double b2, double area2, double error2) {
    const int i_max = i, i_new = size;
    const int new_level = level[i_max] + 1;
    if (error2 > error1) {
      alist[i_max] = a2; rlist[i_max] = area2; elist[i_max] = error2; level[i_max] = (short)new_level;
      alist[i_new] = a1; blist[i_new] = b1; rlist[i_new] = area1; elist[i_new] = error1; level[i_new] = (short)new_level;
    } else {
      blist[i_max] = b1; rlist[i_max] = area1; elist[i_max] = error1; level[i_max] = (short)new_level;
      alist[i_new] = a2; blist[i_new] = b2; rlist[i_new] = area2; elist[i_new] = error2; level[i_new] = (short)new_level;
    }
    size++;
    if (new_level > maximum_level) maximum_level = new_level;
    qpsrt();
  }

  AIRICE_Q_FN bool increase_nrmax() {
    const int id = nrmax;
    const int last = size - 1;
    const int jupbnd = (last > (1 + kLimit / 2)) ? kLimit + 1 - last : last;
    for (int k = id; k <= jupbnd; k++) {
      const int i_max = order[nrmax];
      i = i_max;
      if (level[i_max] < maximum_level) return true;
      nrmax++;
    }
    return false;
  }
};

struct AirIceQelgTable { int n; double rlist2[52]; int nres; double res3la[3]; };

// qelg.c: Wynn's epsilon algorithm on the table of totals
AIRICE_Q_NOINLINE void airice_qelg(AirIceQelgTable& table, double& result, double& abserr) {
  double* epstab = table.rlist2;
  double* res3la = table.res3la;
  const int n = table.n - 1;
  const double current = epstab[n];
  double absolute = AIRICE_QAGS_DBL_MAX;
  double relative = 5 * AIRICE_QAGS_DBL_EPS * fabs(current);
  const int newelm = n / 2;
  const int n_orig = n;
  int n_final = n;
  const int nres_orig = table.nres;
  result = current;
  abserr = AIRICE_QAGS_DBL_MAX;
  if (n < 2) {
    result = current;
    abserr = absolute > relative ? absolute : relative;
    return;
  }
  epstab[n + 2] = epstab[n];
  epstab[n] = AIRICE_QAGS_DBL_MAX;
  for (int i = 0; i < newelm; i++) {
    double res = epstab[n - 2 * i + 2];
    const double e0 = epstab[n - 2 * i - 2];
    const double e1 = epstab[n - 2 * i - 1];
    const double e2 = res;
    const double e1abs = fabs(e1);
    const double delta2 = e2 - e1;
    const double err2 = fabs(delta2);
    const double tol2 = (fabs(e2) > e1abs ? fabs(e2) : e1abs) * AIRICE_QAGS_DBL_EPS;
    const double delta3 = e1 - e0;
    const double err3 = fabs(delta3);
    const double tol3 = (e1abs > fabs(e0) ? e1abs : fabs(e0)) * AIRICE_QAGS_DBL_EPS;
    if (err2 <= tol2 && err3 <= tol3) {
      result = res;
      absolute = err2 + err3;
      relative = 5 * AIRICE_QAGS_DBL_EPS * fabs(res);
      abserr = absolute > relative ? absolute : relative;
      return;
    }
    const double e3 = epstab[n - 2 * i];
    epstab[n - 2 * i] = e1;
    const double delta1 = e1 - e3;
    const double err1 = fabs(delta1);
    const double tol1 = (e1abs > fabs(e3) ? e1abs : fabs(e3)) * AIRICE_QAGS_DBL_EPS;
    if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
      n_final = 2 * i;
      break;
    }
    const double ss = (1 / delta1 + 1 / delta2) - 1 / delta3;
    if (fabs(ss * e1) <= 0.0001) {
      n_final = 2 * i;
      break;
    }
    res = e1 + 1 / ss;
    epstab[n - 2 * i] = res;
    const double error = err2 + fabs(res - e2) + err3;
    if (error <= abserr) {
      abserr = error;
      result = res;
    }
  }
  {
    const int limexp = 50 - 1;
    if (n_final == limexp) n_final = 2 * (limexp / 2);
  }
  if (n_orig % 2 == 1) {
    for (int i = 0; i <= newelm; i++) epstab[1 + i * 2] = epstab[i * 2 + 3];
  } else {
    for (int i = 0; i <= newelm; i++) epstab[i * 2] = epstab[i * 2 + 2];
  }
  if (n_orig != n_final) {
    for (int i = 0; i <= n_final; i++) epstab[i] = epstab[n_orig - n_final + i];
  }
  table.n = n_final + 1;
  if (nres_orig < 3) {
    res3la[nres_orig] = result;
    abserr = AIRICE_QAGS_DBL_MAX;
  } else {
    abserr = (fabs(result - res3la[2]) + fabs(result - res3la[1]) + fabs(result - res3la[0]));
    res3la[0] = res3la[1];
    res3la[1] = res3la[2];
    res3la[2] = result;
  }
  table.nres = nres_orig + 1;
  const double floor_err = 5 * AIRICE_QAGS_DBL_EPS * fabs(result);
  abserr = abserr > floor_err ? abserr : floor_err;
}

// gsl_integration_qags(f, a, b, 0, epsrel, 1000, ...) -> result.  status: 0 success, 1..6 GSL's error classes (the
// reference ignores them and uses the result), -1 more than kCap intervals (result = the running total; never seen).
// intervals: how many the integral used (diagnostics).
template <class F, int kCap = 64>
AIRICE_Q_NOINLINE double airice_qags(const F& f, const double a, const double b, const double epsrel, int& status, int& intervals) {
  const double epsabs = 0.0;
  const int limit = AirIceQagsWork<kCap>::kLimit;
  AirIceQagsWork<kCap> w;
  w.size = 0; w.nrmax = 0; w.i = 0;
  w.alist[0] = a; w.blist[0] = b; w.rlist[0] = 0.0; w.elist[0] = 0.0; w.order[0] = 0; w.level[0] = 0; w.maximum_level = 0;
  status = 0;
  intervals = 1;
  const AirIceQk q0 = airice_qk21(f, a, b);
  const double result0 = q0.result, abserr0 = q0.abserr, resabs0 = q0.resabs, resasc0 = q0.resasc;
  w.size = 1; w.rlist[0] = result0; w.elist[0] = abserr0;
  double tolerance = epsrel * fabs(result0);
  if (epsabs > tolerance) tolerance = epsabs;
  if (abserr0 <= 100 * AIRICE_QAGS_DBL_EPS * resabs0 && abserr0 > tolerance) { status = 2; return result0; }
  else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) return result0;

  AirIceQelgTable table;
  table.n = 0; table.nres = 0;
  table.rlist2[table.n] = result0; table.n++;
  double area = result0, errsum = abserr0;
  double res_ext = result0, err_ext = AIRICE_QAGS_DBL_MAX;
  const bool positive_integrand = (fabs(result0) >= (1 - 50 * AIRICE_QAGS_DBL_EPS) * resabs0);
  double ertest = 0, error_over_large_intervals = 0, reseps = 0, abseps = 0, correc = 0;
  int ktmin = 0, roundoff_type1 = 0, roundoff_type2 = 0, roundoff_type3 = 0, error_type = 0, error_type2 = 0;
  bool extrapolate = false, disallow_extrapolation = false;
  int iteration = 1;
  bool compute_result = false;

#pragma unroll 1
  do {
    if (w.size >= kCap) { status = -1; compute_result = true; break; }
    const double a_i = w.alist[w.i], b_i = w.blist[w.i], r_i = w.rlist[w.i], e_i = w.elist[w.i];
    const int current_level = w.level[w.i] + 1;
    const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
    iteration++;
    const AirIceQk q1 = airice_qk21(f, a1, b1);
    const AirIceQk q2 = airice_qk21(f, a2, b2);
    const double area1 = q1.result, error1 = q1.abserr, resasc1 = q1.resasc;
    const double area2 = q2.result, error2 = q2.abserr, resasc2 = q2.resasc;
    const double area12 = area1 + area2;
    const double error12 = error1 + error2;
    const double last_e_i = e_i;
    errsum = errsum + error12 - e_i;
    area = area + area12 - r_i;
    tolerance = epsrel * fabs(area);
    if (epsabs > tolerance) tolerance = epsabs;
    if (resasc1 != error1 && resasc2 != error2) {
      const double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) {
        if (!extrapolate) roundoff_type1++;
        else roundoff_type2++;
      }
      if (iteration > 10 && error12 > e_i) roundoff_type3++;
    }
    if (roundoff_type1 + roundoff_type2 >= 10 || roundoff_type3 >= 20) error_type = 2;
    if (roundoff_type2 >= 5) error_type2 = 1;
    {
      const double tmp = (1 + 100 * AIRICE_QAGS_DBL_EPS) * (fabs(a2) + 1000 * AIRICE_QAGS_DBL_MIN);
      if (fabs(a1) <= tmp && fabs(b2) <= tmp) error_type = 4;
    }
    w.update(a1, b1, area1, error1, a2, b2, area2, error2);
    intervals = w.size;
    if (errsum <= tolerance) { compute_result = true; break; }
    if (error_type) break;
    if (iteration >= limit - 1) { error_type = 1; break; }
    if (iteration == 2) {
      error_over_large_intervals = errsum;
      ertest = tolerance;
      table.rlist2[table.n] = area; table.n++;
      continue;
    }
    if (disallow_extrapolation) continue;
    error_over_large_intervals += -last_e_i;
    if (current_level < w.maximum_level) error_over_large_intervals += error12;
    if (!extrapolate) {
      if (w.level[w.i] < w.maximum_level) continue;
      extrapolate = true;
      w.nrmax = 1;
    }
    if (!error_type2 && error_over_large_intervals > ertest) {
      if (w.increase_nrmax()) continue;
    }
    table.rlist2[table.n] = area; table.n++;
    airice_qelg(table, reseps, abseps);
    ktmin++;
    if (ktmin > 5 && err_ext < 0.001 * errsum) error_type = 5;
    if (abseps < err_ext) {
      ktmin = 0;
      err_ext = abseps;
      res_ext = reseps;
      correc = error_over_large_intervals;
      ertest = epsrel * fabs(reseps);
      if (epsabs > ertest) ertest = epsabs;
      if (err_ext <= ertest) break;
    }
    if (table.n == 1) disallow_extrapolation = true;
    if (error_type == 5) break;
    w.nrmax = 0; w.i = w.order[0];
    extrapolate = false;
    error_over_large_intervals = errsum;
  } while (iteration < limit);

  double result = res_ext;
  bool return_error = false;
  if (!compute_result) {
    if (err_ext == AIRICE_QAGS_DBL_MAX) compute_result = true;
    else {
      if (error_type || error_type2) {
        if (error_type2) err_ext += correc;
        if (error_type == 0) error_type = 3;
        if (res_ext != 0.0 && area != 0.0) {
          if (err_ext / fabs(res_ext) > errsum / fabs(area)) compute_result = true;
        } else if (err_ext > errsum) {
          compute_result = true;
        } else if (area == 0.0) {
          return_error = true;
        }
      }
      if (!compute_result && !return_error) {
        const double max_area = fabs(res_ext) > fabs(area) ? fabs(res_ext) : fabs(area);
        if (!positive_integrand && max_area < 0.01 * resabs0) return_error = true;
      }
      if (!compute_result && !return_error) {
        const double ratio = res_ext / area;
        if (ratio < 0.01 || ratio > 100.0 || errsum > fabs(area)) error_type = 6;
      }
    }
  }
  if (compute_result) {
    double s = 0;
    for (int k = 0; k < w.size; k++) s += w.rlist[k];
    result = s;
  }
  if (error_type > 2) error_type--;
  if (status == 0) status = error_type;
  return result;
}
