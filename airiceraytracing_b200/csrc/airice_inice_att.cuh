// airice_inice_att.cuh -- in-ice attenuation, focusing factor and the in-ice interpolation table.
//
// Reference (/root/reference/IceRayTracing.cc): GetIceTemperature :135-139, GetIceAttenuationLength :142-162,
// AttenuationIntegrand :165-176, IntegrateOverLAttn :179-200 (gsl_integration_qags, epsrel 1e-7),
// GetTotalAttenuationDirect/Reflected/Refracted :203-219, the AttRay bookkeeping of GetRayTracingSolutions :2970-3148,
// GetFocusingFactor :3218-3293, MakeTable :2614-2724, GetInterpolatedValue :2727-2905.
#pragma once
#include "airice_inice.cuh"
#include "airice_qags.cuh"

// A0 / L_att(x, f) * sqrt(1 + tan(asin(L / n(x)))^2), the reference's expression term by term.  w0 = log(0.0001),
// w2 = log(3.16), w = log(frequency) are call constants made on the host (libm).
struct InIceAttIntegrand {
  double A0, frequency, L, A, B, C, w0, w2, w;
  AIRICE_HD double operator()(double x) const {
    const double depth = fabs(x);
    const double t = 1.83415e-09 * INICE_POW(depth, 3) + (-1.59061e-08 * (depth * depth)) + 0.00267687 * depth + (-51.0696);
    const double w1 = 0.0;
    const double b0 = -6.74890 + t * (0.026709 - t * 0.000884);
    const double b1 = -6.22121 - t * (0.070927 + t * 0.001773);
    const double b2 = -4.09468 - t * (0.002213 + t * 0.000332);
    double a, bb;
    if (frequency < 1.) {
      a = (b1 * w0 - b0 * w1) / (w0 - w1);
      bb = (b1 - b0) / (w1 - w0);
    } else {
      a = (b2 * w1 - b1 * w2) / (w1 - w2);
      bb = (b2 - b1) / (w2 - w1);
    }
    const double Lval = 1. / INICE_EXP(a + bb * w);
    const double n = A + B * INICE_EXP(-C * depth);
    const double tn = tan(asin(L / n));
    return (A0 / Lval) * sqrt(1 + tn * tn);
  }
};

struct InIceAttModel { double A0, frequency, w0, w2, w; };   // w* = log(0.0001), log(3.16), log(frequency)

// IntegrateOverLAttn: fabs(qags(z0 -> z1)).  worst: running maximum of the interval count, status ORed into flags
AIRICE_HD double inice_integrate_att(const AirIceInIce& m, const InIceAttModel& am, double z0, double z1, double L, int& flags,
                                     int& worst) {
  InIceAttIntegrand f = {am.A0, am.frequency, L, m.A, m.B, m.C, am.w0, am.w2, am.w};
  int status, intervals;
  const double r = airice_qags<InIceAttIntegrand, 64>(f, z0, z1, 1e-7, status, intervals);
  if (status < 0) flags |= 1;
  if (intervals > worst) worst = intervals;
  return fabs(r);
}
// kind 0 direct, 1 reflected, 2 refracted (zmax used by 2 only)
AIRICE_HD double inice_total_attenuation(const AirIceInIce& m, const InIceAttModel& am, int kind, double z0, double z1,
                                         double zmax, double L, int& flags, int& worst) {
  z0 = fabs(z0);
  z1 = fabs(z1);
  if (kind == 0) return inice_integrate_att(m, am, z0, z1, L, flags, worst);
  const double end = (kind == 1) ? 0.000001 : zmax;
  const double a = inice_integrate_att(m, am, z0, end, L, flags, worst);
  const double b = inice_integrate_att(m, am, z1, end, L, flags, worst);
  return a + b;
}

// AttD, AttR, AttRa[0..1] of GetRayTracingSolutions (IceRayTracing.cc:2970-2987) from the 29 slots of IceRayTracing():
// 1 - total attenuation for the branches that exist, 0 otherwise
AIRICE_HD void inice_candidate_attenuations(const AirIceInIce& m, const InIceAttModel& am, const double* o, double rx_depth,
                                            double tx_depth, double* att4, int& flags, int& worst) {
  att4[0] = att4[1] = att4[2] = att4[3] = 0;
  if (o[8] != -1000) att4[0] = 1 - inice_total_attenuation(m, am, 0, tx_depth, rx_depth, 0.0, o[19], flags, worst);
  if (o[9] != -1000) att4[1] = 1 - inice_total_attenuation(m, am, 1, tx_depth, rx_depth, 0.0, o[20], flags, worst);
  if (o[10] != -1000) att4[2] = 1 - inice_total_attenuation(m, am, 2, tx_depth, rx_depth, o[23], o[21], flags, worst);
  if (o[11] != -1000) att4[3] = 1 - inice_total_attenuation(m, am, 2, tx_depth, rx_depth, o[24], o[22], flags, worst);
}

// GetFocusingFactor (IceRayTracing.cc:3218-3293) from the two-ray solutions at the receiver depth zR (a: path, launch,
// receive angle of both rays) and at zR - 0.01 (b); f0[2] = the caller's initial values (MakeTable passes {1, 1})
AIRICE_HD void inice_focusing(const AirIceInIce& m, double zT, double zR, const double* path_a, const double* launch_a,
                              const double* recv_a, const double* launch_b, const double* recv_b, double* focusing) {
  const double kpi180 = m.pi / 180;
  const double nTx = inice_nz(m, zT), nRx = inice_nz(m, zR);
  const double recPos0 = zR, recPos2 = zR - 0.01;
  for (int r = 0; r < 2; r++) {
    if (recv_a[r] != -1000 && recv_b[r] != -1000) {
      const double recAng = recv_a[r] * kpi180, lauA = launch_a[r] * kpi180, lauB = launch_b[r] * kpi180;
      focusing[r] = sqrt(((path_a[r] / (sin(recAng) * fabs((recPos2 - recPos0) / (lauB - lauA)))) * (nTx / nRx)));
    }
  }
  if (zR == zT && focusing[0] == 0) focusing[0] = 1.;
}

// GetInterpolatedValue (IceRayTracing.cc:2727-2905): bilinear inside a cell whose four nodes exist, inverse-distance
// weighting over the existing ones otherwise; -1000 outside the grid.  Positions are FLOATs in the reference
// (IceRayTracing.hh:29-30) and enter the arithmetic converted back to double.
AIRICE_HD double inice_table_interp(const float* pos_x, const float* pos_z, int n_x, int n_z, double step_x, double step_z,
                                    const double* col, double xT, double zT) {
  double sum1 = 0, sum2 = 0, NewZValue = -1000;
  const double start_x = pos_x[0], start_z = pos_z[0], stop_x = pos_x[n_x - 1], stop_z = pos_z[n_z - 1];
  if (xT >= start_x && xT <= stop_x && zT >= start_z && zT <= stop_z) {
    const double x = xT, y = zT;
    const int minXbin = (int)floor((xT - start_x) / step_x);
    const int minZbin = (int)floor(fabs(zT - start_z) / step_z);
    if (minXbin + 1 <= n_x - 1 && minZbin + 1 <= n_z - 1) {
      const double x1 = pos_x[minXbin], y1 = pos_z[minZbin], x2 = pos_x[minXbin + 1], y2 = pos_z[minZbin + 1];
      double f11 = col[(int64_t)minXbin * n_z + minZbin];
      double f12 = col[(int64_t)minXbin * n_z + (minZbin + 1)];
      double f21 = col[(int64_t)(minXbin + 1) * n_z + minZbin];
      double f22 = col[(int64_t)(minXbin + 1) * n_z + (minZbin + 1)];
      if (f11 == -1000 || f12 == -1000 || f21 == -1000 || f22 == -1000) {
        const bool all_absent = f11 == -1000 && f12 == -1000 && f21 == -1000 && f22 == -1000;
        if (f11 != -1000) { const double w = 1.0 / ((x1 - x) * (x1 - x) + (y1 - y) * (y1 - y)); sum1 += w * f11; sum2 += w; }
        if (f12 != -1000) { const double w = 1.0 / ((x1 - x) * (x1 - x) + (y2 - y) * (y2 - y)); sum1 += w * f12; sum2 += w; }
        if (f21 != -1000) { const double w = 1.0 / ((x2 - x) * (x2 - x) + (y1 - y) * (y1 - y)); sum1 += w * f21; sum2 += w; }
        if (f22 != -1000) { const double w = 1.0 / ((x2 - x) * (x2 - x) + (y2 - y) * (y2 - y)); sum1 += w * f22; sum2 += w; }
        NewZValue = sum1 / sum2;
        // the reference zeroes the absent values before its all-absent test, so that test never fires; 0/0 = NaN does
        (void)all_absent;
        if (NewZValue != NewZValue) NewZValue = -1000;
      } else {
        const double w11 = ((x2 - x) * (y2 - y)) / ((x2 - x1) * (y2 - y1));
        const double w12 = ((x2 - x) * (y - y1)) / ((x2 - x1) * (y2 - y1));
        const double w21 = ((x - x1) * (y2 - y)) / ((x2 - x1) * (y2 - y1));
        const double w22 = ((x - x1) * (y - y1)) / ((x2 - x1) * (y2 - y1));
        NewZValue = w11 * f11 + w12 * f12 + w21 * f21 + w22 * f22;
      }
    }
  }
  return NewZValue;
}
