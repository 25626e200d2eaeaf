// IceRayTracing.cc -- host-side mirror of the reference's in-ice entry point on top of include/airice_b200.h.
#include "IceRayTracing.hh"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "airice_b200.h"

namespace IceRayTracing {

double A_ice = 1.78, B_ice = -0.43, C_ice = 0.0132;
double GridStepSizeX_O = 0.1, GridStepSizeZ_O = 0.1, GridWidthX = 40, GridWidthZ = 20;   // IceRayTracing.hh:33-36

namespace detail {
struct State {
  airice_ctx *ctx = nullptr;
  int device = 0;
  std::string atmosphere = "Atmosphere.dat";
  double ice_set[3] = {1.78, -0.43, 0.0132};
  std::vector<airice_inice_table *> tables;     // GridZValueb, one per antenna
};
inline State &state() {
  static State s;
  return s;
}
inline bool ensure_ctx() {
  State &s = state();
  if (!s.ctx) {
    // ice-only context: the reference's IceRayTracing needs no atmosphere file either
    if (airice_create(nullptr, AIRICE_VARIANT_MULTIRAY, s.device, &s.ctx) != 0) {
      std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
      s.ctx = nullptr;
      return false;
    }
  }
  if (s.ice_set[0] != A_ice || s.ice_set[1] != B_ice || s.ice_set[2] != C_ice) {
    airice_set_ice_model(s.ctx, A_ice, B_ice, C_ice);
    s.ice_set[0] = A_ice; s.ice_set[1] = B_ice; s.ice_set[2] = C_ice;
  }
  return true;
}
}  // namespace detail

void SetA(double &A) { A_ice = A; }
void SetB(double &B) { B_ice = B; }
void SetC(double &C) { C_ice = C; }
void SetDevice(int device) { detail::state().device = device; }
void SetAtmosphereFile(const std::string &path) { detail::state().atmosphere = path; }

int IceRayTracingBatch(long n, const double *z0, const double *x1, const double *z1, double *out, unsigned char *mask) {
  if (!detail::ensure_ctx()) return 1;
  int rc = airice_inice_solve_host(detail::state().ctx, n, z0, x1, z1, out, mask);
  if (rc != 0) std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
  return rc;
}

double *IceRayTracing(double x0, double z0, double x1, double z1) {
  (void)x0;
  double *output = new double[29];
  unsigned char mask = 0;
  if (IceRayTracingBatch(1, &z0, &x1, &z1, output, &mask) != 0) {
    for (int i = 0; i < 29; i++) output[i] = 0;
    for (int i = 8; i < 12; i++) output[i] = -1000;
  }
  return output;
}

int GetRayTracingSolutionsBatch(long n, const double *RxDepth, const double *Distance, const double *TxDepth, double *out,
                                int *ignore) {
  static_assert(sizeof(int) == sizeof(int32_t), "IgnoreCh is int32 in the C ABI");
  if (!detail::ensure_ctx()) return 1;
  int rc = airice_inice_two_rays_host(detail::state().ctx, n, RxDepth, Distance, TxDepth, out, (int32_t *)ignore);
  if (rc != 0) std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
  return rc;
}

int GetRayTracingSolutionsBatch(long n, const double *RxDepth, const double *Distance, const double *TxDepth, double A0,
                                double frequency, double *out, double *att, int *ignore) {
  if (!detail::ensure_ctx()) return 1;
  int rc = airice_inice_two_rays_att_host(detail::state().ctx, n, RxDepth, Distance, TxDepth, A0, frequency, out, att,
                                          (int32_t *)ignore);
  if (rc != 0) std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
  return rc;
}

void GetRayTracingSolutions(double RxDepth, double Distance, double TxDepth, double TimeRay[2], double PathRay[2],
                            double LaunchAngle[2], double RecieveAngle[2], int IgnoreCh[2], double IncidenceAngleInIce[2],
                            double A0, double frequency, double AttRay[2]) {
  double out[10], att[2] = {0, 0};
  int ig[2] = {0, 0};
  if (GetRayTracingSolutionsBatch(1, &RxDepth, &Distance, &TxDepth, A0, frequency, out, att, ig) != 0) {
    for (int k = 0; k < 10; k++) out[k] = 0;
    out[6] = out[7] = -1000;
  }
  for (int k = 0; k < 2; k++) {
    TimeRay[k] = out[0 + k]; PathRay[k] = out[2 + k]; LaunchAngle[k] = out[4 + k]; RecieveAngle[k] = out[6 + k];
    IncidenceAngleInIce[k] = out[8 + k]; IgnoreCh[k] = ig[k];
    if (AttRay) AttRay[k] = att[k];
  }
}

// IceRayTracing.cc:135-162: closed forms, evaluated where they are called (host)
double GetIceTemperature(double z) {
  double depth = fabs(z);
  double t = 1.83415e-09 * pow(depth, 3) + (-1.59061e-08 * pow(depth, 2)) + 0.00267687 * depth + (-51.0696);
  return t;
}
double GetIceAttenuationLength(double z, double frequency) {
  double t = GetIceTemperature(z);
  const double f0 = 0.0001, f2 = 3.16;
  const double w0 = log(f0), w1 = 0.0, w2 = log(f2), w = log(frequency);
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
  return 1. / exp(a + bb * w);
}

namespace detail {
// one ray through airice_inice_attenuation_device (scalar calls are batch-of-1 launches: slow, correct)
inline double attenuation1(int kind, double A0, double frequency, double z0, double z1, double zmax, double L) {
  if (!ensure_ctx()) return NAN;
  State &s = state();
  const double in[4] = {z0, z1, zmax, L};
  double out = NAN;
  if (airice_inice_attenuation_host(s.ctx, 1, kind, A0, frequency, &in[0], &in[1], &in[2], &in[3], &out) != 0)
    std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
  return out;
}
}  // namespace detail
double GetTotalAttenuationDirect(double A0, double frequency, double z0, double z1, double Lvalue) {
  return detail::attenuation1(0, A0, frequency, z0, z1, 0.0, Lvalue);
}
double GetTotalAttenuationReflected(double A0, double frequency, double z0, double z1, double Lvalue) {
  return detail::attenuation1(1, A0, frequency, z0, z1, 0.0, Lvalue);
}
double GetTotalAttenuationRefracted(double A0, double frequency, double z0, double z1, double zmax, double Lvalue) {
  return detail::attenuation1(2, A0, frequency, z0, z1, zmax, Lvalue);
}

int GetFocusingFactorBatch(long n, const double *zT, const double *xR, const double *zR, double *out) {
  if (!detail::ensure_ctx()) return 1;
  int rc = airice_inice_focusing_host(detail::state().ctx, n, zT, xR, zR, out);
  if (rc != 0) std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
  return rc;
}
void GetFocusingFactor(double zT, double xR, double zR, double focusing[2]) {
  // the kernel starts from {1, 1} like the reference's callers; a ray whose factor is not computed (a branch missing at
  // zR or at zR - 0.01) keeps the CALLER's initial value, as in the reference
  double out[2] = {1, 1};
  if (GetFocusingFactorBatch(1, &zT, &xR, &zR, out) != 0) return;
  double ra[10];
  int ig[2];
  // which rays were computed: both solutions present (IceRayTracing.cc:3270, 3277)
  const double zb = zR - 0.01;
  double rb[10];
  int igb[2];
  const bool have = GetRayTracingSolutionsBatch(1, &zR, &xR, &zT, ra, ig) == 0 && GetRayTracingSolutionsBatch(1, &zb, &xR, &zT, rb, igb) == 0;
  for (int k = 0; k < 2; k++)
    if (!have || (ra[6 + k] != -1000 && rb[6 + k] != -1000)) focusing[k] = out[k];
  if (zR == zT && focusing[0] == 0) focusing[0] = 1.;
}

void SetNumberOfAntennas(int numberOfAntennas) {
  detail::State &s = detail::state();
  for (size_t i = numberOfAntennas > 0 ? (size_t)numberOfAntennas : 0; i < s.tables.size(); i++) airice_inice_table_destroy(s.tables[i]);
  s.tables.resize(numberOfAntennas > 0 ? numberOfAntennas : 0, nullptr);
}
void MakeTable(double ShowerHitDistance, double ShowerDepth, double zR, int AntNum) {
  detail::State &s = detail::state();
  if (!detail::ensure_ctx()) return;
  if (AntNum < 0 || AntNum >= (int)s.tables.size()) { std::cerr << "IceRayTracing (B200): MakeTable: call SetNumberOfAntennas first" << std::endl; return; }
  if (s.tables[AntNum]) { airice_inice_table_destroy(s.tables[AntNum]); s.tables[AntNum] = nullptr; }
  if (airice_inice_table_create(s.ctx, ShowerHitDistance, ShowerDepth, zR, GridStepSizeX_O, GridStepSizeZ_O, GridWidthX, GridWidthZ,
                                &s.tables[AntNum]) != 0) {
    std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
    s.tables[AntNum] = nullptr;
  }
}
int GetInterpolatedValueBatch(long n, const double *xT, const double *zT, int rtParameter, int AntNum, double *out) {
  detail::State &s = detail::state();
  if (AntNum < 0 || AntNum >= (int)s.tables.size() || !s.tables[AntNum]) return 1;
  int rc = airice_inice_table_interp_host(s.ctx, s.tables[AntNum], n, xT, zT, rtParameter, out);
  if (rc != 0) std::cerr << "IceRayTracing (B200): " << airice_last_error() << std::endl;
  return rc;
}
double GetInterpolatedValue(double xT, double zT, int rtParameter, int AntNum) {
  double out = -1000;
  if (GetInterpolatedValueBatch(1, &xT, &zT, rtParameter, AntNum, &out) != 0) return -1000;
  return out;
}
int GetTableColumn(int AntNum, int col, std::vector<double> &out) {
  detail::State &s = detail::state();
  if (AntNum < 0 || AntNum >= (int)s.tables.size() || !s.tables[AntNum]) return 1;
  int64_t info[3];
  airice_inice_table_info(s.tables[AntNum], info);
  out.resize(info[2]);
  return airice_inice_table_copy_column(s.tables[AntNum], col, out.data());
}

double *IceRayTracing(double x0, double z0, double x1, double z1, bool PlotRayPaths) {
  (void)PlotRayPaths;  // ray-path dumps are outside the hot path
  return IceRayTracing(x0, z0, x1, z1);
}

}  // namespace IceRayTracing
