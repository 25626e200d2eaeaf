// IceRayTracing.cc -- host-side mirror of the reference's in-ice entry point on top of include/airice_b200.h.
#include "IceRayTracing.hh"

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "airice_b200.h"

namespace IceRayTracing {

double A_ice = 1.78, B_ice = -0.43, C_ice = 0.0132;

namespace detail {
struct State {
  airice_ctx *ctx = nullptr;
  int device = 0;
  std::string atmosphere = "Atmosphere.dat";
  double ice_set[3] = {1.78, -0.43, 0.0132};
};
inline State &state() {
  static State s;
  return s;
}
inline bool ensure_ctx() {
  State &s = state();
  if (!s.ctx) {
    const char *env = std::getenv("AIRICE_ATMOSPHERE");
    if (airice_create(env ? env : s.atmosphere.c_str(), AIRICE_VARIANT_MULTIRAY, s.device, &s.ctx) != 0) {
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

void GetRayTracingSolutions(double RxDepth, double Distance, double TxDepth, double TimeRay[2], double PathRay[2],
                            double LaunchAngle[2], double RecieveAngle[2], int IgnoreCh[2], double IncidenceAngleInIce[2],
                            double A0, double frequency, double AttRay[2]) {
  (void)A0; (void)frequency;
  double out[10];
  int ig[2] = {0, 0};
  if (GetRayTracingSolutionsBatch(1, &RxDepth, &Distance, &TxDepth, out, ig) != 0) {
    for (int k = 0; k < 10; k++) out[k] = 0;
    out[6] = out[7] = -1000;
  }
  for (int k = 0; k < 2; k++) {
    TimeRay[k] = out[0 + k]; PathRay[k] = out[2 + k]; LaunchAngle[k] = out[4 + k]; RecieveAngle[k] = out[6 + k];
    IncidenceAngleInIce[k] = out[8 + k]; IgnoreCh[k] = ig[k];
    if (AttRay) AttRay[k] = 0;
  }
}

double *IceRayTracing(double x0, double z0, double x1, double z1, bool PlotRayPaths) {
  (void)PlotRayPaths;  // ray-path dumps are outside the hot path
  return IceRayTracing(x0, z0, x1, z1);
}

}  // namespace IceRayTracing
