// pywrap.cc -- libAirIceRayTracing.so: the reference python wrapper's C ABI on top of the B200 library.
//
// Replaces /root/reference/pythonwrapper/TraceIceToAir.C:5-79 (TraceIceToAir + extern "C" Py_TraceIceToAir), which
// the reference's ctypes loader binds with argtypes [c_double x4, c_double*10] (pythonwrapper/AirIceRayTracing.py:8).
// Same signature, units (metres / degrees, depth negative in ice), output slots, -1000 sentinels and stdout lines.
// Difference: the reference re-parses ./Atmosphere.dat and rebuilds a 23k-point spline on EVERY call
// (TraceIceToAir.C:25); here the context is created on first use and reused (AIRICE_ATMOSPHERE overrides the path).
#include <cstdlib>
#include <iostream>

#include "airice_b200.h"

namespace {
airice_ctx *g_ctx = nullptr;
bool ensure_ctx() {
  if (g_ctx) return true;
  const char *path = std::getenv("AIRICE_ATMOSPHERE");
  const char *dev = std::getenv("AIRICE_DEVICE");
  if (airice_create(path ? path : "Atmosphere.dat", AIRICE_VARIANT_PYWRAP, dev ? std::atoi(dev) : 0, &g_ctx) != 0) {
    std::cerr << "libAirIceRayTracing (B200): " << airice_last_error() << std::endl;
    g_ctx = nullptr;
    return false;
  }
  return true;
}
}  // namespace

extern "C" {

// Batched form (new): n (Tx height, distance) pairs against one receiver; out is [n][10] in the slot order below.
int Py_TraceIceToAirBatch(double AntennaDepth, double IceLayerHeight, long n, const double *AirTxHeight,
                          const double *HorizontalDistance, double *out) {
  if (!ensure_ctx()) return 1;
  // the six of the thirteen metre/degree values TraceIceToAir.C:31-68 reads: only these cross the PCIe link
  const size_t nn = (size_t)(n > 0 ? n : 1);
  double *cols = (double *)std::malloc(sizeof(double) * 6 * nn);
  unsigned char *ok = (unsigned char *)std::malloc(nn);
  double *x_air = cols, *theta = cols + nn, *recv = cols + 2 * nn, *p_air = cols + 3 * nn, *p_ice = cols + 4 * nn,
         *refr = cols + 5 * nn;
  double *want[AIRICE_SOLVE_COLS] = {nullptr};
  want[1] = x_air; want[5] = theta; want[6] = recv; want[9] = p_air; want[10] = p_ice; want[12] = refr;
  int rc = airice_solve_host_columns(g_ctx, n, AirTxHeight, HorizontalDistance, nullptr, AntennaDepth, IceLayerHeight,
                                     AIRICE_UNITS_M_DEG_C, want, ok);
  if (rc == 0) {
    for (long i = 0; i < n; i++) {
      double *o = out + 10 * i;
      if (ok[i]) {
        o[0] = AirTxHeight[i]; o[1] = HorizontalDistance[i];
        o[2] = p_ice[i];                // geometricalPathLengthInIce
        o[3] = p_air[i];                // geometricalPathLengthInAir
        o[4] = recv[i];                 // "launchAngle" after std::swap (TraceIceToAir.C:33)
        o[5] = 180 - theta[i];          // "receivedAngle" = 180 - air launch angle (TraceIceToAir.C:34)
        o[6] = x_air[i];                // horidist2interpnt = X_air
        o[7] = refr[i];                 // AngleOfIncidenceOnIce = refracted angle below the surface (AirIceRayTracing.cc:1081)
        o[8] = 0; o[9] = 0;
      } else {
        for (int k = 0; k < 10; k++) o[k] = -1000;
      }
    }
  } else {
    std::cerr << "libAirIceRayTracing (B200): " << airice_last_error() << std::endl;
  }
  std::free(cols); std::free(ok);
  return rc;
}

void Py_TraceIceToAir(double AntennaDepth, double IceLayerHeight, double AirTxHeight, double HorizontalDistance,
                      double ArrayParameters[10]) {
  if (Py_TraceIceToAirBatch(AntennaDepth, IceLayerHeight, 1, &AirTxHeight, &HorizontalDistance, ArrayParameters) != 0) {
    for (int k = 0; k < 10; k++) ArrayParameters[k] = -1000;
  }
  if (ArrayParameters[0] != -1000) {
    std::cout << " We have a solution!!!" << std::endl;
    std::cout << "AirTxHeight: " << AirTxHeight << std::endl;
    std::cout << "HorizontalDistance: " << HorizontalDistance << std::endl;
    std::cout << "geometricalPathLengthInIce: " << ArrayParameters[2] << std::endl;
    std::cout << "geometricalPathLengthInAir: " << ArrayParameters[3] << std::endl;
    std::cout << "launchAngle: " << ArrayParameters[4] << std::endl;
    std::cout << "RecievedAngle: " << ArrayParameters[5] << std::endl;
    std::cout << "horidist2interpnt: " << ArrayParameters[6] << std::endl;
    std::cout << "AngleOfIncidenceOnIce: " << ArrayParameters[7] << std::endl;
  } else {
    std::cout << " We do NOT have a solution!!!" << std::endl;
  }
}
}
