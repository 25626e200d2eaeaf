// airice_path.cuh -- ray-path emission: the (x, height) polyline of a forward-traced ray at 1 m height steps, i.e. the
// RayPathinAirnIce.txt dump of the reference's CLI (SingleRayAirIceRefraction.C:226-299; BASELINE config 1 writes
// 17 206 points), batched over rays.
//
// Semantics as written there: in every air layer the height runs i = start, start-1, ... while i > stop-1, the last
// value clamped to stop; x(i) = F(i) - F(start) + (x of the last point of the layer above), F the closed form fDnfR
// with the layer's B, C and the ray's L (the L of the first layer is carried through all layers and into the ice,
// SingleRayAirIceRefraction.C:133-152); the next layer starts 1e-5 m below the boundary.  In the ice the depth runs
// i = 0, -1, ... while i > -(depth+1) (no clamp: a non-integer depth is overshot, as in the reference) and
// x(i) = x_surface - F_ice(i) + F_ice(0), reported at height i + ice.
//
// Two steps: a per-ray plan (segments, point counts, x offsets; <= 6 closed-form evaluations) and a per-point fill
// (one closed-form evaluation and two 8-byte stores per point: HBM-write bound).
#pragma once
#include "airice_core.cuh"

#define AIRICE_PATH_MAX_SEGS (AIRICE_MAX_LAYERS + 1)

struct AirIcePathPlan {                 // one ray
  double L;
  double start[AIRICE_PATH_MAX_SEGS];   // first height of the segment (ice: 0)
  double stop[AIRICE_PATH_MAX_SEGS];    // clamp height (ice: unused)
  double x0[AIRICE_PATH_MAX_SEGS];      // x of the last point of the segment above
  double f0[AIRICE_PATH_MAX_SEGS];      // F(start)
  double pre[AIRICE_PATH_MAX_SEGS];     // (L/C) (1/sqrt(A^2-L^2)) of the segment's medium: the ray-constant factor of F
  double sA[AIRICE_PATH_MAX_SEGS];      // sqrt(A^2-L^2)
  int first[AIRICE_PATH_MAX_SEGS + 1];  // index of the segment's first point; first[nseg] = total points
  int layer[AIRICE_PATH_MAX_SEGS];      // air layer index, or AIRICE_ICE_SLOT
  int nseg;
};

// fDnfR as the CLI calls it in air (RayTracingFunctions, x = -height, C = +C_layer; identical in value to the M.cc
// convention x = +height, C = -C_layer) and in the ice (x = i <= 0, C = +C_ice), split into the factor that depends on
// the ray only -- evaluated once per segment in the plan -- and the part that depends on the point:
//   F = [(L/C) (1/sqrt(A^2-L^2))] * (C x - log(A n - L^2 + sqrt(A^2-L^2) sqrt(n^2-L^2)))
AIRICE_HD void airice_path_F_factor(double A, double C, double L, double& pre, double& sA) {
  sA = sqrt(A * A - L * L);
  pre = (L / C) * (1.0 / sA);
}
// exp / log per point: glibc's own algorithms on the device (airice_glibc_math.cuh: the reference's libm, and with their
// coefficients in the constant bank about half the instructions of CUDA's versions); one exp, one log, one sqrt per point
#if defined(__CUDA_ARCH__)
#define AIRICE_PATH_EXP(x) airice_glibc_exp(x)
#define AIRICE_PATH_LOG(x) airice_glibc_log(x)
#else
#define AIRICE_PATH_EXP(x) exp(x)
#define AIRICE_PATH_LOG(x) log(x)
#endif
AIRICE_HD double airice_path_F_air(const AirIceMedium& m, int k, double L, double pre, double sA, double height) {
  const double C = m.C[k], x = -height;
  const double n = 1.0 + m.B[k] * AIRICE_PATH_EXP(-m.C[k] * fabs(x));
  return pre * (C * x - AIRICE_PATH_LOG(1.0 * n - L * L + sA * sqrt(n * n - L * L)));
}
AIRICE_HD double airice_path_F_ice(const AirIceMedium& m, double L, double pre, double sA, double i) {
  const double A = m.A_ice, C = m.C_ice;
  const double n = A + m.B_ice * AIRICE_PATH_EXP(-C * fabs(i));
  return pre * (C * i - AIRICE_PATH_LOG(A * n - L * L + sA * sqrt(n * n - L * L)));
}

// number of loop trips of "for (i = start; i > stop - 1; i = i - 1)": the k >= 0 with start - k > stop - 1
AIRICE_HD int airice_path_trips(double start, double stop) {
  if (!(start > stop - 1)) return 0;
  double c = ceil((start - stop) + 1.0);
  if (!(c < 2.0e9)) return -1;
  int k = (int)c;
  while (k > 0 && !(start - (double)(k - 1) > stop - 1)) k--;   // rounding guards: k-1 is the last trip, k is not one
  while (start - (double)k > stop - 1) k++;
  return k;
}

// height of point q of an air segment
AIRICE_HD double airice_path_height(double start, double stop, int q) {
  double i = start - (double)q;
  if (i < stop) i = stop;
  return i;
}

// Plan of one ray launched at `theta` (deg from the upward vertical, > 90) from height h.  Returns the number of
// points (0: the ray does not exist -- Tx outside the layers or below the surface, or L >= 1; -1: more than 2e9).
AIRICE_HD int airice_path_plan(const AirIceMedium& m, const AirIcePlan& p, double theta, double h, AirIcePathPlan& pl) {
  pl.nseg = 0; pl.first[0] = 0; pl.L = 0.0;
  const int kt = airice_top_layer(m, h);
  if (kt < 0 || kt < p.kb || !(h >= p.ice_h)) return 0;
  const double n_tx = airice_n_air(m, kt, h);
  const double L = n_tx * sin((180 - theta) * m.deg2rad);
  if (!(1.0 - L * L > 0.0) || !(L == L)) return 0;
  pl.L = L;
  double last_x = 0.0;
  int total = 0, s = 0;
  for (int k = kt; k >= p.kb; k--, s++) {
    const double start = (k == kt) ? h : p.seg[k].start_x, stop = p.seg[k].stop_x;
    const int trips = airice_path_trips(start, stop);
    if (trips < 0 || total > 2000000000 - trips) return -1;
    pl.start[s] = start; pl.stop[s] = stop; pl.x0[s] = last_x; pl.layer[s] = k;
    airice_path_F_factor(1.0, m.C[k], L, pl.pre[s], pl.sA[s]);
    pl.f0[s] = airice_path_F_air(m, k, L, pl.pre[s], pl.sA[s], start);
    pl.first[s] = total;
    total += trips;
    if (trips > 0)
      last_x = airice_path_F_air(m, k, L, pl.pre[s], pl.sA[s], airice_path_height(start, stop, trips - 1)) - pl.f0[s] + last_x;
  }
  if (p.has_ice) {
    const int trips = airice_path_trips(0.0, -p.depth);        // i = 0, -1, ... while i > -(depth + 1)
    if (trips < 0 || total > 2000000000 - trips) return -1;
    pl.start[s] = 0.0; pl.stop[s] = -p.depth; pl.x0[s] = last_x; pl.layer[s] = AIRICE_ICE_SLOT;
    airice_path_F_factor(m.A_ice, m.C_ice, L, pl.pre[s], pl.sA[s]);
    pl.f0[s] = airice_path_F_ice(m, L, pl.pre[s], pl.sA[s], 0.0);
    pl.first[s] = total;
    total += trips;
    s++;
  }
  pl.nseg = s;
  pl.first[s] = total;
  return total;
}

// point q (0 <= q < total) of a planned ray
AIRICE_HD void airice_path_point(const AirIceMedium& m, const AirIcePlan& p, const AirIcePathPlan& pl, int q, double& x, double& z) {
  int s = 0;
  while (s + 1 < pl.nseg && q >= pl.first[s + 1]) s++;
  const int local = q - pl.first[s];
  if (pl.layer[s] == AIRICE_ICE_SLOT) {
    const double i = 0.0 - (double)local;
    x = pl.x0[s] - airice_path_F_ice(m, pl.L, pl.pre[s], pl.sA[s], i) + pl.f0[s];
    z = i + p.ice_h;
  } else {
    const double i = airice_path_height(pl.start[s], pl.stop[s], local);
    x = airice_path_F_air(m, pl.layer[s], pl.L, pl.pre[s], pl.sA[s], i) - pl.f0[s] + pl.x0[s];
    z = i;
  }
}
