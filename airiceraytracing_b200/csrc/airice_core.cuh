// airice_core.cuh -- closed-form air->ice ray arithmetic shared by the three CUDA kernels.
//
// Everything here is FP64 scalar math on registers; the medium model and the per-launch plan arrive
// as kernel parameters (constant bank), so a thread touches HBM only for its own inputs/outputs.
//
// Reference behaviour being reproduced (file:line under /root/reference):
//   n(h) of air / ice ............ MultiRayAirIceRefraction.cc:150-263
//   fDnfR / ftimeD / fpathD ...... MultiRayAirIceRefraction.cc:377-447
//   layer walk ................... MultiRayAirIceRefraction.cc:661-804 (solver), 1796-1879 (table)
//   ice leg ...................... MultiRayAirIceRefraction.cc:807-869, 1893-1922
//   Fresnel Trans_S/Trans_P ...... MultiRayAirIceRefraction.cc:285-337
//
// Algebra used (exact identities of the reference formulas, n = A + B e^{C'x}, C' = -C_layer,
// sA = sqrt(A^2-L^2), D = n^2-L^2, R = sqrt(D), T = A n - L^2 + sA R, G = C'x - ln T, H = ln(n+R)):
//   fDnfR  = (L/C')(1/sA) G
//   ftimeD = (1/(c C' R)) (D + G A^2 R / sA + A R H)
//   fpathD = (H + (A/sA) G) / C'
// so one layer end costs 1 sqrt + 1 log (distance only) or 1 sqrt + 2 log (distance, time, path), and
// n at every layer end except the transmitter is ray-independent (host-computed into AirIcePlan).
// The reference re-evaluates n(x) 9x in ftimeD and exp() 13x in fpathD; nothing of that survives.
//
// Code-size discipline: a ray is a list of at most 6 segments (<=5 air layers, then the ice leg) that all run
// through ONE loop body (never unrolled), so the hot loop of each kernel is a few hundred instructions and
// stays resident in the SM's instruction cache.  (The first version unrolled the layer loop and inlined every
// evaluation: 230 KB of SASS, and ncu showed 10 "no instruction" stall cycles per issued instruction.)
#pragma once
#include <math.h>
#include <stdint.h>

#include "airice_math.cuh"
#include "airice_glibc_math.cuh"

#if defined(__CUDACC__)
#define AIRICE_HD __host__ __device__ __forceinline__
#define AIRICE_HD_NOINLINE __host__ __device__ __noinline__
#else
#define AIRICE_HD inline
#define AIRICE_HD_NOINLINE
#endif

#define AIRICE_MAX_LAYERS 5
#ifndef AIRICE_PEEL_TOP
#define AIRICE_PEEL_TOP 1
#endif
#ifndef AIRICE_UNROLL_XFAST
#define AIRICE_UNROLL_XFAST 1
#endif
#ifndef AIRICE_UNROLL_FULL
#define AIRICE_UNROLL_FULL 1
#endif
#ifndef AIRICE_UNROLL_RELAY
#define AIRICE_UNROLL_RELAY 2
#endif

// Product rounded on its own (never contracted into an FMA with a following add).  Used where the reference forms
// F(stop)-F(start) from two separately rounded products: for a zero-thickness segment (Tx exactly on the ice surface,
// the last table row) the reference gets an exact 0, and a*b-c*d fused as fma(a,b,-(c*d)) would leave rounding dust.
#if defined(__CUDA_ARCH__)
#define AIRICE_MUL(a, b) __dmul_rn((a), (b))
#else
#define AIRICE_MUL(a, b) ((a) * (b))
#endif

// Per-context medium model (one Atmosphere.dat + ice model + which copy of the reference, i.e. which pi).
struct AirIceMedium {
  int nlayers;                        // MaxLayers (M.cc:142)
  int variant;                        // 0 = MultiRayAirIceRefraction, 1 = pythonwrapper/AirIceRayTracing, 2 = the CLIs'
                                      // RayTracingFunctions copy (Air2IceRayTracing.C: Brent, its own bracket rule)
  double hlo[AIRICE_MAX_LAYERS + 1];  // ATMLAY[k]/100 in metres; hlo[nlayers] closes the top layer
  double B[AIRICE_MAX_LAYERS];        // B_air
  double C[AIRICE_MAX_LAYERS];        // C_air
  double A_ice, B_ice, C_ice;
  double pi;                          // 3.1415927 (M.h:29) or 4*atan(1) (P.h:25)
  double deg2rad;                     // pi/180.0, rounded once like the reference expression
  double rad2deg;                     // 180/pi
  double c;                           // 299792458 (M.h:30)
  double tan16;                       // tan(16 deg) in the variant's pi: tangent step of the 16-deg bracket (M.cc:1487)
  // the clamped bracket's scan (M.cc:1490-1511) visits lo_j = 90.001 (+ 0.05 j times, accumulated as the reference
  // does) and needs sin((180 - lo_j) deg2rad): both are ray-independent, tabulated on the host with the libm the
  // reference uses.  [AIRICE_CLAMP_N][2] = {lo_j, sin_j}; device memory on the GPU, host memory in host builds.
  const double* clamp_tab;
};
#define AIRICE_CLAMP_N 336             // lo_j up to 106.75 deg: the scan never goes past hi - 0.1 < 105.901
#if defined(__CUDA_ARCH__)
#define AIRICE_LDG(p) __ldg(p)
#else
#define AIRICE_LDG(p) (*(p))
#endif

// Per-(ice height, receiver depth) plan: every ray-independent number of the layer walk, computed on
// the host with the same libm the reference uses.  Slot AIRICE_MAX_LAYERS of the per-segment arrays is the ice leg.
#define AIRICE_ICE_SLOT AIRICE_MAX_LAYERS
// One segment (air layer k, or the ice leg in slot AIRICE_ICE_SLOT).  The values a loop trip reads together sit together
// and 16-byte aligned, so that a trip's plan reads are 128-bit constant-bank loads (as separate arrays every value was
// its own LDC: 9.5 % of the solve kernel's instructions).
struct alignas(16) AirIceSeg {
  double neg_c, inv_neg_c;    // C' = -C of the segment's medium, and 1/C'
  double stop_x, stop_n;      // lower end: ice_h for k==kb else hlo[k] (M.cc:722-728); ice leg: depth
  double start_x, start_n;    // upper end when entered from above: hlo[k+1]-1e-5 (M.cc:715); ice leg: 0
  // solver path (one L throughout): n_stop[k] - n_start[k-1] and n_stop[k]^2 - n_start[k-1]^2 across the lower boundary
  // of layer k (~4e-13; 0 for k <= kb): first-order hand-over of R, ln T, H in airice_ray_air<false>
  double ho_dn, ho_dn2;
  double relay, ln_relay;     // n_start[k]/n_stop[k+1]: Snell hand-over of the table path (M.cc:1871), and its log (~3e-13)
  // single-precision companions for the FP32 pre-iteration of the solver (host-computed in double, then rounded):
  // q = n^2 - A^2 and pa = A (n - A) at both ends of a segment keep the small differences (n-1 ~ 3e-4 in air) exact,
  // so that R^2 = q + sA^2 and T = pa + sA (sA + R) stay accurate in float even for grazing rays.
  float f_q_stop, f_pa_stop, f_q_start, f_pa_start;
  float f_cdx, f_inv_neg_c;   // C' (x_stop - x_start) of a full segment, (float)(1/C')
  float pad_[2];
};
struct AirIcePlan {
  int kb;          // layer that contains the ice surface (= SkipLayersBelow, M.cc:680-690)
  int has_ice;     // receiver below the surface (depth != 0), M.cc:901
  double ice_h;    // ice-surface height after the depth>=0 fold (M.cc:1472-1476)
  double depth;    // receiver depth, positive, 0 when the receiver sits in air
  AirIceSeg seg[AIRICE_MAX_LAYERS + 1];
};

AIRICE_HD double airice_n_air(const AirIceMedium& m, int k, double z) { return 1.0 + m.B[k] * exp(-m.C[k] * z); }
// n(h) of a transmitter as the reference's x86 build rounds it (Getnz_air, M.cc:258: 1 + B exp(-C h) with the product
// and the sum rounded separately, exp from glibc): on the device glibc's own exp (airice_glibc_math.cuh, 25 instructions
// with constant-bank coefficients against 53 for CUDA's) and an unfused product, so that L = n sin(theta) starts from the
// reference's bits.
AIRICE_HD double airice_n_tx(const AirIceMedium& m, int k, double h) {
#if defined(__CUDA_ARCH__)
  return 1.0 + AIRICE_MUL(m.B[k], airice_glibc_exp(-m.C[k] * h));
#else
  return 1.0 + m.B[k] * exp(-m.C[k] * h);
#endif
}

// Top layer of a transmitter height for the walk (SkipLayersAbove, M.cc:666-676); -1 = in no layer.
AIRICE_HD int airice_top_layer(const AirIceMedium& m, double h) {
  // the k with hlo[k] <= h < hlo[k+1] (half-open, ">=" on the lower edge): the layer edges ascend, so inside
  // [hlo[0], hlo[nlayers]) that k is the number of interior edges at or below h
  if (!(h >= m.hlo[0] && h < m.hlo[m.nlayers])) return -1;
  int kt = 0;
#pragma unroll
  for (int k = 1; k < AIRICE_MAX_LAYERS; k++) kt += (k < m.nlayers && h >= m.hlo[k]) ? 1 : 0;
  return kt;
}

struct AirIceRay {   // everything the reference reports for one ray (metres, seconds, degrees)
  double x_air, x_ice, t_air, t_ice, p_air, p_ice;
  double inc_ice_deg, recv_deg, refr_deg;
  double trans_s, trans_p;
};

// Horizontal distance X(L) with reference-like rounding (the root function's X, MinimizeforLaunchAngle M.cc:873-917):
// F(stop) and F(start) are formed and rounded separately as GetRayHorizontalPath does (M.cc:463).  The solver path
// carries L unchanged through all layers and into the ice (M.cc:757-771, 894-902).  Used for the rare real
// evaluations inside the bisection replay, where the SIGN of d - X must agree with the reference's.
AIRICE_HD double airice_x_exact(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx, double L,
                                bool with_ice = true) {
  const double L2 = L * L;
  const double sAir = AIRICE_SQRT(1.0 * 1.0 - L2), sIce = AIRICE_SQRT(m.A_ice * m.A_ice - L2);
  const int nair = (kt >= p.kb) ? (kt - p.kb + 1) : 0;
  const int nseg = nair + ((p.has_ice && with_ice) ? 1 : 0);
  double X = 0.0;
#pragma unroll 1
  for (int j = 0; j < nseg; j++) {
    const bool air = j < nair;
    const int k = air ? (kt - j) : AIRICE_ICE_SLOT;
    const double A = air ? 1.0 : m.A_ice;
    const double sA = air ? sAir : sIce;
    const double Cn = p.seg[k].neg_c;
    const bool top = (j == 0) && air;
    const double xt = top ? h : p.seg[k].start_x;
    const double nt = top ? n_tx : p.seg[k].start_n;
    const double xb = p.seg[k].stop_x, nb = p.seg[k].stop_n;
    const double Rb = AIRICE_SQRT(nb * nb - L2), Rt = AIRICE_SQRT(nt * nt - L2);
    const double Gb = Cn * xb - AIRICE_LOG_POS(A * nb - L2 + sA * Rb), Gt = Cn * xt - AIRICE_LOG_POS(A * nt - L2 + sA * Rt);
    const double mult = AIRICE_DIV(L, Cn) * AIRICE_RCP(sA);
    const double seg = AIRICE_MUL(mult, Gb) - AIRICE_MUL(mult, Gt);
    X += air ? -seg : seg;
  }
  return X;
}

// One segment of X(L): (L / (C' sA)) (C' (xb - xt) - ln(T_b / T_t)).  AIR fixes A = 1 at compile time (the products
// A n are then exact copies of n, so the values are those of the generic form).
template <bool AIR>
AIRICE_HD double airice_seg_x(double A, double sA, double inv_sA, double L, double L2, double Cn, double iC, double xt,
                              double nt, double xb, double nb) {
  const double Rb = AIRICE_SQRT(nb * nb - L2), Rt = AIRICE_SQRT(nt * nt - L2);
  const double Tb = (AIR ? nb : A * nb) - L2 + sA * Rb, Tt = (AIR ? nt : A * nt) - L2 + sA * Rt;
  const double dG = Cn * (xb - xt) - AIRICE_LOG_POS(Tb * AIRICE_RCP(Tt));
  return (L * (iC * inv_sA)) * dG;
}

// X(L) only, FP64, arranged like airice_x_dx below but without derivative terms: the work horse of the solver's
// chord iteration (the slope comes from the FP32 pre-iteration or from a secant).  The air layers run BOTTOM-UP:
// every lane of a warp starts at the surface layer kb (uniform), so the per-layer plan values are warp-uniform
// constant-bank reads (top-down, lanes whose transmitters sit in different layers read different slots in the same
// trip and the reads serialise); the transmitter's own layer is the last trip of each lane.  The ice leg is a
// separate instance of the segment, which removes the per-trip air/ice selects.
AIRICE_HD double airice_x_fast(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx, double L) {
  const double L2 = L * L;
  double X = 0.0;
  if (kt >= p.kb) {
    const double sAir = AIRICE_SQRT(1.0 * 1.0 - L2);
    const double yAir = AIRICE_RCP(sAir);
#pragma unroll 1
    for (int k = p.kb; k <= kt; k++) {
      const bool top = (k == kt);
      const double xt = top ? h : p.seg[k].start_x;
      const double nt = top ? n_tx : p.seg[k].start_n;
      X -= airice_seg_x<true>(1.0, sAir, yAir, L, L2, p.seg[k].neg_c, p.seg[k].inv_neg_c, xt, nt, p.seg[k].stop_x, p.seg[k].stop_n);
    }
  }
  if (p.has_ice) {
    const double sIce = AIRICE_SQRT(m.A_ice * m.A_ice - L2);
    const double yIce = AIRICE_RCP(sIce);
    const int k = AIRICE_ICE_SLOT;
    X += airice_seg_x<false>(m.A_ice, sIce, yIce, L, L2, p.seg[k].neg_c, p.seg[k].inv_neg_c, p.seg[k].start_x, p.seg[k].start_n,
                             p.seg[k].stop_x, p.seg[k].stop_n);
  }
  return X;
}

// X(L) and dX/dL in FP64, arranged like airice_x_fast (bottom-up air layers, separate ice leg): the ONE evaluation the
// solver normally spends after the single-precision pre-iteration -- with an analytic slope a Newton step from a point
// ~1e-6 deg from the root lands ~1e-13 deg from it, which a chord step with the single-precision slope cannot do.
// dG/dL = L (sA+R)^2 / (T sA R) per end follows from dT/dL = -L (sA+R)^2 / (sA R); 1/R comes from the square root's own
// refined seed (2^-40), which is ample for a slope.
template <bool AIR>
AIRICE_HD void airice_seg_x_dx(double A, double sA, double inv_sA, double L, double L2, double Cn, double iC, double xt,
                               double nt, double xb, double nb, double& seg, double& dseg) {
  double Rb, yb, Rt, yt;
  AIRICE_SQRT_RSQRT_NZ(nb * nb - L2, Rb, yb);
  AIRICE_SQRT_RSQRT_NZ(nt * nt - L2, Rt, yt);
  const double Tb = (AIR ? nb : A * nb) - L2 + sA * Rb, Tt = (AIR ? nt : A * nt) - L2 + sA * Rt;
  // 1/T_b only enters the slope, but the 20-bit MUFU seed is NOT enough for it: a slope off by 1e-6 moves the Newton
  // step by 1e-6 of its length, which for the pairs whose single-precision landing point is 1e-4 deg off exceeds the
  // replay's guard band (measured: one bisection-cell miss in 128 000 solves, tests/test_gpu_parity.py C5 case)
  const double rTt = AIRICE_RCP(Tt), rTb = AIRICE_RCP(Tb);
  const double dG = Cn * (xb - xt) - AIRICE_LOG_POS(Tb * rTt);
  const double c1 = iC * inv_sA;
  seg = (L * c1) * dG;
  const double qb = (sA + Rb) * (sA + Rb) * (rTb * yb), qt = (sA + Rt) * (sA + Rt) * (rTt * yt);
  const double AA = AIR ? 1.0 : A * A;
  dseg = c1 * (AA * inv_sA * inv_sA * dG + L2 * inv_sA * (qb - qt));
}
AIRICE_HD double airice_x_dx(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx, double L,
                             double& dXdL) {
  const double L2 = L * L;
  double X = 0.0, dX = 0.0;
  if (kt >= p.kb) {
    const double sAir = AIRICE_SQRT_NZ(1.0 * 1.0 - L2);
    const double yAir = AIRICE_RCP(sAir);
#if AIRICE_PEEL_TOP
#pragma unroll 1
    for (int k = p.kb; k < kt; k++) {
      double seg, dseg;
      airice_seg_x_dx<true>(1.0, sAir, yAir, L, L2, p.seg[k].neg_c, p.seg[k].inv_neg_c, p.seg[k].start_x, p.seg[k].start_n, p.seg[k].stop_x, p.seg[k].stop_n, seg, dseg);
      X -= seg; dX -= dseg;
    }
    {
      double seg, dseg;
      airice_seg_x_dx<true>(1.0, sAir, yAir, L, L2, p.seg[kt].neg_c, p.seg[kt].inv_neg_c, h, n_tx, p.seg[kt].stop_x, p.seg[kt].stop_n, seg, dseg);
      X -= seg; dX -= dseg;
    }
#else
#pragma unroll 1
    for (int k = p.kb; k <= kt; k++) {
      const bool top = (k == kt);
      const double xt = top ? h : p.seg[k].start_x;
      const double nt = top ? n_tx : p.seg[k].start_n;
      double seg, dseg;
      airice_seg_x_dx<true>(1.0, sAir, yAir, L, L2, p.seg[k].neg_c, p.seg[k].inv_neg_c, xt, nt, p.seg[k].stop_x, p.seg[k].stop_n, seg, dseg);
      X -= seg; dX -= dseg;
    }
#endif
  }
  if (p.has_ice) {
    const double sIce = AIRICE_SQRT_NZ(m.A_ice * m.A_ice - L2);
    const double yIce = AIRICE_RCP(sIce);
    const int k = AIRICE_ICE_SLOT;
    double seg, dseg;
    airice_seg_x_dx<false>(m.A_ice, sIce, yIce, L, L2, p.seg[k].neg_c, p.seg[k].inv_neg_c, p.seg[k].start_x, p.seg[k].start_n, p.seg[k].stop_x,
                           p.seg[k].stop_n, seg, dseg);
    X += seg; dX += dseg;
  }
  dXdL = dX;
  return X;
}

// Single-precision X(t) and dX/dt, t = tan(incidence at the transmitter): the pre-iteration that brings the solver
// within ~1e-4 deg of the root on the FP32/MUFU pipes, which this FP64-bound kernel leaves idle.  dn_tx = n(h_Tx) - 1.
#if defined(__CUDA_ARCH__)
// The bare MUFU approximations (1-2 ulp, flush-to-zero): rsqrtf / __frcp_rn / __logf / sqrtf wrap them in subnormal
// scaling or a correctly rounded refinement, 5-8 instructions each and 19 layer ends per pair -- 6 % of the solve
// kernel's instructions for accuracy a pre-iteration with a 3e-5 slope allowance has no use for.
__device__ __forceinline__ float airice_f_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float airice_f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float airice_f_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float airice_f_log(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r * 0.693147180559945f; }
#define AIRICE_F_RSQRT(x) airice_f_rsqrt(x)
#define AIRICE_F_RCP(x) airice_f_rcp(x)
#define AIRICE_F_LOG(x) airice_f_log(x)
#define AIRICE_F_SQRT(x) airice_f_sqrt(x)
#else
#define AIRICE_F_RSQRT(x) (1.0f / sqrtf(x))
#define AIRICE_F_RCP(x) (1.0f / (x))
#define AIRICE_F_LOG(x) logf(x)
#define AIRICE_F_SQRT(x) sqrtf(x)
#endif
// One layer END in single precision: R = sqrt(q + sA^2), y = 1/R, T = pa + sA (sA + R), rT = 1/T and the derivative
// term (sA + R)^2 / (T R).
struct AirIceEndF32 { float T, rT, g; };
AIRICE_HD AirIceEndF32 airice_end_f32(float sA2, float sA, float q, float pa) {
  const float R2 = q + sA2;
  const float y = AIRICE_F_RSQRT(R2);
  const float R = R2 * y;
  AirIceEndF32 e;
  e.T = pa + sA * (sA + R);
  e.rT = AIRICE_F_RCP(e.T);
  e.g = (sA + R) * (sA + R) * (e.rT * y);
  return e;
}
// segment sums from its two ends: X += +-(L c1) dG, dX += +-c1 (A^2 dG / sA^2 + L^2 (g_b - g_t) / sA)
template <bool AIR>
AIRICE_HD void airice_seg_f32(float A, float inv_sA, float L, float L2, const AirIceEndF32& eb, const AirIceEndF32& et,
                              float cdx, float icn, float& X, float& dX) {
  const float dG = cdx - AIRICE_F_LOG(eb.T * et.rT);
  const float c1 = icn * inv_sA;
  const float seg = (L * c1) * dG;
  const float AA = AIR ? 1.0f : A * A;
  const float dseg = c1 * (AA * inv_sA * inv_sA * dG + L2 * inv_sA * (eb.g - et.g));
  if (AIR) { X -= seg; dX -= dseg; } else { X += seg; dX += dseg; }
}

AIRICE_HD float airice_x_newton_f32(const AirIceMedium& m, const AirIcePlan& p, int kt, float h_minus_stop_top_cn,
                                    float dn_tx, float t, float& dXdt) {
  const float n_tx = 1.0f + dn_tx;
  const float q_tx = dn_tx * (2.0f + dn_tx);                  // n_tx^2 - 1
  const float w2 = AIRICE_F_RCP(1.0f + t * t), w = AIRICE_F_SQRT(w2);
  const float L = n_tx * t * w, L2 = L * L;
  float X = 0.0f, dX = 0.0f;
  if (kt >= p.kb) {
    const float sA2 = w2 * (1.0f - t * t * q_tx);             // 1 - L^2 without cancellation
    const float y = AIRICE_F_RSQRT(sA2), sA = sA2 * y;
    // bottom-up (see airice_x_fast).  The upper end of a layer and the lower end of the layer above it are the same
    // point to single precision (1e-5 m and 3e-13 in n apart), so each trip evaluates ONE end and keeps it for the next.
    AirIceEndF32 eb = airice_end_f32(sA2, sA, p.seg[p.kb].f_q_stop, p.seg[p.kb].f_pa_stop);
#if AIRICE_PEEL_TOP
#pragma unroll 1
    for (int k = p.kb; k < kt; k++) {
      const AirIceEndF32 et = airice_end_f32(sA2, sA, p.seg[k].f_q_start, p.seg[k].f_pa_start);
      airice_seg_f32<true>(1.0f, y, L, L2, eb, et, p.seg[k].f_cdx, p.seg[k].f_inv_neg_c, X, dX);
      eb = et;
    }
    {
      const AirIceEndF32 et = airice_end_f32(sA2, sA, q_tx, dn_tx);
      airice_seg_f32<true>(1.0f, y, L, L2, eb, et, h_minus_stop_top_cn, p.seg[kt].f_inv_neg_c, X, dX);
    }
#else
#pragma unroll 1
    for (int k = p.kb; k <= kt; k++) {
      const bool top = (k == kt);
      const float qt = top ? q_tx : p.seg[k].f_q_start, pat = top ? dn_tx : p.seg[k].f_pa_start;
      const float cdx = top ? h_minus_stop_top_cn : p.seg[k].f_cdx;
      const AirIceEndF32 et = airice_end_f32(sA2, sA, qt, pat);
      airice_seg_f32<true>(1.0f, y, L, L2, eb, et, cdx, p.seg[k].f_inv_neg_c, X, dX);
      eb = et;
    }
#endif
  }
  if (p.has_ice) {
    const float Ai = (float)m.A_ice;
    const float sA2 = Ai * Ai - L2;
    const float y = AIRICE_F_RSQRT(sA2), sA = sA2 * y;
    const int k = AIRICE_ICE_SLOT;
    const AirIceEndF32 eb = airice_end_f32(sA2, sA, p.seg[k].f_q_stop, p.seg[k].f_pa_stop);
    const AirIceEndF32 et = airice_end_f32(sA2, sA, p.seg[k].f_q_start, p.seg[k].f_pa_start);
    airice_seg_f32<false>(Ai, y, L, L2, eb, et, p.seg[k].f_cdx, p.seg[k].f_inv_neg_c, X, dX);
  }
  dXdt = dX * n_tx * w2 * w;   // dL/dt = n_tx / (1+t^2)^{3/2}
  return X;
}

// Distance, time and geometric path of one segment from the quantities at its two ends (R = sqrt(n^2 - L^2),
// ln T, H = ln(n + R)); products rounded as the reference forms them.  AIR fixes A = 1 at compile time (exact).
template <bool AIR>
AIRICE_HD void airice_seg_sums(double A, double inv_sA, double mult, double cC, double Cn, double iC, double xt, double xb,
                               double Dt, double Db, double Rt, double Rb, double lnTt, double lnTb, double Ht, double Hb,
                               double& xs, double& ts, double& gs) {
  const double Gb = Cn * xb - lnTb, Gt = Cn * xt - lnTt;
  xs = AIRICE_MUL(mult, Gb) - AIRICE_MUL(mult, Gt);
  const double AARb = AIR ? Rb : A * A * Rb, AARt = AIR ? Rt : A * A * Rt;
  const double ARb = AIR ? Rb : A * Rb, ARt = AIR ? Rt : A * Rt;
  const double Ay = AIR ? inv_sA : A * inv_sA;
  const double tb = AIRICE_MUL(AIRICE_RCP(cC * Rb), (Db + AIRICE_MUL(Gb * AARb, inv_sA)) + ARb * Hb);
  const double tt = AIRICE_MUL(AIRICE_RCP(cC * Rt), (Dt + AIRICE_MUL(Gt * AARt, inv_sA)) + ARt * Ht);
  ts = tb - tt;
  gs = AIRICE_MUL(Hb + Ay * Gb, iC) - AIRICE_MUL(Ht + Ay * Gt, iC);
}

// incidence angle on the surface, refracted angle below it and the Fresnel coefficients (air side only)
AIRICE_HD void airice_ray_surface(const AirIceMedium& m, const AirIcePlan& p, int kt, double Lk, double Rsurf, bool want_inc,
                                  bool want_refr, AirIceRay& r) {
  // incidence on the ice surface: receive angle of the bottom air segment, asin(L/n(surface)) (M.cc:760, 583-589).
  // NB: with RELAY the ice leg re-derives L as n_air(surface) sin(incidence) (M.cc:1913, 565-589), which is Lk again.
  const double n1 = p.seg[p.kb < AIRICE_MAX_LAYERS ? p.kb : 0].stop_n;
  const double n2 = p.seg[AIRICE_ICE_SLOT].start_n;  // n_ice(0)
  const double Lsurf = Lk;
  const double si = AIRICE_DIV(Lsurf, n1);
  // asin(L / n1) = atan(L / sqrt(n1^2 - L^2)) when the ray crossed the air (Rsurf is that square root)
  r.inc_ice_deg = want_inc ? ((kt >= p.kb) ? AIRICE_ATAN_Q(Lsurf, Rsurf) : asin(si)) * m.rad2deg : 0.0;
  // Fresnel field transmission, air->ice at the surface (M.cc:285-301, 321-337) without trig:
  // sin(theta_i) = L/n1, n1 cos(theta_i) = sqrt(n1^2-L^2) = R of the bottom end.
  const double n12 = AIRICE_DIV(n1, n2);
  const double u = n12 * si;
  const double sq = AIRICE_SQRT(1.0 - u * u);
  const double c1 = Rsurf;
  double trs = 1.0 + AIRICE_DIV(c1 - n2 * sq, c1 + n2 * sq);
  const double c2 = AIRICE_DIV(c1, n1);  // cos(theta_i)
  double trp = (1.0 - AIRICE_DIV(n1 * sq - n2 * c2, n1 * sq + n2 * c2)) * n12;
  if (trs != trs) trs = 0.0;
  if (trp != trp) trp = 0.0;
  r.trans_s = trs; r.trans_p = trp;
  r.refr_deg = want_refr ? AIRICE_ATAN_Q(u, sq) * m.rad2deg : 0.0;  // refracted angle just below the surface (P.cc:1081)
}

// ---- a ray in three pieces, so that the multi-antenna table can run the air walk once and the ice leg per antenna
// (the air leg does not depend on the receiver depth, M.cc:887-905)
struct AirIceAirLeg { double x, t, g, L, Rsurf; };   // sums over the air segments, L at the surface, n1 cos(incidence)
struct AirIceIceTop { double L, L2, sA, inv_sA, Dt, Rt, lnTt, Ht; };   // ray-dependent, depth-independent part of the ice leg

template <bool RELAY>
AIRICE_HD AirIceAirLeg airice_ray_air(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx, double L) {
  double xa = 0.0, ta = 0.0, ga = 0.0;
  double Lk = L, Rsurf = 0.0;
  if (kt >= p.kb) {
    if (RELAY) {
      // forward tracer: top-down, L handed from layer to layer
      double sA = 0.0, inv_sA = 0.0, prevR = 0.0, prevH = 0.0, prevT = 1.0, prevLnT = 0.0;
      constexpr int kUnrollF = AIRICE_UNROLL_RELAY;
#pragma unroll kUnrollF
      for (int k = kt; k >= p.kb; k--) {
        const bool top = (k == kt);
        if (!top) Lk = Lk * p.seg[k].relay;
        const double L2 = Lk * Lk;
        sA = AIRICE_SQRT_NZ(1.0 * 1.0 - L2); inv_sA = AIRICE_RCP(sA);   // changes with every relayed L
        const double Cn = p.seg[k].neg_c, iC = p.seg[k].inv_neg_c;
        const double xt = top ? h : p.seg[k].start_x;
        const double nt = top ? n_tx : p.seg[k].start_n;
        const double xb = p.seg[k].stop_x, nb = p.seg[k].stop_n;
        const double Db = nb * nb - L2, Dt = nt * nt - L2;
        const double Rb = AIRICE_SQRT_NZ(Db);
        const double Tb = nb - L2 + sA * Rb;
        const double lnTb = AIRICE_LOG_POS(Tb), Hb = AIRICE_LOG_POS(nb + Rb);
        double Rt, lnTt, Ht;
        if (!top) {
          // Snell hand-over at an interior boundary: L' = L rho with rho = n'/n, i.e. the direction L/n is kept, so
          // R' = sqrt(n'^2 - L'^2) = rho R and ln(n' + R') = ln(n + R) + ln(rho) hold exactly; T' = n' - L'^2 + sA' R'
          // differs from the T of the layer above by ~3e-13 relative, so ln T' = ln T + log1p((T' - T)/T) needs the
          // quotient to 3-4 digits only.  Saves one sqrt and two logs per interior boundary (6 of the 20 logs of a cell).
          Rt = p.seg[k].relay * prevR;
          Ht = prevH + p.seg[k].ln_relay;
          const double u = ((nt - L2 + sA * Rt) - prevT) * AIRICE_RCP_APPROX(prevT);
          lnTt = prevLnT + (u - 0.5 * u * u);
        } else {
          Rt = AIRICE_SQRT_NZ(Dt);
          lnTt = AIRICE_LOG_POS(nt - L2 + sA * Rt);
          Ht = AIRICE_LOG_POS(nt + Rt);
        }
        prevR = Rb; prevH = Hb; prevT = Tb; prevLnT = lnTb;
        double xs, ts, gs;
        airice_seg_sums<true>(1.0, inv_sA, (Lk * iC) * inv_sA, m.c * Cn, Cn, iC, xt, xb, Dt, Db, Rt, Rb, lnTt, lnTb, Ht, Hb,
                              xs, ts, gs);
        xa += -xs; ta += -ts; ga += -gs;
        Rsurf = Rb;  // after the last air segment: sqrt(n_air(surface)^2 - L^2) = n1 cos(incidence)
      }
    } else {
      // tail of a solve: one L throughout, so the order of the layers is free; bottom-up keeps the per-layer plan
      // reads warp-uniform (see airice_x_fast)
      const double L2 = Lk * Lk;
      const double sA = AIRICE_SQRT_NZ(1.0 * 1.0 - L2), inv_sA = AIRICE_RCP(sA);
      constexpr int kUnrollF = AIRICE_UNROLL_FULL;
      // The lower end of layer k and the upper end of layer k-1 are 1e-5 m and ~4e-13 in n apart (M.cc:715), so
      // R, ln T and H at the lower end follow from the upper end below it to first order in the host-made
      // differences dn = n_stop[k] - n_start[k-1] and d(n^2): one sqrt and two logs less per interior boundary.
      // (Second order is (d(n^2)/R^2)^2/8 < 1e-13 while R > 1e-3, i.e. unless the ray grazes that very boundary.)
      double pR = 0.0, pY = 0.0, pT = 1.0, pLnT = 0.0, pH = 0.0, pN = 1.0;
#pragma unroll kUnrollF
      for (int k = p.kb; k <= kt; k++) {
        const bool top = (k == kt);
        const double Cn = p.seg[k].neg_c, iC = p.seg[k].inv_neg_c;
        const double xt = top ? h : p.seg[k].start_x;
        const double nt = top ? n_tx : p.seg[k].start_n;
        const double xb = p.seg[k].stop_x, nb = p.seg[k].stop_n;
        const double Db = nb * nb - L2, Dt = nt * nt - L2;
        double Rt, yt;
        AIRICE_SQRT_RSQRT_NZ(Dt, Rt, yt);
        const double Tt = nt - L2 + sA * Rt;
        const double lnTt = AIRICE_LOG_POS(Tt), Ht = AIRICE_LOG_POS(nt + Rt);
        double Rb, lnTb, Hb;
        if (k > p.kb && pR > 1.0e-3) {
          const double dR = p.seg[k].ho_dn2 * (0.5 * pY);
          Rb = pR + dR;
          lnTb = pLnT + (p.seg[k].ho_dn + sA * dR) * AIRICE_RCP_APPROX(pT);
          Hb = pH + (p.seg[k].ho_dn + dR) * AIRICE_RCP_APPROX(pN + pR);
        } else {
          Rb = AIRICE_SQRT_NZ(Db);
          lnTb = AIRICE_LOG_POS(nb - L2 + sA * Rb);
          Hb = AIRICE_LOG_POS(nb + Rb);
        }
        pR = Rt; pY = yt; pT = Tt; pLnT = lnTt; pH = Ht; pN = nt;
        double xs, ts, gs;
        airice_seg_sums<true>(1.0, inv_sA, (Lk * iC) * inv_sA, m.c * Cn, Cn, iC, xt, xb, Dt, Db, Rt, Rb, lnTt, lnTb, Ht, Hb,
                              xs, ts, gs);
        xa += -xs; ta += -ts; ga += -gs;
        if (k == p.kb) Rsurf = Rb;
      }
    }
  }
  AirIceAirLeg o;
  o.x = xa; o.t = ta; o.g = ga; o.L = Lk; o.Rsurf = Rsurf;
  return o;
}

// upper end of the ice leg (z = 0): everything that does not depend on the receiver depth
AIRICE_HD AirIceIceTop airice_ice_top(const AirIceMedium& m, const AirIcePlan& p, double Lk) {
  const int k = AIRICE_ICE_SLOT;
  const double A = m.A_ice, nt = p.seg[k].start_n;
  AirIceIceTop o;
  o.L = Lk; o.L2 = Lk * Lk;
  o.sA = AIRICE_SQRT(A * A - o.L2); o.inv_sA = AIRICE_RCP(o.sA);
  o.Dt = nt * nt - o.L2;
  o.Rt = AIRICE_SQRT(o.Dt);
  o.lnTt = AIRICE_LOG_POS(A * nt - o.L2 + o.sA * o.Rt); o.Ht = AIRICE_LOG_POS(nt + o.Rt);
  return o;
}
// ice leg from the surface to depth xb (n(xb) = nb): GetIcePropagationPar (M.cc:807-869)
AIRICE_HD void airice_ice_leg(const AirIceMedium& m, const AirIcePlan& p, const AirIceIceTop& o, double xb, double nb,
                              double& xi, double& ti, double& gi, double& recv_deg) {
  const int k = AIRICE_ICE_SLOT;
  const double A = m.A_ice;
  const double Cn = p.seg[k].neg_c, iC = p.seg[k].inv_neg_c;
  const double Db = nb * nb - o.L2;
  const double Rb = AIRICE_SQRT(Db);
  const double lnTb = AIRICE_LOG_POS(A * nb - o.L2 + o.sA * Rb), Hb = AIRICE_LOG_POS(nb + Rb);
  airice_seg_sums<false>(A, o.inv_sA, (o.L * iC) * o.inv_sA, m.c * Cn, Cn, iC, p.seg[k].start_x, xb, o.Dt, Db, o.Rt, Rb, o.lnTt,
                         lnTb, o.Ht, Hb, xi, ti, gi);
  // receive angle asin(L / n(depth)) (M.cc:824 / 583-589) as atan(L / sqrt(n^2 - L^2)): the square root is Rb
  recv_deg = AIRICE_ATAN_Q(o.L, Rb) * m.rad2deg;
}

template <bool RELAY>
AIRICE_HD void airice_ray_full(const AirIceMedium& m, const AirIcePlan& p, int kt, double h, double n_tx, double L,
                               bool in_ice, bool want_inc, bool want_refr, AirIceRay& r) {
  const AirIceAirLeg al = airice_ray_air<RELAY>(m, p, kt, h, n_tx, L);
  r.x_air = al.x; r.t_air = al.t; r.p_air = al.g;
  r.x_ice = 0.0; r.t_ice = 0.0; r.p_ice = 0.0;
  r.recv_deg = 0.0;
  if (in_ice) {
    const AirIceIceTop it = airice_ice_top(m, p, al.L);
    airice_ice_leg(m, p, it, p.seg[AIRICE_ICE_SLOT].stop_x, p.seg[AIRICE_ICE_SLOT].stop_n, r.x_ice, r.t_ice, r.p_ice, r.recv_deg);
  }
  airice_ray_surface(m, p, kt, al.L, al.Rsurf, want_inc, want_refr, r);
}
