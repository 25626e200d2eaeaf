// airice_glibc_math.cuh -- exp / log / pow that round like the libm of the reference's x86-64 build.
//
// Why: the in-ice solver (IceRayTracing.cc) returns whatever iterate its GSL false-position / Newton loops stop on
// (|f| < 1e-6 m, |dx| < 1e-6 |x|, +-1e-6 m on the turning depth), and decides "this branch exists" with |f| < 0.5 on a
// function whose turning-depth term amplifies that 1e-6 to ~1e-2 m.  Which iterate the loop stops on can turn on the
// last bit of one exp() -- CUDA's exp/log/pow are 1-2 ulp routines with their own rounding, and round 1 measured one
// solution-branch flip per 20 000 pairs and 1e-9 noise in the refracted roots because of it, while a host build of the
// same headers (glibc's libm) was bit-equal to the reference.  north_star asks for bit-exact branch counts, so the
// device gets the same arithmetic instead: these are the double-precision exp, log and pow of glibc >= 2.28
// (sysdeps/ieee754/dbl-64/e_exp.c, e_log.c, e_pow.c -- the table-driven routines of ARM's Optimized Routines) in the
// operation order of the FMA build that x86-64 glibc dispatches to on every CPU with FMA+AVX2 (__exp_fma, __log_fma,
// __pow_fma of Ubuntu glibc 2.39; the compiler contracted a*b+c there, so the contraction pattern is part of the
// function), written with explicit fused and unfused operations so that neither nvcc nor gcc re-associates them.
// The tables are libm's own (tools/gen_glibc_tables.py).  tests/test_glibc_math.py checks the host build of this header
// against the libm of the machine it runs on, bit for bit, on 3e7 arguments per function; tests/test_gpu_math.py does the
// same for the device build.
//
// Coverage: the fast paths, which is every argument the solver produces (|x| < 512 for exp; positive normal x for log and
// pow, |y log x| < 512) plus the cheap special cases (tiny |x|, log of 0 / negative / inf / NaN / subnormal).  The rare
// remainder (results that overflow or are subnormal) falls back to the platform's function.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define AIRICE_G_FMA(a, b, c) __fma_rn((a), (b), (c))
#define AIRICE_G_MUL(a, b) __dmul_rn((a), (b))
#define AIRICE_G_ADD(a, b) __dadd_rn((a), (b))
#define AIRICE_G_SUB(a, b) __dsub_rn((a), (b))
#define AIRICE_G_BITS(x) ((unsigned long long)__double_as_longlong(x))
#define AIRICE_G_DBL(u) __longlong_as_double((long long)(u))
#define AIRICE_G_LD(p) __ldg(p)
#else
// host build (tests): compiled with -ffp-contract=off, so plain operators stay unfused
#define AIRICE_G_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define AIRICE_G_MUL(a, b) ((a) * (b))
#define AIRICE_G_ADD(a, b) ((a) + (b))
#define AIRICE_G_SUB(a, b) ((a) - (b))
static inline unsigned long long airice_g_bits(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
static inline double airice_g_dbl(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
#define AIRICE_G_BITS(x) airice_g_bits(x)
#define AIRICE_G_DBL(u) airice_g_dbl(u)
#define AIRICE_G_LD(p) (*(p))
#endif

#if defined(__CUDACC__)
#define AIRICE_G_FN __host__ __device__ __forceinline__
#else
#define AIRICE_G_FN static inline
#endif

// The tables: __device__ memory for kernels (8 KB, L1/L2 resident, read with ld.global.nc), plain statics for host code
#if defined(__CUDACC__)
#define AIRICE_GLIBC_TABLE(type, name, n) static __device__ const type name##_d[n]
#include "airice_glibc_tables.inc"
#undef AIRICE_GLIBC_TABLE
#endif
#define AIRICE_GLIBC_TABLE(type, name, n) static const type name##_h[n] __attribute__((unused))
#include "airice_glibc_tables.inc"
#undef AIRICE_GLIBC_TABLE
#if defined(__CUDA_ARCH__)
#define AIRICE_G_TAB(name) name##_d
#else
#define AIRICE_G_TAB(name) name##_h
#endif

// 64-bit literals cost two moves each in SASS (every coefficient of a polynomial became a UMOV pair: 14 moves next to the
// 12 FP64 operations of one exp); from __constant__ memory they are constant-bank operands of the DFMA itself.  The rare
// out-of-range fallbacks are real calls, so that CUDA's exp / pow are not inlined (55 dead instructions) into every use.
#if defined(__CUDACC__)
static __constant__ double airice_gk[24] = {
    0x1.71547652b82fep+7, -0x1.62e42fefa0000p-8, -0x1.cf79abc9e3b3ap-47,                              // 0-2 exp reduction
    0x1.ffffffffffdbdp-2, 0x1.555555555543cp-3, 0x1.55555cf172b91p-5, 0x1.1111167a4d017p-7,           // 3-6 exp C2..C5
    0x1.62e42fefa3800p-1, 0x1.ef35793c76730p-45,                                                      // 7-8 ln2 hi, lo
    0x1.555555551305bp-2, -0x1.fffffffeb4590p-3, 0x1.999b324f10111p-3, -0x1.55575e506c89fp-3,         // 9-12 log A1..A4
    -0x1.0000000000001p-1,                                                                            // 13 log A0
    -0x1.5555555555560p-1, 0x1.0000000000006p-1, 0x1.999999959554ep-1, -0x1.555555529a47ap-1,         // 14-17 pow A1..A4
    -0x1.2495b9b4845e9p+0, 0x1.0002b8b263fc3p+0, 0, 0, 0, 0};                                         // 18-19 pow A5, A6
static __host__ __device__ __noinline__ double airice_glibc_exp_cold(double x) { return exp(x); }
static __host__ __device__ __noinline__ double airice_glibc_pow_cold(double x, double y) { return pow(x, y); }
#else
static inline double airice_glibc_exp_cold(double x) { return exp(x); }
static inline double airice_glibc_pow_cold(double x, double y) { return pow(x, y); }
#endif
#if defined(__CUDA_ARCH__)
#define AIRICE_GK(i, lit) airice_gk[i]
#else
#define AIRICE_GK(i, lit) (lit)
#endif

// ---- exp (e_exp.c: exp(x) = 2^(k/128) exp(r), |r| <= ln2/256, degree-5 polynomial, 2^(i/128) = scale (1 + tail))
// The polynomial tail shared by exp and pow: r is the reduced argument, ki the bits of k + Shift.
AIRICE_G_FN double airice_glibc_exp_tail(double r, unsigned long long ki) {
  const int idx = 2 * (int)(ki & 127);
  const double tail = AIRICE_G_DBL(AIRICE_G_LD(&AIRICE_G_TAB(airice_glibc_exp_tab)[idx]));
  const unsigned long long sbits = AIRICE_G_LD(&AIRICE_G_TAB(airice_glibc_exp_tab)[idx + 1]) + (ki << 45);
  const double p23 = AIRICE_G_FMA(AIRICE_GK(4, 0x1.555555555543cp-3), r, AIRICE_GK(3, 0x1.ffffffffffdbdp-2));   // C2 + r C3
  const double tr = AIRICE_G_ADD(r, tail);
  const double r2 = AIRICE_G_MUL(r, r);
  const double p45 = AIRICE_G_FMA(r, AIRICE_GK(6, 0x1.1111167a4d017p-7), AIRICE_GK(5, 0x1.55555cf172b91p-5));   // C4 + r C5
  const double t = AIRICE_G_FMA(p23, r2, tr);
  const double r4 = AIRICE_G_MUL(r2, r2);
  const double tmp = AIRICE_G_FMA(p45, r4, t);
  const double scale = AIRICE_G_DBL(sbits);
  return AIRICE_G_FMA(scale, tmp, scale);
}

AIRICE_G_FN double airice_glibc_exp(double x) {
  const unsigned abstop = (unsigned)(AIRICE_G_BITS(x) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop < 0x3c9u) return AIRICE_G_ADD(1.0, x);      // |x| < 2^-54
    return airice_glibc_exp_cold(x);                        // |x| >= 512, inf, NaN: results no ray produces
  }
  const double kd0 = AIRICE_G_FMA(x, AIRICE_GK(0, 0x1.71547652b82fep+7), 0x1.8p+52);      // x N/ln2 + Shift, fused
  const unsigned long long ki = AIRICE_G_BITS(kd0);
  const double kd = AIRICE_G_SUB(kd0, 0x1.8p+52);
  const double r = AIRICE_G_FMA(kd, AIRICE_GK(2, -0x1.cf79abc9e3b3ap-47), AIRICE_G_FMA(kd, AIRICE_GK(1, -0x1.62e42fefa0000p-8), x));
  return airice_glibc_exp_tail(r, ki);
}

// ---- log (e_log.c: x = 2^k z, z/c_i - 1 = r exactly with FMA, k ln2 + log c_i + r as hi + lo, degree-6 polynomial; a
// degree-12 polynomial in x - 1 with a split square term within [1 - 2^-4, 1 + 0x1.09p-4))
AIRICE_G_FN double airice_glibc_log(double x) {
  unsigned long long ix = AIRICE_G_BITS(x);
  if (ix - 0x3fee000000000000ull <= 0x308ffffffffffull) {
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = AIRICE_G_SUB(x, 1.0);
    const double r2 = AIRICE_G_MUL(r, r);
    const double r3 = AIRICE_G_MUL(r, r2);
    // y = r3 (B1 + r B2 + r2 B3 + r3 (B4 + r B5 + r2 B6 + r3 (B7 + r B8 + r2 B9 + r3 B10)))
    const double q1 = AIRICE_G_FMA(r2, 0x1.999999995dd0cp-3, AIRICE_G_FMA(r, -0x1.ffffffffffdcbp-3, 0x1.5555555555577p-2));
    const double q4 = AIRICE_G_FMA(r2, -0x1.fffffa4423d65p-4, AIRICE_G_FMA(r, 0x1.24924a344de30p-3, -0x1.55555556745a7p-3));
    double q7 = AIRICE_G_FMA(r, -0x1.999eb43b068ffp-4, 0x1.c7184282ad6cap-4);
    q7 = AIRICE_G_FMA(r2, 0x1.78182f7afd085p-4, q7);
    q7 = AIRICE_G_FMA(r3, -0x1.5521375d145cdp-4, q7);
    const double p = AIRICE_G_FMA(AIRICE_G_FMA(q7, r3, q4), r3, q1);
    const double rw = AIRICE_G_FMA(r, 0x1p27, r);                 // r + r 2^27, fused
    const double rhi = AIRICE_G_FMA(-0x1p27, r, rw);
    const double rhi2 = AIRICE_G_MUL(rhi, rhi);
    const double rlo = AIRICE_G_SUB(r, rhi);
    const double hi = AIRICE_G_FMA(rhi2, -0.5, r);
    double lo = AIRICE_G_FMA(rhi2, -0.5, AIRICE_G_SUB(r, hi));
    lo = AIRICE_G_FMA(AIRICE_G_MUL(-0.5, rlo), AIRICE_G_ADD(r, rhi), lo);
    const double y = AIRICE_G_FMA(p, r3, lo);
    return AIRICE_G_ADD(hi, y);
  }
  const unsigned top = (unsigned)(ix >> 48);
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
    if (ix * 2 == 0) return -INFINITY;                             // log(+-0)
    if (ix == 0x7ff0000000000000ull) return x;                     // log(inf)
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return (x != x) ? x : NAN;   // negative, NaN
    ix = AIRICE_G_BITS(AIRICE_G_MUL(x, 0x1p52)) - (52ull << 52);   // subnormal: normalise
  }
  const unsigned long long tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)(tmp >> 45) & 127;
  const int k = (int)((long long)tmp >> 52);
  const double z = AIRICE_G_DBL(ix - (tmp & 0xfff0000000000000ull));
  const double invc = AIRICE_G_LD(&AIRICE_G_TAB(airice_glibc_log_tab)[2 * i]);
  const double logc = AIRICE_G_LD(&AIRICE_G_TAB(airice_glibc_log_tab)[2 * i + 1]);
  const double kd = (double)k;
  const double w = AIRICE_G_FMA(kd, AIRICE_GK(7, 0x1.62e42fefa3800p-1), logc);
  const double r = AIRICE_G_FMA(z, invc, -1.0);
  const double p12 = AIRICE_G_FMA(r, AIRICE_GK(10, -0x1.fffffffeb4590p-3), AIRICE_GK(9, 0x1.555555551305bp-2));    // A1 + r A2
  const double hi = AIRICE_G_ADD(r, w);
  const double r2 = AIRICE_G_MUL(r, r);
  double lo = AIRICE_G_ADD(AIRICE_G_SUB(w, hi), r);
  lo = AIRICE_G_FMA(kd, AIRICE_GK(8, 0x1.ef35793c76730p-45), lo);
  const double r3 = AIRICE_G_MUL(r, r2);
  const double p34 = AIRICE_G_FMA(r, AIRICE_GK(12, -0x1.55575e506c89fp-3), AIRICE_GK(11, 0x1.999b324f10111p-3));    // A3 + r A4
  lo = AIRICE_G_FMA(r2, AIRICE_GK(13, -0x1.0000000000001p-1), lo);                                   // + r2 A0
  const double p = AIRICE_G_FMA(p34, r2, p12);
  return AIRICE_G_ADD(AIRICE_G_FMA(r3, p, lo), hi);
}

// ---- pow (e_pow.c: log x as hi + lo with ~68 bits, y log x as ehi + elo, exp of that), x > 0 normal
AIRICE_G_FN double airice_glibc_pow(double x, double y) {
  const unsigned long long ix = AIRICE_G_BITS(x), iy = AIRICE_G_BITS(y);
  const unsigned topx = (unsigned)(ix >> 52), topy = (unsigned)(iy >> 52) & 0x7ff;
  if (topx - 1u >= 0x7fdu || topy - 0x3beu >= 0x80u) return airice_glibc_pow_cold(x, y);   // x not a positive normal, or |y| tiny/huge
  const unsigned long long tmp = ix - 0x3fe6955500000000ull;
  const int i = (int)(tmp >> 45) & 127;
  const int k = (int)((long long)tmp >> 52);
  const double z = AIRICE_G_DBL(ix - (tmp & 0xfff0000000000000ull));
  const double kd = (double)k;
  const double invc = AIRICE_G_LD(&AIRICE_G_TAB(airice_glibc_powlog_tab)[3 * i]);
  const double logc = AIRICE_G_LD(&AIRICE_G_TAB(airice_glibc_powlog_tab)[3 * i + 1]);
  const double logctail = AIRICE_G_LD(&AIRICE_G_TAB(airice_glibc_powlog_tab)[3 * i + 2]);
  const double t1 = AIRICE_G_FMA(kd, AIRICE_GK(7, 0x1.62e42fefa3800p-1), logc);
  const double lo1 = AIRICE_G_FMA(kd, AIRICE_GK(8, 0x1.ef35793c76730p-45), logctail);
  const double r = AIRICE_G_FMA(z, invc, -1.0);
  const double ar = AIRICE_G_MUL(r, -0.5);
  const double p12 = AIRICE_G_FMA(r, AIRICE_GK(15, 0x1.0000000000006p-1), AIRICE_GK(14, -0x1.5555555555560p-1));    // A1 + r A2
  const double p34 = AIRICE_G_FMA(r, AIRICE_GK(17, -0x1.555555529a47ap-1), AIRICE_GK(16, 0x1.999999959554ep-1));    // A3 + r A4
  const double t2 = AIRICE_G_ADD(r, t1);
  const double lo2 = AIRICE_G_ADD(AIRICE_G_SUB(t1, t2), r);
  const double ar2 = AIRICE_G_MUL(r, ar);
  const double ar3 = AIRICE_G_MUL(r, ar2);
  const double lo3 = AIRICE_G_FMA(ar, r, -ar2);
  const double hi = AIRICE_G_ADD(t2, ar2);
  const double p56 = AIRICE_G_FMA(r, AIRICE_GK(19, 0x1.0002b8b263fc3p+0), AIRICE_GK(18, -0x1.2495b9b4845e9p+0));    // A5 + r A6
  const double lo4 = AIRICE_G_ADD(AIRICE_G_SUB(t2, hi), ar2);
  const double p36 = AIRICE_G_FMA(p56, ar2, p34);
  const double p16 = AIRICE_G_FMA(ar2, p36, p12);
  double lo = AIRICE_G_ADD(lo1, lo2);
  lo = AIRICE_G_ADD(lo, lo3);
  lo = AIRICE_G_ADD(lo, lo4);
  lo = AIRICE_G_FMA(ar3, p16, lo);
  const double lhi = AIRICE_G_ADD(hi, lo);
  const double llo = AIRICE_G_ADD(AIRICE_G_SUB(hi, lhi), lo);
  const double ehi = AIRICE_G_MUL(y, lhi);
  const double elo = AIRICE_G_FMA(y, llo, AIRICE_G_FMA(lhi, y, -ehi));
  const unsigned abstop = (unsigned)(AIRICE_G_BITS(ehi) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop < 0x3c9u) return AIRICE_G_ADD(1.0, ehi);    // |y log x| < 2^-54 (sign bias 0: x > 0)
    return airice_glibc_pow_cold(x, y);                     // overflow / underflow range
  }
  const double kd0 = AIRICE_G_FMA(ehi, AIRICE_GK(0, 0x1.71547652b82fep+7), 0x1.8p+52);
  const unsigned long long ki = AIRICE_G_BITS(kd0);
  const double kde = AIRICE_G_SUB(kd0, 0x1.8p+52);
  double re = AIRICE_G_FMA(kde, AIRICE_GK(1, -0x1.62e42fefa0000p-8), ehi);
  re = AIRICE_G_FMA(kde, AIRICE_GK(2, -0x1.cf79abc9e3b3ap-47), re);
  re = AIRICE_G_ADD(elo, re);
  return airice_glibc_exp_tail(re, ki);
}
