// TEST: (1) the libm of the machine the tests run on, applied to arrays (numpy's exp/log are its own SIMD code, not
// libm); (2) the HOST build of airice_glibc_math.cuh on the same arrays.  tests/test_glibc_math.py requires (1) == (2)
// bit for bit; tests/test_gpu_math.py requires the device build == (1).
#include <math.h>

#include "airice_glibc_math.cuh"

extern "C" {
void libm_v(int op, long n, const double* a, const double* b, double* out) {
  for (long i = 0; i < n; i++) out[i] = op == 0 ? exp(a[i]) : op == 1 ? log(a[i]) : pow(a[i], b[i]);
}
void glibc_host_v(int op, long n, const double* a, const double* b, double* out) {
  for (long i = 0; i < n; i++)
    out[i] = op == 0 ? airice_glibc_exp(a[i]) : op == 1 ? airice_glibc_log(a[i]) : airice_glibc_pow(a[i], b[i]);
}
// sin(64 deg) with IceRayTracing.hh's pi, evaluated by libm at run time (volatile: no constant folding)
double libm_sin64(void) { volatile double a = 64.0 * (3.14159265359 / 180.0); return sin(a); }
double model_sin64(void) { return 0x1.cc2ebbb5639ecp-1; }
}
