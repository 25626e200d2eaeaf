// TEST: applies the product's device math (airice_math.cuh: log, sqrt, rcp, div, atan, /100) to arrays, so that
// tests/test_gpu_math.py can measure their error against high-precision references.  Built at test time with nvcc.
#include <cuda_runtime.h>
#include <math.h>

#include "airice_math.cuh"
#include "airice_glibc_math.cuh"

__global__ void probe_kernel(int op, long n, const double* a, const double* b, double* out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r;
  switch (op) {
    case 0: r = AIRICE_LOG(a[i]); break;
    case 1: r = AIRICE_SQRT(a[i]); break;
    case 2: r = AIRICE_RCP(a[i]); break;
    case 3: r = AIRICE_DIV(a[i], b[i]); break;
    case 4: r = AIRICE_ATAN_Q(a[i], b[i]); break;
    case 5: r = AIRICE_DIV100(a[i]); break;
    case 6: r = airice_glibc_exp(a[i]); break;
    case 7: r = airice_glibc_log(a[i]); break;
    default: r = airice_glibc_pow(a[i], b[i]); break;
  }
  out[i] = r;
}

extern "C" int math_probe(int op, long n, const double* a, const double* b, double* out) {
  double *da = nullptr, *db = nullptr, *dout = nullptr;
  if (cudaMalloc(&da, n * 8) || cudaMalloc(&db, n * 8) || cudaMalloc(&dout, n * 8)) return 1;
  cudaMemcpy(da, a, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b, n * 8, cudaMemcpyHostToDevice);
  probe_kernel<<<(unsigned)((n + 255) / 256), 256>>>(op, n, da, db, dout);
  const int rc = cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess;
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return rc;
}
