// inice_kernels.cu -- the in-ice direct / reflected / refracted solver kernels.
//
// Compiled with -fmad=false (see build.py): this solver reproduces the reference's ITERATIONS (GSL falsepos / Newton
// stopped at loose tolerances), so its arithmetic has to round like the reference's x86 build does -- every product
// and sum on its own.  A fused n*n - L*L, for instance, turns the exact 0 the reference gets at the bracket end
// L = min(n(z0), n(z1)) into -1e-17, sqrt() of that into NaN, and the refracted-ray search of the pair into a
// different branch count.  Transcendentals still come from the CUDA math library (<= 1-2 ulp from glibc).
#include "airice_inice.cuh"
#include "kernels.cuh"

namespace airice {

namespace {
constexpr int kThreads = 128;

__device__ __forceinline__ AirIceInIce inice_model(const InIceArgs& a) {
  AirIceInIce m;
  m.A = a.A; m.B = a.B; m.C = a.C; m.pi = 3.14159265359; m.c = 299792458.0;  // IceRayTracing.hh:41-43
  return m;
}

// pass 1: direct + reflected ray for every pair; all 29 columns written (refracted ones as absent)
__global__ void __launch_bounds__(kThreads) airice_inice_dr_kernel(const InIceArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n) return;
  const AirIceInIce m = inice_model(a);
  double o[AIRICE_INICE_NCOLS];
  bool needs_ra;
  const int mask = inice_solve_dr(m, a.z0[i], a.x1[i], a.z1[i], o, needs_ra);
  if (a.mask) a.mask[i] = (uint8_t)mask;
#pragma unroll
  for (int k = 0; k < AIRICE_INICE_NCOLS; k++)
    if (a.out[k]) a.out[k][i] = o[k];
  if (needs_ra) {
    // warp-aggregated append
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(active) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(a.ra_count, __popc(active));
    base = __shfl_sync(active, base, leader);
    a.ra_list[base + __popc(active & ((1u << lane) - 1))] = (int32_t)i;
  }
}

// pass 2: the refracted-ray ladder for the listed pairs only
__global__ void __launch_bounds__(kThreads) airice_inice_ra_kernel(const InIceArgs a) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= (int64_t)(*a.ra_count)) return;
  const int64_t i = a.ra_list[j];
  const AirIceInIce m = inice_model(a);
  const int mask_dr = a.mask[i];
  double o[AIRICE_INICE_NCOLS];
  const int mask = inice_solve_ra(m, a.z0[i], a.x1[i], a.z1[i], (mask_dr & 1) == 0, (mask_dr & 2) == 0, a.out[20][i], o);
  a.mask[i] = (uint8_t)(mask_dr | mask);
  const int cols[] = {2, 3, 6, 7, 10, 11, 14, 15, 16, 17, 21, 22, 23, 24, 27, 28};
#pragma unroll
  for (int c = 0; c < 16; c++)
    if (a.out[cols[c]]) a.out[cols[c]][i] = o[cols[c]];
}

}  // namespace

cudaError_t launch_inice(const InIceArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  const int64_t blocks = (a.n + kThreads - 1) / kThreads;
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  if (!a.ra_list || !a.ra_count || !a.mask || !a.out[20]) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(a.ra_count, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  airice_inice_dr_kernel<<<dim3((unsigned)blocks), kThreads, 0, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // pass 2 is launched over the worst-case length; blocks beyond the list exit at once
  airice_inice_ra_kernel<<<dim3((unsigned)blocks), kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace airice
