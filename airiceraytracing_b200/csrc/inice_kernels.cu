// inice_kernels.cu -- the in-ice direct / reflected / refracted solver kernels.
//
// Compiled with -fmad=false (see build.py): this solver reproduces the reference's ITERATIONS (GSL falsepos / Newton
// stopped at loose tolerances), so its arithmetic has to round like the reference's x86 build does -- every product
// and sum on its own.  A fused n*n - L*L, for instance, turns the exact 0 the reference gets at the bracket end
// L = min(n(z0), n(z1)) into -1e-17, sqrt() of that into NaN, and the refracted-ray search of the pair into a
// different branch count.  Transcendentals still come from the CUDA math library (<= 1-2 ulp from glibc).
#include "airice_inice.cuh"
#include "airice_inice_machine.cuh"
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

// pass 2: the refracted-ray root-search ladder for the listed pairs.  Persistent lanes: each lane steps the state
// machine of ONE pair (airice_inice_machine.cuh) -- the warp's common loop body is a single evaluation of fRaa -- and
// takes the next list entry as soon as its pair is finished, so neither the search a lane is in nor the number of
// evaluations its pair needs (median 22, mean 78, maximum ~700) leaves the other lanes idle.
__global__ void __launch_bounds__(kThreads) airice_inice_ladder_kernel(const InIceArgs a) {
  const AirIceInIce m = inice_model(a);
  const int count = a.ra_count[0];
  InIceRaMachine M;
  InIcePair g;
  int j = 0;
  bool has = false, exhausted = false;
  M.ph = InIceRaMachine::DONE; M.xq = 0;
  for (;;) {
    if (!has && !exhausted) {
      j = atomicAdd(a.ra_count + 1, 1);
      if (j < count) {
        const int64_t i = a.ra_list[j];
        bool flip;
        g = inice_make_pair(m, a.z0[i], a.x1[i], a.z1[i], flip);
        const int mask_dr = a.mask[i];
        M.init(m, g, flip, (mask_dr & 1) == 0, (mask_dr & 2) == 0, a.out[20][i]);
        has = true;
      } else {
        exhausted = true;
      }
    }
    if (__all_sync(0xffffffffu, !has)) break;
    if (has) {
      if (!M.done()) {
        double zm;
        const double y = inice_fraa_eval(g, M.xq, zm);
        M.advance(y, zm);
      }
      if (M.done()) {
        const int64_t n = a.n;
        a.ra_lad[0 * n + j] = M.lv0; a.ra_lad[1 * n + j] = M.lv1;
        a.ra_lad[2 * n + j] = M.cz0; a.ra_lad[3 * n + j] = M.cz1;
        a.ra_lad[4 * n + j] = M.zm0; a.ra_lad[5 * n + j] = M.zm1;
        has = false;
      }
    }
  }
}

// pass 3: times, paths and angles of the refracted rays found by pass 2
__global__ void __launch_bounds__(kThreads) airice_inice_ra_finish_kernel(const InIceArgs a) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= (int64_t)a.ra_count[0]) return;
  const int64_t i = a.ra_list[j], n = a.n;
  const AirIceInIce m = inice_model(a);
  const int mask_dr = a.mask[i];
  bool flip;
  const InIcePair g = inice_make_pair(m, a.z0[i], a.x1[i], a.z1[i], flip);
  InIceRaLadder lad;
  lad.lv[0] = a.ra_lad[0 * n + j]; lad.lv[1] = a.ra_lad[1 * n + j];
  lad.cz[0] = a.ra_lad[2 * n + j]; lad.cz[1] = a.ra_lad[3 * n + j];
  lad.zm[0] = a.ra_lad[4 * n + j]; lad.zm[1] = a.ra_lad[5 * n + j];
  double o[AIRICE_INICE_NCOLS];
  const int mask = inice_ra_finish(m, g, flip, (mask_dr & 1) == 0, (mask_dr & 2) == 0, lad, o);
  a.mask[i] = (uint8_t)(mask_dr | mask);
  const int cols[] = {2, 3, 6, 7, 10, 11, 14, 15, 16, 17, 21, 22, 23, 24, 27, 28};
#pragma unroll
  for (int c = 0; c < 16; c++)
    if (a.out[cols[c]]) a.out[cols[c]][i] = o[cols[c]];
}

int ladder_grid() {
  static int blocks = 0;    // per process; every context of a process sits on the same GPU model
  if (blocks == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, airice_inice_ladder_kernel, kThreads, 0);
    blocks = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
  }
  return blocks;
}

}  // namespace

cudaError_t launch_inice(const InIceArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  const int64_t blocks = (a.n + kThreads - 1) / kThreads;
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  if (!a.ra_list || !a.ra_count || !a.ra_lad || !a.mask || !a.out[20]) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(a.ra_count, 0, 2 * sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  airice_inice_dr_kernel<<<dim3((unsigned)blocks), kThreads, 0, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // the list length stays on the device: pass 2 is a persistent grid (one wave), pass 3 is launched over the worst-case
  // length and its blocks beyond the list exit at once
  int64_t lb = ladder_grid();
  if (lb > blocks) lb = blocks;
  airice_inice_ladder_kernel<<<dim3((unsigned)lb), kThreads, 0, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  airice_inice_ra_finish_kernel<<<dim3((unsigned)blocks), kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace airice
