// oldtable_kernels.cu -- the reference's old solve-per-cell table (MakeTable, MultiRayAirIceRefraction.cc:1618-1696) and its
// inverse-distance lookup (GetInterpolatedValue, MultiRayAirIceRefraction.cc:1700-1794) on the device.
//
// The grid is kernel 2 driven by a grid generator: one launch-angle solve per (Tx height, straight-line angle) node with
// the node's angle as the StraightAngle argument (M.cc:1665), 34.9 M solves at the reference's defaults.  These kernels
// only generate the nodes, turn the solver's columns into the reference's nine GridZValue columns (-1000 where the solve
// misses, M.cc:1667-1688), and answer batches of GetInterpolatedValue queries against the device-resident columns.
// Compiled with -fmad=false: the lookup is plain arithmetic on the stored doubles and rounds like the reference's build.
#include "kernels.cuh"

namespace airice {

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) airice_oldgrid_cells_kernel(const OldGridArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n) return;
  const int64_t cell = a.cell0 + i;
  const int ih = (int)(cell / a.n_th), ith = (int)(cell - (int64_t)ih * a.n_th);
  const double h = a.pos_h[ih];
  a.h[i] = h;
  a.th[i] = a.pos_th[ith];
  a.d[i] = (h - a.ice + a.depth) * a.col_tan[ith];          // M.cc:1662, tan() of the node's angle host-made (libm)
}

// solver columns (M_DEG layout of kernel 2) -> GridZValue[0..8] (M.cc:1667-1688)
__global__ void __launch_bounds__(kThreads) airice_oldgrid_pack_kernel(const OldPackArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n) return;
  const double d = a.d[i], X = a.x[i];
  // the reference accepts on the distance test alone here, without the X < 0 veto of the CoREAS entry point
  const bool accept = (fabs(X - d) / d < 0.01 && d <= 100) || (fabs(X - d) < 1 && d > 100);
  double v[9];
  if (accept) {
    v[0] = a.h[i]; v[1] = X; v[2] = a.t_ice[i] * a.c; v[3] = a.t_air[i] * a.c; v[4] = a.launch[i]; v[5] = a.x_air[i];
    v[6] = a.ts[i]; v[7] = a.tp[i]; v[8] = a.inc[i];
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) v[k] = -1000;
  }
#pragma unroll
  for (int k = 0; k < 9; k++) a.col[k][i] = v[k];
}

// GetInterpolatedValue, reproduced as written: the 2x2 nodes below/left of the rounded (and clamped) bin, running value
// overwritten per node, an exact hit short-circuits (M.cc:1700-1794)
__global__ void __launch_bounds__(kThreads) airice_old_interp_kernel(const OldInterpArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n) return;
  const double hR = a.h[i], thR = a.th[i];
  double sum1 = 0, sum2 = 0, NewZValue = -1000;
  double minHbin = round((hR - a.start_h) / a.step_h);
  double minThbin = round((thR - a.start_th) / a.step_th);
  if (minHbin <= 1) minHbin = 1;
  if (minThbin <= 1) minThbin = 1;
  if (minHbin + 1 > a.n_h) minHbin = a.n_h - 2;
  if (minThbin + 1 > a.n_th) minThbin = a.n_th - 2;
  const int startbinH = (int)(minHbin - 1), endbinH = (int)(minHbin + 1);
  const int startbinTh = (int)(minThbin - 1), endbinTh = (int)(minThbin + 1);
  const int64_t points = (int64_t)a.n_h * a.n_th;
  bool done = false;
  for (int ixn = startbinH; ixn < endbinH && !done; ixn++) {
    for (int izn = startbinTh; izn < endbinTh && !done; izn++) {
      const int64_t ich = (int64_t)ixn * a.n_th + izn;
      if (ich >= 0 && ich < points && ixn < a.n_h && izn < a.n_th && ixn >= 0 && izn >= 0) {
        const double dh = hR - a.pos_h[ixn], dt = thR - a.pos_th[izn];
        const double dist = fabs(dh * dh + dt * dt);
        const double z = a.z[ich];
        if (z != -1000) {
          sum1 += (1.0 / dist) * z;
          sum2 += (1.0 / dist);
          NewZValue = sum1 / sum2;
        } else {
          NewZValue = -1000;
        }
        if (dist == 0) {
          NewZValue = (z != -1000) ? z : -1000;
          done = true;               // the reference jumps both loop counters past their ends
        }
      }
    }
  }
  a.out[i] = NewZValue;
}
}  // namespace

cudaError_t launch_oldgrid_cells(const OldGridArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  airice_oldgrid_cells_kernel<<<dim3((unsigned)((a.n + kThreads - 1) / kThreads)), kThreads, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_oldgrid_pack(const OldPackArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  airice_oldgrid_pack_kernel<<<dim3((unsigned)((a.n + kThreads - 1) / kThreads)), kThreads, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_old_interp(const OldInterpArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  airice_old_interp_kernel<<<dim3((unsigned)((a.n + kThreads - 1) / kThreads)), kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace airice
