// inice_kernels.cu -- the in-ice direct / reflected / refracted solver kernels.
//
// Compiled with -fmad=false (see build.py): this solver reproduces the reference's ITERATIONS (GSL falsepos / Newton
// stopped at loose tolerances), so its arithmetic has to round like the reference's x86 build does -- every product
// and sum on its own.  A fused n*n - L*L, for instance, turns the exact 0 the reference gets at the bracket end
// L = min(n(z0), n(z1)) into -1e-17, sqrt() of that into NaN, and the refracted-ray search of the pair into a
// different branch count.  exp / log / pow are glibc's own algorithms (airice_glibc_math.cuh) for the same reason.
#include "airice_inice.cuh"
#include "airice_inice_machine.cuh"
#include "airice_inice_att.cuh"
#include "kernels.cuh"

namespace airice {

namespace {
constexpr int kThreads = 128;
// CTA sizes of pass 1 and pass 2 (instruction fetch is a top stall of both; see the note at airice_solve_kernel)
#ifndef AIRICE_INICE_DR_THREADS
#define AIRICE_INICE_DR_THREADS 512
#endif
#ifndef AIRICE_INICE_LADDER_THREADS
#define AIRICE_INICE_LADDER_THREADS 128
#endif
// The search state of a lane's pair (InIceRaMachine, 200 B) lives in shared memory: in registers it made the kernel a
// 144-register one (3 CTAs of 128 threads per SM, 2.8 warps per scheduler to hide FP64 chains with); in shared memory the
// kernel fits 4 CTAs at 114 registers without spilling: 26.8 -> 23.0 ms per 2e6 pairs (5 / 6 CTAs spill: 23.3 / 23.7 ms).
#ifndef AIRICE_INICE_LADDER_MINBLOCKS
#define AIRICE_INICE_LADDER_MINBLOCKS 4
#endif
#ifndef AIRICE_INICE_SMEM_MACHINE
#define AIRICE_INICE_SMEM_MACHINE 1
#endif
constexpr int kDrThreads = AIRICE_INICE_DR_THREADS;
constexpr int kLadderThreads = AIRICE_INICE_LADDER_THREADS;

__device__ __forceinline__ AirIceInIce inice_model(const InIceArgs& a) {
  return inice_make_model(a.A, a.B, a.C);
}

// pass 1: direct + reflected ray for every pair; all 29 columns written (refracted ones as absent)
#ifndef AIRICE_INICE_DR_MINBLOCKS
#define AIRICE_INICE_DR_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(kDrThreads, AIRICE_INICE_DR_MINBLOCKS) airice_inice_dr_kernel(const InIceArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kDrThreads + threadIdx.x;
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
    // warp-aggregated append.  Pairs that lack BOTH the direct and the reflected ray search for two refracted roots
    // (mean 150 evaluations against 84) and go to the FRONT of the list, the others to the back: pass 2 hands the list
    // out front to back, so the long searches start first and the short ones fill the tail.
    const bool front = (mask == 0);
    const int lane = threadIdx.x & 31;
    const unsigned active = __activemask();
    const unsigned grp = __ballot_sync(active, front);
    const unsigned mine = front ? grp : (active & ~grp);
    const int leader = __ffs(mine) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(a.ra_count + (front ? 0 : 2), __popc(mine));
    base = __shfl_sync(mine, base, leader);
    const int pos = base + __popc(mine & ((1u << lane) - 1));
    a.ra_list[front ? pos : (int)(a.n - 1 - pos)] = (int32_t)i;
  }
}

// list position j (0 <= j < front + back) -> slot of ra_list
__device__ __forceinline__ int64_t ra_slot(const InIceArgs& a, int j, int n_front) {
  return j < n_front ? (int64_t)j : a.n - 1 - (int64_t)(j - n_front);
}

// pass 2: the refracted-ray root-search ladder for the listed pairs.  Persistent lanes: each lane owns the search state
// of ONE pair (airice_inice_machine.cuh) and takes the next list entry as soon as its pair is finished.  Every step, the
// lanes of a warp pool the evaluations of fRaa their pairs ask for (1, 2, 4 or 5 each) in shared memory and share them
// out evenly, so the loop body all lanes run together is one evaluation of fRaa, whatever search each pair is in, and
// a pair with a long search (the number of evaluations per pair spans 3 ... several thousand) uses idle lanes.
struct InIceWarpPool {
  static constexpr int kCap = 32;
  double x[kCap], y[kCap], zm[kCap];   // request points and results
  double pair[5][32];                  // z0, z1, x1, n(z0), n(z1) of each lane's pair
  unsigned char owner[kCap];
};

__global__ void __launch_bounds__(kLadderThreads, AIRICE_INICE_LADDER_MINBLOCKS) airice_inice_ladder_kernel(const InIceArgs a) {
  __shared__ InIceWarpPool pools[kLadderThreads / 32];
#if AIRICE_INICE_SMEM_MACHINE
  __shared__ InIceRaMachine machines[kLadderThreads];
  InIceRaMachine& M = machines[threadIdx.x];
#else
  InIceRaMachine M;
#endif
  InIceWarpPool& pool = pools[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const AirIceInIce m = inice_model(a);
  const int n_front = a.ra_count[0], count = n_front + a.ra_count[2];
  const double e5000 = INICE_EXP(-a.C * 5000.0);
  int j = 0;
  bool has = false, exhausted = false;
  M.ph = InIceRaMachine::DONE; M.nq = 0; M.xq = 0;
  double yk[InIceRaMachine::kMaxReq] = {0, 0, 0, 0, 0}, zk[InIceRaMachine::kMaxReq] = {0, 0, 0, 0, 0};
  int slot[InIceRaMachine::kMaxReq] = {0, 0, 0, 0, 0};
  int rot = 0;
  double my_x1 = 0.0;
  unsigned n_eval = 0, n_zstep = 0;      // work done by this lane: fRaa evaluations and turning-depth falsepos steps
  for (;;) {
    if (!has && !exhausted) {
      j = atomicAdd(a.ra_count + 1, 1);
      if (j < count) {
        const int64_t i = a.ra_list[ra_slot(a, j, n_front)];
        bool flip;
        const InIcePair g = inice_make_pair(m, a.z0[i], a.x1[i], a.z1[i], flip);
        const int mask_dr = a.mask[i];
        M.init(m, g, flip, (mask_dr & 1) == 0, (mask_dr & 2) == 0, a.out[20][i]);
        pool.pair[0][lane] = g.z0; pool.pair[1][lane] = g.z1; pool.pair[2][lane] = g.x1;
        pool.pair[3][lane] = g.n0; pool.pair[4][lane] = g.n1;
        my_x1 = g.x1;
        has = true;
      } else {
        exhausted = true;
      }
    }
    // Requests outside the physical range of L are settled by their owner: for L = NaN (a Newton search that left the
    // domain keeps asking for it: 9% of all requests) and for L > A (5%) every fL term is NaN, fRaa is its
    // "1e9 - 2e9 - x1" penalty, and the turning-depth search ends after one step at a value known in closed form.
    // The others ("hard": 9-16 falsepos steps for the turning depth plus three fL) go to the warp's pool.
    const int nq = has ? M.nq : 0;
    double xk[InIceRaMachine::kMaxReq];
    int nhard = 0;
    unsigned hard_bits = 0;
#pragma unroll
    for (int k = 0; k < InIceRaMachine::kMaxReq; k++) {
      xk[k] = 0.0;
      if (k < nq) {
        const double L = xk[k] = M.query(k);
        const bool hard = !inice_fraa_shortcut(a.A, a.B, e5000, my_x1, L, yk[k], zk[k]);
        if (hard) { hard_bits |= 1u << k; nhard++; }
      }
    }
    // The pool takes at most 32 hard requests per trip -- one for every lane, so that the evaluation below is a single
    // fully occupied pass -- packed greedily in lane order from a start lane that rotates; a pair whose requests do
    // not fit any more simply waits for the next trip.
    if (!__any_sync(full, has)) {        // no lane holds a pair and the list is used up
      // work census of the launch (bench.py turns it into the kernel's algorithmic-work figure): one atomic per warp
      const unsigned ev = __reduce_add_sync(full, n_eval), zs = __reduce_add_sync(full, n_zstep);
      if (lane == 0) {
        unsigned long long* w = (unsigned long long*)(a.ra_count + 4);
        atomicAdd(w, (unsigned long long)ev);
        atomicAdd(w + 1, (unsigned long long)zs);
      }
      break;
    }
    rot = (rot + 11) & 31;
    int incl = __shfl_sync(full, nhard, (lane + rot) & 31);      // lane p holds the count of virtual position p
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(full, incl, d);
      if (lane >= d) incl += t;
    }
    incl = __shfl_sync(full, incl, (lane - rot) & 31);           // back to the owner: inclusive sum up to its position
    const bool accepted = incl <= 32 || nhard == 0;
    const int total = __reduce_max_sync(full, incl <= 32 ? incl : 0);
    if (accepted) {
      int hb = incl - nhard;
#pragma unroll
      for (int k = 0; k < InIceRaMachine::kMaxReq; k++)
        if (hard_bits & (1u << k)) { pool.x[hb] = xk[k]; pool.owner[hb] = (unsigned char)lane; slot[k] = hb++; }
    }
    __syncwarp();
    if (lane < total) {
      const int o = pool.owner[lane];
      InIcePair g;
      g.A = a.A; g.B = a.B; g.C = a.C;
      g.z0 = pool.pair[0][o]; g.z1 = pool.pair[1][o]; g.x1 = pool.pair[2][o]; g.n0 = pool.pair[3][o]; g.n1 = pool.pair[4][o];
      g.ns = 0;
      const double x = pool.x[lane];
      InIceZmaxIter Z;
      Z.init(a.A, a.B, e5000, x);
      n_eval++;
#pragma unroll 1
      do { n_zstep++; } while (!Z.step(a.A, a.B, a.C));
      const double zm = Z.root + 1e-7;
      pool.zm[lane] = zm;
      pool.y[lane] = inice_fraa_given_zmax(g, x, zm);
    }
    __syncwarp();
    if (accepted) {
#pragma unroll
      for (int k = 0; k < InIceRaMachine::kMaxReq; k++)
        if (hard_bits & (1u << k)) { yk[k] = pool.y[slot[k]]; zk[k] = pool.zm[slot[k]]; }
    }
    if (has && accepted) {
      M.advance(yk, zk);
      if (M.done()) {
        const int64_t n = a.n;
        a.ra_lad[0 * n + j] = M.lv0; a.ra_lad[1 * n + j] = M.lv1;
        a.ra_lad[2 * n + j] = M.cz0; a.ra_lad[3 * n + j] = M.cz1;
        a.ra_lad[4 * n + j] = M.zm0; a.ra_lad[5 * n + j] = M.zm1;
        has = false;
      }
    }
    __syncwarp();
  }
}

// pass 3: times, paths and angles of the refracted rays found by pass 2
__global__ void __launch_bounds__(kThreads) airice_inice_ra_finish_kernel(const InIceArgs a) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int n_front = a.ra_count[0];
  if (j >= (int64_t)n_front + (int64_t)a.ra_count[2]) return;
  const int64_t i = a.ra_list[ra_slot(a, (int)j, n_front)], n = a.n;
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

// the two physical rays of each pair out of the four candidates (elementwise; HBM bound: 29 + 3 columns in, 10 + 4 out)
template <bool ATT>     // ATT: with the attenuation integrals (a QAGS run per candidate ray: its own instantiation keeps the plain
                        // selection at its register count)
__global__ void __launch_bounds__(256) airice_inice_pick_kernel(const InIcePickArgs a) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  const AirIceInIce m = inice_make_model(a.A, a.B, a.C);
  double o[AIRICE_INICE_NCOLS];
#pragma unroll
  for (int k = 0; k < AIRICE_INICE_NCOLS; k++) o[k] = a.in[k][i];
  double res[AIRICE_INICE_RAYS_NCOLS];
  int ig[2], ty[2];
  if (ATT) {
    const InIceAttModel am = {a.A0, a.frequency, a.w0, a.w2, a.w};
    double att4[4], att2[2];
    int flags = 0, worst = 0;
    inice_candidate_attenuations(m, am, o, a.rx_depth[i], a.tx_depth[i], att4, flags, worst);
    inice_pick_two_rays(m, o, a.rx_depth[i], a.distance[i], a.tx_depth[i], res, ig, ty, att4, att2);
    a.att[0][i] = att2[0]; a.att[1][i] = att2[1];
    if (a.quad_stats) {
      if (flags) atomicAdd(a.quad_stats, 1);
      if (worst > 24) atomicMax(a.quad_stats + 1, worst);
    }
  } else {
    inice_pick_two_rays(m, o, a.rx_depth[i], a.distance[i], a.tx_depth[i], res, ig, ty);
  }
#pragma unroll
  for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++)
    if (a.out[k]) a.out[k][i] = res[k];
  a.ignore[0][i] = ig[0]; a.ignore[1][i] = ig[1];
  if (a.type[0]) a.type[0][i] = ty[0];
  if (a.type[1]) a.type[1][i] = ty[1];
}

__global__ void __launch_bounds__(128) airice_inice_attenuation_kernel(const InIceAttArgs a) {
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= a.n) return;
  const AirIceInIce m = inice_make_model(a.A, a.B, a.C);
  const InIceAttModel am = {a.A0, a.frequency, a.w0, a.w2, a.w};
  int flags = 0, worst = 0;
  a.out[i] = inice_total_attenuation(m, am, a.kind, a.z0[i], a.z1[i], a.kind == 2 ? a.zmax[i] : 0.0, a.L[i], flags, worst);
  if (a.quad_stats) {
    if (flags) atomicAdd(a.quad_stats, 1);
    if (worst > 24) atomicMax(a.quad_stats + 1, worst);
  }
}

__global__ void __launch_bounds__(256) airice_inice_focusing_kernel(const InIceFocusArgs a) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  const AirIceInIce m = inice_make_model(a.A, a.B, a.C);
  const double path_a[2] = {a.sol_a[2][i], a.sol_a[3][i]}, launch_a[2] = {a.sol_a[4][i], a.sol_a[5][i]};
  const double recv_a[2] = {a.sol_a[6][i], a.sol_a[7][i]};
  const double launch_b[2] = {a.sol_b[4][i], a.sol_b[5][i]}, recv_b[2] = {a.sol_b[6][i], a.sol_b[7][i]};
  double f[2] = {1, 1};                      // the reference's callers pass {1, 1} (IceRayTracing.cc:2666)
  inice_focusing(m, a.zT[i], a.zR[i], path_a, launch_a, recv_a, launch_b, recv_b, f);
  a.out[0][i] = f[0]; a.out[1][i] = f[1];
}

__global__ void __launch_bounds__(256) airice_inice_shift_kernel(int64_t n, const double* in, double shift, double* out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = in[i] + shift;
}

// grid nodes of IceRayTracing::MakeTable: node = ix * n_z + iz, xT = start_x + step_x * ix, zT = start_z + step_z * iz
__global__ void __launch_bounds__(256) airice_inice_table_nodes_kernel(const InIceTableNodeArgs a) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  const int64_t node = a.node0 + i;
  const int ix = (int)(node / a.n_z), iz = (int)(node - (int64_t)ix * a.n_z);
  a.xT[i] = a.start_x + a.step_x * ix;
  a.zT[i] = a.start_z + a.step_z * iz;
  a.rx[i] = a.zR;
}

// GridZValueb[AntNum][0..12] (IceRayTracing.cc:2680-2715): ray 1 {time, path, launch, receive, attenuation, focusing},
// ray 2 {same six, incidence angle on the surface}; -1000 where the ray is absent; NaN focusing -> 1
__global__ void __launch_bounds__(256) airice_inice_table_pack_kernel(const InIceTablePackArgs a) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  double f0 = a.focusing[0][i], f1 = a.focusing[1][i];
  if (f0 != f0) f0 = 1;
  if (f1 != f1) f1 = 1;
  double v[AIRICE_INICE_TABLE_NCOLS];
  if (a.ignore[0][i] != 0) {
    v[0] = a.sol[0][i]; v[1] = a.sol[2][i]; v[2] = a.sol[4][i]; v[3] = a.sol[6][i]; v[4] = a.att[0][i]; v[5] = f0;
  } else {
    for (int k = 0; k < 6; k++) v[k] = -1000;
  }
  if (a.ignore[1][i] != 0) {
    v[6] = a.sol[1][i]; v[7] = a.sol[3][i]; v[8] = a.sol[5][i]; v[9] = a.sol[7][i]; v[10] = a.att[1][i]; v[11] = f1;
    const double inc = a.sol[9][i];
    v[12] = (inc != 100) ? inc : -1000;
  } else {
    for (int k = 6; k < 13; k++) v[k] = -1000;
  }
#pragma unroll
  for (int k = 0; k < AIRICE_INICE_TABLE_NCOLS; k++) a.col[k][i] = v[k];
}

__global__ void __launch_bounds__(256) airice_inice_table_interp_kernel(const InIceTableInterpArgs a) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  a.out[i] = inice_table_interp(a.pos_x, a.pos_z, a.n_x, a.n_z, a.step_x, a.step_z, a.col, a.x[i], a.z[i]);
}

int ladder_grid() {
  static int blocks = 0;    // per process; every context of a process sits on the same GPU model
  if (blocks == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, airice_inice_ladder_kernel, kLadderThreads, 0);
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
  cudaError_t e = cudaMemsetAsync(a.ra_count, 0, 8 * sizeof(int32_t), s);     // 3 list counters, pad, 2 x u64 work census
  if (e != cudaSuccess) return e;
  const int64_t dr_blocks = (a.n + kDrThreads - 1) / kDrThreads;
  airice_inice_dr_kernel<<<dim3((unsigned)dr_blocks), kDrThreads, 0, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // the list length stays on the device: pass 2 is a persistent grid (one wave), pass 3 is launched over the worst-case
  // length and its blocks beyond the list exit at once
  int64_t lb = ladder_grid();
  if (lb > blocks) lb = blocks;
  airice_inice_ladder_kernel<<<dim3((unsigned)lb), kLadderThreads, 0, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  airice_inice_ra_finish_kernel<<<dim3((unsigned)blocks), kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_inice_attenuation(const InIceAttArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  if (!a.z0 || !a.z1 || !a.L || !a.out || (a.kind == 2 && !a.zmax) || a.kind < 0 || a.kind > 2) return cudaErrorInvalidValue;
  airice_inice_attenuation_kernel<<<dim3((unsigned)((a.n + 127) / 128)), 128, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_inice_focusing(const InIceFocusArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  airice_inice_focusing_kernel<<<dim3((unsigned)((a.n + 255) / 256)), 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_inice_shift(int64_t n, const double* in, double shift, double* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  airice_inice_shift_kernel<<<dim3((unsigned)((n + 255) / 256)), 256, 0, s>>>(n, in, shift, out);
  return cudaGetLastError();
}
cudaError_t launch_inice_table_nodes(const InIceTableNodeArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  airice_inice_table_nodes_kernel<<<dim3((unsigned)((a.n + 255) / 256)), 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_inice_table_pack(const InIceTablePackArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  airice_inice_table_pack_kernel<<<dim3((unsigned)((a.n + 255) / 256)), 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_inice_table_interp(const InIceTableInterpArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  airice_inice_table_interp_kernel<<<dim3((unsigned)((a.n + 255) / 256)), 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_inice_pick(const InIcePickArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  const int64_t blocks = (a.n + 255) / 256;
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  for (int k = 0; k < AIRICE_INICE_NCOLS; k++)
    if (!a.in[k]) return cudaErrorInvalidValue;
  if (!a.ignore[0] || !a.ignore[1]) return cudaErrorInvalidValue;
  if (a.att[0] && a.att[1]) airice_inice_pick_kernel<true><<<dim3((unsigned)blocks), 256, 0, s>>>(a);
  else airice_inice_pick_kernel<false><<<dim3((unsigned)blocks), 256, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace airice
