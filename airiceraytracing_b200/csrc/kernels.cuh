// kernels.cuh -- launch-side declarations of the three sm_100a kernels (+ the FP64 FMA peak probe).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "airice_core.cuh"

struct AirIcePathPlan;   // airice_path.cuh

#define AIRICE_TABLE_NCOLS64 17
#define AIRICE_TABLE_NCOLS32 11
#define AIRICE_SOLVE_NCOLS 13
#define AIRICE_LOOKUP_NCOLS 9

namespace airice {

// ---- kernel 1: forward table.  One thread per (Tx height, launch angle) cell, flat cell index
// cell = ihei*n_th + iang exactly as MakeRayTracingTable orders its push_backs (M.cc:2079-2118).
struct TableArgs {
  int64_t cell0;      // first flat cell of this launch
  int64_t ncells;     // cells in this launch
  int64_t n_th;       // TotalAngleSteps
  int64_t row0;       // row index of row_h[0]
  const double* row_h;    // per-row Tx height (host-built: M.cc:2080,2089-2091)
  const double* row_ntx;  // per-row n(h) (host libm)
  const int* row_kt;      // per-row top layer (-1: outside every layer)
  double th_start, th_step, th_stop;
  int in_ice;
  // f64 columns = dummy[1..17] of GetRayTracingSolutions (M.cc:1999-2016):
  //  0 h, 1 X, 2 X_air, 3 X_ice, 4 opt, 5 opt_air, 6 opt_ice, 7 t[ns], 8 t_air[ns], 9 t_ice[ns],
  //  10 launch, 11 incident, 12 received, 13 T_S, 14 T_P, 15 geo_air, 16 geo_ice.  nullptr = not wanted.
  double* c64[AIRICE_TABLE_NCOLS64];
  // float columns in the reference's AllTableAllAntData order (M.cc:2101-2111):
  //  0 h, 1 X, 2 opt_ice, 3 opt_air, 4 launch, 5 X_air, 6 T_S, 7 T_P, 8 geo_air, 9 geo_ice, 10 received.
  float* c32[AIRICE_TABLE_NCOLS32];
};
cudaError_t launch_table(const AirIceMedium& m, const AirIcePlan& p, const TableArgs& a, cudaStream_t s);

// kernel 1c: the float tables of several in-ice antennas in one pass: the air walk of a cell does not
// depend on the receiver depth (M.cc:887-905, 1796-1879), so it runs once and only the ice leg is repeated per antenna.
// `base` supplies the grid (c64/c32 unused); antenna q gets its 11 float columns at blocks[q] + k * ncols_stride.
struct TableMultiArgs {
  TableArgs base;
  int n_ant;
  const double* ant;       // device [n_ant][2]: receiver depth (positive, m) and n_ice(depth) (host libm)
  float* const* blocks;    // device [n_ant]: column-major float block of each antenna's table; nullptr = lookup layout only
  int64_t col_stride;      // elements between two columns of a block (= cells of the whole table)
  // lookup layout of each antenna's table written in the same pass (what airice_pack_kernel derives from the columns);
  // device [n_ant] each, all three null = columns only.  Indexed by the GLOBAL cell (base.cell0 + i).
  float4* const* rec;
  float* const* x;
  float* const* row_h;
};
// per-row products of a packed table (trim ranges, position table, header blocks): what the fused pass cannot do per cell.
// Same-shape tables are prepared together, grid.y = table.
#define AIRICE_ROWPREP_MAX 32
struct RowPrepTab {
  const float* x; const float* row_h;
  int* row_first; int* row_last;
  float* rowblk; float4* rowpar; uint16_t* lut;
};
struct RowPrepBatch {
  int n_tab;
  int64_t cells;
  int n_h, n_th, lut_shift;
  RowPrepTab tab[AIRICE_ROWPREP_MAX];
};
cudaError_t launch_row_prep(const RowPrepBatch& b, cudaStream_t s);
cudaError_t launch_row_ranges(const float* x, const float* row_h, int64_t cells, int n_h, int n_th, int* row_first, int* row_last,
                              float* rowblk, float4* rowpar, uint16_t* lut, int lut_shift, cudaStream_t s);
int lut_shift_for(int64_t n_th);   // fraction bits of the position table (-1: rows too long for u16 positions, no table)
cudaError_t launch_table_multi(const AirIceMedium& m, const AirIcePlan& p, const TableMultiArgs& a, cudaStream_t s);

// kernel 1b: the same forward tracer on arbitrary (theta, h) cells (batched GetRayTracingSolutions, M.cc:1796-2017);
// n(h) of the transmitter is evaluated on the device here.
struct ForwardArgs {
  int64_t n;
  const double* theta;
  const double* h;
  int in_ice;
  double* c64[AIRICE_TABLE_NCOLS64];
};
cudaError_t launch_forward(const AirIceMedium& m, const AirIcePlan& p, const ForwardArgs& a, cudaStream_t s);

// ---- kernel 2: batched launch-angle solve.  One thread per Tx->Rx pair.
enum { AIRICE_UNITS_M_DEG = 0, AIRICE_UNITS_CM_RAD = 1 };
struct SolveArgs {
  int64_t n;
  const double* h;  // Tx heights a.s.l.
  const double* d;  // horizontal distances
  const double* straight;  // optional: caller-supplied straight-line angle per pair [deg] (the StraightAngle argument of
                           // Air2IceRayTracing, M.h:191); nullptr = computed as at M.cc:952-958
  double ice;       // ice-surface height (same units as h)
  double depth;     // signed receiver depth (negative = in ice), same units
  int units;        // AIRICE_UNITS_M_DEG: metres in, 13 columns out (m, s, deg);
                    // AIRICE_UNITS_CM_RAD: cm in, the 9 outputs of GetHorizontalDistanceToIntersectionPoint (M.h:170)
  // M_DEG columns: 0 X_total, 1 X_air, 2 X_ice, 3 t_air[s], 4 t_ice[s], 5 launch, 6 received(ice), 7 T_S, 8 T_P,
  //                9 geo_air, 10 geo_ice, 11 incident on ice, 12 refracted angle below the surface (P.cc:1081)
  // CM_RAD columns: 0 opt_ice, 1 opt_air, 2 geo_ice, 3 geo_air, 4 launch[rad], 5 X_air, 6 T_S, 7 T_P, 8 received[rad]
  double* out[AIRICE_SOLVE_NCOLS];
  uint8_t* ok;        // solution flag (M.cc:974-983)
  int32_t* nevals;    // optional: distance evaluations spent (Newton + replay), diagnostics
  // scratch of the two-pass launch (both or neither): a device counter and room for n pair indices.  With them,
  // launch_solve lists the pairs that need a rare slow path in a first pass and solves those in a second, dense one.
  int32_t* defer_count;
  int32_t* defer_list;
};
cudaError_t launch_solve(const AirIceMedium& m, const AirIcePlan& p, const SolveArgs& a, cudaStream_t s);

// ---- kernel 3: table lookup (GetHorizontalDistanceToIntersectionPoint_Table, M.cc:1305-1462)
// Lookup-side layout (built once per table by launch_pack_table): the reference's column-major float table makes one
// query touch 40+ different 32-byte sectors (10 parameters x 2 bins x 2 rows, one useful float per sector; ncu showed
// 2.3 KB of DRAM traffic per lookup).  Here the THD column stays dense for the index search, and the ten
// interpolated parameters of a cell sit together in one 48-byte record, so the two bins of a row are one contiguous
// 96-byte read.
struct LookupTable {
  const float* x;          // column 1 (total horizontal distance), dense, for FindClosestTHD
  const float4* rec;       // 3 float4 per cell: {X, opt_ice, opt_air, launch | X_air, T_S, T_P, geo_air | geo_ice, received, 0, 0}
  const float* row_h;      // Tx height of each row (column 0 is constant along a row)
  int64_t cells;
  int n_h, n_th;
  double loop_stop_h, h_step;
  const int* row_first;  // per row: first/last bin with a usable X (trim of M.cc:1050-1072), precomputed
  const int* row_last;
  // per-row header, AIRICE_ROWBLK floats (one 64-byte block): {s1, e1 (int bits), X[s1], X[s2], h(s1), h(s2), col0[row],
  // flags (int bits: 1 = the row's window can be answered from the position table, 2 = the second row's) |
  // xm, u_lo, scale, 0 of the row | the same of the second row}
  const float* rowblk;
  // per-row position table (round 2): AIRICE_LUT_EDGES u16 per PHYSICAL row.  Entry k is the (fractional, lut_shift
  // fraction bits) bin of the row at which u(X) = X / (X + xm) falls through u_lo + k / scale; a query's bin of
  // FindClosestTHD is predicted by linear interpolation between two entries and then VERIFIED on the records it is
  // going to read anyway (X[i2] <= d < X[i1] on a strictly decreasing window is the literal search's result), so the
  // eight dependent index halvings and the linear scan in the dense column are the fallback, not the rule.
  const uint16_t* lut;
  int lut_shift;
};
#define AIRICE_ROWBLK 16
#ifndef AIRICE_LUT_EDGES
#define AIRICE_LUT_EDGES 128
#endif
cudaError_t launch_unpack_table(const float4* rec, const float* row_h, int64_t cells, int n_th, float* c0, int64_t stride, cudaStream_t s);
cudaError_t launch_pack_table(const float* const* cols32, int64_t cells, int n_h, int n_th, float* x, float4* rec,
                              float* row_h, int* row_first, int* row_last, float* rowblk, float4* rowpar, uint16_t* lut,
                              int lut_shift, cudaStream_t s);
struct LookupArgs {
  int64_t n;
  const double* h_cm;
  const double* d_cm;
  double* out[AIRICE_LOOKUP_NCOLS];  // same 9 slots as the CM_RAD solve
  uint8_t* ok;                       // solution flag (M.cc:1356-1449)
  int literal;                       // test hook: 1 = always run the literal index search (AIRICE_LOOKUP_LITERAL=1)
};
cudaError_t launch_lookup(const AirIceMedium& m, const LookupTable& t, const LookupArgs& a, cudaStream_t s);

// ---- kernel 4: in-ice direct / reflected / refracted solver (IceRayTracing::IceRayTracing, IceRayTracing.cc:1745-1919)
#define AIRICE_INICE_NCOLS 29
struct InIceArgs {
  int64_t n;
  const double* z0;   // Tx depth (negative), Rx depth (negative), horizontal distance; Tx at x = 0
  const double* x1;
  const double* z1;
  double A, B, C;     // ice model n(z) = A + B exp(-C |z|)
  double* out[AIRICE_INICE_NCOLS];  // the 29 slots of the reference's output array, SoA; nullptr = skip
  uint8_t* mask;      // bit0 D, bit1 R, bit2 Ra1, bit3 Ra2: which branches exist (receive-angle slot != -1000)
  // scratch owned by the caller (context): compaction list of the pairs whose refracted-ray ladder must run, and its
  // length.  Pass 1 (all pairs: direct + reflected) appends to it; pass 2 (persistent lanes stepping the root-search
  // state machine, each taking the next list entry when its own is finished) and pass 3 (times, paths, angles) walk it.
  int32_t* ra_list;   // [n]
  int32_t* ra_count;  // [8], zeroed by launch_inice: [0] entries at the front of the list (pairs searching for two
                      // refracted rays), [1] next list position to hand out (pass 2), [2] entries at the back; [4..7] =
                      // two u64: fRaa evaluations and turning-depth falsepos steps pass 2 ran (work census)
  double* ra_lad;     // [6][n] ladder results (L, f(L), z_max of the two candidate roots) per list entry, pass 2 -> pass 3
};
cudaError_t launch_inice(const InIceArgs& a, cudaStream_t s);

// ---- kernel 4b: the two physical rays of a pair (IceRayTracing::GetRayTracingSolutions, IceRayTracing.cc:2907-3210,
// without its attenuation integrals), from the 29 columns of kernel 4
#define AIRICE_INICE_RAYS_NCOLS 10
struct InIcePickArgs {
  int64_t n;
  const double* rx_depth;
  const double* distance;
  const double* tx_depth;
  double A, B, C;
  const double* in[AIRICE_INICE_NCOLS];    // kernel 4 output, SoA
  double* out[AIRICE_INICE_RAYS_NCOLS];    // TimeRay[2], PathRay[2], LaunchAngle[2], RecieveAngle[2], IncidenceAngleInIce[2]
  int32_t* ignore[2];                      // IgnoreCh[2]: 1 = ray present
  int32_t* type[2];                        // RayType[2] (1 D, 2 R, 3 Ra1, 4 Ra2); nullptr = skip
  // attenuation (IceRayTracing.cc:2970-2987): att[0] != nullptr -> AttRay[2] = 1 - integral of A0 / L_att along the ray
  double* att[2];
  double A0, frequency, w0, w2, w;         // w0 = log(0.0001), w2 = log(3.16), w = log(frequency): host libm
  int32_t* quad_stats;                     // [2] device: integrals that ran out of interval storage, largest interval count
};
cudaError_t launch_inice_pick(const InIcePickArgs& a, cudaStream_t s);

// ---- kernel 4c: GetTotalAttenuationDirect / Reflected / Refracted (IceRayTracing.cc:203-219) for n rays
struct InIceAttArgs {
  int64_t n;
  int kind;                                // 0 direct, 1 reflected, 2 refracted
  const double *z0, *z1, *zmax, *L;        // zmax: refracted only
  double A, B, C, A0, frequency, w0, w2, w;
  double* out;
  int32_t* quad_stats;
};
cudaError_t launch_inice_attenuation(const InIceAttArgs& a, cudaStream_t s);

// ---- kernel 4d: GetFocusingFactor (IceRayTracing.cc:3218-3293) from the two-ray solutions at zR (a) and zR - 0.01 (b)
struct InIceFocusArgs {
  int64_t n;
  const double *zT, *zR;
  const double* sol_a[AIRICE_INICE_RAYS_NCOLS];   // kernel 4b columns
  const double* sol_b[AIRICE_INICE_RAYS_NCOLS];
  double A, B, C;
  double* out[2];
};
cudaError_t launch_inice_focusing(const InIceFocusArgs& a, cudaStream_t s);
// rx - 0.01 for the second solution; grid nodes and the 13 columns of IceRayTracing::MakeTable (IceRayTracing.cc:2614-2724)
cudaError_t launch_inice_shift(int64_t n, const double* in, double shift, double* out, cudaStream_t s);
struct InIceTableNodeArgs {
  int64_t node0, n;
  int n_z;
  double start_x, start_z, step_x, step_z, zR;
  double *xT, *zT, *rx;                    // outputs [n]
};
cudaError_t launch_inice_table_nodes(const InIceTableNodeArgs& a, cudaStream_t s);
#define AIRICE_INICE_TABLE_NCOLS 13
struct InIceTablePackArgs {
  int64_t n;
  const double* sol[AIRICE_INICE_RAYS_NCOLS];
  const double* att[2];
  const int32_t* ignore[2];
  const double* focusing[2];
  double* col[AIRICE_INICE_TABLE_NCOLS];   // at this chunk's first node
};
cudaError_t launch_inice_table_pack(const InIceTablePackArgs& a, cudaStream_t s);
struct InIceTableInterpArgs {
  int64_t n;
  const double *x, *z;
  const float *pos_x, *pos_z;
  int n_x, n_z;
  double step_x, step_z;
  const double* col;
  double* out;
};
cudaError_t launch_inice_table_interp(const InIceTableInterpArgs& a, cudaStream_t s);

// ---- kernel 5: ray-path emission (SingleRayAirIceRefraction.C:226-299), airice_path.cuh
struct PathArgs {
  int64_t n;             // rays
  const double* theta;   // launch angle, deg from the upward vertical (> 90)
  const double* h;       // Tx height, m
  int64_t max_points;    // row length of x / z
  double* x;             // [n][max_points] horizontal distance from the Tx; unused tail entries are NaN
  double* z;             // [n][max_points] height above sea level
  int32_t* count;        // [n] points of the full path (may exceed max_points: the row then holds the first max_points)
  ::AirIcePathPlan* plans; // [n] scratch
};
size_t path_plan_bytes();
cudaError_t launch_ray_path(const AirIceMedium& m, const AirIcePlan& p, const PathArgs& a, cudaStream_t s);

// ---- kernel 6: the old solve-per-cell table (MakeTable, M.cc:1618-1696) and its inverse-distance lookup
// (GetInterpolatedValue, M.cc:1700-1794); oldtable_kernels.cu
struct OldGridArgs {
  int64_t cell0, n;        // nodes [cell0, cell0 + n) of the grid, node = ih * n_th + ith
  int n_th;
  const double* pos_h;     // GridPositionH [n_h] (host-made: start + step * ih, last snapped, M.cc:1646-1655)
  const double* pos_th;    // GridPositionTh [n_th]
  const double* col_tan;   // tan((180 - th) * (pi / 180)) per angle node (host libm)
  double ice, depth;       // m; depth signed as the caller passed it
  double *h, *d, *th;      // outputs [n]: the solver's inputs (Tx height, distance, StraightAngle)
};
cudaError_t launch_oldgrid_cells(const OldGridArgs& a, cudaStream_t s);
struct OldPackArgs {
  int64_t n;
  const double *h, *d;                                          // node height and requested distance
  const double *x, *x_air, *t_air, *t_ice, *launch, *ts, *tp, *inc;   // kernel 2, M_DEG columns 0, 1, 3, 4, 5, 7, 8, 11
  double c;                                                     // spedc: optical path = time * c
  double* col[9];                                               // GridZValue[0..8] at this chunk's first node
};
cudaError_t launch_oldgrid_pack(const OldPackArgs& a, cudaStream_t s);
struct OldInterpArgs {
  int64_t n;
  const double *h, *th;    // queries (m, deg)
  const double* z;         // the GridZValue column asked for
  const double *pos_h, *pos_th;
  int n_h, n_th;
  double start_h, start_th, step_h, step_th;
  double* out;
};
cudaError_t launch_old_interp(const OldInterpArgs& a, cudaStream_t s);

// ---- FP64 FMA peak probe (roofline denominator; MEASURED_PEAKS.json has no FP64 figure)
cudaError_t fp64_peak_probe(double* tflops_out, int iters, cudaStream_t s);

}  // namespace airice
