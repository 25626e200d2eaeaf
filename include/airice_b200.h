/* airice_b200.h -- C ABI of the B200-native air->ice ray solver (libairice_b200.so).
 *
 * This is the drop-in boundary for the hot path of uzairlatif90/AirIceRayTracing: plain pointers and sizes,
 * no C++/torch types.  Each entry point names the reference interface it replaces (paths are relative to the
 * reference repository root).  Device entry points take DEVICE pointers and a cudaStream_t passed as void*
 * (NULL = default stream) and do not synchronise; host entry points take HOST pointers (pinned or pageable), copy
 * chunk by chunk on the context's own two streams and return when the results are in the caller's buffers.
 *
 * All functions return 0 on success or a negative error code; airice_last_error() describes the last failure
 * on the calling thread.  A context is bound to one GPU and may be used from one thread at a time.
 */
#ifndef AIRICE_B200_H
#define AIRICE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct airice_ctx airice_ctx;
typedef struct airice_table airice_table;

enum { AIRICE_VARIANT_MULTIRAY = 0, /* MultiRayAirIceRefraction.{h,cc}: pi = 3.1415927 (MultiRayAirIceRefraction.h:29) */
       AIRICE_VARIANT_PYWRAP = 1,   /* pythonwrapper/AirIceRayTracing.{h,cc}: pi = 4*atan(1) (AirIceRayTracing.h:25) */
       AIRICE_VARIANT_CLI = 2       /* the command-line solver Air2IceRayTracing.C on RayTracingFunctions.{h,cc}: pi = 3.1415927
                                     * (RayTracingFunctions.h:26), gsl_root_fsolver_brent with tolerance 1e-9 and at most 20
                                     * iterations (Air2IceRayTracing.C:137, RayTracingFunctions.cc:259), bracket rule of
                                     * Air2IceRayTracing.C:101-129 (lo < 90.00 -> 90.05, stepping while lo <= hi - 1).  Affects
                                     * airice_solve_*; everything else behaves as variant 0. */ };
enum { AIRICE_UNITS_M_DEG_C = 0, AIRICE_UNITS_CM_RAD_C = 1 };

#define AIRICE_TABLE_COLS64 17
#define AIRICE_TABLE_COLS32 11
#define AIRICE_SOLVE_COLS 13
#define AIRICE_SOLVE_COLS_CM_RAD 9
#define AIRICE_LOOKUP_COLS 9

/* ---- context: replaces MakeAtmosphere() (MultiRayAirIceRefraction.cc:920-942, .h:157) and
 * AirIceRayTracing::MakeAtmosphere(file) (pythonwrapper/AirIceRayTracing.cc:860-882).  Parses the GDAS file once. */
/* atmosphere_path NULL or "": an ice-only context (the reference's IceRayTracing namespace needs no atmosphere file);
 * the air->ice entry points then fail with -9. */
int airice_create(const char *atmosphere_path, int variant, int device, airice_ctx **out);
void airice_destroy(airice_ctx *ctx);
const char *airice_last_error(void);
int airice_device_count(void);
/* out[0]=MaxLayers, out[1..5]=ATMLAY[cm], out[6..10]=B_air, out[11..15]=C_air, out[16..18]=A,B,C of ice,
 * out[19]=pi, out[20]=spline n(h=0), out[21]=#spline knots */
int airice_get_medium(const airice_ctx *ctx, double out[24]);
/* the reference's mutable ice-model globals A_ice/B_ice/C_ice (MultiRayAirIceRefraction.h:72-74) */
int airice_set_ice_model(airice_ctx *ctx, double A, double B, double C);

/* ---- kernel 1: forward table = MakeRayTracingTable (MultiRayAirIceRefraction.cc:2019-2158, .h:204) and its
 * per-cell worker GetRayTracingSolutions (MultiRayAirIceRefraction.cc:1796-2017, .h:198).
 * Grid: heights h_top, h_top-h_step, ... down to the surface (last row snapped), angles th_start, +th_step, ...
 * (last bin snapped to th_stop); cell = row*n_th + bin.  All lengths in metres, angles in degrees, depth negative
 * for a receiver in ice.  The reference's own grid is (100000, 10, 90.1, 0.1, 180)  (MultiRayAirIceRefraction.cc:12-18,2044). */
int airice_table_dims(const airice_ctx *ctx, double depth_m, double ice_m, double h_top, double h_step,
                      double th_start, double th_step, double th_stop, int64_t *n_h, int64_t *n_th);
/* Builds rows [row_begin,row_end) into caller-owned DEVICE columns (SoA; each column holds (row_end-row_begin)*n_th
 * entries).  cols64: 17 pointers = dummy[1..17] of GetRayTracingSolutions, cols32: 11 pointers in the reference's
 * AllTableAllAntData float layout; either array may be NULL, individual f64 columns may be NULL. */
int airice_table_build_device(airice_ctx *ctx, double depth_m, double ice_m, double h_top, double h_step,
                              double th_start, double th_step, double th_stop, int64_t row_begin, int64_t row_end,
                              double *const *cols64, float *const *cols32, void *stream);
/* Forward-traces arbitrary (theta, h) cells (batched GetRayTracingSolutions): DEVICE inputs, 17 f64 SoA outputs. */
int airice_forward_device(airice_ctx *ctx, int64_t n, const double *d_theta, const double *d_h, double depth_m,
                          double ice_m, double *const *cols64, void *stream);

/* Same through HOST buffers: out is a dense SoA block out[col*n + i] with 17 columns. */
int airice_forward_host(airice_ctx *ctx, int64_t n, const double *theta, const double *h, double depth_m, double ice_m,
                        double *out);

/* Library-owned float table for lookups = one entry of AllTableAllAntData (MultiRayAirIceRefraction.cc:9,2136).
 * (A receiver in the ice is built by the fused pass of airice_table_create_multi with one antenna.) */
int airice_table_create(airice_ctx *ctx, double depth_m, double ice_m, double h_top, double h_step, double th_start,
                        double th_step, double th_stop, airice_table **out);
/* The tables of n_ant in-ice antennas (depths_m[q] < 0) in ONE pass: the air walk of a cell does not depend on the
 * receiver depth (MultiRayAirIceRefraction.cc:887-905, 1796-1879), so it runs once per cell and only the ice leg is
 * repeated per antenna.  Replaces n_ant calls of MakeRayTracingTable (MultiRayAirIceRefraction.cc:2019-2158; one
 * AllTableAllAntData entry each, :2136); out[q] is bit-identical to airice_table_create(depths_m[q], ...). */
int airice_table_create_multi(airice_ctx *ctx, int n_ant, const double *depths_m, double ice_m, double h_top,
                              double h_step, double th_start, double th_step, double th_stop, airice_table **out);
/* Wraps 11 caller-owned DEVICE float columns (e.g. a gathered multi-GPU table or a reference-built table). */
int airice_table_wrap(airice_ctx *ctx, const float *const *d_cols32, int64_t n_h, int64_t n_th, double loop_stop_h,
                      double h_step, airice_table **out);
void airice_table_destroy(airice_table *t);
/* info[0]=n_h, info[1]=n_th, info[2]=cells, info[3]=11 */
int airice_table_info(const airice_table *t, int64_t info[4]);
int airice_table_copy_column(const airice_table *t, int col, float *host_out);
int airice_table_column_ptr(const airice_table *t, int col, const float **d_ptr);
/* Persistence: the reference keeps AllTableAllAntData in memory only and rebuilds every table per
 * process (50-90 s each on a CPU).  save writes the 11 float columns with a 64-byte versioned header (dimensions,
 * loop_stop_h, h_step, checksum); load rebuilds the device table and its lookup layout from such a file -- lookups on the
 * loaded table return the bits of the original.  Errors: -7 I/O, -8 not a table file / wrong version / damaged. */
int airice_table_save(const airice_table *t, const char *path);
int airice_table_load(airice_ctx *ctx, const char *path, airice_table **out);
/* row trim ranges (FindClosestAirTxHeight's StartBin/EndBin scans, MultiRayAirIceRefraction.cc:1050-1072), per row */
int airice_table_copy_row_ranges(const airice_table *t, int32_t *host_first, int32_t *host_last);

/* ---- kernel 2: batched launch-angle solve = Air2IceRayTracing (MultiRayAirIceRefraction.cc:1464-1616, .h:191) under
 * GetHorizontalDistanceToIntersectionPoint (MultiRayAirIceRefraction.cc:945-989, .h:170); variant 1 follows
 * AirIceRayTracing::GetRayTracingSolution (pythonwrapper/AirIceRayTracing.cc:884-1086).
 * units = CM_RAD: inputs in cm, 9 output columns in the order of the reference's by-reference arguments
 *   (opt ice, opt air, geo ice, geo air, launch[rad], X_air, T_S, T_P, received[rad]);
 * units = M_DEG: inputs in m, 13 columns (X, X_air, X_ice, t_air[s], t_ice[s], launch, received, T_S, T_P, geo air,
 *   geo ice, incident on ice, refracted below surface).  out: array of column pointers, NULL entries are skipped.
 * ok: the reference's bool (|X-d| test, MultiRayAirIceRefraction.cc:974-983).  nevals: optional diagnostics.
 * d_straight: optional per-pair straight-line angle (deg for M_DEG, rad for CM_RAD) = the StraightAngle argument of
 *   Air2IceRayTracing; NULL = computed from the geometry as GetHorizontalDistanceToIntersectionPoint does.
 * Launches of 6e6 pairs and more run as two kernels on `stream` (the second solves the ~0.6 % of pairs that need a
 * rare slow path, listed by the first); their scratch -- 4 bytes per pair -- belongs to the context, one per stream. */
int airice_solve_device(airice_ctx *ctx, int64_t n, const double *d_h, const double *d_dist, const double *d_straight,
                        double depth, double ice, int units, double *const *d_out, uint8_t *d_ok, int32_t *d_nevals,
                        void *stream);
/* Multi-antenna form (CoREAS: n_points shower points x n_ant receivers, the shape of RunMultiRayCode_loop.C with
 * several AntennaDepths): d_dist and every output column are antenna-major [n_ant][n_points]; depths_host[n_ant] is a
 * HOST array of signed receiver depths.  One kernel launch per receiver; the launches alternate between two streams
 * of the context that are forked from `stream` and joined back to it by events before the call returns. */
int airice_solve_multi_device(airice_ctx *ctx, int64_t n_points, int n_ant, const double *d_h, const double *d_dist,
                              const double *depths_host, double ice, int units, double *const *d_out, uint8_t *d_ok,
                              void *stream);
/* Same through HOST buffers: out is a dense SoA block out[col*n + i] with 9 (CM_RAD) or 13 (M_DEG) columns. */
int airice_solve_host(airice_ctx *ctx, int64_t n, const double *h, const double *dist, const double *straight,
                      double depth, double ice, int units, double *out, uint8_t *ok);
/* The same with one HOST pointer per output column (9 for CM_RAD, 13 for M_DEG, in the column order above): a NULL column
 * is neither stored by the kernel nor copied back, and `ok` may be NULL.  The host path is bound by the PCIe link
 * (73 B of results per pair when every column travels), so a caller that reads a subset -- TraceIceToAir.C:31-68 uses
 * 6 of the 13 metre/degree values of Air2IceRayTracing (AirIceRayTracing.cc:1066-1084) -- gets its answers that much
 * sooner. */
int airice_solve_host_columns(airice_ctx *ctx, int64_t n, const double *h, const double *dist, const double *straight,
                              double depth, double ice, int units, double *const *cols, uint8_t *ok);

/* ---- kernel 3: table lookup = GetHorizontalDistanceToIntersectionPoint_Table
 * (MultiRayAirIceRefraction.cc:1305-1462, .h:189) with FindClosestAirTxHeight / FindClosestTHD / GetParValues
 * (MultiRayAirIceRefraction.cc:1033-1302).  cm/rad in and out, 9 columns as for CM_RAD solves.  The bin FindClosestTHD's
 * index halvings + scan end on is predicted from a per-row position table and verified on the records the interpolation
 * reads (same indices, same bits); AIRICE_LOOKUP_LITERAL=1 in the environment forces the literal search (test hook). */
int airice_lookup_device(airice_ctx *ctx, const airice_table *t, int64_t n, const double *d_h_cm,
                         const double *d_dist_cm, double *const *d_out, uint8_t *d_ok, void *stream);
int airice_lookup_host(airice_ctx *ctx, const airice_table *t, int64_t n, const double *h_cm, const double *dist_cm,
                       double *out, uint8_t *ok);
/* The same with one HOST pointer per output column (NULL = not wanted: neither stored nor copied back; ok may be NULL),
 * like airice_solve_host_columns. */
int airice_lookup_host_columns(airice_ctx *ctx, const airice_table *t, int64_t n, const double *h_cm,
                               const double *dist_cm, double *const *cols, uint8_t *ok);

/* ---- kernel 4: in-ice solver = IceRayTracing::IceRayTracing(0, z0, x1, z1) (IceRayTracing.cc:1745-1919; direct,
 * reflected and up to two refracted rays by GetDirectRayPar :626, GetReflectedRayPar :745, GetRefractedRayPar :923).
 * Depths negative, metres; out: 29 column pointers = the 29 slots of the reference's output array (launch angles 0-3,
 * times 4-7, receive angles 8-11 with -1000 = branch absent, sub-times 12-17, incidence 18, L 19-22, z_max 23-24,
 * geometric paths 25-28); mask bit0..3 = D, R, Ra1, Ra2 present (the solution-branch count is its popcount).
 * The ice model is the context's (airice_set_ice_model; IceRayTracing::SetA/SetB/SetC in the reference). */
#define AIRICE_INICE_COLS 29
int airice_inice_solve_device(airice_ctx *ctx, int64_t n, const double *d_z0, const double *d_x1, const double *d_z1,
                              double *const *d_out, uint8_t *d_mask, void *stream);
int airice_inice_solve_host(airice_ctx *ctx, int64_t n, const double *z0, const double *x1, const double *z1, double *out,
                            uint8_t *mask);

/* ---- kernel 4b: the two physical rays of a pair = IceRayTracing::GetRayTracingSolutions(RxDepth, Distance, TxDepth, ...)
 * (IceRayTracing.cc:2907-3210; declared IceRayTracing.hh) without its attenuation integrals (A0, frequency, AttRay are
 * post-processing outside the hot path): runs kernel 4 for (0, TxDepth, Distance, RxDepth), selects two of the
 * candidates D, R, Ra1, Ra2, orders them by arrival time and applies the same-depth straight-line patch.
 * out: 10 column pointers = TimeRay[0..1] [s], PathRay[0..1] [m], LaunchAngle[0..1], RecieveAngle[0..1] [deg, -1000 =
 * absent], IncidenceAngleInIce[0..1] [deg, 100 = none]; ignore: 2 int32 columns = IgnoreCh (1 = ray present);
 * type: optional 2 int32 columns = the reference's internal RayType (1 D, 2 R, 3 Ra1, 4 Ra2), may be NULL. */
#define AIRICE_INICE_RAYS_COLS 10
int airice_inice_two_rays_device(airice_ctx *ctx, int64_t n, const double *d_rx_depth, const double *d_distance,
                                 const double *d_tx_depth, double *const *d_out, int32_t *const *d_ignore,
                                 int32_t *const *d_type, void *stream);
/* HOST buffers: out[col*n + i] with 10 columns, ignore[k*n + i] with k = 0, 1 */
int airice_inice_two_rays_host(airice_ctx *ctx, int64_t n, const double *rx_depth, const double *distance,
                               const double *tx_depth, double *out, int32_t *ignore);

/* Work census of the LAST airice_inice_solve_device / two_rays call of the context (its refracted-ray ladder, pass 2):
 * out[0], out[1] = pairs that searched for two / one refracted root, out[2] = evaluations of the root function fRaa,
 * out[3] = falsepos steps of the nested turning-depth search.  Diagnostics for the roofline accounting; synchronises. */
int airice_inice_ladder_stats(airice_ctx *ctx, int64_t out[4]);

/* ---- kernels 4c-4e: attenuation, focusing factor and the in-ice interpolation table.
 * airice_inice_two_rays_att_*: IceRayTracing::GetRayTracingSolutions WITH its A0 / frequency / AttRay arguments
 * (IceRayTracing.cc:2907-3210): AttRay[k] = 1 - integral of A0 / L_att(z, f) along ray k (GetTotalAttenuationDirect /
 * Reflected / Refracted, IceRayTracing.cc:203-219; integrand :165-176 with the AraSim temperature and attenuation-length
 * model :135-162).  The reference integrates with gsl_integration_qags(epsabs 0, epsrel 1e-7, limit 1000)
 * (IceRayTracing.cc:179-200); the kernel runs the same QAGS procedure per ray (21-point Gauss-Kronrod, bisection by
 * largest error, epsilon extrapolation).  frequency in GHz.  att: 2 column pointers (device) / att[k*n + i] (host). */
int airice_inice_two_rays_att_device(airice_ctx *ctx, int64_t n, const double *d_rx_depth, const double *d_distance,
                                     const double *d_tx_depth, double A0, double frequency_ghz, double *const *d_out,
                                     double *const *d_att, int32_t *const *d_ignore, int32_t *const *d_type, void *stream);
int airice_inice_two_rays_att_host(airice_ctx *ctx, int64_t n, const double *rx_depth, const double *distance,
                                   const double *tx_depth, double A0, double frequency_ghz, double *out, double *att,
                                   int32_t *ignore);
/* GetTotalAttenuationDirect (kind 0) / Reflected (1) / Refracted (2; d_zmax = turning depth) for n rays with Snell
 * parameter L between depths z0 and z1 (IceRayTracing.cc:203-219). */
int airice_inice_attenuation_device(airice_ctx *ctx, int64_t n, int kind, double A0, double frequency_ghz,
                                    const double *d_z0, const double *d_z1, const double *d_zmax, const double *d_L,
                                    double *d_out, void *stream);
int airice_inice_attenuation_host(airice_ctx *ctx, int64_t n, int kind, double A0, double frequency_ghz, const double *z0,
                                  const double *z1, const double *zmax, const double *L, double *out);
/* out[0] = integrals that would have needed more than the 64 intervals a thread holds (the reference allows 1000; their
 * result is the running total instead of the extrapolated one), out[1] = largest interval count seen above 24.  Both 0 in
 * every test; a non-zero out[0] means results that may differ from the reference's. */
int airice_inice_quadrature_stats(airice_ctx *ctx, int64_t out[2]);
/* IceRayTracing::GetFocusingFactor(zT, xR, zR, focusing[2]) (IceRayTracing.cc:3218-3293) with the initial {1, 1} its
 * caller passes: two two-ray solutions (receiver at zR and at zR - 0.01 m) -> sqrt(path / (sin(recv) |dz / dlaunch|) nTx / nRx)
 * per ray.  out: 2 column pointers (device) / out[k*n + i] (host). */
int airice_inice_focusing_device(airice_ctx *ctx, int64_t n, const double *d_zT, const double *d_xR, const double *d_zR,
                                 double *const *d_out, void *stream);
int airice_inice_focusing_host(airice_ctx *ctx, int64_t n, const double *zT, const double *xR, const double *zR, double *out);
/* IceRayTracing::MakeTable(ShowerHitDistance, ShowerDepth, zR, AntNum) (IceRayTracing.cc:2614-2724): emitter positions
 * (xT, zT) on a grid around the shower (reference globals IceRayTracing.hh:33-36: steps 0.1 m, widths 40 m x 20 m =
 * 401 x 201 nodes), one receiver depth zR; per node the two-ray solution with attenuation (A0 = 1, 0.1 GHz) and the
 * focusing factors.  13 f64 columns = GridZValueb[AntNum][0..12]: ray 1 {time, path, launch, receive, AttRay, focusing},
 * ray 2 {the same six, incidence angle on the surface}; -1000 = absent.  node = ix * n_z + iz; positions are float. */
typedef struct airice_inice_table airice_inice_table;
#define AIRICE_INICE_TABLE_COLS 13
int airice_inice_table_create(airice_ctx *ctx, double shower_hit_distance, double shower_depth, double zR, double step_x,
                              double step_z, double width_x, double width_z, airice_inice_table **out);
void airice_inice_table_destroy(airice_inice_table *t);
/* info[0] = TotalStepsX_O, info[1] = TotalStepsZ_O, info[2] = GridPoints */
int airice_inice_table_info(const airice_inice_table *t, int64_t info[3]);
int airice_inice_table_copy_column(const airice_inice_table *t, int col, double *host_out);
int airice_inice_table_copy_positions(const airice_inice_table *t, float *host_x, float *host_z);
/* IceRayTracing::GetInterpolatedValue(xT, zT, rtParameter, AntNum) (IceRayTracing.cc:2727-2905), batched: bilinear in a
 * cell whose four nodes exist, inverse-distance weighting over the existing ones otherwise, -1000 outside the grid. */
int airice_inice_table_interp_device(airice_ctx *ctx, const airice_inice_table *t, int64_t n, const double *d_x,
                                     const double *d_z, int rt_parameter, double *d_out, void *stream);
int airice_inice_table_interp_host(airice_ctx *ctx, const airice_inice_table *t, int64_t n, const double *x, const double *z,
                                   int rt_parameter, double *out);

/* ---- kernel 5: ray-path emission = the RayPathinAirnIce.txt dump of the reference's CLI
 * (SingleRayAirIceRefraction.C:226-299 on the layer walk of :133-152; `./SingleRayAirIceRefraction 200 170 20000 3000`
 * writes 17206 points), batched: for every ray (launch angle theta in deg from the upward vertical, Tx height h in m)
 * the polyline at 1 m height steps through the air layers (last point of a layer clamped to the layer edge, next layer
 * entered 1e-5 m lower) and, for depth_m < 0, on at 1 m depth steps to the receiver depth (depth_m >= 0: the path ends on
 * the ice surface).
 * x, z: [n][max_points] row-major (horizontal distance from the Tx, height above sea level); entries past the end of a
 * path are NaN.  count[i] = points of ray i's full path (0: the ray does not exist; larger than max_points: truncated).
 * Call with max_points = 0 (x = z = NULL) to get the counts only. */
int airice_ray_path_device(airice_ctx *ctx, int64_t n, const double *d_theta, const double *d_h, double depth_m, double ice_m,
                           int64_t max_points, double *d_x, double *d_z, int32_t *d_count, void *stream);
int airice_ray_path_host(airice_ctx *ctx, int64_t n, const double *theta, const double *h, double depth_m, double ice_m,
                         int64_t max_points, double *x, double *z, int32_t *count);

/* ---- kernel 6: the old solve-per-cell table = MakeTable (MultiRayAirIceRefraction.cc:1618-1696, .h:193; grid globals
 * .h:38-54) and its lookup GetInterpolatedValue (MultiRayAirIceRefraction.cc:1700-1794, .h:195), device resident.
 * Nodes: Tx heights ice_m + 1, + step_h, ... 100000 m (last snapped) x straight-line angles start_th, + step_th, ...
 * stop_th (last snapped); the reference's defaults are (90.05, 179.95, 25, 0.01) = 3880 x 8990 = 34.9 M launch-angle
 * solves.  Node counts are the reference's int truncations of width/step + 1.  Nine f64 columns GridZValue[0..8] =
 * {h, THD, optical path ice, optical path air, launch angle, THD in air, T_S, T_P, incident angle}, -1000 where the solve
 * misses; node = ih * n_th + ith.  depth_m signed as MakeTable's AntennaDepth (negative = in ice). */
typedef struct airice_oldtable airice_oldtable;
#define AIRICE_OLDTABLE_COLS 9
int airice_oldtable_create(airice_ctx *ctx, double ice_m, double depth_m, double start_th, double stop_th, double step_h,
                           double step_th, airice_oldtable **out);
/* Wraps nine HOST columns of a grid made elsewhere (e.g. by the reference) for device lookups. */
int airice_oldtable_wrap_host(airice_ctx *ctx, double ice_m, double start_th, double stop_th, double step_h, double step_th,
                              const double *cols9, airice_oldtable **out);
void airice_oldtable_destroy(airice_oldtable *t);
/* info[0] = TotalStepsH_O, info[1] = TotalStepsTh_O, info[2] = GridPoints */
int airice_oldtable_info(const airice_oldtable *t, int64_t info[3]);
int airice_oldtable_copy_column(const airice_oldtable *t, int col, double *host_out);
/* GridPositionH [info[0]] and GridPositionTh [info[1]] */
int airice_oldtable_copy_positions(const airice_oldtable *t, double *host_h, double *host_th);
/* Batched GetInterpolatedValue(hR, thR, rtParameter): DEVICE arrays in and out, one thread per query. */
int airice_oldtable_interp_device(airice_ctx *ctx, const airice_oldtable *t, int64_t n, const double *d_h, const double *d_th,
                                  int rt_parameter, double *d_out, void *stream);
int airice_oldtable_interp_host(airice_ctx *ctx, const airice_oldtable *t, int64_t n, const double *h, const double *th,
                                int rt_parameter, double *out);

/* ---- peer memory: multi-GPU reassembly without a collective ("one gather per batch").
 * One process per GPU; the consumer rank allocates the result block with airice_peer_alloc and hands the 64-byte handle
 * to the producers (any transport: torch.distributed object broadcast, MPI, a file); a producer maps it with
 * airice_peer_open and passes pointers INTO it as the output columns of airice_solve_device / airice_lookup_device /
 * airice_table_build_device / ...: its kernel's stores travel over NVLink straight to their final place in the
 * consumer's HBM while the kernel computes -- the gather is fused into the compute kernel, no staging, no extra launch.
 * The reference has no counterpart (single process).  The consumer must order its reads after the producers' kernels
 * (a barrier on the producers' streams, e.g. an NCCL barrier / event + message).  airice_peer_copy is a plain
 * asynchronous device-to-device copy between any two such pointers for callers that want the block replicated. */
#define AIRICE_PEER_HANDLE_BYTES 64
int airice_peer_alloc(airice_ctx *ctx, size_t bytes, void **d_ptr, unsigned char handle[AIRICE_PEER_HANDLE_BYTES]);
int airice_peer_free(airice_ctx *ctx, void *d_ptr);
int airice_peer_open(airice_ctx *ctx, const unsigned char handle[AIRICE_PEER_HANDLE_BYTES], void **d_ptr);
int airice_peer_close(airice_ctx *ctx, void *d_ptr);
int airice_peer_copy(airice_ctx *ctx, void *d_dst, const void *d_src, size_t bytes, void *stream);

/* Host calls of up to 64 pairs or queries (airice_solve_host*, airice_lookup_host*: the reference's scalar functions are
 * batch-of-1 calls) do not issue copies at all: the kernel reads its inputs from and writes its results to a page-locked,
 * device-mapped block of the context -- one launch and one synchronisation per call.  AIRICE_NO_MAPPED=1 in the environment
 * keeps the chunked copy path for every size (test hook). */

/* ---- page-locked host memory.  The host entry points (airice_*_host*) copy straight from / to the caller's buffers; with
 * pageable memory every copy is staged through the driver and blocks the calling thread, with page-locked memory the
 * upload of chunk k+1, the kernel of chunk k and the download of chunk k-1 overlap and the call runs at the PCIe rate.
 * A C or C++ caller without the CUDA toolkit (CoREAS including MultiRayAirIceRefraction.cc) page-locks the buffers it
 * reuses with airice_host_register (cudaHostRegister) or takes them from airice_host_alloc (cudaHostAlloc).  The
 * reference has no counterpart: its arguments are scalars (MultiRayAirIceRefraction.h:186-189). */
int airice_host_register(void *p, size_t bytes);
int airice_host_unregister(void *p);
int airice_host_alloc(size_t bytes, void **p);
int airice_host_free(void *p);

/* ---- measurement helpers */
int airice_fp64_peak_tflops(airice_ctx *ctx, double *tflops); /* dependent-free DFMA probe, roofline denominator */
int airice_sync(airice_ctx *ctx);
/* Gives the device memory the context keeps for reuse back to the driver: the buffers of destroyed tables (up to a
 * third of the GPU's memory -- 64 antennas' reference-grid tables are 54 GB) and the solve kernel's per-stream
 * scratch.  Synchronises the device.  The reference has no counterpart (its tables live until the process ends). */
int airice_trim(airice_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
