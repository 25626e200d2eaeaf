// MultiRayAirIceRefraction.cc -- host-side mirror of the reference's hot-path API on top of the C ABI
// (include/airice_b200.h).  Included by callers exactly like the reference file (RunMultiRayCode.C:1), or linked
// from libMultiRayAirIceRefraction.so.  No arithmetic of the ray solve happens here: this file converts arguments,
// owns the table handles (the reference's AllTableAllAntData, MultiRayAirIceRefraction.cc:9) and maps output slots.
#include "MultiRayAirIceRefraction.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/stat.h>

#include "airice_b200.h"

double MaxAirTxHeight = 0, MinAirTxHeight = 0;
double AngleStepSize = 0.1, LoopStartAngle = 90.1, LoopStopAngle = 180.0;
int TotalAngleSteps = floor((LoopStopAngle - LoopStartAngle) / AngleStepSize) + 1;
double HeightStepSize = 10, LoopStartHeight = 0, LoopStopHeight = 0;
int TotalHeightSteps = 0;

namespace MultiRayAirIceRefraction {

double A_ice = 1.78, B_ice = -0.43, C_ice = 0.0132;
double GridStartTh = 90.05, GridStopTh = 179.95, GridStepSizeH_O = 25, GridStepSizeTh_O = 0.01, GridWidthH = 1000;
double GridWidthTh = GridStopTh - GridStartTh, GridStartH = 1000, GridStopH = 100000;
int GridPoints = 100, TotalStepsH_O = 100, TotalStepsTh_O = 100;
std::vector<double> GridPositionH, GridPositionTh;
std::vector<double> GridZValue[10];

namespace detail {
struct State {
  airice_ctx *ctx = nullptr;
  int device = 0;
  std::string atmosphere = "Atmosphere.dat";  // CWD-relative like the reference (MultiRayAirIceRefraction.cc:27,80)
  std::vector<airice_table *> tables;          // index = order of MakeRayTracingTable calls
  double ice_set[3] = {1.78, -0.43, 0.0132};
  long mtime = -1, size = -1;
  airice_oldtable *old_table = nullptr;        // device copy of GridZValue (last MakeTable)
  // MakeRayTracingTable calls for in-ice antennas that have not been built yet: the reference's callers build one
  // table per antenna in a loop (RunMultiRayCode.C) and only then start looking things up, so the calls are collected
  // and the first use builds them in ONE pass over the grid with the air walk shared (airice_table_create_multi;
  // 64 reference-grid tables: 7 ms instead of 44 ms one by one; the same bits).  AIRICE_EAGER_TABLES=1 builds per call.
  struct Pending { size_t slot; double depth_m, ice_m, h_step, th_start, th_step, th_stop; };
  std::vector<Pending> pending;
};
inline State &state() {
  static State s;
  return s;
}
inline bool flush_pending();
inline bool ensure_ctx() {
  State &s = state();
  // tables asked for under the ice model the context still has are built before the model changes
  if (s.ctx && !s.pending.empty() && (s.ice_set[0] != A_ice || s.ice_set[1] != B_ice || s.ice_set[2] != C_ice)) flush_pending();
  if (!s.ctx) {
    if (airice_create(s.atmosphere.c_str(), AIRICE_VARIANT_MULTIRAY, s.device, &s.ctx) != 0) {
      std::cerr << "MultiRayAirIceRefraction (B200): " << airice_last_error() << std::endl;
      s.ctx = nullptr;
      return false;
    }
  }
  if (s.ice_set[0] != A_ice || s.ice_set[1] != B_ice || s.ice_set[2] != C_ice) {
    airice_set_ice_model(s.ctx, A_ice, B_ice, C_ice);
    s.ice_set[0] = A_ice; s.ice_set[1] = B_ice; s.ice_set[2] = C_ice;
  }
  return true;
}
inline void report(const char *what) { std::cerr << "MultiRayAirIceRefraction (B200): " << what << ": " << airice_last_error() << std::endl; }
// build the collected tables: one shared-air pass per group of calls with the same surface height and grid
inline bool flush_pending() {
  State &s = state();
  bool ok = true;
  while (!s.pending.empty()) {
    const State::Pending f = s.pending.front();
    std::vector<State::Pending> group, rest;
    for (const State::Pending &q : s.pending) {
      const bool same = q.ice_m == f.ice_m && q.h_step == f.h_step && q.th_start == f.th_start && q.th_step == f.th_step &&
                        q.th_stop == f.th_stop;
      (same ? group : rest).push_back(q);
    }
    s.pending.swap(rest);
    std::vector<double> depths;
    for (const State::Pending &q : group) depths.push_back(q.depth_m);
    std::vector<airice_table *> ts(group.size(), nullptr);
    if (airice_table_create_multi(s.ctx, (int)group.size(), depths.data(), f.ice_m, 100000, f.h_step, f.th_start, f.th_step,
                                  f.th_stop, ts.data()) != 0) {
      report("MakeRayTracingTable (deferred build)");
      ok = false;
      continue;
    }
    for (size_t i = 0; i < group.size(); i++) s.tables[group[i].slot] = ts[i];
  }
  return ok;
}
}  // namespace detail

// ---- medium accessors (MultiRayAirIceRefraction.cc:150-263): the layer constants come from the context's parsed
// atmosphere (airice_get_medium); the layer rule is the reference's half-open [ATMLAY[k], ATMLAY[k+1]) in metres
namespace detail {
struct Medium { int nlayers; double lay[5], B[5], C[5]; bool ok; };
inline Medium medium() {
  Medium m;
  m.ok = false; m.nlayers = 0;
  double v[24];
  if (!ensure_ctx() || airice_get_medium(state().ctx, v) != 0) return m;
  m.nlayers = (int)v[0];
  for (int k = 0; k < 5; k++) { m.lay[k] = v[1 + k]; m.B[k] = v[6 + k]; m.C[k] = v[11 + k]; }
  m.ok = true;
  return m;
}
inline int air_layer(const Medium &m, double zabs) {
  int which = 0;
  for (int il = 0; il < m.nlayers - 1; il++)
    if (zabs < m.lay[il + 1] / 100 && zabs >= m.lay[il] / 100) { which = il; break; }
  if (zabs >= m.lay[m.nlayers - 1] / 100) which = m.nlayers - 1;
  return which;
}
}  // namespace detail

double GetB_ice(double z) { (void)z; return B_ice; }
double GetC_ice(double z) { (void)z; return C_ice; }
double Getnz_ice(double z) { z = fabs(z); return A_ice + GetB_ice(z) * exp(-GetC_ice(z) * z); }
double GetB_air(double z) {
  const detail::Medium m = detail::medium();
  return m.ok ? m.B[detail::air_layer(m, fabs(z))] : NAN;
}
double GetC_air(double z) {
  const detail::Medium m = detail::medium();
  return m.ok ? m.C[detail::air_layer(m, fabs(z))] : NAN;
}
double Getnz_air(double z) {
  const detail::Medium m = detail::medium();
  if (!m.ok) return NAN;
  const double zabs = fabs(z);
  const int k = detail::air_layer(m, zabs);
  return 1.0 + m.B[k] * exp(-m.C[k] * zabs);
}
namespace detail {
inline void fresnel_terms(double thetai, double IceLayerHeight, double &n1, double &n2, double &sqterm) {
  n1 = Getnz_air(IceLayerHeight);
  n2 = Getnz_ice(0);
  sqterm = sqrt(1 - pow((n1 / n2) * (sin(thetai)), 2));
}
}  // namespace detail
double Refl_S(double thetai, double IceLayerHeight) {
  double n1, n2, sq;
  detail::fresnel_terms(thetai, IceLayerHeight, n1, n2, sq);
  const double rS = (n1 * cos(thetai) - n2 * sq) / (n1 * cos(thetai) + n2 * sq);
  return std::isnan(rS) ? 1 : rS;
}
double Trans_S(double thetai, double IceLayerHeight) {
  double n1, n2, sq;
  detail::fresnel_terms(thetai, IceLayerHeight, n1, n2, sq);
  const double tS = 1 + ((n1 * cos(thetai) - n2 * sq) / (n1 * cos(thetai) + n2 * sq));
  return std::isnan(tS) ? 0 : tS;
}
double Refl_P(double thetai, double IceLayerHeight) {
  double n1, n2, sq;
  detail::fresnel_terms(thetai, IceLayerHeight, n1, n2, sq);
  const double rP = -(n1 * sq - n2 * cos(thetai)) / (n1 * sq + n2 * cos(thetai));
  return std::isnan(rP) ? 1 : rP;
}
double Trans_P(double thetai, double IceLayerHeight) {
  double n1, n2, sq;
  detail::fresnel_terms(thetai, IceLayerHeight, n1, n2, sq);
  const double tP = (1 - ((n1 * sq - n2 * cos(thetai)) / (n1 * sq + n2 * cos(thetai)))) * (n1 / n2);
  return std::isnan(tP) ? 0 : tP;
}

void SetDevice(int device) { detail::state().device = device; }
void SetAtmosphereFile(const std::string &path) { detail::state().atmosphere = path; }

// MultiRayAirIceRefraction.cc:920-942.  The reference re-parses the file (and rebuilds a 23k-point spline) on every
// call, i.e. once per table; here the parsed medium is kept and only refreshed when the file changed on disk.
int MakeAtmosphere() {
  detail::State &s = detail::state();
  struct stat st;
  const bool have = ::stat(s.atmosphere.c_str(), &st) == 0;
  const bool changed = have && ((long)st.st_mtime != s.mtime || (long)st.st_size != s.size);
  if (s.ctx && changed && s.tables.empty()) {
    airice_destroy(s.ctx);
    s.ctx = nullptr;
    s.ice_set[0] = 1.78; s.ice_set[1] = -0.43; s.ice_set[2] = 0.0132;
  }
  if (!detail::ensure_ctx()) return 1;
  if (have) { s.mtime = (long)st.st_mtime; s.size = (long)st.st_size; }
  std::cout << "Atmosphere has been generated " << std::endl;
  return 0;
}

// MultiRayAirIceRefraction.cc:2019-2158 (cm in).  The table lives on the GPU; index = order of calls.
int MakeRayTracingTable(double AntennaDepth, double IceLayerHeight, int AntennaNumber) {
  (void)AntennaNumber;
  if (MakeAtmosphere() != 0) return 1;
  detail::State &s = detail::state();
  AntennaDepth = AntennaDepth / 100;
  IceLayerHeight = IceLayerHeight / 100;
  const double AirTxHeight = 100000;
  TotalAngleSteps = floor((LoopStopAngle - LoopStartAngle) / AngleStepSize) + 1;
  LoopStartHeight = AirTxHeight;
  LoopStopHeight = (AntennaDepth < 0) ? IceLayerHeight : IceLayerHeight + AntennaDepth;
  TotalHeightSteps = floor((LoopStartHeight - LoopStopHeight) / HeightStepSize) + 1;
  static const bool eager = [] { const char *e = std::getenv("AIRICE_EAGER_TABLES"); return e && e[0] == '1'; }();
  if (AntennaDepth < 0 && !eager) {      // in-ice antenna: built with the others on first use (detail::flush_pending)
    s.pending.push_back({s.tables.size(), AntennaDepth, IceLayerHeight, HeightStepSize, LoopStartAngle, AngleStepSize, LoopStopAngle});
    s.tables.push_back(nullptr);
    return 0;
  }
  airice_table *t = nullptr;
  if (airice_table_create(s.ctx, AntennaDepth, IceLayerHeight, AirTxHeight, HeightStepSize, LoopStartAngle, AngleStepSize,
                          LoopStopAngle, &t) != 0) {
    detail::report("MakeRayTracingTable");
    return 1;
  }
  s.tables.push_back(t);
  return 0;
}

// Batch form of the reference's per-antenna loop `for (i...) MakeRayTracingTable(AntennaDepths[i], IceLayerHeight, i)`
// (RunMultiRayCode.C): all in-ice antennas in one pass over the grid, the air walk shared (airice_table_create_multi).
// Tables are appended in the order of `AntennaDepths`, exactly as the loop would; an antenna in air (depth >= 0, cm)
// falls back to its own MakeRayTracingTable call.
int MakeRayTracingTables(const std::vector<double> &AntennaDepths_cm, double IceLayerHeight) {
  if (MakeAtmosphere() != 0) return 1;
  detail::State &s = detail::state();
  bool all_in_ice = !AntennaDepths_cm.empty();
  for (double dcm : AntennaDepths_cm) all_in_ice = all_in_ice && (dcm < 0);
  if (!all_in_ice) {
    for (size_t i = 0; i < AntennaDepths_cm.size(); i++)
      if (MakeRayTracingTable(AntennaDepths_cm[i], IceLayerHeight, (int)i) != 0) return 1;
    return 0;
  }
  std::vector<double> depths_m;
  for (double dcm : AntennaDepths_cm) depths_m.push_back(dcm / 100);
  IceLayerHeight = IceLayerHeight / 100;
  const double AirTxHeight = 100000;
  TotalAngleSteps = floor((LoopStopAngle - LoopStartAngle) / AngleStepSize) + 1;
  LoopStartHeight = AirTxHeight;
  LoopStopHeight = IceLayerHeight;
  TotalHeightSteps = floor((LoopStartHeight - LoopStopHeight) / HeightStepSize) + 1;
  std::vector<airice_table *> ts(depths_m.size(), nullptr);
  if (airice_table_create_multi(s.ctx, (int)depths_m.size(), depths_m.data(), IceLayerHeight, AirTxHeight, HeightStepSize,
                                LoopStartAngle, AngleStepSize, LoopStopAngle, ts.data()) != 0) {
    detail::report("MakeRayTracingTables");
    return 1;
  }
  for (airice_table *t : ts) s.tables.push_back(t);
  return 0;
}

// Persistence (new): the reference rebuilds every table in every process.
int SaveRayTracingTable(int AntennaNumber, const std::string &path) {
  detail::State &s = detail::state();
  detail::flush_pending();
  if (AntennaNumber < 0 || AntennaNumber >= (int)s.tables.size() || !s.tables[AntennaNumber]) return 1;
  if (airice_table_save(s.tables[AntennaNumber], path.c_str()) != 0) { detail::report("SaveRayTracingTable"); return 1; }
  return 0;
}
int LoadRayTracingTable(const std::string &path) {
  if (MakeAtmosphere() != 0) return -1;
  detail::State &s = detail::state();
  airice_table *t = nullptr;
  if (airice_table_load(s.ctx, path.c_str(), &t) != 0) { detail::report("LoadRayTracingTable"); return -1; }
  s.tables.push_back(t);
  return (int)s.tables.size() - 1;
}

int GetTableColumn(int AntennaNumber, int col, std::vector<float> &out) {
  detail::State &s = detail::state();
  detail::flush_pending();
  if (AntennaNumber < 0 || AntennaNumber >= (int)s.tables.size() || !s.tables[AntennaNumber]) return 1;
  int64_t info[4];
  airice_table_info(s.tables[AntennaNumber], info);
  out.resize(info[2]);
  return airice_table_copy_column(s.tables[AntennaNumber], col, out.data());
}

int PinHostBuffer(void *p, size_t bytes) {
  int rc = airice_host_register(p, bytes);
  if (rc != 0) detail::report("PinHostBuffer");
  return rc;
}
int UnpinHostBuffer(void *p) {
  int rc = airice_host_unregister(p);
  if (rc != 0) detail::report("UnpinHostBuffer");
  return rc;
}

int GetHorizontalDistanceToIntersectionPointBatch(long n, const double *SrcHeightASL, const double *HorizontalDistanceToRx,
                                                  double RxDepthBelowIceBoundary, double IceLayerHeight, double *out,
                                                  unsigned char *ok) {
  if (!detail::ensure_ctx()) return 1;
  int rc = airice_solve_host(detail::state().ctx, n, SrcHeightASL, HorizontalDistanceToRx, nullptr, RxDepthBelowIceBoundary,
                             IceLayerHeight, AIRICE_UNITS_CM_RAD_C, out, ok);
  if (rc != 0) detail::report("GetHorizontalDistanceToIntersectionPointBatch");
  return rc;
}

int GetHorizontalDistanceToIntersectionPointBatch(long n, const double *SrcHeightASL, const double *HorizontalDistanceToRx,
                                                  double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                  double *opticalPathLengthInIce, double *opticalPathLengthInAir,
                                                  double *geometricalPathLengthInIce, double *geometricalPathLengthInAir,
                                                  double *launchAngle, double *horizontalDistanceToIntersectionPoint,
                                                  double *transmissionCoefficientS, double *transmissionCoefficientP,
                                                  double *RecievedAngleInIce, unsigned char *ok) {
  if (!detail::ensure_ctx()) return 1;
  double *const cols[AIRICE_SOLVE_COLS_CM_RAD] = {opticalPathLengthInIce, opticalPathLengthInAir, geometricalPathLengthInIce,
                                                  geometricalPathLengthInAir, launchAngle, horizontalDistanceToIntersectionPoint,
                                                  transmissionCoefficientS, transmissionCoefficientP, RecievedAngleInIce};
  int rc = airice_solve_host_columns(detail::state().ctx, n, SrcHeightASL, HorizontalDistanceToRx, nullptr,
                                     RxDepthBelowIceBoundary, IceLayerHeight, AIRICE_UNITS_CM_RAD_C, cols, ok);
  if (rc != 0) detail::report("GetHorizontalDistanceToIntersectionPointBatch");
  return rc;
}

// MultiRayAirIceRefraction.cc:945-989
bool GetHorizontalDistanceToIntersectionPoint(double SrcHeightASL, double HorizontalDistanceToRx,
                                              double RxDepthBelowIceBoundary, double IceLayerHeight,
                                              double &opticalPathLengthInIce, double &opticalPathLengthInAir,
                                              double &geometricalPathLengthInIce, double &geometricalPathLengthInAir,
                                              double &launchAngle, double &horizontalDistanceToIntersectionPoint,
                                              double &transmissionCoefficientS, double &transmissionCoefficientP,
                                              double &RecievedAngleInIce) {
  double o[9];
  unsigned char ok = 0;
  if (GetHorizontalDistanceToIntersectionPointBatch(1, &SrcHeightASL, &HorizontalDistanceToRx, RxDepthBelowIceBoundary,
                                                    IceLayerHeight, o, &ok) != 0)
    return false;
  opticalPathLengthInIce = o[0]; opticalPathLengthInAir = o[1];
  geometricalPathLengthInIce = o[2]; geometricalPathLengthInAir = o[3];
  launchAngle = o[4]; horizontalDistanceToIntersectionPoint = o[5];
  transmissionCoefficientS = o[6]; transmissionCoefficientP = o[7]; RecievedAngleInIce = o[8];
  return ok != 0;
}

namespace detail {
// antenna -> table remap by depth equality (MultiRayAirIceRefraction.cc:1348-1352), pending builds flushed; nullptr = no table
inline airice_table *table_of_antenna(int AntennaNumber) {
  State &s = state();
  for (size_t j = 0; j < AntennaTableAlreadyMade.size(); j++) {
    if (AntennaNumber < (int)AntennaDepths.size() && AntennaTableAlreadyMade[j] < (int)AntennaDepths.size() &&
        AntennaDepths[AntennaNumber] == AntennaDepths[AntennaTableAlreadyMade[j]])
      AntennaNumber = (int)j;
  }
  if (!s.pending.empty()) flush_pending();
  if (AntennaNumber < 0 || AntennaNumber >= (int)s.tables.size() || !s.tables[AntennaNumber]) {
    std::cerr << "MultiRayAirIceRefraction (B200): no table for antenna " << AntennaNumber << std::endl;
    return nullptr;
  }
  return s.tables[AntennaNumber];
}
}  // namespace detail

int GetHorizontalDistanceToIntersectionPoint_TableBatch(long n, const double *SrcHeightASL,
                                                        const double *HorizontalDistanceToRx,
                                                        double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                        int AntennaNumber, double *out, unsigned char *ok) {
  (void)RxDepthBelowIceBoundary; (void)IceLayerHeight;
  if (!detail::ensure_ctx()) return 1;
  airice_table *t = detail::table_of_antenna(AntennaNumber);
  if (!t) return 1;
  int rc = airice_lookup_host(detail::state().ctx, t, n, SrcHeightASL, HorizontalDistanceToRx, out, ok);
  if (rc != 0) detail::report("GetHorizontalDistanceToIntersectionPoint_TableBatch");
  return rc;
}

int GetHorizontalDistanceToIntersectionPoint_TableBatch(long n, const double *SrcHeightASL,
                                                        const double *HorizontalDistanceToRx,
                                                        double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                        int AntennaNumber, double *opticalPathLengthInIce,
                                                        double *opticalPathLengthInAir, double *geometricalPathLengthInIce,
                                                        double *geometricalPathLengthInAir, double *launchAngle,
                                                        double *horizontalDistanceToIntersectionPoint,
                                                        double *transmissionCoefficientS, double *transmissionCoefficientP,
                                                        double *RecievedAngleInIce, unsigned char *ok) {
  (void)RxDepthBelowIceBoundary; (void)IceLayerHeight;
  if (!detail::ensure_ctx()) return 1;
  airice_table *t = detail::table_of_antenna(AntennaNumber);
  if (!t) return 1;
  double *const cols[AIRICE_LOOKUP_COLS] = {opticalPathLengthInIce, opticalPathLengthInAir, geometricalPathLengthInIce,
                                             geometricalPathLengthInAir, launchAngle, horizontalDistanceToIntersectionPoint,
                                             transmissionCoefficientS, transmissionCoefficientP, RecievedAngleInIce};
  int rc = airice_lookup_host_columns(detail::state().ctx, t, n, SrcHeightASL, HorizontalDistanceToRx, cols, ok);
  if (rc != 0) detail::report("GetHorizontalDistanceToIntersectionPoint_TableBatch");
  return rc;
}

// MultiRayAirIceRefraction.cc:1305-1462
bool GetHorizontalDistanceToIntersectionPoint_Table(double SrcHeightASL, double HorizontalDistanceToRx,
                                                    double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                    int AntennaNumber, double &opticalPathLengthInIce,
                                                    double &opticalPathLengthInAir, double &geometricalPathLengthInIce,
                                                    double &geometricalPathLengthInAir, double &launchAngle,
                                                    double &horizontalDistanceToIntersectionPoint,
                                                    double &transmissionCoefficientS, double &transmissionCoefficientP,
                                                    double &RecievedAngleInIce) {
  double o[9];
  unsigned char ok = 0;
  if (GetHorizontalDistanceToIntersectionPoint_TableBatch(1, &SrcHeightASL, &HorizontalDistanceToRx, RxDepthBelowIceBoundary,
                                                          IceLayerHeight, AntennaNumber, o, &ok) != 0)
    return false;
  opticalPathLengthInIce = o[0]; opticalPathLengthInAir = o[1];
  geometricalPathLengthInIce = o[2]; geometricalPathLengthInAir = o[3];
  launchAngle = o[4]; horizontalDistanceToIntersectionPoint = o[5];
  transmissionCoefficientS = o[6]; transmissionCoefficientP = o[7]; RecievedAngleInIce = o[8];
  return ok != 0;
}

// MultiRayAirIceRefraction.cc:1464-1616 (metres / degrees; dummy[0..16])
void Air2IceRayTracing(double AirTxHeight, double HorizontalDistance, double IceLayerHeight, double AntennaDepth,
                       double StraightAngle, double dummy[20]) {
  std::cout << "main function parameters are " << AirTxHeight << " " << HorizontalDistance << " " << IceLayerHeight << " "
            << AntennaDepth << std::endl;  // the reference prints this per call (MultiRayAirIceRefraction.cc:1466)
  for (int i = 0; i < 20; i++) dummy[i] = 0;
  if (!detail::ensure_ctx()) return;
  double o[AIRICE_SOLVE_COLS];
  unsigned char ok = 0;
  if (airice_solve_host(detail::state().ctx, 1, &AirTxHeight, &HorizontalDistance, &StraightAngle, AntennaDepth,
                        IceLayerHeight, AIRICE_UNITS_M_DEG_C, o, &ok) != 0) {
    detail::report("Air2IceRayTracing");
    return;
  }
  const double t_air = o[3], t_ice = o[4], t = t_ice + t_air;
  dummy[0] = AirTxHeight; dummy[1] = o[0]; dummy[2] = o[1]; dummy[3] = o[2];
  dummy[4] = t * spedc; dummy[5] = t_ice * spedc; dummy[6] = t_air * spedc;
  dummy[7] = t; dummy[8] = t_ice; dummy[9] = t_air;
  dummy[10] = o[5]; dummy[11] = o[6]; dummy[12] = o[7]; dummy[13] = o[8];
  dummy[14] = o[9]; dummy[15] = o[10]; dummy[16] = o[11];
}

// MultiRayAirIceRefraction.cc:1796-2017 (metres / degrees; dummy[0..17])
void GetRayTracingSolutions(double RayLaunchAngleInAir, double AirTxHeight, double IceLayerHeight, double AntennaDepth,
                            double dummy[20], bool &InIce) {
  for (int i = 0; i < 18; i++) dummy[i] = 0;
  if (!detail::ensure_ctx()) return;
  double o[AIRICE_TABLE_COLS64];
  // InIce decides whether the ice leg is traced; the C ABI encodes it in the sign of the depth
  const double depth = InIce ? (AntennaDepth < 0 ? AntennaDepth : -AntennaDepth) : 0.0;
  if (airice_forward_host(detail::state().ctx, 1, &RayLaunchAngleInAir, &AirTxHeight, depth, IceLayerHeight, o) != 0) {
    detail::report("GetRayTracingSolutions");
    return;
  }
  for (int i = 0; i < AIRICE_TABLE_COLS64; i++) dummy[1 + i] = o[i];
}

// MultiRayAirIceRefraction.cc:1618-1696: the old solve-per-cell table.  The grid is generated, solved and kept on the GPU
// (airice_oldtable_create: one launch-angle solve per node in batched launches); the reference's public GridZValue /
// GridPosition vectors are filled from it so that code reading them keeps working.  cm in; nine double columns, -1000 where
// the solve misses.
void MakeTable(double IceLayerHeight, double AntennaDepth) {
  IceLayerHeight = IceLayerHeight / 100;
  AntennaDepth = AntennaDepth / 100;
  std::cout << "making the table now " << AntennaDepth << " " << IceLayerHeight << std::endl;
  if (MakeAtmosphere() != 0) return;
  detail::State &st = detail::state();
  if (st.old_table) { airice_oldtable_destroy(st.old_table); st.old_table = nullptr; }
  GridStartH = IceLayerHeight + 1;
  GridStopH = 100000;
  GridWidthH = GridStopH - GridStartH;
  GridWidthTh = GridStopTh - GridStartTh;
  for (int c = 0; c < 10; c++) GridZValue[c].clear();
  if (airice_oldtable_create(st.ctx, IceLayerHeight, AntennaDepth, GridStartTh, GridStopTh, GridStepSizeH_O, GridStepSizeTh_O,
                             &st.old_table) != 0) {
    detail::report("MakeTable");
    st.old_table = nullptr;
    return;
  }
  int64_t info[3];
  airice_oldtable_info(st.old_table, info);
  TotalStepsH_O = (int)info[0];
  TotalStepsTh_O = (int)info[1];
  GridPoints = (int)info[2];
  GridPositionH.resize(TotalStepsH_O);
  GridPositionTh.resize(TotalStepsTh_O);
  airice_oldtable_copy_positions(st.old_table, GridPositionH.data(), GridPositionTh.data());
  for (int c = 0; c < AIRICE_OLDTABLE_COLS; c++) {
    GridZValue[c].resize(GridPoints);
    if (airice_oldtable_copy_column(st.old_table, c, GridZValue[c].data()) != 0) detail::report("MakeTable");
  }
}

// Batched GetInterpolatedValue on the device-resident grid of the last MakeTable (new, not in the reference): one thread
// per query, same arithmetic as the scalar function below.  Returns 0 on success.
int GetInterpolatedValueBatch(long n, const double *hR, const double *thR, int rtParameter, double *out) {
  detail::State &st = detail::state();
  if (!st.old_table) { std::cerr << "GetInterpolatedValueBatch: MakeTable has not been called" << std::endl; return -1; }
  const int rc = airice_oldtable_interp_host(st.ctx, st.old_table, n, hR, thR, rtParameter, out);
  if (rc != 0) detail::report("GetInterpolatedValueBatch");
  return rc;
}

// MultiRayAirIceRefraction.cc:1700-1794: inverse-distance weighting over the 2x2 nodes below/left of the rounded bin,
// reproduced as written (the running value is overwritten per node; an exact hit short-circuits).
double GetInterpolatedValue(double hR, double thR, int rtParameter) {
  double sum1 = 0, sum2 = 0, NewZValue = -1000;
  double minHbin = round((hR - GridStartH) / GridStepSizeH_O);
  double minThbin = round((thR - GridStartTh) / GridStepSizeTh_O);
  if (minHbin <= 1) minHbin = 1;
  if (minThbin <= 1) minThbin = 1;
  if (minHbin + 1 > TotalStepsH_O) minHbin = TotalStepsH_O - 2;
  if (minThbin + 1 > TotalStepsTh_O) minThbin = TotalStepsTh_O - 2;
  const int startbinH = minHbin - 1, endbinH = minHbin + 1, startbinTh = minThbin - 1, endbinTh = minThbin + 1;
  const std::vector<double> &Z = GridZValue[rtParameter];
  for (int ixn = startbinH; ixn < endbinH; ixn++) {
    for (int izn = startbinTh; izn < endbinTh; izn++) {
      const int ich = ixn * TotalStepsTh_O + izn;
      if (ich >= 0 && ich < GridPoints && ixn < TotalStepsH_O && izn < TotalStepsTh_O && ixn >= 0 && izn >= 0) {
        const double dist = fabs((hR - GridPositionH[ixn]) * (hR - GridPositionH[ixn]) +
                                 (thR - GridPositionTh[izn]) * (thR - GridPositionTh[izn]));
        if (Z[ich] != -1000) {
          sum1 += (1.0 / dist) * Z[ich];
          sum2 += (1.0 / dist);
          NewZValue = sum1 / sum2;
        } else {
          NewZValue = -1000;
        }
        if (dist == 0) {
          NewZValue = (Z[ich] != -1000) ? Z[ich] : -1000;
          izn = minThbin + 3;
          ixn = minHbin + 3;
        }
      }
    }
  }
  return NewZValue;
}

}  // namespace MultiRayAirIceRefraction
