// TEST INFRASTRUCTURE ONLY.
// Thin extern "C" shim around the UNMODIFIED reference translation unit
// /root/reference/MultiRayAirIceRefraction.cc (compiled in place via -I/root/reference,
// never copied into this repository).  Built by oracle/Makefile into
// oracle/_ref/libmultiray_ref.so against oracle/gsl_standin.  The reference source
// uses unqualified cout/isnan (MultiRayAirIceRefraction.cc:939,1505) and therefore only
// compiles after the three lines below, exactly as under ROOT/cling.
#include <iostream>
#include <cmath>
using namespace std;
#include "MultiRayAirIceRefraction.cc"

#include <cstring>

// The reference leaves these two for its caller to define (MultiRayAirIceRefraction.h:23-24).
std::vector<double> AntennaDepths;
std::vector<int> AntennaTableAlreadyMade;

namespace M = MultiRayAirIceRefraction;

extern "C" {

// The reference prints one line per solve (MultiRayAirIceRefraction.cc:1466); batch
// drivers mute std::cout instead of paying for the formatting + write.
void ref_quiet(int on) {
  if (on) std::cout.setstate(std::ios_base::failbit);
  else std::cout.clear();
}

int ref_make_atmosphere() { return M::MakeAtmosphere(); }

// out[0]=MaxLayers, out[1..5]=ATMLAY(cm), out[6..10]=B_air, out[11..15]=C_air,
// out[16..18]=A_ice,B_ice,C_ice, out[19]=pi, out[20]=spline(0)
void ref_constants(double *out) {
  out[0] = M::MaxLayers;
  for (int i = 0; i < 5; i++) { out[1 + i] = M::ATMLAY[i]; out[6 + i] = M::B_air[i]; out[11 + i] = M::C_air[i]; }
  out[16] = M::A_ice; out[17] = M::B_ice; out[18] = M::C_ice; out[19] = M::pi;
  out[20] = gsl_spline_eval(M::spline, 0, M::accelerator);
}

void ref_set_ice(double A, double B, double C) { M::A_ice = A; M::B_ice = B; M::C_ice = C; }

double ref_nz_air(double z) { return M::Getnz_air(z); }
double ref_nz_ice(double z) { return M::Getnz_ice(z); }

// GetRayTracingSolutions (forward tracer / one table cell): out[18]
void ref_forward(double theta, double h, double ice, double depth, int inice, double *out) {
  double dummy[20];
  bool InIce = inice != 0;
  M::GetRayTracingSolutions(theta, h, ice, depth, dummy, InIce);
  std::memcpy(out, dummy, 18 * sizeof(double));
}
void ref_forward_batch(long n, const double *theta, const double *h, double ice, double depth, int inice, double *out) {
  for (long i = 0; i < n; i++) ref_forward(theta[i], h[i], ice, depth, inice, out + 18 * i);
}

// Air2IceRayTracing (metres/degrees): out[17]
void ref_air2ice(double h, double d, double ice, double depth, double thR, double *out) {
  double dummy[20];
  for (int i = 0; i < 20; i++) dummy[i] = 0;
  M::Air2IceRayTracing(h, d, ice, depth, thR, dummy);
  std::memcpy(out, dummy, 17 * sizeof(double));
}

// GetHorizontalDistanceToIntersectionPoint (cm/rad API): out[9] in the argument order of
// MultiRayAirIceRefraction.h:170; returns the bool.
int ref_solve_cm(double h_cm, double d_cm, double depth_cm, double ice_cm, double *out) {
  bool ok = M::GetHorizontalDistanceToIntersectionPoint(h_cm, d_cm, depth_cm, ice_cm, out[0], out[1], out[2], out[3],
                                                        out[4], out[5], out[6], out[7], out[8]);
  return ok ? 1 : 0;
}
void ref_solve_cm_batch(long n, const double *h_cm, const double *d_cm, double depth_cm, double ice_cm, double *out,
                        unsigned char *ok) {
  for (long i = 0; i < n; i++) ok[i] = (unsigned char)ref_solve_cm(h_cm[i], d_cm[i], depth_cm, ice_cm, out + 9 * i);
}

// root function f(theta) = d - X(theta)  (MinimizeforLaunchAngle)
double ref_rootfn(double theta, double h, double ice, double depth_pos, double d) {
  M::MinforLAng_params p = {h, ice, depth_pos, d};
  return M::MinimizeforLaunchAngle(theta, &p);
}

// forward-table grid globals (MultiRayAirIceRefraction.cc:12-21); TotalAngleSteps is
// recomputed with the reference's own expression (line 15).
void ref_set_grid(double angle_step, double angle_start, double angle_stop, double height_step) {
  AngleStepSize = angle_step; LoopStartAngle = angle_start; LoopStopAngle = angle_stop; HeightStepSize = height_step;
  TotalAngleSteps = floor((LoopStopAngle - LoopStartAngle) / AngleStepSize) + 1;
}
void ref_clear_tables() { AllTableAllAntData.clear(); AntennaDepths.clear(); AntennaTableAlreadyMade.clear(); }
// Mirrors the caller protocol of RunMultiRayCode.C:31-52 (one table per distinct depth).
int ref_make_table(double depth_cm, double ice_cm) {
  int ant = (int)AntennaDepths.size();
  AntennaDepths.push_back(depth_cm);
  M::MakeRayTracingTable(depth_cm, ice_cm, ant);
  AntennaTableAlreadyMade.push_back(ant);
  return ant;
}
// info[0]=TotalHeightSteps info[1]=TotalAngleSteps info[2]=cells info[3]=ncols
void ref_table_info(int ant, long *info) {
  info[0] = TotalHeightSteps; info[1] = TotalAngleSteps;
  info[2] = (long)AllTableAllAntData[ant][0].size(); info[3] = (long)AllTableAllAntData[ant].size();
}
void ref_table_col(int ant, int col, float *out) {
  const std::vector<float> &v = AllTableAllAntData[ant][col];
  std::memcpy(out, v.data(), v.size() * sizeof(float));
}
// Overwrite a reference-held table column (lets a test feed OUR table to the reference lookup).
void ref_table_set_col(int ant, int col, const float *in) {
  std::vector<float> &v = AllTableAllAntData[ant][col];
  std::memcpy(v.data(), in, v.size() * sizeof(float));
}
int ref_lookup_cm(double h_cm, double d_cm, double depth_cm, double ice_cm, int ant, double *out) {
  for (int i = 0; i < 9; i++) out[i] = 0;
  bool ok = M::GetHorizontalDistanceToIntersectionPoint_Table(h_cm, d_cm, depth_cm, ice_cm, ant, out[0], out[1], out[2],
                                                              out[3], out[4], out[5], out[6], out[7], out[8]);
  return ok ? 1 : 0;
}
void ref_lookup_cm_batch(long n, const double *h_cm, const double *d_cm, double depth_cm, double ice_cm, int ant,
                         double *out, unsigned char *ok) {
  for (long i = 0; i < n; i++) ok[i] = (unsigned char)ref_lookup_cm(h_cm[i], d_cm[i], depth_cm, ice_cm, ant, out + 9 * i);
}
// index helpers of the lookup (integer results must be bit-exact)
void ref_find_rows(double h_m, int ant, int *idx /*4*/, double *cv /*2*/) {
  M::FindClosestAirTxHeight(h_m, idx[0], idx[1], cv[0], idx[2], idx[3], cv[1], ant);
}
void ref_find_thd(double d_m, int start, int end, int ant, int *idx /*2*/, double *cv) {
  M::FindClosestTHD(d_m, start, end, idx[0], idx[1], *cv, ant);
}

// old solve-per-cell table (MakeTable / GetInterpolatedValue).  The grid statics live in the
// reference header (MultiRayAirIceRefraction.h:42-48) and are visible in this TU.
void ref_old_grid(double start_th, double stop_th, double step_h, double step_th) {
  M::GridStartTh = start_th; M::GridStopTh = stop_th; M::GridStepSizeH_O = step_h; M::GridStepSizeTh_O = step_th;
  M::GridWidthTh = M::GridStopTh - M::GridStartTh;
}
void ref_old_set_stoph(double) {}
void ref_old_make_table(double ice_cm, double depth_cm) {
  for (int i = 0; i < 10; i++) M::GridZValue[i].clear();
  M::MakeTable(ice_cm, depth_cm);
}
void ref_old_info(long *info) { info[0] = M::TotalStepsH_O; info[1] = M::TotalStepsTh_O; info[2] = M::GridPoints; info[3] = (long)M::GridZValue[0].size(); }
void ref_old_col(int col, double *out) { std::memcpy(out, M::GridZValue[col].data(), M::GridZValue[col].size() * sizeof(double)); }
double ref_old_interp(double h, double th, int par) { return M::GetInterpolatedValue(h, th, par); }
}
