// TEST INFRASTRUCTURE ONLY.
// extern "C" shim around the UNMODIFIED /root/reference/IceRayTracing.cc (in-ice direct / reflected / refracted
// launch-angle solver), built into oracle/_ref/libiceray_ref.so.  As shipped the pair does not compile on its own:
// IceRayTracing.hh:186 declares IceRayTracing(double,double,double,double,bool) while IceRayTracing.cc:1745 defines
// the 4-argument form.  The missing declaration is supplied HERE, in this translation unit, between the two
// includes; the reference files themselves are untouched.
#include "IceRayTracing.hh"
namespace IceRayTracing { double *IceRayTracing(double x0, double z0, double x1, double z1); }
#include "IceRayTracing.cc"

static long g_fevals[4];

extern "C" {
// IceRayTracing::IceRayTracing(0, z0, x1, z1) -> out[29] (IceRayTracing.cc:1745-1919).  Slots 12..17 are only written
// by the reference when the corresponding branch exists (IceRayTracing.cc:1876-1889); absent ones are zeroed here.
void iceref_solve(double z0, double x1, double z1, double *out) {
  double *r = IceRayTracing::IceRayTracing(0.0, z0, x1, z1);
  for (int i = 0; i < 29; i++) out[i] = r[i];
  if (out[9] == -1000) { out[12] = 0; out[13] = 0; }
  if (out[10] == -1000) { out[14] = 0; out[15] = 0; }
  if (out[11] == -1000) { out[16] = 0; out[17] = 0; }
  delete[] r;
}
void iceref_solve_batch(long n, const double *z0, const double *x1, const double *z1, double *out) {
  for (long i = 0; i < n; i++) iceref_solve(z0[i], x1[i], z1[i], out + 29 * i);
}
void iceref_direct(double z0, double x1, double z1, double *out6) {
  double *r = IceRayTracing::GetDirectRayPar(z0, x1, z1);
  for (int i = 0; i < 6; i++) out6[i] = r[i];
  delete[] r;
}
void iceref_reflected(double z0, double x1, double z1, double *out11) {
  double *r = IceRayTracing::GetReflectedRayPar(z0, x1, z1);
  for (int i = 0; i < 11; i++) out11[i] = r[i];
  delete[] r;
}
// IceRayTracing::GetRayTracingSolutions (IceRayTracing.cc:2907-3210).  Its attenuation outputs go through the stand-in's
// quadrature and are not compared; out10 = TimeRay, PathRay, LaunchAngle, RecieveAngle, IncidenceAngleInIce (2 each).
void iceref_two_rays_batch(long n, const double *rx, const double *dist, const double *tx, double *out10, int *ignore2) {
  for (long i = 0; i < n; i++) {
    double T[2], P[2], La[2], Ra[2], Inc[2], Att[2];
    int Ig[2];
    IceRayTracing::GetRayTracingSolutions(rx[i], dist[i], tx[i], T, P, La, Ra, Ig, Inc, 1.0, 0.3, Att);
    double *o = out10 + 10 * i;
    o[0] = T[0]; o[1] = T[1]; o[2] = P[0]; o[3] = P[1]; o[4] = La[0]; o[5] = La[1]; o[6] = Ra[0]; o[7] = Ra[1];
    o[8] = Inc[0]; o[9] = Inc[1];
    ignore2[2 * i] = Ig[0]; ignore2[2 * i + 1] = Ig[1];
  }
}
// The same with the attenuation outputs (A0, frequency [GHz] -> AttRay = 1 - integral; IceRayTracing.cc:2976-2987,
// 179-219 through the stand-in's gsl_integration_qags, itself checked against QUADPACK in tests/test_oracle.py)
void iceref_two_rays_att_batch(long n, const double *rx, const double *dist, const double *tx, double A0, double frequency,
                               double *out10, double *att2, int *ignore2) {
  for (long i = 0; i < n; i++) {
    double T[2], P[2], La[2], Ra[2], Inc[2], Att[2];
    int Ig[2];
    IceRayTracing::GetRayTracingSolutions(rx[i], dist[i], tx[i], T, P, La, Ra, Ig, Inc, A0, frequency, Att);
    double *o = out10 + 10 * i;
    o[0] = T[0]; o[1] = T[1]; o[2] = P[0]; o[3] = P[1]; o[4] = La[0]; o[5] = La[1]; o[6] = Ra[0]; o[7] = Ra[1];
    o[8] = Inc[0]; o[9] = Inc[1];
    att2[2 * i] = Att[0]; att2[2 * i + 1] = Att[1];
    ignore2[2 * i] = Ig[0]; ignore2[2 * i + 1] = Ig[1];
  }
}
// GetTotalAttenuationDirect / Reflected / Refracted (IceRayTracing.cc:203-219); kind 0, 1, 2
double iceref_attenuation(int kind, double A0, double frequency, double z0, double z1, double zmax, double L) {
  if (kind == 0) return IceRayTracing::GetTotalAttenuationDirect(A0, frequency, z0, z1, L);
  if (kind == 1) return IceRayTracing::GetTotalAttenuationReflected(A0, frequency, z0, z1, L);
  return IceRayTracing::GetTotalAttenuationRefracted(A0, frequency, z0, z1, zmax, L);
}
double iceref_attenuation_length(double z, double frequency) { return IceRayTracing::GetIceAttenuationLength(z, frequency); }
// GetFocusingFactor(zT, xR, zR, focusing[2]) (IceRayTracing.cc:3218-3293); the caller's initial {1, 1} as in MakeTable
void iceref_focusing_batch(long n, const double *zT, const double *xR, const double *zR, double *out2) {
  for (long i = 0; i < n; i++) {
    double f[2] = {1, 1};
    IceRayTracing::GetFocusingFactor(zT[i], xR[i], zR[i], f);
    out2[2 * i] = f[0]; out2[2 * i + 1] = f[1];
  }
}
// MakeTable / GetInterpolatedValue (IceRayTracing.cc:2614-2905) on a grid of the caller's choosing (the globals are
// file-static in IceRayTracing.hh, reachable from this translation unit)
void iceref_set_grid(double step_x, double step_z, double width_x, double width_z) {
  IceRayTracing::GridStepSizeX_O = step_x; IceRayTracing::GridStepSizeZ_O = step_z;
  IceRayTracing::GridWidthX = width_x; IceRayTracing::GridWidthZ = width_z;
}
void iceref_make_table(int n_ant, double hit_distance, double shower_depth, double zR, int ant) {
  if ((int)IceRayTracing::GridZValueb.size() != n_ant) IceRayTracing::SetNumberOfAntennas(n_ant);
  IceRayTracing::GridZValueb[ant].clear();
  IceRayTracing::MakeTable(hit_distance, shower_depth, zR, ant);
}
void iceref_table_info(int ant, long *info3) {
  info3[0] = (long)IceRayTracing::GridPositionXb[ant].size();
  info3[1] = (long)IceRayTracing::GridPositionZb[ant].size();
  info3[2] = (long)IceRayTracing::GridZValueb[ant][0].size();
}
void iceref_table_column(int ant, int col, double *out) {
  const std::vector<double> &v = IceRayTracing::GridZValueb[ant][col];
  for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
}
void iceref_table_positions(int ant, float *x, float *z) {
  for (size_t i = 0; i < IceRayTracing::GridPositionXb[ant].size(); i++) x[i] = IceRayTracing::GridPositionXb[ant][i];
  for (size_t i = 0; i < IceRayTracing::GridPositionZb[ant].size(); i++) z[i] = IceRayTracing::GridPositionZb[ant][i];
}
double iceref_interp(double xT, double zT, int par, int ant) { return IceRayTracing::GetInterpolatedValue(xT, zT, par, ant); }

double iceref_zmax(double A, double L) { return IceRayTracing::GetZmax(A, L); }
double iceref_fraa(double L, double z0, double x1, double z1) {
  IceRayTracing::fDanfRa_params p = {IceRayTracing::A_ice, z0, x1, z1};
  return IceRayTracing::fRaa(L, &p);
}
double iceref_nz(double z) { return IceRayTracing::Getnz(z); }
}
