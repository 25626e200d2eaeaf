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
double iceref_zmax(double A, double L) { return IceRayTracing::GetZmax(A, L); }
double iceref_fraa(double L, double z0, double x1, double z1) {
  IceRayTracing::fDanfRa_params p = {IceRayTracing::A_ice, z0, x1, z1};
  return IceRayTracing::fRaa(L, &p);
}
double iceref_nz(double z) { return IceRayTracing::Getnz(z); }
}
