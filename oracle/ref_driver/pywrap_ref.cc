// TEST INFRASTRUCTURE ONLY.
// Compiles the UNMODIFIED /root/reference/pythonwrapper/TraceIceToAir.C (which itself
// #includes AirIceRayTracing.cc) into oracle/_ref/pywrap/libAirIceRayTracing.so, i.e. the same
// shared object the reference's build comment (TraceIceToAir.C:3) produces, plus two batch
// helpers used by the parity tests.
#include <iostream>
#include <cmath>
using namespace std;
#include "TraceIceToAir.C"

extern "C" {
void pyref_quiet(int on) {
  if (on) std::cout.setstate(std::ios_base::failbit);
  else std::cout.clear();
}
int pyref_make_atmosphere(const char *path) { return AirIceRayTracing::MakeAtmosphere(path); }
// AirIceRayTracing::Air2IceRayTracing (metres/degrees) without re-parsing the file: out[15]
void pyref_air2ice(double h, double d, double ice, double depth, double thR, double *out) {
  double dummy[20];
  for (int i = 0; i < 20; i++) dummy[i] = 0;
  AirIceRayTracing::Air2IceRayTracing(h, d, ice, depth, thR, dummy);
  for (int i = 0; i < 15; i++) out[i] = dummy[i];
}
// GetRayTracingSolution: out[8] in the header's argument order, returns the bool.
int pyref_solution(double h, double d, double depth, double ice, double *out) {
  bool ok = AirIceRayTracing::GetRayTracingSolution(h, d, depth, ice, out[0], out[1], out[2], out[3], out[4], out[5],
                                                    out[6], out[7]);
  return ok ? 1 : 0;
}
// The python-wrapper copy's constant-refractive-index air option, switched on the way the commented-out lines of
// TraceIceToAir.C:27-29 do it (A_const = n_air(ice surface), flag, A_air = A_const); on = 0 restores the defaults.
// root_fn[3] receives MinimizeforLaunchAngle at the lower bracket end the option fixes (startanglelim = 90,
// AirIceRayTracing.cc:986-988) and at two interior angles, for the caller's (h, d, ice, depth > 0).
void pyref_set_constant_air(int on, double ice) {
  namespace P = AirIceRayTracing;
  if (on) {
    P::UseConstantRefractiveIndex = false;
    P::A_air = 1.00;
    P::A_const = P::Getnz_air(ice);
    P::UseConstantRefractiveIndex = true;
    P::A_air = P::A_const;
  } else {
    P::UseConstantRefractiveIndex = false;
    P::A_air = 1.00;
    P::A_const = 1.00;
  }
}
void pyref_rootfn3(double h, double d, double ice, double depth, const double *theta, double *root_fn) {
  struct AirIceRayTracing::MinforLAng_params p = {h, ice, depth, d};
  for (int i = 0; i < 3; i++) root_fn[i] = AirIceRayTracing::MinimizeforLaunchAngle(theta[i], &p);
}
void pyref_constants(double *out) {
  namespace P = AirIceRayTracing;
  out[0] = P::MaxLayers;
  for (int i = 0; i < 5; i++) { out[1 + i] = P::ATMLAY[i]; out[6 + i] = P::B_air[i]; out[11 + i] = P::C_air[i]; }
  out[16] = P::A_ice; out[17] = -0.43; out[18] = 0.0132; out[19] = P::pi;
}
}
