// TEST INFRASTRUCTURE ONLY.
// The UNMODIFIED reference command-line solver /root/reference/Air2IceRayTracing.C (on RayTracingFunctions.cc: Brent,
// tolerance 1e-9, <= 20 iterations, its own bracket rule) as a library call: its main() is compiled under another name and
// run with std::cout redirected into a string at 17 significant digits; the numbers it prints are parsed back.  Built into
// oracle/_ref/libcli_ref.so.  The file reads ./Atmosphere.dat on every call, like the CLI.
#include <cstdio>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

#define main air2ice_cli_main
#include "Air2IceRayTracing.C"
#undef main

static bool grab(const std::string &text, const char *key, double *v) {
  const size_t at = text.find(key);
  if (at == std::string::npos) return false;
  return std::sscanf(text.c_str() + at + std::strlen(key), "%lf", v) == 1;
}

extern "C" {
// out[9] = bracket lo, bracket hi, X_air, incident angle on ice, L, t_air [ns], X_ice, receive angle, t_ice [ns];
// returns the number of values found (9 when the run printed all of them)
int cliref_air2ice(double h, double d, double ice, double depth_positive, double *out) {
  char a1[40], a2[40], a3[40], a4[40], a0[] = "Air2IceRayTracing";
  std::snprintf(a1, sizeof a1, "%.17g", h); std::snprintf(a2, sizeof a2, "%.17g", d);
  std::snprintf(a3, sizeof a3, "%.17g", ice); std::snprintf(a4, sizeof a4, "%.17g", depth_positive);
  char *argv[] = {a0, a1, a2, a3, a4, nullptr};
  std::ostringstream cap;
  cap << std::setprecision(17);
  std::streambuf *old = std::cout.rdbuf(cap.rdbuf());
  const std::streamsize oldp = std::cout.precision(17);
  air2ice_cli_main(5, argv);
  std::cout.precision(oldp);
  std::cout.rdbuf(old);
  const std::string t = cap.str();
  int n = 0;
  n += grab(t, "Startangle ", out + 0);
  n += grab(t, ",Endangle ", out + 1);
  n += grab(t, "TotalHorizontalDistanceinAir ", out + 2);
  n += grab(t, "IncidentAngleonIce ", out + 3);
  n += grab(t, "LvalueAir for ", out + 4);
  n += grab(t, "PropagationTimeAir ", out + 5);
  n += grab(t, "TotalHorizontalDistanceinIce ", out + 6);
  n += grab(t, "IncidentAngleonAntenna ", out + 7);
  n += grab(t, "PropagationTimeIce ", out + 8);
  return n;
}
}
