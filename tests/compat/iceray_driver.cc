// TEST: a caller of the in-ice entry point written like the reference's users (MakeMultiRayPlot.C style):
// includes "IceRayTracing.cc", calls IceRayTracing::IceRayTracing(0, z0, x1, z1), deletes the result.
#include "IceRayTracing.cc"

#include <cstdio>

int main(int argc, char **argv) {
  if (argc > 1) IceRayTracing::SetAtmosphereFile(argv[1]);
  const double cases[3][3] = {{-180, 100, -5}, {-1000, 2000, -200}, {-200, 1500, -150}};
  for (auto &c : cases) {
    double *r = IceRayTracing::IceRayTracing(0, c[0], c[1], c[2]);
    std::printf("case");
    for (int i = 0; i < 29; i++) std::printf(" %.17g", r[i]);
    std::printf("\n");
    delete[] r;
  }
  // the two-ray selection of the same three pairs, with the reference's own argument list
  for (auto &c : cases) {
    double T[2], P[2], La[2], Ra[2], Inc[2], Att[2];
    int Ig[2];
    IceRayTracing::GetRayTracingSolutions(c[2], c[1], c[0], T, P, La, Ra, Ig, Inc, 1.0, 0.3, Att);
    std::printf("rays %d %d", Ig[0], Ig[1]);
    for (int k = 0; k < 2; k++) std::printf(" %.17g %.17g %.17g %.17g %.17g", T[k], P[k], La[k], Ra[k], Inc[k]);
    std::printf("\n");
  }
  return 0;
}
