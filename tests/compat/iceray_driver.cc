// TEST: a caller of the in-ice entry point written like the reference's users (MakeMultiRayPlot.C style):
// includes "IceRayTracing.cc", calls IceRayTracing::IceRayTracing(0, z0, x1, z1), deletes the result.
#include "IceRayTracing.cc"

#include <cstdio>

int main(int argc, char **argv) {
  (void)argc; (void)argv;   // no atmosphere file: the in-ice context is created ice-only
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
    std::printf("att %.17g %.17g\n", Att[0], Att[1]);
    double f[2] = {1, 1};
    IceRayTracing::GetFocusingFactor(c[0], c[1], c[2], f);
    std::printf("focus %.17g %.17g\n", f[0], f[1]);
  }
  std::printf("attlen %.17g %.17g\n", IceRayTracing::GetIceAttenuationLength(-100.0, 0.3), IceRayTracing::GetIceTemperature(-1000.0));
  std::printf("attdirect %.17g\n", IceRayTracing::GetTotalAttenuationDirect(1.0, 0.3, -180.0, -5.0, 0.80023300831165));
  // the in-ice interpolation table on a small grid (tests/golden/inice_att.npz, table t1) and three lookups
  IceRayTracing::GridStepSizeX_O = 1.0; IceRayTracing::GridStepSizeZ_O = 1.0; IceRayTracing::GridWidthX = 12.0; IceRayTracing::GridWidthZ = 8.0;
  IceRayTracing::SetNumberOfAntennas(2);
  IceRayTracing::MakeTable(60.0, -30.0, -20.0, 1);
  std::vector<double> col;
  IceRayTracing::GetTableColumn(1, 0, col);
  std::printf("tablecol0");
  for (double v : col) std::printf(" %.17g", v);
  std::printf("\n");
  std::printf("interp %.17g %.17g %.17g\n", IceRayTracing::GetInterpolatedValue(57.3, -31.2, 0, 1), IceRayTracing::GetInterpolatedValue(60.0, -30.0, 4, 1),
              IceRayTracing::GetInterpolatedValue(10.0, -30.0, 0, 1));
  return 0;
}
