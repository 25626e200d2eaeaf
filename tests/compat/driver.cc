// TEST: a caller written the way the reference's own drivers are (RunMultiRayCode.C:1-79): it includes
// "MultiRayAirIceRefraction.cc", defines the two vectors the header declares extern, fills the grid globals, builds
// one table per distinct antenna depth, then asks for one table lookup and one direct solve.  Output: one
// "key value" pair per line (%.17g), parsed by tests/test_gpu_compat.py.
#include "MultiRayAirIceRefraction.cc"

#include <cstdio>
#include <cstring>

std::vector<double> AntennaDepths;
std::vector<int> AntennaTableAlreadyMade;

int main(int argc, char **argv) {
  if (argc > 1) MultiRayAirIceRefraction::SetAtmosphereFile(argv[1]);
  double AntennaDepth = -200, IceLayerHeight = 3000, AirTxHeight = 5000, HorizontalDistance = 1000;
  // coarse grid so the oracle side of the test stays quick
  AngleStepSize = 0.5; LoopStartAngle = 92.0; LoopStopAngle = 180.0; HeightStepSize = 2000.0;

  AntennaDepths.push_back(AntennaDepth * 100);
  AntennaDepths.push_back(-150.0 * 100);
  AntennaDepths.push_back(AntennaDepth * 100);  // same depth as antenna 0: must reuse table 0
  for (size_t i = 0; i < AntennaDepths.size(); i++) {
    bool make = true;
    for (size_t j = 0; j < AntennaTableAlreadyMade.size(); j++)
      if (AntennaDepths[i] == AntennaDepths[AntennaTableAlreadyMade[j]]) make = false;
    if (make) {
      MultiRayAirIceRefraction::MakeRayTracingTable(AntennaDepths[i], IceLayerHeight * 100, (int)i);
      AntennaTableAlreadyMade.push_back((int)i);
    }
  }
  std::printf("tables %zu\n", AntennaTableAlreadyMade.size());
  std::printf("TotalHeightSteps %d\nTotalAngleSteps %d\n", TotalHeightSteps, TotalAngleSteps);

  double oi, oa, gi, ga, la, hx, ts, tp, ra;
  for (int ant = 0; ant < 3; ant++) {
    bool ok = MultiRayAirIceRefraction::GetHorizontalDistanceToIntersectionPoint_Table(
        AirTxHeight * 100, HorizontalDistance * 100, AntennaDepths[ant], IceLayerHeight * 100, ant, oi, oa, gi, ga, la, hx, ts, tp, ra);
    std::printf("table%d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", ant, (int)ok, oi, oa, gi, ga, la, hx, ts, tp, ra);
  }
  bool ok = MultiRayAirIceRefraction::GetHorizontalDistanceToIntersectionPoint(
      AirTxHeight * 100, HorizontalDistance * 100, AntennaDepth * 100, IceLayerHeight * 100, oi, oa, gi, ga, la, hx, ts, tp, ra);
  std::printf("direct %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", (int)ok, oi, oa, gi, ga, la, hx, ts, tp, ra);

  {
    // per-argument batch form: only the launch angle, the distance to the intersection point and the flag are asked for
    const double hs[2] = {AirTxHeight * 100, 20000.0 * 100}, ds[2] = {HorizontalDistance * 100, 3000.0 * 100};
    double bl[2] = {0, 0}, bx[2] = {0, 0};
    unsigned char bk[2] = {0, 0};
    const int rc = MultiRayAirIceRefraction::GetHorizontalDistanceToIntersectionPointBatch(
        2, hs, ds, AntennaDepth * 100, IceLayerHeight * 100, nullptr, nullptr, nullptr, nullptr, bl, bx, nullptr, nullptr, nullptr, bk);
    std::printf("direct_cols %d %d %.17g %.17g %d %.17g %.17g\n", rc, (int)bk[0], bl[0], bx[0], (int)bk[1], bl[1], bx[1]);
  }

  {
    // the table lookup in the per-argument batch form, antenna 1: launch angle + flag only
    const double hs[1] = {AirTxHeight * 100}, ds[1] = {HorizontalDistance * 100};
    double bl[1] = {0};
    unsigned char bk[1] = {0};
    const int rc = MultiRayAirIceRefraction::GetHorizontalDistanceToIntersectionPoint_TableBatch(
        1, hs, ds, AntennaDepths[1], IceLayerHeight * 100, 1, nullptr, nullptr, nullptr, nullptr, bl, nullptr, nullptr, nullptr, nullptr, bk);
    std::printf("table_cols %d %d %.17g\n", rc, (int)bk[0], bl[0]);
  }

  double dummy[20];
  bool inice = true;
  MultiRayAirIceRefraction::GetRayTracingSolutions(170, 20000, 3000, -200, dummy, inice);
  std::printf("forward");
  for (int i = 0; i < 18; i++) std::printf(" %.17g", dummy[i]);
  std::printf("\n");
  const double thR = 180 - (atan(HorizontalDistance / (AirTxHeight - IceLayerHeight - AntennaDepth)) * (180.0 / MultiRayAirIceRefraction::pi));
  MultiRayAirIceRefraction::Air2IceRayTracing(AirTxHeight, HorizontalDistance, IceLayerHeight, AntennaDepth, thR, dummy);
  std::printf("air2ice");
  for (int i = 0; i < 17; i++) std::printf(" %.17g", dummy[i]);
  std::printf("\n");

  // mutable ice model (MultiRayAirIceRefraction.h:72-74)
  MultiRayAirIceRefraction::A_ice = 1.775;
  ok = MultiRayAirIceRefraction::GetHorizontalDistanceToIntersectionPoint(
      AirTxHeight * 100, HorizontalDistance * 100, AntennaDepth * 100, IceLayerHeight * 100, oi, oa, gi, ga, la, hx, ts, tp, ra);
  std::printf("direct_A1775 %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", (int)ok, oi, oa, gi, ga, la, hx, ts, tp, ra);
  MultiRayAirIceRefraction::A_ice = 1.78;

  // medium accessors and Fresnel coefficients (MultiRayAirIceRefraction.h:90-119)
  const double zs[] = {0.0, 2999.99, 3000.0, 4000.0, 4000.01, 10000.0, 10000.5, 23141.75, 23141.76, 40000.0, 100000.0, -5000.0};
  for (double z : zs)
    std::printf("medium %.17g %.17g %.17g %.17g %.17g\n", z, MultiRayAirIceRefraction::GetB_air(z), MultiRayAirIceRefraction::GetC_air(z),
                MultiRayAirIceRefraction::Getnz_air(z), MultiRayAirIceRefraction::Getnz_ice(-z / 100));
  const double ths[] = {0.0, 0.3, 1.0, 1.5, 1.5707};
  for (double th : ths)
    std::printf("fresnel %.17g %.17g %.17g %.17g %.17g\n", th, MultiRayAirIceRefraction::Refl_S(th, 3000), MultiRayAirIceRefraction::Trans_S(th, 3000),
                MultiRayAirIceRefraction::Refl_P(th, 3000), MultiRayAirIceRefraction::Trans_P(th, 3000));

  // batch table build (all antennas in one pass, shared air walk): tables 2..4 must equal the per-antenna tables 0, 1
  // bit for bit, and a third depth comes along
  {
    std::vector<double> depths = {AntennaDepth * 100, -150.0 * 100, -75.0 * 100};
    const int rc = MultiRayAirIceRefraction::MakeRayTracingTables(depths, IceLayerHeight * 100);
    long differ = 0, cells = 0;
    std::vector<float> a, b;
    for (int k = 0; k < 2 && rc == 0; k++)
      for (int col = 0; col < 11; col++) {
        MultiRayAirIceRefraction::GetTableColumn(k, col, a);
        MultiRayAirIceRefraction::GetTableColumn(2 + k, col, b);
        cells = (long)a.size();
        if (a.size() != b.size()) { differ += 1000000; continue; }
        for (size_t q = 0; q < a.size(); q++) differ += std::memcmp(&a[q], &b[q], sizeof(float)) != 0;
      }
    std::vector<float> c;
    const int rc2 = MultiRayAirIceRefraction::GetTableColumn(4, 10, c);
    std::printf("multitables %d %ld %ld %d %zu\n", rc, differ, cells, rc2, c.size());
  }

  // old solve-per-cell table on the grid of tests/golden/old_table.npz
  MultiRayAirIceRefraction::GridStepSizeH_O = 4000.0;
  MultiRayAirIceRefraction::GridStepSizeTh_O = 1.5;
  MultiRayAirIceRefraction::MakeTable(IceLayerHeight * 100, AntennaDepth * 100);
  std::printf("old_dims %d %d %d\n", MultiRayAirIceRefraction::TotalStepsH_O, MultiRayAirIceRefraction::TotalStepsTh_O,
              (int)MultiRayAirIceRefraction::GridZValue[0].size());
  for (int c = 0; c < 9; c++) {
    std::printf("old_col%d", c);
    for (double v : MultiRayAirIceRefraction::GridZValue[c]) std::printf(" %.17g", v);
    std::printf("\n");
  }
  if (argc > 2) {
    FILE *f = std::fopen(argv[2], "r");
    double qh, qt;
    std::vector<double> qhs, qts;
    while (f && std::fscanf(f, "%lf %lf", &qh, &qt) == 2) {
      qhs.push_back(qh); qts.push_back(qt);
      std::printf("old_q");
      for (int p = 0; p < 9; p++) std::printf(" %.17g", MultiRayAirIceRefraction::GetInterpolatedValue(qh, qt, p));
      std::printf("\n");
    }
    if (f) std::fclose(f);
    // the batched form on the GPU-resident grid: same bits as the scalar host loop above
    long differ = 0;
    std::vector<double> o(qhs.size());
    for (int p = 0; p < 9; p++) {
      const int rc = MultiRayAirIceRefraction::GetInterpolatedValueBatch((long)qhs.size(), qhs.data(), qts.data(), p, o.data());
      if (rc != 0) differ += 1000000;
      for (size_t q = 0; q < qhs.size(); q++) {
        const double want = MultiRayAirIceRefraction::GetInterpolatedValue(qhs[q], qts[q], p);
        differ += std::memcmp(&want, &o[q], sizeof(double)) != 0;
      }
    }
    std::printf("old_batch_differ %ld %zu\n", differ, qhs.size());
  }
  // persistence: table 0 saved, loaded as a new table, same lookup bits
  if (argc > 3) {
    const int rcs = MultiRayAirIceRefraction::SaveRayTracingTable(0, argv[3]);
    const int idx = MultiRayAirIceRefraction::LoadRayTracingTable(argv[3]);
    double a[9], b[9];
    const bool ka = MultiRayAirIceRefraction::GetHorizontalDistanceToIntersectionPoint_Table(
        AirTxHeight * 100, HorizontalDistance * 100, AntennaDepth * 100, IceLayerHeight * 100, 0, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
    // idx lies beyond AntennaDepths, so the depth remap leaves it alone and the loaded table itself answers
    const bool kb = idx >= 0 && MultiRayAirIceRefraction::GetHorizontalDistanceToIntersectionPoint_Table(
        AirTxHeight * 100, HorizontalDistance * 100, AntennaDepth * 100, IceLayerHeight * 100, idx, b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8]);
    std::printf("persist %d %d %d %d %d\n", rcs, idx, (int)ka, (int)kb, (int)(std::memcmp(a, b, sizeof(a)) == 0));
  }
  // tables asked for under one ice model, used after the model changed back (and an antenna in air in between): the
  // deferred build must use the model of the time of the MakeRayTracingTable call, like the per-call build does
  {
    MultiRayAirIceRefraction::A_ice = 1.775;
    MultiRayAirIceRefraction::MakeRayTracingTable(-120.0 * 100, IceLayerHeight * 100, 0);   // in ice: collected
    MultiRayAirIceRefraction::MakeRayTracingTable(50.0 * 100, IceLayerHeight * 100, 0);     // in air: built at once
    MultiRayAirIceRefraction::MakeRayTracingTable(-60.0 * 100, IceLayerHeight * 100, 0);    // in ice: collected
    MultiRayAirIceRefraction::A_ice = 1.78;
    std::vector<float> c;
    double sum[3] = {0, 0, 0};
    int rcs[3];
    const int first = (int)AntennaTableAlreadyMade.size() + 3 + (argc > 3 ? 1 : 0);   // tables 0,1 + batch 2..4 (+ loaded)
    for (int k = 0; k < 3; k++) {
      rcs[k] = MultiRayAirIceRefraction::GetTableColumn(first + k, 2, c);               // optical path in ice
      for (float v : c) if (v == v) sum[k] += v;
    }
    std::printf("icemodel_tables %d %d %d %.17g %.17g %.17g\n", rcs[0], rcs[1], rcs[2], sum[0], sum[1], sum[2]);
  }
  return 0;
}
