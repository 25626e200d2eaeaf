// airice_host.hpp -- host-side pieces of the B200 air->ice solver: medium model loading, per-launch plans,
// grid geometry, and the launch wrappers the C ABI calls.  No torch, no GSL.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "airice_core.cuh"

namespace airice {

// Atmosphere.dat -> AirIceMedium.  Follows readATMpar / readnhFromFile / MakeAtmosphere /
// FillInAirRefractiveIndex (MultiRayAirIceRefraction.cc:24-147, 193-213, 920-942); `n0_out` receives the natural
// cubic spline value at h=0 (the only use the reference makes of its 23k-point GSL spline, M.cc:203).
// Returns 0 or a negative error code; `err` gets a message.
int load_medium(const char* path, int variant, AirIceMedium* out, double* n0_out, int* npoints_out, std::string* err);
void make_clamp_table(const AirIceMedium& m, double* tab /* [AIRICE_CLAMP_N][2] */);

// Layer index of a height as GetB_air/GetC_air see it (M.cc:216-256).
int layer_of(const AirIceMedium& m, double z);
double n_air(const AirIceMedium& m, double z);  // Getnz_air, host libm
double n_ice(const AirIceMedium& m, double z);  // Getnz_ice, host libm

// Ray-independent numbers of the layer walk for one (ice height, signed receiver depth) pair.
// depth_signed < 0: receiver in ice; >= 0: receiver in air, folded into the surface height (M.cc:1472-1479).
void make_plan(const AirIceMedium& m, double ice_h, double depth_signed, AirIcePlan* plan);

// Forward-table geometry (MakeRayTracingTable, M.cc:2019-2094).
struct TableGrid {
  double h_top, h_step, th_start, th_step, th_stop;  // LoopStartHeight, HeightStepSize, LoopStartAngle, AngleStepSize, LoopStopAngle
  double loop_stop_h;                                // ice height (receiver in ice) or ice+depth (receiver in air)
  double depth_signed, ice_h;
  int in_ice;
  int64_t n_h, n_th;                                 // TotalHeightSteps, TotalAngleSteps (reference double expressions)
  int64_t first_skipped_row;                         // rows with h<=0 are skipped by the reference (M.cc:2082); n_h if none
};
int make_grid(double depth_m, double ice_m, double h_top, double h_step, double th_start, double th_step,
              double th_stop, TableGrid* g, std::string* err);
// Per-row transmitter data for rows [r0,r1): height, n(h) (host libm), top layer.
void grid_rows(const AirIceMedium& m, const TableGrid& g, int64_t r0, int64_t r1, std::vector<double>* h,
               std::vector<double>* ntx, std::vector<int>* kt);

}  // namespace airice
