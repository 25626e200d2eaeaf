// MultiRayAirIceRefraction.h -- source-compatible host API of the B200-native air->ice ray solver.
//
// Same namespace, names, argument order, units (cm / rad at the CoREAS entry points, m / deg underneath) and
// failure conventions (bool + sentinels, no exceptions) as the reference's MultiRayAirIceRefraction.h, for the hot
// path: MakeAtmosphere, MakeRayTracingTable, GetHorizontalDistanceToIntersectionPoint[_Table], Air2IceRayTracing,
// GetRayTracingSolutions, MakeTable, GetInterpolatedValue.  Each scalar call is a batch-of-1 launch on the GPU
// (correct, slow); the *Batch entry points below carry the throughput.  Callers keep including
// "MultiRayAirIceRefraction.cc" exactly as RunMultiRayCode.C:1 does, and keep defining the two extern vectors.
//
// Reference declarations this replaces: /root/reference/MultiRayAirIceRefraction.h:23-24 (externs), :157, :170,
// :189, :191, :193, :195, :198, :204; tunable globals MultiRayAirIceRefraction.cc:4-21 and .h:42-54, :72-74.
#ifndef _INCLUDE_MULTIRAYAIRICEREFRACTION_H_
#define _INCLUDE_MULTIRAYAIRICEREFRACTION_H_

#include <cmath>
#include <iostream>
#include <string>
#include <vector>

extern std::vector<double> AntennaDepths;            // defined by the caller (MultiRayAirIceRefraction.h:23)
extern std::vector<int> AntennaTableAlreadyMade;     // defined by the caller (MultiRayAirIceRefraction.h:24)

// forward-table grid, as in the reference (MultiRayAirIceRefraction.cc:12-21); change before MakeRayTracingTable
extern double AngleStepSize, LoopStartAngle, LoopStopAngle, HeightStepSize, LoopStartHeight, LoopStopHeight;
extern int TotalAngleSteps, TotalHeightSteps;
extern double MaxAirTxHeight, MinAirTxHeight;

namespace MultiRayAirIceRefraction {

static const double pi = 3.1415927;        // MultiRayAirIceRefraction.h:29 (sic)
static const double spedc = 299792458.0;   // MultiRayAirIceRefraction.h:30

// mutable ice model n(z) = A + B exp(-C z) (MultiRayAirIceRefraction.h:64-74); read at every call
extern double A_ice, B_ice, C_ice;
// old solve-per-cell grid (MultiRayAirIceRefraction.h:42-54)
extern double GridStartTh, GridStopTh, GridStepSizeH_O, GridStepSizeTh_O, GridStartH, GridStopH, GridWidthH, GridWidthTh;
extern int GridPoints, TotalStepsH_O, TotalStepsTh_O;
extern std::vector<double> GridPositionH, GridPositionTh;
extern std::vector<double> GridZValue[10];

// B200 extras: which GPU to use and where Atmosphere.dat lives (default: ./Atmosphere.dat like the reference)
void SetDevice(int device);
void SetAtmosphereFile(const std::string &path);

int MakeAtmosphere();
int MakeRayTracingTable(double AntennaDepth, double IceLayerHeight, int AntennaNumber);
// all antennas in one pass (shared air walk); appends one table per depth, in order (new, not in the reference)
int MakeRayTracingTables(const std::vector<double> &AntennaDepths_cm, double IceLayerHeight);
bool GetHorizontalDistanceToIntersectionPoint(double SrcHeightASL, double HorizontalDistanceToRx,
                                              double RxDepthBelowIceBoundary, double IceLayerHeight,
                                              double &opticalPathLengthInIce, double &opticalPathLengthInAir,
                                              double &geometricalPathLengthInIce, double &geometricalPathLengthInAir,
                                              double &launchAngle, double &horizontalDistanceToIntersectionPoint,
                                              double &transmissionCoefficientS, double &transmissionCoefficientP,
                                              double &RecievedAngleInIce);
bool GetHorizontalDistanceToIntersectionPoint_Table(double SrcHeightASL, double HorizontalDistanceToRx,
                                                    double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                    int AntennaNumber, double &opticalPathLengthInIce,
                                                    double &opticalPathLengthInAir, double &geometricalPathLengthInIce,
                                                    double &geometricalPathLengthInAir, double &launchAngle,
                                                    double &horizontalDistanceToIntersectionPoint,
                                                    double &transmissionCoefficientS, double &transmissionCoefficientP,
                                                    double &RecievedAngleInIce);
void Air2IceRayTracing(double AirTxHeight, double HorizontalDistance, double IceLayerHeight, double AntennaDepth,
                       double StraightAngle, double dummy[20]);
void GetRayTracingSolutions(double RayLaunchAngleInAir, double AirTxHeight, double IceLayerHeight, double AntennaDepth,
                            double dummy[20], bool &InIce);
void MakeTable(double IceLayerHeight, double AntennaDepth);
double GetInterpolatedValue(double hR, double thR, int rtParameter);
// batched form on the GPU-resident grid of the last MakeTable (new): out[i] = GetInterpolatedValue(hR[i], thR[i], rtParameter)
int GetInterpolatedValueBatch(long n, const double *hR, const double *thR, int rtParameter, double *out);
// persistence of a forward table (new; the reference rebuilds every table per process): writes / appends table AntennaNumber
int SaveRayTracingTable(int AntennaNumber, const std::string &path);
int LoadRayTracingTable(const std::string &path);   // appended like a MakeRayTracingTable call; returns its index or -1

// ---- the medium model as the reference exposes it (MultiRayAirIceRefraction.h:90-119): parameters of the parsed
// atmosphere held by the GPU context (needs MakeAtmosphere(), i.e. a GPU) and of the mutable ice model.  z in metres.
double GetB_ice(double z);
double GetC_ice(double z);
double Getnz_ice(double z);
double GetB_air(double z);
double GetC_air(double z);
double Getnz_air(double z);
// field Fresnel coefficients air -> ice at the surface, incident angle in rad (MultiRayAirIceRefraction.cc:267-337)
double Refl_S(double thetai, double IceLayerHeight);
double Trans_S(double thetai, double IceLayerHeight);
double Refl_P(double thetai, double IceLayerHeight);
double Trans_P(double thetai, double IceLayerHeight);

// ---- batch entry points (new): n pairs per call, SoA outputs out[col*n + i] in the order of the by-reference
// arguments above (opt ice, opt air, geo ice, geo air, launch, X_air, T_S, T_P, received), flags in ok[i].
int GetHorizontalDistanceToIntersectionPointBatch(long n, const double *SrcHeightASL, const double *HorizontalDistanceToRx,
                                                  double RxDepthBelowIceBoundary, double IceLayerHeight, double *out,
                                                  unsigned char *ok);
// The same with one array per by-reference argument of the scalar call; a NULL array (or NULL ok) is an output the
// caller does not read: it is neither computed into a column nor copied back over PCIe, which is what bounds the call.
int GetHorizontalDistanceToIntersectionPointBatch(long n, const double *SrcHeightASL, const double *HorizontalDistanceToRx,
                                                  double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                  double *opticalPathLengthInIce, double *opticalPathLengthInAir,
                                                  double *geometricalPathLengthInIce, double *geometricalPathLengthInAir,
                                                  double *launchAngle, double *horizontalDistanceToIntersectionPoint,
                                                  double *transmissionCoefficientS, double *transmissionCoefficientP,
                                                  double *RecievedAngleInIce, unsigned char *ok);
int GetHorizontalDistanceToIntersectionPoint_TableBatch(long n, const double *SrcHeightASL,
                                                        const double *HorizontalDistanceToRx,
                                                        double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                        int AntennaNumber, double *out, unsigned char *ok);
int GetHorizontalDistanceToIntersectionPoint_TableBatch(long n, const double *SrcHeightASL,
                                                        const double *HorizontalDistanceToRx,
                                                        double RxDepthBelowIceBoundary, double IceLayerHeight,
                                                        int AntennaNumber, double *opticalPathLengthInIce,
                                                        double *opticalPathLengthInAir, double *geometricalPathLengthInIce,
                                                        double *geometricalPathLengthInAir, double *launchAngle,
                                                        double *horizontalDistanceToIntersectionPoint,
                                                        double *transmissionCoefficientS, double *transmissionCoefficientP,
                                                        double *RecievedAngleInIce, unsigned char *ok);
// page-locks / releases a buffer the caller passes to the *Batch functions again and again (cudaHostRegister): the batch
// calls then run at the PCIe rate instead of staging every copy through the driver
int PinHostBuffer(void *p, size_t bytes);
int UnpinHostBuffer(void *p);
// copies column `col` (0..10, reference AllTableAllAntData order) of antenna table `AntennaNumber` to the host
int GetTableColumn(int AntennaNumber, int col, std::vector<float> &out);

}  // namespace MultiRayAirIceRefraction
#endif
