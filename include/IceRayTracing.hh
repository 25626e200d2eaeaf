// IceRayTracing.hh -- source-compatible host API of the in-ice solver on top of the B200 C ABI.
//
// Replaces, for the hot path only, /root/reference/IceRayTracing.hh / IceRayTracing.cc: the 4-argument
// IceRayTracing::IceRayTracing(x0, z0, x1, z1) defined at IceRayTracing.cc:1745 (the reference header declares a
// 5-argument form at IceRayTracing.hh:186; both are offered here), and the ice-model setters SetA/SetB/SetC
// (IceRayTracing.cc:7-17), and GetRayTracingSolutions (IceRayTracing.cc:2907; two-ray selection).  The returned array has the reference's 29 slots and is owned by the caller (delete[]),
// as in the reference.  Attenuation (GetTotalAttenuation*, the AttRay output), GetFocusingFactor and the in-ice interpolation
// table (SetNumberOfAntennas / MakeTable / GetInterpolatedValue) are provided as well; ray-path dumps and
// the constant-n variants are not.  No atmosphere file is needed: the context behind this namespace is created ice-only.
#ifndef IRT_HEAD_B200
#define IRT_HEAD_B200
#include <string>
#include <vector>

namespace IceRayTracing {

static constexpr double pi = 3.14159265359;       // IceRayTracing.hh:41 (sic)
static constexpr double c_light_ms = 299792458;   // IceRayTracing.hh:43
extern double A_ice, B_ice, C_ice;                // IceRayTracing.hh:54-56

void SetA(double &A);
void SetB(double &B);
void SetC(double &C);
void SetDevice(int device);                       // B200 extra
void SetAtmosphereFile(const std::string &path);  // kept for source compatibility with round 1; the in-ice context needs none

// out[0..3] launch angles D,R,Ra1,Ra2; [4..7] times; [8..11] receive angles (-1000 = branch absent); [12..17] sub-times;
// [18] incidence on the surface; [19..22] L; [23..24] z_max; [25..28] geometric paths.  x0 must be 0 (as in the reference's
// own callers); it is accepted for signature compatibility.
double *IceRayTracing(double x0, double z0, double x1, double z1);
double *IceRayTracing(double x0, double z0, double x1, double z1, bool PlotRayPaths);

// batch form (new): n pairs, out[col*n + i] with 29 columns, mask[i] bit0..3 = D,R,Ra1,Ra2 present
int IceRayTracingBatch(long n, const double *z0, const double *x1, const double *z1, double *out, unsigned char *mask);

// The two physical rays of a pair, ordered by arrival time (IceRayTracing.cc:2907-3210, same argument list), with
// AttRay[k] = 1 - attenuation of ray k at `frequency` [GHz] for amplitude A0 (QAGS per ray on the GPU).
void GetRayTracingSolutions(double RxDepth, double Distance, double TxDepth, double TimeRay[2], double PathRay[2],
                            double LaunchAngle[2], double RecieveAngle[2], int IgnoreCh[2], double IncidenceAngleInIce[2],
                            double A0, double frequency, double AttRay[2]);
// batch form (new): out[col*n + i], columns TimeRay[0..1], PathRay[0..1], LaunchAngle[0..1], RecieveAngle[0..1],
// IncidenceAngleInIce[0..1]; ignore[k*n + i] = IgnoreCh[k]
int GetRayTracingSolutionsBatch(long n, const double *RxDepth, const double *Distance, const double *TxDepth, double *out,
                                int *ignore);
// the same with attenuation: att[k*n + i] = AttRay[k]
int GetRayTracingSolutionsBatch(long n, const double *RxDepth, const double *Distance, const double *TxDepth, double A0,
                                double frequency, double *out, double *att, int *ignore);

// IceRayTracing.cc:135-219
double GetIceTemperature(double z);
double GetIceAttenuationLength(double z, double frequency);
double GetTotalAttenuationDirect(double A0, double frequency, double z0, double z1, double Lvalue);
double GetTotalAttenuationReflected(double A0, double frequency, double z0, double z1, double Lvalue);
double GetTotalAttenuationRefracted(double A0, double frequency, double z0, double z1, double zmax, double Lvalue);
// IceRayTracing.cc:3218-3293; focusing[] is read (initial values) and written like the reference's
void GetFocusingFactor(double zT, double xR, double zR, double focusing[2]);
int GetFocusingFactorBatch(long n, const double *zT, const double *xR, const double *zR, double *out);   // out[k*n + i]

// in-ice interpolation table (IceRayTracing.cc:2614-2905; grid globals IceRayTracing.hh:33-36), kept on the GPU
extern double GridStepSizeX_O, GridStepSizeZ_O, GridWidthX, GridWidthZ;
void SetNumberOfAntennas(int numberOfAntennas);
void MakeTable(double ShowerHitDistance, double ShowerDepth, double zR, int AntNum);
double GetInterpolatedValue(double xT, double zT, int rtParameter, int AntNum);
int GetInterpolatedValueBatch(long n, const double *xT, const double *zT, int rtParameter, int AntNum, double *out);
int GetTableColumn(int AntNum, int col, std::vector<double> &out);   // GridZValueb[AntNum][col]

}  // namespace IceRayTracing
#endif
