// Builds libMultiRayAirIceRefraction.so for callers that include only the header and link, instead of including
// MultiRayAirIceRefraction.cc the way the reference's drivers do.  The two caller-owned vectors get weak definitions
// here so the library also loads stand-alone (an executable's own definitions take precedence).
#include "MultiRayAirIceRefraction.cc"

__attribute__((weak)) std::vector<double> AntennaDepths;
__attribute__((weak)) std::vector<int> AntennaTableAlreadyMade;
