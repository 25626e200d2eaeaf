// atmosphere.cc -- Atmosphere.dat -> layered exponential air model + ice model (host, one-off per context).
//
// Behaviour reproduced from the reference (paths under /root/reference):
//   readATMpar          MultiRayAirIceRefraction.cc:24-71   line 2 = ATMLAY[5] (cm), lines 3-5 = a,b,c; then
//                       abc[4]=abc[3], ATMLAY[4]=150000 m
//   readnhFromFile      MultiRayAirIceRefraction.cc:73-147  rows "h n" from line 7 on, kept when h>-1; a new per-layer
//                       vector is started each time h*100 crosses the next ATMLAY edge; MaxLayers = #vectors+1
//   MakeAtmosphere      MultiRayAirIceRefraction.cc:920-942 natural cubic spline through all kept rows
//   FillInAirRefractiveIndex  MultiRayAirIceRefraction.cc:193-213  C_k = 1/(c_k/100); B_0 from spline(0); B_k by continuity
// The pythonwrapper copy (pythonwrapper/AirIceRayTracing.cc:4-165, 860-882) is identical apart from pi.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "airice_host.hpp"

namespace airice {

namespace {

// Value at xq of the natural cubic spline through (x,y).  The second-derivative system is symmetric
// tridiagonal and diagonally dominant; it is solved with the L D L^T recurrences in the same operation
// order as GSL's solve_tridiag so that spline(0) -- which seeds every B_air -- carries the reference's bits.
double natural_spline_value(const std::vector<double>& x, const std::vector<double>& y, double xq) {
  const long n = (long)x.size();
  if (n < 3) return NAN;
  const long m = n - 2;  // interior unknowns c[1..n-2]
  std::vector<double> c(n, 0.0), rhs(m), dg(m), od(m), gam(m), alp(m), z(m);
  for (long i = 0; i < m; i++) {
    const double hl = x[i + 1] - x[i], hr = x[i + 2] - x[i + 1];
    const double dl = y[i + 1] - y[i], dr = y[i + 2] - y[i + 1];
    const double il = (hl != 0.0) ? 1.0 / hl : 0.0, ir = (hr != 0.0) ? 1.0 / hr : 0.0;
    od[i] = hr;
    dg[i] = 2.0 * (hr + hl);
    rhs[i] = 3.0 * (dr * ir - dl * il);
  }
  if (m == 1) {
    c[1] = rhs[0] / dg[0];
  } else {
    alp[0] = dg[0];
    gam[0] = od[0] / alp[0];
    for (long i = 1; i < m - 1; i++) {
      alp[i] = dg[i] - od[i - 1] * gam[i - 1];
      gam[i] = od[i] / alp[i];
    }
    alp[m - 1] = dg[m - 1] - od[m - 2] * gam[m - 2];
    z[0] = rhs[0];
    for (long i = 1; i < m; i++) z[i] = rhs[i] - gam[i - 1] * z[i - 1];
    for (long i = 0; i < m; i++) z[i] = z[i] / alp[i];
    c[m] = z[m - 1];
    for (long i = m - 2; i >= 0; i--) c[i + 1] = z[i] - gam[i] * c[i + 2];
  }
  long lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    const long mid = (hi + lo) / 2;
    if (x[mid] > xq) hi = mid; else lo = mid;
  }
  const double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo], t = xq - x[lo];
  const double b = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
  const double d3 = (c[lo + 1] - c[lo]) / (3.0 * dx);
  return y[lo] + t * (b + t * (c[lo] + t * d3));
}

bool parse_doubles(const std::string& line, int want, double* out) {
  const char* s = line.c_str();
  for (int i = 0; i < want; i++) {
    char* end;
    out[i] = std::strtod(s, &end);
    if (end == s) return false;
    s = end;
  }
  return true;
}

}  // namespace

int load_medium(const char* path, int variant, AirIceMedium* out, double* n0_out, int* npoints_out, std::string* err) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) {
    if (err) *err = std::string("cannot open atmosphere file: ") + path;
    return -1;
  }
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();
  std::vector<std::string> lines;
  {
    size_t pos = 0;
    while (pos < text.size()) {
      size_t nl = text.find('\n', pos);
      if (nl == std::string::npos) { lines.push_back(text.substr(pos)); break; }
      lines.push_back(text.substr(pos, nl - pos));
      pos = nl + 1;
    }
  }
  const bool ends_with_newline = !text.empty() && text.back() == '\n';
  if (lines.size() < 8) {
    if (err) *err = "atmosphere file too short";
    return -2;
  }
  double atmlay_cm[5], abc[3][5];
  if (!parse_doubles(lines[1], 5, atmlay_cm) || !parse_doubles(lines[2], 5, abc[0]) ||
      !parse_doubles(lines[3], 5, abc[1]) || !parse_doubles(lines[4], 5, abc[2])) {
    if (err) *err = "atmosphere header (ATMLAY / a / b / c lines) is malformed";
    return -3;
  }
  abc[2][4] = abc[2][3];      // abc[4] = abc[3]
  atmlay_cm[4] = 150000 * 100;  // top of the model, cm

  // tabulated n(h): line 7 onwards
  std::vector<double> hs, ns;
  int layer = 0, vectors = 0;
  bool open_vector = false;
  for (size_t i = 6; i < lines.size(); i++) {
    double hv[2];
    if (!parse_doubles(lines[i], 2, hv)) continue;  // blank tail line
    if (hv[0] > -1) {
      hs.push_back(hv[0]);
      ns.push_back(hv[1]);
      open_vector = true;
      if (layer < 5 && hv[0] * 100 >= atmlay_cm[layer]) {
        if (layer > 0) { vectors++; open_vector = false; }
        layer++;
      }
    }
  }
  if (layer > 0) vectors++;
  (void)open_vector;
  // The reference's read loop sees the final row twice when the file ends in a newline and then erases one
  // copy (M.cc:137-140); without the trailing newline the erase removes the genuine last row instead.
  if (!ends_with_newline && !hs.empty()) { hs.pop_back(); ns.pop_back(); }
  if (hs.size() < 3) {
    if (err) *err = "atmosphere table has fewer than 3 usable rows";
    return -4;
  }

  AirIceMedium m;
  std::memset(&m, 0, sizeof(m));
  m.variant = variant;
  m.nlayers = vectors + 1;
  if (m.nlayers > AIRICE_MAX_LAYERS) m.nlayers = AIRICE_MAX_LAYERS;
  m.pi = (variant == 1) ? 4.0 * std::atan(1.0) : 3.1415927;
  m.deg2rad = m.pi / 180.0;
  m.rad2deg = 180 / m.pi;
  m.c = 299792458.0;
  m.tan16 = std::tan(16 * m.deg2rad);
  m.clamp_tab = nullptr;   // set by the owner of the medium (make_clamp_table)
  m.A_ice = 1.78; m.B_ice = -0.43; m.C_ice = 0.0132;
  for (int k = 0; k < 5; k++) m.hlo[k] = atmlay_cm[k] / 100;
  m.hlo[5] = atmlay_cm[4] / 100;

  const double n0 = natural_spline_value(hs, ns, 0.0);
  double N0 = 0;
  for (int k = 0; k < 5; k++) {
    const double hlow = atmlay_cm[k] / 100;
    m.C[k] = 1.0 / (abc[2][k] / 100);
    if (k > 0) N0 = 1.00 + m.B[k - 1] * std::exp(-hlow * m.C[k - 1]);
    if (k == 0) N0 = n0;
    m.B[k] = ((N0 - 1) / std::exp(-hlow * m.C[k]));
  }
  *out = m;
  if (n0_out) *n0_out = n0;
  if (npoints_out) *npoints_out = (int)hs.size();
  return 0;
}

int layer_of(const AirIceMedium& m, double z) {
  const double za = std::fabs(z);
  int which = 0;
  for (int k = 0; k < m.nlayers - 1; k++) {
    if (za < m.hlo[k + 1] && za >= m.hlo[k]) { which = k; break; }
  }
  if (za >= m.hlo[m.nlayers - 1]) which = m.nlayers - 1;
  return which;
}

double n_air(const AirIceMedium& m, double z) {
  const double za = std::fabs(z);
  const int k = layer_of(m, za);
  return 1.00 + m.B[k] * std::exp(-m.C[k] * za);
}

double n_ice(const AirIceMedium& m, double z) {
  z = std::fabs(z);
  return m.A_ice + m.B_ice * std::exp(-m.C_ice * z);
}

void make_plan(const AirIceMedium& m, double ice_h, double depth_signed, AirIcePlan* plan) {
  AirIcePlan p;
  std::memset(&p, 0, sizeof(p));
  if (depth_signed >= 0) {
    ice_h = depth_signed + ice_h;
    p.depth = 0;
    p.has_ice = 0;
  } else {
    p.depth = -depth_signed;
    p.has_ice = (p.depth != 0) ? 1 : 0;
  }
  p.ice_h = ice_h;
  // SkipLayersBelow (M.cc:680-690): index of the layer holding the surface, nlayers if none does
  p.kb = m.nlayers;
  for (int k = 0; k < m.nlayers; k++) {
    if (ice_h >= m.hlo[k] && ice_h < m.hlo[k + 1]) { p.kb = k; break; }
  }
  for (int k = 0; k <= AIRICE_MAX_LAYERS; k++) {
    p.seg[k].neg_c = -1; p.seg[k].inv_neg_c = -1; p.seg[k].stop_x = 0; p.seg[k].stop_n = 1; p.seg[k].start_x = 0; p.seg[k].start_n = 1; p.seg[k].relay = 1; p.seg[k].ln_relay = 0;
  }
  for (int k = 0; k < m.nlayers; k++) { p.seg[k].neg_c = -m.C[k]; p.seg[k].inv_neg_c = 1.0 / p.seg[k].neg_c; }
  for (int k = p.kb; k < m.nlayers; k++) {
    p.seg[k].stop_x = (k == p.kb) ? ice_h : m.hlo[k];
    p.seg[k].stop_n = n_air(m, p.seg[k].stop_x);
    p.seg[k].start_x = m.hlo[k + 1] - 0.00001;
    p.seg[k].start_n = n_air(m, p.seg[k].start_x);
  }
  for (int k = p.kb; k + 1 < m.nlayers; k++) {
    p.seg[k].relay = p.seg[k].start_n / p.seg[k + 1].stop_n;
    p.seg[k].ln_relay = std::log(p.seg[k].relay);
  }
  for (int k = 0; k <= AIRICE_MAX_LAYERS; k++) { p.seg[k].ho_dn = 0; p.seg[k].ho_dn2 = 0; }
  for (int k = p.kb + 1; k < m.nlayers; k++) {
    p.seg[k].ho_dn = p.seg[k].stop_n - p.seg[k - 1].start_n;
    p.seg[k].ho_dn2 = (p.seg[k].stop_n - p.seg[k - 1].start_n) * (p.seg[k].stop_n + p.seg[k - 1].start_n);
  }
  // ice leg: surface (x=0) down to the receiver (x=depth), GetIcePropagationPar (M.cc:807-869)
  p.seg[AIRICE_ICE_SLOT].neg_c = -m.C_ice;
  p.seg[AIRICE_ICE_SLOT].inv_neg_c = 1.0 / p.seg[AIRICE_ICE_SLOT].neg_c;
  p.seg[AIRICE_ICE_SLOT].start_x = 0.0;
  p.seg[AIRICE_ICE_SLOT].start_n = n_ice(m, 0.0);
  p.seg[AIRICE_ICE_SLOT].stop_x = p.depth;
  p.seg[AIRICE_ICE_SLOT].stop_n = n_ice(m, p.depth);
  for (int k = 0; k <= AIRICE_MAX_LAYERS; k++) {
    const double A = (k == AIRICE_ICE_SLOT) ? m.A_ice : 1.0;
    p.seg[k].f_q_stop = (float)((p.seg[k].stop_n - A) * (p.seg[k].stop_n + A));
    p.seg[k].f_pa_stop = (float)(A * (p.seg[k].stop_n - A));
    p.seg[k].f_q_start = (float)((p.seg[k].start_n - A) * (p.seg[k].start_n + A));
    p.seg[k].f_pa_start = (float)(A * (p.seg[k].start_n - A));
    p.seg[k].f_cdx = (float)(p.seg[k].neg_c * (p.seg[k].stop_x - p.seg[k].start_x));
    p.seg[k].f_inv_neg_c = (float)p.seg[k].inv_neg_c;
  }
  *plan = p;
}

int make_grid(double depth_m, double ice_m, double h_top, double h_step, double th_start, double th_step,
              double th_stop, TableGrid* g, std::string* err) {
  if (!(h_step > 0) || !(th_step > 0) || !(th_stop >= th_start)) {
    if (err) *err = "table grid: steps must be positive and th_stop >= th_start";
    return -1;
  }
  TableGrid t;
  t.h_top = h_top; t.h_step = h_step; t.th_start = th_start; t.th_step = th_step; t.th_stop = th_stop;
  t.depth_signed = depth_m; t.ice_h = ice_m;
  t.in_ice = depth_m < 0 ? 1 : 0;
  t.loop_stop_h = t.in_ice ? ice_m : ice_m + depth_m;                     // M.cc:2054-2059
  t.n_th = (int)(std::floor((th_stop - th_start) / th_step) + 1);          // M.cc:15
  t.n_h = (int)(std::floor((h_top - t.loop_stop_h) / h_step) + 1);         // M.cc:2061
  if (t.n_th < 1 || t.n_h < 1) {
    if (err) *err = "table grid is empty";
    return -2;
  }
  t.first_skipped_row = t.n_h;
  for (int64_t r = 0; r < t.n_h; r++) {
    const double h = h_top - h_step * (double)r;
    if (!(h > 0)) { t.first_skipped_row = r; break; }
  }
  *g = t;
  return 0;
}

void grid_rows(const AirIceMedium& m, const TableGrid& g, int64_t r0, int64_t r1, std::vector<double>* h,
               std::vector<double>* ntx, std::vector<int>* kt) {
  const int64_t n = r1 - r0;
  h->resize(n); ntx->resize(n); kt->resize(n);
  for (int64_t r = r0; r < r1; r++) {
    double hv = g.h_top - g.h_step * (double)r;                         // M.cc:2080
    if (hv != g.loop_stop_h && r == g.n_h - 1) hv = g.loop_stop_h;      // M.cc:2089-2091
    int k = -1;
    for (int i = 0; i < m.nlayers; i++)
      if (hv >= m.hlo[i] && hv < m.hlo[i + 1]) { k = i; break; }       // SkipLayersAbove, M.cc:666-676
    (*h)[r - r0] = hv;
    (*kt)[r - r0] = k;
    (*ntx)[r - r0] = n_air(m, hv);
  }
}

// lo_j and sin((180 - lo_j) deg2rad) of the clamped bracket's scan (M.cc:1490-1511): lo accumulates 0.05 exactly as
// the reference's `lo = lo + 0.05` does
void make_clamp_table(const AirIceMedium& m, double* tab) {
  double lo = 90.001;
  for (int j = 0; j < AIRICE_CLAMP_N; j++) {
    tab[2 * j] = lo;
    tab[2 * j + 1] = std::sin((180 - lo) * m.deg2rad);
    lo = lo + 0.05;
  }
}

}  // namespace airice
