// TEST INFRASTRUCTURE ONLY: compiles the product's device math headers (airice_core.cuh, airice_solve.cuh)
// for the HOST so that `pytest -m "not gpu"` can check the kernel arithmetic against the oracle in a
// container without a GPU.  It is never loaded by the product; the shipped path is CUDA only.
#include <cstring>
#include <string>

#include "airice_host.hpp"
#include "airice_solve.cuh"
#include "airice_inice.cuh"
#include "airice_path.cuh"
#include "airice_inice_machine.cuh"
#include "airice_inice_att.cuh"

using namespace airice;

static AirIceMedium g_m;

extern "C" {
int sim_load(const char* path, int variant) {
  std::string err; double n0; int np;
  const int rc = load_medium(path, variant, &g_m, &n0, &np, &err);
  static double clamp[2 * AIRICE_CLAMP_N];
  if (rc == 0) { make_clamp_table(g_m, clamp); g_m.clamp_tab = clamp; }
  return rc;
}
void sim_medium(double* out) {
  out[0] = g_m.nlayers;
  for (int i = 0; i < 5; i++) { out[1 + i] = g_m.hlo[i] * 100; out[6 + i] = g_m.B[i]; out[11 + i] = g_m.C[i]; }
  out[16] = g_m.A_ice; out[17] = g_m.B_ice; out[18] = g_m.C_ice; out[19] = g_m.pi;
}
static int top_layer(double h) {
  for (int i = 0; i < g_m.nlayers; i++) if (h >= g_m.hlo[i] && h < g_m.hlo[i + 1]) return i;
  return -1;
}
// GetRayTracingSolutions layout: out[18]
void sim_forward(double theta, double h, double ice, double depth, int inice, double* out) {
  // GetRayTracingSolutions takes the surface height as given; depth only matters for the ice leg
  AirIcePlan p; make_plan(g_m, ice, inice ? depth : 0.0, &p);
  const int kt = top_layer(h);
  const double ntx = n_air(g_m, h);
  const double L = airice_L_of_theta(g_m, ntx, theta);
  AirIceRay r;
  airice_ray_full<true>(g_m, p, kt, h, ntx, L, inice != 0, true, true, r);
  for (int i = 0; i < 18; i++) out[i] = 0;
  out[1] = h; out[2] = r.x_air + r.x_ice; out[3] = r.x_air; out[4] = r.x_ice;
  out[5] = (r.t_ice + r.t_air) * g_m.c; out[6] = r.t_air * g_m.c; out[7] = r.t_ice * g_m.c;
  out[8] = (r.t_ice + r.t_air) * 1e9; out[9] = r.t_air * 1e9; out[10] = r.t_ice * 1e9;
  out[11] = theta; out[12] = r.inc_ice_deg; out[13] = r.recv_deg; out[14] = r.trans_s; out[15] = r.trans_p;
  out[16] = r.p_air; out[17] = r.p_ice;
}
// the command-line solver (variant 2: Brent, Air2IceRayTracing.C): out[7] = launch angle, X_air, incident angle on ice, L,
// t_air [ns], X_ice, receive angle, then t_ice [ns] in out[7]; depth negative in ice
int sim_solve_cli(double h, double d, double ice, double depth, double* out) {
  AirIcePlan p; make_plan(g_m, ice, depth, &p);
  const int kt = top_layer(h);
  const double ntx = 1.0 + g_m.B[kt < 0 ? 0 : kt] * exp(-g_m.C[kt < 0 ? 0 : kt] * h);
  double ta;
  const double thR = airice_straight_angle(g_m, h, d, ice, depth, ta);
  int nev = 0;
  const double theta = airice_solve_theta_cli(g_m, p, kt, h, ntx, d, thR, nev);
  const double L = airice_L_of_theta(g_m, ntx, theta);
  AirIceRay r;
  airice_ray_full<false>(g_m, p, kt, h, ntx, L, p.has_ice != 0, true, true, r);
  out[0] = theta; out[1] = r.x_air; out[2] = r.inc_ice_deg; out[3] = L; out[4] = r.t_air * 1e9; out[5] = r.x_ice; out[6] = r.recv_deg;
  out[7] = r.t_ice * 1e9;
  return nev;
}
// GetHorizontalDistanceToIntersectionPoint layout (cm/rad): out[9]; stats[0..2] = newton evals, replay evals, theta*
int sim_solve_cm(double h_cm, double d_cm, double depth_cm, double ice_cm, double* out, double* stats) {
  const double h = h_cm / 100, d = d_cm / 100, ice = ice_cm / 100, depth = depth_cm / 100;
  AirIcePlan p; make_plan(g_m, ice, depth, &p);
  const int kt = top_layer(h);
  const double ntx = 1.0 + g_m.B[kt < 0 ? 0 : kt] * exp(-g_m.C[kt < 0 ? 0 : kt] * h);
  double ta;
  const double thR = airice_straight_angle(g_m, h, d, ice, depth, ta);
  AirIceSolveStat st; double ths;
  const double theta = airice_solve_theta(g_m, p, kt, h, ntx, d, thR, ta, ths, st);
  const double L = airice_L_of_theta(g_m, ntx, theta);
  AirIceRay r;
  airice_ray_full<false>(g_m, p, kt, h, ntx, L, p.has_ice != 0, true, true, r);
  out[0] = (r.t_ice * g_m.c) * 100; out[1] = (r.t_air * g_m.c) * 100; out[2] = r.p_ice * 100; out[3] = r.p_air * 100;
  out[4] = theta * (g_m.pi / 180); out[5] = r.x_air * 100; out[6] = r.trans_s; out[7] = r.trans_p;
  out[8] = r.recv_deg * (g_m.pi / 180);
  if (stats) { stats[0] = st.n_newton; stats[1] = st.n_replay; stats[2] = ths; }
  return airice_check_solution(r.x_ice + r.x_air, d) ? 1 : 0;
}
void sim_solve_cm_batch(long n, const double* h_cm, const double* d_cm, double depth_cm, double ice_cm, double* out,
                        unsigned char* ok, double* stats) {
  for (long i = 0; i < n; i++) ok[i] = (unsigned char)sim_solve_cm(h_cm[i], d_cm[i], depth_cm, ice_cm, out + 9 * i, stats ? stats + 3 * i : nullptr);
}
// first pass of the two-pass launch (DEFER = true) next to the complete solve: theta of both and the `hard` flag
void sim_solve_defer_batch(long n, const double* h_cm, const double* d_cm, double depth_cm, double ice_cm, double* theta_defer,
                           double* theta_full, unsigned char* hard) {
  const double ice = ice_cm / 100, depth = depth_cm / 100;
  AirIcePlan p; make_plan(g_m, ice, depth, &p);
  for (long i = 0; i < n; i++) {
    const double h = h_cm[i] / 100, d = d_cm[i] / 100;
    const int kt = top_layer(h);
    const double ntx = 1.0 + g_m.B[kt < 0 ? 0 : kt] * exp(-g_m.C[kt < 0 ? 0 : kt] * h);
    double ta;
    const double thR = airice_straight_angle(g_m, h, d, ice, depth, ta);
    AirIceSolveStat st; double ths; bool hd;
    theta_defer[i] = airice_solve_theta_t<true>(g_m, p, kt, h, ntx, d, thR, ta, ths, st, hd);
    hard[i] = hd ? 1 : 0;
    theta_full[i] = airice_solve_theta(g_m, p, kt, h, ntx, d, thR, ta, ths, st);
  }
}
void sim_forward_batch(long n, const double* th, const double* h, double ice, double depth, int inice, double* out) {
  for (long i = 0; i < n; i++) sim_forward(th[i], h[i], ice, depth, inice, out + 18 * i);
}
// d/dL check: returns X and analytic dX/dL
double sim_x_total(double h, double ice, double depth, double L, double* dXdL) {
  AirIcePlan p; make_plan(g_m, ice, depth, &p);
  return airice_x_dx(g_m, p, top_layer(h), h, n_air(g_m, h), L, *dXdL);
}
// in-ice solver: IceRayTracing::IceRayTracing(0,z0,x1,z1) layout, out[29] per pair; returns nothing
void sim_inice_batch(long n, const double* z0, const double* x1, const double* z1, double* out, int* mask) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  for (long i = 0; i < n; i++) mask[i] = inice_solve(m, z0[i], x1[i], z1[i], out + 29 * i);
}
}
extern "C" {
// literal ladder vs stepped state machine for every pair whose refracted search runs: writes both 6-value results,
// returns how many pairs ran the ladder; evals[i] = fRaa evaluations the machine requested (-1: ladder not needed),
// steps[i] = machine steps (critical path length)
long sim_inice_ladder_compare(long n, const double* z0, const double* x1, const double* z1, double* direct, double* stepped,
                              int* evals, int* steps) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  long ran = 0;
  for (long i = 0; i < n; i++) {
    double o[29]; bool needs;
    const int mask = inice_solve_dr(m, z0[i], x1[i], z1[i], o, needs);
    evals[i] = -1; steps[i] = -1;
    for (int k = 0; k < 6; k++) { direct[6 * i + k] = 0; stepped[6 * i + k] = 0; }
    if (!needs) continue;
    bool flip;
    const InIcePair g = inice_make_pair(m, z0[i], x1[i], z1[i], flip);
    const InIceRaLadder a = inice_ra_ladder(m, g, flip, (mask & 1) == 0, (mask & 2) == 0, o[20]);
    const InIceRaLadder b = inice_ra_ladder_stepped(m, g, flip, (mask & 1) == 0, (mask & 2) == 0, o[20], &evals[i], &steps[i]);
    const double da[6] = {a.lv[0], a.lv[1], a.cz[0], a.cz[1], a.zm[0], a.zm[1]};
    const double db[6] = {b.lv[0], b.lv[1], b.cz[0], b.cz[1], b.zm[0], b.zm[1]};
    for (int k = 0; k < 6; k++) { direct[6 * i + k] = da[k]; stepped[6 * i + k] = db[k]; }
    ran++;
  }
  return ran;
}
}
extern "C" {
double sim_inice_zmax(double L) { return inice_zmax(1.78, -0.43, 0.0132, L); }
// fRaa at L for the pair (z0, x1, z1): full evaluation into y[0], zm[0]; shortcut into y[1], zm[1]; returns 1 if the
// shortcut applied
int sim_inice_fraa_shortcut(double L, double z0, double x1, double z1, double* y, double* zm) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  bool flip;
  const InIcePair g = inice_make_pair(m, z0, x1, z1, flip);
  y[0] = inice_fraa_eval(g, L, zm[0]);
  y[1] = 0; zm[1] = 0;
  return inice_fraa_shortcut(m.A, m.B, exp(-m.C * 5000.0), g.x1, L, y[1], zm[1]) ? 1 : 0;
}
// GetRayTracingSolutions(RxDepth, Distance, TxDepth) through the host build of the device code
void sim_inice_two_rays_batch(long n, const double* rx, const double* dist, const double* tx, double* out10, int* ignore2,
                              int* type2) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  for (long i = 0; i < n; i++) {
    double o[29];
    inice_solve(m, tx[i], dist[i], rx[i], o);
    inice_pick_two_rays(m, o, rx[i], dist[i], tx[i], out10 + 10 * i, ignore2 + 2 * i, type2 + 2 * i);
  }
}
// the same with the attenuation outputs (QAGS per candidate ray); stats[0] = interval-storage overflows, [1] = most intervals
void sim_inice_two_rays_att_batch(long n, const double* rx, const double* dist, const double* tx, double A0, double frequency,
                                  double* out10, double* att2, int* ignore2, int* stats) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  const InIceAttModel am = {A0, frequency, log(0.0001), log(3.16), log(frequency)};
  int flags = 0, worst = 0, over = 0;
  for (long i = 0; i < n; i++) {
    double o[29], att4[4];
    int ty[2];
    inice_solve(m, tx[i], dist[i], rx[i], o);
    flags = 0;
    inice_candidate_attenuations(m, am, o, rx[i], tx[i], att4, flags, worst);
    over += flags != 0;
    inice_pick_two_rays(m, o, rx[i], dist[i], tx[i], out10 + 10 * i, ignore2 + 2 * i, ty, att4, att2 + 2 * i);
  }
  stats[0] = over; stats[1] = worst;
}
double sim_inice_attenuation(int kind, double A0, double frequency, double z0, double z1, double zmax, double L) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  const InIceAttModel am = {A0, frequency, log(0.0001), log(3.16), log(frequency)};
  int flags = 0, worst = 0;
  return inice_total_attenuation(m, am, kind, z0, z1, zmax, L, flags, worst);
}
// GetFocusingFactor(zT, xR, zR) with the initial {1, 1}
void sim_inice_focusing_batch(long n, const double* zT, const double* xR, const double* zR, double* out2) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  for (long i = 0; i < n; i++) {
    double o[29], a[10], b[10];
    int ig[2], ty[2];
    inice_solve(m, zT[i], xR[i], zR[i], o);
    inice_pick_two_rays(m, o, zR[i], xR[i], zT[i], a, ig, ty);
    const double zb = zR[i] - 0.01;
    inice_solve(m, zT[i], xR[i], zb, o);
    inice_pick_two_rays(m, o, zb, xR[i], zT[i], b, ig, ty);
    double f[2] = {1, 1};
    inice_focusing(m, zT[i], zR[i], a + 2, a + 4, a + 6, b + 4, b + 6, f);
    out2[2 * i] = f[0]; out2[2 * i + 1] = f[1];
  }
}
double sim_inice_table_interp(const float* pos_x, const float* pos_z, int n_x, int n_z, double step_x, double step_z,
                              const double* col, double x, double z) {
  return inice_table_interp(pos_x, pos_z, n_x, n_z, step_x, step_z, col, x, z);
}
// ray-path polyline through the host build of airice_path.cuh; depth negative in ice; returns the point count
long sim_ray_path(double theta, double h, double ice, double depth, long max_points, double* x, double* z) {
  AirIcePlan p; make_plan(g_m, ice, depth < 0 ? depth : 0.0, &p);
  AirIcePathPlan pl;
  const int total = airice_path_plan(g_m, p, theta, h, pl);
  for (long q = 0; q < total && q < max_points; q++) airice_path_point(g_m, p, pl, (int)q, x[q], z[q]);
  return total;
}
double sim_inice_zmax_literal(double L) { return inice_zmax_literal(1.78, -0.43, 0.0132, L); }
double sim_inice_fraa(double L, double z0, double x1, double z1) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  InIcePair g; g.A = m.A; g.B = m.B; g.C = m.C; g.z0 = z0; g.z1 = z1; g.x1 = x1;
  g.n0 = inice_nz(m, z0); g.n1 = inice_nz(m, z1); g.ns = inice_nz(m, 1e-7);
  InIceFRaa f = {g};
  return f(L);
}
double sim_inice_dfraa(double L, double z0, double x1, double z1) {
  const AirIceInIce m = inice_make_model(1.78, -0.43, 0.0132);
  InIcePair g; g.A = m.A; g.B = m.B; g.C = m.C; g.z0 = z0; g.z1 = z1; g.x1 = x1;
  g.n0 = inice_nz(m, z0); g.n1 = inice_nz(m, z1); g.ns = inice_nz(m, 1e-7);
  InIceFRaa f = {g};
  return inice_deriv_central(f, L, 1e-8);
}
}
