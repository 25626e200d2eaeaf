// capi.cu -- extern "C" layer of libairice_b200.so (declarations + reference citations in include/airice_b200.h).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/airice_b200.h"
#include "airice_host.hpp"
#include "kernels.cuh"

using namespace airice;

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  // the runtime also remembers a failed call as its "last error", which the next launch's cudaGetLastError() check would
  // report as its own (e.g. a refused airice_host_register followed by a solve): it has been reported here, forget it
  cudaGetLastError();
  return -100 - (int)e;
}
#define NEED_AIR(c)                                                                                              \
  do {                                                                                                           \
    if ((c)->no_air) return fail(-9, "this context was created without an atmosphere file: in-ice entry points only"); \
  } while (0)
#define CK(call)                                  \
  do {                                            \
    cudaError_t e__ = (call);                     \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)
}  // namespace

struct airice_ctx {
  int device = 0;
  bool no_air = false;               // created without an atmosphere file: only the in-ice entry points work
  AirIceMedium medium;
  double n0 = 0;
  int npoints = 0;
  std::map<std::pair<double, double>, AirIcePlan> plans;
  // host-API pipeline: 2 device staging slots on 2 streams (copies go straight from/to the caller's buffers;
  // when those are pinned the H2D of chunk k+1, the kernel of chunk k and the D2H of chunk k-1 overlap)
  static const int kSlots = 2;
  cudaStream_t streams[kSlots] = {nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[kSlots] = {nullptr, nullptr};   // airice_solve_multi_device's fork/join
  void* dev[kSlots] = {nullptr, nullptr};
  size_t slot_bytes = 0;
  // scalar and small calls of the host API (the reference's scalar functions are batch-of-1 calls): one page-locked,
  // device-mapped block that the kernel reads its inputs from and writes its results to directly -- one launch and one
  // synchronisation per call instead of four copy calls around the launch
  static const int kSmall = 64;
  void* zc_host = nullptr;
  void* zc_dev = nullptr;
  // per-row transmitter data (height, n(h), top layer) of the last table grid built: uploaded once, reused by
  // every rebuild of the same rows (MakeRayTracingTable is called once per antenna depth on the same grid)
  // in-ice solver scratch: compaction list + counter (+ private mask / L_R columns when the caller passes none)
  void* inice_scratch[kSlots] = {nullptr, nullptr};   // one per staging slot, so that chunks of the host API overlap
  size_t inice_bytes[kSlots] = {0, 0};
  double* inice_cols = nullptr;      // 29 columns of one chunk for airice_inice_two_rays_*
  void* path_plans = nullptr;        // per-ray plans of airice_ray_path_*
  int32_t* quad_stats = nullptr;     // attenuation quadrature: overflow count, largest interval count
  double* focus_scratch = nullptr;   // airice_inice_focusing_device / airice_inice_table_create intermediates
  int64_t focus_cap = 0;
  double* clamp_tab = nullptr;       // device copy of the clamped-bracket table (medium.clamp_tab points at it)
  // buffers of destroyed tables kept for the next table of the same size (a per-antenna loop creates and destroys 64
  // tables of 384 + 455 MB: cudaMalloc / cudaFree of that size cost more than building the table, and jitter wildly)
  // events: recorded at release time on every stream that used the table; the next owner's build stream waits on them
  // (no host synchronisation when a table is destroyed)
  struct SpareBuf { void* p; size_t bytes; std::vector<cudaEvent_t> events; };
  std::vector<SpareBuf> spare;
  std::vector<airice_table*> tables;   // live tables; airice_destroy detaches them
  // scratch of the solve kernel's two-pass launch, one per stream that has run a solve (launches on one stream are
  // ordered, so they can share it): a counter in the first 256 bytes, then room for `cap` deferred pair indices
  struct DeferScratch { int32_t* buf = nullptr; int64_t cap = 0; };
  std::map<cudaStream_t, DeferScratch> defer;
  size_t spare_cap = 0;              // bytes of freed table buffers the context may hold (set on first release)
  size_t path_plan_cap = 0;
  int64_t inice_cols_n = 0;
  // per-row transmitter data (height, n(h), top layer) of the grids tables were built on lately: a caller that alternates
  // between a few grids (row shards of one grid, two receiver depths in air) neither re-uploads nor synchronises
  struct RowCache {
    double key[7] = {0, 0, 0, 0, 0, 0, 0};
    int64_t r0 = -1, r1 = -1;
    double* d_rows = nullptr;
    int* d_kt = nullptr;
    uint64_t used = 0;          // value of row_clock at the last use
  };
  static constexpr int kRowCaches = 4;
  RowCache rows[kRowCaches];
  uint64_t row_clock = 0;

  const AirIcePlan& plan(double ice_m, double depth_m) {
    auto key = std::make_pair(ice_m, depth_m);
    auto it = plans.find(key);
    if (it == plans.end()) {
      AirIcePlan p;
      make_plan(medium, ice_m, depth_m, &p);
      it = plans.emplace(key, p).first;
    }
    return it->second;
  }
};

struct airice_inice_table {
  int device = 0;
  int n_x = 0, n_z = 0;
  int64_t points = 0;
  double step_x = 0, step_z = 0;
  std::vector<float> pos_x, pos_z;       // GridPositionXb / GridPositionZb (float in the reference)
  double* block = nullptr;               // 13 columns, then the two float position arrays
  float* d_pos_x = nullptr;
  float* d_pos_z = nullptr;
};

struct airice_oldtable {
  airice_ctx* ctx = nullptr;
  int device = 0;
  int n_h = 0, n_th = 0;
  int64_t points = 0;
  double start_h = 0, stop_h = 0, step_h = 0, start_th = 0, stop_th = 0, step_th = 0;
  std::vector<double> pos_h, pos_th;     // GridPositionH / GridPositionTh
  double* block = nullptr;               // 9 columns, then the two position arrays (one allocation)
  double* d_pos_h = nullptr;
  double* d_pos_th = nullptr;
};

struct airice_table {
  airice_ctx* ctx = nullptr;
  bool owns = false;
  // reference layout (column-major), owned or wrapped.  Tables built by the fused multi-antenna pass hold the lookup
  // layout only (a lookup never reads the columns, and writing them is 46 % of the pass's HBM traffic); the columns are
  // then made from the records on first use (airice_table_copy_column / _column_ptr / _save): same floats.
  mutable float* cols[AIRICE_TABLE_NCOLS32] = {nullptr};
  // lookup layout, always owned: dense X, 48-byte records, per-row height, per-row trim ranges (one allocation)
  void* pack = nullptr;
  mutable size_t cols_bytes = 0;
  size_t pack_bytes = 0;
  float* x = nullptr;
  float4* rec = nullptr;
  float* row_h = nullptr;
  int* row_first = nullptr;
  int* row_last = nullptr;
  float* rowblk = nullptr;     // per-row header blocks, kernels.cuh
  float4* rowpar = nullptr;    // per-row coordinate parameters of the position table
  uint16_t* lut = nullptr;     // position table: AIRICE_LUT_EDGES u16 per row (nullptr: rows too long for it)
  int lut_shift = -1;
  int64_t n_h = 0, n_th = 0, cells = 0;
  double loop_stop_h = 0, h_step = 0;
  mutable std::vector<cudaStream_t> used;   // streams that ran lookups on this table (besides the build stream 0)
  void note_stream(cudaStream_t s) const {
    for (cudaStream_t u : used) if (u == s) return;
    used.push_back(s);
  }
  LookupTable view() const {
    LookupTable t;
    t.x = x; t.rec = rec; t.row_h = row_h;
    t.cells = cells; t.n_h = (int)n_h; t.n_th = (int)n_th;
    t.loop_stop_h = loop_stop_h; t.h_step = h_step;
    t.row_first = row_first; t.row_last = row_last; t.rowblk = rowblk;
    t.lut = lut_shift >= 0 ? lut : nullptr; t.lut_shift = lut_shift >= 0 ? lut_shift : 0;
    return t;
  }
};

namespace {

void drop_spares(airice_ctx* c) {
  for (auto& b : c->spare) {
    for (cudaEvent_t ev : b.events) cudaEventDestroy(ev);
    cudaFree(b.p);                 // cudaFree waits for the device
  }
  c->spare.clear();
}

// table-sized device buffers: reuse a spare of (nearly) the right size, else allocate
cudaError_t table_alloc(airice_ctx* c, void** p, size_t bytes) {
  // best fit: the column block (384 MB) and the lookup layout (455 MB) of the reference grid fall into each other's
  // window, and handing the larger one to the smaller request leaves the next larger request without a spare
  size_t best = c->spare.size();
  for (size_t i = 0; i < c->spare.size(); i++) {
    if (c->spare[i].bytes >= bytes && c->spare[i].bytes <= bytes + bytes / 4 &&
        (best == c->spare.size() || c->spare[i].bytes < c->spare[best].bytes))
      best = i;
  }
  if (best < c->spare.size()) {
    *p = c->spare[best].p;
    // tables are built on the legacy default stream: it waits for the previous owner's last users
    for (cudaEvent_t ev : c->spare[best].events) { cudaStreamWaitEvent(nullptr, ev, 0); cudaEventDestroy(ev); }
    c->spare.erase(c->spare.begin() + best);
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess && !c->spare.empty()) {      // out of memory: give the spares back and retry
    cudaGetLastError();
    drop_spares(c);
    e = cudaMalloc(p, bytes);
  }
  return e;
}
void table_release(airice_ctx* c, void* p, size_t bytes, const std::vector<cudaStream_t>& used) {
  if (!p) return;
  // keep freed table buffers for the next tables of that size (cudaMalloc/cudaFree of a 400 MB block cost more than
  // building the table): up to a third of the device's memory, 64 antennas' tables (54 GB) included on a B200;
  // table_alloc hands everything back to the driver when an allocation fails
  size_t held = 0;
  for (auto& b : c->spare) held += b.bytes;
  if (c->spare_cap == 0) {
    size_t free_b = 0, total_b = 0;
    c->spare_cap = (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) ? total_b / 3 : ((size_t)2 << 30);
  }
  if (bytes >= ((size_t)1 << 20) && held + bytes <= c->spare_cap && c->spare.size() < 512) {
    // what cudaFree would have waited for, as events instead of a device-wide synchronisation
    airice_ctx::SpareBuf b{p, bytes, {}};
    auto mark = [&](cudaStream_t s) {
      cudaEvent_t ev;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return false;
      if (cudaEventRecord(ev, s) != cudaSuccess) { cudaEventDestroy(ev); return false; }
      b.events.push_back(ev);
      return true;
    };
    bool okm = mark(nullptr);
    for (cudaStream_t s : used) if (s) okm = mark(s) && okm;
    if (!okm) { cudaGetLastError(); cudaDeviceSynchronize(); }
    c->spare.push_back(b);
  } else {
    cudaFree(p);
  }
}

// The reference-layout columns of a table that was built in the lookup layout only.
int ensure_columns(const airice_table* t) {
  if (t->cols[0]) return 0;
  if (!t->ctx || !t->rec) return fail(-1, "table has no data");
  float* block = nullptr;
  const size_t bytes = sizeof(float) * (size_t)t->cells * AIRICE_TABLE_NCOLS32;
  cudaError_t e = table_alloc(t->ctx, (void**)&block, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(table columns)");
  e = launch_unpack_table(t->rec, t->row_h, t->cells, (int)t->n_th, block, t->cells, nullptr);
  if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
  if (e != cudaSuccess) { cudaFree(block); return cuda_fail(e, "unpack table"); }
  for (int k = 0; k < AIRICE_TABLE_NCOLS32; k++) t->cols[k] = block + (int64_t)k * t->cells;
  t->cols_bytes = bytes;
  return 0;
}

// Allocate the lookup layout of a table (dense X, 48-byte records, per-row height and trim ranges: one block).
int pack_alloc(airice_table* t) {
  const size_t rec_bytes = sizeof(float4) * 3 * (size_t)t->cells;
  const size_t x_bytes = (sizeof(float) * (size_t)t->cells + 255) / 256 * 256;
  const size_t rowh_bytes = (sizeof(float) * (size_t)t->n_h + 255) / 256 * 256;
  const size_t range_bytes = (sizeof(int) * (size_t)t->n_h + 255) / 256 * 256;
  const size_t blk_bytes = (sizeof(float) * AIRICE_ROWBLK * (size_t)t->n_h + 255) / 256 * 256;
  const size_t par_bytes = (sizeof(float4) * (size_t)t->n_h + 255) / 256 * 256;
  const size_t lut_bytes = (sizeof(uint16_t) * AIRICE_LUT_EDGES * (size_t)t->n_h + 255) / 256 * 256;
  t->pack_bytes = rec_bytes + x_bytes + rowh_bytes + 2 * range_bytes + blk_bytes + par_bytes + lut_bytes;
  cudaError_t e = table_alloc(t->ctx, &t->pack, t->pack_bytes);
  if (e != cudaSuccess) { t->pack = nullptr; return cuda_fail(e, "cudaMalloc(lookup layout)"); }
  char* base = (char*)t->pack;
  t->rec = (float4*)base;
  t->x = (float*)(base + rec_bytes);
  t->row_h = (float*)(base + rec_bytes + x_bytes);
  t->row_first = (int*)(base + rec_bytes + x_bytes + rowh_bytes);
  t->row_last = (int*)(base + rec_bytes + x_bytes + rowh_bytes + range_bytes);
  t->rowblk = (float*)(base + rec_bytes + x_bytes + rowh_bytes + 2 * range_bytes);
  t->rowpar = (float4*)(base + rec_bytes + x_bytes + rowh_bytes + 2 * range_bytes + blk_bytes);
  t->lut = (uint16_t*)(base + rec_bytes + x_bytes + rowh_bytes + 2 * range_bytes + blk_bytes + par_bytes);
  t->lut_shift = lut_shift_for(t->n_th);
  return 0;
}
// ... and fill it from the column-major form.
int pack_table(airice_table* t) {
  int rc = pack_alloc(t);
  if (rc) return rc;
  cudaError_t e = launch_pack_table(t->cols, t->cells, (int)t->n_h, (int)t->n_th, t->x, t->rec, t->row_h, t->row_first, t->row_last,
                                    t->rowblk, t->rowpar, t->lut, t->lut_shift, nullptr);
  if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
  if (e != cudaSuccess) return cuda_fail(e, "pack table");
  return 0;
}

int ensure_slots(airice_ctx* c, size_t bytes) {
  if (c->slot_bytes >= bytes) return 0;
  for (int s = 0; s < airice_ctx::kSlots; s++) {
    if (c->dev[s]) cudaFree(c->dev[s]);
    c->dev[s] = nullptr;
  }
  c->slot_bytes = 0;
  for (int s = 0; s < airice_ctx::kSlots; s++) {
    CK(cudaMalloc(&c->dev[s], bytes));
    if (!c->streams[s]) CK(cudaStreamCreateWithFlags(&c->streams[s], cudaStreamNonBlocking));
  }
  c->slot_bytes = bytes;
  return 0;
}

int ensure_mapped(airice_ctx* c) {
  if (c->zc_host) return 0;
  const size_t bytes = (size_t)airice_ctx::kSmall * (sizeof(double) * (3 + AIRICE_SOLVE_COLS) + 8);
  CK(cudaHostAlloc(&c->zc_host, bytes, cudaHostAllocMapped));
  cudaError_t e = cudaHostGetDevicePointer(&c->zc_dev, c->zc_host, 0);
  if (e != cudaSuccess) { cudaFreeHost(c->zc_host); c->zc_host = nullptr; return cuda_fail(e, "cudaHostGetDevicePointer"); }
  return 0;
}
bool mapped_small_calls() {
  static const bool on = std::getenv("AIRICE_NO_MAPPED") == nullptr;   // test hook: the chunked copy path for every size
  return on;
}

// Upload per-row transmitter data for rows [r0,r1) and launch kernel 1 on them.
int build_rows(airice_ctx* ctx, const TableGrid& g, int64_t r0, int64_t r1, double* const* cols64, float* const* cols32,
               cudaStream_t s, TableMultiArgs* multi = nullptr) {
  const int64_t rows_avail = g.first_skipped_row < g.n_h ? g.first_skipped_row : g.n_h;
  if (r0 < 0 || r1 > rows_avail || r0 > r1) return fail(-3, "table rows out of range");
  if (r1 == r0) return 0;
  const int64_t nr = r1 - r0;
  const double key[7] = {g.h_top, g.h_step, g.loop_stop_h, (double)g.n_h, ctx->medium.B[0], ctx->medium.C[0], (double)ctx->medium.nlayers};
  airice_ctx::RowCache* hit = nullptr;
  airice_ctx::RowCache* victim = &ctx->rows[0];
  for (airice_ctx::RowCache& q : ctx->rows) {
    if (q.d_rows && q.r0 == r0 && q.r1 == r1 && std::memcmp(q.key, key, sizeof(key)) == 0) hit = &q;
    if (q.used < victim->used) victim = &q;      // empty slots have used == 0
  }
  if (!hit) {
    airice_ctx::RowCache& rc = *victim;
    std::vector<double> h, ntx;
    std::vector<int> kt;
    grid_rows(ctx->medium, g, r0, r1, &h, &ntx, &kt);
    if (rc.d_rows) {
      CK(cudaDeviceSynchronize());  // launches on any stream may still read the evicted arrays
      cudaFree(rc.d_rows);
      if (rc.d_kt) cudaFree(rc.d_kt);
    }
    rc.d_rows = nullptr; rc.d_kt = nullptr; rc.r0 = rc.r1 = -1;
    CK(cudaMalloc((void**)&rc.d_rows, sizeof(double) * 2 * nr));
    CK(cudaMalloc((void**)&rc.d_kt, sizeof(int) * nr));
    CK(cudaMemcpy(rc.d_rows, h.data(), sizeof(double) * nr, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rc.d_rows + nr, ntx.data(), sizeof(double) * nr, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rc.d_kt, kt.data(), sizeof(int) * nr, cudaMemcpyHostToDevice));
    std::memcpy(rc.key, key, sizeof(key));
    rc.r0 = r0; rc.r1 = r1;
    hit = &rc;
  }
  airice_ctx::RowCache& rc = *hit;
  rc.used = ++ctx->row_clock;
  double* d_rows = rc.d_rows;
  int* d_kt = rc.d_kt;
  TableArgs a;
  std::memset(&a, 0, sizeof(a));
  a.cell0 = r0 * g.n_th; a.ncells = nr * g.n_th; a.n_th = g.n_th; a.row0 = r0;
  a.row_h = d_rows; a.row_ntx = d_rows + nr; a.row_kt = d_kt;
  a.th_start = g.th_start; a.th_step = g.th_step; a.th_stop = g.th_stop;
  a.in_ice = g.in_ice;
  if (cols64) for (int k = 0; k < AIRICE_TABLE_NCOLS64; k++) a.c64[k] = cols64[k];
  if (cols32) for (int k = 0; k < AIRICE_TABLE_NCOLS32; k++) a.c32[k] = cols32[k];
  if (cols32) for (int k = 0; k < AIRICE_TABLE_NCOLS32; k++) if (!cols32[k]) return fail(-4, "all 11 float columns are required");
  // the table path hands the surface height of the grid (ice, or ice+depth for a receiver in air) to the walk
  const AirIcePlan& p = ctx->plan(g.ice_h, g.depth_signed);
  cudaError_t e;
  if (multi) { multi->base = a; e = launch_table_multi(ctx->medium, p, *multi, s); }
  else e = launch_table(ctx->medium, p, a, s);
  if (e != cudaSuccess) return cuda_fail(e, "launch_table");
  return 0;
}

// Points a SolveArgs at the stream's deferred-list scratch (grown as needed); without it (allocation failure, >= 2^31
// pairs) launch_solve falls back to the single-pass kernel.
void attach_defer_scratch(airice_ctx* c, cudaStream_t s, SolveArgs* a) {
  if (a->n >= 2147483647LL || a->n < 6000000) return;   // launch_solve runs smaller batches in one pass (kTwoPassMinPairs)
  airice_ctx::DeferScratch& d = c->defer[s];
  if (d.cap < a->n) {
    if (d.buf) { cudaStreamSynchronize(s); cudaFree(d.buf); d.buf = nullptr; d.cap = 0; }
    const int64_t cap = a->n + a->n / 8 + 1024;
    if (cudaMalloc((void**)&d.buf, sizeof(int32_t) * (size_t)(64 + cap)) != cudaSuccess) {
      cudaGetLastError();
      d.buf = nullptr;
      return;
    }
    d.cap = cap;
  }
  a->defer_count = d.buf;
  a->defer_list = d.buf + 64;
}

}  // namespace

extern "C" {

const char* airice_last_error(void) { return g_err.c_str(); }

int airice_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int airice_create(const char* atmosphere_path, int variant, int device, airice_ctx** out) {
  if (!out) return fail(-1, "null argument");
  const bool ice_only = !atmosphere_path || !atmosphere_path[0];   // in-ice entry points only (no GDAS file needed)
  if (variant != AIRICE_VARIANT_MULTIRAY && variant != AIRICE_VARIANT_PYWRAP && variant != AIRICE_VARIANT_CLI)
    return fail(-1, "unknown variant");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(-2, std::string("no CUDA device available (this library has no CPU path): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= ndev) return fail(-2, "device index out of range");
  airice_ctx* c = new airice_ctx();
  c->device = device;
  std::string err;
  if (ice_only) {
    std::memset(&c->medium, 0, sizeof(c->medium));
    c->medium.nlayers = 0; c->medium.variant = variant;
    c->medium.A_ice = 1.78; c->medium.B_ice = -0.43; c->medium.C_ice = 0.0132;
    c->medium.pi = (variant == AIRICE_VARIANT_PYWRAP) ? 4.0 * atan(1.0) : 3.1415927;   // (the CLI copy shares M's pi)
    c->medium.deg2rad = c->medium.pi / 180.0; c->medium.rad2deg = 180.0 / c->medium.pi; c->medium.c = 299792458.0;
    c->no_air = true;
  } else {
    int rc = load_medium(atmosphere_path, variant, &c->medium, &c->n0, &c->npoints, &err);
    if (rc != 0) { delete c; return fail(rc, err); }
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaSetDevice"); }
  if (!ice_only) {
    std::vector<double> tab(2 * AIRICE_CLAMP_N);
    make_clamp_table(c->medium, tab.data());
    e = cudaMalloc((void**)&c->clamp_tab, sizeof(double) * tab.size());
    if (e == cudaSuccess) e = cudaMemcpy(c->clamp_tab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { if (c->clamp_tab) cudaFree(c->clamp_tab); delete c; return cuda_fail(e, "clamp table"); }
    c->medium.clamp_tab = c->clamp_tab;
  }
  *out = c;
  return 0;
}

void airice_destroy(airice_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  // tables that outlive their context: free their device memory now and leave the handles as empty shells that
  // airice_table_destroy can still be called on (every other call on them fails with "null table")
  for (airice_table* t : c->tables) {
    if (t->owns && t->cols[0]) cudaFree(t->cols[0]);
    if (t->pack) cudaFree(t->pack);
    for (auto& col : t->cols) col = nullptr;
    t->pack = nullptr; t->x = nullptr; t->rec = nullptr; t->row_h = nullptr; t->row_first = t->row_last = nullptr;
    t->cells = 0; t->n_h = 0;
    t->ctx = nullptr;
  }
  c->tables.clear();
  for (int s = 0; s < airice_ctx::kSlots; s++) {
    if (c->dev[s]) cudaFree(c->dev[s]);
    if (c->streams[s]) cudaStreamDestroy(c->streams[s]);
  }
  for (airice_ctx::RowCache& q : c->rows) {
    if (q.d_rows) cudaFree(q.d_rows);
    if (q.d_kt) cudaFree(q.d_kt);
  }
  for (int k = 0; k < airice_ctx::kSlots; k++)
    if (c->inice_scratch[k]) cudaFree(c->inice_scratch[k]);
  if (c->inice_cols) cudaFree(c->inice_cols);
  if (c->zc_host) cudaFreeHost(c->zc_host);
  if (c->path_plans) cudaFree(c->path_plans);
  if (c->quad_stats) cudaFree(c->quad_stats);
  if (c->focus_scratch) cudaFree(c->focus_scratch);
  if (c->clamp_tab) cudaFree(c->clamp_tab);
  drop_spares(c);
  for (auto& d : c->defer) if (d.second.buf) cudaFree(d.second.buf);
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  for (int s = 0; s < airice_ctx::kSlots; s++) if (c->join_ev[s]) cudaEventDestroy(c->join_ev[s]);
  delete c;
}

int airice_get_medium(const airice_ctx* c, double out[24]) {
  if (!c || !out) return fail(-1, "null argument");
  const AirIceMedium& m = c->medium;
  out[0] = m.nlayers;
  for (int i = 0; i < 5; i++) { out[1 + i] = m.hlo[i] * 100; out[6 + i] = m.B[i]; out[11 + i] = m.C[i]; }
  out[1 + 4] = 150000 * 100;
  out[16] = m.A_ice; out[17] = m.B_ice; out[18] = m.C_ice; out[19] = m.pi; out[20] = c->n0; out[21] = c->npoints;
  out[22] = 0; out[23] = 0;
  return 0;
}

int airice_set_ice_model(airice_ctx* c, double A, double B, double C) {
  if (!c) return fail(-1, "null context");
  c->medium.A_ice = A; c->medium.B_ice = B; c->medium.C_ice = C;
  c->plans.clear();
  return 0;
}

int airice_table_dims(const airice_ctx* c, double depth_m, double ice_m, double h_top, double h_step, double th_start,
                      double th_step, double th_stop, int64_t* n_h, int64_t* n_th) {
  if (!c) return fail(-1, "null context");
  TableGrid g; std::string err;
  int rc = make_grid(depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop, &g, &err);
  if (rc) return fail(rc, err);
  if (n_h) *n_h = g.first_skipped_row < g.n_h ? g.first_skipped_row : g.n_h;
  if (n_th) *n_th = g.n_th;
  return 0;
}

int airice_table_build_device(airice_ctx* c, double depth_m, double ice_m, double h_top, double h_step,
                              double th_start, double th_step, double th_stop, int64_t row_begin, int64_t row_end,
                              double* const* cols64, float* const* cols32, void* stream) {
  if (!c) return fail(-1, "null context");
  NEED_AIR(c);
  if (!cols64 && !cols32) return fail(-1, "no output columns");
  CK(cudaSetDevice(c->device));
  TableGrid g; std::string err;
  int rc = make_grid(depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop, &g, &err);
  if (rc) return fail(rc, err);
  return build_rows(c, g, row_begin, row_end, cols64, cols32, (cudaStream_t)stream);
}

int airice_table_create(airice_ctx* c, double depth_m, double ice_m, double h_top, double h_step, double th_start,
                        double th_step, double th_stop, airice_table** out) {
  if (!c || !out) return fail(-1, "null argument");
  NEED_AIR(c);
  // a receiver in the ice: the fused pass of airice_table_create_multi (columns + lookup layout in one kernel, same bits)
  if (depth_m < 0) return airice_table_create_multi(c, 1, &depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop, out);
  CK(cudaSetDevice(c->device));
  TableGrid g; std::string err;
  int rc = make_grid(depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop, &g, &err);
  if (rc) return fail(rc, err);
  airice_table* t = new airice_table();
  t->ctx = c; t->owns = true;
  c->tables.push_back(t);
  t->n_h = g.first_skipped_row < g.n_h ? g.first_skipped_row : g.n_h;
  t->n_th = g.n_th; t->cells = t->n_h * t->n_th;
  t->loop_stop_h = g.loop_stop_h; t->h_step = g.h_step;
  if (t->cells >= 2147483647LL) { delete t; return fail(-5, "lookup tables are limited to 2^31-1 cells (the reference indexes them with int)"); }
  float* block = nullptr;
  t->cols_bytes = sizeof(float) * (size_t)t->cells * AIRICE_TABLE_NCOLS32;
  cudaError_t e = table_alloc(c, (void**)&block, t->cols_bytes);
  if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaMalloc(table)"); }
  for (int k = 0; k < AIRICE_TABLE_NCOLS32; k++) t->cols[k] = block + (int64_t)k * t->cells;
  rc = build_rows(c, g, 0, t->n_h, nullptr, t->cols, nullptr);
  if (rc == 0) rc = pack_table(t);
  if (rc) { airice_table_destroy(t); return rc; }
  *out = t;
  return 0;
}

int airice_table_create_multi(airice_ctx* c, int n_ant, const double* depths_m, double ice_m, double h_top, double h_step,
                              double th_start, double th_step, double th_stop, airice_table** out) {
  if (!c || !out || !depths_m || n_ant <= 0) return fail(-1, "null argument");
  NEED_AIR(c);
  for (int q = 0; q < n_ant; q++)
    if (!(depths_m[q] < 0)) return fail(-6, "airice_table_create_multi: every antenna must sit in the ice (depth < 0); "
                                            "a receiver in air changes the surface height of the air walk");
  CK(cudaSetDevice(c->device));
  TableGrid g; std::string err;
  int rc = make_grid(depths_m[0], ice_m, h_top, h_step, th_start, th_step, th_stop, &g, &err);   // same grid for every depth < 0
  if (rc) return fail(rc, err);
  const int64_t n_h = g.first_skipped_row < g.n_h ? g.first_skipped_row : g.n_h;
  const int64_t cells = n_h * g.n_th;
  if (cells >= 2147483647LL) return fail(-5, "lookup tables are limited to 2^31-1 cells (the reference indexes them with int)");
  for (int q = 0; q < n_ant; q++) out[q] = nullptr;
  std::vector<double> ant(2 * (size_t)n_ant);
  double* d_ant = nullptr;
  void** d_blocks = nullptr;   // [3][n_ant]: records, dense X, row heights
  auto cleanup = [&](int code) {
    for (int q = 0; q < n_ant; q++) { if (out[q]) airice_table_destroy(out[q]); out[q] = nullptr; }
    if (d_ant) cudaFree(d_ant);
    if (d_blocks) cudaFree(d_blocks);
    return code;
  };
  for (int q = 0; q < n_ant; q++) {
    airice_table* t = new airice_table();
    out[q] = t;
    t->ctx = c; t->owns = true;
    c->tables.push_back(t);
    t->n_h = n_h; t->n_th = g.n_th; t->cells = cells;
    t->loop_stop_h = g.loop_stop_h; t->h_step = g.h_step;
    ant[2 * q] = -depths_m[q];                       // the plan's positive depth (make_plan)
    ant[2 * q + 1] = n_ice(c->medium, -depths_m[q]);
  }
  // the pass writes the lookup layout of every table and nothing else: the reference-layout columns are made from it
  // on first use (ensure_columns)
  std::vector<void*> ptrs(3 * (size_t)n_ant);
  for (int q = 0; q < n_ant; q++) {
    rc = pack_alloc(out[q]);
    if (rc) return cleanup(rc);
    ptrs[q] = out[q]->rec;
    ptrs[(size_t)n_ant + q] = out[q]->x;
    ptrs[2 * (size_t)n_ant + q] = out[q]->row_h;
  }
  cudaError_t e = cudaMalloc((void**)&d_ant, sizeof(double) * ant.size());
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_blocks, sizeof(void*) * ptrs.size());
  if (e == cudaSuccess) e = cudaMemcpy(d_ant, ant.data(), sizeof(double) * ant.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_blocks, ptrs.data(), sizeof(void*) * ptrs.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cleanup(cuda_fail(e, "antenna arrays"));
  TableMultiArgs ma;
  std::memset(&ma, 0, sizeof(ma));
  ma.n_ant = n_ant; ma.ant = d_ant; ma.blocks = nullptr; ma.col_stride = cells;
  ma.rec = (float4* const*)d_blocks;
  ma.x = (float* const*)(d_blocks + n_ant);
  ma.row_h = (float* const*)(d_blocks + 2 * (size_t)n_ant);
  rc = build_rows(c, g, 0, n_h, nullptr, nullptr, nullptr, &ma);
  for (int q0 = 0; q0 < n_ant && rc == 0; q0 += AIRICE_ROWPREP_MAX) {
    RowPrepBatch b;
    std::memset(&b, 0, sizeof(b));
    b.n_tab = (n_ant - q0 < AIRICE_ROWPREP_MAX) ? (n_ant - q0) : AIRICE_ROWPREP_MAX;
    b.cells = cells; b.n_h = (int)n_h; b.n_th = (int)g.n_th; b.lut_shift = out[q0]->lut_shift;
    for (int k = 0; k < b.n_tab; k++) {
      const airice_table* t = out[q0 + k];
      b.tab[k].x = t->x; b.tab[k].row_h = t->row_h; b.tab[k].row_first = t->row_first; b.tab[k].row_last = t->row_last;
      b.tab[k].rowblk = t->rowblk; b.tab[k].rowpar = t->rowpar; b.tab[k].lut = t->lut;
    }
    e = launch_row_prep(b, nullptr);
    if (e != cudaSuccess) rc = cuda_fail(e, "row ranges");
  }
  if (rc == 0 && (e = cudaStreamSynchronize(nullptr)) != cudaSuccess) rc = cuda_fail(e, "multi-antenna tables");
  if (rc) return cleanup(rc);
  cudaFree(d_ant); cudaFree(d_blocks);
  return 0;
}

int airice_table_wrap(airice_ctx* c, const float* const* d_cols32, int64_t n_h, int64_t n_th, double loop_stop_h,
                      double h_step, airice_table** out) {
  if (!c || !out || !d_cols32) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  if (n_h <= 0 || n_th <= 0) return fail(-5, "table dimensions must be positive");
  if (n_h >= 2147483647LL || n_th >= 2147483647LL || n_h * n_th >= 2147483647LL) return fail(-5, "lookup tables are limited to 2^31-1 cells");
  airice_table* t = new airice_table();
  t->ctx = c; t->owns = false;
  c->tables.push_back(t);
  for (int k = 0; k < AIRICE_TABLE_NCOLS32; k++) t->cols[k] = const_cast<float*>(d_cols32[k]);
  t->n_h = n_h; t->n_th = n_th; t->cells = n_h * n_th; t->loop_stop_h = loop_stop_h; t->h_step = h_step;
  int rc = pack_table(t);
  if (rc) { airice_table_destroy(t); return rc; }
  *out = t;
  return 0;
}

void airice_table_destroy(airice_table* t) {
  if (!t) return;
  if (t->ctx) {                      // else: the context was destroyed first and took the device memory with it
    airice_ctx* c = t->ctx;
    cudaSetDevice(c->device);
    if (t->owns && t->cols[0]) table_release(c, t->cols[0], t->cols_bytes, t->used);
    if (t->pack) table_release(c, t->pack, t->pack_bytes, t->used);
    for (size_t i = 0; i < c->tables.size(); i++)
      if (c->tables[i] == t) { c->tables.erase(c->tables.begin() + i); break; }
  }
  delete t;
}

// ---- table persistence.  File: 64-byte header {"AIRICETB", u32 version = 1, u32 columns = 11, i64 n_h, i64 n_th,
// f64 loop_stop_h, f64 h_step, u64 FNV-1a of the payload, 8 reserved bytes}, then the 11 float columns in
// AllTableAllAntData order, column-major, little-endian.
namespace {
struct TableFileHeader {
  char magic[8];
  uint32_t version, ncols;
  int64_t n_h, n_th;
  double loop_stop_h, h_step;
  uint64_t checksum;
  uint64_t reserved;
};
static_assert(sizeof(TableFileHeader) == 64, "header layout");
uint64_t fnv1a(const void* p, size_t n) {
  const unsigned char* b = (const unsigned char*)p;
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
}  // namespace

int airice_table_save(const airice_table* t, const char* path) {
  if (!t || !t->ctx || !path) return fail(-1, "null argument");
  CK(cudaSetDevice(t->ctx->device));
  if (int rc = ensure_columns(t)) return rc;
  const size_t n = (size_t)t->cells * AIRICE_TABLE_NCOLS32;
  std::vector<float> host(n);
  for (int k = 0; k < AIRICE_TABLE_NCOLS32; k++)
    CK(cudaMemcpy(host.data() + (size_t)k * t->cells, t->cols[k], sizeof(float) * (size_t)t->cells, cudaMemcpyDeviceToHost));
  TableFileHeader hd;
  std::memset(&hd, 0, sizeof(hd));
  std::memcpy(hd.magic, "AIRICETB", 8);
  hd.version = 1; hd.ncols = AIRICE_TABLE_NCOLS32; hd.n_h = t->n_h; hd.n_th = t->n_th;
  hd.loop_stop_h = t->loop_stop_h; hd.h_step = t->h_step;
  hd.checksum = fnv1a(host.data(), n * sizeof(float));
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(-7, std::string("cannot open for writing: ") + path);
  const bool okw = std::fwrite(&hd, sizeof(hd), 1, f) == 1 && std::fwrite(host.data(), sizeof(float), n, f) == n;
  if (std::fclose(f) != 0 || !okw) return fail(-7, std::string("short write: ") + path);
  return 0;
}

int airice_table_load(airice_ctx* c, const char* path, airice_table** out) {
  if (!c || !path || !out) return fail(-1, "null argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(-7, std::string("cannot open: ") + path);
  TableFileHeader hd;
  if (std::fread(&hd, sizeof(hd), 1, f) != 1) { std::fclose(f); return fail(-8, "table file: truncated header"); }
  if (std::memcmp(hd.magic, "AIRICETB", 8) != 0) { std::fclose(f); return fail(-8, "table file: bad magic"); }
  if (hd.version != 1 || hd.ncols != AIRICE_TABLE_NCOLS32) { std::fclose(f); return fail(-8, "table file: unsupported version / column count"); }
  // each factor bounded first: the product of two untrusted 64-bit values must not overflow before it is tested
  if (hd.n_h <= 0 || hd.n_th <= 0 || hd.n_h >= 2147483647LL || hd.n_th >= 2147483647LL || hd.n_h * hd.n_th >= 2147483647LL) {
    std::fclose(f);
    return fail(-8, "table file: bad dimensions");
  }
  const size_t cells = (size_t)(hd.n_h * hd.n_th), n = cells * AIRICE_TABLE_NCOLS32;
  std::vector<float> host(n);
  const size_t got = std::fread(host.data(), sizeof(float), n, f);
  const bool extra = std::fgetc(f) != EOF;
  std::fclose(f);
  if (got != n || extra) return fail(-8, "table file: payload size does not match the header");
  if (fnv1a(host.data(), n * sizeof(float)) != hd.checksum) return fail(-8, "table file: checksum mismatch");
  CK(cudaSetDevice(c->device));
  airice_table* t = new airice_table();
  t->ctx = c; t->owns = true;
  c->tables.push_back(t);
  t->n_h = hd.n_h; t->n_th = hd.n_th; t->cells = (int64_t)cells; t->loop_stop_h = hd.loop_stop_h; t->h_step = hd.h_step;
  float* block = nullptr;
  t->cols_bytes = sizeof(float) * n;
  cudaError_t e = table_alloc(c, (void**)&block, t->cols_bytes);
  if (e != cudaSuccess) { airice_table_destroy(t); return cuda_fail(e, "cudaMalloc(table)"); }
  for (int k = 0; k < AIRICE_TABLE_NCOLS32; k++) t->cols[k] = block + (int64_t)k * t->cells;
  e = cudaMemcpyAsync(block, host.data(), t->cols_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e != cudaSuccess) { airice_table_destroy(t); return cuda_fail(e, "upload table"); }
  int rc = pack_table(t);          // synchronises the stream: `host` may go
  if (rc) { airice_table_destroy(t); return rc; }
  *out = t;
  return 0;
}

int airice_table_info(const airice_table* t, int64_t info[4]) {
  if (!t || !t->ctx) return fail(-1, "null table (or its context was destroyed)");
  info[0] = t->n_h; info[1] = t->n_th; info[2] = t->cells; info[3] = AIRICE_TABLE_NCOLS32;
  return 0;
}

int airice_table_copy_column(const airice_table* t, int col, float* host_out) {
  if (!t || !t->ctx || col < 0 || col >= AIRICE_TABLE_NCOLS32) return fail(-1, "bad table/column");
  CK(cudaSetDevice(t->ctx->device));
  if (int rc = ensure_columns(t)) return rc;
  CK(cudaMemcpy(host_out, t->cols[col], sizeof(float) * t->cells, cudaMemcpyDeviceToHost));
  return 0;
}

int airice_table_column_ptr(const airice_table* t, int col, const float** d_ptr) {
  if (!t || !t->ctx || col < 0 || col >= AIRICE_TABLE_NCOLS32) return fail(-1, "bad table/column");
  CK(cudaSetDevice(t->ctx->device));
  if (int rc = ensure_columns(t)) return rc;
  *d_ptr = t->cols[col];
  return 0;
}

int airice_table_copy_row_ranges(const airice_table* t, int32_t* host_first, int32_t* host_last) {
  if (!t || !t->ctx) return fail(-1, "null table");
  CK(cudaSetDevice(t->ctx->device));
  CK(cudaMemcpy(host_first, t->row_first, sizeof(int) * t->n_h, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(host_last, t->row_last, sizeof(int) * t->n_h, cudaMemcpyDeviceToHost));
  return 0;
}

int airice_forward_device(airice_ctx* c, int64_t n, const double* d_theta, const double* d_h, double depth_m,
                          double ice_m, double* const* cols64, void* stream) {
  if (!c || !cols64) return fail(-1, "null argument");
  NEED_AIR(c);
  CK(cudaSetDevice(c->device));
  // GetRayTracingSolutions takes the surface height as given (no depth fold); depth>=0 just means "no ice leg"
  const int in_ice = depth_m < 0 ? 1 : 0;
  const AirIcePlan& p = c->plan(ice_m, in_ice ? depth_m : 0.0);
  ForwardArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.theta = d_theta; a.h = d_h; a.in_ice = in_ice;
  for (int k = 0; k < AIRICE_TABLE_NCOLS64; k++) a.c64[k] = cols64[k];
  cudaError_t e = launch_forward(c->medium, p, a, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_forward");
  return 0;
}

int airice_forward_host(airice_ctx* c, int64_t n, const double* theta, const double* h, double depth_m, double ice_m,
                        double* out) {
  if (!c) return fail(-1, "null context");
  NEED_AIR(c);
  if (n == 0) return 0;
  if (!theta || !h || !out) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int nc = AIRICE_TABLE_NCOLS64;
  const int64_t chunk = n < (1 << 20) ? n : (1 << 20);
  int rc = ensure_slots(c, (size_t)chunk * sizeof(double) * (2 + nc) + 64);
  if (rc) return rc;
  const int in_ice = depth_m < 0 ? 1 : 0;
  const AirIcePlan& p = c->plan(ice_m, in_ice ? depth_m : 0.0);
  int slot = 0;
  for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    double* dh = (double*)c->dev[slot];
    cudaStream_t s = c->streams[slot];
    CK(cudaMemcpyAsync(dh, theta + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, h + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    ForwardArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = m; a.theta = dh; a.h = dh + chunk; a.in_ice = in_ice;
    for (int k = 0; k < nc; k++) a.c64[k] = dh + (2 + k) * chunk;
    cudaError_t e = launch_forward(c->medium, p, a, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_forward");
    for (int k = 0; k < nc; k++)
      CK(cudaMemcpyAsync(out + (int64_t)k * n + off, a.c64[k], sizeof(double) * m, cudaMemcpyDeviceToHost, s));
  }
  for (int s = 0; s < airice_ctx::kSlots; s++)
    if (c->streams[s]) CK(cudaStreamSynchronize(c->streams[s]));
  return 0;
}

int airice_solve_device(airice_ctx* c, int64_t n, const double* d_h, const double* d_dist, const double* d_straight,
                        double depth, double ice, int units, double* const* d_out, uint8_t* d_ok, int32_t* d_nevals,
                        void* stream) {
  if (!c || !d_out) return fail(-1, "null argument");
  NEED_AIR(c);
  if (units != AIRICE_UNITS_M_DEG && units != AIRICE_UNITS_CM_RAD) return fail(-1, "unknown units");
  if (n == 0) return 0;
  CK(cudaSetDevice(c->device));
  const double sc = (units == AIRICE_UNITS_CM_RAD) ? 100.0 : 1.0;
  const AirIcePlan& p = c->plan(ice / sc, depth / sc);  // same "/100" the reference applies (M.cc:949-950)
  SolveArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.h = d_h; a.d = d_dist; a.straight = d_straight; a.ice = ice; a.depth = depth; a.units = units;
  const int nc = units == AIRICE_UNITS_CM_RAD ? AIRICE_SOLVE_COLS_CM_RAD : AIRICE_SOLVE_COLS;
  for (int k = 0; k < nc; k++) a.out[k] = d_out[k];
  a.ok = d_ok; a.nevals = d_nevals;
  attach_defer_scratch(c, (cudaStream_t)stream, &a);
  cudaError_t e = launch_solve(c->medium, p, a, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_solve");
  return 0;
}

// Multi-antenna form (BASELINE config 5: shower points x in-ice receivers): one launch per receiver depth over the
// same Tx heights; distances and outputs are antenna-major ([antenna][point]).
int airice_solve_multi_device(airice_ctx* c, int64_t n_points, int n_ant, const double* d_h, const double* d_dist,
                              const double* depths_host, double ice, int units, double* const* d_out, uint8_t* d_ok,
                              void* stream) {
  if (!c) return fail(-1, "null context");
  if (n_ant < 0 || n_points < 0) return fail(-1, "negative size");
  if (n_ant == 0 || n_points == 0) return 0;
  if (!d_h || !d_dist || !depths_host || !d_out) return fail(-1, "null argument");
  const int nc = units == AIRICE_UNITS_CM_RAD ? AIRICE_SOLVE_COLS_CM_RAD : AIRICE_SOLVE_COLS;
  // The per-antenna launches alternate between the context's two streams, forked from and joined to the caller's
  // stream by events: a launch of 1e6 pairs is 6.6 waves of CTAs, and the next antenna's CTAs fill its tail.
  CK(cudaSetDevice(c->device));
  const bool fork = n_ant > 1;
  cudaStream_t user = (cudaStream_t)stream;
  if (fork) {
    for (int s = 0; s < airice_ctx::kSlots; s++)
      if (!c->streams[s]) CK(cudaStreamCreateWithFlags(&c->streams[s], cudaStreamNonBlocking));
    if (!c->fork_ev) CK(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    for (int s = 0; s < airice_ctx::kSlots; s++)
      if (!c->join_ev[s]) CK(cudaEventCreateWithFlags(&c->join_ev[s], cudaEventDisableTiming));
    CK(cudaEventRecord(c->fork_ev, user));
    for (int s = 0; s < airice_ctx::kSlots; s++) CK(cudaStreamWaitEvent(c->streams[s], c->fork_ev, 0));
  }
  int rc = 0;
  for (int a = 0; a < n_ant && rc == 0; a++) {
    double* cols[AIRICE_SOLVE_NCOLS];
    for (int k = 0; k < nc; k++) cols[k] = d_out[k] ? d_out[k] + (int64_t)a * n_points : nullptr;
    rc = airice_solve_device(c, n_points, d_h, d_dist + (int64_t)a * n_points, nullptr, depths_host[a], ice, units, cols,
                             d_ok ? d_ok + (int64_t)a * n_points : nullptr, nullptr,
                             fork ? (void*)c->streams[a % airice_ctx::kSlots] : stream);
  }
  if (fork) {   // join even after an error: the caller's stream must not run ahead of what was launched
    for (int s = 0; s < airice_ctx::kSlots; s++) {
      cudaEventRecord(c->join_ev[s], c->streams[s]);
      cudaStreamWaitEvent(user, c->join_ev[s], 0);
    }
  }
  return rc;
}

// Host-buffer path: the batch is cut into chunks that alternate between two streams, each with its own device
// staging slot.  Copies go directly from/to the caller's buffers (no extra host memcpy); with pinned caller memory
// the H2D of chunk k+1, the kernel of chunk k and the D2H of chunk k-1 run concurrently.
}  // extern "C"
namespace {
// Host-buffer solve.  Either `out` (a dense [nc][n] block, every column wanted) or `cols` (nc host pointers, NULL = the
// caller does not use that column: it is neither stored by the kernel nor copied back) is given; `ok` may be NULL with `cols`.
int solve_host_impl(airice_ctx* c, int64_t n, const double* h, const double* dist, const double* straight, double depth,
                    double ice, int units, double* out, double* const* cols, uint8_t* ok) {
  if (!c) return fail(-1, "null context");
  NEED_AIR(c);
  if (units != AIRICE_UNITS_M_DEG && units != AIRICE_UNITS_CM_RAD) return fail(-1, "unknown units");
  if (n == 0) return 0;
  if (!h || !dist || (!out && !cols) || (out && !ok)) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int nc = units == AIRICE_UNITS_CM_RAD ? AIRICE_SOLVE_COLS_CM_RAD : AIRICE_SOLVE_COLS;
  if (n <= airice_ctx::kSmall && mapped_small_calls()) {
    if (int rc = ensure_slots(c, 256)) return rc;      // makes the streams
    if (int rc = ensure_mapped(c)) return rc;
    const int cap = airice_ctx::kSmall;
    double* hh = (double*)c->zc_host;
    double* dd = (double*)c->zc_dev;
    std::memcpy(hh, h, sizeof(double) * n);
    std::memcpy(hh + cap, dist, sizeof(double) * n);
    if (straight) std::memcpy(hh + 2 * cap, straight, sizeof(double) * n);
    const double sc1 = (units == AIRICE_UNITS_CM_RAD) ? 100.0 : 1.0;
    const AirIcePlan& p1 = c->plan(ice / sc1, depth / sc1);
    SolveArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = n; a.h = dd; a.d = dd + cap; a.ice = ice; a.depth = depth; a.units = units;
    if (straight) a.straight = dd + 2 * cap;
    for (int k = 0; k < nc; k++) a.out[k] = (out || cols[k]) ? dd + (3 + k) * cap : nullptr;
    a.ok = ok ? (uint8_t*)(dd + (3 + AIRICE_SOLVE_COLS) * cap) : nullptr;
    cudaStream_t s = c->streams[0];
    cudaError_t e = launch_solve(c->medium, p1, a, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_solve");
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < nc; k++) {
      double* dst = out ? out + (size_t)k * n : cols[k];
      if (dst) std::memcpy(dst, hh + (3 + k) * cap, sizeof(double) * n);
    }
    if (ok) std::memcpy(ok, (const uint8_t*)(hh + (3 + AIRICE_SOLVE_COLS) * cap), (size_t)n);
    return 0;
  }
  // pairs per pipeline chunk (H2D -> kernel -> D2H on alternating streams).  The first chunk's upload and kernel are the
  // only part the D2H link waits for: measured 14.0 / 13.8 / 13.7 / 13.6 ms per 1e7 pairs with 2M / 1M / 512K / 256K
  // pairs per chunk; AIRICE_HOST_CHUNK overrides
  static const int64_t kChunk = [] { const char* e = std::getenv("AIRICE_HOST_CHUNK"); const long v = e ? std::atol(e) : 0; return (int64_t)(v >= 4096 ? v : (1 << 19)); }();
  const int64_t chunk = n < kChunk ? (n > 0 ? n : 1) : kChunk;
  int rc = ensure_slots(c, (size_t)chunk * (sizeof(double) * (3 + nc) + 1) + 64);
  if (rc) return rc;
  const double sc = (units == AIRICE_UNITS_CM_RAD) ? 100.0 : 1.0;
  const AirIcePlan& p = c->plan(ice / sc, depth / sc);
  int slot = 0;
  for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    double* dh = (double*)c->dev[slot];
    cudaStream_t s = c->streams[slot];
    CK(cudaMemcpyAsync(dh, h + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, dist + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    SolveArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = m; a.h = dh; a.d = dh + chunk; a.ice = ice; a.depth = depth; a.units = units;
    for (int k = 0; k < nc; k++) a.out[k] = (out || cols[k]) ? dh + (2 + k) * chunk : nullptr;
    a.ok = ok ? (uint8_t*)(dh + (3 + nc) * chunk) : nullptr;
    if (straight) {
      double* ds = dh + (2 + nc) * chunk;
      CK(cudaMemcpyAsync(ds, straight + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
      a.straight = ds;
    }
    attach_defer_scratch(c, s, &a);
    cudaError_t e = launch_solve(c->medium, p, a, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_solve");
    if (out) {
      // the nc columns of the chunk in ONE strided copy (device pitch = chunk, host pitch = n) instead of nc copies
      CK(cudaMemcpy2DAsync(out + off, sizeof(double) * (size_t)n, a.out[0], sizeof(double) * (size_t)chunk, sizeof(double) * (size_t)m,
                           (size_t)nc, cudaMemcpyDeviceToHost, s));
    } else {
      for (int k = 0; k < nc; k++)
        if (cols[k]) CK(cudaMemcpyAsync(cols[k] + off, a.out[k], sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, s));
    }
    if (ok) CK(cudaMemcpyAsync(ok + off, a.ok, (size_t)m, cudaMemcpyDeviceToHost, s));
  }
  for (int s = 0; s < airice_ctx::kSlots; s++)
    if (c->streams[s]) CK(cudaStreamSynchronize(c->streams[s]));
  return 0;
}
}  // namespace
extern "C" {

int airice_solve_host(airice_ctx* c, int64_t n, const double* h, const double* dist, const double* straight,
                      double depth, double ice, int units, double* out, uint8_t* ok) {
  if (n != 0 && !out) return fail(-1, "null argument");
  return solve_host_impl(c, n, h, dist, straight, depth, ice, units, out, nullptr, ok);
}

int airice_solve_host_columns(airice_ctx* c, int64_t n, const double* h, const double* dist, const double* straight,
                              double depth, double ice, int units, double* const* cols, uint8_t* ok) {
  if (n != 0 && !cols) return fail(-1, "null argument");
  return solve_host_impl(c, n, h, dist, straight, depth, ice, units, nullptr, cols, ok);
}

static int lookup_literal() {
  const char* v = std::getenv("AIRICE_LOOKUP_LITERAL");
  return (v && v[0] == '1') ? 1 : 0;
}

int airice_lookup_device(airice_ctx* c, const airice_table* t, int64_t n, const double* d_h_cm, const double* d_dist_cm,
                         double* const* d_out, uint8_t* d_ok, void* stream) {
  if (!c || !t) return fail(-1, "null argument");
  if (n == 0) return 0;
  if (!d_out || !d_ok) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  LookupArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.h_cm = d_h_cm; a.d_cm = d_dist_cm; a.ok = d_ok;
  for (int k = 0; k < AIRICE_LOOKUP_NCOLS; k++) a.out[k] = d_out[k];
  a.literal = lookup_literal();
  if (t->ctx != c) return fail(-1, "table belongs to another (or a destroyed) context");
  t->note_stream((cudaStream_t)stream);
  cudaError_t e = launch_lookup(c->medium, t->view(), a, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_lookup");
  return 0;
}

}  // extern "C"
namespace {
// Host-buffer lookup: `out` (dense [9][n] block) or `cols` (9 host pointers, NULL = not wanted: not stored, not copied
// back; `ok` may then be NULL too).
int lookup_host_impl(airice_ctx* c, const airice_table* t, int64_t n, const double* h_cm, const double* dist_cm, double* out,
                     double* const* cols, uint8_t* ok) {
  if (!c || !t) return fail(-1, "null argument");
  if (n == 0) return 0;
  if (!h_cm || !dist_cm || (!out && !cols) || (out && !ok)) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int nc = AIRICE_LOOKUP_NCOLS;
  if (n <= airice_ctx::kSmall && mapped_small_calls()) {
    if (int rc = ensure_slots(c, 256)) return rc;      // makes the streams
    if (int rc = ensure_mapped(c)) return rc;
    const int cap = airice_ctx::kSmall;
    double* hh = (double*)c->zc_host;
    double* dd = (double*)c->zc_dev;
    std::memcpy(hh, h_cm, sizeof(double) * n);
    std::memcpy(hh + cap, dist_cm, sizeof(double) * n);
    LookupArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = n; a.h_cm = dd; a.d_cm = dd + cap; a.ok = (uint8_t*)(dd + (2 + nc) * cap);
    for (int k = 0; k < nc; k++) a.out[k] = (out || cols[k]) ? dd + (2 + k) * cap : nullptr;
    a.literal = lookup_literal();
    cudaStream_t s = c->streams[0];
    t->note_stream(s);
    cudaError_t e = launch_lookup(c->medium, t->view(), a, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_lookup");
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < nc; k++) {
      double* dst = out ? out + (size_t)k * n : cols[k];
      if (dst) std::memcpy(dst, hh + (2 + k) * cap, sizeof(double) * n);
    }
    if (ok) std::memcpy(ok, (const uint8_t*)(hh + (2 + nc) * cap), (size_t)n);
    return 0;
  }
  const int64_t chunk = n < (1 << 20) ? (n > 0 ? n : 1) : (1 << 20);
  int rc = ensure_slots(c, (size_t)chunk * (sizeof(double) * (2 + nc) + 1) + 64);
  if (rc) return rc;
  int slot = 0;
  for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    double* dh = (double*)c->dev[slot];
    cudaStream_t s = c->streams[slot];
    CK(cudaMemcpyAsync(dh, h_cm + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, dist_cm + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    LookupArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = m; a.h_cm = dh; a.d_cm = dh + chunk; a.ok = (uint8_t*)(dh + (2 + nc) * chunk);
    for (int k = 0; k < nc; k++) a.out[k] = (out || cols[k]) ? dh + (2 + k) * chunk : nullptr;
    a.literal = lookup_literal();
    t->note_stream(s);
    cudaError_t e = launch_lookup(c->medium, t->view(), a, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_lookup");
    if (out) {
      // the nc columns of the chunk in ONE strided copy (device pitch = chunk, host pitch = n) instead of nc copies
      CK(cudaMemcpy2DAsync(out + off, sizeof(double) * (size_t)n, a.out[0], sizeof(double) * (size_t)chunk, sizeof(double) * (size_t)m,
                           (size_t)nc, cudaMemcpyDeviceToHost, s));
    } else {
      for (int k = 0; k < nc; k++)
        if (cols[k]) CK(cudaMemcpyAsync(cols[k] + off, a.out[k], sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, s));
    }
    if (ok) CK(cudaMemcpyAsync(ok + off, a.ok, (size_t)m, cudaMemcpyDeviceToHost, s));
  }
  for (int s = 0; s < airice_ctx::kSlots; s++)
    if (c->streams[s]) CK(cudaStreamSynchronize(c->streams[s]));
  return 0;
}
}  // namespace
extern "C" {

int airice_lookup_host(airice_ctx* c, const airice_table* t, int64_t n, const double* h_cm, const double* dist_cm,
                       double* out, uint8_t* ok) {
  if (n != 0 && !out) return fail(-1, "null argument");
  return lookup_host_impl(c, t, n, h_cm, dist_cm, out, nullptr, ok);
}

int airice_lookup_host_columns(airice_ctx* c, const airice_table* t, int64_t n, const double* h_cm, const double* dist_cm,
                               double* const* cols, uint8_t* ok) {
  if (n != 0 && !cols) return fail(-1, "null argument");
  return lookup_host_impl(c, t, n, h_cm, dist_cm, nullptr, cols, ok);
}

namespace {
// scratch layout: [counters:int32 x2 pad to 256][list:int32 x n][mask:u8 x n][L_R:f64 x n][ladder:f64 x 6n]
int inice_scratch(airice_ctx* c, int slot, int64_t n, InIceArgs* a, cudaStream_t s) {
  const size_t list_b = ((size_t)n * 4 + 255) / 256 * 256, mask_b = ((size_t)n + 255) / 256 * 256, lr_b = (size_t)n * 8;
  const size_t need = 256 + list_b + mask_b + lr_b + 6 * lr_b;
  if (c->inice_bytes[slot] < need) {
    CK(cudaStreamSynchronize(s));
    if (c->inice_scratch[slot]) cudaFree(c->inice_scratch[slot]);
    c->inice_scratch[slot] = nullptr; c->inice_bytes[slot] = 0;
    CK(cudaMalloc(&c->inice_scratch[slot], need));
    c->inice_bytes[slot] = need;
  }
  char* base = (char*)c->inice_scratch[slot];
  a->ra_count = (int32_t*)base;
  a->ra_list = (int32_t*)(base + 256);
  if (!a->mask) a->mask = (uint8_t*)(base + 256 + list_b);
  if (!a->out[20]) a->out[20] = (double*)(base + 256 + list_b + mask_b);
  a->ra_lad = (double*)(base + 256 + list_b + mask_b + lr_b);
  return 0;
}
}  // namespace

int airice_inice_solve_device(airice_ctx* c, int64_t n, const double* d_z0, const double* d_x1, const double* d_z1,
                              double* const* d_out, uint8_t* d_mask, void* stream) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!d_z0 || !d_x1 || !d_z1 || !d_out) return fail(-1, "null argument");
  if (n >= 2147483647LL) return fail(-5, "at most 2^31-1 pairs per call");
  CK(cudaSetDevice(c->device));
  InIceArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.z0 = d_z0; a.x1 = d_x1; a.z1 = d_z1;
  a.A = c->medium.A_ice; a.B = c->medium.B_ice; a.C = c->medium.C_ice;
  for (int k = 0; k < AIRICE_INICE_NCOLS; k++) a.out[k] = d_out[k];
  a.mask = d_mask;
  { int rc = inice_scratch(c, 0, n, &a, (cudaStream_t)stream); if (rc) return rc; }
  cudaError_t e = launch_inice(a, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_inice");
  return 0;
}

int airice_inice_solve_host(airice_ctx* c, int64_t n, const double* z0, const double* x1, const double* z1, double* out,
                            uint8_t* mask) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!z0 || !x1 || !z1 || !out || !mask) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int nc = AIRICE_INICE_NCOLS;
  const int64_t chunk = n < (1 << 19) ? n : (1 << 19);
  int rc = ensure_slots(c, (size_t)chunk * (sizeof(double) * (3 + nc) + 1) + 64);
  if (rc) return rc;
  int slot = 0;
  for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    double* dh = (double*)c->dev[slot];
    cudaStream_t s = c->streams[slot];
    CK(cudaMemcpyAsync(dh, z0 + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, x1 + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + 2 * chunk, z1 + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    InIceArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = m; a.z0 = dh; a.x1 = dh + chunk; a.z1 = dh + 2 * chunk;
    a.A = c->medium.A_ice; a.B = c->medium.B_ice; a.C = c->medium.C_ice;
    for (int k = 0; k < nc; k++) a.out[k] = dh + (3 + k) * chunk;
    a.mask = (uint8_t*)(dh + (3 + nc) * chunk);
    { int rcs = inice_scratch(c, slot, chunk, &a, s); if (rcs) return rcs; }   // per-slot scratch: chunks overlap
    cudaError_t e = launch_inice(a, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_inice");
    for (int k = 0; k < nc; k++)
      CK(cudaMemcpyAsync(out + (int64_t)k * n + off, a.out[k], sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(mask + off, a.mask, (size_t)m, cudaMemcpyDeviceToHost, s));
  }
  for (int s = 0; s < airice_ctx::kSlots; s++)
    if (c->streams[s]) CK(cudaStreamSynchronize(c->streams[s]));
  return 0;
}

namespace {
constexpr int64_t kRaysChunk = 1 << 20;

// solve + pick for one chunk of DEVICE inputs on stream s; the 29 intermediate columns live in the context
struct AttSpec { bool on = false; double A0 = 0, frequency = 0; double* att[2] = {nullptr, nullptr}; };

// the context's quadrature statistics: [0] integrals that ran out of interval storage, [1] largest interval count (> 24)
int quad_stats(airice_ctx* c, int32_t** out) {
  if (!c->quad_stats) {
    CK(cudaMalloc((void**)&c->quad_stats, 2 * sizeof(int32_t)));
    CK(cudaMemset(c->quad_stats, 0, 2 * sizeof(int32_t)));
  }
  *out = c->quad_stats;
  return 0;
}

int two_rays_chunk(airice_ctx* c, int64_t m, const double* d_rx, const double* d_dist, const double* d_tx, double* const* out10,
                   int32_t* const* ignore2, int32_t* const* type2, cudaStream_t s, const AttSpec* att = nullptr) {
  if (c->inice_cols_n < m) {
    CK(cudaStreamSynchronize(s));
    if (c->inice_cols) cudaFree(c->inice_cols);
    c->inice_cols = nullptr; c->inice_cols_n = 0;
    const int64_t cap = m < kRaysChunk ? m : kRaysChunk;
    CK(cudaMalloc((void**)&c->inice_cols, sizeof(double) * AIRICE_INICE_NCOLS * (size_t)cap));
    c->inice_cols_n = cap;
  }
  InIceArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = m; a.z0 = d_tx; a.x1 = d_dist; a.z1 = d_rx;     // IceRayTracing(0, TxDepth, Distance, RxDepth), IceRayTracing.cc:2917
  a.A = c->medium.A_ice; a.B = c->medium.B_ice; a.C = c->medium.C_ice;
  for (int k = 0; k < AIRICE_INICE_NCOLS; k++) a.out[k] = c->inice_cols + (size_t)k * c->inice_cols_n;
  { int rc = inice_scratch(c, 0, m, &a, s); if (rc) return rc; }
  cudaError_t e = launch_inice(a, s);
  if (e != cudaSuccess) return cuda_fail(e, "launch_inice");
  InIcePickArgs p;
  std::memset(&p, 0, sizeof(p));
  p.n = m; p.rx_depth = d_rx; p.distance = d_dist; p.tx_depth = d_tx;
  p.A = a.A; p.B = a.B; p.C = a.C;
  for (int k = 0; k < AIRICE_INICE_NCOLS; k++) p.in[k] = a.out[k];
  for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++) p.out[k] = out10[k];
  p.ignore[0] = ignore2[0]; p.ignore[1] = ignore2[1];
  p.type[0] = type2 ? type2[0] : nullptr; p.type[1] = type2 ? type2[1] : nullptr;
  if (att && att->on) {
    p.att[0] = att->att[0]; p.att[1] = att->att[1];
    p.A0 = att->A0; p.frequency = att->frequency;
    p.w0 = log(0.0001); p.w2 = log(3.16); p.w = log(att->frequency);     // IceRayTracing.cc:146: libm, like the reference
    { int rc = quad_stats(c, &p.quad_stats); if (rc) return rc; }
  }
  e = launch_inice_pick(p, s);
  if (e != cudaSuccess) return cuda_fail(e, "launch_inice_pick");
  return 0;
}
}  // namespace

int airice_inice_two_rays_device(airice_ctx* c, int64_t n, const double* d_rx_depth, const double* d_distance,
                                 const double* d_tx_depth, double* const* d_out, int32_t* const* d_ignore,
                                 int32_t* const* d_type, void* stream) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!d_rx_depth || !d_distance || !d_tx_depth || !d_out || !d_ignore || !d_ignore[0] || !d_ignore[1])
    return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  for (int64_t off = 0; off < n; off += kRaysChunk) {
    const int64_t m = (n - off < kRaysChunk) ? (n - off) : kRaysChunk;
    double* o[AIRICE_INICE_RAYS_NCOLS];
    for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++) o[k] = d_out[k] ? d_out[k] + off : nullptr;
    int32_t* ig[2] = {d_ignore[0] + off, d_ignore[1] + off};
    int32_t* ty[2] = {d_type && d_type[0] ? d_type[0] + off : nullptr, d_type && d_type[1] ? d_type[1] + off : nullptr};
    const int rc = two_rays_chunk(c, m, d_rx_depth + off, d_distance + off, d_tx_depth + off, o, ig, ty, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int airice_inice_two_rays_host(airice_ctx* c, int64_t n, const double* rx_depth, const double* distance, const double* tx_depth,
                               double* out, int32_t* ignore) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!rx_depth || !distance || !tx_depth || !out || !ignore) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int nc = AIRICE_INICE_RAYS_NCOLS;
  const int64_t chunk = n < kRaysChunk ? n : kRaysChunk;
  int rc = ensure_slots(c, (size_t)chunk * (sizeof(double) * (3 + nc) + 2 * sizeof(int32_t)) + 64);
  if (rc) return rc;
  cudaStream_t s = c->streams[0];             // one stream: the 29-column intermediate is shared between chunks
  double* dh = (double*)c->dev[0];
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    CK(cudaMemcpyAsync(dh, rx_depth + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, distance + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + 2 * chunk, tx_depth + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    double* o[nc];
    for (int k = 0; k < nc; k++) o[k] = dh + (3 + k) * chunk;
    int32_t* ig0 = (int32_t*)(dh + (3 + nc) * chunk);
    int32_t* ig[2] = {ig0, ig0 + chunk};
    rc = two_rays_chunk(c, m, dh, dh + chunk, dh + 2 * chunk, o, ig, nullptr, s);
    if (rc) return rc;
    for (int k = 0; k < nc; k++)
      CK(cudaMemcpyAsync(out + (int64_t)k * n + off, o[k], sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    for (int k = 0; k < 2; k++)
      CK(cudaMemcpyAsync(ignore + (int64_t)k * n + off, ig[k], sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}

// ---- attenuation, focusing, in-ice table
int airice_inice_two_rays_att_device(airice_ctx* c, int64_t n, const double* d_rx_depth, const double* d_distance,
                                     const double* d_tx_depth, double A0, double frequency_ghz, double* const* d_out,
                                     double* const* d_att, int32_t* const* d_ignore, int32_t* const* d_type, void* stream) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!d_rx_depth || !d_distance || !d_tx_depth || !d_out || !d_att || !d_att[0] || !d_att[1] || !d_ignore || !d_ignore[0] ||
      !d_ignore[1])
    return fail(-1, "null argument");
  if (!(frequency_ghz > 0)) return fail(-3, "frequency must be positive (GHz)");
  CK(cudaSetDevice(c->device));
  for (int64_t off = 0; off < n; off += kRaysChunk) {
    const int64_t m = (n - off < kRaysChunk) ? (n - off) : kRaysChunk;
    double* o[AIRICE_INICE_RAYS_NCOLS];
    for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++) o[k] = d_out[k] ? d_out[k] + off : nullptr;
    int32_t* ig[2] = {d_ignore[0] + off, d_ignore[1] + off};
    int32_t* ty[2] = {d_type && d_type[0] ? d_type[0] + off : nullptr, d_type && d_type[1] ? d_type[1] + off : nullptr};
    AttSpec as;
    as.on = true; as.A0 = A0; as.frequency = frequency_ghz; as.att[0] = d_att[0] + off; as.att[1] = d_att[1] + off;
    const int rc = two_rays_chunk(c, m, d_rx_depth + off, d_distance + off, d_tx_depth + off, o, ig, ty, (cudaStream_t)stream, &as);
    if (rc) return rc;
  }
  return 0;
}

int airice_inice_two_rays_att_host(airice_ctx* c, int64_t n, const double* rx_depth, const double* distance,
                                   const double* tx_depth, double A0, double frequency_ghz, double* out, double* att,
                                   int32_t* ignore) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!rx_depth || !distance || !tx_depth || !out || !att || !ignore) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int nc = AIRICE_INICE_RAYS_NCOLS;
  const int64_t chunk = n < kRaysChunk ? n : kRaysChunk;
  int rc = ensure_slots(c, (size_t)chunk * (sizeof(double) * (3 + nc + 2) + 2 * sizeof(int32_t)) + 64);
  if (rc) return rc;
  cudaStream_t s = c->streams[0];
  double* dh = (double*)c->dev[0];
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    CK(cudaMemcpyAsync(dh, rx_depth + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, distance + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + 2 * chunk, tx_depth + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    double* o[nc];
    for (int k = 0; k < nc; k++) o[k] = dh + (3 + k) * chunk;
    double* at[2] = {dh + (3 + nc) * chunk, dh + (4 + nc) * chunk};
    int32_t* ig0 = (int32_t*)(dh + (5 + nc) * chunk);
    int32_t* ig[2] = {ig0, ig0 + chunk};
    rc = airice_inice_two_rays_att_device(c, m, dh, dh + chunk, dh + 2 * chunk, A0, frequency_ghz, o, at, ig, nullptr, s);
    if (rc) return rc;
    for (int k = 0; k < nc; k++)
      CK(cudaMemcpyAsync(out + (int64_t)k * n + off, o[k], sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    for (int k = 0; k < 2; k++) {
      CK(cudaMemcpyAsync(att + (int64_t)k * n + off, at[k], sizeof(double) * m, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(ignore + (int64_t)k * n + off, ig[k], sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s));
    }
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}

int airice_inice_attenuation_device(airice_ctx* c, int64_t n, int kind, double A0, double frequency_ghz, const double* d_z0,
                                    const double* d_z1, const double* d_zmax, const double* d_L, double* d_out, void* stream) {
  if (!c) return fail(-1, "null context");
  if (kind < 0 || kind > 2) return fail(-1, "kind: 0 direct, 1 reflected, 2 refracted");
  if (n == 0) return 0;
  if (!d_z0 || !d_z1 || !d_L || !d_out || (kind == 2 && !d_zmax)) return fail(-1, "null argument");
  if (!(frequency_ghz > 0)) return fail(-3, "frequency must be positive (GHz)");
  CK(cudaSetDevice(c->device));
  InIceAttArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.kind = kind; a.z0 = d_z0; a.z1 = d_z1; a.zmax = d_zmax; a.L = d_L; a.out = d_out;
  a.A = c->medium.A_ice; a.B = c->medium.B_ice; a.C = c->medium.C_ice;
  a.A0 = A0; a.frequency = frequency_ghz; a.w0 = log(0.0001); a.w2 = log(3.16); a.w = log(frequency_ghz);
  { int rc = quad_stats(c, &a.quad_stats); if (rc) return rc; }
  cudaError_t e = launch_inice_attenuation(a, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_inice_attenuation");
  return 0;
}

int airice_inice_attenuation_host(airice_ctx* c, int64_t n, int kind, double A0, double frequency_ghz, const double* z0,
                                  const double* z1, const double* zmax, const double* L, double* out) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!z0 || !z1 || !L || !out || (kind == 2 && !zmax)) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int64_t chunk = n < (1 << 20) ? n : (1 << 20);
  int rc = ensure_slots(c, (size_t)chunk * sizeof(double) * 5 + 64);
  if (rc) return rc;
  cudaStream_t s = c->streams[0];
  double* dh = (double*)c->dev[0];
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    CK(cudaMemcpyAsync(dh, z0 + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, z1 + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    if (kind == 2) CK(cudaMemcpyAsync(dh + 2 * chunk, zmax + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + 3 * chunk, L + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    rc = airice_inice_attenuation_device(c, m, kind, A0, frequency_ghz, dh, dh + chunk, dh + 2 * chunk, dh + 3 * chunk, dh + 4 * chunk, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out + off, dh + 4 * chunk, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}

int airice_inice_ladder_stats(airice_ctx* c, int64_t out[4]) {
  if (!c || !out) return fail(-1, "null argument");
  out[0] = out[1] = out[2] = out[3] = 0;
  if (!c->inice_scratch[0]) return 0;
  CK(cudaSetDevice(c->device));
  int32_t h[8];
  CK(cudaMemcpy(h, c->inice_scratch[0], sizeof(h), cudaMemcpyDeviceToHost));    // synchronises with the kernels that wrote it
  unsigned long long w[2];
  std::memcpy(w, h + 4, sizeof(w));
  out[0] = h[0]; out[1] = h[2]; out[2] = (int64_t)w[0]; out[3] = (int64_t)w[1];
  return 0;
}

int airice_inice_quadrature_stats(airice_ctx* c, int64_t out[2]) {
  if (!c || !out) return fail(-1, "null argument");
  out[0] = out[1] = 0;
  if (!c->quad_stats) return 0;
  CK(cudaSetDevice(c->device));
  int32_t h[2];
  CK(cudaMemcpy(h, c->quad_stats, sizeof(h), cudaMemcpyDeviceToHost));   // synchronises with the kernels that wrote it
  out[0] = h[0]; out[1] = h[1];
  return 0;
}

namespace {
// scratch of one focusing chunk: [rx - 0.01][10 cols A][10 cols B][ignore A x2, ignore B x2 as int32]
int focus_scratch(airice_ctx* c, int64_t m, cudaStream_t s) {
  if (c->focus_cap >= m) return 0;
  CK(cudaStreamSynchronize(s));
  if (c->focus_scratch) cudaFree(c->focus_scratch);
  c->focus_scratch = nullptr; c->focus_cap = 0;
  CK(cudaMalloc((void**)&c->focus_scratch, sizeof(double) * 23 * (size_t)m));
  c->focus_cap = m;
  return 0;
}

// GetFocusingFactor for m DEVICE triples.  sol_a / ign_a: optional, the caller's own two-ray solution at zR (MakeTable
// has just computed it: the reference computes it twice); else solved here.
int focusing_chunk(airice_ctx* c, int64_t m, const double* d_zT, const double* d_xR, const double* d_zR, double* const* out2,
                   double* const* sol_a_in, cudaStream_t s) {
  int rc = focus_scratch(c, m, s);
  if (rc) return rc;
  const int64_t cap = c->focus_cap;
  double* base = c->focus_scratch;
  double* rxb = base;
  double* A[AIRICE_INICE_RAYS_NCOLS];
  double* B[AIRICE_INICE_RAYS_NCOLS];
  for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++) { A[k] = base + (1 + k) * cap; B[k] = base + (11 + k) * cap; }
  int32_t* ig = (int32_t*)(base + 21 * cap);
  int32_t* igA[2] = {ig, ig + cap};
  int32_t* igB[2] = {ig + 2 * cap, ig + 3 * cap};
  if (!sol_a_in) {
    rc = two_rays_chunk(c, m, d_zR, d_xR, d_zT, A, igA, nullptr, s);     // GetRayTracingSolutions(zR, xR, zT, ...)
    if (rc) return rc;
  }
  cudaError_t e = launch_inice_shift(m, d_zR, -0.01, rxb, s);
  if (e != cudaSuccess) return cuda_fail(e, "launch_inice_shift");
  rc = two_rays_chunk(c, m, rxb, d_xR, d_zT, B, igB, nullptr, s);        // GetRayTracingSolutions(zR - 0.01, xR, zT, ...)
  if (rc) return rc;
  InIceFocusArgs f;
  std::memset(&f, 0, sizeof(f));
  f.n = m; f.zT = d_zT; f.zR = d_zR;
  for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++) { f.sol_a[k] = sol_a_in ? sol_a_in[k] : A[k]; f.sol_b[k] = B[k]; }
  f.A = c->medium.A_ice; f.B = c->medium.B_ice; f.C = c->medium.C_ice;
  f.out[0] = out2[0]; f.out[1] = out2[1];
  e = launch_inice_focusing(f, s);
  if (e != cudaSuccess) return cuda_fail(e, "launch_inice_focusing");
  return 0;
}
}  // namespace

int airice_inice_focusing_device(airice_ctx* c, int64_t n, const double* d_zT, const double* d_xR, const double* d_zR,
                                 double* const* d_out, void* stream) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!d_zT || !d_xR || !d_zR || !d_out || !d_out[0] || !d_out[1]) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  for (int64_t off = 0; off < n; off += kRaysChunk) {
    const int64_t m = (n - off < kRaysChunk) ? (n - off) : kRaysChunk;
    double* o[2] = {d_out[0] + off, d_out[1] + off};
    const int rc = focusing_chunk(c, m, d_zT + off, d_xR + off, d_zR + off, o, nullptr, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int airice_inice_focusing_host(airice_ctx* c, int64_t n, const double* zT, const double* xR, const double* zR, double* out) {
  if (!c) return fail(-1, "null context");
  if (n == 0) return 0;
  if (!zT || !xR || !zR || !out) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int64_t chunk = n < kRaysChunk ? n : kRaysChunk;
  int rc = ensure_slots(c, (size_t)chunk * sizeof(double) * 5 + 64);
  if (rc) return rc;
  cudaStream_t s = c->streams[0];
  double* dh = (double*)c->dev[0];
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    CK(cudaMemcpyAsync(dh, zT + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, xR + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + 2 * chunk, zR + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    double* o[2] = {dh + 3 * chunk, dh + 4 * chunk};
    rc = focusing_chunk(c, m, dh, dh + chunk, dh + 2 * chunk, o, nullptr, s);
    if (rc) return rc;
    for (int k = 0; k < 2; k++) CK(cudaMemcpyAsync(out + (int64_t)k * n + off, o[k], sizeof(double) * m, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}

// IceRayTracing::MakeTable (IceRayTracing.cc:2614-2724)
int airice_inice_table_create(airice_ctx* c, double shower_hit_distance, double shower_depth, double zR, double step_x,
                              double step_z, double width_x, double width_z, airice_inice_table** out) {
  if (!c || !out) return fail(-1, "null argument");
  if (!(step_x > 0) || !(step_z > 0) || !(width_x > 0) || !(width_z > 0)) return fail(-3, "bad in-ice table grid");
  CK(cudaSetDevice(c->device));
  airice_inice_table* t = new airice_inice_table();
  t->device = c->device;
  t->step_x = step_x; t->step_z = step_z;
  const double nx = (width_x / step_x) + 1, nz = (width_z / step_z) + 1;      // the reference's double -> int truncation
  if (!(nx >= 2) || !(nz >= 2) || nx * nz >= 2147483647.0) { delete t; return fail(-5, "in-ice table needs 2 .. 2^31-1 nodes"); }
  t->n_x = (int)nx; t->n_z = (int)nz;
  t->points = (int64_t)t->n_x * t->n_z;
  double start_x = shower_hit_distance - (width_x / 2);
  if (shower_hit_distance <= width_x / 2) start_x = 0;
  double start_z = shower_depth - (width_z / 2);
  const double stop_z = shower_depth + (width_z / 2);
  if (fabs(shower_depth) <= 10 || stop_z >= 0) start_z = -20;
  t->pos_x.resize(t->n_x); t->pos_z.resize(t->n_z);
  for (int ix = 0; ix < t->n_x; ix++) t->pos_x[ix] = (float)(start_x + step_x * ix);    // float, IceRayTracing.hh:29-30
  for (int iz = 0; iz < t->n_z; iz++) t->pos_z[iz] = (float)(start_z + step_z * iz);
  const size_t cols_b = sizeof(double) * (size_t)t->points * AIRICE_INICE_TABLE_NCOLS;
  cudaError_t e = cudaMalloc((void**)&t->block, cols_b + sizeof(float) * ((size_t)t->n_x + t->n_z));
  if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaMalloc(in-ice table)"); }
  t->d_pos_x = (float*)((char*)t->block + cols_b);
  t->d_pos_z = t->d_pos_x + t->n_x;
  e = cudaMemcpy(t->d_pos_x, t->pos_x.data(), sizeof(float) * t->n_x, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_pos_z, t->pos_z.data(), sizeof(float) * t->n_z, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { airice_inice_table_destroy(t); return cuda_fail(e, "upload positions"); }
  // per chunk: xT, zT, rx | 10 solution columns | 2 attenuation | 2 focusing | 2 ignore (int32 pairs)
  const int64_t chunk = t->points < kRaysChunk ? t->points : kRaysChunk;
  double* scratch = nullptr;
  e = cudaMalloc((void**)&scratch, sizeof(double) * 18 * (size_t)chunk);
  if (e != cudaSuccess) { airice_inice_table_destroy(t); return cuda_fail(e, "cudaMalloc(in-ice table scratch)"); }
  auto done = [&](int code) { cudaFree(scratch); if (code) airice_inice_table_destroy(t); return code; };
  cudaStream_t s = nullptr;
  for (int64_t off = 0; off < t->points; off += chunk) {
    const int64_t m = t->points - off < chunk ? t->points - off : chunk;
    InIceTableNodeArgs g;
    std::memset(&g, 0, sizeof(g));
    g.node0 = off; g.n = m; g.n_z = t->n_z; g.start_x = start_x; g.start_z = start_z; g.step_x = step_x; g.step_z = step_z; g.zR = zR;
    g.xT = scratch; g.zT = scratch + chunk; g.rx = scratch + 2 * chunk;
    e = launch_inice_table_nodes(g, s);
    if (e != cudaSuccess) return done(cuda_fail(e, "launch_inice_table_nodes"));
    double* sol[AIRICE_INICE_RAYS_NCOLS];
    for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++) sol[k] = scratch + (3 + k) * chunk;
    double* at[2] = {scratch + 13 * chunk, scratch + 14 * chunk};
    double* fo[2] = {scratch + 15 * chunk, scratch + 16 * chunk};
    int32_t* ig0 = (int32_t*)(scratch + 17 * chunk);
    int32_t* ig[2] = {ig0, ig0 + chunk};
    // GetRayTracingSolutions(zR, xT, zT, ...) with A0 = 1, frequency = 0.1 GHz (IceRayTracing.cc:2654-2664)
    int rc = airice_inice_two_rays_att_device(c, m, g.rx, g.xT, g.zT, 1.0, 0.1, sol, at, ig, nullptr, s);
    if (rc) return done(rc);
    // GetFocusingFactor(zT, xT, zR): its first solution is the one above
    rc = focusing_chunk(c, m, g.zT, g.xT, g.rx, fo, sol, s);
    if (rc) return done(rc);
    InIceTablePackArgs pk;
    std::memset(&pk, 0, sizeof(pk));
    pk.n = m;
    for (int k = 0; k < AIRICE_INICE_RAYS_NCOLS; k++) pk.sol[k] = sol[k];
    pk.att[0] = at[0]; pk.att[1] = at[1]; pk.ignore[0] = ig[0]; pk.ignore[1] = ig[1]; pk.focusing[0] = fo[0]; pk.focusing[1] = fo[1];
    for (int k = 0; k < AIRICE_INICE_TABLE_NCOLS; k++) pk.col[k] = t->block + (size_t)k * t->points + off;
    e = launch_inice_table_pack(pk, s);
    if (e != cudaSuccess) return done(cuda_fail(e, "launch_inice_table_pack"));
  }
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return done(cuda_fail(e, "in-ice table"));
  *out = t;
  return done(0);
}

void airice_inice_table_destroy(airice_inice_table* t) {
  if (!t) return;
  if (t->block) { cudaSetDevice(t->device); cudaFree(t->block); }
  delete t;
}

int airice_inice_table_info(const airice_inice_table* t, int64_t info[3]) {
  if (!t || !info) return fail(-1, "null argument");
  info[0] = t->n_x; info[1] = t->n_z; info[2] = t->points;
  return 0;
}

int airice_inice_table_copy_column(const airice_inice_table* t, int col, double* host_out) {
  if (!t || !host_out || col < 0 || col >= AIRICE_INICE_TABLE_NCOLS) return fail(-1, "bad table/column");
  CK(cudaSetDevice(t->device));
  CK(cudaMemcpy(host_out, t->block + (size_t)col * t->points, sizeof(double) * (size_t)t->points, cudaMemcpyDeviceToHost));
  return 0;
}

int airice_inice_table_copy_positions(const airice_inice_table* t, float* host_x, float* host_z) {
  if (!t || !host_x || !host_z) return fail(-1, "null argument");
  std::memcpy(host_x, t->pos_x.data(), sizeof(float) * t->n_x);
  std::memcpy(host_z, t->pos_z.data(), sizeof(float) * t->n_z);
  return 0;
}

int airice_inice_table_interp_device(airice_ctx* c, const airice_inice_table* t, int64_t n, const double* d_x, const double* d_z,
                                     int rt_parameter, double* d_out, void* stream) {
  if (!c || !t) return fail(-1, "null argument");
  if (rt_parameter < 0 || rt_parameter >= AIRICE_INICE_TABLE_NCOLS) return fail(-1, "rtParameter out of range (0..12)");
  if (n == 0) return 0;
  if (!d_x || !d_z || !d_out) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  InIceTableInterpArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.x = d_x; a.z = d_z; a.pos_x = t->d_pos_x; a.pos_z = t->d_pos_z; a.n_x = t->n_x; a.n_z = t->n_z;
  a.step_x = t->step_x; a.step_z = t->step_z; a.col = t->block + (size_t)rt_parameter * t->points; a.out = d_out;
  cudaError_t e = launch_inice_table_interp(a, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_inice_table_interp");
  return 0;
}

int airice_inice_table_interp_host(airice_ctx* c, const airice_inice_table* t, int64_t n, const double* x, const double* z,
                                   int rt_parameter, double* out) {
  if (!c || !t) return fail(-1, "null argument");
  if (n == 0) return 0;
  if (!x || !z || !out) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int64_t chunk = n < (1 << 20) ? n : (1 << 20);
  int rc = ensure_slots(c, (size_t)chunk * sizeof(double) * 3 + 64);
  if (rc) return rc;
  int slot = 0;
  for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    double* dh = (double*)c->dev[slot];
    cudaStream_t s = c->streams[slot];
    CK(cudaMemcpyAsync(dh, x + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, z + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    rc = airice_inice_table_interp_device(c, t, m, dh, dh + chunk, rt_parameter, dh + 2 * chunk, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out + off, dh + 2 * chunk, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
  }
  for (int s = 0; s < airice_ctx::kSlots; s++)
    if (c->streams[s]) CK(cudaStreamSynchronize(c->streams[s]));
  return 0;
}

namespace {
// plans: per-ray plan scratch for this launch, or nullptr = the context's own (grow-only; callers on one stream)
int ray_path_launch(airice_ctx* c, int64_t n, const double* d_theta, const double* d_h, double depth_m, double ice_m,
                    int64_t max_points, double* d_x, double* d_z, int32_t* d_count, void* plans, cudaStream_t s) {
  const int in_ice = depth_m < 0 ? 1 : 0;
  const AirIcePlan& p = c->plan(ice_m, in_ice ? depth_m : 0.0);
  PathArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.theta = d_theta; a.h = d_h; a.max_points = max_points; a.x = d_x; a.z = d_z; a.count = d_count;
  if (!plans) {
    const size_t need = path_plan_bytes() * (size_t)n;
    if (c->path_plan_cap < need) {
      CK(cudaDeviceSynchronize());      // an earlier call on another stream may still read the old scratch
      if (c->path_plans) cudaFree(c->path_plans);
      c->path_plans = nullptr; c->path_plan_cap = 0;
      CK(cudaMalloc(&c->path_plans, need));
      c->path_plan_cap = need;
    }
    plans = c->path_plans;
  }
  a.plans = (AirIcePathPlan*)plans;
  cudaError_t e = launch_ray_path(c->medium, p, a, s);
  if (e != cudaSuccess) return cuda_fail(e, "launch_ray_path");
  return 0;
}
}  // namespace

int airice_ray_path_device(airice_ctx* c, int64_t n, const double* d_theta, const double* d_h, double depth_m, double ice_m,
                           int64_t max_points, double* d_x, double* d_z, int32_t* d_count, void* stream) {
  if (!c) return fail(-1, "null context");
  NEED_AIR(c);
  if (n == 0) return 0;
  if (!d_theta || !d_h || !d_count || (max_points > 0 && (!d_x || !d_z))) return fail(-1, "null argument");
  if (max_points < 0) return fail(-3, "max_points < 0");
  CK(cudaSetDevice(c->device));
  return ray_path_launch(c, n, d_theta, d_h, depth_m, ice_m, max_points, d_x, d_z, d_count, nullptr, (cudaStream_t)stream);
}

int airice_ray_path_host(airice_ctx* c, int64_t n, const double* theta, const double* h, double depth_m, double ice_m,
                         int64_t max_points, double* x, double* z, int32_t* count) {
  if (!c) return fail(-1, "null context");
  NEED_AIR(c);
  if (n == 0) return 0;
  if (!theta || !h || !count || (max_points > 0 && (!x || !z))) return fail(-1, "null argument");
  if (max_points < 0) return fail(-3, "max_points < 0");
  CK(cudaSetDevice(c->device));
  // rays per chunk so that one chunk's x and z rows stay within ~256 MB
  int64_t chunk = max_points > 0 ? (int64_t)(16 << 20) / max_points : n;
  if (chunk < 1) chunk = 1;
  if (chunk > n) chunk = n;
  // each staging slot carries its own plan scratch: consecutive chunks run on two streams, and the plan kernel of chunk
  // k+1 must not overwrite the plans the fill kernel of chunk k is still reading
  const size_t plan_off = ((size_t)chunk * (sizeof(double) * (2 + 2 * (size_t)max_points) + sizeof(int32_t)) + 255) / 256 * 256;
  int rc = ensure_slots(c, plan_off + path_plan_bytes() * (size_t)chunk + 64);
  if (rc) return rc;
  int slot = 0;
  for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    double* dh = (double*)c->dev[slot];
    cudaStream_t s = c->streams[slot];
    double* dx = dh + 2 * chunk;
    double* dz = dx + chunk * max_points;
    int32_t* dc = (int32_t*)(dz + chunk * max_points);
    void* plans = (char*)c->dev[slot] + plan_off;
    CK(cudaMemcpyAsync(dh, theta + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, h + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    rc = ray_path_launch(c, m, dh, dh + chunk, depth_m, ice_m, max_points, dx, dz, dc, plans, s);
    if (rc) return rc;
    if (max_points > 0) {
      CK(cudaMemcpyAsync(x + off * max_points, dx, sizeof(double) * m * max_points, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(z + off * max_points, dz, sizeof(double) * m * max_points, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(count + off, dc, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s));
  }
  for (int s = 0; s < airice_ctx::kSlots; s++)
    if (c->streams[s]) CK(cudaStreamSynchronize(c->streams[s]));
  return 0;
}

// ---- kernel 6: old solve-per-cell table + inverse-distance lookup (MakeTable / GetInterpolatedValue)
namespace {
int oldtable_shape(airice_oldtable* t, double ice_m, double start_th, double stop_th, double step_h, double step_th) {
  if (!(step_h > 0) || !(step_th > 0) || !(stop_th > start_th)) return fail(-3, "bad old-table grid");
  t->start_h = ice_m + 1;                       // MultiRayAirIceRefraction.cc:1629-1636
  t->stop_h = 100000;
  t->start_th = start_th; t->stop_th = stop_th; t->step_h = step_h; t->step_th = step_th;
  const double wh = t->stop_h - t->start_h, wt = stop_th - start_th;
  const double nh = (wh / step_h) + 1, nt = (wt / step_th) + 1;
  if (!(nh >= 2) || !(nt >= 2) || nh * nt >= 2147483647.0) return fail(-5, "old-table grid needs 2 .. 2^31-1 nodes (the reference indexes them with int)");
  t->n_h = (int)nh; t->n_th = (int)nt;          // the reference's double -> int truncation
  t->points = (int64_t)t->n_h * t->n_th;
  t->pos_h.resize(t->n_h); t->pos_th.resize(t->n_th);
  for (int ih = 0; ih < t->n_h; ih++) t->pos_h[ih] = (ih == t->n_h - 1) ? t->stop_h : t->start_h + step_h * ih;
  for (int it = 0; it < t->n_th; it++) t->pos_th[it] = (it == t->n_th - 1) ? stop_th : start_th + step_th * it;
  return 0;
}
int oldtable_alloc(airice_oldtable* t) {
  const size_t cols_b = sizeof(double) * (size_t)t->points * AIRICE_OLDTABLE_COLS;
  const size_t pos_b = sizeof(double) * ((size_t)t->n_h + (size_t)t->n_th);
  CK(cudaMalloc((void**)&t->block, cols_b + pos_b));
  t->d_pos_h = t->block + (size_t)t->points * AIRICE_OLDTABLE_COLS;
  t->d_pos_th = t->d_pos_h + t->n_h;
  CK(cudaMemcpy(t->d_pos_h, t->pos_h.data(), sizeof(double) * t->n_h, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(t->d_pos_th, t->pos_th.data(), sizeof(double) * t->n_th, cudaMemcpyHostToDevice));
  return 0;
}
}  // namespace

int airice_oldtable_create(airice_ctx* c, double ice_m, double depth_m, double start_th, double stop_th, double step_h,
                           double step_th, airice_oldtable** out) {
  if (!c || !out) return fail(-1, "null argument");
  NEED_AIR(c);
  CK(cudaSetDevice(c->device));
  airice_oldtable* t = new airice_oldtable();
  t->ctx = c; t->device = c->device;
  int rc = oldtable_shape(t, ice_m, start_th, stop_th, step_h, step_th);
  if (rc == 0) rc = oldtable_alloc(t);
  if (rc) { airice_oldtable_destroy(t); return rc; }
  // tan() of the node angles on the host: the distance each node asks for is formed from glibc's tan like the reference's
  std::vector<double> tn(t->n_th);
  for (int it = 0; it < t->n_th; it++) tn[it] = tan((180 - t->pos_th[it]) * (c->medium.pi / 180.0));
  const int64_t chunk = t->points < (4 << 20) ? t->points : (4 << 20);
  double* scratch = nullptr;     // [tan n_th] then 11 columns of one chunk: h, d, th | X, X_air, t_air, t_ice, launch, T_S, T_P, incident
  cudaError_t e = cudaMalloc((void**)&scratch, sizeof(double) * ((size_t)t->n_th + 11 * (size_t)chunk));
  if (e != cudaSuccess) { airice_oldtable_destroy(t); return cuda_fail(e, "cudaMalloc(old table scratch)"); }
  auto done = [&](int code) { cudaFree(scratch); if (code) airice_oldtable_destroy(t); return code; };
  e = cudaMemcpy(scratch, tn.data(), sizeof(double) * t->n_th, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return done(cuda_fail(e, "upload tangents"));
  double* col = scratch + t->n_th;
  for (int64_t off = 0; off < t->points && rc == 0; off += chunk) {
    const int64_t m = t->points - off < chunk ? t->points - off : chunk;
    OldGridArgs g;
    std::memset(&g, 0, sizeof(g));
    g.cell0 = off; g.n = m; g.n_th = t->n_th; g.pos_h = t->d_pos_h; g.pos_th = t->d_pos_th; g.col_tan = scratch;
    g.ice = ice_m; g.depth = depth_m; g.h = col; g.d = col + chunk; g.th = col + 2 * chunk;
    e = launch_oldgrid_cells(g, nullptr);
    if (e != cudaSuccess) return done(cuda_fail(e, "launch_oldgrid_cells"));
    double* cols[AIRICE_SOLVE_NCOLS] = {nullptr};
    const int want[8] = {0, 1, 3, 4, 5, 7, 8, 11};
    for (int k = 0; k < 8; k++) cols[want[k]] = col + (3 + k) * chunk;
    // ok flag not needed (MakeTable applies its own acceptance test); the solver wants a flag buffer: reuse the h column's tail? no -- own bytes
    rc = airice_solve_device(c, m, g.h, g.d, g.th, depth_m, ice_m, AIRICE_UNITS_M_DEG, cols, nullptr, nullptr, nullptr);
    if (rc) return done(rc);
    OldPackArgs pk;
    std::memset(&pk, 0, sizeof(pk));
    pk.n = m; pk.h = g.h; pk.d = g.d; pk.x = cols[0]; pk.x_air = cols[1]; pk.t_air = cols[3]; pk.t_ice = cols[4];
    pk.launch = cols[5]; pk.ts = cols[7]; pk.tp = cols[8]; pk.inc = cols[11]; pk.c = c->medium.c;
    for (int k = 0; k < AIRICE_OLDTABLE_COLS; k++) pk.col[k] = t->block + (size_t)k * t->points + off;
    e = launch_oldgrid_pack(pk, nullptr);
    if (e != cudaSuccess) return done(cuda_fail(e, "launch_oldgrid_pack"));
  }
  e = cudaStreamSynchronize(nullptr);
  if (e != cudaSuccess) return done(cuda_fail(e, "old table"));
  *out = t;
  return done(0);
}

int airice_oldtable_wrap_host(airice_ctx* c, double ice_m, double start_th, double stop_th, double step_h, double step_th,
                              const double* cols9, airice_oldtable** out) {
  if (!c || !out || !cols9) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  airice_oldtable* t = new airice_oldtable();
  t->ctx = c; t->device = c->device;
  int rc = oldtable_shape(t, ice_m, start_th, stop_th, step_h, step_th);
  if (rc == 0) rc = oldtable_alloc(t);
  if (rc == 0) {
    cudaError_t e = cudaMemcpy(t->block, cols9, sizeof(double) * (size_t)t->points * AIRICE_OLDTABLE_COLS, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "upload old table");
  }
  if (rc) { airice_oldtable_destroy(t); return rc; }
  *out = t;
  return 0;
}

void airice_oldtable_destroy(airice_oldtable* t) {
  if (!t) return;
  if (t->block) { cudaSetDevice(t->device); cudaFree(t->block); }
  delete t;
}

int airice_oldtable_info(const airice_oldtable* t, int64_t info[3]) {
  if (!t || !info) return fail(-1, "null argument");
  info[0] = t->n_h; info[1] = t->n_th; info[2] = t->points;
  return 0;
}

int airice_oldtable_copy_column(const airice_oldtable* t, int col, double* host_out) {
  if (!t || !host_out || col < 0 || col >= AIRICE_OLDTABLE_COLS) return fail(-1, "bad table/column");
  CK(cudaSetDevice(t->device));
  CK(cudaMemcpy(host_out, t->block + (size_t)col * t->points, sizeof(double) * (size_t)t->points, cudaMemcpyDeviceToHost));
  return 0;
}

int airice_oldtable_copy_positions(const airice_oldtable* t, double* host_h, double* host_th) {
  if (!t || !host_h || !host_th) return fail(-1, "null argument");
  std::memcpy(host_h, t->pos_h.data(), sizeof(double) * t->n_h);
  std::memcpy(host_th, t->pos_th.data(), sizeof(double) * t->n_th);
  return 0;
}

int airice_oldtable_interp_device(airice_ctx* c, const airice_oldtable* t, int64_t n, const double* d_h, const double* d_th,
                                  int rt_parameter, double* d_out, void* stream) {
  if (!c || !t) return fail(-1, "null argument");
  if (rt_parameter < 0 || rt_parameter >= AIRICE_OLDTABLE_COLS) return fail(-1, "rtParameter out of range (0..8)");
  if (n == 0) return 0;
  if (!d_h || !d_th || !d_out) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  OldInterpArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = n; a.h = d_h; a.th = d_th; a.z = t->block + (size_t)rt_parameter * t->points; a.pos_h = t->d_pos_h; a.pos_th = t->d_pos_th;
  a.n_h = t->n_h; a.n_th = t->n_th; a.start_h = t->start_h; a.start_th = t->start_th; a.step_h = t->step_h; a.step_th = t->step_th;
  a.out = d_out;
  cudaError_t e = launch_old_interp(a, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_old_interp");
  return 0;
}

int airice_oldtable_interp_host(airice_ctx* c, const airice_oldtable* t, int64_t n, const double* h, const double* th,
                                int rt_parameter, double* out) {
  if (!c || !t) return fail(-1, "null argument");
  if (n == 0) return 0;
  if (!h || !th || !out) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  const int64_t chunk = n < (1 << 20) ? n : (1 << 20);
  int rc = ensure_slots(c, (size_t)chunk * sizeof(double) * 3 + 64);
  if (rc) return rc;
  int slot = 0;
  for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    double* dh = (double*)c->dev[slot];
    cudaStream_t s = c->streams[slot];
    CK(cudaMemcpyAsync(dh, h + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dh + chunk, th + off, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    rc = airice_oldtable_interp_device(c, t, m, dh, dh + chunk, rt_parameter, dh + 2 * chunk, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out + off, dh + 2 * chunk, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
  }
  for (int s = 0; s < airice_ctx::kSlots; s++)
    if (c->streams[s]) CK(cudaStreamSynchronize(c->streams[s]));
  return 0;
}

// ---- peer memory (CUDA IPC): see include/airice_b200.h
int airice_peer_alloc(airice_ctx* c, size_t bytes, void** d_ptr, unsigned char handle[AIRICE_PEER_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == AIRICE_PEER_HANDLE_BYTES, "handle size");
  if (!c || !d_ptr || !handle) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  void* p = nullptr;
  CK(cudaMalloc(&p, bytes ? bytes : 1));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  std::memcpy(handle, &h, sizeof(h));
  *d_ptr = p;
  return 0;
}
int airice_peer_free(airice_ctx* c, void* d_ptr) {
  if (!c) return fail(-1, "null context");
  CK(cudaSetDevice(c->device));
  CK(cudaFree(d_ptr));
  return 0;
}
int airice_peer_open(airice_ctx* c, const unsigned char handle[AIRICE_PEER_HANDLE_BYTES], void** d_ptr) {
  if (!c || !d_ptr || !handle) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int airice_peer_close(airice_ctx* c, void* d_ptr) {
  if (!c) return fail(-1, "null context");
  CK(cudaSetDevice(c->device));
  CK(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}
int airice_peer_copy(airice_ctx* c, void* d_dst, const void* d_src, size_t bytes, void* stream) {
  if (!c) return fail(-1, "null context");
  if (!bytes) return 0;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int airice_fp64_peak_tflops(airice_ctx* c, double* tflops) {
  if (!c || !tflops) return fail(-1, "null argument");
  CK(cudaSetDevice(c->device));
  cudaError_t e = fp64_peak_probe(tflops, 4096, nullptr);
  if (e != cudaSuccess) return cuda_fail(e, "fp64_peak_probe");
  return 0;
}

// ---- page-locked host buffers for the host entry points
int airice_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return fail(-1, "null argument");
  CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}
int airice_host_unregister(void* p) {
  if (!p) return fail(-1, "null argument");
  CK(cudaHostUnregister(p));
  return 0;
}
int airice_host_alloc(size_t bytes, void** p) {
  if (!p || bytes == 0) return fail(-1, "null argument");
  *p = nullptr;
  CK(cudaHostAlloc(p, bytes, cudaHostAllocPortable));
  return 0;
}
int airice_host_free(void* p) {
  if (!p) return 0;
  CK(cudaFreeHost(p));
  return 0;
}

int airice_sync(airice_ctx* c) {
  if (!c) return fail(-1, "null context");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  return 0;
}

int airice_trim(airice_ctx* c) {
  if (!c) return fail(-1, "null context");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  drop_spares(c);
  for (auto& d : c->defer) if (d.second.buf) cudaFree(d.second.buf);
  c->defer.clear();
  return 0;
}

}  // extern "C"
