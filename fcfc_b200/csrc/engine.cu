// fcfc_b200/csrc/engine.cu -- host side of libfcfc_b200.so: the C ABI of include/fcfc_gpu.h.
//
// Replaces tree_create / tree_destroy / count_pairs of the reference
// (src/fcfc/2pt_box/build_tree.c:36-219, count_func.c:4847-7724 and the survey twins) with
//   catalogue upload -> cell list (cell index, radix sort, gather)   [replaces src/tree/*.c]
//   neighbour stencil + work items -> persistent counting kernel      [replaces dual_tree.c,
//                                                                       metric_*.c, OpenMP/MPI]
// There is no CPU fallback anywhere in this file: without a usable device every call fails.
#include "../../include/fcfc_gpu.h"
#include "count_kernel.cuh"
#include "dispatch.h"

#include <cub/cub.cuh>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// Fixed-point scales 2^ks, 2^km of the fast bins (count_kernel.cuh, fast_bins): the flag band 2^-k must cover the error
// budget with a factor >= 2 to spare, and bin * 2^k must stay below 2^23.  Pure host arithmetic, exported so that the
// CPU tests can check the error budget against an emulation of the device arithmetic (tests/test_fastbin_budget.py).
extern "C" void fcfc_gpu_fastbin_scales(int ns, int nmu, int periodic, int *ks, int *km) {
  auto pick = [](double need, int nbin) { int k = 20; while (k > 4 && (std::ldexp(1.0, -k) < need || (double) (nbin + 2) * std::ldexp(1.0, k) >= 8388608.0)) k--; return k; };
  // (survey (s,mu): mu comes from two approximate reciprocal square roots and three converted terms, 9.5e-7 relative)
  if (ks) *ks = pick(2.5 * 3e-7 * (ns + 1), ns);
  if (km) *km = pick(2.2 * (periodic ? 4.5e-7 : 9.5e-7) * (nmu + 1), nmu);
}

// Limits of the division-free pre-tests of the survey (s_perp, pi) metric (count_kernel.cuh, eval_pair), each rounded
// up to the build's `real`: out[0] the sphere s^2 < (s2max + p2max)(1 + 8 eps) searched before anything else;
// out[1] p2max (1 + 8 eps) for d*d < (s + t) p2max'; out[2] c = s2max (1 + 32 eps) + 32 eps out[0] for
// (s^2 - c)(s + t) < d*d: c >= s2max (1 + 2 eps) + 3.2 eps max(s^2) keeps it a necessary condition of the exact test
// through every rounding (derivation in DESIGN.md), padded tenfold.  Exported for tests/test_fastbin_budget.py.
extern "C" void fcfc_gpu_survey_pretest_limits(double s2max, double p2max, int is_float, double out[3]) {
  const double eps = is_float ? (double) FLT_EPSILON : DBL_EPSILON;
  auto up = [&](double v) {
    if (!is_float) return v;
    float f = (float) v;
    if ((double) f < v) f = std::nextafter(f, INFINITY);
    return (double) f;
  };
  const double pm = (s2max + p2max) * (1 + 8 * eps);
  out[0] = up(pm);
  out[1] = up(p2max * (1 + 8 * eps));
  out[2] = up(s2max * (1 + 32 * eps) + 32 * eps * pm);
}

// Limits of the single-precision pre-filter of the double-precision kernels (count_kernel_pf.cuh), padded so that the
// filter never drops a pair the exact double-precision tests accept.  Inputs: the limits of the exact tests (s2max:
// squared separation; box (s_perp,pi): s_perp^2 and pmax; survey (s_perp,pi): s_perp^2 and p2max = pi^2), M = the largest
// |coordinate| that reaches the filter (periodic shifts included), smax_sq / smin_sq = largest / smallest |x|^2 of the two
// catalogues (survey).  With u = 2^-24:
//   a coordinate difference of an accepted pair, formed from float copies:  |dx_f - dx| <= e = 2 u M + 1.01 u r
//   its squared separation (3 squares, 2 sums, fused or not):               |d2_f - d2| <= E_d = 2 sqrt(k) r e + k e^2 + 6 u r^2
//   (k = 3, or 2 for s_perp^2 of a box; r = the largest accepted separation), the exact d2 itself carries `slack`;
//   survey (s_perp,pi), with sd = s1 - s2, st = |x1 + x2|^2 >= st_min = (2 sqrt(smin_sq) - r)^2:
//     |sd_f - sd| <= E_s = 2.02 u smax_sq,   |st_f - st| <= E_st = 12.1 u smax_sq + E_d
//     test 1  dd_f < st_f * plim:            plim  = (p2max (1 + 1e-9) + 2 sqrt(p2max / st_min) E_s + E_s^2 / st_min) (1 + 8 u) / (1 - E_st / st_min)
//     test 2  (d2_f - s2lim) st_f <= dd_f:   s2lim = s2max (1 + 1e-9) + slack + E_d + 4 u p2max + 1.01 p2max E_st / st_min + 2 sqrt(p2max / st_min) E_s
//   (derivation in DESIGN.md); every e / E is doubled for safety, results are rounded up to float.
// out[0] = d2 limit (survey (s_perp,pi): the sphere s2max + p2max), out[1] = plim, out[2] = s2lim, out[3] = the relative
// padding of the tightest limit; returns 1 when the filter is usable (padding below 5 %), 2 when in addition the
// cylinder tests of the survey (s_perp,pi) filter are, 0 otherwise.
extern "C" int fcfc_gpu_prefilter_limits(int periodic, int bintype, double s2max, double pmax, double M, double smax_sq,
                                         double smin_sq, double out[4]) {
  const double u = std::ldexp(1.0, -24);
  auto up = [](double v) { float f = (float) v; if ((double) f < v) f = std::nextafter(f, INFINITY); return (double) f; };
  const bool cyl_box = periodic && bintype == FCFC_GPU_BIN_SPI, cyl_svy = !periodic && bintype == FCFC_GPU_BIN_SPI;
  const double r2 = cyl_box ? s2max + pmax * pmax : (cyl_svy ? s2max + pmax : s2max);
  const double r = std::sqrt(r2);
  const double slack = (periodic || bintype == FCFC_GPU_BIN_ISO) ? 64 * DBL_EPSILON * M * M : 64 * DBL_EPSILON * (smax_sq + 1);
  const double e = 2.0 * (2 * u * M + 1.01 * u * r);                    // doubled
  const int k = cyl_box ? 2 : 3;
  const double rr = cyl_box ? std::sqrt(s2max) : r;
  const double Ed = 2 * std::sqrt((double) k) * rr * e + k * e * e + 12 * u * rr * rr;
  const double lim = (cyl_box ? s2max : r2) * (1 + 1e-9) + slack + Ed;
  out[0] = up(lim * (1 + 8 * u));
  out[1] = out[2] = 0;
  double pad = (out[0] / (cyl_box ? s2max : r2)) - 1;
  int mode = 1;
  if (cyl_box) { out[1] = up((pmax + e) * (1 + 8 * u)); pad = std::max(pad, out[1] / pmax - 1); }
  if (cyl_svy) {
    const double rmin = std::sqrt(std::max(smin_sq, 0.0));
    const double stm = (2 * rmin > 1.05 * r) ? (2 * rmin - r) * (2 * rmin - r) : 0;
    if (stm > 0) {
      const double Es = 2.0 * 2.02 * u * smax_sq, Est = 2.0 * 12.1 * u * smax_sq + Ed;
      if (Est < 0.01 * stm) {
        const double p2 = pmax;   // pi^2
        const double plim = (p2 * (1 + 1e-9) + 2 * std::sqrt(p2 / stm) * Es + Es * Es / stm) * (1 + 8 * u) / (1 - Est / stm);
        const double s2lim = s2max * (1 + 1e-9) + slack + Ed + 4 * u * p2 + 1.01 * p2 * Est / stm + 2 * std::sqrt(p2 / stm) * Es;
        if (plim < 1.05 * p2 && s2lim < 1.05 * s2max + 1e-300) { out[1] = up(plim); out[2] = up(s2lim * (1 + 8 * u)); mode = 2; }
      }
    }
    if (mode != 2) { out[1] = 0; out[2] = 0; }          // cylinder tests off: sphere only
  }
  out[3] = pad;
  return pad < 0.05 ? mode : 0;
}

// Limits of the classification of staged points against the tile's bounding box in the single-precision kernels
// (count_kernel_cl.cuh): a point is dropped when the squared distance to the nearest point of the box, computed in
// float, exceeds out[0], and binned in place ("dense") when the squared distance to the farthest corner is below
// out[1].  Only the first must be safe (the dense loop keeps its range test): the box centre and the image-shifted
// coordinates carry up to 3 ulp(M) per axis (M = largest coordinate magnitude, shifts included), i.e. 12 M 2^-23 / r
// relative in d^2 at the range limit r, and every evaluation a few ulps of d^2 itself; the margin is three times that.
extern "C" void fcfc_gpu_classify_limits(double s2max, double maxabs, float out[2]) {
  const double margin = 4e-6 * (maxabs / std::sqrt(std::max(s2max, 1e-300))) + 4e-6;
  out[0] = std::nextafter((float) (s2max * (1.0 + margin)), INFINITY);
  out[1] = (float) (s2max * (1.0 - margin));
}

// Error budget of the double-precision-at-float-speed kernel (count_kernel_df.cuh).  Coordinates reach the float
// arithmetic relative to the centre of the tile's cell: a primary is within cs/2 of it per axis, the secondary of a pair of
// interest within cs/2 + r (r = the largest accepted separation, padded), so with u = 2^-24 a coordinate difference
// formed in float differs from the exact one by at most e = u (cs/2) + u (cs/2 + r) + u r (two roundings to float and
// the subtraction), doubled for safety.  From it:
//   |d_f - d| <= sqrt(3) e            error of the separation vector, hence of s, on top of the arithmetic error of the
//                                     computed bins (3e-7 relative per s bin, 4.5e-7 per mu bin: fcfc_gpu_fastbin_scales)
//   |d2_f - d2| <= 2 sqrt(3) r e + 3 e^2 + 8 u r^2        -> padded range limit d2lim; everything the padding admits lies
//                                     within the band of the last s edge and is flagged
//   |nmu mu_f - nmu mu| <= nmu (1 + sqrt(3)) e / s        -> grows at small separations: the band 2^-km must cover it for
//                                     s >= s1, and pairs with s < s1 are flagged wholesale; km is the value that minimises
//                                     the flagged fraction 2 * 2^-km + (s1 / r)^3 of uniformly distributed pairs.
// Returns 1 when the kernel is usable (enough fixed-point bits, flagged fraction below 3 %), else 0.
extern "C" int fcfc_gpu_df_budget(int ns, int nmu, double s2max, double cs_max, int *ks_out, int *km_out, double *d2lim_out,
                                  double *s1sq_out, double *flagged_out) {
  const double u = std::ldexp(1.0, -24);
  const double r = std::sqrt(s2max) * (1 + 1e-3);
  const double e = 2.0 * (u * 0.5 * cs_max + u * (0.5 * cs_max + r) + u * r);
  auto fits = [](int k, int nbin) { return (double) (nbin + 2) * std::ldexp(1.0, k) < 8388608.0; };
  // s: band unit 2^-ks >= 2.5 x (arithmetic + coordinate error)
  const double err_s = 3e-7 * (ns + 1) + 1.7321 * e + 4 * u * r;
  int ks = 20;
  while (ks > 4 && (std::ldexp(1.0, -ks) < 2.5 * err_s || !fits(ks, ns))) ks--;
  int km = 0;
  double s1 = 0, best = 1e300;
  if (nmu > 1) {
    for (int k = 6; k <= 20; k++) {
      if (!fits(k, nmu)) break;
      const double band = std::ldexp(1.0, -k) - 2.2 * 4.5e-7 * (nmu + 1);
      if (band <= 0) break;
      const double s1k = 2.2 * nmu * 2.7321 * e / band;
      const double f = 2 * std::ldexp(1.0, -k) + std::pow(std::min(1.0, s1k / r), 3);
      if (f < best) { best = f; km = k; s1 = s1k; }
    }
  } else { km = 20; best = 0; }
  const double Ed = 2 * 1.7321 * r * e + 3 * e * e + 8 * u * r * r;
  float lim = (float) ((s2max + Ed) * (1 + 8 * u));
  if ((double) lim < (s2max + Ed) * (1 + 8 * u)) lim = std::nextafter(lim, INFINITY);
  float s1sq = (float) (s1 * s1);
  if ((double) s1sq < s1 * s1) s1sq = std::nextafter(s1sq, INFINITY);
  const double flagged = best + 4 * std::ldexp(1.0, -ks);
  if (ks_out) *ks_out = ks;
  if (km_out) *km_out = km;
  if (d2lim_out) *d2lim_out = lim;
  if (s1sq_out) *s1sq_out = (nmu > 1) ? s1sq : 0.0;
  if (flagged_out) *flagged_out = flagged;
  return (ks >= 6 && km >= 6 && flagged < 0.03) ? 1 : 0;
}

namespace fcfc {

// ------------------------------------------------------------------------------------------
// error handling
static thread_local std::string g_err;
static thread_local fcfc_gpu_stats g_stats;
static int g_verbose = 0;

static void set_err(const char *fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  if (g_verbose) fprintf(stderr, "[fcfc_gpu] error: %s\n", buf);
}
#define CUDA_TRY(x, code)                                                                \
  do { cudaError_t e_ = (x); if (e_ != cudaSuccess) {                                    \
    set_err("%s failed at %s:%d: %s", #x, __FILE__, __LINE__, cudaGetErrorString(e_));   \
    cudaGetLastError(); return code; } } while (0)

// ------------------------------------------------------------------------------------------
// Tuning / diagnostic options.  They are NOT read from the environment on the counting path: the process-wide
// defaults are parsed once by fcfc_gpu_init from FCFC_GPU_TUNE="name=value,..." and can be changed explicitly
// with fcfc_gpu_set_option (tests, A/B scripts).  Every count call works on a snapshot taken at its entry.
struct Options {
  int k = 0;                // force cells of reach / k (0: cost model of choose_grid)
  int nsplit = 0;           // force the number of pieces every tile's sweep list is cut into (0: automatic)
  int items_per_warp = 32;  // automatic nsplit: work items per resident warp and shard (measured: 8 leaves a 5 % tail at 8 shards, 32 leaves 0.6 %)
  int cost_bits = 1;        // mantissa bits of the cost classes of the work-item order (23: exact cost order)
  int no_subsort = 0;       // 1: no Morton order inside the cells
  int no_table_math = 0;    // 1: always look the bins up, never compute them
  int no_hist_copies = 0;   // 1: one weighted shared-memory histogram instead of 32 lane-private copies
  int qdepth = 0;           // cap of the per-lane stack depth (0: as deep as shared memory allows, <= 64)
  int qkeep = -1;           // entries a drain leaves on the fullest stack (-1: automatic)
  int force_generic = 0;    // 1: run the generic variant (cross-check of the fast paths)
  int global_hist = 0;      // 1: histogram in global memory
  int no_dense = 0;         // 1: dense-cell path off
  int no_prefilter = 0;     // 1: double-precision kernels without the float pre-filter
  int force_prefilter = 0;  // 1: take the pre-filter kernel whenever it is usable (A/B runs), not only where it was measured faster
  int no_df = 0;            // 1: double-precision box / isotropic counts without the float-speed kernel (count_kernel_df.cuh)
  int no_classify = 0;      // 1: single-precision box / isotropic counts without the classified staging (count_kernel_cl.cuh)
  int sorted_copies = 3;    // cell-sorted copies kept per catalogue and precision (one per grid in use)
  int nccl = 0;             // 1: one process, several devices: combine the histograms with an NCCL all-reduce instead of on the host
};
static Options g_opt, g_opt_base;       // current values; the process defaults (built-in, then FCFC_GPU_TUNE)
static std::mutex g_opt_mutex;

static int set_option(const char *name, long value) {
  if (!name) return FCFC_GPU_ERR_ARG;
  std::lock_guard<std::mutex> lock(g_opt_mutex);
  const struct { const char *n; int *p; } tab[] = {
      {"k", &g_opt.k}, {"nsplit", &g_opt.nsplit}, {"items_per_warp", &g_opt.items_per_warp}, {"cost_bits", &g_opt.cost_bits},
      {"no_subsort", &g_opt.no_subsort}, {"no_table_math", &g_opt.no_table_math}, {"no_hist_copies", &g_opt.no_hist_copies},
      {"qdepth", &g_opt.qdepth}, {"qkeep", &g_opt.qkeep}, {"force_generic", &g_opt.force_generic},
      {"global_hist", &g_opt.global_hist}, {"no_dense", &g_opt.no_dense}, {"no_prefilter", &g_opt.no_prefilter}, {"force_prefilter", &g_opt.force_prefilter}, {"no_df", &g_opt.no_df}, {"no_classify", &g_opt.no_classify},
      {"sorted_copies", &g_opt.sorted_copies}, {"nccl", &g_opt.nccl}};
  if (!strcmp(name, "defaults")) { g_opt = g_opt_base; return 0; }
  for (auto &t : tab) if (!strcmp(name, t.n)) { *t.p = (int) value; return 0; }
  return FCFC_GPU_ERR_ARG;
}
static Options options_snapshot() { std::lock_guard<std::mutex> lock(g_opt_mutex); return g_opt; }
// FCFC_GPU_TUNE="k=3,no_dense=1": parsed once, by fcfc_gpu_init
static void options_from_env() {
  const char *e = getenv("FCFC_GPU_TUNE");
  if (!e) return;
  std::string s(e);
  size_t pos = 0;
  while (pos < s.size()) {
    size_t end = s.find(',', pos); if (end == std::string::npos) end = s.size();
    const std::string item = s.substr(pos, end - pos);
    const size_t eq = item.find('=');
    const std::string key = item.substr(0, eq);
    const long val = (eq == std::string::npos) ? 1 : atol(item.c_str() + eq + 1);
    if (!key.empty() && set_option(key.c_str(), val) != 0) fprintf(stderr, "[fcfc_gpu] FCFC_GPU_TUNE: unknown option '%s' ignored\n", key.c_str());
    pos = end + 1;
  }
  std::lock_guard<std::mutex> lock(g_opt_mutex);
  g_opt_base = g_opt;
}

// ------------------------------------------------------------------------------------------
// device context (one per process; multi-GPU jobs run one process per GPU, or list several
// devices here and let fcfc_gpu_count loop over them)
struct Context {
  std::vector<int> devices;
  int sm_count = 0;
  bool ready = false;
};
static Context g_ctx;

static int ensure_init() {
  if (g_ctx.ready) return 0;
  return fcfc_gpu_init(0, nullptr, 0) > 0 ? 0 : FCFC_GPU_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------
// Device memory pool.  cudaMalloc / cudaFree cost milliseconds and synchronise the device; a count step
// allocates a dozen buffers of recurring sizes (catalogue columns, sort scratch, cell tables), so freed
// blocks are kept and handed out again (first fit within 2x).  Everything is returned by fcfc_gpu_finalize.
struct PoolBlock { void *p; size_t bytes; int dev; };
static std::vector<PoolBlock> g_pool_free, g_pool_used;
static std::mutex g_pool_mutex;    // fcfc_gpu_count drives several devices from worker threads
static size_t g_pool_cached = 0;
static const size_t kPoolMaxCached = (size_t) 48 << 30;

static cudaError_t pool_alloc(void **out, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  int dev = 0; cudaGetDevice(&dev);
  bytes = (bytes + 511) & ~(size_t) 511;
  if (bytes == 0) bytes = 512;
  size_t best = (size_t) -1;
  for (size_t i = 0; i < g_pool_free.size(); i++)
    if (g_pool_free[i].dev == dev && g_pool_free[i].bytes >= bytes && g_pool_free[i].bytes <= 2 * bytes &&
        (best == (size_t) -1 || g_pool_free[i].bytes < g_pool_free[best].bytes)) best = i;
  if (best != (size_t) -1) {
    PoolBlock b = g_pool_free[best];
    g_pool_free.erase(g_pool_free.begin() + best);
    g_pool_cached -= b.bytes;
    g_pool_used.push_back(b);
    *out = b.p;
    return cudaSuccess;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {               // out of memory: give the cache back and retry once
    cudaGetLastError();
    for (auto &b : g_pool_free) cudaFree(b.p);
    g_pool_free.clear(); g_pool_cached = 0;
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
  }
  g_pool_used.push_back({p, bytes, dev});
  *out = p;
  return cudaSuccess;
}
static void pool_free(void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (size_t i = 0; i < g_pool_used.size(); i++)
    if (g_pool_used[i].p == p) {
      PoolBlock b = g_pool_used[i];
      g_pool_used.erase(g_pool_used.begin() + i);
      if (g_pool_cached + b.bytes > kPoolMaxCached) { cudaFree(b.p); return; }
      g_pool_free.push_back(b); g_pool_cached += b.bytes;
      return;
    }
  cudaFree(p);
}
static void pool_release_all() {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (auto &b : g_pool_free) cudaFree(b.p);
  g_pool_free.clear(); g_pool_cached = 0;
}
template <class P> static cudaError_t pool_alloc(P **out, size_t bytes) { return pool_alloc(reinterpret_cast<void **>(out), bytes); }

// Scope guards: pool blocks and events are released on every exit path of a call.
struct PoolScope {
  std::vector<void *> held;
  template <class P> cudaError_t alloc(P **out, size_t bytes) {
    cudaError_t e = pool_alloc(out, bytes);
    if (e == cudaSuccess) held.push_back(*out);
    return e;
  }
  void free_now(void *p) {            // early release of one block
    for (size_t i = 0; i < held.size(); i++) if (held[i] == p) { held.erase(held.begin() + i); pool_free(p); return; }
  }
  template <class P> P *keep(P *p) {  // ownership passes to the caller
    for (size_t i = 0; i < held.size(); i++) if (held[i] == p) { held.erase(held.begin() + i); break; }
    return p;
  }
  ~PoolScope() { for (void *p : held) pool_free(p); }
};
struct EventScope {
  std::vector<cudaEvent_t> ev;
  explicit EventScope(int n) : ev(n) { for (auto &e : ev) cudaEventCreate(&e); }
  cudaEvent_t operator[](int i) const { return ev[i]; }
  ~EventScope() { for (auto &e : ev) cudaEventDestroy(e); }
};

// ------------------------------------------------------------------------------------------
// cell grid
struct Grid {
  int nc[3] = {1, 1, 1};
  double origin[3] = {0, 0, 0};
  double cs[3] = {0, 0, 0};     // cell size
  double box[3] = {0, 0, 0};    // periodic box (0 if not periodic)
  int periodic = 0;
  bool operator==(const Grid &o) const {
    return !memcmp(nc, o.nc, sizeof nc) && !memcmp(origin, o.origin, sizeof origin) &&
           !memcmp(cs, o.cs, sizeof cs) && !memcmp(box, o.box, sizeof box) && periodic == o.periodic;
  }
  long long ncell() const { return (long long) nc[0] * nc[1] * nc[2]; }
};

template <class T> struct Sorted {
  bool valid = false;
  unsigned long long stamp = 0; // last use (the least recently used copy is evicted)
  Grid grid;
  int tile = 0;                 // points per work item
  Vec4<T> *pos = nullptr;
  T *w = nullptr;
  int *cell_start = nullptr;    // ncell + 1
  int *item_cell = nullptr, *item_off = nullptr, *item_cnt = nullptr;
  int nitem = 0;
  void release() {
    pool_free(pos); pool_free(w); pool_free(cell_start); pool_free(item_cell); pool_free(item_off); pool_free(item_cnt);
    pos = nullptr; w = nullptr; cell_start = nullptr; item_cell = item_off = item_cnt = nullptr;
    valid = false; nitem = 0;
  }
};

}  // namespace fcfc

// One device's copy of a catalogue.
struct DevCat {
  int is_float = 0;
  size_t n = 0;
  int device = 0;
  void *x = nullptr, *y = nullptr, *z = nullptr, *s = nullptr, *w = nullptr;   // device SoA of `real`
  bool has_s = false, has_w = false;
  double bmin[3], bmax[3];      // bounding box of the (rescaled) coordinates
  double smax = 0, smin = 0;    // max / min of x^2+y^2+z^2
  double wsum = 0;
  // cell-sorted copies, one per (grid, tile) in use: DD, DR and RR of a survey run on different grids (the grid
  // follows the extents and sizes of both catalogues), and each keeps its copy instead of re-sorting on every call
  std::vector<fcfc::Sorted<float>> sf;
  std::vector<fcfc::Sorted<double>> sd;
  unsigned long long clock = 0;
};

// The public handle: one replica per device in use (the secondary catalogue must be resident everywhere).
struct fcfc_gpu_catalog {
  int is_float = 0;
  size_t n = 0;
  double wsum = 0;
  bool has_w = false;
  std::vector<DevCat *> dev;
};

namespace fcfc {

template <class T> static std::vector<Sorted<T>> &sorted_of(DevCat *c);
template <> std::vector<Sorted<float>> &sorted_of<float>(DevCat *c) { return c->sf; }
template <> std::vector<Sorted<double>> &sorted_of<double>(DevCat *c) { return c->sd; }

// ------------------------------------------------------------------------------------------
// catalogue kernels
__device__ __forceinline__ unsigned long long enc_f64(double v) {       // order-preserving encoding
  unsigned long long u = (unsigned long long) __double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
static inline double dec_f64(unsigned long long u) {
  u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  double d; memcpy(&d, &u, 8); return d;
}

// stats[0..2] = min xyz, [3..5] = max xyz, [6] = max s, [9] = min s (all order-encoded), [7] = #non-finite, [8] = sum of weights
// Rescales in `real` precision (build_tree.c:121-131) and, when requested, fills the survey's 4th
// coordinate (2pt/build_tree.c:59 scalar order / :75-82 FMA order).
template <class T>
__global__ void prep_kernel(T *x, T *y, T *z, T *s, size_t n, T rescale, int do_rescale, int sumsq,
                            unsigned long long *stats, double *wsum, const T *w) {
  using A = Ar<T>;
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300}, smx = 0, smn = 1e300, ws = 0;
  unsigned long long bad = 0;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    T a = x[i], b = y[i], c = z[i];
    if (do_rescale) { a = A::mul(a, rescale); b = A::mul(b, rescale); c = A::mul(c, rescale); x[i] = a; y[i] = b; z[i] = c; }
    T ss;
    if (sumsq == 0) ss = A::add(A::add(A::mul(a, a), A::mul(b, b)), A::mul(c, c));
    else if (sumsq == 1) ss = A::fma(c, c, A::fma(b, b, A::mul(a, a)));
    else ss = s ? s[i] : A::add(A::add(A::mul(a, a), A::mul(b, b)), A::mul(c, c));
    if (sumsq >= 0 && s) s[i] = ss;
    if (!(isfinite((double) a) && isfinite((double) b) && isfinite((double) c))) { bad++; continue; }
    mn[0] = fmin(mn[0], (double) a); mn[1] = fmin(mn[1], (double) b); mn[2] = fmin(mn[2], (double) c);
    mx[0] = fmax(mx[0], (double) a); mx[1] = fmax(mx[1], (double) b); mx[2] = fmax(mx[2], (double) c);
    smx = fmax(smx, (double) ss); smn = fmin(smn, (double) ss);
    if (w) ws += (double) w[i];
  }
  typedef cub::BlockReduce<double, 256> BR;
  typedef cub::BlockReduce<unsigned long long, 256> BRU;
  __shared__ typename BR::TempStorage tmp;
  __shared__ typename BRU::TempStorage tmpu;
  for (int k = 0; k < 3; k++) {
    double v = BR(tmp).Reduce(mn[k], cub::Min()); __syncthreads();
    if (threadIdx.x == 0) atomicMin(&stats[k], enc_f64(v));
    v = BR(tmp).Reduce(mx[k], cub::Max()); __syncthreads();
    if (threadIdx.x == 0) atomicMax(&stats[3 + k], enc_f64(v));
  }
  double v = BR(tmp).Reduce(smx, cub::Max()); __syncthreads();
  if (threadIdx.x == 0) atomicMax(&stats[6], enc_f64(v));
  v = BR(tmp).Reduce(smn, cub::Min()); __syncthreads();
  if (threadIdx.x == 0) atomicMin(&stats[9], enc_f64(v));
  v = BR(tmp).Sum(ws); __syncthreads();
  if (threadIdx.x == 0 && w) atomicAdd(wsum, v);
  unsigned long long nb = BRU(tmpu).Sum(bad);
  if (threadIdx.x == 0 && nb) atomicAdd(&stats[7], nb);
}

// cell id of every point (double arithmetic so that float and double catalogues bin identically)
template <class T>
__global__ void cellid_kernel(const T *x, const T *y, const T *z, int n, Grid g, int sub, unsigned int *key, int *idx, int *err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {(double) x[i], (double) y[i], (double) z[i]};
  int c[3];
  unsigned int q[3];            // position inside the cell in quarters (0..3)
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double u = (p[k] - g.origin[k]) / g.cs[k];
    int ci = (int) floor(u);
    // one period of the box, wherever it starts (the reference never range-checks: only differences and +-L shifts enter
    // the metric, so a box stored as [-L/2, L/2] is as good as [0, L]); x == origin + L happens after rounding to float
    if (g.periodic && !(p[k] >= g.origin[k] && p[k] <= g.origin[k] + g.box[k])) atomicExch(err, 1);
    c[k] = min(max(ci, 0), g.nc[k] - 1);
    q[k] = (unsigned int) min(max((int) ((u - c[k]) * 4.0), 0), 3);
  }
  // Sort key: the cell, then (low bits) a Morton code of the position inside the cell.  The counting kernel hands
  // lane l of a tile the points l, l+32, l+64, ...: with the points of a cell in Morton order every lane gets
  // primaries from different parts of the cell, which evens out the rate at which the lanes accept pairs.
  unsigned int m = 0;
#pragma unroll
  for (int b = 1; b >= 0; b--) m = (m << 3) | (((q[0] >> b) & 1u) << 2) | (((q[1] >> b) & 1u) << 1) | ((q[2] >> b) & 1u);
  key[i] = ((unsigned int) ((c[0] * g.nc[1] + c[1]) * g.nc[2] + c[2]) << sub) | (m >> (6 - sub));
  idx[i] = i;
}

template <class T>
__global__ void gather_kernel(const T *x, const T *y, const T *z, const T *s, const T *w, const int *idx, int n,
                              Vec4<T> *pos, T *wout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = idx[i];
  Vec4<T> v; v.x = x[j]; v.y = y[j]; v.z = z[j]; v.s = s ? s[j] : (T) 0;
  pos[i] = v;
  if (wout) wout[i] = w ? w[j] : (T) 1;
}

// cell_start[c] = first sorted index with key >= c (keys sorted ascending), cell_start[ncell] = n
__global__ void cellstart_kernel(const unsigned int *key, int n, int ncell, int sub, int *cell_start) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int prev = (i == 0) ? -1 : (int) (key[i - 1] >> sub);
  int cur = (i == n) ? ncell : (int) (key[i] >> sub);
  for (int c = prev + 1; c <= cur; c++) cell_start[c] = i;
}

// number of tiles per cell (balanced split of the cell into ceil(n/tile) items)
__global__ void ntile_kernel(const int *cell_start, int ncell, int tile, int *ntile) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int n = cell_start[c + 1] - cell_start[c];
  ntile[c] = (n + tile - 1) / tile;
}
__global__ void items_kernel(const int *cell_start, const int *tile_off, int ncell, int tile,
                             int *item_cell, int *item_off, int *item_cnt) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int b = cell_start[c], n = cell_start[c + 1] - b;
  if (n == 0) return;
  int nt = (n + tile - 1) / tile, o = tile_off[c];
  int base = n / nt, rem = n % nt, p = b;
  for (int t = 0; t < nt; t++) {
    int m = base + (t < rem ? 1 : 0);
    item_cell[o + t] = c; item_off[o + t] = p; item_cnt[o + t] = m;
    p += m;
  }
}

// Estimated cost of every work item: its points times the secondary points its stencil sweeps (same range logic
// as count_kernel).  Items are then processed longest first, and shards take every nparts-th item of that
// order, which balances clustered catalogues across warps and across GPUs.
__global__ void item_cost_kernel(const int *item_cell, const int *item_cnt, int ntile, int nsplit, const int *cell_start2,
                                 const int4 *rows, int nrows, int ncx, int ncy, int ncz, int periodic, int isauto,
                                 unsigned int keep_mask, float *cost, int *index) {
  int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= ntile * nsplit) return;
  const int tile = it / nsplit, split = it - tile * nsplit;
  const int cell = item_cell[tile];
  const int iz = cell % ncz, iy = (cell / ncz) % ncy, ix = cell / (ncz * ncy);
  const int nq = (periodic ? 3 : 1) * nrows, qfirst = isauto ? -1 : 0;
  const int qlo = qfirst + (int) ((long long) (nq - qfirst) * split / nsplit);
  const int qhi = qfirst + (int) ((long long) (nq - qfirst) * (split + 1) / nsplit);
  long long tot = 0;
  for (int q = qlo; q < qhi; q++) {
    if (q < 0) { tot += (cell_start2[cell + 1] - cell_start2[cell]) / 2; continue; }
    const int ri = periodic ? q / 3 : q, img = periodic ? q - 3 * ri : 1;
    const int4 row = rows[ri];
    int jx = ix + row.x, jy = iy + row.y;
    int zlo = iz + row.z, zhi = iz + row.w;
    if (periodic) {
      jx = (jx % ncx + ncx) % ncx; jy = (jy % ncy + ncy) % ncy;
      if (img == 0) { zhi = min(zhi, -1) + ncz; zlo += ncz; }
      else if (img == 1) { zlo = max(zlo, 0); zhi = min(zhi, ncz - 1); }
      else { zlo = max(zlo, ncz) - ncz; zhi -= ncz; }
    } else {
      if (jx < 0 || jx >= ncx || jy < 0 || jy >= ncy) continue;
      zlo = max(zlo, 0); zhi = min(zhi, ncz - 1);
    }
    if (zlo > zhi) continue;
    const int rowbase = (jx * ncy + jy) * ncz;
    tot += cell_start2[rowbase + zhi + 1] - cell_start2[rowbase + zlo];
  }
  // coarse cost classes (exponent + a few mantissa bits): the stable sort keeps the cell order inside a class,
  // so that warps working at the same time sweep neighbouring cells (L2 reuse of the secondary catalogue)
  cost[it] = __uint_as_float(__float_as_uint((float) tot * (float) item_cnt[tile] + 1.0f) & keep_mask);
  index[it] = it;
}

// ------------------------------------------------------------------------------------------
// The cell-sorted copy of a catalogue for grid g (built on first use, then kept: see DevCat).
template <class T>
static int build_sorted(DevCat *cat, const Grid &g, int tile, bool need_w, const Options &opt, float *ms_out, Sorted<T> **out) {
  std::vector<Sorted<T>> &all = sorted_of<T>(cat);
  for (auto &S : all)
    if (S.valid && S.grid == g && S.tile == tile && (!need_w || S.w)) { S.stamp = ++cat->clock; *out = &S; return 0; }
  // evict the least recently used copies beyond the cap (and stale ones without weights for this grid)
  for (size_t i = 0; i < all.size();)
    if (!all[i].valid || (all[i].grid == g && all[i].tile == tile)) { all[i].release(); all.erase(all.begin() + i); } else i++;
  while ((int) all.size() >= std::max(1, opt.sorted_copies)) {
    size_t lru = 0;
    for (size_t i = 1; i < all.size(); i++) if (all[i].stamp < all[lru].stamp) lru = i;
    all[lru].release(); all.erase(all.begin() + lru);
  }
  Sorted<T> S;
  const int n = (int) cat->n;
  const long long ncell = g.ncell();
  EventScope ev(2);
  PoolScope pool;
  cudaEventRecord(ev[0]);
  unsigned int *key = nullptr, *key2 = nullptr; int *idx = nullptr, *idx2 = nullptr, *err = nullptr;
  int *ntile = nullptr, *tile_off = nullptr; void *tmp = nullptr;
#define TRY_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_err("%s: %s", #x, cudaGetErrorString(e_)); cudaGetLastError(); return FCFC_GPU_ERR_TREE; } } while (0)
  const size_t nn = n ? n : 1;
  TRY_(pool.alloc(&key, nn * 4)); TRY_(pool.alloc(&key2, nn * 4)); TRY_(pool.alloc(&idx, nn * 4)); TRY_(pool.alloc(&idx2, nn * 4));
  TRY_(pool.alloc(&err, 4)); TRY_(cudaMemset(err, 0, 4));
  TRY_(pool.alloc(&S.pos, nn * sizeof(Vec4<T>)));
  if (need_w) TRY_(pool.alloc(&S.w, nn * sizeof(T)));
  TRY_(pool.alloc(&S.cell_start, (ncell + 1) * sizeof(int)));
  const int nb = (n + 255) / 256;
  int bits = 1; while ((1ll << bits) < ncell) bits++;
  const int sub = opt.no_subsort ? 0 : std::min(6, 31 - bits);   // Morton bits inside the cell
  if (n) {
    cellid_kernel<T><<<nb, 256>>>((const T *) cat->x, (const T *) cat->y, (const T *) cat->z, n, g, sub, key, idx, err);
    g_stats.kernel_launches++;
    bits += sub;
    size_t tb = 0;
    TRY_(cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, idx, idx2, n, 0, bits));
    TRY_(pool.alloc(&tmp, tb ? tb : 1));
    TRY_(cub::DeviceRadixSort::SortPairs(tmp, tb, key, key2, idx, idx2, n, 0, bits));
    gather_kernel<T><<<nb, 256>>>((const T *) cat->x, (const T *) cat->y, (const T *) cat->z,
                                  cat->has_s ? (const T *) cat->s : nullptr, cat->has_w ? (const T *) cat->w : nullptr,
                                  idx2, n, S.pos, S.w);
    g_stats.kernel_launches += 2;
  }
  cellstart_kernel<<<(n + 1 + 255) / 256, 256>>>(key2, n, (int) ncell, sub, S.cell_start);
  g_stats.kernel_launches++;
  int herr = 0;
  TRY_(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  if (herr) {
    set_err("catalogue has points outside one period of the box: FCFC_2PT_BOX expects max - min <= BOX_SIZE on every axis");
    return FCFC_GPU_ERR_DATA;
  }
  // work items
  TRY_(pool.alloc(&ntile, (ncell + 1) * sizeof(int))); TRY_(pool.alloc(&tile_off, (ncell + 1) * sizeof(int)));
  TRY_(cudaMemset(ntile, 0, (ncell + 1) * sizeof(int)));
  ntile_kernel<<<(int) ((ncell + 255) / 256), 256>>>(S.cell_start, (int) ncell, tile, ntile);
  pool.free_now(tmp); tmp = nullptr;
  size_t tb = 0;
  TRY_(cub::DeviceScan::ExclusiveSum(nullptr, tb, ntile, tile_off, (int) ncell + 1));
  TRY_(pool.alloc(&tmp, tb ? tb : 1));
  TRY_(cub::DeviceScan::ExclusiveSum(tmp, tb, ntile, tile_off, (int) ncell + 1));
  int nitem = 0;
  TRY_(cudaMemcpy(&nitem, tile_off + ncell, 4, cudaMemcpyDeviceToHost));
  S.nitem = nitem;
  const size_t ni = nitem ? nitem : 1;
  TRY_(pool.alloc(&S.item_cell, ni * 4)); TRY_(pool.alloc(&S.item_off, ni * 4)); TRY_(pool.alloc(&S.item_cnt, ni * 4));
  items_kernel<<<(int) ((ncell + 255) / 256), 256>>>(S.cell_start, tile_off, (int) ncell, tile, S.item_cell, S.item_off, S.item_cnt);
  g_stats.kernel_launches += 3;
  TRY_(cudaGetLastError());
  cudaEventRecord(ev[1]); TRY_(cudaEventSynchronize(ev[1]));
  float ms = 0; cudaEventElapsedTime(&ms, ev[0], ev[1]); if (ms_out) *ms_out += ms;
#undef TRY_
  // success: the sorted copy keeps its blocks, everything else goes back to the pool with `pool`
  pool.keep(S.pos); pool.keep(S.w); pool.keep(S.cell_start); pool.keep(S.item_cell); pool.keep(S.item_off); pool.keep(S.item_cnt);
  S.valid = true; S.grid = g; S.tile = tile; S.stamp = ++cat->clock;
  all.push_back(S);
  *out = &all.back();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Search radii.  The cell sweep must cover every pair the per-pair tests can accept, including
// pairs that only pass because of rounding in `real` arithmetic, hence the small inflation.
struct Reach { double r2_xy, r_z; bool cylinder; double r2; };

static Reach compute_reach(const fcfc_gpu_bins *b, double s2max, double pmax, double maxabs, double smax_sq, bool is_float) {
  const double eps = is_float ? FLT_EPSILON : DBL_EPSILON;
  Reach R;
  R.cylinder = (b->periodic && b->bintype == FCFC_GPU_BIN_SPI);
  double slack;
  if (b->periodic || b->bintype == FCFC_GPU_BIN_ISO) slack = 64 * eps * maxabs * maxabs;
  else slack = 64 * eps * (smax_sq + 1);          // survey: s^2 by cancellation of |x|^2-sized terms
  if (R.cylinder) {
    R.r2_xy = s2max * (1 + 1e-6) + slack;
    R.r_z = pmax * (1 + 1e-6) + 64 * eps * maxabs;
    R.r2 = R.r2_xy + R.r_z * R.r_z;
  } else {
    double r2 = s2max;
    if (!b->periodic && b->bintype == FCFC_GPU_BIN_SPI) r2 = s2max + pmax;   // pmax is pi^2 here (count_func.c:5296)
    R.r2 = r2 * (1 + 1e-6) + slack;
    R.r2_xy = R.r2; R.r_z = std::sqrt(R.r2);
  }
  return R;
}

// Stencil rows (dx, dy, zlo, zhi) of the cells that can hold a pair within reach.  For auto counts
// only the lexicographically positive half is kept (every unordered cell pair once).
static std::vector<int4> build_stencil(const Grid &g, const Reach &R, bool half) {
  std::vector<int4> rows;
  auto gap = [](int d, double cs) { int a = std::abs(d) - 1; return a > 0 ? a * cs : 0.0; };
  const double rxy = std::sqrt(R.r2_xy);
  const int kx = (int) std::ceil(rxy / g.cs[0]) + 1, ky = (int) std::ceil(rxy / g.cs[1]) + 1;
  for (int dx = -kx; dx <= kx; dx++)
    for (int dy = -ky; dy <= ky; dy++) {
      if (half && (dx < 0 || (dx == 0 && dy < 0))) continue;
      double gx = gap(dx, g.cs[0]), gy = gap(dy, g.cs[1]);
      double d2 = gx * gx + gy * gy;
      if (d2 >= R.r2_xy) continue;
      int zmax = 0;
      while (true) {
        double gz = gap(zmax + 1, g.cs[2]);
        bool ok = R.cylinder ? (gz < R.r_z) : (d2 + gz * gz < R.r2);
        if (!ok) break;
        zmax++;
      }
      // NB: no clamping to nc/2 on periodic axes: offsets d and d - nc address the same cell through
      // *different* periodic images, and both can hold in-range pairs; the gap test alone prunes.
      int zlo = -zmax, zhi = zmax;
      if (half && dx == 0 && dy == 0) { zlo = 1; if (zhi < zlo) continue; }
      rows.push_back(make_int4(dx, dy, zlo, zhi));
    }
  return rows;
}

// Choose the number of cells: cells of side reach/k, k picked by a simple cost model
// (candidate evaluations per primary point, tile fill of the primary cells, per-range overhead).
// For every stencil row the sub-range of dz whose cells lie entirely within the maximum separation of every point of
// the tile's cell: the largest distance between a point of cell 0 and a point of cell (dx, dy, dz) is the norm of
// ((|dx|+1) csx, (|dy|+1) csy, (|dz|+1) csz).  A small margin keeps points that sit on a cell face out of doubt; the
// kernel keeps its range test anyway (count_kernel.cuh, do_chunk_dense).  Empty ranges are (1, 0).
static std::vector<int2> stencil_inside(const Grid &g, const std::vector<int4> &rows, double s2max) {
  std::vector<int2> in(rows.size());
  const double lim = s2max * (1.0 - 1e-4);
  for (size_t i = 0; i < rows.size(); i++) {
    const double ex = (std::abs(rows[i].x) + 1) * g.cs[0], ey = (std::abs(rows[i].y) + 1) * g.cs[1];
    int m = -1;                 // largest |dz| that is inside
    while (true) {
      const double ez = (m + 2) * g.cs[2];
      if (ex * ex + ey * ey + ez * ez < lim) m++; else break;
    }
    const int lo = std::max(-m, rows[i].z), hi = std::min(m, rows[i].w);
    in[i] = (m >= 0 && lo <= hi) ? make_int2(lo, hi) : make_int2(1, 0);
  }
  return in;
}

}  // namespace fcfc
// Diagnostics for the CPU tests (tests/test_stencil.py): the neighbour stencil and its dense sub-ranges for cells of size
// cs[3], a spherical reach r2 (squared) and the maximum separation s2max, exactly as the counting path builds them.
// rows_out[4 i .. 4 i + 3] = (dx, dy, dz_lo, dz_hi), inside_out[2 i .. 2 i + 1] = (dz_lo, dz_hi) of the dense cells
// ((1, 0) = none).  Returns the number of rows (nothing is written beyond max_rows).
extern "C" int fcfc_gpu_debug_stencil(const double cs[3], double r2, double s2max, int half, int *rows_out, int *inside_out, int max_rows) {
  fcfc::Grid g;
  for (int d = 0; d < 3; d++) g.cs[d] = cs[d];
  fcfc::Reach R;
  R.cylinder = false; R.r2 = r2; R.r2_xy = r2; R.r_z = std::sqrt(r2);
  const std::vector<int4> rows = fcfc::build_stencil(g, R, half != 0);
  const std::vector<int2> in = fcfc::stencil_inside(g, rows, s2max);
  for (size_t i = 0; i < rows.size() && (int) i < max_rows; i++) {
    rows_out[4 * i] = rows[i].x; rows_out[4 * i + 1] = rows[i].y; rows_out[4 * i + 2] = rows[i].z; rows_out[4 * i + 3] = rows[i].w;
    inside_out[2 * i] = in[i].x; inside_out[2 * i + 1] = in[i].y;
  }
  return (int) rows.size();
}
namespace fcfc {

static Grid choose_grid(const fcfc_gpu_bins *b, const Reach &R, const double lo[3], const double hi[3],
                        double n1, double n2, int tile, bool half, const Options &opt) {
  Grid best; double best_cost = 1e300;
  const double rxy = std::sqrt(R.r2_xy), rz = R.r_z;
  int kmin = 1, kmax = 12;
  if (opt.k > 0) kmin = kmax = opt.k;     // forced reach / cell
  for (int k = kmin; k <= kmax; k++) {
    Grid g; g.periodic = b->periodic;
    bool ok = true;
    for (int d = 0; d < 3; d++) {
      const double want = ((d == 2) ? rz : rxy) / k;
      if (b->periodic) {
        const double L = b->bsize[d];
        int nc = (int) std::floor(L / want);
        if (nc < 1) nc = 1;
        while (nc > 1 && L / nc < want) nc--;
        // the period starts at 0 when the catalogues live in [0, L] (the usual case: one grid for every pair of
        // catalogues), otherwise at their common minimum
        g.nc[d] = nc; g.origin[d] = (lo[d] >= 0 && hi[d] <= L) ? 0 : lo[d]; g.cs[d] = L / nc; g.box[d] = L;
      } else {
        const double ext = std::max(hi[d] - lo[d], 1e-30);
        int nc = (int) std::floor(ext / want) + 1;
        g.nc[d] = nc; g.origin[d] = lo[d]; g.cs[d] = want;
      }
      if (g.nc[d] > 2048) ok = false;
    }
    if (k == kmin && (!ok || g.ncell() > (1ll << 26))) {
      // tiny reach in a huge volume: even cells of one reach are too many.  Coarser cells are always valid
      // (the stencil is derived from the actual cell size).
      double shrink = 1.0;
      for (int it = 0; it < 200 && (!ok || g.ncell() > (1ll << 26)); it++) {
        shrink *= 0.9; ok = true;
        for (int d = 0; d < 3; d++) {
          const double ext = b->periodic ? b->bsize[d] : std::max(hi[d] - lo[d], 1e-30);
          const double want = ((d == 2) ? rz : rxy) / k;
          int nc = std::max(1, std::min(2048, (int) std::floor(ext / want * shrink) + (b->periodic ? 0 : 1)));
          g.nc[d] = nc;
          if (b->periodic) g.cs[d] = ext / nc; else g.cs[d] = std::max(want, ext / nc * (1 + 1e-9));
        }
      }
    }
    if (!ok || g.ncell() > (1ll << 26)) break;
    const double ncell = (double) g.ncell();
    const double m1 = std::max(n1 / ncell, 1e-3), m2 = std::max(n2 / ncell, 1e-3);
    std::vector<int4> st = build_stencil(g, R, half);
    double cand = 0, nseg = 0;
    for (auto &r : st) { double pts = (r.w - r.z + 1) * m2; cand += std::ceil(pts / 32) * 32 + 24; nseg += 1; }
    if (half) cand += m2 / 2 + 24;
    const double tiles = std::ceil(m1 / tile);
    const double fill = m1 / (tiles * tile);
    const double cost = cand / std::min(1.0, fill + 1e-9) + 1e-3 * ncell / std::max(n1, 1.0);
    if (g_verbose > 1) fprintf(stderr, "[fcfc_gpu] k=%d nc=%dx%dx%d rows=%zu m1=%.1f cand=%.0f fill=%.2f cost=%.0f\n", k, g.nc[0], g.nc[1], g.nc[2], st.size(), m1, cand, fill, cost);
    if (cost < best_cost && st.size() <= (size_t) kMaxRows) { best_cost = cost; best = g; }
    if (m1 < 4 && m2 < 4) break;
  }
  return best;
}

static int table_is_sqrt(const void *tab, int width, long n) {
  if (!tab || n <= 0 || n > (1 << 18)) return 0;
  for (long i = 0; i < n; i++) {
    long v = width ? ((const uint16_t *) tab)[i] : ((const uint8_t *) tab)[i];
    long r = (long) std::floor(std::sqrt((double) i)); while (r * r > i) r--; while ((r + 1) * (r + 1) <= i) r++;
    if (v != r) return 0;
  }
  return 1;
}

// ------------------------------------------------------------------------------------------
template <class T>
static int count_impl(DevCat *c1, DevCat *c2, const fcfc_gpu_bins *b, int isauto, int withwt,
                      int part, int nparts, int64_t *cnt_i, double *cnt_d, void *dev_hist) {
  constexpr bool is_float = sizeof(T) == 4;
  const int bintype = b->bintype;
  const int ns = b->ns, np = (bintype == BIN_SPI) ? b->np : 0, nmu = (bintype == BIN_SMU) ? b->nmu : 1;
  const size_t ntot = (size_t) ns * (bintype == BIN_ISO ? 1 : (bintype == BIN_SMU ? nmu : np));
  if (ns < 1 || ntot < 1 || !b->s2bin || !b->stab) { set_err("invalid bins"); return FCFC_GPU_ERR_ARG; }
  if (bintype == BIN_SPI && (!b->pbin || !b->ptab || np < 1)) { set_err("(s_perp, pi) bins need pbin/ptab"); return FCFC_GPU_ERR_ARG; }
  if (bintype == BIN_SMU && (!b->mutab || nmu < 1)) { set_err("(s, mu) bins need mutab"); return FCFC_GPU_ERR_ARG; }
  if (!b->periodic && bintype != BIN_ISO && (!c1->has_s || !c2->has_s)) { set_err("survey (s,mu)/(s_perp,pi) counts need the sum of squares (x2sum)"); return FCFC_GPU_ERR_ARG; }
  const T *s2bin = (const T *) b->s2bin, *pbin = (const T *) b->pbin;
  const double s2min = s2bin[0], s2max = s2bin[ns];
  const double pmin = np ? pbin[0] : 0, pmax = np ? pbin[np] : 0;
  // table lengths: util/create_lut.c:62-64 (integer) and :105-107 (hybrid)
  auto tablen = [&](const T *e, int n) -> long {
    long mn = (long) e[0];
    long mx = (b->tabtype == FCFC_GPU_TAB_INT) ? (long) e[n] : (long) std::ceil((double) e[n]);
    return mx - mn;
  };
  const long nstab = b->nstab ? (long) b->nstab : tablen(s2bin, ns);
  const long nptab = np ? (b->nptab ? (long) b->nptab : tablen(pbin, np)) : 0;
  if (nstab <= 0 || nstab > (1 << 20) || nptab < 0 || nptab > (1 << 20)) { set_err("invalid lookup table length"); return FCFC_GPU_ERR_ARG; }

  memset(&g_stats, 0, sizeof g_stats);

  // ---- grid, cell lists, stencil ----
  double lo[3], hi[3], maxabs = 0;
  for (int d = 0; d < 3; d++) {
    lo[d] = std::min(c1->bmin[d], c2->bmin[d]); hi[d] = std::max(c1->bmax[d], c2->bmax[d]);
    maxabs = std::max(maxabs, std::max(std::fabs(lo[d]), std::fabs(hi[d])));
    if (b->periodic) maxabs = std::max(maxabs, 2 * b->bsize[d]);
  }
  if (b->periodic) {
    for (int d = 0; d < 3; d++) if (!(b->bsize[d] > 0)) { set_err("invalid box size"); return FCFC_GPU_ERR_ARG; }
    const double rmax = std::sqrt(s2max);
    for (int d = 0; d < 3; d++)
      if (rmax >= 0.5 * b->bsize[d] || (bintype == BIN_SPI && pmax >= 0.5 * b->bsize[d])) {
        set_err("the maximum separation must be smaller than half the box size"); return FCFC_GPU_ERR_ARG;
      }
  }
  const Reach R = compute_reach(b, s2max, pmax, maxabs, std::max(c1->smax, c2->smax), is_float);
  const Options opt = options_snapshot();
  // tables that are pure functions of the index are computed in registers instead of looked up
  const int mu_is_sqrt = (bintype == BIN_SMU) ? table_is_sqrt(b->mutab, 0, (long) nmu * nmu) : 0;
  const int stab_is_sqrt = (b->tabtype == FCFC_GPU_TAB_INT && s2bin[0] == 0) ? table_is_sqrt(b->stab, b->swidth, nstab) : 0;
  // counts that will take the classified-staging kernel (count_kernel_cl.cuh; the shared-memory plan confirms it below)
  // get tiles of 32 kClR points: fewer registers for the tile, measured 2 % faster on the bench workload
  const bool cl_candidate = is_float && !withwt && bintype != BIN_SPI && (b->periodic || bintype == BIN_ISO) && s2bin[0] == 0 && b->swidth == 0 &&
                            b->tabtype == FCFC_GPU_TAB_INT && stab_is_sqrt && (bintype == BIN_ISO || mu_is_sqrt) &&
                            !opt.no_table_math && !opt.no_dense && !opt.no_classify && !opt.force_generic && !opt.global_hist;
  const int tile = 32 * (cl_candidate ? kClR : kR);    // (tiles of 96 points were measured for count_kernel_df too: 468.5 -> 474.6 ms)
  const bool half = isauto != 0;
  if (c1->n == 0 || c2->n == 0) {
    if (withwt) { if (cnt_d) memset(cnt_d, 0, ntot * 8); } else if (cnt_i) memset(cnt_i, 0, ntot * 8);
    if (dev_hist) cudaMemset(dev_hist, 0, ntot * 8);
    return 0;
  }
  const Grid g = choose_grid(b, R, lo, hi, (double) c1->n, (double) c2->n, tile, half, opt);
  if (g.cs[0] <= 0) { set_err("failed to choose a cell grid"); return FCFC_GPU_ERR_TREE; }
  EventScope evs(4);            // released on every exit path, like the pool blocks of `pool`
  PoolScope pool;
  cudaEventRecord(evs[0]);
  float ms_sort = 0;
  Sorted<T> *pS1 = nullptr, *pS2 = nullptr;
  int e = build_sorted<T>(c1, g, tile, withwt != 0, opt, &ms_sort, &pS1);
  if (e) return e;
  if (c2 != c1) { e = build_sorted<T>(c2, g, tile, withwt != 0, opt, &ms_sort, &pS2); if (e) return e; } else pS2 = pS1;
  Sorted<T> &S1 = *pS1, &S2 = *pS2;
  std::vector<int4> rows = build_stencil(g, R, half);
  if (rows.size() > (size_t) kMaxRows) { set_err("neighbour stencil too large (%zu rows)", rows.size()); return FCFC_GPU_ERR_TREE; }

  // ---- device copies of the tables ----
  const int sw = b->swidth ? 2 : 1, pw = b->pwidth ? 2 : 1;
  const size_t sz_stab = (size_t) nstab * sw, sz_ptab = (size_t) nptab * pw, sz_mu = (bintype == BIN_SMU) ? (size_t) nmu * nmu : 0;
  const size_t sz_s2 = (ns + 1) * sizeof(T), sz_pb = np ? (np + 1) * sizeof(T) : 0, sz_rows = rows.size() * sizeof(int4);
  auto al = [](size_t v) { return (v + 255) & ~(size_t) 255; };
  const size_t sz_rin = rows.size() * sizeof(int2);
  const size_t o_stab = 0, o_ptab = o_stab + al(sz_stab), o_mu = o_ptab + al(sz_ptab), o_s2 = o_mu + al(sz_mu),
               o_pb = o_s2 + al(sz_s2), o_rows = o_pb + al(sz_pb), o_rin = o_rows + al(sz_rows), o_hist = o_rin + al(sz_rin),
               o_cnt = o_hist + al(ntot * 8), o_end = o_cnt + 256;
  std::vector<unsigned char> hbuf(o_end, 0);
  memcpy(&hbuf[o_stab], b->stab, sz_stab);
  if (sz_ptab) memcpy(&hbuf[o_ptab], b->ptab, sz_ptab);
  if (sz_mu) memcpy(&hbuf[o_mu], b->mutab, sz_mu);
  memcpy(&hbuf[o_s2], s2bin, sz_s2);
  if (sz_pb) memcpy(&hbuf[o_pb], pbin, sz_pb);
  if (sz_rows) memcpy(&hbuf[o_rows], rows.data(), sz_rows);
  int dense_rows = 0;
  if (sz_rin && bintype != BIN_SPI) {
    const std::vector<int2> rin = stencil_inside(g, rows, s2max);
    memcpy(&hbuf[o_rin], rin.data(), sz_rin);
    for (const int2 &r : rin) dense_rows += r.x <= r.y;
  }
  unsigned char *dbuf = nullptr;
  CUDA_TRY(pool.alloc(&dbuf, o_end), FCFC_GPU_ERR_MEMORY);
  CUDA_TRY(cudaMemcpy(dbuf, hbuf.data(), o_end, cudaMemcpyHostToDevice), FCFC_GPU_ERR_CUDA);

  // ---- kernel parameters ----
  CountParams<T> P;
  memset(&P, 0, sizeof P);
  P.pos1 = S1.pos; P.w1 = S1.w; P.pos2 = S2.pos; P.w2 = S2.w; P.cell_start2 = S2.cell_start;
  P.item_cell = S1.item_cell; P.item_off = S1.item_off; P.item_cnt = S1.item_cnt;
  // longest-first order of the work items (cost depends on the secondary catalogue and the stencil); shard `part` of
  // `nparts` takes the items part, part + nparts, ... of that order (count_kernel.cuh: persistent warp loop)
  int *d_order = nullptr, nsplit = 1;
  {
    // small problems: cut every tile's sweep list so that each warp of the grid still gets several work items
    const long long warps_total = (long long) g_ctx.sm_count * BlockShape<T>::kWarps;
    const int nq_all = (int) rows.size() * (b->periodic ? 3 : 1) + (isauto ? 1 : 0);
    const long long ipw = std::max(1, opt.items_per_warp);
    nsplit = (int) std::min<long long>(std::max<long long>(1, (ipw * warps_total * nparts + S1.nitem - 1) / std::max(S1.nitem, 1)), std::max(nq_all, 1));
    if (opt.nsplit > 0) nsplit = std::max(1, std::min(opt.nsplit, std::max(nq_all, 1)));
    const int ni = S1.nitem * nsplit;
    float *cost = nullptr, *cost2 = nullptr; int *idx = nullptr; void *tmp = nullptr;
    auto fail = [&](const char *what) { set_err("%s", what); return FCFC_GPU_ERR_TREE; };
    if (pool.alloc(&cost, (size_t) ni * 4) || pool.alloc(&cost2, (size_t) ni * 4) || pool.alloc(&idx, (size_t) ni * 4) ||
        pool.alloc(&d_order, (size_t) ni * 4)) return fail("out of device memory for the work-item order");
    const int cost_bits = std::max(0, std::min(23, opt.cost_bits));   // mantissa bits of the cost classes (23 = exact cost order)
    const unsigned int keep_mask = 0xffffffffu << (23 - cost_bits);
    if (ni) {
      item_cost_kernel<<<(ni + 255) / 256, 256>>>(S1.item_cell, S1.item_cnt, S1.nitem, nsplit, S2.cell_start, reinterpret_cast<const int4 *>(dbuf + o_rows),
                                                  (int) rows.size(), g.nc[0], g.nc[1], g.nc[2], b->periodic, isauto, keep_mask, cost, idx);
      g_stats.kernel_launches++;
      size_t tb = 0;
      if (cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, cost, cost2, idx, d_order, ni) != cudaSuccess) return fail("cub sort failed");
      if (pool.alloc(&tmp, tb ? tb : 1)) return fail("out of device memory for the work-item order");
      if (cub::DeviceRadixSort::SortPairsDescending(tmp, tb, cost, cost2, idx, d_order, ni) != cudaSuccess) return fail("cub sort failed");
    }
    pool.free_now(cost); pool.free_now(cost2); pool.free_now(idx); pool.free_now(tmp);
  }
  P.item_order = d_order; P.nitem = S1.nitem * nsplit; P.nsplit = nsplit; P.part = part; P.nparts = nparts;
  P.work_counter = reinterpret_cast<unsigned int *>(dbuf + o_cnt);
  for (int d = 0; d < 3; d++) { P.nc[d] = g.nc[d]; P.bsize[d] = (T) b->bsize[d]; }
  P.periodic = b->periodic;
  P.rows = reinterpret_cast<const int4 *>(dbuf + o_rows); P.nrows = (int) rows.size();
  P.s2min = (T) s2min; P.s2max = (T) s2max; P.pmin = (T) pmin; P.pmax = (T) pmax;
  {
    // survey (s_perp, pi): limits of the division-free pre-tests of eval_pair, already rounded up to `real`
    double lim[3];
    fcfc_gpu_survey_pretest_limits(s2max, pmax, is_float ? 1 : 0, lim);
    P.premax = (T) lim[0]; P.pmax_pre = (T) lim[1]; P.s2max_pre = (T) lim[2];
  }
  P.nmu2 = nmu * nmu; P.nmu2f = (T) (nmu * nmu);
  P.ns = ns; P.np = np; P.ntot = (int) ntot;
  P.soff = (int) s2bin[0]; P.poff = np ? (int) pbin[0] : 0;
  P.tab_hybrid = (b->tabtype == FCFC_GPU_TAB_HYBRID); P.swidth = b->swidth; P.pwidth = b->pwidth;
  P.with_mu_one = b->with_mu_one; P.smin0 = (s2bin[0] == 0); P.pmin0 = np ? (pbin[0] == 0) : 1;
  P.stab = dbuf + o_stab; P.ptab = dbuf + o_ptab; P.mutab = dbuf + o_mu;
  P.nstab = (int) nstab; P.nptab = (int) nptab;
  P.s2bin = reinterpret_cast<const T *>(dbuf + o_s2); P.pbin = reinterpret_cast<const T *>(dbuf + o_pb);
  P.isauto = isauto;
  P.ghist_i = reinterpret_cast<unsigned long long *>(dbuf + o_hist);
  P.ghist_d = reinterpret_cast<double *>(dbuf + o_hist);
  P.gevals = reinterpret_cast<unsigned long long *>(dbuf + o_cnt + 8);

  const bool generic = !(P.smin0 && P.pmin0 && !P.tab_hybrid && b->swidth == 0 && (np == 0 || b->pwidth == 0));
  Variant v;
  v.is_float = is_float; v.bintype = bintype; v.box = b->periodic != 0; v.wt = withwt != 0;
  v.arith = b->arith ? 1 : 0; v.generic = generic;
  // shared-memory budget decides whether the histogram lives in shared memory
  int dev = 0; cudaGetDevice(&dev);
  int smem_max = 0; cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int nmutab = (bintype == BIN_SMU) ? nmu * nmu : 0;
  // tables that are pure functions of the index are computed in registers instead of looked up
  {
    P.mu_is_sqrt = mu_is_sqrt;
    P.stab_is_sqrt = stab_is_sqrt;
    P.ptab_is_ident = 0;
    if (bintype == BIN_SPI && b->periodic && b->tabtype == FCFC_GPU_TAB_INT && pbin[0] == 0 && nptab == np) {
      P.ptab_is_ident = 1;
      for (long i = 0; i < nptab; i++) { long vv = b->pwidth ? ((const uint16_t *) b->ptab)[i] : ((const uint8_t *) b->ptab)[i]; if (vv != i) P.ptab_is_ident = 0; }
    }
    if (opt.no_table_math) P.mu_is_sqrt = P.stab_is_sqrt = P.ptab_is_ident = 0;
    // fixed-point scales of the fast bins (count_kernel.cuh, fast_bins): the flag band 2^-k must cover the error
    // budget with a factor >= 2 to spare, and bin * 2^k must stay below 2^23
    int ks = 0, km = 0;
    fcfc_gpu_fastbin_scales(ns, nmu, b->periodic, &ks, &km);
    P.fb_sscale = (float) std::ldexp(1.0, ks); P.fb_mscale = (float) std::ldexp((double) nmu, km);
    P.fb_smask = (1u << ks) - 4u; P.fb_mmask = (1u << km) - 2u; P.fb_sshift = (unsigned) ks; P.fb_mshift = (unsigned) km;
    P.fb_smul = 1u << (32 - ks); P.fb_mmul = 1u << (32 - km);
    P.fb_bias = (0x4B000000u >> ks) + ((bintype == BIN_SMU) ? (0x4B000000u >> km) * (unsigned int) ns : 0u);
    if (ks < 6 || km < 6) P.stab_is_sqrt = 0;   // too many bins for the fixed-point trick: use the exact path
  }
  // accepted-pair queues: the deepest depth (a multiple of 4, <= 64) that fits next to the histogram and tables
  const int qwords = (bintype == BIN_ISO) ? (withwt ? 2 : 1) : (b->periodic ? (withwt ? 4 : 2) : 4);
  // weighted sums with few bins: 32 lane-private copies of the shared histogram, so that the FP64 (CAS) atomics
  // of the lanes of a warp never collide on a bin
  const int hist_copies = (withwt && ntot <= 256 && !opt.no_hist_copies) ? 32 : 1;
  P.hist_copies = hist_copies;
  auto plan = [&](bool sh, int depth, bool tg) {
    return withwt ? make_smem_plan<T, true>((int) ntot, (int) sz_stab, (int) sz_ptab, nmutab, ns, np, (int) rows.size(), sh, qwords, depth, tg, hist_copies)
                  : make_smem_plan<T, false>((int) ntot, (int) sz_stab, (int) sz_ptab, nmutab, ns, np, (int) rows.size(), sh, qwords, depth, tg);
  };
  int qdepth_max = 64;
  if (opt.qdepth > 0) qdepth_max = std::max(8, opt.qdepth & ~3);
  // preference order: everything in shared memory with deep queues > tables in global memory >
  // shallow queues > histogram in global memory (large tables / histograms are rare)
  SmemPlan pl{}; int depth = 0; bool tabs_global = false; v.smem_hist = true;
  const struct { bool sh, tg; int dmin; } tries[] = {{true, false, 16}, {true, true, 16}, {true, false, 8}, {true, true, 8},
                                                     {false, false, 16}, {false, true, 8}};
  // fast variants whose s and mu bins are computed never read the tables in the hot loop: leave them in
  // global memory (the exact re-binning of flagged pairs reads them there) and give the space to the stacks
  const bool fast_variant = !generic && !opt.force_generic && (b->periodic || bintype == BIN_ISO);
  const bool tables_unused = fast_variant && P.stab_is_sqrt && ((bintype == BIN_SMU && P.mu_is_sqrt) || bintype == BIN_ISO);
  const bool force_ghist = opt.global_hist != 0;
  // the packed pair loop (float, box or isotropic, unweighted: PairLoop::kPacked) takes 2 x 4 entries per step and
  // needs that much room above what a drain leaves behind
  const int dmin_variant = (is_float && (b->periodic || bintype == BIN_ISO) && !withwt) ? 12 : 8;
  qdepth_max = std::max(qdepth_max, dmin_variant);
  for (auto &t : tries) {
    if (force_ghist && t.sh) continue;
    if (tables_unused && !t.tg && t.sh) continue;       // computed bins: deeper stacks beat resident tables
    for (int d = qdepth_max; d >= std::max(t.dmin, dmin_variant) && !depth; d -= 4) {
      pl = plan(t.sh, d, t.tg);
      if (pl.total + 1024 <= smem_max) { depth = d; v.smem_hist = t.sh; tabs_global = t.tg; }
    }
    if (depth) break;
  }
  if (!depth) { set_err("shared-memory plan does not fit (%d bytes)", pl.total); return FCFC_GPU_ERR_CF; }
  P.tabs_global = tabs_global;
  if (tabs_global && !tables_unused) v.generic = true;   // otherwise only the generic variant reads tables through global pointers
  if (opt.force_generic) v.generic = true;   // cross-check of the fast paths against the generic one
  // dense cells are binned in place by the packed float pair loop of the variants whose drain computes its bins
  // (count_kernel.cuh: kDense, do_chunk_dense)
  // (periodic boxes only: the survey's isotropic float counts could take it too, but no test exercises it there yet)
  const bool dense = is_float && !withwt && bintype != BIN_SPI && b->periodic && !v.generic && v.smem_hist &&
                     P.stab_is_sqrt && (bintype == BIN_ISO || P.mu_is_sqrt) && sz_rin && !opt.no_dense;
  P.rows_in = dense ? reinterpret_cast<const int2 *>(dbuf + o_rin) : nullptr;
  P.qdepth = depth;
  P.qkeep = (depth >= 32) ? depth / 8 : depth / 4;     // measured on the bench workload (depth 32): 1/8 beats 1/4 and 0; shallow stacks prefer 1/4
  if (opt.qkeep >= 0) P.qkeep = std::max(0, std::min(opt.qkeep, depth / 2));
  P.qkeep = std::max(0, std::min(P.qkeep, depth - 1 - (dmin_variant == 12 ? 8 : 4)));    // a drained stack must have room for the next step
  // double precision, box (s,mu) / isotropic and survey isotropic counts with computed bins: all bulk work in FP32 on
  // cell-relative coordinates, the few pairs within rounding distance of an edge re-evaluated in FP64 (count_kernel_df.cuh)
  bool use_df = false;
  DfPlan dpl{};
  if (!is_float && !opt.no_df && !opt.force_prefilter && !v.generic && v.smem_hist && bintype != BIN_SPI && (b->periodic || bintype == BIN_ISO) &&
      P.stab_is_sqrt && (bintype == BIN_ISO || P.mu_is_sqrt)) {
    int ks = 0, km = 0;
    double d2lim = 0, s1sq = 0, flagged = 0;
    const double cs_max = std::max(g.cs[0], std::max(g.cs[1], g.cs[2]));
    if (fcfc_gpu_df_budget(ns, bintype == BIN_SMU ? nmu : 1, s2max, cs_max, &ks, &km, &d2lim, &s1sq, &flagged)) {
      // (its stacks are small: spare shared memory goes to copies of the weighted histogram, as for the pre-filter kernel below)
      int dcopies = hist_copies;
      for (int hc = (withwt && !opt.no_hist_copies) ? 8 : 1; hc > hist_copies; hc >>= 1)
        if (make_df_plan<true>((int) ntot, ns, (int) rows.size(), hc).total + 1024 <= smem_max) { dcopies = hc; break; }
      dpl = withwt ? make_df_plan<true>((int) ntot, ns, (int) rows.size(), dcopies) : make_df_plan<false>((int) ntot, ns, (int) rows.size(), 1);
      if (dpl.total + 1024 <= smem_max) {
        use_df = true;
        if (withwt) P.hist_copies = dcopies;
        P.fb_sscale = (float) std::ldexp(1.0, ks); P.fb_mscale = (float) std::ldexp((double) nmu, km);
        P.fb_smask = (1u << ks) - 4u; P.fb_mmask = (1u << km) - 2u; P.fb_sshift = (unsigned) ks; P.fb_mshift = (unsigned) km;
        P.fb_smul = 1u << (32 - ks); P.fb_mmul = 1u << (32 - km);
        P.fb_bias = (0x4B000000u >> ks) + ((bintype == BIN_SMU) ? (0x4B000000u >> km) * (unsigned int) ns : 0u);
        P.df_d2lim = (float) d2lim; P.df_s1sq = (float) s1sq;
        for (int d = 0; d < 3; d++) { P.gorg[d] = (T) g.origin[d]; P.gcs[d] = (T) g.cs[d]; }
        P.tabs_global = 1;
      }
    }
  }
  // double precision otherwise: the float pre-filter (count_kernel_pf.cuh) where it was measured faster than the plain
  // double kernel (weighted counts, survey (s,mu) and (s_perp,pi)), whenever its padded limits stay tight and its
  // shared-memory plan fits next to a shared-memory histogram; the plain double kernel in every other case
  bool use_pf = false;
  PfPlan ppl{};
  if (!is_float && !use_df && !opt.no_prefilter && v.smem_hist && (withwt || (!b->periodic && bintype != BIN_ISO) || opt.force_prefilter)) {
    double lim[4], M = 0;
    for (int d = 0; d < 3; d++) M = std::max(M, std::max(std::fabs(lo[d]), std::fabs(hi[d])) + (b->periodic ? b->bsize[d] : 0.0));
    const int mode = fcfc_gpu_prefilter_limits(b->periodic, bintype, s2max, pmax, M, std::max(c1->smax, c2->smax), std::min(c1->smin, c2->smin), lim);
    if (mode) {
      // this kernel has no stacks: what shared memory is left goes to copies of the weighted histogram (lane l adds to copy
      // l mod copies), which thins out the collisions of the CAS-based FP64 atomics inside a warp
      for (int tg = tables_unused ? 1 : 0; tg < 2 && !use_pf; tg++) {
        for (int hc = (withwt && !opt.no_hist_copies) ? 8 : 1; hc >= 1 && !use_pf; hc >>= 1) {
          const int copies = std::max(hc, hist_copies);
          ppl = withwt ? make_pf_plan<true>((int) ntot, (int) sz_stab, (int) sz_ptab, nmutab, ns, np, (int) rows.size(), true, tg != 0, copies, pf_warps(bintype, b->periodic != 0))
                       : make_pf_plan<false>((int) ntot, (int) sz_stab, (int) sz_ptab, nmutab, ns, np, (int) rows.size(), true, tg != 0, 1, pf_warps(bintype, b->periodic != 0));
          if (ppl.total + 1024 <= smem_max) {
            use_pf = true;
            P.tabs_global = tg;
            P.hist_copies = withwt ? copies : 1;
            v.generic = generic || opt.force_generic || (tg && !tables_unused);
          }
        }
      }
      if (use_pf) { P.pf_d2lim = (float) lim[0]; P.pf_plim = (float) lim[1]; P.pf_s2lim = (float) lim[2]; }
    }
  }
  // single precision, the same family of counts: classified staging (count_kernel_cl.cuh) -- every staged secondary point is
  // tested against the tile's bounding box, dropped, binned in place or sent through the ordinary pair loop
  bool use_cl = false;
  ClPlan cpl{};
  if constexpr (is_float) {
    // (box (s,mu) / isotropic: the counts that also qualify for the dense cells of count_kernel; survey isotropic: same kernel
    // without image shifts)
    const bool cl_ok = !withwt && bintype != BIN_SPI && (b->periodic || bintype == BIN_ISO) && !v.generic && v.smem_hist &&
                       P.stab_is_sqrt && (bintype == BIN_ISO || P.mu_is_sqrt) && !opt.no_dense;
    if (cl_ok && !opt.no_classify) {
      int cdepth = 0;
      for (int d = std::min(qdepth_max, 64) & ~3; d >= 12 && !cdepth; d -= 4) {
        cpl = make_cl_plan((int) ntot, ns, (int) rows.size(), qwords, d);
        if (cpl.total + 1024 <= smem_max) cdepth = d;
      }
      if (cdepth) {
        use_cl = true;
        P.qdepth = cdepth;
        P.qkeep = (cdepth >= 32) ? cdepth / 8 : cdepth / 4;
        if (opt.qkeep >= 0) P.qkeep = std::max(0, std::min(opt.qkeep, cdepth / 2));
        P.qkeep = std::max(0, std::min(P.qkeep, cdepth - 1 - 8));
        // (margins: fcfc_gpu_classify_limits)
        float cl[2];
        fcfc_gpu_classify_limits(s2max, maxabs, cl);
        P.cl_skip = cl[0]; P.cl_dense = cl[1];
        P.tabs_global = 1;
      }
    }
  }
  g_stats.prefilter = use_df ? 2 : (use_pf ? 1 : 0);
  cudaEventRecord(evs[1]);
  const int my_items = (S1.nitem * nsplit - part + nparts - 1) / nparts;
  const int warps_blk = use_df ? kDfWarps : (use_pf ? pf_warps(bintype, b->periodic != 0) : (use_cl ? kClWarps : BlockShape<T>::kWarps));
  const int nblocks = std::max(1, std::min(g_ctx.sm_count, (my_items + warps_blk - 1) / warps_blk));
  cudaError_t le;
  if constexpr (!is_float) le = use_df ? launch_count_df(v, P, nblocks, dpl.total)
                                   : (use_pf ? launch_count_pf(v, P, nblocks, ppl.total) : launch_count<T>(v, P, nblocks, pl.total));
  else le = use_cl ? launch_count_cl(v, P, nblocks, cpl.total) : launch_count<T>(v, P, nblocks, pl.total);
  g_stats.kernel_launches++;
  cudaEventRecord(evs[2]);
  if (le != cudaSuccess) { set_err("count kernel launch failed: %s", cudaGetErrorString(le)); cudaGetLastError(); return FCFC_GPU_ERR_CF; }
  // ---- results ----
  cudaError_t ce = cudaMemcpy(withwt ? (void *) cnt_d : (void *) cnt_i, dbuf + o_hist, ntot * 8, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) { set_err("count kernel failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return FCFC_GPU_ERR_CF; }
  if (dev_hist) cudaMemcpy(dev_hist, dbuf + o_hist, ntot * 8, cudaMemcpyDeviceToDevice);
  unsigned long long ev = 0;
  cudaMemcpy(&ev, dbuf + o_cnt + 8, 8, cudaMemcpyDeviceToHost);
#ifdef FCFC_PF_STATS
  if (use_pf) {
    unsigned long long dbg[2] = {0, 0};
    cudaMemcpy(dbg, dbuf + o_cnt + 16, 16, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[fcfc_gpu] pre-filter exact pass: %llu lane-steps, %llu useful (%.3f), %.3f per evaluation\n", dbg[0], dbg[1],
            dbg[0] ? (double) dbg[1] / dbg[0] : 0.0, ev ? (double) dbg[0] / ev : 0.0);
  }
#endif
  cudaEventRecord(evs[3]); cudaEventSynchronize(evs[3]);
  float ms_count = 0, ms_total = 0;
  cudaEventElapsedTime(&ms_count, evs[1], evs[2]); cudaEventElapsedTime(&ms_total, evs[0], evs[3]);
  g_stats.pair_evals = ev;
  g_stats.pair_evals_computed = ev;
  g_stats.classified = (use_cl || (use_pf && !b->periodic)) ? 1 : 0;
  if (use_cl || use_pf) { unsigned long long ec = 0; cudaMemcpy(&ec, dbuf + o_cnt + 32, 8, cudaMemcpyDeviceToHost); g_stats.pair_evals_computed = ec; }
  if (!withwt && cnt_i) { unsigned long long t = 0; for (size_t i = 0; i < ntot; i++) t += (unsigned long long) cnt_i[i]; g_stats.pairs_in = t; }
  g_stats.ms_sort = ms_sort; g_stats.ms_count = ms_count; g_stats.ms_total = ms_total;
  for (int d = 0; d < 3; d++) g_stats.ncell[d] = g.nc[d];
  g_stats.nitem = my_items;
  g_stats.dense_rows = dense ? dense_rows : 0;
  return 0;
}

}  // namespace fcfc

// ==========================================================================================
// C ABI
using namespace fcfc;

extern "C" int fcfc_gpu_abi_version(void) { return FCFC_GPU_ABI_VERSION; }
extern "C" const char *fcfc_gpu_last_error(void) { return g_err.c_str(); }

static void nccl_reset();

extern "C" int fcfc_gpu_set_option(const char *name, long value) {
  const int e = set_option(name, value);
  if (e) set_err("unknown option '%s'", name ? name : "(null)");
  return e;
}

extern "C" int fcfc_gpu_init(int ndev, const int *devices, int verbose) {
  g_verbose = verbose;
  static std::once_flag env_once;
  std::call_once(env_once, options_from_env);      // FCFC_GPU_TUNE: read once per process, never on the counting path
  nccl_reset();          // communicators belong to a device set
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_err("no CUDA device available (%s); fcfc_b200 has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return FCFC_GPU_ERR_CUDA;
  }
  g_ctx.devices.clear();
  if (ndev <= 0) ndev = devices ? 0 : count;
  for (int i = 0; i < ndev; i++) {
    int d = devices ? devices[i] : i;
    if (d < 0 || d >= count) { set_err("invalid device %d", d); return FCFC_GPU_ERR_ARG; }
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, d), FCFC_GPU_ERR_CUDA);
    if (p.major != 10) { set_err("device %d (%s, sm_%d%d) is not a Blackwell sm_100 part", d, p.name, p.major, p.minor); return FCFC_GPU_ERR_CUDA; }
    g_ctx.devices.push_back(d);
    g_ctx.sm_count = p.multiProcessorCount;
    if (verbose) fprintf(stderr, "[fcfc_gpu] device %d: %s, %d SMs\n", d, p.name, p.multiProcessorCount);
  }
  if (g_ctx.devices.empty()) { set_err("no device selected"); return FCFC_GPU_ERR_ARG; }
  // direct NVLink paths between the devices in use (catalogue replicas are copied device to device); where peer
  // access is not available the copies are staged by the driver
  for (int a : g_ctx.devices)
    for (int b2 : g_ctx.devices) {
      int can = 0;
      if (a != b2 && cudaDeviceCanAccessPeer(&can, a, b2) == cudaSuccess && can) {
        cudaSetDevice(a);
        if (cudaDeviceEnablePeerAccess(b2, 0) != cudaSuccess) cudaGetLastError();     // (already enabled: fine)
      }
    }
  CUDA_TRY(cudaSetDevice(g_ctx.devices[0]), FCFC_GPU_ERR_CUDA);
  g_ctx.ready = true;
  return (int) g_ctx.devices.size();
}

static void stage_release_all();       // pinned staging block of the streamed ingest (below)
extern "C" void fcfc_gpu_finalize(void) {
  for (int d : g_ctx.devices) { cudaSetDevice(d); cudaDeviceSynchronize(); }
  nccl_reset();
  stage_release_all();
  pool_release_all(); g_ctx.ready = false; g_ctx.devices.clear();
}

// The second half of an upload, on columns that are already in device memory (c->x, y, z[, s][, w], n rows): rescale in
// `real`, the survey's 4th coordinate when the metric needs it, bounding box, extreme |x|^2, sum of weights, the check
// for non-finite coordinates -- one pass of prep_kernel (build_tree.c:121-140, 2pt/build_tree.c:35-133).  `s_given`: the
// caller supplied x^2+y^2+z^2 (it is kept as it is).
template <class T>
static int catalog_prepare(DevCat *c, bool s_given, double rescale, int sumsq) {
  const size_t n = c->n;
  PoolScope pool;
  unsigned long long *stats = nullptr; double *wsum = nullptr;
  CUDA_TRY(pool.alloc(&stats, 10 * 8), FCFC_GPU_ERR_MEMORY);
  wsum = reinterpret_cast<double *>(stats + 8);
  unsigned long long init[10];
  for (int k = 0; k < 3; k++) { init[k] = ~0ull; init[3 + k] = 0; }
  init[6] = 0; init[7] = 0; init[8] = 0; init[9] = ~0ull;
  CUDA_TRY(cudaMemcpy(stats, init, sizeof init, cudaMemcpyHostToDevice), FCFC_GPU_ERR_CUDA);
  if (n) {
    const int nb = (int) std::min<size_t>((n + 255) / 256, 148 * 8);
    prep_kernel<T><<<nb, 256>>>((T *) c->x, (T *) c->y, (T *) c->z, (T *) c->s, n, (T) rescale, rescale != 1.0,
                                s_given ? -1 : sumsq, stats, wsum, (const T *) c->w);
    g_stats.kernel_launches++;
  }
  unsigned long long out[10];
  CUDA_TRY(cudaMemcpy(out, stats, sizeof out, cudaMemcpyDeviceToHost), FCFC_GPU_ERR_CUDA);
  if (out[7]) { set_err("catalogue contains %llu non-finite coordinates", out[7]); return FCFC_GPU_ERR_DATA; }
  for (int k = 0; k < 3; k++) { c->bmin[k] = n ? dec_f64(out[k]) : 0; c->bmax[k] = n ? dec_f64(out[3 + k]) : 0; }
  c->smax = n ? dec_f64(out[6]) : 0;
  c->smin = n ? dec_f64(out[9]) : 0;
  double ws; memcpy(&ws, &out[8], 8);
  c->wsum = c->has_w ? ws : (double) n;
  return 0;
}

// The catalogue columns arrive as host pointers (the FCFC host's malloc'd reader arrays) or as device pointers (a
// launcher that all-gathered slices over NVLink): cudaMemcpyDefault resolves either through unified addressing.
template <class T>
static int catalog_upload(DevCat *c, const void *x, const void *y, const void *z, const void *s, const void *w,
                          double rescale, int sumsq) {
  const size_t n = c->n, bytes = (n ? n : 1) * sizeof(T);
  CUDA_TRY(pool_alloc(&c->x, bytes), FCFC_GPU_ERR_MEMORY);
  CUDA_TRY(pool_alloc(&c->y, bytes), FCFC_GPU_ERR_MEMORY);
  CUDA_TRY(pool_alloc(&c->z, bytes), FCFC_GPU_ERR_MEMORY);
  CUDA_TRY(cudaMemcpy(c->x, x, n * sizeof(T), cudaMemcpyDefault), FCFC_GPU_ERR_CUDA);
  CUDA_TRY(cudaMemcpy(c->y, y, n * sizeof(T), cudaMemcpyDefault), FCFC_GPU_ERR_CUDA);
  CUDA_TRY(cudaMemcpy(c->z, z, n * sizeof(T), cudaMemcpyDefault), FCFC_GPU_ERR_CUDA);
  c->has_s = (s != nullptr) || sumsq >= 0;
  if (c->has_s) {
    CUDA_TRY(pool_alloc(&c->s, bytes), FCFC_GPU_ERR_MEMORY);
    if (s) CUDA_TRY(cudaMemcpy(c->s, s, n * sizeof(T), cudaMemcpyDefault), FCFC_GPU_ERR_CUDA);
  }
  c->has_w = w != nullptr;
  if (w) {
    CUDA_TRY(pool_alloc(&c->w, bytes), FCFC_GPU_ERR_MEMORY);
    CUDA_TRY(cudaMemcpy(c->w, w, n * sizeof(T), cudaMemcpyDefault), FCFC_GPU_ERR_CUDA);
  }
  return catalog_prepare<T>(c, s != nullptr, rescale, sumsq);
}

// Replica of an uploaded (rescaled, checked) catalogue on another device: device-to-device copies over NVLink /
// NVSwitch instead of one more pass over the host arrays per GPU -- the replacement of kdtree_broadcast
// (src/tree/kdtree.c:529-619).  All copies of one replica are asynchronous on that device's default stream, so the
// replicas of a catalogue fill concurrently.
static int catalog_replicate(DevCat *dst, const DevCat *src, size_t real_bytes) {
  const size_t bytes = (src->n ? src->n : 1) * real_bytes, used = src->n * real_bytes;
  void *const from[5] = {src->x, src->y, src->z, src->s, src->w};
  void **const to[5] = {&dst->x, &dst->y, &dst->z, &dst->s, &dst->w};
  for (int k = 0; k < 5; k++) {
    if (!from[k]) continue;
    CUDA_TRY(pool_alloc(to[k], bytes), FCFC_GPU_ERR_MEMORY);
    if (used) CUDA_TRY(cudaMemcpyPeerAsync(*to[k], dst->device, from[k], src->device, used, 0), FCFC_GPU_ERR_CUDA);
  }
  dst->has_s = src->has_s; dst->has_w = src->has_w;
  memcpy(dst->bmin, src->bmin, sizeof dst->bmin); memcpy(dst->bmax, src->bmax, sizeof dst->bmax);
  dst->smax = src->smax; dst->smin = src->smin; dst->wsum = src->wsum;
  return 0;
}

// Replicas of a prepared first copy (c->dev[0]) on every other device in use, copied concurrently and complete on return.
static int catalog_spread(fcfc_gpu_catalog *c) {
  for (size_t i = 1; i < g_ctx.devices.size(); i++) {
    DevCat *dc = new DevCat();
    dc->is_float = c->is_float; dc->n = c->n; dc->device = g_ctx.devices[i];
    c->dev.push_back(dc);
    cudaSetDevice(dc->device);
    const int e = catalog_replicate(dc, c->dev[0], c->is_float ? sizeof(float) : sizeof(double));
    if (e) return e;
  }
  for (size_t i = 1; i < g_ctx.devices.size(); i++) {       // the replicas are complete before anyone counts on them
    cudaSetDevice(g_ctx.devices[i]);
    if (cudaStreamSynchronize(0) != cudaSuccess) { set_err("catalogue replication failed: %s", cudaGetErrorString(cudaGetLastError())); return FCFC_GPU_ERR_CUDA; }
  }
  return 0;
}

extern "C" fcfc_gpu_catalog *fcfc_gpu_catalog_create(const void *x, const void *y, const void *z, const void *x2sum,
                                                     const void *w, size_t n, int is_float, double rescale, int sumsq_arith) {
  if (ensure_init()) return nullptr;
  if (n && (!x || !y || !z)) { set_err("NULL coordinate array"); return nullptr; }
  if (n >= (1ull << 31) - 64) { set_err("catalogue too large for 32-bit point indices"); return nullptr; }
  fcfc_gpu_catalog *c = new fcfc_gpu_catalog();
  c->is_float = is_float; c->n = n; c->has_w = w != nullptr;
  // one replica per device: the first is uploaded from the caller's arrays and prepared (rescale, sums, bounding box,
  // checks) once; the others are copied from it device to device, concurrently
  DevCat *dc = new DevCat();
  dc->is_float = is_float; dc->n = n; dc->device = g_ctx.devices[0];
  c->dev.push_back(dc);
  cudaSetDevice(dc->device);
  int e = is_float ? catalog_upload<float>(dc, x, y, z, x2sum, w, rescale, sumsq_arith)
                   : catalog_upload<double>(dc, x, y, z, x2sum, w, rescale, sumsq_arith);
  if (!e) e = catalog_spread(c);
  if (e) { fcfc_gpu_catalog_destroy(c); cudaSetDevice(g_ctx.devices[0]); return nullptr; }
  c->wsum = c->dev[0]->wsum;
  cudaSetDevice(g_ctx.devices[0]);
  return c;
}

// ------------------------------------------------------------------------------------------
// Streamed ingest.  The reference's readers fill the catalogue columns chunk by chunk
// (src/io/read_ascii.c:750-950: fread a chunk, parse its lines, append to res[i]) and only then does tree_create rescale
// and index them (2pt_box/build_tree.c:84-156).  With this interface the reader hands every parsed chunk over as soon as it
// exists: the rows are copied into one of a few pinned staging slots and sent to the device with cudaMemcpyAsync on a
// private stream, so the transfer of chunk k overlaps the parsing of chunk k + 1 and the catalogue is resident when the
// file ends.  finish() then runs the same prep_kernel pass as the one-shot upload and spreads the replicas.
struct fcfc_gpu_catalog_stream {
  int is_float = 0, has_w = 0, device = 0;
  size_t rb = 8;                  // bytes per value
  size_t n = 0, cap = 0;          // rows appended / rows the device columns hold
  void *col[4] = {nullptr, nullptr, nullptr, nullptr};      // x, y, z, w on the device (pool blocks)
  int ncol = 3;
  cudaStream_t stream = nullptr;
  static constexpr int kSlots = 4;
  static constexpr size_t kSlotRows = (size_t) 1 << 16;     // 64 Ki rows per slot and column: transfers of 256-512 KB
  unsigned char *pinned = nullptr;
  size_t pinned_bytes = 0;
  cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  bool busy[kSlots] = {false, false, false, false};
  int next = 0;
  bool failed = false;
};

// The pinned staging block of the last finished stream is kept for the next one (cudaMallocHost costs ~1 ms per MB: a
// program reads two or three catalogues in a row); fcfc_gpu_finalize gives it back.
static std::mutex g_stage_mutex;
static unsigned char *g_stage_cached = nullptr;
static size_t g_stage_cached_bytes = 0;
static unsigned char *stage_take(size_t bytes, size_t *got) {
  {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    if (g_stage_cached && g_stage_cached_bytes >= bytes) {
      unsigned char *p = g_stage_cached; *got = g_stage_cached_bytes;
      g_stage_cached = nullptr; g_stage_cached_bytes = 0;
      return p;
    }
  }
  unsigned char *p = nullptr;
  if (cudaMallocHost(reinterpret_cast<void **>(&p), bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  *got = bytes;
  return p;
}
static void stage_give(unsigned char *p, size_t bytes) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    if (!g_stage_cached || g_stage_cached_bytes < bytes) { std::swap(p, g_stage_cached); std::swap(bytes, g_stage_cached_bytes); }
  }
  if (p) cudaFreeHost(p);
}
static void stage_release_all() {
  std::lock_guard<std::mutex> lock(g_stage_mutex);
  if (g_stage_cached) cudaFreeHost(g_stage_cached);
  g_stage_cached = nullptr; g_stage_cached_bytes = 0;
}

static void stream_release(fcfc_gpu_catalog_stream *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (auto &e : b->done) if (e) cudaEventDestroy(e);
  if (b->stream) cudaStreamDestroy(b->stream);
  stage_give(b->pinned, b->pinned_bytes);
  for (auto &p : b->col) pool_free(p);
  cudaGetLastError();
  delete b;
}

// device columns for at least `rows` rows; rows already appended are carried over (device to device, stream order)
static int stream_reserve(fcfc_gpu_catalog_stream *b, size_t rows) {
  if (rows <= b->cap) return 0;
  size_t cap = std::max<size_t>(rows, std::max<size_t>(2 * b->cap, 1024));
  void *fresh[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < b->ncol; k++) {
    cudaError_t e = pool_alloc(&fresh[k], cap * b->rb);
    if (e != cudaSuccess) {
      for (int j = 0; j < k; j++) pool_free(fresh[j]);
      set_err("out of device memory while growing a streamed catalogue to %zu rows: %s", cap, cudaGetErrorString(e));
      cudaGetLastError();
      return FCFC_GPU_ERR_MEMORY;
    }
    if (b->n && e == cudaSuccess) e = cudaMemcpyAsync(fresh[k], b->col[k], b->n * b->rb, cudaMemcpyDeviceToDevice, b->stream);
    if (e != cudaSuccess) {
      for (int j = 0; j <= k; j++) pool_free(fresh[j]);
      set_err("streamed ingest: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return FCFC_GPU_ERR_CUDA;
    }
  }
  if (b->cap && cudaStreamSynchronize(b->stream) != cudaSuccess) {      // the old blocks go back to the pool: nothing may still read them
    for (int k = 0; k < b->ncol; k++) pool_free(fresh[k]);
    set_err("streamed ingest: %s", cudaGetErrorString(cudaGetLastError()));
    return FCFC_GPU_ERR_CUDA;
  }
  for (int k = 0; k < b->ncol; k++) { pool_free(b->col[k]); b->col[k] = fresh[k]; }
  b->cap = cap;
  return 0;
}

extern "C" fcfc_gpu_catalog_stream *fcfc_gpu_catalog_stream_begin(size_t n_hint, int is_float, int with_weight) {
  if (ensure_init()) return nullptr;
  fcfc_gpu_catalog_stream *b = new fcfc_gpu_catalog_stream();
  b->is_float = is_float != 0; b->has_w = with_weight != 0; b->rb = is_float ? sizeof(float) : sizeof(double);
  b->ncol = with_weight ? 4 : 3;
  b->device = g_ctx.devices[0];
  cudaSetDevice(b->device);
  const size_t slot_bytes = fcfc_gpu_catalog_stream::kSlotRows * b->rb * (size_t) b->ncol;
  cudaError_t e = cudaStreamCreate(&b->stream);      // (blocking flavour: ordered after whatever the default stream still does with recycled pool blocks)
  if (e == cudaSuccess && !(b->pinned = stage_take(slot_bytes * fcfc_gpu_catalog_stream::kSlots, &b->pinned_bytes))) e = cudaErrorMemoryAllocation;
  for (int k = 0; e == cudaSuccess && k < fcfc_gpu_catalog_stream::kSlots; k++) e = cudaEventCreateWithFlags(&b->done[k], cudaEventDisableTiming);
  if (e != cudaSuccess) { set_err("streamed ingest: %s", cudaGetErrorString(e)); cudaGetLastError(); stream_release(b); return nullptr; }
  n_hint = std::min<size_t>(n_hint, ((size_t) 1 << 31) - 65);       // (a hint, not a promise: append() enforces the limit)
  if (n_hint && stream_reserve(b, n_hint)) { stream_release(b); return nullptr; }
  return b;
}

extern "C" int fcfc_gpu_catalog_stream_append(fcfc_gpu_catalog_stream *b, const void *x, const void *y, const void *z,
                                              const void *w, size_t n) {
  if (!b) { set_err("NULL catalogue stream"); return FCFC_GPU_ERR_ARG; }
  if (b->failed) { set_err("catalogue stream is in a failed state"); return FCFC_GPU_ERR_ARG; }
  if (n == 0) return 0;
  if (!x || !y || !z || (b->has_w && !w)) { set_err("NULL column in a chunk of %zu rows", n); return FCFC_GPU_ERR_ARG; }
  if (b->n + n >= (1ull << 31) - 64) { set_err("catalogue too large for 32-bit point indices"); return FCFC_GPU_ERR_ARG; }
  cudaSetDevice(b->device);
  int e = stream_reserve(b, b->n + n);
  if (e) { b->failed = true; return e; }
  const void *src[4] = {x, y, z, w};
  const size_t slot_rows = fcfc_gpu_catalog_stream::kSlotRows, slot_bytes = slot_rows * b->rb * (size_t) b->ncol;
  for (size_t off = 0; off < n; off += slot_rows) {
    const size_t m = std::min(slot_rows, n - off);
    const int slot = b->next;
    b->next = (b->next + 1) % fcfc_gpu_catalog_stream::kSlots;
    if (b->busy[slot]) {          // the transfer that last used this slot must have left it
      cudaError_t ce = cudaEventSynchronize(b->done[slot]);
      if (ce != cudaSuccess) { b->failed = true; set_err("streamed ingest: %s", cudaGetErrorString(ce)); cudaGetLastError(); return FCFC_GPU_ERR_CUDA; }
    }
    unsigned char *stage = b->pinned + (size_t) slot * slot_bytes;
    for (int k = 0; k < b->ncol; k++) {
      unsigned char *h = stage + (size_t) k * slot_rows * b->rb;
      memcpy(h, static_cast<const unsigned char *>(src[k]) + off * b->rb, m * b->rb);
      cudaError_t ce = cudaMemcpyAsync(static_cast<unsigned char *>(b->col[k]) + (b->n + off) * b->rb, h, m * b->rb, cudaMemcpyHostToDevice, b->stream);
      if (ce != cudaSuccess) { b->failed = true; set_err("streamed ingest: %s", cudaGetErrorString(ce)); cudaGetLastError(); return FCFC_GPU_ERR_CUDA; }
    }
    cudaEventRecord(b->done[slot], b->stream);
    b->busy[slot] = true;
  }
  b->n += n;
  return 0;
}

extern "C" size_t fcfc_gpu_catalog_stream_size(const fcfc_gpu_catalog_stream *b) { return b ? b->n : 0; }

extern "C" void fcfc_gpu_catalog_stream_abort(fcfc_gpu_catalog_stream *b) {
  stream_release(b);
  if (g_ctx.ready && !g_ctx.devices.empty()) cudaSetDevice(g_ctx.devices[0]);
}

extern "C" fcfc_gpu_catalog *fcfc_gpu_catalog_stream_finish(fcfc_gpu_catalog_stream *b, double rescale, int sumsq_arith) {
  if (!b) { set_err("NULL catalogue stream"); return nullptr; }
  if (b->failed) { set_err("catalogue stream is in a failed state"); stream_release(b); return nullptr; }
  cudaSetDevice(b->device);
  if (stream_reserve(b, 1)) { stream_release(b); return nullptr; }      // (an empty catalogue still owns valid columns)
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) {
    set_err("streamed ingest: %s", cudaGetErrorString(cudaGetLastError()));
    stream_release(b);
    return nullptr;
  }
  fcfc_gpu_catalog *c = new fcfc_gpu_catalog();
  c->is_float = b->is_float; c->n = b->n; c->has_w = b->has_w != 0;
  DevCat *dc = new DevCat();
  dc->is_float = b->is_float; dc->n = b->n; dc->device = b->device;
  dc->x = b->col[0]; dc->y = b->col[1]; dc->z = b->col[2]; dc->w = b->col[3];     // ownership of the columns passes to the catalogue
  for (auto &p : b->col) p = nullptr;
  dc->has_w = b->has_w != 0;
  c->dev.push_back(dc);
  const int is_float = b->is_float;
  stream_release(b);
  cudaSetDevice(dc->device);
  int e = 0;
  dc->has_s = sumsq_arith >= 0;
  if (dc->has_s) {
    cudaError_t ce = pool_alloc(&dc->s, (dc->n ? dc->n : 1) * (is_float ? sizeof(float) : sizeof(double)));
    if (ce != cudaSuccess) { set_err("out of device memory: %s", cudaGetErrorString(ce)); cudaGetLastError(); e = FCFC_GPU_ERR_MEMORY; }
  }
  if (!e) e = is_float ? catalog_prepare<float>(dc, false, rescale, sumsq_arith) : catalog_prepare<double>(dc, false, rescale, sumsq_arith);
  if (!e) e = catalog_spread(c);
  if (e) { fcfc_gpu_catalog_destroy(c); cudaSetDevice(g_ctx.devices[0]); return nullptr; }
  c->wsum = c->dev[0]->wsum;
  cudaSetDevice(g_ctx.devices[0]);
  return c;
}

extern "C" void fcfc_gpu_catalog_destroy(fcfc_gpu_catalog *c) {
  if (!c) return;
  for (DevCat *dc : c->dev) {
    cudaSetDevice(dc->device);
    pool_free(dc->x); pool_free(dc->y); pool_free(dc->z); pool_free(dc->s); pool_free(dc->w);
    for (auto &S : dc->sf) S.release();
    for (auto &S : dc->sd) S.release();
    delete dc;
  }
  if (g_ctx.ready && !g_ctx.devices.empty()) cudaSetDevice(g_ctx.devices[0]);
  delete c;
}
extern "C" size_t fcfc_gpu_catalog_size(const fcfc_gpu_catalog *c) { return c ? c->n : 0; }
extern "C" double fcfc_gpu_catalog_wsum(const fcfc_gpu_catalog *c) { return c ? c->wsum : 0; }

static int check_count_args(fcfc_gpu_catalog *c1, fcfc_gpu_catalog *c2, const fcfc_gpu_bins *b, int isauto, int withwt,
                            const int64_t *cnt_i, const double *cnt_d) {
  if (ensure_init()) return FCFC_GPU_ERR_CUDA;
  if (!c1 || !c2 || !b) { set_err("NULL argument"); return FCFC_GPU_ERR_ARG; }
  if ((withwt && !cnt_d) || (!withwt && !cnt_i)) { set_err("missing output array"); return FCFC_GPU_ERR_ARG; }
  if (c1->is_float != c2->is_float || c1->is_float != (b->is_float != 0)) { set_err("precision mismatch between catalogues and bins"); return FCFC_GPU_ERR_ARG; }
  if (isauto && c1 != c2) { set_err("auto count needs cat1 == cat2"); return FCFC_GPU_ERR_ARG; }
  if (c1->dev.empty() || c1->dev.size() != c2->dev.size()) { set_err("catalogues are resident on different device sets"); return FCFC_GPU_ERR_ARG; }
  return 0;
}

extern "C" int fcfc_gpu_count_partial(fcfc_gpu_catalog *c1, fcfc_gpu_catalog *c2, const fcfc_gpu_bins *b, int isauto,
                                      int withwt, int part, int nparts, int64_t *cnt_i, double *cnt_d, void *dev_hist) {
  int e = check_count_args(c1, c2, b, isauto, withwt, cnt_i, cnt_d);
  if (e) return e;
  if (nparts < 1 || part < 0 || part >= nparts) { set_err("invalid shard %d/%d", part, nparts); return FCFC_GPU_ERR_ARG; }
  if (!isauto && c2->n > c1->n) std::swap(c1, c2);      // cross counts are symmetric: the larger catalogue supplies the tiles
  cudaSetDevice(c1->dev[0]->device);
  return b->is_float ? count_impl<float>(c1->dev[0], c2->dev[0], b, isauto, withwt, part, nparts, cnt_i, cnt_d, dev_hist)
                     : count_impl<double>(c1->dev[0], c2->dev[0], b, isauto, withwt, part, nparts, cnt_i, cnt_d, dev_hist);
}

// ---- NCCL, bound at run time (only needed when one process drives several devices) -------------------
namespace {
struct Nccl {
  void *lib = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::vector<void *> comms;
  std::vector<int> devices;           // the device set the communicators were made for
  bool ready = false, tried = false;
};
Nccl g_nccl;
bool nccl_setup() {
  if (g_nccl.tried) return g_nccl.ready;
  g_nccl.tried = true;
  g_nccl.devices = g_ctx.devices;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) { g_nccl.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return false;
#define FCFC_SYM(f, n) g_nccl.f = reinterpret_cast<decltype(g_nccl.f)>(dlsym(g_nccl.lib, n)); if (!g_nccl.f) return false;
  FCFC_SYM(CommInitAll, "ncclCommInitAll") FCFC_SYM(AllReduce, "ncclAllReduce") FCFC_SYM(GroupStart, "ncclGroupStart")
  FCFC_SYM(GroupEnd, "ncclGroupEnd") FCFC_SYM(CommDestroy, "ncclCommDestroy") FCFC_SYM(GetErrorString, "ncclGetErrorString")
#undef FCFC_SYM
  g_nccl.comms.assign(g_nccl.devices.size(), nullptr);
  if (g_nccl.CommInitAll(g_nccl.comms.data(), (int) g_nccl.devices.size(), g_nccl.devices.data()) != 0) return false;
  g_nccl.ready = true;
  return true;
}
}  // namespace

static void nccl_reset() {
  if (g_nccl.ready) for (void *c : g_nccl.comms) if (c) g_nccl.CommDestroy(c);
  g_nccl.comms.clear(); g_nccl.ready = false; g_nccl.tried = false;
}

// count_pairs() over every device in use: device d counts shard d of the primary catalogue's work items
// against its replica of the secondary catalogue (one host thread per device), then the per-device
// histograms are summed with one ncclAllReduce (int64 / float64, ntot elements) and read back once.
extern "C" int fcfc_gpu_count(fcfc_gpu_catalog *c1, fcfc_gpu_catalog *c2, const fcfc_gpu_bins *b, int isauto, int withwt,
                              int64_t *cnt_i, double *cnt_d) {
  int e = check_count_args(c1, c2, b, isauto, withwt, cnt_i, cnt_d);
  if (e) return e;
  const int ndev = (int) c1->dev.size();
  if (ndev == 1) return fcfc_gpu_count_partial(c1, c2, b, isauto, withwt, 0, 1, cnt_i, cnt_d, nullptr);
  if (!isauto && c2->n > c1->n) std::swap(c1, c2);      // cross counts are symmetric: the larger catalogue supplies the tiles
  const int bt = b->bintype;
  const size_t ntot = (size_t) b->ns * (bt == FCFC_GPU_BIN_ISO ? 1 : (bt == FCFC_GPU_BIN_SMU ? b->nmu : b->np));
  std::vector<void *> dhist(ndev, nullptr);
  std::vector<std::vector<unsigned char>> part(ndev, std::vector<unsigned char>(ntot * 8));
  std::vector<int> rc(ndev, 0);
  std::vector<std::string> errs(ndev);
  std::vector<fcfc_gpu_stats> st(ndev);
  for (int d = 0; d < ndev; d++) {
    cudaSetDevice(c1->dev[d]->device);
    if (pool_alloc(&dhist[d], ntot * 8) != cudaSuccess) {
      for (int k = 0; k < d; k++) { cudaSetDevice(c1->dev[k]->device); pool_free(dhist[k]); }
      cudaSetDevice(c1->dev[0]->device);
      set_err("out of device memory"); return FCFC_GPU_ERR_MEMORY;
    }
  }
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; d++)
    th.emplace_back([&, d]() {
      cudaSetDevice(c1->dev[d]->device);
      int64_t *pi = withwt ? nullptr : reinterpret_cast<int64_t *>(part[d].data());
      double *pd = withwt ? reinterpret_cast<double *>(part[d].data()) : nullptr;
      rc[d] = b->is_float ? count_impl<float>(c1->dev[d], c2->dev[d], b, isauto, withwt, d, ndev, pi, pd, dhist[d])
                          : count_impl<double>(c1->dev[d], c2->dev[d], b, isauto, withwt, d, ndev, pi, pd, dhist[d]);
      if (rc[d]) errs[d] = g_err;
      st[d] = g_stats;
    });
  for (auto &t : th) t.join();
  int bad = 0;
  for (int d = 0; d < ndev; d++) if (rc[d]) { bad = rc[d]; g_err = errs[d]; }
  if (!bad) {
    // One process driving several devices: the ntot-element partial histograms are already on the host (count_impl reads
    // them back), and adding them there costs microseconds, whereas ncclCommInitAll costs 1-5 s per process (measured:
    // a 2-GPU DD count of the 10^7 box took 1.6 s with it, 0.2 s without).  The all-reduce over NVLink is therefore opt-in
    // here (option nccl = 1); launchers with one process per GPU (bench.py under torchrun) reduce with NCCL as a matter of
    // course, their communicators exist anyway.
    if (options_snapshot().nccl && nccl_setup()) {
      int ne = g_nccl.GroupStart();
      for (int d = 0; d < ndev && !ne; d++) {
        cudaSetDevice(c1->dev[d]->device);
        ne = g_nccl.AllReduce(dhist[d], dhist[d], ntot, withwt ? 8 /* ncclFloat64 */ : 4 /* ncclInt64 */, 0 /* ncclSum */, g_nccl.comms[d], 0);
      }
      if (!ne) ne = g_nccl.GroupEnd(); else g_nccl.GroupEnd();
      cudaSetDevice(c1->dev[0]->device);
      if (ne || cudaMemcpy(withwt ? (void *) cnt_d : (void *) cnt_i, dhist[0], ntot * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_err("NCCL all-reduce of the histograms failed: %s", ne ? g_nccl.GetErrorString(ne) : "copy back"); bad = FCFC_GPU_ERR_CUDA;
      }
    } else {
      // add the ntot-element partial histograms on the host (integer sums are exact either way)
      for (size_t k = 0; k < ntot; k++) {
        if (withwt) { double v = 0; for (int d = 0; d < ndev; d++) v += reinterpret_cast<double *>(part[d].data())[k]; cnt_d[k] = v; }
        else { int64_t v = 0; for (int d = 0; d < ndev; d++) v += reinterpret_cast<int64_t *>(part[d].data())[k]; cnt_i[k] = v; }
      }
    }
  }
  for (int d = 0; d < ndev; d++) { cudaSetDevice(c1->dev[d]->device); pool_free(dhist[d]); }
  cudaSetDevice(c1->dev[0]->device);
  // statistics: sums over the devices, times are the slowest device
  g_stats = st[0];
  for (int d = 1; d < ndev; d++) {
    g_stats.pair_evals += st[d].pair_evals; g_stats.kernel_launches += st[d].kernel_launches; g_stats.nitem += st[d].nitem;
    g_stats.ms_sort = std::max(g_stats.ms_sort, st[d].ms_sort); g_stats.ms_count = std::max(g_stats.ms_count, st[d].ms_count);
    g_stats.ms_total = std::max(g_stats.ms_total, st[d].ms_total);
  }
  if (!bad && !withwt) { unsigned long long t = 0; for (size_t k = 0; k < ntot; k++) t += (unsigned long long) cnt_i[k]; g_stats.pairs_in = t; }
  return bad;
}

extern "C" int fcfc_gpu_get_stats(fcfc_gpu_stats *out) {
  if (!out) return FCFC_GPU_ERR_ARG;
  *out = g_stats;
  return 0;
}
