// fcfc_b200/csrc/host_setup.cpp -- host-side mirror of the reference's bin setup, for callers that
// use libfcfc_b200.so without the FCFC host (bench.py, the Python API, third-party codes).
//
// When the engine is linked into FCFC itself none of this runs: cf_setup() has already produced
// cf->rescale, cf->s2bin, cf->pbin/p2bin, cf->stab, cf->ptab, cf->mutab and the shim passes them
// through.  The functions below rebuild the same objects from the same inputs:
//   least_fac2 / least_fac4      src/fcfc/2pt_box/setup_cf.c:308-376
//   create_tab_sbin              src/fcfc/2pt_box/setup_cf.c:385-430
//   create_tab_sp_pi             src/fcfc/2pt_box/setup_cf.c:439-503, src/fcfc/2pt/setup_cf.c:432-503
//   create_tab_mu                src/fcfc/2pt_box/setup_cf.c:513-531
//   create_lut_int / _hybrid     src/util/create_lut.c:58-140
// tests/test_bins_setup.py checks them against dumps of the compiled reference.
#include "../../include/fcfc_gpu.h"

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

struct fcfc_gpu_bins_owner {
  fcfc_gpu_bins bins;
  double rescale = 1;
  std::vector<unsigned char> s2bin, pbin, stab, ptab, mutab;
};

namespace {

template <class T> struct Lim;
template <> struct Lim<float> { static constexpr double maxint = 16777216.0; static constexpr double tol = 1e-5; };
template <> struct Lim<double> { static constexpr double maxint = 9007199254740992.0; static constexpr double tol = 1e-10; };

size_t gcd_sz(size_t a, size_t b) { while (b) { size_t t = b; b = a % b; a = t; } return a; }

// Smallest factor that turns all values into non-negative integers (<= 4 decimal digits tried).
template <class T> T least_factor(std::vector<T> v) {
  for (T a : v) if (a < 0 || a > (T) Lim<T>::maxint) return 0;
  if (v.size() == 2) {
    if (v[0] == 0) return 1 / v[1];
    if (v[1] == 0) return 1 / v[0];
  }
  size_t tries = 0, ifac = 1;
  for (;;) {
    bool integral = true;
    for (T a : v) if (!(std::fabs(std::round(a) - a) < (T) Lim<T>::tol)) integral = false;
    if (integral) break;
    for (T &a : v) a *= 10;
    ifac *= 10;
    if (++tries > 4) return 0;
  }
  for (T a : v) if (a > (T) Lim<T>::maxint) return 0;
  size_t g = 0;
  if (v.size() == 2) g = gcd_sz((size_t) std::round(v[0]), (size_t) std::round(v[1]));
  else {
    for (size_t k = 0; k < v.size(); k++) {
      if (v[k] == 0) continue;
      // the reference truncates instead of rounding once a divisor is known (setup_cf.c:366-371)
      g = (g == 0) ? (size_t) std::round(v[k]) : gcd_sz(g, (size_t) v[k]);
    }
    if (g == 0) g = 1;
  }
  return (T) ifac / (T) g;
}

template <class T> bool int_table(const T *edges, int num, std::vector<unsigned char> &out, int &width, uint64_t &len) {
  const long lo = (long) edges[0], hi = (long) edges[num], n = hi - lo;
  if (num > 65535 || n < 0) return false;
  width = (num <= 255) ? 0 : 1;
  out.assign((size_t) (n + 8) * (width ? 2 : 1), 0);
  int bin = 0;
  for (long i = 0; i < n; i++) {
    while (!((T) (i + lo) < edges[bin + 1])) if (++bin >= num) return false;
    if (width) reinterpret_cast<uint16_t *>(out.data())[i] = (uint16_t) bin; else out[i] = (unsigned char) bin;
  }
  len = (uint64_t) n;
  return true;
}

template <class T> bool hybrid_table(const T *edges, int num, std::vector<unsigned char> &out, int &width, uint64_t &len) {
  const long lo = (long) edges[0], hi = (long) std::ceil((double) edges[num]), n = hi - lo;
  if (num > 65535 / 2 || n < 0) return false;
  width = (num <= 255 / 2) ? 0 : 1;
  out.assign((size_t) (n + 8) * (width ? 2 : 1), 0);
  auto put = [&](long i, int v) { if (width) reinterpret_cast<uint16_t *>(out.data())[i] = (uint16_t) v; else out[i] = (unsigned char) v; };
  int k = 1;                    // candidate upper edge index
  long edge = (long) edges[k];
  long i = 0;
  for (; i < n; i++) {
    if (i + lo < edge) { put(i, k - 1); continue; }
    // this integer cell contains at least one edge: store "start from bin k'-1 and walk down"
    bool found = false;
    while (++k <= num) {
      edge = (long) edges[k];
      if (i + lo < edge) { put(i, num + k - 1); found = true; break; }
    }
    if (!found) break;
  }
  for (; i < n; i++) put(i, num * 2);
  len = (uint64_t) n;
  return true;
}

template <class T>
fcfc_gpu_bins_owner *make_bins(int periodic, int bintype, int linear, double smin, double ds, double pmin, double dpi,
                               const double *sedge, int ns, const double *pedge, int np, int nmu, const double box[3],
                               int with_mu_one, int arith) {
  if (ns < 1 || (bintype == FCFC_GPU_BIN_SPI && np < 1) || (bintype == FCFC_GPU_BIN_SMU && (nmu < 1 || nmu > 255))) return nullptr;
  if (!linear && (!sedge || (bintype == FCFC_GPU_BIN_SPI && !pedge))) return nullptr;
  auto *o = new fcfc_gpu_bins_owner();
  std::vector<T> sb(ns + 1), s2(ns + 1), pb, p2;
  for (int i = 0; i <= ns; i++) sb[i] = linear ? (T) (smin + ds * i) : (T) sedge[i];        // setup_cf.c:612
  const bool spi = bintype == FCFC_GPU_BIN_SPI;
  if (spi) { pb.resize(np + 1); p2.resize(np + 1); for (int i = 0; i <= np; i++) pb[i] = linear ? (T) (pmin + dpi * i) : (T) pedge[i]; }
  T fac = 0;
  int tabtype = FCFC_GPU_TAB_HYBRID;
  if (linear) {
    fac = spi ? least_factor<T>({(T) smin, (T) ds, (T) pmin, (T) dpi}) : least_factor<T>({(T) smin, (T) ds});
    if (fac != 0) {
      T a = sb[0] * fac; a = std::round(a * a);
      T b = sb[ns] * fac; b = std::round(b * b);
      bool ok = a <= (T) Lim<T>::maxint && b <= (T) Lim<T>::maxint && (size_t) b - (size_t) a <= 40960;
      if (ok && spi) { T p1 = pb[0] * fac, p2v = pb[np] * fac; ok = (size_t) p2v - (size_t) p1 <= 40960; }
      if (ok) tabtype = FCFC_GPU_TAB_INT;
    }
  }
  if (tabtype == FCFC_GPU_TAB_INT) {
    for (int i = 0; i <= ns; i++) { sb[i] = std::round(sb[i] * fac); s2[i] = sb[i] * sb[i]; }
    if (spi) for (int i = 0; i <= np; i++) { pb[i] = std::round(pb[i] * fac); p2[i] = periodic ? pb[i] : pb[i] * pb[i]; }
  } else {
    T smax = sb[ns]; smax *= smax;
    T larger = smax, smaller = smax;
    if (spi) {
      T pmax = pb[np]; if (!periodic) pmax *= pmax;
      if (smax >= pmax) smaller = pmax; else { larger = pmax; smaller = smax; }
    }
    fac = (T) 32768 / larger;
    fac = (T) std::pow((T) 2, std::logb(fac));          // REAL_TRUNC_FRAC
    while (fac * smaller > (T) (8192 * 2)) fac *= (T) 0.5;
    for (int i = 0; i <= ns; i++) { sb[i] *= fac; s2[i] = sb[i] * sb[i]; }
    if (spi) for (int i = 0; i <= np; i++) { pb[i] *= fac; p2[i] = periodic ? pb[i] : pb[i] * pb[i]; }
  }
  fcfc_gpu_bins &B = o->bins;
  memset(&B, 0, sizeof B);
  int sw = 0, pw = 0; uint64_t nst = 0, npt = 0;
  bool ok = (tabtype == FCFC_GPU_TAB_INT) ? int_table<T>(s2.data(), ns, o->stab, sw, nst) : hybrid_table<T>(s2.data(), ns, o->stab, sw, nst);
  if (ok && spi) ok = (tabtype == FCFC_GPU_TAB_INT) ? int_table<T>(p2.data(), np, o->ptab, pw, npt) : hybrid_table<T>(p2.data(), np, o->ptab, pw, npt);
  if (!ok) { delete o; return nullptr; }
  if (bintype == FCFC_GPU_BIN_SMU) {
    o->mutab.resize((size_t) nmu * nmu);
    int j = 0;
    for (int i = 0; i < nmu * nmu; i++) { while ((j + 1) * (j + 1) <= i) j++; o->mutab[i] = (unsigned char) j; }
  }
  o->s2bin.assign(reinterpret_cast<unsigned char *>(s2.data()), reinterpret_cast<unsigned char *>(s2.data() + ns + 1));
  if (spi) o->pbin.assign(reinterpret_cast<unsigned char *>(p2.data()), reinterpret_cast<unsigned char *>(p2.data() + np + 1));
  o->rescale = (double) fac;
  B.bintype = bintype; B.periodic = periodic; B.is_float = sizeof(T) == 4; B.tabtype = tabtype;
  B.ns = ns; B.np = spi ? np : 0; B.nmu = (bintype == FCFC_GPU_BIN_SMU) ? nmu : 1;
  B.swidth = sw; B.pwidth = pw; B.with_mu_one = with_mu_one; B.arith = arith;
  B.s2bin = o->s2bin.data(); B.pbin = spi ? o->pbin.data() : nullptr;
  B.stab = o->stab.data(); B.ptab = spi ? o->ptab.data() : nullptr;
  B.mutab = (bintype == FCFC_GPU_BIN_SMU) ? o->mutab.data() : nullptr;
  B.nstab = nst; B.nptab = npt;
  for (int d = 0; d < 3; d++) B.bsize[d] = periodic ? (double) ((T) box[d] * fac) : 0.0;       // setup_cf.c:679
  return o;
}

}  // namespace

extern "C" fcfc_gpu_bins_owner *fcfc_gpu_bins_create(int periodic, int is_float, int bintype, int linear, double smin, double ds,
                                                     double pmin, double dpi, const double *sedge, int ns, const double *pedge,
                                                     int np, int nmu, const double box[3], int with_mu_one, int arith) {
  static const double zero[3] = {0, 0, 0};
  if (!box) box = zero;
  return is_float ? make_bins<float>(periodic, bintype, linear, smin, ds, pmin, dpi, sedge, ns, pedge, np, nmu, box, with_mu_one, arith)
                  : make_bins<double>(periodic, bintype, linear, smin, ds, pmin, dpi, sedge, ns, pedge, np, nmu, box, with_mu_one, arith);
}
extern "C" const fcfc_gpu_bins *fcfc_gpu_bins_get(const fcfc_gpu_bins_owner *o) { return o ? &o->bins : nullptr; }
extern "C" double fcfc_gpu_bins_rescale(const fcfc_gpu_bins_owner *o) { return o ? o->rescale : 0; }
extern "C" void fcfc_gpu_bins_free(fcfc_gpu_bins_owner *o) { delete o; }
