// fcfc_b200/csrc/count_kernel.cuh -- the pair-counting kernels (sm_100a).
//
// Replaces the reference's hot loops
//   count_dual_node / count_single_node   src/fcfc/2pt_box/metric_common.c:980-1482, 2003-2295
//   compute_dist_hist_scalar              src/fcfc/2pt_box/metric_common.c:140-235
//   compute_dist_vector (+update_hist)    src/fcfc/2pt_box/metric_common.c:377-534, 774-957
//   survey variants                       src/fcfc/2pt/metric_common.c:142-259, 283-460
// and the dual-tree traversal around them (dual_tree.c:245-373) by a cell-list sweep:
//
//   * points are sorted by grid cell (z fastest), one float4/double4 (x, y, z, |x|^2) per point;
//   * a work item is a tile of <= 32*R consecutive points of one primary cell; each *warp* pulls
//     items from a global queue and keeps its tile in registers (R points per lane);
//   * for every row (dx, dy) of the neighbour stencil the secondary points of the cells
//     (dz_lo..dz_hi) are contiguous in memory; the warp stages them 32 at a time through its
//     private shared-memory buffer (coalesced 128-bit loads, periodic shift applied while
//     staging) and every lane reads them back with broadcast LDS.128;
//   * per pair: 3 subtractions + 1 multiply + 2 FMA (6 FP32 instructions, FMA mode) or
//     3 sub + 3 mul + 2 add (scalar-parity mode), one compare; accepted pairs go through the
//     reference's lookup tables into a per-block shared-memory histogram (32-bit counters,
//     flushed lock-free to 64-bit global counters; FP64 sums for weighted counts).
//
// No tensor cores: the work is FP32/FP64 CUDA-core arithmetic plus shared-memory atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace fcfc {

enum { BIN_ISO = 0, BIN_SMU = 1, BIN_SPI = 2 };
enum { ARITH_SCALAR = 0, ARITH_FMA = 1 };

constexpr int kWarpsPerBlock = 16;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxRows = 1024;          // stencil rows kept in shared memory
constexpr int kSegPieceMax = 1 << 19;   // secondary points per overflow-accounting piece

template <class T> struct Vec4;
template <> struct __align__(16) Vec4<float> { float x, y, z, s; };
template <> struct __align__(16) Vec4<double> { double x, y, z, s; };

// Arithmetic with explicit rounding: the _rn intrinsics are never contracted into FMAs by nvcc,
// which is what the scalar parity mode needs (the reference is built with -std=c99, i.e.
// -ffp-contract=off: SURVEY.md section 5).
template <class T> struct Ar;
template <> struct Ar<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float divrz(float a, float b) { return __fdiv_rz(a, b); }
  static __device__ __forceinline__ int toint(float a) { return __float2int_rz(a); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float eps() { return FLT_EPSILON; }
  static __device__ __forceinline__ float far() { return 1e15f; }
  static __device__ __forceinline__ float huge() { return 1e30f; }
};
template <> struct Ar<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double divrz(double a, double b) { return __ddiv_rz(a, b); }
  static __device__ __forceinline__ int toint(double a) { return __double2int_rz(a); }
  static __device__ __forceinline__ double abs(double a) { return fabs(a); }
  static __device__ __forceinline__ double eps() { return DBL_EPSILON; }
  static __device__ __forceinline__ double far() { return 1e100; }
  static __device__ __forceinline__ double huge() { return 1e200; }
};

// Kernel arguments (plain data, passed by value).
template <class T> struct CountParams {
  // cell-sorted catalogues: primary (1) and secondary (2)
  const Vec4<T> *pos1; const T *w1;
  const Vec4<T> *pos2; const T *w2; const int *cell_start2;
  // work items of the primary catalogue: cell id, first point, number of points (<= 32*R)
  const int *item_cell; const int *item_off; const int *item_cnt;
  int item_begin, item_end;             // this launch (shard) processes items [begin, end)
  unsigned int *work_counter;           // global queue head (starts at 0)
  // cell grid (cell id = (ix*nc[1] + iy)*nc[2] + iz)
  int nc[3]; int periodic;
  T bsize[3];
  // neighbour stencil: rows (dx, dy, dz_lo, dz_hi)
  const int4 *rows; int nrows;
  // binning
  T s2min, s2max, pmin, pmax, premax, nmu2f;
  int ns, np, nmu2, ntot, soff, poff;
  int tab_hybrid, swidth, pwidth, with_mu_one, smin0, pmin0, mu_is_sqrt;
  const uint8_t *stab; const uint8_t *ptab; const uint8_t *mutab;
  int nstab, nptab;                     // entries
  const T *s2bin; const T *pbin;
  int isauto;
  // outputs
  unsigned long long *ghist_i; double *ghist_d;
  unsigned long long *gevals;           // [0] candidate pair evaluations
};

// Shared-memory plan (dynamic): [hist][tables][edges][rows][per-warp staging]
struct SmemPlan {
  int off_hist, off_stab, off_ptab, off_mutab, off_s2bin, off_pbin, off_rows, off_stage, stage_per_warp, total;
};

template <class T, bool WT>
__host__ __device__ inline SmemPlan make_smem_plan(int ntot, int nstab_bytes, int nptab_bytes, int nmutab_bytes,
                                                   int ns, int np, int nrows, bool smem_hist) {
  SmemPlan p;
  int o = 0;
  auto al = [](int v) { return (v + 15) & ~15; };
  p.off_hist = o; o += smem_hist ? al(ntot * (WT ? 8 : 4)) : 0;
  p.off_stab = o; o += al(nstab_bytes);
  p.off_ptab = o; o += al(nptab_bytes);
  p.off_mutab = o; o += al(nmutab_bytes);
  p.off_s2bin = o; o += al((ns + 1) * (int) sizeof(T));
  p.off_pbin = o; o += al((np + 1) * (int) sizeof(T));
  p.off_rows = o; o += al(nrows * 16);
  p.off_stage = o;
  p.stage_per_warp = 32 * (int) sizeof(Vec4<T>) + (WT ? 32 * (int) sizeof(T) : 0);
  o += kWarpsPerBlock * p.stage_per_warp;
  p.total = o;
  return p;
}

template <class T> struct BlockCtx {
  unsigned int *hist_u;         // shared or global 32/64-bit counters (unweighted)
  double *hist_d;               // weighted sums
  const uint8_t *stab, *ptab, *mutab;
  const T *s2bin, *pbin;
  unsigned int *blk_evals;      // shared overflow accounting
};

// ---------------------------------------------------------------------------------------------
// Table lookup with the hybrid walk-down (metric_common.c:56-64).
template <class T>
__device__ __forceinline__ int lut(const uint8_t *tab, int width, int hybrid, int idx, int nbin, T val, const T *edges) {
  int v = width ? (int) reinterpret_cast<const uint16_t *>(tab)[idx] : (int) tab[idx];
  if (hybrid && v >= nbin) {
    v -= nbin;
    while (v != 0 && val < edges[v]) v--;
  }
  return v;
}

// Second half of a pair: everything after the cheap range test.  Returns the histogram bin or -1.
//   d2  : squared separation (ISO, SMU; survey SPI: s^2 = s - t before the pi^2 subtraction)
//   aux : box SMU dz^2 | box SPI pi | survey SMU/SPI t (twice the dot product)
template <class T, int BIN, bool BOX, int ARITH, bool GENERIC>
__device__ __forceinline__ int finish_pair(const CountParams<T> &P, const BlockCtx<T> &C, T d2, T aux, T as, T bs) {
  using A = Ar<T>;
  int pb = 0;
  T pival = aux;
  if (BIN == BIN_SMU || (BIN == BIN_SPI && !BOX)) {
    T num = aux;
    if (!BOX) {               // survey: pi^2 = (s1 - s2)^2 / (s + t), 2pt/metric_common.c:180-181
      T s = A::add(as, bs);
      T d = A::sub(as, bs);
      num = A::div(A::mul(d, d), A::add(s, aux));
    }
    if (BIN == BIN_SPI) {     // survey SPI: range tests of 2pt/metric_common.c:185-205
      if (num >= P.pmax || (!P.pmin0 && num < P.pmin)) return -1;
      d2 = A::sub(d2, num);
      if (d2 >= P.s2max || (!P.smin0 && d2 < P.s2min)) return -1;
      pival = num;
    } else {
      int m;
      if (ARITH == ARITH_SCALAR) {      // metric_common.c:184
        m = (d2 < A::eps()) ? 0 : A::toint(A::mul(A::div(num, d2), P.nmu2f));
      } else {                          // metric_common.c:472-494 (AVX-512 branch)
        T q = A::mul(num, P.nmu2f);
        q = (d2 >= A::eps()) ? A::divrz(q, d2) : (T) 0;
        m = (q < P.nmu2f) ? A::toint(q) : P.nmu2;
      }
      if (m >= P.nmu2) {
        if (P.with_mu_one) m = P.nmu2 - 1; else return -1;
      }
      pb = C.mutab[m];
    }
  }
  int sidx = A::toint(d2) - P.soff;
  if (!BOX) sidx = max(sidx, 0);        // survey s_perp^2 can be slightly negative ((int) maps (-1,0) to 0)
  int sb;
  if (GENERIC) sb = lut<T>(C.stab, P.swidth, P.tab_hybrid, sidx, P.ns, d2, C.s2bin);
  else sb = C.stab[sidx];
  if (BIN == BIN_SPI) {
    int pidx = A::toint(pival) - P.poff;
    if (GENERIC) pb = lut<T>(C.ptab, P.pwidth, P.tab_hybrid, pidx, P.np, pival, C.pbin);
    else pb = C.ptab[pidx];
  }
  return sb + pb * P.ns;
}

// First half: distances and the cheap range test.  Outputs d2 and aux as described above.
template <class T, int BIN, bool BOX, int ARITH, bool GENERIC>
__device__ __forceinline__ bool eval_pair(const CountParams<T> &P, T ax, T ay, T az, T as,
                                          const Vec4<T> &b, T &d2, T &aux) {
  using A = Ar<T>;
  bool ok;
  if (BOX || BIN == BIN_ISO) {
    T dx = A::sub(ax, b.x), dy = A::sub(ay, b.y), dz = A::sub(az, b.z);
    if (BIN == BIN_SPI) {                       // box (s_perp, pi): metric_common.c:157-165, 416-424
      aux = A::abs(dz);
      d2 = (ARITH == ARITH_SCALAR) ? A::add(A::mul(dx, dx), A::mul(dy, dy)) : A::fma(dy, dy, A::mul(dx, dx));
      ok = (aux < P.pmax) && (d2 < P.s2max);
      if (GENERIC) ok = ok && (P.pmin0 || aux >= P.pmin);
    } else {
      T dz2 = A::mul(dz, dz);
      if (ARITH == ARITH_SCALAR) d2 = A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), dz2);     // :170-172
      else if (BOX) d2 = A::fma(dy, dy, A::fma(dx, dx, dz2));                                   // :426-430
      else d2 = A::fma(dz, dz, A::fma(dy, dy, A::mul(dx, dx)));                                 // 2pt/:330-333
      aux = dz2;
      ok = d2 < P.s2max;
    }
  } else {                                      // survey (s,mu) / (s_perp,pi): 2pt/metric_common.c:169-172, 341-357
    T t;
    if (ARITH == ARITH_SCALAR) t = A::mul(A::add(A::add(A::mul(ax, b.x), A::mul(ay, b.y)), A::mul(az, b.z)), (T) 2);
    else t = A::mul(A::fma(az, b.z, A::fma(ay, b.y, A::mul(ax, b.x))), (T) 2);
    T s = A::add(as, b.s);
    d2 = A::sub(s, t);
    aux = t;
    ok = d2 < ((BIN == BIN_SPI) ? P.premax : P.s2max);
  }
  if (GENERIC && BIN != BIN_SPI) ok = ok && (P.smin0 || d2 >= P.s2min);
  if (GENERIC && BIN == BIN_SPI && BOX) ok = ok && (P.smin0 || d2 >= P.s2min);
  return ok;
}

template <class T, bool WT, bool SMEMHIST>
__device__ __forceinline__ void hist_add(const BlockCtx<T> &C, int bin, T wa, T wb) {
  if (WT) {
    T w = Ar<T>::mul(wa, wb);                   // product in `real`, sum in double: metric_common.c:216-231
    atomicAdd(&C.hist_d[bin], (double) w);
  } else if (SMEMHIST) {
    atomicAdd(&C.hist_u[bin], 1u);
  } else {
    atomicAdd(reinterpret_cast<unsigned long long *>(C.hist_u) + bin, 1ull);
  }
}

// Lock-free drain of the 32-bit shared counters into the 64-bit global histogram: safe while other
// warps keep counting because every word is taken with an atomic exchange.
__device__ __forceinline__ void sweep_hist(unsigned int *h, unsigned long long *g, int ntot, int lane) {
  for (int i = lane; i < ntot; i += 32) {
    unsigned int v = atomicExch(&h[i], 0u);
    if (v) atomicAdd(&g[i], (unsigned long long) v);
  }
}

// One chunk of <= 32 staged secondary points against the R register-resident primaries of each lane.
template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, bool SMEMHIST, int R, bool SELF>
__device__ __forceinline__ void do_chunk(const CountParams<T> &P, const BlockCtx<T> &C,
                                         const Vec4<T> *sbuf, const T *wbuf, int nj,
                                         const T (&ax)[R], const T (&ay)[R], const T (&az)[R], const T (&as)[R],
                                         const T (&aw)[R], int jglob0, int iglob0) {
#pragma unroll 2
  for (int j = 0; j < nj; j++) {
    const Vec4<T> b = sbuf[j];
    T bw = (T) 1;
    if (WT) bw = wbuf[j];
#pragma unroll
    for (int r = 0; r < R; r++) {
      T d2, aux;
      bool ok = eval_pair<T, BIN, BOX, ARITH, GENERIC>(P, ax[r], ay[r], az[r], as[r], b, d2, aux);
      if (SELF) ok = ok && (jglob0 + j > iglob0 + r * 32);      // unordered pairs once: metric_common.c:2017-2018
      if (ok) {
        int bin = finish_pair<T, BIN, BOX, ARITH, GENERIC>(P, C, d2, aux, as[r], b.s);
        if (bin >= 0) hist_add<T, WT, SMEMHIST>(C, bin, aw[r], bw);
      }
    }
  }
}

template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, bool SMEMHIST, int R>
__global__ void __launch_bounds__(kThreads, 1) count_kernel(const CountParams<T> P) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ unsigned int s_blk_evals;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nmutab = (BIN == BIN_SMU) ? P.nmu2 : 0;
  const SmemPlan pl = make_smem_plan<T, WT>(P.ntot, P.nstab * (P.swidth ? 2 : 1), P.nptab * (P.pwidth ? 2 : 1),
                                            nmutab, P.ns, P.np, P.nrows, SMEMHIST);
  BlockCtx<T> C;
  C.hist_u = SMEMHIST ? reinterpret_cast<unsigned int *>(smem + pl.off_hist) : reinterpret_cast<unsigned int *>(P.ghist_i);
  C.hist_d = SMEMHIST ? reinterpret_cast<double *>(smem + pl.off_hist) : P.ghist_d;
  uint8_t *s_stab = smem + pl.off_stab, *s_ptab = smem + pl.off_ptab, *s_mutab = smem + pl.off_mutab;
  T *s_s2bin = reinterpret_cast<T *>(smem + pl.off_s2bin), *s_pbin = reinterpret_cast<T *>(smem + pl.off_pbin);
  int4 *s_rows = reinterpret_cast<int4 *>(smem + pl.off_rows);
  C.stab = s_stab; C.ptab = s_ptab; C.mutab = s_mutab; C.s2bin = s_s2bin; C.pbin = s_pbin;
  C.blk_evals = &s_blk_evals;

  // ---- block prologue: zero the histogram, stage tables / edges / stencil rows ----
  if (SMEMHIST) {
    if (WT) for (int i = threadIdx.x; i < P.ntot; i += kThreads) C.hist_d[i] = 0.0;
    else for (int i = threadIdx.x; i < P.ntot; i += kThreads) C.hist_u[i] = 0u;
  }
  for (int i = threadIdx.x; i < P.nstab * (P.swidth ? 2 : 1); i += kThreads) s_stab[i] = P.stab[i];
  if (BIN == BIN_SPI) for (int i = threadIdx.x; i < P.nptab * (P.pwidth ? 2 : 1); i += kThreads) s_ptab[i] = P.ptab[i];
  if (BIN == BIN_SMU) for (int i = threadIdx.x; i < nmutab; i += kThreads) s_mutab[i] = P.mutab[i];
  for (int i = threadIdx.x; i <= P.ns; i += kThreads) s_s2bin[i] = P.s2bin[i];
  if (BIN == BIN_SPI) for (int i = threadIdx.x; i <= P.np; i += kThreads) s_pbin[i] = P.pbin[i];
  for (int i = threadIdx.x; i < P.nrows; i += kThreads) s_rows[i] = P.rows[i];
  if (threadIdx.x == 0) s_blk_evals = 0;
  __syncthreads();

  Vec4<T> *sbuf = reinterpret_cast<Vec4<T> *>(smem + pl.off_stage + warp * pl.stage_per_warp);
  T *wbuf = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(sbuf) + 32 * sizeof(Vec4<T>));
  unsigned long long my_evals = 0;
  const int ncy = P.nc[1], ncz = P.nc[2];

  // ---- persistent warp loop over work items ----
  while (true) {
    int item = 0;
    if (lane == 0) item = P.item_begin + (int) atomicAdd(P.work_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= P.item_end) break;
    const int cell = P.item_cell[item], t0 = P.item_off[item], cnt = P.item_cnt[item];
    const int iz = cell % ncz, iy = (cell / ncz) % ncy, ix = cell / (ncz * ncy);

    // primaries: lane holds points t0 + r*32 + lane; padding lanes sit far away (never in range)
    T px[R], py[R], pz[R], ps[R], pw[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int k = r * 32 + lane;
      if (k < cnt) {
        Vec4<T> v = P.pos1[t0 + k];
        px[r] = v.x; py[r] = v.y; pz[r] = v.z; ps[r] = v.s;
        pw[r] = WT ? P.w1[t0 + k] : (T) 1;
      } else {
        if (BOX || BIN == BIN_ISO) { px[r] = py[r] = pz[r] = Ar<T>::far(); ps[r] = 0; }
        else { px[r] = py[r] = pz[r] = 0; ps[r] = Ar<T>::huge(); }
        pw[r] = 0;
      }
    }

    // One contiguous range [b, e) of secondary points with per-axis image shifts.
    // sa: shift added to the primaries, sb: shift added to the secondaries (the lower point gets +L).
    auto sweep_range = [&](int b, int e, T sax, T say, T saz, T sbx, T sby, T sbz, bool self) {
      T ax[R], ay[R], az[R];
#pragma unroll
      for (int r = 0; r < R; r++) {
        ax[r] = BOX ? Ar<T>::add(px[r], sax) : px[r];
        ay[r] = BOX ? Ar<T>::add(py[r], say) : py[r];
        az[r] = BOX ? Ar<T>::add(pz[r], saz) : pz[r];
      }
      while (b < e) {
        const int piece_end = min(e, b + kSegPieceMax);
        // overflow accounting of the 32-bit shared counters (see sweep_hist)
        if (SMEMHIST && !WT) {
          unsigned int add = (unsigned int) (piece_end - b) * (unsigned int) cnt, old = 0;
          if (lane == 0) old = atomicAdd(C.blk_evals, add);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old + add >= 0x80000000u || old + add < old) {
            if (lane == 0) atomicExch(C.blk_evals, 0u);
            sweep_hist(C.hist_u, P.ghist_i, P.ntot, lane);
          }
        }
        {
          unsigned long long ev = (unsigned long long) (piece_end - b) * (unsigned long long) cnt;
          if (self && b == t0) ev -= (unsigned long long) cnt * (unsigned long long) (cnt + 1) / 2;   // i < j only
          my_evals += ev;
        }
        // register-prefetched staging: one point per lane
        Vec4<T> nxt; T nxtw = 0;
        int jn = b + lane;
        if (jn < piece_end) { nxt = P.pos2[jn]; if (WT) nxtw = P.w2[jn]; }
        for (int c0 = b; c0 < piece_end; c0 += 32) {
          __syncwarp();
          if (BOX) { nxt.x = Ar<T>::add(nxt.x, sbx); nxt.y = Ar<T>::add(nxt.y, sby); nxt.z = Ar<T>::add(nxt.z, sbz); }
          sbuf[lane] = nxt;
          if (WT) wbuf[lane] = nxtw;
          __syncwarp();
          jn = c0 + 32 + lane;
          if (jn < piece_end) { nxt = P.pos2[jn]; if (WT) nxtw = P.w2[jn]; }
          const int nj = min(32, piece_end - c0);
          if (self && c0 < t0 + cnt)
            do_chunk<T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST, R, true>(P, C, sbuf, wbuf, nj, ax, ay, az, ps, pw, c0, t0 + lane);
          else
            do_chunk<T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST, R, false>(P, C, sbuf, wbuf, nj, ax, ay, az, ps, pw, 0, 0);
        }
        b = piece_end;
      }
    };

    // Sweep list: q = -1 is the tile's own cell (auto counts: pairs i < j only, metric_common.c:2017);
    // q = 3*row + image enumerates, for every stencil row, the three periodic images of its z run
    // (below the box: the primaries get +L in z; inside; above the box: the secondaries get +L).
    // A single call site keeps the unrolled pair loops in the instruction cache.
    const int nq = (P.periodic ? 3 : 1) * P.nrows;
    for (int q = P.isauto ? -1 : 0; q < nq; q++) {
      int b, e;
      T sax = 0, say = 0, saz = 0, sbx = 0, sby = 0, sbz = 0;
      if (q < 0) { b = t0; e = P.cell_start2[cell + 1]; }
      else {
        const int ri = P.periodic ? q / 3 : q, img = P.periodic ? q - 3 * ri : 1;
        const int4 row = s_rows[ri];
        int jx = ix + row.x, jy = iy + row.y;
        int zlo = iz + row.z, zhi = iz + row.w;
        if (P.periodic) {
          if (jx >= P.nc[0]) { jx -= P.nc[0]; sbx = P.bsize[0]; } else if (jx < 0) { jx += P.nc[0]; sax = P.bsize[0]; }
          if (jy >= ncy) { jy -= ncy; sby = P.bsize[1]; } else if (jy < 0) { jy += ncy; say = P.bsize[1]; }
          if (img == 0) { zhi = min(zhi, -1) + ncz; zlo += ncz; saz = P.bsize[2]; }
          else if (img == 1) { zlo = max(zlo, 0); zhi = min(zhi, ncz - 1); }
          else { zlo = max(zlo, ncz) - ncz; zhi -= ncz; sbz = P.bsize[2]; }
        } else {
          if (jx < 0 || jx >= P.nc[0] || jy < 0 || jy >= ncy) continue;
          zlo = max(zlo, 0); zhi = min(zhi, ncz - 1);
        }
        if (zlo > zhi) continue;
        const int rowbase = (jx * ncy + jy) * ncz;
        b = P.cell_start2[rowbase + zlo]; e = P.cell_start2[rowbase + zhi + 1];
      }
      if (b >= e) continue;
      sweep_range(b, e, sax, say, saz, sbx, sby, sbz, q < 0);
    }
  }

  // ---- block epilogue: flush the histogram ----
  __syncthreads();
  if (SMEMHIST) {
    if (WT) {
      for (int i = threadIdx.x; i < P.ntot; i += kThreads) { double v = C.hist_d[i]; if (v != 0.0) atomicAdd(&P.ghist_d[i], v); }
    } else {
      for (int i = threadIdx.x; i < P.ntot; i += kThreads) { unsigned int v = C.hist_u[i]; if (v) atomicAdd(&P.ghist_i[i], (unsigned long long) v); }
    }
  }
  // pair-evaluation counter: one atomic per warp
  if (lane == 0 && my_evals) atomicAdd(P.gevals, my_evals);
}

}  // namespace fcfc
