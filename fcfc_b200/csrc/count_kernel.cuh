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
//     3 sub + 3 mul + 2 add (scalar-parity mode) and one compare.  Pairs that pass the range test
//     are NOT binned in place (that would leave most lanes idle in a divergent branch): each lane
//     pushes (d^2, aux[, w]) on its own stack in shared memory (one predicated store + pointer bump);
//   * when a stack is nearly full the warp drains: every lane pops its own entries, four per
//     iteration, so all 32 lanes run the expensive part (bins, histogram update) together.  Box and
//     isotropic counts compute both bins from one rsqrt.approx with fixed-point edge detection; pairs
//     within the error band of a bin edge are re-binned with the exact IEEE sequence of the reference
//     (results are bit-exact).  The histogram is per block in shared memory (32-bit counters flushed
//     lock-free to 64-bit global counters; FP64 sums for weighted counts);
//   * secondary cells that lie entirely within range of the tile's cell ("dense" cells, marked by the host per
//     stencil row) skip the stacks: the pair loop bins their pairs in place (do_chunk_dense).
//
// No tensor cores: the work is FP32/FP64 CUDA-core arithmetic plus shared-memory atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

// Shared-memory loads written as inline PTX are volatile (see the note above QOps); 0 restores plain asm for A/B builds.
#ifndef FCFC_LDS_VOLATILE
#define FCFC_LDS_VOLATILE 1
#endif
#if FCFC_LDS_VOLATILE
#define FCFC_LDS_ASM asm volatile
#else
#define FCFC_LDS_ASM asm
#endif

namespace fcfc {

enum { BIN_ISO = 0, BIN_SMU = 1, BIN_SPI = 2 };
enum { ARITH_SCALAR = 0, ARITH_FMA = 1 };

// Warps per block (one block per SM): as many as the register file allows -- the kernels are latency bound
// at 4 warps per scheduler.  float variants fit in 80 registers (24 warps), double variants <= 128 (16 warps).
#ifndef FCFC_WARPS_F32
#define FCFC_WARPS_F32 24
#endif
template <class T> struct BlockShape { static constexpr int kWarps = (sizeof(T) == 4) ? FCFC_WARPS_F32 : 16; static constexpr int kThreads = kWarps * 32; };
constexpr int kMaxRows = 1024;          // stencil rows kept in shared memory
#ifndef FCFC_EVAL_UNROLL
#define FCFC_EVAL_UNROLL 1
#endif
constexpr int kEvalUnroll = FCFC_EVAL_UNROLL;   // secondary points per iteration of the pair loop
constexpr int kSegPieceMax = 1 << 19;   // secondary points per overflow-accounting piece
#ifndef FCFC_DENSE
#define FCFC_DENSE 1                    // 0 compiles the dense-cell path out (experiments)
#endif

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
  const int *item_order;                // items sorted by decreasing estimated cost (longest first)
  int nitem, part, nparts;              // this launch (shard) processes order[part], order[part + nparts], ...
  int nsplit;                           // every tile's sweep list is cut into nsplit work items (small problems)
  unsigned int *work_counter;           // global queue head (starts at 0)
  // cell grid (cell id = (ix*nc[1] + iy)*nc[2] + iz)
  int nc[3]; int periodic;
  T bsize[3];
  // neighbour stencil: rows (dx, dy, dz_lo, dz_hi)
  const int4 *rows; int nrows;
  // per row: the sub-range (dz_lo, dz_hi) of its cells that lie entirely within the maximum separation of every point of
  // the tile's cell ("dense" cells: every pair is in range, see do_chunk_dense); null when the dense path is not used
  const int2 *rows_in;
  // binning
  T s2max_pre;                          // survey (s_perp,pi): padded s2max of the division-free pre-test (eval_pair)
  T s2min, s2max, pmin, pmax, premax, pmax_pre, nmu2f;
  int ns, np, nmu2, ntot, soff, poff;
  int tab_hybrid, swidth, pwidth, with_mu_one, smin0, pmin0;
  int mu_is_sqrt, stab_is_sqrt, ptab_is_ident;     // tables that equal floor(sqrt(i)) / i are computed, not looked up
  // fast-bin fixed-point scales: s and nmu*mu are truncated after multiplication by 2^ks / 2^km (see fast_bins)
  float fb_sscale, fb_mscale; unsigned int fb_smask, fb_mmask, fb_sshift, fb_mshift;
  unsigned int fb_smul, fb_mmul, fb_bias;        // 2^(32 - shift) (shifts as IMAD.HI) and the bias of the fast bins, (0x4B000000 >> ks) + (0x4B000000 >> km) * ns
  const uint8_t *stab; const uint8_t *ptab; const uint8_t *mutab;
  int nstab, nptab;                     // entries
  const T *s2bin; const T *pbin;
  int isauto;
  int tabs_global;                      // lookup tables too large for shared memory: read them from global memory
  int hist_copies;                      // weighted shared-memory histogram: copies (a power of two <= 32), lane l adds to copy l mod copies
  int qdepth;                           // entries per lane of the accepted-pair queues (a multiple of 4)
  int qkeep;                            // entries a drain leaves on the fullest stack
  // float pre-filter of the double-precision kernels (count_kernel_pf.cuh): limits padded by the worst-case float error
  float pf_d2lim;                       // d^2 (box (s_perp,pi): s_perp^2); survey (s_perp,pi): the searched sphere
  float pf_plim;                        // box (s_perp,pi): |dz|; survey (s_perp,pi): pi^2 (cylinder test 1)
  float pf_s2lim;                       // survey (s_perp,pi): s_perp^2 (cylinder test 2)
  // double precision at float speed (count_kernel_df.cuh): the cell grid (local origins), the padded range limit of the
  // float d^2 and the squared separation below which (s,mu) pairs always take the exact path
  T gorg[3], gcs[3];
  float df_d2lim, df_s1sq;
  // classified staging of the float kernels (count_kernel_cl.cuh): a staged point is dropped when its squared distance to
  // the nearest point of the tile's box exceeds cl_skip, and binned in place when the farthest corner is within cl_dense
  float cl_skip, cl_dense;
  // outputs
  unsigned long long *ghist_i; double *ghist_d;
  unsigned long long *gevals;           // [0] candidate pair evaluations (n_a * n_b over the swept cell ranges), [3] distance evaluations made (count_kernel_cl.cuh)
};

// Shared-memory plan (dynamic): [hist][tables][edges][rows][per-warp staging]
struct SmemPlan {
  int off_hist, off_stab, off_ptab, off_mutab, off_s2bin, off_pbin, off_rows, off_misc, off_stage, stage_per_warp, off_queue, queue_per_warp, total;
};

// Words of `real` per queue entry: ISO d2[,w]; box (s,mu)/(s_perp,pi) d2,aux[,w]; survey t,s1,s2[,w].
template <int BIN, bool BOX, bool WT> struct QFmt {
  static constexpr int NW = (BIN == BIN_ISO) ? (WT ? 2 : 1) : (BOX ? (WT ? 4 : 2) : 4);
};

template <class T, bool WT>
__host__ __device__ inline SmemPlan make_smem_plan(int ntot, int nstab_bytes, int nptab_bytes, int nmutab_bytes,
                                                   int ns, int np, int nrows, bool smem_hist, int qwords, int qdepth, bool tabs_global, int hist_copies = 1) {
  constexpr int kWarpsPerBlock = BlockShape<T>::kWarps;
  SmemPlan p;
  int o = 0;
  auto al = [](int v) { return (v + 15) & ~15; };
  // unweighted: bins 0 .. ntot + ns (the fast drain may land a flagged pair one row / column outside) + 32 per-lane dump slots
  p.off_hist = o; o += smem_hist ? al(WT ? ntot * 8 * hist_copies : (ntot + ns + 1 + 32) * 4) : 0;
  p.off_stab = o; o += tabs_global ? 0 : al(nstab_bytes);
  p.off_ptab = o; o += tabs_global ? 0 : al(nptab_bytes);
  p.off_mutab = o; o += tabs_global ? 0 : al(nmutab_bytes);
  p.off_s2bin = o; o += al((ns + 1) * (int) sizeof(T));
  p.off_pbin = o; o += al((np + 1) * (int) sizeof(T));
  p.off_rows = o; o += al(nrows * 16);
  p.off_misc = o; o += 16;                      // overflow counter of the 32-bit histogram
  p.off_stage = o;
  p.stage_per_warp = 32 * (int) sizeof(Vec4<T>) + (WT ? 32 * (int) sizeof(T) : 0);
  o += kWarpsPerBlock * p.stage_per_warp;
  p.off_queue = o;
  p.queue_per_warp = qdepth * 32 * qwords * (int) sizeof(T);
  o += kWarpsPerBlock * p.queue_per_warp;
  o += 4 * 32 * qwords * (int) sizeof(T);       // the four-entry drain reads (and ignores) up to three slots past a queue
  p.total = o;
  return p;
}

template <class T> struct BlockCtx {
  unsigned int *hist_u;         // shared or global 32/64-bit counters (unweighted)
  double *hist_d;               // weighted sums
  int hmul, hoff;               // weighted bin -> slot: bin * hmul + hoff (lane-private copies when bins are few)
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
//   aux : box SMU dz (squared again here: same rounding as in eval_pair) | box SPI pi | survey SMU/SPI t
template <class T, int BIN, bool BOX, int ARITH, bool GENERIC>
__device__ __forceinline__ int finish_pair(const CountParams<T> &P, const BlockCtx<T> &C, T d2, T aux, T as, T bs) {
  using A = Ar<T>;
  int pb = 0;
  T pival = aux;
  if (BIN == BIN_SMU || (BIN == BIN_SPI && !BOX)) {
    T num = BOX ? A::mul(aux, aux) : aux;
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
                                          const Vec4<T> &b, const T s2lim, T &d2, T &aux) {
  using A = Ar<T>;
  bool ok;
  if (BOX || BIN == BIN_ISO) {
    T dx = A::sub(ax, b.x), dy = A::sub(ay, b.y), dz = A::sub(az, b.z);
    if (BIN == BIN_SPI) {                       // box (s_perp, pi): metric_common.c:157-165, 416-424
      aux = A::abs(dz);
      d2 = (ARITH == ARITH_SCALAR) ? A::add(A::mul(dx, dx), A::mul(dy, dy)) : A::fma(dy, dy, A::mul(dx, dx));
      ok = (aux < P.pmax) && (d2 < s2lim);
      if (GENERIC) ok = ok && (P.pmin0 || aux >= P.pmin);
    } else {
      T dz2 = A::mul(dz, dz);
      if (ARITH == ARITH_SCALAR) d2 = A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), dz2);     // :170-172
      else if (BOX) d2 = A::fma(dy, dy, A::fma(dx, dx, dz2));                                   // :426-430
      else d2 = A::fma(dz, dz, A::fma(dy, dy, A::mul(dx, dx)));                                 // 2pt/:330-333
      aux = (BOX && BIN == BIN_SMU) ? dz : dz2;
      ok = d2 < s2lim;
    }
  } else {                                      // survey (s,mu) / (s_perp,pi): 2pt/metric_common.c:169-172, 341-357
    // t = 2 (x1 . x2): the caller passes the primary already doubled (exact), which saves the final multiplication;
    // every product and sum is then exactly twice the reference's, rounding included
    T t;
    if (ARITH == ARITH_SCALAR) t = A::add(A::add(A::mul(ax, b.x), A::mul(ay, b.y)), A::mul(az, b.z));
    else t = A::fma(az, b.z, A::fma(ay, b.y, A::mul(ax, b.x)));
    T s = A::add(as, b.s);
    d2 = A::sub(s, t);
    aux = t;
    ok = d2 < s2lim;                            // survey (s_perp, pi): s2max + p2max, see engine.cu
    if (BIN == BIN_SPI) {
      // cheap necessary condition for pi^2 = d*d / (s + t) < p2max, without the division (the exact test is
      // repeated in finish_pair): d*d < (s + t) * p2max * (1 + 8 eps).  Cuts the queue traffic of survey
      // (s_perp, pi) counts, whose accepted region is a thin cylinder inside the searched sphere.
      const T d = A::sub(as, b.s), dd = A::mul(d, d), st = A::add(s, t);
      ok = ok && (dd < A::mul(st, P.pmax_pre));
      // and for s_perp^2 = s^2 - pi^2 < s2max, again without the division: (s^2 - s2max') (s + t) < d*d, with s2max'
      // padded by the host for every rounding on the way.  Three quarters of the sphere-shaped candidates fail it.
      ok = ok && (A::mul(A::sub(d2, P.s2max_pre), st) < dd);
    }
  }
  if (GENERIC && BIN != BIN_SPI) ok = ok && (P.smin0 || d2 >= P.s2min);
  if (GENERIC && BIN == BIN_SPI && BOX) ok = ok && (P.smin0 || d2 >= P.s2min);
  return ok;
}

template <class T, bool WT, bool SMEMHIST>
__device__ __forceinline__ void hist_add(const BlockCtx<T> &C, int bin, T w) {
  if (WT) {
    atomicAdd(&C.hist_d[bin * C.hmul + C.hoff], (double) w);    // product formed in `real`, summed in double: metric_common.c:216-231
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
    const int v = (int) atomicExch(&h[i], 0u);  // signed: a pending correction of the fast drain may leave -1 behind
    if (v) atomicAdd(&g[i], (unsigned long long) (long long) v);
  }
}

// ---------------------------------------------------------------------------------------------
// Per-lane stacks of accepted pairs, addressed with 32-bit shared-window addresses.
// Every shared-memory access written as inline PTX in this file is `asm volatile`: to the compiler a plain asm is a pure
// function of its operands, and a load whose address register holds the same value in every chunk could legally be
// merged with an earlier one across the stores that refill the buffer.  Volatile asms keep their order among
// themselves and against __syncwarp() (itself a volatile asm with a memory clobber, which also orders the ordinary
// C++ stores of the staging code), and that is all the ordering these buffers need; a "memory" clobber on the loads
// would in addition make the compiler re-read kernel parameters from the constant bank inside the pair loops.
// Layout: [slot][lane][NW words]; a warp-wide push or pop touches 32 consecutive entries (no bank conflicts).
template <class T, int NW> struct QOps;
// push: predicated store + predicated pointer bump in one asm block (a stack: no wrap-around arithmetic)
#define FCFC_PUSH_ASM(ST) "{.reg .pred q; setp.ne.s32 q, %1, 0; " ST " @q add.u32 %0, %0, %2;}"
template <int NW> struct QOps<float, NW> {
  static __device__ __forceinline__ unsigned lane_offset(int lane) { return (unsigned) lane * (unsigned) (NW * sizeof(float)); }
  static __device__ __forceinline__ void push(unsigned &w, const float (&v)[NW], bool p) {
    constexpr unsigned S = 32u * NW * sizeof(float);
    if (NW == 1) asm volatile(FCFC_PUSH_ASM("@q st.shared.f32 [%0], %3;") : "+r"(w) : "r"((int) p), "n"(S), "f"(v[0]));
    else if (NW == 2) asm volatile(FCFC_PUSH_ASM("@q st.shared.v2.f32 [%0], {%3, %4};") : "+r"(w) : "r"((int) p), "n"(S), "f"(v[0]), "f"(v[1 % NW]));
    else asm volatile(FCFC_PUSH_ASM("@q st.shared.v4.f32 [%0], {%3, %4, %5, %6};") : "+r"(w) : "r"((int) p), "n"(S), "f"(v[0]), "f"(v[1 % NW]), "f"(v[2 % NW]), "f"(v[3 % NW]));
  }
  static __device__ __forceinline__ void load(unsigned a, float (&v)[NW]) {
    if (NW == 1) FCFC_LDS_ASM("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(a));
    else if (NW == 2) FCFC_LDS_ASM("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1 % NW]) : "r"(a));
    else FCFC_LDS_ASM("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1 % NW]), "=f"(v[2 % NW]), "=f"(v[3 % NW]) : "r"(a));
  }
};
template <int NW> struct QOps<double, NW> {
  static __device__ __forceinline__ unsigned lane_offset(int lane) { return (unsigned) lane * (unsigned) (NW * sizeof(double)); }
  static __device__ __forceinline__ void push(unsigned &w, const double (&v)[NW], bool p) {
    constexpr unsigned S = 32u * NW * sizeof(double);
    if (NW == 1) asm volatile(FCFC_PUSH_ASM("@q st.shared.f64 [%0], %3;") : "+r"(w) : "r"((int) p), "n"(S), "d"(v[0]));
    else if (NW == 2) asm volatile(FCFC_PUSH_ASM("@q st.shared.v2.f64 [%0], {%3, %4};") : "+r"(w) : "r"((int) p), "n"(S), "d"(v[0]), "d"(v[1 % NW]));
    else asm volatile(FCFC_PUSH_ASM("@q st.shared.v2.f64 [%0], {%3, %4}; @q st.shared.v2.f64 [%0+16], {%5, %6};") : "+r"(w) : "r"((int) p), "n"(S), "d"(v[0]), "d"(v[1 % NW]), "d"(v[2 % NW]), "d"(v[3 % NW]));
  }
  static __device__ __forceinline__ void load(unsigned a, double (&v)[NW]) {
    if (NW == 1) FCFC_LDS_ASM("ld.shared.f64 %0, [%1];" : "=d"(v[0]) : "r"(a));
    else {
      FCFC_LDS_ASM("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1 % NW]) : "r"(a));
      if (NW == 4) FCFC_LDS_ASM("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(v[2 % NW]), "=d"(v[3 % NW]) : "r"(a));
    }
  }
};
#undef FCFC_PUSH_ASM

// Per-lane LIFO of accepted pairs (the order in which pairs reach the histogram is irrelevant).
// The stacks are columns of a [slot][column] array, and a column is not tied to a lane: the packed pair loop hands every
// column to the next lane after each step (one SHFL of `top`), so that every column collects the accepted pairs of all
// lanes in turn.  A lane's acceptance rate follows the position of its primaries in the tile and stays high or low for a
// whole stencil row; without the rotation the fullest stack fills at the highest rate, the average one at the average rate,
// and a drain -- every lane pops as many entries as the fullest stack must lose -- runs with idle lanes in that
// proportion (measured: 55 % of the pop slots useful on the shell of partially accepted points; a simulation of the
// rotation gives 85-90 %).  Entries are self-contained, so whoever holds a column when the warp drains pops it.
#ifndef FCFC_ROTATE
#define FCFC_ROTATE 1
#endif
template <class T, int NW> struct LaneQueue {
  unsigned int top;             // shared address of the next free slot of the column this lane holds
  unsigned int base;            // shared address of slot 0 of that column (valid outside the rotating pair loop)
  unsigned int wbase;           // shared address of the warp's queue area (slot 0 of column 0)
  static constexpr unsigned int kStride = 32u * NW * sizeof(T);
  __device__ __forceinline__ void push(const T (&v)[NW], bool ok) { QOps<T, NW>::push(top, v, ok); }
  __device__ __forceinline__ unsigned int fill_bytes() const { return top - base; }
  __device__ __forceinline__ void rotate(int next_lane) { top = __shfl_sync(0xffffffffu, top, next_lane); }
  __device__ __forceinline__ void rebase() { base = wbase + ((top - wbase) & (kStride - 1u)); }
};

// Truncation of 0 <= x < 2^23 (float) / 2^31 (double) without the quarter-rate F2I: add 2^23 (2^52)
// rounding toward zero and read the low mantissa bits.
__device__ __forceinline__ int trunc_pos(float x) { return __float_as_int(__fadd_rz(x, 8388608.0f)) - 0x4B000000; }
__device__ __forceinline__ int trunc_pos(double x) { return __double2loint(__dadd_rz(x, 4503599627370496.0)); }
// floor(sqrt(m)) for an integer 0 <= m < 2^18: sqrt(m + 1/2) is at least 0.25/sqrt(m) away from every
// integer, far more than the error of the approximate square root.
__device__ __forceinline__ int isqrt_small(int m) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float) m + 0.5f));
  return trunc_pos(r);
}
// Quotient rounded toward zero for a >= 0, b > 0: the residual of the round-to-nearest quotient is exact.
__device__ __forceinline__ float div_rz_pos(float a, float b) {
  float q = __fdiv_rn(a, b);
  return (__fmaf_rn(-q, b, a) < 0.0f) ? __int_as_float(__float_as_int(q) - 1) : q;
}
__device__ __forceinline__ double div_rz_pos(double a, double b) {
  double q = __ddiv_rn(a, b);
  return (__fma_rn(-q, b, a) < 0.0) ? __longlong_as_double(__double_as_longlong(q) - 1) : q;
}

// Histogram bin of one queued pair, -1 if it is dropped after all.  The fast variant (GENERIC = false:
// zero lower bounds, 8-bit integer tables in shared memory) is branch-free for box / isotropic counts;
// everything else goes through finish_pair.
template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, int NW>
__device__ __forceinline__ int bin_entry(const CountParams<T> &P, const BlockCtx<T> &C, const T (&e)[NW], T &w) {
  using A = Ar<T>;
  w = (T) 1;
  if (!GENERIC && (BOX || BIN == BIN_ISO)) {
    const T d2 = e[0];
    if (WT) w = e[(BIN == BIN_ISO) ? 1 % NW : 2 % NW];
    const int si = trunc_pos(d2);
    int sb;
    if (P.stab_is_sqrt) sb = isqrt_small(si); else sb = C.stab[si];
    int pb = 0;
    bool ok = true;
    if (BIN == BIN_SMU) {
      const T dz2 = A::mul(e[1 % NW], e[1 % NW]);    // the queue holds dz
      int m;
      if (ARITH == ARITH_SCALAR) {              // metric_common.c:184
        m = trunc_pos(A::mul(A::div(dz2, d2), P.nmu2f));
        m = (d2 < A::eps()) ? 0 : m;
      } else {                                  // metric_common.c:472-494 (AVX-512 branch)
        const T q = div_rz_pos(A::mul(dz2, P.nmu2f), d2);
        m = (q < P.nmu2f) ? trunc_pos(q) : P.nmu2;
        m = (d2 >= A::eps()) ? m : 0;
      }
      if (m >= P.nmu2) { ok = P.with_mu_one != 0; m = P.nmu2 - 1; }
      if (P.mu_is_sqrt) pb = isqrt_small(m); else pb = C.mutab[m];
    } else if (BIN == BIN_SPI) {
      const int pi_i = trunc_pos(e[1 % NW]);
      if (P.ptab_is_ident) pb = pi_i; else pb = C.ptab[pi_i];
    }
    return ok ? sb + pb * P.ns : -1;
  } else {
    T d2, aux = 0, as = 0, bs = 0;
    if (BIN == BIN_ISO) { d2 = e[0]; if (WT) w = e[1 % NW]; }
    else if (BOX) { d2 = e[0]; aux = e[1 % NW]; if (WT) w = e[2 % NW]; }
    else {      // survey: rebuild s^2 = (s1 + s2) - t with the same two operations as eval_pair
      aux = e[0]; as = e[1 % NW]; bs = e[2 % NW]; if (WT) w = e[3 % NW];
      d2 = A::sub(A::add(as, bs), aux);
    }
    return finish_pair<T, BIN, BOX, ARITH, GENERIC>(P, C, d2, aux, as, bs);
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-window (32-bit address) helpers for the drain: no generic-address arithmetic in the hot loop.
__device__ __forceinline__ void red_shared_u32(unsigned a, bool p) {
  asm volatile("{.reg .pred q; setp.ne.s32 q, %0, 0; @q red.shared.add.u32 [%1], 1;}" ::"r"((int) p), "r"(a));
}
__device__ __forceinline__ void red_shared_f64(unsigned a, double v, bool p) {
  asm volatile("{.reg .pred q; setp.ne.s32 q, %0, 0; @q red.shared.add.f64 [%1], %2;}" ::"r"((int) p), "r"(a), "d"(v));
}
__device__ __forceinline__ int lds_u8(unsigned a, bool p) {
  unsigned v = 0;
  asm volatile("{.reg .pred q; setp.ne.s32 q, %1, 0; @q ld.shared.u8 %0, [%2];}" : "+r"(v) : "r"((int) p), "r"(a));
  return (int) v;
}
struct FastCtx { unsigned hist_s, stab_s, ptab_s, mutab_s; unsigned hstride, hlane; };   // weighted: bytes per bin, byte offset of this lane's copy

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(double x) { return __double2float_rn(x); }

// Fast bins of a box / isotropic pair from ONE approximate reciprocal square root r = rsqrt(d2):
//   s        = d2 * r            -> s bin  floor(s)   (floor(sqrt(floor(d2))) == floor(sqrt(d2)) exactly)
//   nmu * mu = nmu * |dz| * r    -> mu bin floor(.)   (== floor(sqrt(floor(fl(fl(dz2/d2)*nmu^2)))) away from bin edges)
// Both are computed scaled by 2^ks / 2^km (chosen on the host from ns / nmu) and truncated by a fused multiply-add
// of 2^23 + 1 rounded toward zero: the mantissa then holds floor(2^k x) + 1 = bin * 2^k + k fraction bits.
// Error budget: rsqrt.approx 2^-22 rel. + two roundings move s by < ns * 3e-7; the same plus the reference's own
// three roundings move nmu*mu by < nmu * 4.5e-7; the host picks 2^-ks >= 2.5 * ns * 3e-7 and 2^-km >= 2.2 * nmu * 4.5e-7.
// A pair must be re-binned with the exact IEEE sequence when
//   s  : fraction bits in {2^ks - 1, 0, 1, 2}  -> s within [-1, 3) / 2^ks of an edge; also catches d2 < EPS (s ~ 0)
//   mu : fraction bits in {2^km - 1, 0}        -> nmu*mu within [-1, 1) / 2^km of an integer; also catches mu >= 1,
//        since nmu*mu cannot exceed nmu by more than the error budget.
// `t` (output) is zero for such a pair: the product of the masked fraction bits (it may also wrap to zero, which
// only costs a needless exact evaluation).  Integer work is kept off the half-rate ALU pipe where possible
// (IMAD.HI shifts, IMAD product).  Returns the bin biased by bias_s + bias_m * ns, bias = 0x4B000000 >> k; the
// caller folds the bias into the base address.  The biased bin of ANY pair that passed the range test lies in
// [bias, bias + ntot + ns]: s <= ns (1 + 3e-7), nmu*mu <= nmu (1 + 4.5e-7).
template <int BIN, bool CLAMP = false>
__device__ __forceinline__ int fast_bins(float d2, float dz, const float sscale, const float mscale_nmu,
                                         const unsigned int smask, const unsigned int mmask,
                                         const unsigned int sshift_mul, const unsigned int mshift_mul, int ns, unsigned int &t) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d2 + 1e-30f));      // d2 = 0 (coincident points): finite r, s = 0 -> flagged
  // one fused multiply-add rounded toward zero: the mantissa is exactly floor(2^ks s) + 1
  float sr, mr = 0.0f;
  if (BIN == BIN_SMU) {         // (d2, dz) * (r, r) as one packed multiply (sm_100a FMUL2: half the issue slots)
    unsigned long long e2, r2, p2;
    asm("mov.b64 %0, {%1, %2};" : "=l"(e2) : "f"(d2), "f"(dz));
    asm("mov.b64 %0, {%1, %1};" : "=l"(r2) : "f"(r));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p2) : "l"(e2), "l"(r2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(sr), "=f"(mr) : "l"(p2));
  } else sr = d2 * r;
  if (CLAMP) mr = fminf(fabsf(mr), 1.000002f);  // survey: mu from separately rounded terms is not bounded by construction
  const unsigned int us = (unsigned int) __float_as_int(__fmaf_rz(sr, sscale, 8388609.0f));
  t = us & smask;
  int bin = (int) __umulhi(us, sshift_mul);     // us >> ks
  if (BIN == BIN_SMU) {
    const unsigned int um = (unsigned int) __float_as_int(__fmaf_rz(fabsf(mr), mscale_nmu, 8388609.0f));
    t *= (um & mmask);
    bin += (int) __umulhi(um, mshift_mul) * ns; // (um >> km) * ns
  }
  return bin;
}

__device__ __forceinline__ void red_shared_u32_add(unsigned a, unsigned v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v));
}

// (d2, aux) of a queued pair as the floats fast_bins works on.  Box / isotropic entries hold them directly.
// Survey (s,mu) entries hold (t, s1, s2[, w]) (2pt/metric_common.c:169-172): d2 = (s1 + s2) - t with the same two
// operations as eval_pair, and the line-of-sight separation pi = |s1 - s2| / sqrt(s1 + s2 + t) plays the role of dz
// (mu = pi / s).  Two more roundings and one more rsqrt.approx than the box form: the host widens the mu band.
template <class T, int BIN, bool BOX, int NW>
__device__ __forceinline__ void fast_inputs(const T (&e)[NW], float &d2f, float &auxf) {
  if (BOX || BIN == BIN_ISO) { d2f = to_f32(e[0]); auxf = (BIN == BIN_SMU) ? to_f32(e[1 % NW]) : 0.0f; }
  else {
    using A = Ar<T>;
    const T s = A::add(e[1 % NW], e[2 % NW]);
    d2f = fmaxf(to_f32(A::sub(s, e[0])), 0.0f);         // (s - t can round below zero for coincident points)
    const float df = to_f32(A::sub(e[1 % NW], e[2 % NW])), stf = to_f32(A::add(s, e[0]));
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(stf));
    auxf = df * r;
  }
}
template <int BIN, bool BOX, int NW> struct QWeight { static constexpr int kIndex = (BIN == BIN_ISO) ? 1 % NW : (BOX ? 2 % NW : 3 % NW); };

// Exact re-binning of one flagged pair, out of line (a few pairs per ten thousand): keeps its registers and code
// out of the drain loop.  Tables are read from global memory (P.stab / P.mutab), which is fine at this frequency.
// Unweighted counts were already added to the fast bin by the drain loop: that increment is taken back here
// (the 32-bit shared counters are signed, see sweep_hist).
template <class T, int BIN, bool BOX, bool WT, int ARITH, int NW>
__device__ __noinline__ void fix_entry(const CountParams<T> &P, unsigned int hist_s, unsigned int hstride, unsigned int hlane,
                                       unsigned int qaddr) {
  BlockCtx<T> G;
  G.hist_u = nullptr; G.hist_d = nullptr; G.blk_evals = nullptr;
  G.stab = P.stab; G.ptab = P.ptab; G.mutab = P.mutab; G.s2bin = P.s2bin; G.pbin = P.pbin;
  T e[NW], w;
  QOps<T, NW>::load(qaddr, e);
  const int b = bin_entry<T, BIN, BOX, WT, ARITH, false, NW>(P, G, e, w);
  if (WT) { if (b >= 0) red_shared_f64(hist_s + hlane + hstride * (unsigned int) b, (double) w, true); }
  else {
    const int bias = (int) (0x4B000000u >> P.fb_sshift) + ((BIN == BIN_SMU) ? (int) (0x4B000000u >> P.fb_mshift) * P.ns : 0);
    unsigned int t;
    float d2f, auxf;
    fast_inputs<T, BIN, BOX, NW>(e, d2f, auxf);
    const int fb = fast_bins<BIN, !(BOX || BIN == BIN_ISO)>(d2f, auxf, P.fb_sscale, P.fb_mscale, P.fb_smask,
                                  P.fb_mmask, 1u << (32 - P.fb_sshift), 1u << (32 - P.fb_mshift), P.ns, t) - bias;
    if (b != fb) {
      red_shared_u32_add(hist_s + 4u * (unsigned int) fb, 0xffffffffu);
      if (b >= 0) red_shared_u32_add(hist_s + 4u * (unsigned int) b, 1u);
    }
  }
}

// Drain of the fast variants whose bins are computed (box (s,mu) or isotropic counts, integer tables that are
// floor(sqrt(i)), zero lower bounds, shared-memory histogram): every lane pops `rounds` entries off its stack,
// four per iteration, with no data-dependent branch.  Unweighted: every popped pair increments its fast bin
// unconditionally (FULL: all lanes still have entries; otherwise lanes that have run dry aim at a private dump
// slot); weighted: predicated.  One bit per pair records whether it was clean (carry chain:
// clean = 2 * clean + (t != 0), two instructions); the rare other pairs are re-binned exactly after the loop.
template <class T, int BIN, bool BOX, bool WT, int NW, bool FULL>
__device__ __forceinline__ void drain_fast_loop(const CountParams<T> &P, unsigned int hist_adj, const unsigned int HS, const unsigned int dump,
                                                unsigned int &rp, int k0, int k1, int mine, unsigned int &clean) {
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
  constexpr int NE = 4;
  const float sscale = P.fb_sscale, mscale = P.fb_mscale;
  const unsigned int smask = P.fb_smask, mmask = P.fb_mmask, smul = 1u << (32 - P.fb_sshift), mmul = 1u << (32 - P.fb_mshift);
#pragma unroll 1
  for (int k = k0; k < k1; k += NE, rp += NE * S) {
    T e[NE][NW];
#pragma unroll
    for (int i = 0; i < NE; i++) QOps<T, NW>::load(rp + i * S, e[i]);   // slots above the old top hold stale entries: ignored
#pragma unroll
    for (int i = 0; i < NE; i++) {
      unsigned int t;
      float d2f, auxf;
      fast_inputs<T, BIN, BOX, NW>(e[i], d2f, auxf);
      const int bin = fast_bins<BIN, !(BOX || BIN == BIN_ISO)>(d2f, auxf, sscale, mscale, smask, mmask, smul, mmul, P.ns, t);
      const bool h = FULL || (k + i < mine);
      unsigned int addr = hist_adj + HS * (unsigned int) bin;
      if (WT) red_shared_f64(addr, (double) e[i][QWeight<BIN, BOX, NW>::kIndex], h && t != 0u);
      else {
        if (!FULL) addr = h ? addr : dump;
        red_shared_u32_add(addr, 1u);
      }
      asm("{.reg .u32 tmp; add.cc.u32 tmp, %1, 0xffffffff; addc.u32 %0, %0, %0;}" : "+r"(clean) : "r"(t));
    }
  }
}

template <class T, int BIN, bool BOX, bool WT, int ARITH, int NW>
__device__ __forceinline__ void drain_fast(const CountParams<T> &P, const BlockCtx<T> &C, const FastCtx &F,
                                           LaneQueue<T, NW> &Q, int rounds) {
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
  const int mine = min((int) (Q.fill_bytes() / S), rounds);     // entries this lane pops (rounds <= 32, a multiple of 4)
  Q.top -= (unsigned int) mine * S;
  unsigned int rp = Q.top;
  const int bias = (int) (0x4B000000u >> P.fb_sshift) + ((BIN == BIN_SMU) ? (int) (0x4B000000u >> P.fb_mshift) * P.ns : 0);
  const unsigned int hs = WT ? F.hstride : 4u;
  const unsigned int hist_adj = F.hist_s + (WT ? F.hlane : 0u) - hs * (unsigned int) bias;
  const unsigned int dump = F.hist_s + 4u * (unsigned int) (P.ntot + P.ns + 1) + 4u * (threadIdx.x & 31u);
  // rounds in which every lane still has four entries (one ragged loop for everything measured 1.5 % slower)
  const int nfull = __reduce_min_sync(0xffffffffu, mine) & ~3;
  unsigned int clean = 0;       // one bit per round, most recent round in bit 0
  drain_fast_loop<T, BIN, BOX, WT, NW, true>(P, hist_adj, hs, dump, rp, 0, nfull, mine, clean);
  drain_fast_loop<T, BIN, BOX, WT, NW, false>(P, hist_adj, hs, dump, rp, nfull, rounds, mine, clean);
  unsigned int flagged = ~clean & (0xffffffffu >> (32 - rounds));
  while (flagged) {                                             // rare
    const int k = rounds - __ffs((int) flagged);
    flagged &= flagged - 1;
    if (k < mine)                                               // (stale entries past this lane's stack may have raised a flag)
      fix_entry<T, BIN, BOX, WT, ARITH, NW>(P, F.hist_s, F.hstride, F.hlane, Q.top + (unsigned int) k * S);
  }
  __syncwarp();
}

// Drain of the fast variants that look their bins up: box (s_perp, pi), or integer tables that are not
// floor(sqrt(i)).  Exact by construction (truncation + table), predicated shared-memory loads.
template <class T, int BIN, bool BOX, bool WT, int ARITH, int NW>
__device__ __forceinline__ void drain_lut(const CountParams<T> &P, const BlockCtx<T> &C, const FastCtx &F,
                                          LaneQueue<T, NW> &Q, int rounds) {
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
  const int mine = min((int) (Q.fill_bytes() / S), rounds);
  Q.top -= (unsigned int) mine * S;
  unsigned int rp = Q.top;
#pragma unroll 1
  for (int k = 0; k < rounds; k += 2, rp += 2 * S) {
    T e[2][NW];
    QOps<T, NW>::load(rp, e[0]);
    QOps<T, NW>::load(rp + S, e[1]);
#pragma unroll
    for (int i = 0; i < 2; i++) {
      bool h = k + i < mine;
      int bin;
      if (BIN == BIN_SPI) {
        int sb, pb = trunc_pos(e[i][1 % NW]);
        const int fl = trunc_pos(e[i][0]);
        if (P.stab_is_sqrt) sb = isqrt_small(fl); else sb = lds_u8(F.stab_s + (unsigned) fl, h);
        if (!P.ptab_is_ident) pb = lds_u8(F.ptab_s + (unsigned) pb, h);
        bin = sb + pb * P.ns;
      } else {
        T ww;
        bin = h ? bin_entry<T, BIN, BOX, WT, ARITH, false, NW>(P, C, e[i], ww) : -1;
        h = bin >= 0;
      }
      if (WT) red_shared_f64(F.hist_s + F.hlane + F.hstride * (unsigned) bin, (double) e[i][(BIN == BIN_ISO) ? 1 % NW : 2 % NW], h);
      else red_shared_u32(F.hist_s + 4u * (unsigned) bin, h);
    }
  }
}

// Drain of every other variant: exact per-entry binning through finish_pair.
template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, bool SMEMHIST, int NW>
__device__ __forceinline__ void drain_generic(const CountParams<T> &P, const BlockCtx<T> &C, LaneQueue<T, NW> &Q, int rounds) {
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
#pragma unroll 1
  for (int k = 0; k < rounds; k++) {
    if (Q.top != Q.base) {
      T e[NW];
      Q.top -= S;
      QOps<T, NW>::load(Q.top, e);
      T w;
      const int b = bin_entry<T, BIN, BOX, WT, ARITH, GENERIC, NW>(P, C, e, w);
      if (b >= 0) hist_add<T, WT, SMEMHIST>(C, b, w);
    }
  }
}

// Pop and bin entries.  Called when the fullest queue may overflow: if it really is close to full, every lane
// pops until the fullest queue is down to `keep` entries; lanes that run dry simply idle.  Returns the new
// upper bound of the fullest queue (entries).
template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, bool SMEMHIST, int NW>
__device__ __forceinline__ int drain_queue(const CountParams<T> &P, const BlockCtx<T> &C, const FastCtx &F,
                                           LaneQueue<T, NW> &Q, int need, int keep) {
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
  const int mx = (int) (__reduce_max_sync(0xffffffffu, Q.fill_bytes()) / S);
  if (mx + need <= P.qdepth - 1) return mx;
  const int rounds = min((mx - keep + 3) & ~3, 32);     // a multiple of 4 (the fast drain pops four entries per iteration)
  if (rounds <= 0) return 0;
  __syncwarp();                 // the columns rotate between the lanes: pushes of other lanes must be visible to whoever pops
  if (!GENERIC && SMEMHIST && (BOX || BIN == BIN_ISO)) {
    if (BIN != BIN_SPI && P.stab_is_sqrt && (BIN != BIN_SMU || P.mu_is_sqrt)) drain_fast<T, BIN, BOX, WT, ARITH, NW>(P, C, F, Q, rounds);
    else drain_lut<T, BIN, BOX, WT, ARITH, NW>(P, C, F, Q, rounds);
  } else if (!GENERIC && SMEMHIST && BIN == BIN_SMU && P.stab_is_sqrt && P.mu_is_sqrt) {
    drain_fast<T, BIN, BOX, WT, ARITH, NW>(P, C, F, Q, rounds);       // survey (s,mu): computed bins too (fast_inputs)
  } else drain_generic<T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST, NW>(P, C, Q, rounds);
  __syncwarp();                 // ... and the pops complete before another lane pushes into the column
  return max(mx - rounds, 0);
}

__device__ __forceinline__ void lds_vec4_raw(unsigned a, float &x, float &y, float &z, float &w) {
  FCFC_LDS_ASM("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
}
__device__ __forceinline__ void lds_vec4_raw(unsigned a, double &x, double &y, double &z, double &w) {
  FCFC_LDS_ASM("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
  FCFC_LDS_ASM("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(z), "=d"(w) : "r"(a));
}
template <class T> __device__ __forceinline__ Vec4<T> lds_vec4(unsigned a) { Vec4<T> v; lds_vec4_raw(a, v.x, v.y, v.z, v.s); return v; }

// ---------------------------------------------------------------------------------------------
// Packed FP32 pairs (sm_100a add/mul/fma.rn.f32x2: the lane throughput of the scalar forms at half the issue
// slots; a scalar operand is broadcast for free).  Every operation is a separate IEEE round-to-nearest
// instruction, exactly like its scalar counterpart: results are bit-identical.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// Product that must stay a separately rounded product (scalar-parity order): ptxas contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 (observed in the SASS; a literal -0 addend is folded away first), unlike the scalar forms.
// x * y + (-0) with a -0 that ptxas cannot see through (read from shared memory at kernel start) is the correctly
// rounded product and leaves nothing to contract.
__device__ __forceinline__ f32x2 mul2_uncontracted(f32x2 a, f32x2 b, float negzero) { return fma2(a, b, pk2(negzero, negzero)); }

// Squared separations (and the second queue word) of one primary against the two secondary points of a staged pair,
// on packed f32x2 arithmetic: bit-identical to the scalar sequences of eval_pair.
template <int BIN, bool BOX, int ARITH, int NW>
__device__ __forceinline__ void packed_dist(const float axr, const float ayr, const float azr, const f32x2 X, const f32x2 Y, const f32x2 Z,
                                            const float (&zz)[2], const float negzero, float (&d2h)[2], float (&auxh)[2]) {
  const f32x2 dx = sub2(pk2(axr, axr), X), dy = sub2(pk2(ayr, ayr), Y);
  if (NW == 1) {                            // isotropic: everything packed, only d2 is kept
    const f32x2 dz = sub2(pk2(azr, azr), Z);
    const f32x2 dz2 = (ARITH == ARITH_SCALAR) ? mul2_uncontracted(dz, dz, negzero) : mul2(dz, dz);
    f32x2 d2;
    if (ARITH == ARITH_SCALAR) d2 = add2(add2(mul2_uncontracted(dx, dx, negzero), mul2_uncontracted(dy, dy, negzero)), dz2);      // :170-172
    else if (BOX) d2 = fma2(dy, dy, fma2(dx, dx, dz2));                               // :426-430
    else d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));                               // 2pt/:330-333
    upk2(d2, d2h[0], d2h[1]);
    auxh[0] = auxh[1] = 0.0f;
  } else {
    // two-word entries (d2, aux) are pushed with one 64-bit store each: the last operation of the d2 chain and
    // the z difference are scalar, so that d2_h and aux_h can be produced side by side in registers
    float dzh[2], dyh[2];
    dzh[0] = __fsub_rn(azr, zz[0]); dzh[1] = __fsub_rn(azr, zz[1]);
    upk2(dy, dyh[0], dyh[1]);
    if (BIN == BIN_SPI) {                   // box (s_perp, pi): metric_common.c:157-165, 416-424
      float mxh[2], myh[2];
      if (ARITH == ARITH_SCALAR) {
        upk2(mul2_uncontracted(dx, dx, negzero), mxh[0], mxh[1]); upk2(mul2_uncontracted(dy, dy, negzero), myh[0], myh[1]);
      } else upk2(mul2(dx, dx), mxh[0], mxh[1]);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        d2h[h] = (ARITH == ARITH_SCALAR) ? __fadd_rn(mxh[h], myh[h]) : __fmaf_rn(dyh[h], dyh[h], mxh[h]);
        auxh[h] = fabsf(dzh[h]);
      }
    } else {                                // box (s, mu): :170-172 / :426-430
      const float dz2h[2] = {__fmul_rn(dzh[0], dzh[0]), __fmul_rn(dzh[1], dzh[1])};
      float uh[2];
      if (ARITH == ARITH_SCALAR) upk2(add2(mul2_uncontracted(dx, dx, negzero), mul2_uncontracted(dy, dy, negzero)), uh[0], uh[1]);
      else upk2(fma2(dx, dx, pk2(dz2h[0], dz2h[1])), uh[0], uh[1]);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        d2h[h] = (ARITH == ARITH_SCALAR) ? __fadd_rn(uh[h], dz2h[h]) : __fmaf_rn(dyh[h], dyh[h], uh[h]);
        auxh[h] = dzh[h];
      }
    }
  }
}

// Variants whose pair loop runs packed: float, isotropic bins, unweighted (one word per queue entry).  Their staging
// buffer holds the 32 secondary points as 16 pairs, [pair][x0 x1 y0 y1 z0 z1 - -] (32 bytes per pair).  With two-word
// entries ((s,mu), (s_perp,pi)) the packed results (d2_0, d2_1), (dz_0, dz_1) would have to be re-paired for the
// 64-bit push: measured slower than the scalar loop (474 vs 460 ms on the bench workload; two 32-bit pushes: 467 ms).
template <class T, int BIN, bool BOX, bool WT> struct PairLoop { static constexpr bool kPacked = sizeof(T) == 4 && (BOX || BIN == BIN_ISO) && !WT; };
__device__ __forceinline__ unsigned staged_pair_addr(unsigned sbuf_s, int j) { return sbuf_s + (unsigned) (j >> 1) * 32u + (unsigned) (j & 1) * 4u; }

// Tile point held by `lane` as its r-th primary (other assignments, e.g. reversed in odd r, measured no better).
__device__ __forceinline__ int tile_slot(int r, int lane) { return r * 32 + lane; }

// "Dense" cells: every point of the secondary cell is within the maximum separation of every point of the tile's cell
// (the host marks them per stencil row: CountParams::rows_in), so every pair is accepted -- a third of the accepted
// pairs of the bench workload.  Such chunks skip the stacks: the pair loop computes the fast bins in place, with all
// lanes busy, and increments the histogram directly (no push, no pop: 4 of the 7.5 shared-memory wavefronts an accepted
// pair costs otherwise).  Pairs within the error band of a bin edge (fast_bins: t == 0) are pushed on the stack
// instead, and the ordinary drain re-bins them exactly.  The range test is kept (one compare): exactness never depends
// on the host's classification of the cells.  Only for the variants whose drain computes its bins (drain_fast).
template <class T, int BIN, bool BOX, int ARITH, int R, int NW, int RMAX>
__device__ __forceinline__ int do_chunk_dense(const CountParams<T> &P, LaneQueue<T, NW> &Q, int &ub, const unsigned int hist_s, const int lane,
                                              const Vec4<T> *sbuf, int j0, int nj,
                                              const T (&ax)[RMAX], const T (&ay)[RMAX], const T (&az)[RMAX],
                                              const T s2lim, const float negzero) {
  static_assert(sizeof(T) == 4, "the dense path runs on the packed float pair loop");
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
  const unsigned int sbuf_s = (unsigned int) __cvta_generic_to_shared(sbuf);
  const unsigned int lim = Q.base + (unsigned int) (P.qdepth - 1 - 2 * R) * S;        // room for 2 R flagged pairs
  unsigned int sa = sbuf_s + (unsigned int) (j0 >> 1) * 32u;
  const unsigned int sa0 = sa, se = sbuf_s + (unsigned int) ((nj + 1) >> 1) * 32u;
  const float sscale = P.fb_sscale, mscale = P.fb_mscale;
  const unsigned int smask = P.fb_smask, mmask = P.fb_mmask, smul = P.fb_smul, mmul = P.fb_mmul;
  unsigned int hist_adj = hist_s - 4u * P.fb_bias;                        // base address of the biased fast bins (as in drain_fast)
  unsigned int dump = hist_s + 4u * (unsigned int) (P.ntot + P.ns + 1) + 4u * (unsigned int) lane;
  // opaque to the compiler from here on: the two addresses stay in registers instead of being rebuilt from the lane index
  // and the constant bank on every step of the loop (7 of its 178 instructions; C2 356.9 -> 352.4 ms)
  asm volatile("" : "+r"(hist_adj), "+r"(dump));
  ub = P.qdepth;
#pragma unroll 1
  for (; sa != se; sa += 32u) {
    if (__any_sync(0xffffffffu, Q.top > lim)) break;
    f32x2 X, Y, Z;
    FCFC_LDS_ASM("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(X), "=l"(Y) : "r"(sa));
    FCFC_LDS_ASM("ld.shared.b64 %0, [%1+16];" : "=l"(Z) : "r"(sa));
    float zz[2];
    upk2(Z, zz[0], zz[1]);
#pragma unroll
    for (int r = 0; r < R; r++) {
      float d2h[2], auxh[2];
      packed_dist<BIN, BOX, ARITH, NW>(ax[r], ay[r], az[r], X, Y, Z, zz, negzero, d2h, auxh);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        unsigned int t, addr;
        const int bin = fast_bins<BIN, false>(d2h[h], auxh[h], sscale, mscale, smask, mmask, smul, mmul, P.ns, t);
        // in range (padding lanes and parked points are not; their bins are never used) and clean: counted here;
        // in range and flagged: pushed.  ATOMS.POPC.INC cannot be predicated without a branch: everything that is
        // not counted aims at the lane's dump slot.  One predicate chain, six instructions.
        if (NW == 1)
          asm volatile("{.reg .pred pk, pc, pf; setp.lt.f32 pk, %2, %3; setp.ne.and.u32 pc, %4, 0, pk; setp.eq.and.u32 pf, %4, 0, pk; "
                       "selp.u32 %0, %5, %6, pc; @pf st.shared.f32 [%1], %2; @pf add.u32 %1, %1, %7;}"
                       : "=r"(addr), "+r"(Q.top) : "f"(d2h[h]), "f"(s2lim), "r"(t), "r"(hist_adj + 4u * (unsigned int) bin), "r"(dump), "n"(S));
        else
          asm volatile("{.reg .pred pk, pc, pf; setp.lt.f32 pk, %2, %3; setp.ne.and.u32 pc, %4, 0, pk; setp.eq.and.u32 pf, %4, 0, pk; "
                       "selp.u32 %0, %5, %6, pc; @pf st.shared.v2.f32 [%1], {%2, %8}; @pf add.u32 %1, %1, %7;}"
                       : "=r"(addr), "+r"(Q.top) : "f"(d2h[h]), "f"(s2lim), "r"(t), "r"(hist_adj + 4u * (unsigned int) bin), "r"(dump), "n"(S), "f"(auxh[h]));
        red_shared_u32_add(addr, 1u);
      }
    }
  }
  return min(j0 + (int) ((sa - sa0) >> 4), nj);
}

// One chunk of <= 32 staged secondary points against the R register-resident primaries of each lane,
// starting at staged point j0.  Returns nj when the chunk is done, or the index of the point at which
// the queues must be drained first (the caller drains at its single call site and resumes).
template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, int R, bool SELF, int NW, int RMAX>
__device__ __forceinline__ int do_chunk(const CountParams<T> &P, LaneQueue<T, NW> &Q, int &ub,
                                        const Vec4<T> *sbuf, const T *wbuf, int j0, int nj,
                                        const T (&ax)[RMAX], const T (&ay)[RMAX], const T (&az)[RMAX], const T (&as)[RMAX],
                                        const T (&aw)[RMAX], int jglob0, int iglob0, int lane, const T s2lim, const float negzero) {
  constexpr bool kPacked = PairLoop<T, BIN, BOX, WT>::kPacked;
  const unsigned int sbuf_s = (unsigned int) __cvta_generic_to_shared(sbuf);
  if constexpr (kPacked && !SELF) {
    // two secondary points per step.  Before every step one vote checks that each lane has room for 2 R more entries:
    // the loop leaves only when a stack is really about to overflow (a worst-case bound on the pushes would
    // leave every two or three steps at a 50 % acceptance rate).
    constexpr unsigned int S = LaneQueue<T, NW>::kStride;
    // proceed while fill + 2 R <= qdepth - 1, i.e. while the slot index of `top` is below qdepth - 2 R (the column offset
    // is less than one slot row, so one limit serves every column)
    const unsigned int lim = Q.wbase + (unsigned int) (P.qdepth - 2 * R) * S;
    const int next_lane = (lane + 1) & 31;
    unsigned int sa = sbuf_s + (unsigned int) (j0 >> 1) * 32u;     // j0 is even: this path always advances by pairs
    const unsigned int sa0 = sa, se = sbuf_s + (unsigned int) ((nj + 1) >> 1) * 32u;
    ub = P.qdepth;                // (no bound is tracked here: whoever needs one next measures the stacks first)
#pragma unroll kEvalUnroll
    for (; sa != se; sa += 32u) {
      if (__any_sync(0xffffffffu, Q.top >= lim)) break;
      f32x2 X, Y, Z;
      FCFC_LDS_ASM("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(X), "=l"(Y) : "r"(sa));
      FCFC_LDS_ASM("ld.shared.b64 %0, [%1+16];" : "=l"(Z) : "r"(sa));
      float zz[2];
      upk2(Z, zz[0], zz[1]);
#pragma unroll
      for (int r = 0; r < R; r++) {
        float d2h[2], auxh[2];
        packed_dist<BIN, BOX, ARITH, NW>(ax[r], ay[r], az[r], X, Y, Z, zz, negzero, d2h, auxh);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          bool ok = d2h[h] < s2lim;
          if (BIN == BIN_SPI) ok = ok && (auxh[h] < P.pmax);
          if (GENERIC) ok = ok && (P.smin0 || d2h[h] >= P.s2min);
          if (GENERIC && BIN == BIN_SPI) ok = ok && (P.pmin0 || auxh[h] >= P.pmin);
          T e[NW];
          e[0] = d2h[h];
          if (NW > 1) e[1 % NW] = auxh[h];
          Q.push(e, ok);
        }
      }
      if (FCFC_ROTATE) Q.rotate(next_lane);
    }
    if (FCFC_ROTATE) Q.rebase();
    return min(j0 + (int) ((sa - sa0) >> 4), nj);
  }
  // one secondary point per step; as in the packed loop a vote before every step checks the real fill of the stacks
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
  const unsigned int lim = Q.wbase + (unsigned int) (P.qdepth - R) * S;         // proceed while fill + R <= qdepth - 1 (slot index of `top` below qdepth - R)
  const int next_lane = (lane + 1) & 31;
  ub = P.qdepth;
  // the staged points are walked with a 32-bit shared-window address (one add per point, no index arithmetic)
  unsigned int sa = (unsigned int) __cvta_generic_to_shared(sbuf + j0);
  const unsigned int se = (unsigned int) __cvta_generic_to_shared(sbuf + nj);
  int j = j0;
#pragma unroll kEvalUnroll
  for (; sa != se; sa += (unsigned int) sizeof(Vec4<T>)) {
    if (__any_sync(0xffffffffu, Q.top >= lim)) break;
    Vec4<T> b;
    if constexpr (kPacked) {      // (only the tile against its own points gets here: SELF) pair layout, one point
      const unsigned int pa = staged_pair_addr(sbuf_s, j);
      FCFC_LDS_ASM("ld.shared.f32 %0, [%1];" : "=f"(b.x) : "r"(pa));
      FCFC_LDS_ASM("ld.shared.f32 %0, [%1+8];" : "=f"(b.y) : "r"(pa));
      FCFC_LDS_ASM("ld.shared.f32 %0, [%1+16];" : "=f"(b.z) : "r"(pa));
      b.s = 0;
    } else b = lds_vec4<T>(sa);
    T bw = (T) 1;
    if (WT) bw = wbuf[j];
#pragma unroll
    for (int r = 0; r < R; r++) {
      T d2, aux;
      bool ok = eval_pair<T, BIN, BOX, ARITH, GENERIC>(P, ax[r], ay[r], az[r], as[r], b, s2lim, d2, aux);
      if (SELF) ok = ok && (jglob0 + j > iglob0 + tile_slot(r, lane));    // unordered pairs once: metric_common.c:2017-2018
      T e[NW];
      if (BIN == BIN_ISO) { e[0] = d2; if (WT) e[1 % NW] = Ar<T>::mul(aw[r], bw); }
      else if (BOX) { e[0] = d2; e[1 % NW] = aux; if (WT) { e[2 % NW] = Ar<T>::mul(aw[r], bw); e[3 % NW] = 0; } }
      else { e[0] = aux; e[1 % NW] = as[r]; e[2 % NW] = b.s; e[3 % NW] = WT ? Ar<T>::mul(aw[r], bw) : (T) 0; }
      Q.push(e, ok);
    }
    if (FCFC_ROTATE) Q.rotate(next_lane);       // the stack columns move on to the next lane (see LaneQueue)
    j++;
  }
  if (FCFC_ROTATE) Q.rebase();
  return j;
}

template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, bool SMEMHIST, int RMAX>
__global__ void __launch_bounds__(BlockShape<T>::kThreads, 1) count_kernel(const __grid_constant__ CountParams<T> P) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kThreads = BlockShape<T>::kThreads;
  constexpr int NW = QFmt<BIN, BOX, WT>::NW;
  constexpr bool kPacked = PairLoop<T, BIN, BOX, WT>::kPacked;
  constexpr bool kDense = FCFC_DENSE && kPacked && !GENERIC && SMEMHIST && BIN != BIN_SPI && RMAX >= 2;     // see do_chunk_dense
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nmutab = (BIN == BIN_SMU) ? P.nmu2 : 0;
  const SmemPlan pl = make_smem_plan<T, WT>(P.ntot, P.nstab * (P.swidth ? 2 : 1), P.nptab * (P.pwidth ? 2 : 1),
                                            nmutab, P.ns, P.np, P.nrows, SMEMHIST, NW, P.qdepth, P.tabs_global != 0, WT ? P.hist_copies : 1);
  BlockCtx<T> C;
  C.hist_u = SMEMHIST ? reinterpret_cast<unsigned int *>(smem + pl.off_hist) : reinterpret_cast<unsigned int *>(P.ghist_i);
  C.hist_d = SMEMHIST ? reinterpret_cast<double *>(smem + pl.off_hist) : P.ghist_d;
  const int hcopies = (WT && SMEMHIST) ? P.hist_copies : 1;
  C.hmul = hcopies; C.hoff = lane & (hcopies - 1);       // (hcopies: a power of two)
  uint8_t *s_stab = smem + pl.off_stab, *s_ptab = smem + pl.off_ptab, *s_mutab = smem + pl.off_mutab;
  T *s_s2bin = reinterpret_cast<T *>(smem + pl.off_s2bin), *s_pbin = reinterpret_cast<T *>(smem + pl.off_pbin);
  int4 *s_rows = reinterpret_cast<int4 *>(smem + pl.off_rows);
  C.stab = s_stab; C.ptab = s_ptab; C.mutab = s_mutab; C.s2bin = s_s2bin; C.pbin = s_pbin;
  if (P.tabs_global) { C.stab = P.stab; C.ptab = P.ptab; C.mutab = P.mutab; }     // generic variant, or fast variant whose tables are computed
  unsigned int *s_blk_evals = reinterpret_cast<unsigned int *>(smem + pl.off_misc);
  C.blk_evals = s_blk_evals;

  // ---- block prologue: zero the histogram, stage tables / edges / stencil rows ----
  if (SMEMHIST) {
    if (WT) for (int i = threadIdx.x; i < P.ntot * hcopies; i += kThreads) C.hist_d[i] = 0.0;
    else for (int i = threadIdx.x; i < P.ntot + P.ns + 33; i += kThreads) C.hist_u[i] = 0u;
  }
  if (!P.tabs_global) {
    for (int i = threadIdx.x; i < P.nstab * (P.swidth ? 2 : 1); i += kThreads) s_stab[i] = P.stab[i];
    if (BIN == BIN_SPI) for (int i = threadIdx.x; i < P.nptab * (P.pwidth ? 2 : 1); i += kThreads) s_ptab[i] = P.ptab[i];
    if (BIN == BIN_SMU) for (int i = threadIdx.x; i < nmutab; i += kThreads) s_mutab[i] = P.mutab[i];
  }
  for (int i = threadIdx.x; i <= P.ns; i += kThreads) s_s2bin[i] = P.s2bin[i];
  if (BIN == BIN_SPI) for (int i = threadIdx.x; i <= P.np; i += kThreads) s_pbin[i] = P.pbin[i];
  for (int i = threadIdx.x; i < P.nrows; i += kThreads) s_rows[i] = P.rows[i];
  if (threadIdx.x == 0) {
    *s_blk_evals = 0;
    *reinterpret_cast<T *>(smem + pl.off_misc + 8) = (BIN == BIN_SPI && !BOX) ? P.premax : P.s2max;
    *reinterpret_cast<float *>(smem + pl.off_misc + 4) = -0.0f;
  }
  // queues start zeroed: slots past a queue's tail are read (and ignored) by the two-entry drain
  for (int i = threadIdx.x * 16; i < pl.total - pl.off_queue; i += kThreads * 16)    // (including the over-read pad behind the last stack)
    *reinterpret_cast<uint4 *>(smem + pl.off_queue + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  FastCtx F;
  F.hist_s = (unsigned int) __cvta_generic_to_shared(smem + pl.off_hist);
  F.hstride = 8u * (unsigned int) hcopies; F.hlane = 8u * (unsigned int) (lane & (hcopies - 1));
  F.stab_s = (unsigned int) __cvta_generic_to_shared(s_stab);
  F.ptab_s = (unsigned int) __cvta_generic_to_shared(s_ptab);
  F.mutab_s = (unsigned int) __cvta_generic_to_shared(s_mutab);

  Vec4<T> *sbuf = reinterpret_cast<Vec4<T> *>(smem + pl.off_stage + warp * pl.stage_per_warp);
  T *wbuf = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(sbuf) + 32 * sizeof(Vec4<T>));
  LaneQueue<T, NW> Q;
  Q.wbase = (unsigned int) __cvta_generic_to_shared(smem + pl.off_queue) + warp * (unsigned int) pl.queue_per_warp;
  Q.base = Q.top = Q.wbase + QOps<T, NW>::lane_offset(lane);
  int ub = 0;                   // warp-uniform upper bound of the fullest lane queue (entries)
  // upper limit of the range test, parked in a vector register (read back from shared memory, which the compiler
  // cannot fold into a constant-bank operand: it would otherwise be re-fetched with LDCU for every secondary point)
  const T s2lim = *reinterpret_cast<volatile T *>(smem + pl.off_misc + 8);
  const float negzero = kPacked ? *reinterpret_cast<volatile float *>(smem + pl.off_misc + 4) : 0.0f;     // see mul2_uncontracted
  unsigned long long my_evals = 0;
  const int ncy = P.nc[1], ncz = P.nc[2];

  // ---- persistent warp loop over work items ----
  while (true) {
    int item = 0;
    if (lane == 0) {
      const long long w = (long long) P.part + (long long) P.nparts * (long long) atomicAdd(P.work_counter, 1u);
      item = (w < (long long) P.nitem) ? P.item_order[w] : -1;
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item < 0) break;
    const int tile_id = item / P.nsplit, split = item - tile_id * P.nsplit;
    const int cell = P.item_cell[tile_id], t0 = P.item_off[tile_id], cnt = P.item_cnt[tile_id];
    const int iz = cell % ncz, iy = (cell / ncz) % ncy, ix = cell / (ncz * ncy);
    const int nr = (cnt + 31) >> 5;             // primaries per lane actually used by this tile (1..RMAX)

    // primaries: lane holds points t0 + r*32 + lane; padding lanes sit far away (never in range)
    T px[RMAX], py[RMAX], pz[RMAX], ps[RMAX], pw[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; r++) {
      const int k = tile_slot(r, lane);
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
    auto sweep_range = [&](int b, int e, T sax, T say, T saz, T sbx, T sby, T sbz, bool self, bool dense) {
      T ax[RMAX], ay[RMAX], az[RMAX];
#pragma unroll
      for (int r = 0; r < RMAX; r++) {
        constexpr bool kDot = !BOX && BIN != BIN_ISO;     // survey (s,mu) / (s_perp,pi): dot-product form, see eval_pair
        ax[r] = BOX ? Ar<T>::add(px[r], sax) : (kDot ? Ar<T>::add(px[r], px[r]) : px[r]);
        ay[r] = BOX ? Ar<T>::add(py[r], say) : (kDot ? Ar<T>::add(py[r], py[r]) : py[r]);
        az[r] = BOX ? Ar<T>::add(pz[r], saz) : (kDot ? Ar<T>::add(pz[r], pz[r]) : pz[r]);
      }
      while (b < e) {
        const int piece_end = min(e, b + kSegPieceMax);
        // overflow accounting of the 32-bit shared counters (see sweep_hist)
        if (SMEMHIST && !WT) {
          unsigned int add = (unsigned int) (piece_end - b) * (unsigned int) cnt, old = 0;
          if (lane == 0) old = atomicAdd(C.blk_evals, add);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old + add >= 0x40000000u || old + add < old) {
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
          if constexpr (kPacked) {        // pairs of points; lanes past the end of the range park a point that is never in range
            const bool live = c0 + lane < piece_end;
            const unsigned int pa = staged_pair_addr((unsigned int) __cvta_generic_to_shared(sbuf), lane);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(pa), "f"(live ? nxt.x : -Ar<T>::far()));
            asm volatile("st.shared.f32 [%0+8], %1;" ::"r"(pa), "f"(live ? nxt.y : -Ar<T>::far()));
            asm volatile("st.shared.f32 [%0+16], %1;" ::"r"(pa), "f"(live ? nxt.z : -Ar<T>::far()));
          } else sbuf[lane] = nxt;
          if (WT) wbuf[lane] = nxtw;
          __syncwarp();
          jn = c0 + 32 + lane;
          if (jn < piece_end) { nxt = P.pos2[jn]; if (WT) nxtw = P.w2[jn]; }
          const int nj = min(32, piece_end - c0);
          const bool sf = self && c0 < t0 + cnt;
#define FCFC_CHUNK_DENSE(RR) do_chunk_dense<T, BIN, BOX, ARITH, RR, NW, RMAX>(P, Q, ub, F.hist_s, lane, sbuf, j, nj, ax, ay, az, s2lim, negzero)
#define FCFC_CHUNK(RR, SF) do_chunk<T, BIN, BOX, WT, ARITH, GENERIC, RR, SF, NW, RMAX>(P, Q, ub, sbuf, wbuf, j, nj, ax, ay, az, ps, pw, c0, t0, lane, s2lim, negzero)
          for (int j = 0;;) {
            if (sf) j = FCFC_CHUNK(RMAX, true);           // rare: the tile against its own points
            else if (kDense && dense && nr >= RMAX - 1) {  // every pair in range: binned in place (full or nearly full tiles)
              if constexpr (kDense) {
                if (nr == RMAX) j = FCFC_CHUNK_DENSE(RMAX); else j = FCFC_CHUNK_DENSE(RMAX - 1);
              }
            } else if (RMAX == 4) {
              switch (nr) {                               // partially filled tiles evaluate only the primaries they hold
                case 1: j = FCFC_CHUNK(1, false); break;
                case 2: j = FCFC_CHUNK(2, false); break;
                case 3: j = FCFC_CHUNK(3, false); break;
                default: j = FCFC_CHUNK(4, false); break;
              }
            } else j = (nr == 1) ? FCFC_CHUNK(1, false) : FCFC_CHUNK(RMAX, false);
            if (j >= nj) break;
            // a queue may overflow: the one place where queued pairs are binned
            ub = drain_queue<T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST, NW>(P, C, F, Q, kPacked ? 2 * RMAX : RMAX, P.qkeep);
          }
#undef FCFC_CHUNK
#undef FCFC_CHUNK_DENSE
        }
        b = piece_end;
      }
    };

    // Sweep list: q = -1 is the tile's own cell (auto counts: pairs i < j only, metric_common.c:2017);
    // q = 3*row + image enumerates, for every stencil row, the three periodic images of its z run
    // (below the box: the primaries get +L in z; inside; above the box: the secondaries get +L).
    // A single call site keeps the unrolled pair loops in the instruction cache.
    const int nq = (P.periodic ? 3 : 1) * P.nrows, qfirst = P.isauto ? -1 : 0;
    const int qlo = qfirst + (int) ((long long) (nq - qfirst) * split / P.nsplit);
    const int qhi = qfirst + (int) ((long long) (nq - qfirst) * (split + 1) / P.nsplit);
    // With dense cells (do_chunk_dense) the z run of a sweep is cut into three pieces -- before, inside and after the
    // row's dense cells -- enumerated by the same loop (qq = 3 q + piece), so that nothing but the loop counter stays
    // live across a sweep: registers are what the pair loops are short of.
    constexpr int kPieces = kDense ? 3 : 1;
    for (int qq = kPieces * qlo; qq < kPieces * qhi; qq++) {
      const int q = kDense ? (qq >= 0 ? qq / 3 : -1) : qq, piece = kDense ? qq - 3 * q : 0;
      int b, e;
      T sax = 0, say = 0, saz = 0, sbx = 0, sby = 0, sbz = 0;
      if (q < 0) {
        if (piece != 0) continue;
        b = t0; e = P.cell_start2[cell + 1];
      } else {
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
        if (kDense) {
          // the dense cells of the row [dlo, dhi], through the same image mapping; none: dlo = zhi + 1 (piece 0 is everything)
          int dlo = zhi + 1, dhi = zhi;
          if (P.rows_in != nullptr) {
            const int2 in = __ldg(&P.rows_in[ri]);
            int lo = iz + in.x, hi = iz + in.y;
            if (P.periodic) {
              if (img == 0) { hi = min(hi, -1) + ncz; lo += ncz; }
              else if (img == 2) { lo = max(lo, ncz) - ncz; hi -= ncz; }
            }
            lo = max(lo, zlo); hi = min(hi, zhi);
            if (lo <= hi) { dlo = lo; dhi = hi; }
          }
          if (piece == 0) zhi = dlo - 1;
          else if (piece == 1) { zlo = dlo; zhi = dhi; }
          else zlo = dhi + 1;
          if (zlo > zhi) continue;
        }
        const int rowbase = (jx * ncy + jy) * ncz;
        b = P.cell_start2[rowbase + zlo]; e = P.cell_start2[rowbase + zhi + 1];
      }
      if (b >= e) continue;
      sweep_range(b, e, sax, say, saz, sbx, sby, sbz, q < 0, piece == 1);       // (a single call site keeps the pair loops in the instruction cache)
    }
  }
  // whatever is still queued
  while (drain_queue<T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST, NW>(P, C, F, Q, P.qdepth, 0) > 0) {}

  // ---- block epilogue: flush the histogram ----
  __syncthreads();
  if (SMEMHIST) {
    if (WT) {
      for (int i = threadIdx.x; i < P.ntot; i += kThreads) {
        double v = 0.0;
        for (int c = 0; c < hcopies; c++) v += C.hist_d[i * hcopies + c];
        if (v != 0.0) atomicAdd(&P.ghist_d[i], v);
      }
    } else {
      for (int i = threadIdx.x; i < P.ntot; i += kThreads) { const int v = (int) C.hist_u[i]; if (v) atomicAdd(&P.ghist_i[i], (unsigned long long) (long long) v); }
    }
  }
  // pair-evaluation counter: one atomic per warp
  if (lane == 0 && my_evals) atomicAdd(P.gevals, my_evals);
}

}  // namespace fcfc
