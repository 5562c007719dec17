// fcfc_b200/csrc/count_kernel_pf.cuh -- double-precision pair counting through a single-precision pre-filter (sm_100a).
//
// The reference's default build is double (src/util/define_comm.h:68-89).  Evaluated directly, every candidate pair
// costs 6-14 FP64 instructions on a pipe with half the FP32 rate, although only the pairs that end up in the histogram
// need double precision.  This kernel splits the work:
//
//   filter (FP32, packed f32x2, every candidate)   float copies of the staged secondaries and of the tile's primaries,
//       d^2 from plain differences, compared against limits PADDED by the worst-case float error (the host derives the
//       padding from the extent of the data, engine.cu: fcfc_gpu_prefilter_limits): a superset of the pairs the exact
//       tests can accept.  A candidate is recorded as ONE BIT: every lane keeps, per primary, a 32-bit mask over the 32
//       secondaries of the staged chunk, in a register.  No stores, no votes, no early exits in the filter loop;
//   exact pass (FP64, candidates only)              after each chunk, per primary slot r: the warp ORs its masks, walks the
//       secondaries any lane selected (broadcast read of the secondary's double4 from the staging buffer) and the
//       selected lanes evaluate the pair with eval_pair / bin_entry of count_kernel.cuh -- the same IEEE sequences as the
//       plain double kernel, which restate the reference's (metric_common.c:140-235, 377-534; 2pt/metric_common.c:142-259,
//       283-460).  The double values decide: results are bit-identical to the plain kernel.  The 32 primaries of a slot
//       are consecutive points of the Morton order inside a cell, so the lanes mostly select the same secondaries.
//
// Survey (s_perp, pi) counts accept a thin cylinder inside the searched sphere; their filter adds float versions of the
// two division-free cylinder tests of eval_pair, again with padded limits.
//
// Work distribution, cell lists, stencil rows, histogram handling: as in count_kernel.cuh (same CountParams).
#pragma once
#include "count_kernel.cuh"
#include <type_traits>

namespace fcfc {

#ifndef FCFC_PF_WARPS
#define FCFC_PF_WARPS 24
#endif
#ifndef FCFC_PF_WARPS_SPI
#define FCFC_PF_WARPS_SPI 28
#endif
// Warps per block (one block per SM).  Measured on survey counts (profiles/dense_block_experiments_r2.log): 16 / 20 / 24 / 28
// warps give 26.6 / 24.0 / 21.9 / 20.7 ms for (s_perp,pi) -- its exact pass is a chain of dependent FP64 operations and a
// division, bound by latency -- and 503.6 / 457.5 / 432.5 / 436.4 ms for (s,mu).
__host__ __device__ constexpr int pf_warps(int bintype, bool box) { return (bintype == BIN_SPI && !box) ? FCFC_PF_WARPS_SPI : FCFC_PF_WARPS; }

struct PfPlan { int off_hist, off_stab, off_ptab, off_mutab, off_s2bin, off_pbin, off_rows, off_misc, off_warp, per_warp, o_stage_d, o_wbuf, o_stage_f, o_box, o_clist, total; };

template <bool WT>
__host__ __device__ inline PfPlan make_pf_plan(int ntot, int nstab_bytes, int nptab_bytes, int nmutab_bytes, int ns, int np, int nrows,
                                               bool smem_hist, bool tabs_global, int hist_copies, int warps) {
  PfPlan p;
  int o = 0;
  auto al = [](int v) { return (v + 15) & ~15; };
  p.off_hist = o; o += smem_hist ? al(WT ? ntot * 8 * hist_copies : (ntot + ns + 1 + 32) * 4) : 0;
  p.off_stab = o; o += tabs_global ? 0 : al(nstab_bytes);
  p.off_ptab = o; o += tabs_global ? 0 : al(nptab_bytes);
  p.off_mutab = o; o += tabs_global ? 0 : al(nmutab_bytes);
  p.off_s2bin = o; o += al((ns + 1) * 8);
  p.off_pbin = o; o += al((np + 1) * 8);
  p.off_rows = o; o += al(nrows * 16);
  p.off_misc = o; o += 16;
  p.off_warp = o;
  int w = 0;
  // ring of 64 staged secondaries (two blocks of 32; survey counts compact the points that survive the classification
  // against the tile into it, box counts stage chunk by chunk into the block at its head)
  p.o_stage_d = w; w += 64 * 32;                // 64 x double4
  p.o_wbuf = w; w += WT ? 64 * 8 : 0;
  p.o_stage_f = w; w += 1024;                   // float copies: 32 pairs x (x0 x1 y0 y1) | 32 pairs x (z0 z1 s0 s1)
  p.o_box = w; w += 48;                         // the tile for the classification: (cx, cy, cz, R) (hx, hy, hz, -) (s_min, s_max, -, -)
  p.o_clist = w; w += 1024 * 2;                 // candidate codes of one primary slot: (lane << 5) | secondary, at most 32 x 32
  p.per_warp = w;
  p.total = o + warps * w;
  return p;
}

// Exact binning of one pair whose fast bins were flagged (a few per ten thousand), out of line.
template <int BIN, bool BOX, bool WT, int ARITH, int NW>
__device__ __noinline__ void pf_fix_pair(const CountParams<double> &P, unsigned int hist_s, unsigned int hstride, unsigned int hlane,
                                         double e0, double e1, double e2, double e3) {
  BlockCtx<double> G;
  G.hist_u = nullptr; G.hist_d = nullptr; G.blk_evals = nullptr;
  G.stab = P.stab; G.ptab = P.ptab; G.mutab = P.mutab; G.s2bin = P.s2bin; G.pbin = P.pbin;
  double e[NW], w;
  e[0] = e0; if (NW > 1) e[1 % NW] = e1; if (NW > 2) { e[2 % NW] = e2; e[3 % NW] = e3; }
  const int b = bin_entry<double, BIN, BOX, WT, ARITH, false, NW>(P, G, e, w);
  if (b < 0) return;
  if (WT) red_shared_f64(hist_s + hlane + hstride * (unsigned int) b, w, true);
  else red_shared_u32_add(hist_s + 4u * (unsigned int) b, 1u);
}

template <int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, bool SMEMHIST, int RMAX>
__global__ void __launch_bounds__(pf_warps(BIN, BOX) * 32, 1) count_kernel_pf(const __grid_constant__ CountParams<double> P) {
  constexpr int kPfWarps = pf_warps(BIN, BOX), kPfThreads = kPfWarps * 32;
  using T = double;
  using A = Ar<double>;
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NW = QFmt<BIN, BOX, WT>::NW;
  constexpr bool kDot = !BOX && BIN != BIN_ISO;         // survey (s,mu) / (s_perp,pi): dot-product form of eval_pair
  constexpr bool kCyl = !BOX && BIN == BIN_SPI;         // cylinder tests in the filter
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nmutab = (BIN == BIN_SMU) ? P.nmu2 : 0;
  const int hcopies = (WT && SMEMHIST) ? P.hist_copies : 1;
  const PfPlan pl = make_pf_plan<WT>(P.ntot, P.nstab * (P.swidth ? 2 : 1), P.nptab * (P.pwidth ? 2 : 1), nmutab, P.ns, P.np, P.nrows,
                                     SMEMHIST, P.tabs_global != 0, hcopies, kPfWarps);
  BlockCtx<T> C;
  C.hist_u = SMEMHIST ? reinterpret_cast<unsigned int *>(smem + pl.off_hist) : reinterpret_cast<unsigned int *>(P.ghist_i);
  C.hist_d = SMEMHIST ? reinterpret_cast<double *>(smem + pl.off_hist) : P.ghist_d;
  C.hmul = hcopies; C.hoff = lane & (hcopies - 1);       // (hcopies: a power of two)
  uint8_t *s_stab = smem + pl.off_stab, *s_ptab = smem + pl.off_ptab, *s_mutab = smem + pl.off_mutab;
  T *s_s2bin = reinterpret_cast<T *>(smem + pl.off_s2bin), *s_pbin = reinterpret_cast<T *>(smem + pl.off_pbin);
  int4 *s_rows = reinterpret_cast<int4 *>(smem + pl.off_rows);
  C.stab = s_stab; C.ptab = s_ptab; C.mutab = s_mutab; C.s2bin = s_s2bin; C.pbin = s_pbin;
  if (P.tabs_global) { C.stab = P.stab; C.ptab = P.ptab; C.mutab = P.mutab; }
  unsigned int *s_blk_evals = reinterpret_cast<unsigned int *>(smem + pl.off_misc);
  C.blk_evals = s_blk_evals;

  // ---- block prologue ----
  if (SMEMHIST) {
    if (WT) for (int i = threadIdx.x; i < P.ntot * hcopies; i += kPfThreads) C.hist_d[i] = 0.0;
    else for (int i = threadIdx.x; i < P.ntot + P.ns + 33; i += kPfThreads) C.hist_u[i] = 0u;
  }
  if (!P.tabs_global) {
    for (int i = threadIdx.x; i < P.nstab * (P.swidth ? 2 : 1); i += kPfThreads) s_stab[i] = P.stab[i];
    if (BIN == BIN_SPI) for (int i = threadIdx.x; i < P.nptab * (P.pwidth ? 2 : 1); i += kPfThreads) s_ptab[i] = P.ptab[i];
    if (BIN == BIN_SMU) for (int i = threadIdx.x; i < nmutab; i += kPfThreads) s_mutab[i] = P.mutab[i];
  }
  for (int i = threadIdx.x; i <= P.ns; i += kPfThreads) s_s2bin[i] = P.s2bin[i];
  if (BIN == BIN_SPI) for (int i = threadIdx.x; i <= P.np; i += kPfThreads) s_pbin[i] = P.pbin[i];
  for (int i = threadIdx.x; i < P.nrows; i += kPfThreads) s_rows[i] = P.rows[i];
  if (threadIdx.x == 0) *s_blk_evals = 0;
  for (int i = threadIdx.x * 16; i < kPfWarps * pl.per_warp; i += kPfThreads * 16)      // staging starts zeroed
    *reinterpret_cast<uint4 *>(smem + pl.off_warp + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();

  unsigned char *wbase = smem + pl.off_warp + warp * pl.per_warp;
  Vec4<T> *stage_d = reinterpret_cast<Vec4<T> *>(wbase + pl.o_stage_d);
  T *wbuf = reinterpret_cast<T *>(wbase + pl.o_wbuf);
  unsigned short *clist = reinterpret_cast<unsigned short *>(wbase + pl.o_clist);
  float *s_box = reinterpret_cast<float *>(wbase + pl.o_box);
  const unsigned int box_s = (unsigned int) __cvta_generic_to_shared(s_box);
  const unsigned int lt_mask = (1u << lane) - 1u;
  constexpr bool kClassify = !BOX;              // survey counts: no image shifts, the ring lives across the stencil rows
  const unsigned int stage_d_s = (unsigned int) __cvta_generic_to_shared(stage_d);
  const unsigned int stage_f_s = (unsigned int) __cvta_generic_to_shared(wbase + pl.o_stage_f);
  const unsigned int hist_s = (unsigned int) __cvta_generic_to_shared(smem + pl.off_hist);
  const unsigned int hstride = 8u * (unsigned int) hcopies, hlane = 8u * (unsigned int) (lane & (hcopies - 1));
  const unsigned int dump = hist_s + 4u * (unsigned int) (P.ntot + P.ns + 1) + 4u * (unsigned int) lane;
  // binning path of the exact pass (warp-uniform): computed bins + exact re-binning of flagged pairs, or bin_entry
  const bool fast = !GENERIC && SMEMHIST && BIN != BIN_SPI && P.stab_is_sqrt && (BIN != BIN_SMU || P.mu_is_sqrt);
  const T s2lim = (BIN == BIN_SPI && !BOX) ? P.premax : P.s2max;
  const float f_d2lim = P.pf_d2lim, f_plim = P.pf_plim, f_s2lim = P.pf_s2lim;
  const float f_s2cl = (float) P.s2max * 1.001f, f_p2cl = (float) P.pmax * 1.001f;      // classification: the exact limits, widened
  const bool cyl_on = kCyl && f_plim > 0.0f;            // (the host switches the cylinder tests off when their padding would be large)
  unsigned long long my_evals = 0, my_made = 0;         // candidates of the swept cell ranges; filter evaluations made
#ifdef FCFC_PF_STATS
  unsigned long long dbg_steps = 0, dbg_useful = 0;
#endif
  const int ncy = P.nc[1], ncz = P.nc[2];

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
    const int nr = (cnt + 31) >> 5;
    int ring_cnt = 0, ring_head = 0;            // pending points [head, head + cnt) mod 64, head is 0 or 32
    if (kClassify) {
      // bounding box (centre, half-widths, half-diagonal) and range of |x|^2 of the tile's points, widened for the float roundings
      float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f}, smn = 3e38f, smx = 0.0f;
      for (int k = lane; k < cnt; k += 32) {
        const Vec4<T> v = P.pos1[t0 + k];
        const float x = __double2float_rn(v.x), y = __double2float_rn(v.y), z = __double2float_rn(v.z), q = __double2float_rn(v.s);
        lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x); lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y); lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
        smn = fminf(smn, q); smx = fmaxf(smx, q);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; d++) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
        smn = fminf(smn, __shfl_xor_sync(0xffffffffu, smn, o)); smx = fmaxf(smx, __shfl_xor_sync(0xffffffffu, smx, o));
      }
      __syncwarp();
      if (lane == 0) {
        const float hx = 0.5f * (hi[0] - lo[0]), hy = 0.5f * (hi[1] - lo[1]), hz = 0.5f * (hi[2] - lo[2]);
        s_box[0] = 0.5f * (lo[0] + hi[0]); s_box[1] = 0.5f * (lo[1] + hi[1]); s_box[2] = 0.5f * (lo[2] + hi[2]);
        const float cmax = fmaxf(fmaxf(fmaxf(fabsf(lo[0]), fabsf(hi[0])), fmaxf(fabsf(lo[1]), fabsf(hi[1]))), fmaxf(fabsf(lo[2]), fabsf(hi[2])));
        s_box[3] = sqrtf(hx * hx + hy * hy + hz * hz) * 1.001f + 1e-6f * cmax;        // half-diagonal + a few ulps of the largest coordinate
        s_box[4] = hx; s_box[5] = hy; s_box[6] = hz;
        s_box[8] = smn * 0.999999f; s_box[9] = smx * 1.000001f;
      }
      __syncwarp();
    }

    auto sweep_range = [&](int b, int e, T sax, T say, T saz, T sbx, T sby, T sbz, bool self, bool flush) {
      // float copies of the (shifted) primaries: the filter's side of the tile
      float fx[RMAX], fy[RMAX], fz[RMAX], fs[RMAX];
#pragma unroll
      for (int r = 0; r < RMAX; r++) {
        const int k = r * 32 + lane;
        if (k < cnt) {
          const Vec4<T> v = P.pos1[t0 + k];
          fx[r] = __double2float_rn(BOX ? A::add(v.x, sax) : v.x);
          fy[r] = __double2float_rn(BOX ? A::add(v.y, say) : v.y);
          fz[r] = __double2float_rn(BOX ? A::add(v.z, saz) : v.z);
          fs[r] = kCyl ? __double2float_rn(v.s) : 0.0f;
        } else { fx[r] = fy[r] = fz[r] = 3e18f; fs[r] = 0.0f; }       // padding lanes: never within the limits
      }
      while (b < e || flush) {
        const int piece_end = min(e, b + kSegPieceMax);
        if (SMEMHIST && !WT && b < e) {         // overflow accounting of the 32-bit shared counters (count_kernel.cuh)
          unsigned int add = (unsigned int) (piece_end - b) * (unsigned int) cnt, old = 0;
          if (lane == 0) old = atomicAdd(C.blk_evals, add);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old + add >= 0x40000000u || old + add < old) {
            if (lane == 0) atomicExch(C.blk_evals, 0u);
            sweep_hist(C.hist_u, P.ghist_i, P.ntot, lane);
          }
        }
        if (b < e) {
          unsigned long long ev = (unsigned long long) (piece_end - b) * (unsigned long long) cnt;
          if (self && b == t0) { ev -= (unsigned long long) cnt * (unsigned long long) (cnt + 1) / 2; my_made -= (unsigned long long) cnt * (unsigned long long) (cnt + 1) / 2; }
          my_evals += ev;
        }
        Vec4<T> nxt; T nxtw = 0;
        nxt.x = nxt.y = nxt.z = nxt.s = 0;
        int jn = b + lane;
        if (jn < piece_end) { nxt = P.pos2[jn]; if (WT) nxtw = P.w2[jn]; }
        const bool last_piece = piece_end >= e;
        for (int c0 = b; c0 < piece_end || (flush && last_piece);) {
          const bool fl = c0 >= piece_end;      // closing pass of a flushed range: nothing to stage, everything pending is processed
          __syncwarp();
          int thr = 32;
          bool sf = false;
          if (fl) {
            thr = 1;
            if ((ring_cnt & 1) && lane == 0) {  // an odd fill gets a parked partner (never within the limits)
              const unsigned int q = (unsigned int) ((ring_head + ring_cnt) & 63);
              const unsigned int pa = stage_f_s + (q >> 1) * 16u + (q & 1u) * 4u;
              asm volatile("st.shared.f32 [%0], %1; st.shared.f32 [%0+8], %1; st.shared.f32 [%0+512], %1; st.shared.f32 [%0+520], %2;" ::"r"(pa), "f"(-3e18f), "f"(0.0f));
            }
          } else {
            // ---- stage the chunk: exact doubles (shift applied) and their float copies in pair layout ----
            if (BOX) { nxt.x = A::add(nxt.x, sbx); nxt.y = A::add(nxt.y, sby); nxt.z = A::add(nxt.z, sbz); }
            const bool live = c0 + lane < piece_end;
            const float bx = __double2float_rn(nxt.x), by = __double2float_rn(nxt.y), bz = __double2float_rn(nxt.z), bs = kCyl ? __double2float_rn(nxt.s) : 0.0f;
            bool keep = true;
            unsigned int q = (unsigned int) (ring_head + lane);
            if (kClassify && !self) {
              // Classification against the tile (a superset test: a dropped point has no partner the exact tests accept).
              //   sphere:   nearest point of the tile's box beyond the padded range limit of the filter;
              //   cylinder (s_perp, pi), with s_perp = 2 |a x b| / |a + b| and pi = | |a|^2 - |b|^2 | / |a + b|:
              //     |a x b| >= |b| (rho - R), rho = distance of the box centre from the line of sight of b, R = half-diagonal,
              //     |a + b|^2 <= 2 (|a|^2 + |b|^2)  =>  s_perp^2 >= 2 |b|^2 (rho - R)^2 / (max|a|^2 + |b|^2)
              //                                         pi^2     >= (gap of |b|^2 to the tile's range of |a|^2)^2 / (2 (max|a|^2 + |b|^2))
              //   each compared with its limit widened by 1e-3 (the float roundings of these tests are below 1e-4).
              float cx_, cy_, cz_, cr_, hx_, hy_, hz_, smn_, smx_, unused, unused2;
              lds_vec4_raw(box_s, cx_, cy_, cz_, cr_);
              lds_vec4_raw(box_s + 16u, hx_, hy_, hz_, unused);
              const float vx = bx - cx_, vy = by - cy_, vz = bz - cz_;
              const float ux = fmaxf(fabsf(vx) - hx_, 0.0f), uy = fmaxf(fabsf(vy) - hy_, 0.0f), uz = fmaxf(fabsf(vz) - hz_, 0.0f);
              keep = !(__fmaf_rn(uz, uz, __fmaf_rn(uy, uy, ux * ux)) > f_d2lim * 1.001f);
              if (kCyl) {
                lds_vec4_raw(box_s + 32u, smn_, smx_, unused, unused2);
                const float kx = cy_ * bz - cz_ * by, ky = cz_ * bx - cx_ * bz, kz = cx_ * by - cy_ * bx;       // c x b
                const float cross2 = __fmaf_rn(kz, kz, __fmaf_rn(ky, ky, kx * kx));
                const float ssum = smx_ + bs;
                const float tt = sqrtf(f_s2cl * ssum / (2.0f * bs)) * 1.001f + cr_;            // T + R
                keep = keep && !(cross2 > tt * tt * bs * 1.001f);
                const float gap = fmaxf(fmaxf(bs - smx_, smn_ - bs), 0.0f);
                keep = keep && !(gap * gap > f_p2cl * 2.0f * ssum);
              }
              keep = keep && live;
              const unsigned int mk = __ballot_sync(0xffffffffu, keep);
              q = (unsigned int) ((ring_head + ring_cnt + __popc(mk & lt_mask)) & 63);
              ring_cnt += __popc(mk);
            } else {
              // box counts (the primaries' image shift belongs to the row) and the tile's own cell (pairs i < j by position):
              // chunk by chunk into the block at the head of the (empty) ring
              ring_cnt = min(32, piece_end - c0);
              thr = 1;
              sf = self && c0 < t0 + cnt;
            }
            if (keep) {
              stage_d[q] = nxt;
              if (WT) wbuf[q] = nxtw;
              const unsigned int pa = stage_f_s + (q >> 1) * 16u + (q & 1u) * 4u;
              const float park = -3e18f;
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(pa), "f"(live ? bx : park));
              asm volatile("st.shared.f32 [%0+8], %1;" ::"r"(pa), "f"(live ? by : park));
              asm volatile("st.shared.f32 [%0+512], %1;" ::"r"(pa), "f"(live ? bz : park));
              if (kCyl) asm volatile("st.shared.f32 [%0+520], %1;" ::"r"(pa), "f"(live ? bs : 0.0f));
            }
            jn = c0 + 32 + lane;
            if (jn < piece_end) { nxt = P.pos2[jn]; if (WT) nxtw = P.w2[jn]; }
          }
          __syncwarp();
          while (ring_cnt >= thr) {             // ---- filter + exact pass on the pending blocks (single site) ----
          const int nj = min(ring_cnt, 32);
          const unsigned int blk_f_s = stage_f_s + (unsigned int) ring_head * 8u, blk_d_s = stage_d_s + (unsigned int) ring_head * 32u;
          const T *blk_w = wbuf + ring_head;
          my_made += (unsigned long long) (nj * cnt);

          // ---- filter: every candidate in FP32, one bit per survivor ----
          unsigned int cand[RMAX];              // per primary: the staged secondaries that passed
#pragma unroll
          for (int r = 0; r < RMAX; r++) cand[r] = 0u;
          auto filter = [&](auto rtag, auto selftag) {
            constexpr int R = decltype(rtag)::value;
            constexpr bool SELF = decltype(selftag)::value;
            unsigned int bit = 1u;                    // mask bit of the first point of the staged pair
            int jv = 0;
            const unsigned int se = blk_f_s + (unsigned int) ((nj + 1) >> 1) * 16u;
#pragma unroll 1
            for (unsigned int sa = blk_f_s; sa != se; sa += 16u, bit <<= 2, jv += 2) {
              f32x2 X, Y, Z, S;
              FCFC_LDS_ASM("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(X), "=l"(Y) : "r"(sa));
              if (kCyl) FCFC_LDS_ASM("ld.shared.v2.b64 {%0, %1}, [%2+512];" : "=l"(Z), "=l"(S) : "r"(sa));
              else { FCFC_LDS_ASM("ld.shared.b64 %0, [%1+512];" : "=l"(Z) : "r"(sa)); S = 0; }
#pragma unroll
              for (int r = 0; r < R; r++) {
                const f32x2 dx = sub2(pk2(fx[r], fx[r]), X), dy = sub2(pk2(fy[r], fy[r]), Y), dz = sub2(pk2(fz[r], fz[r]), Z);
                bool ok[2];
                if (BOX && BIN == BIN_SPI) {            // box (s_perp, pi): s_perp^2 and |dz| against their padded limits
                  const f32x2 d2 = fma2(dy, dy, mul2(dx, dx));
                  float a0, a1, z0, z1;
                  upk2(d2, a0, a1); upk2(dz, z0, z1);
                  ok[0] = (a0 < f_d2lim) && (fabsf(z0) < f_plim);
                  ok[1] = (a1 < f_d2lim) && (fabsf(z1) < f_plim);
                } else {
                  const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                  float a0, a1;
                  upk2(d2, a0, a1);
                  ok[0] = a0 < f_d2lim; ok[1] = a1 < f_d2lim;
                  if (kCyl && cyl_on) {
                    // float versions of the division-free cylinder tests of eval_pair, with padded limits:
                    //   pi^2 = (s1 - s2)^2 / |x1 + x2|^2 < p2max   <=>  dd < st * p2max,   st = 2 (s1 + s2) - d^2
                    //   s_perp^2 = d^2 - pi^2 < s2max              <=>  (d^2 - s2max) st < dd
                    const f32x2 S1 = pk2(fs[r], fs[r]);
                    const f32x2 sdif = sub2(S1, S), ssum = add2(S1, S);
                    const f32x2 st = sub2(add2(ssum, ssum), d2), dd = mul2(sdif, sdif);
                    const f32x2 lhs1 = mul2(st, pk2(f_plim, f_plim)), lhs2 = mul2(sub2(d2, pk2(f_s2lim, f_s2lim)), st);
                    float d0, d1, p0, p1, q0, q1;
                    upk2(dd, d0, d1); upk2(lhs1, p0, p1); upk2(lhs2, q0, q1);
                    ok[0] = ok[0] && (d0 < p0) && (q0 <= d0);
                    ok[1] = ok[1] && (d1 < p1) && (q1 <= d1);
                  }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  bool p = ok[h];
                  if (SELF) p = p && (c0 + jv + h > t0 + r * 32 + lane);        // unordered pairs once: metric_common.c:2017-2018
                  if (p) cand[r] |= bit << h;
                }
              }
            }
          };
          if (sf) filter(std::integral_constant<int, RMAX>(), std::true_type());
          else if (RMAX == 4) {
            switch (nr) {
              case 1: filter(std::integral_constant<int, 1>(), std::false_type()); break;
              case 2: filter(std::integral_constant<int, 2>(), std::false_type()); break;
              case 3: filter(std::integral_constant<int, 3>(), std::false_type()); break;
              default: filter(std::integral_constant<int, 4>(), std::false_type()); break;
            }
          } else filter(std::integral_constant<int, RMAX>(), std::false_type());

          // ---- exact pass: the candidates of this chunk in FP64 ----
          // Per primary slot r, one of two walks (a warp-uniform choice by their step counts):
          //   union  every lane evaluates its own primary against each secondary SOME lane selected (broadcast read of the
          //          secondary); lanes that did not select it idle.  Right when most lanes select most secondaries (wide
          //          spheres: 80-90 % of the lane-steps useful);
          //   dealt  the candidates of all lanes are listed in shared memory as (lane, secondary) codes -- a lane writes its
          //          own at the offset a warp prefix sum gives it -- and dealt out evenly: lane i evaluates the candidates
          //          i, i + 32, ... whoever found them (primary from the tile in global memory / L1, secondary from the
          //          staging buffer).  Right for sparse selections (survey (s_perp,pi): the union walk left 56 % of the
          //          lane-steps without a candidate).
#pragma unroll 1
          for (int r = 0; r < nr; r++) {
            unsigned int mine = cand[0];
#pragma unroll
            for (int q = 1; q < RMAX; q++) mine = (r == q) ? cand[q] : mine;
            unsigned int todo = __reduce_or_sync(0xffffffffu, mine);     // secondaries some lane selected for its primary r
            if (todo == 0u) continue;
            const int total = (int) __reduce_add_sync(0xffffffffu, (unsigned int) __popc(mine));
            // (a dealt step costs about 1.4 union steps -- primary from L1, unaligned secondary -- plus two for the list:
            // measured on survey (s,mu) counts, where the union walk is 75 % useful and must stay)
            const bool dealt = 7 * ((total + 31) >> 5) + 10 < 5 * __popc(todo);
            const int nsteps = dealt ? (total + 31) >> 5 : __popc(todo);
            if (dealt) {
              int off = __popc(mine);
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += v; }
              off -= __popc(mine);
              __syncwarp();
              for (unsigned int m = mine; m; m &= m - 1u)
                clist[off++] = (unsigned short) (((unsigned int) lane << 5) | ((unsigned int) __ffs((int) m) - 1u));
              __syncwarp();
            }
            // union walk: this lane's primary r (its candidates only exist when the lane holds a point)
            T ax = 0, ay = 0, az = 0, as = 0, aw = 1;
            if (!dealt && r * 32 + lane < cnt) {
              const Vec4<T> v = P.pos1[t0 + r * 32 + lane];
              ax = BOX ? A::add(v.x, sax) : (kDot ? A::add(v.x, v.x) : v.x);
              ay = BOX ? A::add(v.y, say) : (kDot ? A::add(v.y, v.y) : v.y);
              az = BOX ? A::add(v.z, saz) : (kDot ? A::add(v.z, v.z) : v.z);
              as = v.s;
              if (WT) aw = P.w1[t0 + r * 32 + lane];
            }
#pragma unroll 1
            for (int st = 0; st < nsteps; st++) {
              bool have;
              unsigned int j;
              if (dealt) {
                const int c = st * 32 + lane;
                have = c < total;
                const unsigned int code = have ? (unsigned int) clist[c] : 0u;
                j = code & 31u;
                const int k = r * 32 + (int) (code >> 5);       // (k < cnt: candidates only exist for points the tile holds)
                const Vec4<T> v = P.pos1[t0 + k];
                ax = BOX ? A::add(v.x, sax) : (kDot ? A::add(v.x, v.x) : v.x);
                ay = BOX ? A::add(v.y, say) : (kDot ? A::add(v.y, v.y) : v.y);
                az = BOX ? A::add(v.z, saz) : (kDot ? A::add(v.z, v.z) : v.z);
                as = v.s;
                if (WT) aw = P.w1[t0 + k];
              } else {
                j = (unsigned int) __ffs((int) todo) - 1u;
                todo &= todo - 1u;
                have = (mine >> j) & 1u;
              }
#ifdef FCFC_PF_STATS
              dbg_steps++; dbg_useful += have;
#endif
              const Vec4<T> bq = lds_vec4<T>(blk_d_s + j * 32u);          // (union walk: a broadcast read)
              T d2, aux;
              bool ok = eval_pair<T, BIN, BOX, ARITH, GENERIC>(P, ax, ay, az, as, bq, s2lim, d2, aux);
              ok = ok && have;
              T e[NW];
              T w = (T) 1;
              if (WT) w = A::mul(aw, blk_w[j]);
              if (BIN == BIN_ISO) { e[0] = d2; if (WT) e[1 % NW] = w; }
              else if (BOX) { e[0] = d2; e[1 % NW] = aux; if (WT) { e[2 % NW] = w; e[3 % NW] = 0; } }
              else { e[0] = aux; e[1 % NW] = as; e[2 % NW] = bq.s; e[3 % NW] = WT ? w : (T) 0; }
              if (fast) {
                if constexpr (!GENERIC && SMEMHIST && BIN != BIN_SPI) {
                  unsigned int t;
                  float d2f, auxf;
                  fast_inputs<T, BIN, BOX, NW>(e, d2f, auxf);
                  const int bin = fast_bins<BIN, !(BOX || BIN == BIN_ISO)>(d2f, auxf, P.fb_sscale, P.fb_mscale, P.fb_smask, P.fb_mmask,
                                                                         P.fb_smul, P.fb_mmul, P.ns, t) - (int) P.fb_bias;
                  const bool clean = ok && t != 0u;
                  if (WT) red_shared_f64(hist_s + hlane + hstride * (unsigned int) bin, w, clean);
                  else red_shared_u32_add(clean ? hist_s + 4u * (unsigned int) bin : dump, 1u);
                  if (ok && t == 0u)
                    pf_fix_pair<BIN, BOX, WT, ARITH, NW>(P, hist_s, hstride, hlane, e[0], e[1 % NW], e[2 % NW], e[3 % NW]);
                }
              } else if (ok) {
                T ww;
                const int bin = bin_entry<T, BIN, BOX, WT, ARITH, GENERIC, NW>(P, C, e, ww);
                if (bin >= 0) hist_add<T, WT, SMEMHIST>(C, bin, ww);
              }
            }
          }
          ring_cnt -= nj; ring_head ^= 32;
          }                                     // (pending blocks)
          if (fl) break;
          c0 += 32;
        }
        b = piece_end;
        if (last_piece) flush = false;
      }
    };

    const int nq = (P.periodic ? 3 : 1) * P.nrows, qfirst = P.isauto ? -1 : 0;
    const int qlo = qfirst + (int) ((long long) (nq - qfirst) * split / P.nsplit);
    const int qhi = qfirst + (int) ((long long) (nq - qfirst) * (split + 1) / P.nsplit);
    for (int q = qlo; q <= qhi; q++) {
      int b = 0, e = 0;
      T sax = 0, say = 0, saz = 0, sbx = 0, sby = 0, sbz = 0;
      if (q == qhi) { if (!kClassify) break; }          // closing pass: empties the ring
      else if (q < 0) { b = t0; e = P.cell_start2[cell + 1]; }
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
      const bool closing = q == qhi;
      if (b >= e && !closing) continue;
      sweep_range(b, e, sax, say, saz, sbx, sby, sbz, q < 0 && !closing, closing);
    }
  }

  __syncthreads();
  if (SMEMHIST) {
    if (WT) {
      for (int i = threadIdx.x; i < P.ntot; i += kPfThreads) {
        double v = 0.0;
        for (int c = 0; c < hcopies; c++) v += C.hist_d[i * hcopies + c];
        if (v != 0.0) atomicAdd(&P.ghist_d[i], v);
      }
    } else {
      for (int i = threadIdx.x; i < P.ntot; i += kPfThreads) { const int v = (int) C.hist_u[i]; if (v) atomicAdd(&P.ghist_i[i], (unsigned long long) (long long) v); }
    }
  }
  if (lane == 0 && my_evals) atomicAdd(P.gevals, my_evals);
  if (lane == 0 && my_made) atomicAdd(P.gevals + 3, my_made);
#ifdef FCFC_PF_STATS          // diagnostics build: lane-steps of the exact pass and the useful ones among them
  atomicAdd(P.gevals + 1, dbg_steps); atomicAdd(P.gevals + 2, dbg_useful);
#endif
}

}  // namespace fcfc
