// fcfc_b200/csrc/count_kernel_df.cuh -- double-precision box / isotropic counts at single-precision speed (sm_100a).
//
// The reference's default build is double (src/util/define_comm.h:68-89), but only a few pairs per thousand need double
// precision to be binned like the reference bins them: the ones that sit within rounding distance of a bin edge or of
// the maximum separation.  This kernel therefore does ALL the bulk work in FP32 and re-evaluates just those pairs in FP64:
//
//   * coordinates are taken RELATIVE TO THE CENTRE OF THE TILE'S CELL before they are rounded to float (primaries once per
//     tile, secondaries while they are staged, periodic shifts folded in in double): the values that reach the float
//     arithmetic are bounded by reach + 1.5 cells instead of the box size, so a coordinate difference carries an absolute
//     error e of a few 1e-6 (rescaled units) -- the host derives it from the grid (engine.cu: fcfc_gpu_df_budget);
//   * the pair loop is the in-place loop of the float kernels (count_kernel.cuh: do_chunk_dense): packed f32x2 distances,
//     one vote per primary and staged pair to skip the binning when no lane is in range, computed s and mu bins from one
//     rsqrt.approx with fixed-point edge detection (fast_bins).  The flag bands 2^-ks, 2^-km are widened to cover e on top
//     of the arithmetic error; since the mu error grows like e / s, pairs below a small separation s1 are flagged wholesale
//     (a few 1e-4 of all pairs).  An unflagged pair gets the bins the exact FP64 sequence would give it -- the claim
//     tests/test_df_budget.py checks by emulating this arithmetic on pairs planted at the edges;
//   * a flagged pair (~0.5 %) is pushed on a small per-lane stack as its identity (secondary index, primary slot, periodic
//     image) and re-evaluated later, all lanes together, with eval_pair / bin_entry of count_kernel.cuh in double: the IEEE
//     sequences of the reference (metric_common.c:140-235 scalar order, :377-534 FMA order).  The range test of those pairs
//     is exact too: the float range limit is padded, and everything within the padding is flagged by the s-bin band.
//
// Results are bit-identical to the plain double kernel and to the reference's double builds (tests/test_fullsize_golden.py).
// Variants: box (s,mu) and isotropic, survey isotropic; unweighted and weighted; zero lower bounds, sqrt-type tables,
// shared-memory histogram (everything else takes count_kernel_pf.cuh or the plain kernel).
#pragma once
#include "count_kernel.cuh"
#include <type_traits>

namespace fcfc {

#ifndef FCFC_DF_WARPS
#define FCFC_DF_WARPS 24
#endif
constexpr int kDfWarps = FCFC_DF_WARPS, kDfThreads = kDfWarps * 32;
constexpr int kDfDepth = 16;            // flagged-pair stack entries per lane (8 bytes each)

struct DfPlan { int off_hist, off_rows, off_misc, off_warp, per_warp, o_stage_f, o_wbuf, o_stack, total; };

template <bool WT>
__host__ __device__ inline DfPlan make_df_plan(int ntot, int ns, int nrows, int hist_copies) {
  DfPlan p;
  int o = 0;
  auto al = [](int v) { return (v + 15) & ~15; };
  p.off_hist = o; o += al(WT ? ntot * 8 * hist_copies : (ntot + ns + 1 + 32) * 4);
  p.off_rows = o; o += al(nrows * 16);
  p.off_misc = o; o += 16;
  p.off_warp = o;
  int w = 0;
  p.o_stage_f = w; w += 512;                    // 16 pairs x (x0 x1 y0 y1) | 16 pairs x (z0 z1 - -)
  p.o_wbuf = w; w += WT ? 32 * 8 : 0;
  p.o_stack = w; w += kDfDepth * 32 * 8;        // [slot][lane] x (secondary index, primary slot | image code)
  p.per_warp = w;
  p.total = o + kDfWarps * w;
  return p;
}

template <int BIN, bool BOX, bool WT, int ARITH, int RMAX>
__global__ void __launch_bounds__(kDfThreads, 1) count_kernel_df(const __grid_constant__ CountParams<double> P) {
  using T = double;
  using A = Ar<double>;
  static_assert(BIN != BIN_SPI, "computed bins only");
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NW = QFmt<BIN, BOX, WT>::NW;
  constexpr unsigned int S = 32u * 8u;          // stack stride: one slot of all lanes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hcopies = WT ? P.hist_copies : 1;
  const DfPlan pl = make_df_plan<WT>(P.ntot, P.ns, P.nrows, hcopies);
  unsigned int *hist_u = reinterpret_cast<unsigned int *>(smem + pl.off_hist);
  double *hist_d = reinterpret_cast<double *>(smem + pl.off_hist);
  int4 *s_rows = reinterpret_cast<int4 *>(smem + pl.off_rows);
  unsigned int *s_blk_evals = reinterpret_cast<unsigned int *>(smem + pl.off_misc);

  if (WT) for (int i = threadIdx.x; i < P.ntot * hcopies; i += kDfThreads) hist_d[i] = 0.0;
  else for (int i = threadIdx.x; i < P.ntot + P.ns + 33; i += kDfThreads) hist_u[i] = 0u;
  for (int i = threadIdx.x; i < P.nrows; i += kDfThreads) s_rows[i] = P.rows[i];
  if (threadIdx.x == 0) *s_blk_evals = 0;
  for (int i = threadIdx.x * 16; i < kDfWarps * pl.per_warp; i += kDfThreads * 16)
    *reinterpret_cast<uint4 *>(smem + pl.off_warp + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();

  unsigned char *wbase = smem + pl.off_warp + warp * pl.per_warp;
  T *wbuf = reinterpret_cast<T *>(wbase + pl.o_wbuf);
  const unsigned int stage_s = (unsigned int) __cvta_generic_to_shared(wbase + pl.o_stage_f);
  const unsigned int qbase = (unsigned int) __cvta_generic_to_shared(wbase + pl.o_stack) + 8u * (unsigned int) lane;
  unsigned int qtop = qbase;
  const unsigned int hist_s = (unsigned int) __cvta_generic_to_shared(smem + pl.off_hist);
  const unsigned int hstride = 8u * (unsigned int) hcopies, hlane = 8u * (unsigned int) (lane & (hcopies - 1));
  const unsigned int hist_adj = hist_s + (WT ? hlane : 0u) - (WT ? hstride : 4u) * P.fb_bias;        // base of the biased fast bins
  const unsigned int dump = hist_s + 4u * (unsigned int) (P.ntot + P.ns + 1) + 4u * (unsigned int) lane;
  // exact context of the flagged pairs: tables through global memory (a few pairs per thousand)
  BlockCtx<T> G;
  G.hist_u = nullptr; G.hist_d = nullptr; G.blk_evals = nullptr; G.hmul = 1; G.hoff = 0;
  G.stab = P.stab; G.ptab = P.ptab; G.mutab = P.mutab; G.s2bin = P.s2bin; G.pbin = P.pbin;
  const float f_lim = P.df_d2lim, f_s1sq = P.df_s1sq;
  const float sscale = P.fb_sscale, mscale = P.fb_mscale;
  const unsigned int smask = P.fb_smask, mmask = P.fb_mmask, smul = P.fb_smul, mmul = P.fb_mmul;
  unsigned long long my_evals = 0;
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
    // the local origin: the centre of the tile's cell
    const T ox = A::add(P.gorg[0], A::mul((T) ix + 0.5, P.gcs[0])), oy = A::add(P.gorg[1], A::mul((T) iy + 0.5, P.gcs[1])),
            oz = A::add(P.gorg[2], A::mul((T) iz + 0.5, P.gcs[2]));
    // primaries: float coordinates relative to the origin (no shift: the secondaries carry the whole image shift)
    float fx[RMAX], fy[RMAX], fz[RMAX];
    T aw[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; r++) {
      const int k = r * 32 + lane;
      aw[r] = 1;
      if (k < cnt) {
        const Vec4<T> v = P.pos1[t0 + k];
        fx[r] = __double2float_rn(A::sub(v.x, ox)); fy[r] = __double2float_rn(A::sub(v.y, oy)); fz[r] = __double2float_rn(A::sub(v.z, oz));
        if (WT) aw[r] = P.w1[t0 + k];
      } else fx[r] = fy[r] = fz[r] = 3e18f;           // padding lanes: never in range
    }

    // Exact re-evaluation of the flagged pairs queued so far (all lanes pop together; a lane's entries are its own pairs).
    auto drain = [&]() {
      const int mine = (int) ((qtop - qbase) / S);
      const int mx = __reduce_max_sync(0xffffffffu, mine);
#pragma unroll 1
      for (int q = 0; q < mx; q++) {
        if (q < mine) {
          unsigned int jg, code;
          FCFC_LDS_ASM("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(jg), "=r"(code) : "r"(qbase + (unsigned int) q * S));
          const int r = (int) (code & 3u);
          const Vec4<T> a = P.pos1[t0 + r * 32 + lane];
          Vec4<T> b = P.pos2[jg];
          T ax = a.x, ay = a.y, az = a.z;
          if (BOX) {              // the image of the sweep the pair came from: the lower point gets +L (count_kernel.cuh)
            if (code & 4u) ax = A::add(ax, P.bsize[0]);
            if (code & 8u) ay = A::add(ay, P.bsize[1]);
            if (code & 16u) az = A::add(az, P.bsize[2]);
            if (code & 32u) b.x = A::add(b.x, P.bsize[0]);
            if (code & 64u) b.y = A::add(b.y, P.bsize[1]);
            if (code & 128u) b.z = A::add(b.z, P.bsize[2]);
          }
          T d2, aux;
          if (eval_pair<T, BIN, BOX, ARITH, false>(P, ax, ay, az, a.s, b, P.s2max, d2, aux)) {
            T e[NW], w;
            e[0] = d2;
            if (BIN == BIN_ISO) { if (WT) e[1 % NW] = A::mul(P.w1[t0 + r * 32 + lane], P.w2[jg]); }
            else { e[1 % NW] = aux; if (WT) { e[2 % NW] = A::mul(P.w1[t0 + r * 32 + lane], P.w2[jg]); e[3 % NW] = 0; } }
            const int bin = bin_entry<T, BIN, BOX, WT, ARITH, false, NW>(P, G, e, w);
            if (bin >= 0) {
              if (WT) red_shared_f64(hist_s + hlane + hstride * (unsigned int) bin, w, true);
              else red_shared_u32_add(hist_s + 4u * (unsigned int) bin, 1u);
            }
          }
        }
      }
      qtop = qbase;
      __syncwarp();
    };

    auto sweep_range = [&](int b, int e, T shx, T shy, T shz, unsigned int code, bool self) {
      while (b < e) {
        const int piece_end = min(e, b + kSegPieceMax);
        if (!WT) {                              // overflow accounting of the 32-bit shared counters (count_kernel.cuh)
          unsigned int add = (unsigned int) (piece_end - b) * (unsigned int) cnt, old = 0;
          if (lane == 0) old = atomicAdd(s_blk_evals, add);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old + add >= 0x40000000u || old + add < old) {
            if (lane == 0) atomicExch(s_blk_evals, 0u);
            sweep_hist(hist_u, P.ghist_i, P.ntot, lane);
          }
        }
        {
          unsigned long long ev = (unsigned long long) (piece_end - b) * (unsigned long long) cnt;
          if (self && b == t0) ev -= (unsigned long long) cnt * (unsigned long long) (cnt + 1) / 2;
          my_evals += ev;
        }
        Vec4<T> nxt; T nxtw = 0;
        nxt.x = nxt.y = nxt.z = nxt.s = 0;
        int jn = b + lane;
        if (jn < piece_end) { nxt = P.pos2[jn]; if (WT) nxtw = P.w2[jn]; }
        for (int c0 = b; c0 < piece_end; c0 += 32) {
          __syncwarp();
          {     // stage: (x2 + image shift) - origin, rounded to float, in pair layout
            const bool live = c0 + lane < piece_end;
            const unsigned int pa = stage_s + (unsigned int) (lane >> 1) * 16u + (unsigned int) (lane & 1) * 4u;
            const float park = -3e18f;
            const float bx = __double2float_rn(A::sub(BOX ? A::add(nxt.x, shx) : nxt.x, ox));
            const float by = __double2float_rn(A::sub(BOX ? A::add(nxt.y, shy) : nxt.y, oy));
            const float bz = __double2float_rn(A::sub(BOX ? A::add(nxt.z, shz) : nxt.z, oz));
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(pa), "f"(live ? bx : park));
            asm volatile("st.shared.f32 [%0+8], %1;" ::"r"(pa), "f"(live ? by : park));
            asm volatile("st.shared.f32 [%0+256], %1;" ::"r"(pa), "f"(live ? bz : park));
            if (WT) wbuf[lane] = nxtw;
          }
          __syncwarp();
          jn = c0 + 32 + lane;
          if (jn < piece_end) { nxt = P.pos2[jn]; if (WT) nxtw = P.w2[jn]; }
          const int nj = min(32, piece_end - c0);
          const bool sf = self && c0 < t0 + cnt;

          // One pass over (a part of) the staged chunk; returns the staged index at which the flagged-pair stacks must
          // be emptied first, or nj.
          auto chunk = [&](auto rtag, auto selftag, int j0) -> int {
            constexpr int R = decltype(rtag)::value;
            constexpr bool SELF = decltype(selftag)::value;
            const unsigned int lim = qbase + (unsigned int) (kDfDepth - 2 * R) * S;     // proceed while fill + 2 R <= depth
            unsigned int sa = stage_s + (unsigned int) (j0 >> 1) * 16u;
            const unsigned int sa0 = sa, se = stage_s + (unsigned int) ((nj + 1) >> 1) * 16u;
            unsigned int jg = (unsigned int) (c0 + j0);
#pragma unroll 1
            for (; sa != se; sa += 16u, jg += 2u) {
              if (__any_sync(0xffffffffu, qtop > lim)) break;
              f32x2 X, Y, Z;
              FCFC_LDS_ASM("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(X), "=l"(Y) : "r"(sa));
              FCFC_LDS_ASM("ld.shared.b64 %0, [%1+256];" : "=l"(Z) : "r"(sa));
#pragma unroll
              for (int r = 0; r < R; r++) {
                const f32x2 dx = sub2(pk2(fx[r], fx[r]), X), dy = sub2(pk2(fy[r], fy[r]), Y), dz = sub2(pk2(fz[r], fz[r]), Z);
                const f32x2 d2 = fma2(dy, dy, fma2(dx, dx, mul2(dz, dz)));
                float d2h[2], dzh[2];
                upk2(d2, d2h[0], d2h[1]); upk2(dz, dzh[0], dzh[1]);
                if (!__any_sync(0xffffffffu, fminf(d2h[0], d2h[1]) < f_lim)) continue;   // no lane has either pair in range
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  unsigned int t;
                  const int bin = fast_bins<BIN, false>(d2h[h], dzh[h], sscale, mscale, smask, mmask, smul, mmul, P.ns, t);
                  bool in = d2h[h] < f_lim;
                  if (SELF) in = in && ((int) jg + h > t0 + r * 32 + lane);             // unordered pairs once
                  if (BIN == BIN_SMU) t = (d2h[h] < f_s1sq) ? 0u : t;                    // mu is not resolved in float below s1
                  const bool clean = in && t != 0u, flagged = in && t == 0u;
                  if (WT) {
                    if (clean) {        // (product in `real`, summed in double: metric_common.c:216-231)
                      const T w = A::mul(aw[r], wbuf[(jg - (unsigned int) c0) + h]);
                      red_shared_f64(hist_adj + hstride * (unsigned int) bin, w, true);
                    }
                  } else red_shared_u32_add(clean ? hist_adj + 4u * (unsigned int) bin : dump, 1u);
                  asm volatile("{.reg .pred q; setp.ne.s32 q, %1, 0; @q st.shared.v2.u32 [%0], {%2, %3}; @q add.u32 %0, %0, 256;}"
                               : "+r"(qtop) : "r"((int) flagged), "r"(jg + (unsigned int) h), "r"(code | (unsigned int) r));
                }
              }
            }
            return min(j0 + (int) ((sa - sa0) >> 3), nj);
          };
          for (int j = 0;;) {
            if (sf) j = chunk(std::integral_constant<int, RMAX>(), std::true_type(), j);
            else if (RMAX >= 3) {
              switch (nr) {
                case 1: j = chunk(std::integral_constant<int, 1>(), std::false_type(), j); break;
                case 2: j = chunk(std::integral_constant<int, 2>(), std::false_type(), j); break;
                case 3: j = chunk(std::integral_constant<int, 3>(), std::false_type(), j); break;
                default: j = chunk(std::integral_constant<int, RMAX>(), std::false_type(), j); break;
              }
            } else j = chunk(std::integral_constant<int, RMAX>(), std::false_type(), j);
            if (j >= nj) break;
            drain();
          }
        }
        b = piece_end;
      }
    };

    const int nq = (P.periodic ? 3 : 1) * P.nrows, qfirst = P.isauto ? -1 : 0;
    const int qlo = qfirst + (int) ((long long) (nq - qfirst) * split / P.nsplit);
    const int qhi = qfirst + (int) ((long long) (nq - qfirst) * (split + 1) / P.nsplit);
    for (int q = qlo; q < qhi; q++) {
      int b, e;
      unsigned int code = 0;                    // bits 2-4: the primaries get +L in x, y, z; bits 5-7: the secondaries do
      T shx = 0, shy = 0, shz = 0;              // what the secondaries get relative to the primaries
      if (q < 0) { b = t0; e = P.cell_start2[cell + 1]; }
      else {
        const int ri = P.periodic ? q / 3 : q, img = P.periodic ? q - 3 * ri : 1;
        const int4 row = s_rows[ri];
        int jx = ix + row.x, jy = iy + row.y;
        int zlo = iz + row.z, zhi = iz + row.w;
        if (P.periodic) {
          if (jx >= P.nc[0]) { jx -= P.nc[0]; shx = P.bsize[0]; code |= 32u; } else if (jx < 0) { jx += P.nc[0]; shx = -P.bsize[0]; code |= 4u; }
          if (jy >= ncy) { jy -= ncy; shy = P.bsize[1]; code |= 64u; } else if (jy < 0) { jy += ncy; shy = -P.bsize[1]; code |= 8u; }
          if (img == 0) { zhi = min(zhi, -1) + ncz; zlo += ncz; shz = -P.bsize[2]; code |= 16u; }
          else if (img == 1) { zlo = max(zlo, 0); zhi = min(zhi, ncz - 1); }
          else { zlo = max(zlo, ncz) - ncz; zhi -= ncz; shz = P.bsize[2]; code |= 128u; }
        } else {
          if (jx < 0 || jx >= P.nc[0] || jy < 0 || jy >= ncy) continue;
          zlo = max(zlo, 0); zhi = min(zhi, ncz - 1);
        }
        if (zlo > zhi) continue;
        const int rowbase = (jx * ncy + jy) * ncz;
        b = P.cell_start2[rowbase + zlo]; e = P.cell_start2[rowbase + zhi + 1];
      }
      if (b >= e) continue;
      sweep_range(b, e, shx, shy, shz, code, q < 0);
    }
    drain();                    // the queued identities refer to this tile
  }

  __syncthreads();
  if (WT) {
    for (int i = threadIdx.x; i < P.ntot; i += kDfThreads) {
      double v = 0.0;
      for (int c = 0; c < hcopies; c++) v += hist_d[i * hcopies + c];
      if (v != 0.0) atomicAdd(&P.ghist_d[i], v);
    }
  } else {
    for (int i = threadIdx.x; i < P.ntot; i += kDfThreads) { const int v = (int) hist_u[i]; if (v) atomicAdd(&P.ghist_i[i], (unsigned long long) (long long) v); }
  }
  if (lane == 0 && my_evals) atomicAdd(P.gevals, my_evals);
}

}  // namespace fcfc
