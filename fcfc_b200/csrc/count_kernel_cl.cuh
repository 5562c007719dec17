// fcfc_b200/csrc/count_kernel_cl.cuh -- single-precision box / isotropic counts with computed bins: every staged
// secondary point is first classified against the bounding box of the tile (sm_100a).
//
// Same job as count_kernel.cuh (reference: count_dual_node / compute_dist_vector / update_hist_vector,
// src/fcfc/2pt_box/metric_common.c:377-534, 774-957, 980-1482, and the node pruning of dual_tree.c:245-373), same
// arithmetic per pair (packed_dist / fast_bins / drain_fast / fix_entry are the ones of count_kernel.cuh: bit-identical
// results), different bookkeeping around it.  The cell stencil prunes at the granularity of cells: at the cell size
// that is fastest for the bench workload it hands the pair loop 2.06 candidates per pair in range, and only a third of
// the accepted pairs sit in cells that are entirely in range.  Here the lane that stages a secondary point also tests it
// against the bounding box of the tile's primaries (centre c, half-widths h, 18 FP32 instructions per staged point,
// i.e. per <= 96 pair evaluations):
//
//   * nearest point of the box farther than the maximum separation  -> the point is dropped (no pair can be in range);
//   * farthest corner of the box within the maximum separation      -> "dense" point: every pair is in range, the pair
//     loop bins it in place with all lanes busy (do_chunk_dense: no stack traffic);
//   * otherwise                                                     -> "partial" point: ordinary pair loop, accepted
//     pairs pushed on the per-lane stacks and binned by the drain.
//
// The warp compacts the two kinds into two ring buffers of 64 points (ballot + population count give every lane its
// slot) and runs the pair loops on full blocks of 32 points, so the loops never see a ragged chunk except when the rings
// are flushed: at the end of a work item and when the periodic image shift of the PRIMARIES changes (the shift of the
// secondaries is applied while staging; the lower point gets +L as in count_kernel.cuh, which keeps the float results
// identical to the other kernels').  For the bench workload the evaluated candidates per pair in range drop from 2.06 to
// about 1.5 and two thirds of the accepted pairs take the dense path.
//
// Exactness does not depend on the classification being sharp: the dense loop keeps its range test, and a point is only
// dropped when the computed distance to the box exceeds the limit by a margin that covers every rounding on the way
// (cl_skip, set by the host from the coordinate magnitudes).
//
// Variants: float, unweighted, zero lower bounds, sqrt-type tables, shared-memory histogram; box (s,mu) / isotropic and
// survey isotropic -- the variants whose drain computes its bins.  Everything else takes count_kernel.cuh.
#pragma once
#include "count_kernel.cuh"

namespace fcfc {

#ifndef FCFC_CL_WARPS
#define FCFC_CL_WARPS 24
#endif
constexpr int kClWarps = FCFC_CL_WARPS, kClThreads = kClWarps * 32;
constexpr int kClR = 3;                 // primaries per lane: tiles of up to 96 points (the other kernels: 4)
constexpr int kClRingBytes = 64 * 16;   // 64 points in pair layout: [pair][x0 x1 y0 y1 z0 z1 - -]

struct ClPlan { int off_hist, off_rows, off_misc, off_warp, per_warp, o_ring_d, o_ring_p, o_box, off_queue, queue_per_warp, total; };

__host__ __device__ inline ClPlan make_cl_plan(int ntot, int ns, int nrows, int qwords, int qdepth) {
  ClPlan p;
  int o = 0;
  auto al = [](int v) { return (v + 15) & ~15; };
  p.off_hist = o; o += al((ntot + ns + 1 + 32) * 4);    // as in make_smem_plan: fast bins may land one row / column outside, 32 dump slots
  p.off_rows = o; o += al(nrows * 16);
  p.off_misc = o; o += 16;
  p.off_warp = o;
  int w = 0;
  p.o_ring_d = w; w += kClRingBytes;
  p.o_ring_p = w; w += kClRingBytes;
  p.o_box = w; w += 32;                                 // (cx, cy, cz, -) (hx, hy, hz, -) of the tile, image shift included
  p.per_warp = w;
  o += kClWarps * w;
  p.off_queue = o;
  p.queue_per_warp = qdepth * 32 * qwords * 4;
  o += kClWarps * p.queue_per_warp;
  o += 4 * 32 * qwords * 4;                             // over-read pad of the four-entry drain
  p.total = o;
  return p;
}

template <int BIN, bool BOX, int ARITH, int RMAX>
__global__ void __launch_bounds__(kClThreads, 1) count_kernel_cl(const __grid_constant__ CountParams<float> P) {
  using T = float;
  static_assert(BIN != BIN_SPI, "computed bins only");
  static_assert(BOX || BIN == BIN_ISO, "survey (s,mu) takes the dot-product form of count_kernel.cuh");
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NW = QFmt<BIN, BOX, false>::NW;
  constexpr unsigned int S = LaneQueue<T, NW>::kStride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const ClPlan pl = make_cl_plan(P.ntot, P.ns, P.nrows, NW, P.qdepth);
  unsigned int *hist_u = reinterpret_cast<unsigned int *>(smem + pl.off_hist);
  int4 *s_rows = reinterpret_cast<int4 *>(smem + pl.off_rows);
  unsigned int *s_blk_evals = reinterpret_cast<unsigned int *>(smem + pl.off_misc);

  for (int i = threadIdx.x; i < P.ntot + P.ns + 33; i += kClThreads) hist_u[i] = 0u;
  for (int i = threadIdx.x; i < P.nrows; i += kClThreads) s_rows[i] = P.rows[i];
  if (threadIdx.x == 0) {
    *s_blk_evals = 0;
    *reinterpret_cast<float *>(smem + pl.off_misc + 8) = P.s2max;
    *reinterpret_cast<float *>(smem + pl.off_misc + 4) = -0.0f;
  }
  for (int i = threadIdx.x * 16; i < pl.total - pl.off_warp; i += kClThreads * 16)
    *reinterpret_cast<uint4 *>(smem + pl.off_warp + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();

  BlockCtx<T> C;
  C.hist_u = hist_u; C.hist_d = nullptr; C.hmul = 1; C.hoff = 0;
  C.stab = P.stab; C.ptab = P.ptab; C.mutab = P.mutab; C.s2bin = P.s2bin; C.pbin = P.pbin; C.blk_evals = s_blk_evals;
  FastCtx F;
  F.hist_s = (unsigned int) __cvta_generic_to_shared(hist_u);
  F.hstride = 8u; F.hlane = 0u; F.stab_s = F.ptab_s = F.mutab_s = 0u;

  unsigned char *wbase = smem + pl.off_warp + warp * pl.per_warp;
  const unsigned int ring_s = (unsigned int) __cvta_generic_to_shared(wbase);       // dense ring; the partial ring follows it
  float *s_box = reinterpret_cast<float *>(wbase + pl.o_box);
  const unsigned int box_s = (unsigned int) __cvta_generic_to_shared(s_box);
  LaneQueue<T, NW> Q;
  Q.wbase = (unsigned int) __cvta_generic_to_shared(smem + pl.off_queue) + warp * (unsigned int) pl.queue_per_warp;
  Q.base = Q.top = Q.wbase + QOps<T, NW>::lane_offset(lane);
  int ub = 0;
  // parked in registers through shared memory (see count_kernel.cuh)
  const T s2lim = *reinterpret_cast<volatile T *>(smem + pl.off_misc + 8);
  const float negzero = *reinterpret_cast<volatile float *>(smem + pl.off_misc + 4);
  unsigned long long my_cand = 0, my_evals = 0;
  const int ncy = P.nc[1], ncz = P.nc[2];
  const unsigned int lt_mask = (1u << lane) - 1u;
  const T ps[RMAX] = {}, pw[RMAX] = {};         // (unused by the unweighted box / isotropic pair loops)

  auto drain = [&](int need, int keep) -> int {
    const int mx = (int) (__reduce_max_sync(0xffffffffu, Q.fill_bytes()) / S);
    if (mx + need <= P.qdepth - 1) return mx;
    const int rounds = min((mx - keep + 3) & ~3, 32);
    if (rounds <= 0) return 0;
    __syncwarp();               // the columns rotate between the lanes: pushes of other lanes must be visible to whoever pops
    drain_fast<T, BIN, BOX, false, ARITH, NW>(P, C, F, Q, rounds);      // (ends with __syncwarp())
    return max(mx - rounds, 0);
  };

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

    // primaries with the current image shift; the bounding box of the unshifted tile
    T ax[RMAX], ay[RMAX], az[RMAX];
    T cx, cy, cz, hx, hy, hz;
    {
      T lox = Ar<T>::huge(), loy = lox, loz = lox, hix = -lox, hiy = -lox, hiz = -lox;
#pragma unroll
      for (int r = 0; r < RMAX; r++) {
        const int k = tile_slot(r, lane);
        if (k < cnt) {
          const Vec4<T> v = P.pos1[t0 + k];
          ax[r] = v.x; ay[r] = v.y; az[r] = v.z;
          lox = fminf(lox, v.x); hix = fmaxf(hix, v.x); loy = fminf(loy, v.y); hiy = fmaxf(hiy, v.y); loz = fminf(loz, v.z); hiz = fmaxf(hiz, v.z);
        } else ax[r] = ay[r] = az[r] = Ar<T>::far();
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
        loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
      }
      cx = 0.5f * (lox + hix); cy = 0.5f * (loy + hiy); cz = 0.5f * (loz + hiz);
      hx = 0.5f * (hix - lox); hy = 0.5f * (hiy - loy); hz = 0.5f * (hiz - loz);
    }
    T cur_x = 0, cur_y = 0, cur_z = 0;          // image shift the primaries carry
    __syncwarp();
    if (lane == 0) { s_box[0] = cx; s_box[1] = cy; s_box[2] = cz; s_box[4] = hx; s_box[5] = hy; s_box[6] = hz; }
    // ring state: pending points [head, head + cnt) mod 64, head is 0 or 32
    int cnt_d = 0, cnt_p = 0, head_d = 0, head_p = 0;

    // One range [b, e) of secondary points with the image shift sb of the secondaries.  `flush` first empties the rings
    // (ragged blocks) and then gives the primaries the image shift (sax, say, saz); `self`: the tile's own cell, staged
    // without classification (pairs i < j only).
    auto sweep_range = [&](int b, int e, T sax, T say, T saz, T sbx, T sby, T sbz, bool self, bool flush) {
      while (flush || b < e) {
        const int piece_end = min(e, b + kSegPieceMax);
        if (b < e) {
          // overflow accounting of the 32-bit shared counters (see sweep_hist): an upper bound of the increments
          unsigned int add = (unsigned int) (piece_end - b) * (unsigned int) cnt, old = 0;
          if (lane == 0) old = atomicAdd(s_blk_evals, add);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old + add >= 0x40000000u || old + add < old) {
            if (lane == 0) atomicExch(s_blk_evals, 0u);
            sweep_hist(hist_u, P.ghist_i, P.ntot, lane);
          }
          unsigned long long ev = (unsigned long long) (piece_end - b) * (unsigned long long) cnt;
          if (self && b == t0) { ev -= (unsigned long long) cnt * (unsigned long long) (cnt + 1) / 2; my_evals -= (unsigned long long) cnt * (unsigned long long) (cnt + 1) / 2; }
          my_cand += ev;
        }
        Vec4<T> nxt; nxt.x = nxt.y = nxt.z = nxt.s = 0;
        int jn = b + lane;
        if (!flush && jn < piece_end) nxt = P.pos2[jn];
        for (int c0 = b; flush || c0 < piece_end;) {
          __syncwarp();
          int thr = 32;
          bool sf = false;
          if (flush) {
            // odd fills get a parked partner (never in range), then everything pending is processed
            if (lane == 0) {
              if (cnt_d & 1) { const unsigned int pa = staged_pair_addr(ring_s, (head_d + cnt_d) & 63); asm volatile("st.shared.f32 [%0], %1; st.shared.f32 [%0+8], %1; st.shared.f32 [%0+16], %1;" ::"r"(pa), "f"(-Ar<T>::far())); }
              if (cnt_p & 1) { const unsigned int pa = staged_pair_addr(ring_s + kClRingBytes, (head_p + cnt_p) & 63); asm volatile("st.shared.f32 [%0], %1; st.shared.f32 [%0+8], %1; st.shared.f32 [%0+16], %1;" ::"r"(pa), "f"(-Ar<T>::far())); }
            }
            thr = 1;
          } else {
            const bool live = c0 + lane < piece_end;
            if (BOX) { nxt.x = Ar<T>::add(nxt.x, sbx); nxt.y = Ar<T>::add(nxt.y, sby); nxt.z = Ar<T>::add(nxt.z, sbz); }
            if (self) {
              // the tile's own cell: all 32 lanes store (lanes past the end park a point), block head_p of the empty partial ring
              const unsigned int pa = staged_pair_addr(ring_s + kClRingBytes, head_p + lane);
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(pa), "f"(live ? nxt.x : -Ar<T>::far()));
              asm volatile("st.shared.f32 [%0+8], %1;" ::"r"(pa), "f"(live ? nxt.y : -Ar<T>::far()));
              asm volatile("st.shared.f32 [%0+16], %1;" ::"r"(pa), "f"(live ? nxt.z : -Ar<T>::far()));
              cnt_p = min(32, piece_end - c0);
              thr = 1;
              sf = c0 < t0 + cnt;
            } else {
              // distance of the point to the nearest / farthest point of the tile's box, squared
              float bcx, bcy, bcz, bhx, bhy, bhz, unused;
              lds_vec4_raw(box_s, bcx, bcy, bcz, unused);
              lds_vec4_raw(box_s + 16u, bhx, bhy, bhz, unused);
              const float tx = fabsf(__fsub_rn(nxt.x, bcx)), ty = fabsf(__fsub_rn(nxt.y, bcy)), tz = fabsf(__fsub_rn(nxt.z, bcz));
              const float ux = fmaxf(__fsub_rn(tx, bhx), 0.0f), uy = fmaxf(__fsub_rn(ty, bhy), 0.0f), uz = fmaxf(__fsub_rn(tz, bhz), 0.0f);
              const float vx = __fadd_rn(tx, bhx), vy = __fadd_rn(ty, bhy), vz = __fadd_rn(tz, bhz);
              const float dmin2 = __fmaf_rn(uz, uz, __fmaf_rn(uy, uy, __fmul_rn(ux, ux)));
              const float dmax2 = __fmaf_rn(vz, vz, __fmaf_rn(vy, vy, __fmul_rn(vx, vx)));
              const bool keep = live && !(dmin2 > P.cl_skip);
              const bool dn = keep && dmax2 < P.cl_dense;
              const unsigned int md = __ballot_sync(0xffffffffu, dn), mp = __ballot_sync(0xffffffffu, keep && !dn);
              const int pos = (dn ? head_d + cnt_d + __popc(md & lt_mask) : head_p + cnt_p + __popc(mp & lt_mask)) & 63;
              const unsigned int pa = staged_pair_addr(ring_s + (dn ? 0u : (unsigned int) kClRingBytes), pos);
              if (keep) {
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(pa), "f"(nxt.x));
                asm volatile("st.shared.f32 [%0+8], %1;" ::"r"(pa), "f"(nxt.y));
                asm volatile("st.shared.f32 [%0+16], %1;" ::"r"(pa), "f"(nxt.z));
              }
              cnt_d += __popc(md); cnt_p += __popc(mp);
            }
            jn = c0 + 32 + lane;
            if (jn < piece_end) nxt = P.pos2[jn];
          }
          __syncwarp();
          // ---- pair loops on the pending blocks (the single call site of the unrolled loops) ----
          while (true) {
            bool dense;
            if (cnt_d >= thr) dense = true; else if (cnt_p >= thr) dense = false; else break;
            const int nj = min(dense ? cnt_d : cnt_p, 32);
            const int head = dense ? head_d : head_p;
            const Vec4<T> *sbuf = reinterpret_cast<const Vec4<T> *>(wbase + (dense ? 0 : kClRingBytes)) + head;
            my_evals += (unsigned long long) (nj * cnt);
#define FCFC_CHUNK_DENSE(RR) do_chunk_dense<T, BIN, BOX, ARITH, RR, NW, RMAX>(P, Q, ub, F.hist_s, lane, sbuf, j, nj, ax, ay, az, s2lim, negzero)
#define FCFC_CHUNK(RR, SF) do_chunk<T, BIN, BOX, false, ARITH, false, RR, SF, NW, RMAX>(P, Q, ub, sbuf, nullptr, j, nj, ax, ay, az, ps, pw, c0, t0, lane, s2lim, negzero)
            for (int j = 0;;) {
              if (sf) j = FCFC_CHUNK(RMAX, true);
              else if (dense) {
                switch (nr) {
                  case 1: j = FCFC_CHUNK_DENSE(1); break;
                  case 2: j = FCFC_CHUNK_DENSE(2); break;
                  case 3: j = FCFC_CHUNK_DENSE((RMAX < 3 ? RMAX : 3)); break;
                  default: j = FCFC_CHUNK_DENSE(RMAX); break;
                }
              } else {
                switch (nr) {
                  case 1: j = FCFC_CHUNK(1, false); break;
                  case 2: j = FCFC_CHUNK(2, false); break;
                  case 3: j = FCFC_CHUNK((RMAX < 3 ? RMAX : 3), false); break;
                  default: j = FCFC_CHUNK(RMAX, false); break;
                }
              }
              if (j >= nj) break;
              ub = drain(2 * RMAX, P.qkeep);
            }
#undef FCFC_CHUNK
#undef FCFC_CHUNK_DENSE
            if (dense) { cnt_d -= nj; head_d ^= 32; } else { cnt_p -= nj; head_p ^= 32; }
          }
          if (flush) {
            flush = false;
            if (sax != cur_x || say != cur_y || saz != cur_z) {
              cur_x = sax; cur_y = say; cur_z = saz;
#pragma unroll
              for (int r = 0; r < RMAX; r++) {
                const int k = tile_slot(r, lane);
                if (k < cnt) {
                  const Vec4<T> v = P.pos1[t0 + k];
                  ax[r] = Ar<T>::add(v.x, sax); ay[r] = Ar<T>::add(v.y, say); az[r] = Ar<T>::add(v.z, saz);
                }
              }
              __syncwarp();
              if (lane == 0) { s_box[0] = cx + sax; s_box[1] = cy + say; s_box[2] = cz + saz; }
            }
            jn = b + lane;
            if (jn < piece_end) nxt = P.pos2[jn];
          } else c0 += 32;
        }
        b = piece_end;
      }
    };

    // Sweep list as in count_kernel.cuh (q = -1: the tile's own cell; q = 3 row + image), plus one closing pass that
    // flushes the rings.
    const int nq = (P.periodic ? 3 : 1) * P.nrows, qfirst = P.isauto ? -1 : 0;
    const int qlo = qfirst + (int) ((long long) (nq - qfirst) * split / P.nsplit);
    const int qhi = qfirst + (int) ((long long) (nq - qfirst) * (split + 1) / P.nsplit);
    for (int q = qlo; q <= qhi; q++) {
      int b = 0, e = 0;
      T sax = 0, say = 0, saz = 0, sbx = 0, sby = 0, sbz = 0;
      bool flush = false;
      if (q == qhi) { flush = true; sax = cur_x; say = cur_y; saz = cur_z; }          // closing pass
      else if (q < 0) { b = t0; e = P.cell_start2[cell + 1]; flush = true; }
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
        if (b >= e) continue;
        flush = (sax != cur_x || say != cur_y || saz != cur_z);
      }
      if (b >= e && !flush) continue;
      sweep_range(b, e, sax, say, saz, sbx, sby, sbz, q < 0 && q != qhi, flush);
    }
  }
  while (drain(P.qdepth, 0) > 0) {}

  __syncthreads();
  for (int i = threadIdx.x; i < P.ntot; i += kClThreads) { const int v = (int) hist_u[i]; if (v) atomicAdd(&P.ghist_i[i], (unsigned long long) (long long) v); }
  if (lane == 0 && my_cand) atomicAdd(P.gevals, my_cand);
  if (lane == 0 && my_evals) atomicAdd(P.gevals + 3, my_evals);
}

}  // namespace fcfc
