/*******************************************************************************
* include/fcfc_gpu.h -- C ABI of the B200-native pair-counting engine (libfcfc_b200.so).
*
* This is the drop-in boundary for the counting step of cheng-zhao/FCFC v1.0.1
* (FCFC_2PT_BOX and FCFC_2PT).  It replaces, behind plain C types only:
*
*   reference seam (paths relative to the reference's src/)      replaced by
*   ----------------------------------------------------------   --------------------------
*   tree_create()   fcfc/2pt_box/build_tree.h:47, build_tree.c:159-177   fcfc_gpu_catalog_create()
*                   (create_kdtree / create_balltree, tree/kdtree.c:434,
*                    tree/balltree.c:712 -- the spatial index)           (GPU cell list)
*   tree_destroy()  fcfc/2pt_box/build_tree.h:62                         fcfc_gpu_catalog_destroy()
*   count_pairs()   fcfc/2pt_box/count_func.h:54, count_func.c:4847;     fcfc_gpu_count()
*                   fcfc/2pt/count_func.c:4846 (dual-tree traversal,
*                   metric_common.c kernels, OpenMP/MPI scheduling and
*                   reductions)
*   the slice of `CF` the counter reads  fcfc/2pt_box/eval_cf.h:54-67    fcfc_gpu_bins
*   MPI_Ireduce of the histograms        count_func.c:7654-7721          partial counts + one
*                                                                        all-reduce (NCCL)
*
* The C99 host shim that implements the reference's three functions on top of
* this ABI is fcfc_b200/host/fcfc_gpu_shim.c; INTEGRATION.md shows the build
* lines a maintainer adds.  There is no CPU fallback: every entry point fails
* with FCFC_GPU_ERR_CUDA when no usable sm_100 device is present.
*
* Conventions (identical to the reference's count_pairs):
*   - coordinates handed over are already multiplied by cf->rescale, unless a
*     `rescale` factor != 1 is passed, in which case the engine performs the
*     same single multiplication in `real` precision (build_tree.c:121-131);
*   - an auto count returns every unordered pair once (the caller doubles,
*     eval_cf.c:126-133); self pairs are never counted;
*   - counts are int64 when !withwt, double sums of (w1*w2 formed in `real`)
*     when withwt; layout cnt[s_idx + p_idx * ns] (metric_common.c:218);
*   - bins are [lo, hi); pairs with mu = 1 are dropped unless with_mu_one.
*******************************************************************************/
#ifndef FCFC_GPU_H
#define FCFC_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCFC_GPU_ABI_VERSION    1

/* Error codes: 0 = success; negative values mirror util/define_comm.h:218-234. */
#define FCFC_GPU_OK             0
#define FCFC_GPU_ERR_MEMORY     (-1)    /* FCFC_ERR_MEMORY  */
#define FCFC_GPU_ERR_ARG        (-2)    /* FCFC_ERR_ARG     */
#define FCFC_GPU_ERR_DATA       (-21)   /* FCFC_ERR_DATA: point outside the periodic box, NaN */
#define FCFC_GPU_ERR_TREE       (-23)   /* FCFC_ERR_TREE: cell-list construction failed */
#define FCFC_GPU_ERR_CF         (-24)   /* FCFC_ERR_CF: counting kernel failed */
#define FCFC_GPU_ERR_CUDA       (-50)   /* no device / CUDA runtime error */

/* Binning schemes, fcfc/2pt_box/define.h:45-47. */
#define FCFC_GPU_BIN_ISO        0
#define FCFC_GPU_BIN_SMU        1
#define FCFC_GPU_BIN_SPI        2
/* Lookup table types, define.h:49-50. */
#define FCFC_GPU_TAB_INT        0
#define FCFC_GPU_TAB_HYBRID     1
/* Arithmetic order of the per-pair formulas. */
#define FCFC_GPU_ARITH_SCALAR   0   /* unfused, order of compute_dist_hist_scalar
                                       (metric_common.c:140-235): the bit-exact parity mode */
#define FCFC_GPU_ARITH_FMA      1   /* FMA chain + (num*nmu^2)/d^2 rounded toward zero, as the
                                       AVX-512 compute_dist_vector (metric_common.c:377-534) */

/* The part of the reference's `CF` structure read by count_pairs.  All arrays are host
 * pointers to `real` (float if is_float, else double) and are copied by the callee. */
typedef struct {
  int32_t bintype;      /* FCFC_GPU_BIN_*  (cf->bintype)                                   */
  int32_t periodic;     /* 1: FCFC_2PT_BOX metric (periodic wrap), 0: FCFC_2PT survey metric */
  int32_t is_float;     /* 1: reference built with -DSINGLE_PREC                            */
  int32_t tabtype;      /* FCFC_GPU_TAB_*  (cf->tabtype)                                    */
  int32_t ns, np, nmu;  /* cf->ns, cf->np, cf->nmu                                          */
  int32_t swidth;       /* cf->swidth: 0 = uint8_t, 1 = uint16_t entries                    */
  int32_t pwidth;       /* cf->pwidth                                                       */
  int32_t with_mu_one;  /* reference built with -DWITH_MU_ONE                               */
  int32_t arith;        /* FCFC_GPU_ARITH_*                                                 */
  int32_t reserved;
  const void *s2bin;    /* real[ns+1]: cf->s2bin                                            */
  const void *pbin;     /* real[np+1]: cf->pbin (box: pi) / cf->p2bin (survey: pi^2); SPI only */
  const void *stab;     /* cf->stab, nstab entries                                          */
  const void *ptab;     /* cf->ptab, nptab entries; SPI only                                */
  const uint8_t *mutab; /* cf->mutab, nmu*nmu entries; SMU only                             */
  uint64_t nstab;       /* 0: derive from the edges as util/create_lut.c:62-64,105-107 does */
  uint64_t nptab;
  double bsize[3];      /* cf->bsize (rescaled box), periodic only                          */
} fcfc_gpu_bins;

typedef struct fcfc_gpu_catalog fcfc_gpu_catalog;      /* opaque, replaces the KDT / BLT handle */

/* Statistics of the last fcfc_gpu_count / fcfc_gpu_count_partial call on this thread. */
typedef struct {
  uint64_t pair_evals;  /* candidate distance evaluations (sum of n_a*n_b over visited cell pairs) */
  uint64_t pairs_in;    /* pairs that landed in the histogram (unweighted) / 0                    */
  double ms_sort;       /* device time of the cell-list (re)builds done by this call             */
  double ms_count;      /* device time of the counting kernel(s)                                  */
  double ms_total;      /* device time of the whole call                                          */
  uint32_t kernel_launches; /* number of kernels of this library launched by the call            */
  int32_t ncell[3];     /* cell grid used                                                         */
  int32_t nitem;        /* number of (cell, tile) work items                                      */
  int32_t dense_rows;   /* stencil rows with cells binned in place (dense-cell path), 0 if unused */
  int32_t prefilter;    /* double-precision counts: 0 plain FP64 kernel, 1 float pre-filter + FP64 pass on the candidates,
                           2 float-speed kernel (FP32 on cell-relative coordinates, FP64 for the pairs near an edge)   */
  int32_t classified;   /* 1: staged points were classified against the tile's bounding box (count_kernel_cl.cuh) */
  uint64_t pair_evals_computed; /* distance evaluations actually made: pair_evals minus the candidates the classification
                                   dropped wholesale (equal to pair_evals for the other kernels)              */
} fcfc_gpu_stats;

/* Bind the calling process to CUDA devices.  ndev <= 0: all visible devices.  `devices` may be
 * NULL (0..ndev-1).  A multi-process launcher (one rank per GPU) passes ndev = 1 and its local
 * device.  Returns the number of devices in use or a negative error. */
int fcfc_gpu_init(int ndev, const int *devices, int verbose);
void fcfc_gpu_finalize(void);
const char *fcfc_gpu_last_error(void);
int fcfc_gpu_abi_version(void);

/* Upload a catalogue (SoA arrays of `real`; host pointers -- pageable or pinned -- or device pointers, resolved by
 * unified addressing) and keep it resident on every device in use: the first device reads the caller's arrays,
 * the others copy from it device to device (replaces kdtree_broadcast, tree/kdtree.c:529-619).
 *   x2sum: survey 4th coordinate x^2+y^2+z^2 as left by data_preprocess (fcfc/2pt/build_tree.c:35);
 *          NULL: computed on the device when the metric needs it, in the order selected by
 *          `sumsq_arith` (FCFC_GPU_ARITH_SCALAR: build_tree.c:59; _FMA: build_tree.c:75-82).
 *   w:     weights or NULL.
 *   rescale: coordinates are multiplied by (real) rescale on the device when != 1.
 * Returns NULL on failure (see fcfc_gpu_last_error). */
fcfc_gpu_catalog *fcfc_gpu_catalog_create(const void *x, const void *y, const void *z,
    const void *x2sum, const void *w, size_t n, int is_float, double rescale, int sumsq_arith);
void fcfc_gpu_catalog_destroy(fcfc_gpu_catalog *cat);
size_t fcfc_gpu_catalog_size(const fcfc_gpu_catalog *cat);
/* Sum of weights accumulated in double (data->wt, build_tree.c:133-140); n if unweighted. */
double fcfc_gpu_catalog_wsum(const fcfc_gpu_catalog *cat);

/* Streamed ingest: the same catalogue built chunk by chunk while the host is still reading the file, replacing the
 * "read everything, then index" order of tree_create (fcfc/2pt_box/build_tree.c:84-156) around the chunk loop of the
 * readers (io/read_ascii.c:750-950: fread a chunk, parse its lines, append to the columns).
 *   _begin:  n_hint = expected number of rows (0: unknown, the device columns grow geometrically); with_weight != 0 when
 *            every chunk carries a weight column.
 *   _append: n rows of `real` per column (host pointers, unrescaled or rescaled as the caller prefers); the rows are
 *            copied into pinned staging memory before the call returns -- the caller may reuse or realloc() its arrays
 *            at once -- and travel to the device asynchronously, overlapping whatever the host does next.  n = 0 is a
 *            no-op.  After an error the stream only accepts _abort / _finish (which then fails).
 *   _finish: waits for the transfers, runs the one pass over the columns that fcfc_gpu_catalog_create runs (rescale in
 *            `real` when rescale != 1, x^2+y^2+z^2 in the order `sumsq_arith` selects when >= 0, bounding box, sum of
 *            weights, non-finite check), replicates the catalogue on the other devices in use and returns the handle
 *            (NULL on failure).  The stream object is released either way.
 *   _abort:  releases a stream without building a catalogue. */
typedef struct fcfc_gpu_catalog_stream fcfc_gpu_catalog_stream;
fcfc_gpu_catalog_stream *fcfc_gpu_catalog_stream_begin(size_t n_hint, int is_float, int with_weight);
int fcfc_gpu_catalog_stream_append(fcfc_gpu_catalog_stream *st, const void *x, const void *y, const void *z,
    const void *w, size_t n);
size_t fcfc_gpu_catalog_stream_size(const fcfc_gpu_catalog_stream *st);
fcfc_gpu_catalog *fcfc_gpu_catalog_stream_finish(fcfc_gpu_catalog_stream *st, double rescale, int sumsq_arith);
void fcfc_gpu_catalog_stream_abort(fcfc_gpu_catalog_stream *st);

/* count_pairs(): full count over all devices in use; the per-device partial histograms are
 * combined with one all-reduce.  Exactly one of cnt_i / cnt_d is written (ntot entries,
 * overwritten, not accumulated): cnt_i when !withwt, cnt_d when withwt. */
int fcfc_gpu_count(fcfc_gpu_catalog *cat1, fcfc_gpu_catalog *cat2, const fcfc_gpu_bins *bins,
    int isauto, int withwt, int64_t *cnt_i, double *cnt_d);

/* One shard of the same count, for launchers that run one process per GPU: the work items of the
 * primary catalogue (cat1) are split into `nparts` cost-balanced parts and part `part` is counted
 * on the first device in use.  The partial histogram is written to host memory (cnt_i / cnt_d)
 * and, if dev_hist is not NULL, left in device memory as ntot x 8 bytes (int64 or double) so the
 * caller can all-reduce it in place (ncclAllReduce / torch.distributed) without a host round trip. */
int fcfc_gpu_count_partial(fcfc_gpu_catalog *cat1, fcfc_gpu_catalog *cat2, const fcfc_gpu_bins *bins,
    int isauto, int withwt, int part, int nparts, int64_t *cnt_i, double *cnt_d, void *dev_hist);

int fcfc_gpu_get_stats(fcfc_gpu_stats *out);

/* Tuning / diagnostic options of the engine (tests and A/B measurements; the defaults are what production uses).
 * The counting path never reads the environment: FCFC_GPU_TUNE="name=value,..." is parsed once by fcfc_gpu_init,
 * and this call changes a value explicitly.  Names: k (cells of reach / k), nsplit, items_per_warp, cost_bits,
 * no_subsort, no_table_math, no_hist_copies, qdepth, qkeep, force_generic, global_hist, no_dense, no_prefilter,  force_prefilter, no_df, no_classify,
 * sorted_copies, nccl; "defaults" restores everything.  Returns FCFC_GPU_ERR_ARG for an unknown name. */
int fcfc_gpu_set_option(const char *name, long value);

/* Optional host helper for callers that do not link the FCFC host: builds the rescale factor,
 * rescaled edges and lookup tables exactly as cf_setup does (fcfc/2pt_box/setup_cf.c:385-531,
 * fcfc/2pt/setup_cf.c:432-503, util/create_lut.c:58-140).  `linear` != 0 uses smin/ds (and
 * pmin/dpi); otherwise explicit edges sedge[ns+1] / pedge[np+1].  The returned object owns the
 * arrays `bins` points to; free with fcfc_gpu_bins_free.  Pure host code, no device needed. */
typedef struct fcfc_gpu_bins_owner fcfc_gpu_bins_owner;
fcfc_gpu_bins_owner *fcfc_gpu_bins_create(int periodic, int is_float, int bintype, int linear,
    double smin, double ds, double pmin, double dpi, const double *sedge, int ns,
    const double *pedge, int np, int nmu, const double box[3], int with_mu_one, int arith);
const fcfc_gpu_bins *fcfc_gpu_bins_get(const fcfc_gpu_bins_owner *o);
double fcfc_gpu_bins_rescale(const fcfc_gpu_bins_owner *o);
void fcfc_gpu_bins_free(fcfc_gpu_bins_owner *o);

/* (ra [deg], dec [deg], redshift) -> comoving Cartesian coordinates, in place (host or device pointers), replacing
 * cnvt_coord_integr (fcfc/2pt/cnvt_coord.c:321-337) for w = -1 dark energy: Legendre-Gauss quadrature of the given order
 * with the caller's abscissas / weights (the order >> 1 non-zero ones, then the x = 0 weight of an odd order: the
 * reference's legauss_x / legauss_w + LEGAUSS_IDX(order), math/legauss.h:49-59).  The comoving distance is bit-identical
 * to the host's; sin / cos are the CUDA math library's (<= 2 ulp each; a coordinate is within 5 ulp of the host's, 85 % identical),
 * which is why the shim only uses this on request. */
int fcfc_gpu_cnvt_coord(void *x, void *y, void *z, size_t n, int is_float, double omega_m, double omega_l,
    double omega_k, int order, const double *gl_x, const double *gl_w);

/* Measured FP32 instruction issue peak of device 0 (lane-instructions per second of a dependent-
 * free FFMA stream) -- the denominator of the pair-evaluation roofline. */
double fcfc_gpu_measure_fp32_peak(double *sm_clock_mhz_out);
/* The same for FP64 (DFMA stream): the denominator for the double-precision kernels. */
double fcfc_gpu_measure_fp64_peak(void);
/* Shared-memory histogram increments per second (red.shared.add.u32 lanes on pseudo-random bins of a 4800-counter
 * histogram, whole device): every accepted pair costs one, so in-range pairs / this rate is the time the histogram
 * update alone needs -- the second bound bench.py states next to the FP32 issue roofline (`roofline.ceiling`). */
double fcfc_gpu_measure_smem_atomic_peak(void);
/* Diagnostics: the fixed-point scales 2^ks, 2^km the counting kernels use for their computed s and mu bins
 * (host arithmetic only; the CPU tests check the error budget behind them). */
void fcfc_gpu_fastbin_scales(int ns, int nmu, int periodic, int *ks, int *km);
/* Diagnostics: limits of the division-free pre-tests of the survey (s_perp, pi) metric, rounded up to the build's
 * real type: out[0] searched sphere, out[1] padded p2max, out[2] padded s2max. */
void fcfc_gpu_survey_pretest_limits(double s2max, double p2max, int is_float, double out[3]);
/* Diagnostics: padded limits of the float pre-filter of the double-precision kernels (see engine.cu for the error
 * budget); returns 0 (filter not usable), 1 (sphere / box tests) or 2 (survey (s_perp,pi): cylinder tests as well). */
int fcfc_gpu_prefilter_limits(int periodic, int bintype, double s2max, double pmax, double maxabs, double smax_sq,
                              double smin_sq, double out[4]);
/* Diagnostics: error budget of the double-precision-at-float-speed kernel (count_kernel_df.cuh) for ns unit-width s bins,
 * nmu mu bins (1: isotropic), the squared maximum separation and the largest cell size: fixed-point scales 2^ks, 2^km, the
 * padded float range limit, the squared separation below which (s,mu) pairs always take the exact path, and the expected
 * fraction of flagged pairs; returns 1 when the kernel is usable. */
int fcfc_gpu_df_budget(int ns, int nmu, double s2max, double cs_max, int *ks, int *km, double *d2lim, double *s1sq,
                       double *flagged);
/* Diagnostics: limits of the classification of staged points against the tile's bounding box in the single-precision
 * kernels (count_kernel_cl.cuh) for the squared maximum separation and the largest coordinate magnitude (image shifts
 * included): out[0] = squared distance to the nearest point of the box above which a point is dropped, out[1] = squared
 * distance to the farthest corner below which it is binned in place. */
void fcfc_gpu_classify_limits(double s2max, double maxabs, float out[2]);
/* Diagnostics: the neighbour-cell stencil (rows (dx, dy, dz_lo, dz_hi)) and the dense sub-range of every row for cells
 * of size cs[3], a spherical reach r2 (squared) and the maximum separation s2max (squared); returns the row count. */
int fcfc_gpu_debug_stencil(const double cs[3], double r2, double s2max, int half, int *rows_out, int *inside_out, int max_rows);

#ifdef __cplusplus
}
#endif
#endif
