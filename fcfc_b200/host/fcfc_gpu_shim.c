/*******************************************************************************
* fcfc_b200/host/fcfc_gpu_shim.c -- C99 host shim: FCFC's counting seam on top of libfcfc_b200.so.
*
* Drop-in for three translation units of cheng-zhao/FCFC v1.0.1:
*     src/tree/kdtree.c, src/tree/balltree.c   (the spatial index; qselect.c / pca.c become unused)
*     src/fcfc/2pt_box/count_func.c  or  src/fcfc/2pt/count_func.c   (the counting core)
* Everything else of the host -- load_conf, cf_setup, the catalogue readers, cnvt_coord,
* tree_create / tree_destroy (build_tree.c), eval_cf, save_res -- is compiled UNMODIFIED.  The
* reference's tree_create() keeps reading, rescaling and weighting the catalogue on the host
* (build_tree.c:84-156) and then calls create_kdtree()/create_balltree(): here that call uploads
* the arrays to the GPU and returns the catalogue handle in place of the tree root.  count_pairs()
* (called from eval_cf.c:119,142) marshals the slice of `CF` the counter reads into fcfc_gpu_bins.
*
* Compile once per program with the reference's include paths (see integration/Makefile):
*     -Isrc/fcfc/2pt_box (or 2pt)  -Isrc/util -Isrc/io -Isrc/lib -Isrc/math -Isrc/tree  -Iinclude
* The same -DSINGLE_PREC / -DWITH_MU_ONE switches as the rest of the host apply.  -DOMP/-DWITH_SIMD
* may stay enabled for the readers; -DMPI is not supported (one process drives all GPUs).
*
* Environment: FCFC_GPU_ARITH=fma|scalar selects the evaluation order (default: the one this host build would run in the
* reference: FMA order of the AVX kernels when compiled with -DWITH_SIMD, scalar order otherwise);
* FCFC_GPU_DEVICES=0,1,... restricts the devices; FCFC_GPU_VERBOSE=1 prints engine diagnostics.
*******************************************************************************/
#ifndef _POSIX_C_SOURCE
#define _POSIX_C_SOURCE 200809L         /* strdup, strtok_r (the reference's Makefile sets the same) */
#endif
#include "define.h"
#include "eval_cf.h"
#include "count_func.h"
#include "kdtree.h"
#include "balltree.h"
#include "fcfc_gpu.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef MPI
  #error the GPU engine replaces MPI: build the host without -DMPI
#endif

#ifdef SINGLE_PREC
  #define SHIM_IS_FLOAT 1
#else
  #define SHIM_IS_FLOAT 0
#endif
#ifdef FCFC_METRIC_PERIODIC
  #define SHIM_PERIODIC 1
#else
  #define SHIM_PERIODIC 0
#endif

static int shim_ready = 0, shim_verbose = 0;

static double shim_now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

static void shim_init(void) {
  if (shim_ready) return;
  int dev[64], ndev = 0;
  const char *s = getenv("FCFC_GPU_DEVICES");
  if (s) {
    char *copy = strdup(s), *save = NULL;
    for (char *t = strtok_r(copy, ",", &save); t && ndev < 64; t = strtok_r(NULL, ",", &save)) dev[ndev++] = atoi(t);
    free(copy);
  }
  const char *v = getenv("FCFC_GPU_VERBOSE");
  shim_verbose = v ? atoi(v) : 0;
  int n = fcfc_gpu_init(ndev, ndev ? dev : NULL, shim_verbose);
  if (n <= 0) {
    P_ERR("GPU engine initialisation failed: %s\n", fcfc_gpu_last_error());
    exit(FCFC_ERR_TREE);
  }
  shim_ready = 1;
}

/* Upload one catalogue; the handle stands in for the tree root (it is only ever passed back to
 * count_pairs() and *_free()).  The coordinates are already rescaled, the survey build's 4th
 * column already holds x^2+y^2+z^2 (2pt/build_tree.c:35-133). */
static void *shim_upload(real *x[static FCFC_XDIM], real *w, const size_t ndata, size_t *nnode) {
  shim_init();
  const void *s = NULL;
#if FCFC_XDIM > 3
  s = x[3];
#endif
  const double t0 = shim_now();
  fcfc_gpu_catalog *cat = fcfc_gpu_catalog_create(x[0], x[1], x[2], s, w, ndata, SHIM_IS_FLOAT, 1.0, -1);
  if (shim_verbose && cat) fprintf(stderr, "[fcfc_gpu] catalogue of %zu objects resident on every device: %.1f ms\n", ndata, 1e3 * (shim_now() - t0));
  if (!cat) {
    P_ERR("failed to build the GPU cell list: %s\n", fcfc_gpu_last_error());
    return NULL;
  }
  if (nnode) *nnode = 1;
  return cat;
}

KDT *create_kdtree(real *x[static FCFC_XDIM], real *w, const size_t ndata,
#ifdef FCFC_METRIC_PERIODIC
    const real bsize[static 3],
#endif
    const size_t nleaf, size_t *nnode) {
#ifdef FCFC_METRIC_PERIODIC
  (void) bsize;
#endif
  (void) nleaf;
  return (KDT *) shim_upload(x, w, ndata, nnode);
}

void kdtree_free(KDT *root) { fcfc_gpu_catalog_destroy((fcfc_gpu_catalog *) root); }

BLT *create_balltree(real *x[static FCFC_XDIM], real *w, const size_t ndata,
#ifdef FCFC_METRIC_PERIODIC
    const real bsize[static 3],
#endif
    const size_t nleaf, size_t *nnode) {
#ifdef FCFC_METRIC_PERIODIC
  (void) bsize;
#endif
  (void) nleaf;
  return (BLT *) shim_upload(x, w, ndata, nnode);
}

void balltree_free(BLT *root) { fcfc_gpu_catalog_destroy((fcfc_gpu_catalog *) root); }

/* count_pairs(): same signature and semantics as count_func.h:54 (auto pairs once, raw counts into
 * cnt[].i or weighted sums into cnt[].d, layout s + p * ns).  eval_cf.c ignores the return value, so
 * errors terminate the program like the reference's own fatal paths (count_func.c:77-95). */
int count_pairs(const void *tree1, const void *tree2, CF *cf, COUNT *cnt,
    const bool isauto, const bool withwt) {
  fcfc_gpu_bins b;
  memset(&b, 0, sizeof b);
  b.bintype = cf->bintype;
  b.periodic = SHIM_PERIODIC;
  b.is_float = SHIM_IS_FLOAT;
  b.tabtype = cf->tabtype;
  b.ns = cf->ns; b.np = cf->np; b.nmu = cf->nmu;
  b.swidth = cf->swidth; b.pwidth = cf->pwidth;
#ifdef WITH_MU_ONE
  b.with_mu_one = 1;
#endif
  /* evaluation order: that of the code path this host build would run in the reference -- the AVX/FMA kernels of a
   * -DWITH_SIMD build (metric_common.c:377-534), the scalar sequence otherwise (:140-235); FCFC_GPU_ARITH overrides */
  const char *ar = getenv("FCFC_GPU_ARITH");
#if defined(WITH_SIMD)
  b.arith = (ar && !strcmp(ar, "scalar")) ? FCFC_GPU_ARITH_SCALAR : FCFC_GPU_ARITH_FMA;
#else
  b.arith = (ar && !strcmp(ar, "fma")) ? FCFC_GPU_ARITH_FMA : FCFC_GPU_ARITH_SCALAR;
#endif
  b.s2bin = cf->s2bin;
#ifdef FCFC_METRIC_PERIODIC
  b.pbin = cf->pbin;
  for (int i = 0; i < 3; i++) b.bsize[i] = cf->bsize[i];
#else
  b.pbin = cf->p2bin;
#endif
  b.stab = cf->stab; b.ptab = cf->ptab; b.mutab = cf->mutab;
  const double t0 = shim_now();
  int e = fcfc_gpu_count((fcfc_gpu_catalog *) tree1, (fcfc_gpu_catalog *) tree2, &b, isauto, withwt,
      withwt ? NULL : (int64_t *) cnt, withwt ? (double *) cnt : NULL);
  if (shim_verbose && !e) {
    fcfc_gpu_stats st;
    fcfc_gpu_get_stats(&st);
    fprintf(stderr, "[fcfc_gpu] count step: %.1f ms wall (cell lists %.1f ms, counting kernel %.1f ms on the slowest device, "
        "%.4g pair evaluations, grid %d x %d x %d)\n", 1e3 * (shim_now() - t0), st.ms_sort, st.ms_count, (double) st.pair_evals,
        st.ncell[0], st.ncell[1], st.ncell[2]);
  }
  if (e) {
    P_ERR("GPU pair counting failed (%d): %s\n", e, fcfc_gpu_last_error());
    exit(FCFC_ERR_CF);
  }
  return 0;
}


#ifndef FCFC_METRIC_PERIODIC
/* Coordinate conversion on the device (opt-in: FCFC_GPU_CNVT=1).  The survey host is linked with
 * -Wl,--wrap=cnvt_coord (integration/Makefile): cnvt_coord() of the unmodified cnvt_coord.c stays the default -- its
 * libm sin / cos are what the parity tests are pinned on -- and this wrapper takes over the Legendre-Gauss branch for
 * w = -1 dark energy on request.  The integration order is chosen exactly as the host does (convergence of the quadrature
 * at FCFC_INT_NUM_ZSP sample redshifts, cnvt_coord.c:262-305), with the host's own abscissas and weights. */
#include "cnvt_coord.h"
#include "legauss.h"
#include <float.h>
#include <limits.h>
#include <math.h>

int __real_cnvt_coord(const CONF *conf, real *x[static 3], const size_t ndata, COORD_CNVT *coord);

static double shim_integrand(const double om, const double ol, const double ok, const double z) {
  const double z1 = z + 1, z2 = z1 * z1;
  double d = om * z2 * z1;
  if (ok) d += ok * z2;
  d += ol;
  return SPEED_OF_LIGHT * 0.01 / sqrt(d);
}

static double shim_legauss(const double om, const double ol, const double ok, const int order, const double z) {
  const double zp = z * 0.5;
  double sum = 0;
  int i = LEGAUSS_IDX(order);
  for (; i < LEGAUSS_IDX(order) + LEGAUSS_LEN_NONZERO(order); i++) {
    const double a = zp * (1 + legauss_x[i]), b = zp * (1 - legauss_x[i]);
    sum += legauss_w[i] * (shim_integrand(om, ol, ok, a) + shim_integrand(om, ol, ok, b));
  }
  if (order & 1) sum += legauss_w[i] * shim_integrand(om, ol, ok, zp);
  return sum * zp;
}

int __wrap_cnvt_coord(const CONF *conf, real *x[static 3], const size_t ndata, COORD_CNVT *coord) {
  const char *on = getenv("FCFC_GPU_CNVT");
  if (!on || !atoi(on) || !conf || conf->fcnvt || conf->dew != -1 || !x || !x[0] || !x[1] || !x[2] || !ndata)
    return __real_cnvt_coord(conf, x, ndata, coord);
  shim_init();
  double zmin = DBL_MAX, zmax = -DBL_MAX;
  for (size_t i = 0; i < ndata; i++) {
    const double z = x[2][i];
    if (z < 0) return __real_cnvt_coord(conf, x, ndata, coord);          /* (the host reports the invalid redshift) */
    if (z > zmax) zmax = z;
    if (z < zmin) zmin = z;
  }
  int order = 0;
  for (int k = 0; k < FCFC_INT_NUM_ZSP; k++) {
    const double z = zmin + k * (zmax - zmin) / (FCFC_INT_NUM_ZSP - 1);
    double oint, nint = 0;
    int n = LEGAUSS_MIN_ORDER - 1;
    do {
      if (n > LEGAUSS_MAX_ORDER) return __real_cnvt_coord(conf, x, ndata, coord);
      oint = nint;
      nint = shim_legauss(conf->omega_m, conf->omega_l, conf->omega_k, ++n, z);
    } while (fabs(nint - oint) > nint * conf->ecnvt);
    if (order < n) order = n;
  }
  if (order > LEGAUSS_MAX_ORDER) return __real_cnvt_coord(conf, x, ndata, coord);
  const double t0 = shim_now();
  const int e = fcfc_gpu_cnvt_coord(x[0], x[1], x[2], ndata, SHIM_IS_FLOAT, conf->omega_m, conf->omega_l, conf->omega_k, order,
      legauss_x + LEGAUSS_IDX(order), legauss_w + LEGAUSS_IDX(order));
  if (e) {
    P_ERR("coordinate conversion on the GPU failed (%d)\n", e);
    return FCFC_ERR_CNVT;
  }
  if (shim_verbose) fprintf(stderr, "[fcfc_gpu] %zu objects converted on the device (Legendre-Gauss order %d): %.1f ms\n", ndata, order, 1e3 * (shim_now() - t0));
  if (conf->verbose) printf("  Coordinates converted using Legendre-Gauss integration with order %d\n", order);
  return 0;
}
#endif
