/*******************************************************************************
* oracle/ref_driver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
*
* Timing / parity driver for the *unmodified* FCFC reference.  It is compiled
* against the reference's headers and linked with the reference's own objects
* (everything except its main(), see oracle/Makefile) and then follows the call
* sequence of the reference's main program and pair-count loop
*   load_conf -> cf_setup            (reference src/fcfc/2pt_box/fcfc.c:34-44)
*   tree_create -> count_pairs       (reference src/fcfc/2pt_box/eval_cf.c:92-142)
* but wraps `count_pairs` alone in clock_gettime(), and dumps the raw (un-doubled)
* counts together with the bin tables the reference built, so that
*   - tests can check the CUDA engine and the CPU restatement (oracle/fcfc_oracle.c)
*     against the real reference on in-memory catalogues, and
*   - bench.py can report the CPU baseline for exactly the replaced step.
*
* Catalogues with the suffix ".fbin" are read from a trivial binary container
* instead of ASCII: the linker option --wrap=read_ascii_data redirects the
* reference's call (src/fcfc/2pt_box/build_tree.c:87) to the function below,
* which honours the reference reader's allocation contract
* (src/io/read_ascii.c:1007-1020: n + FCFC_NUM_REAL zero-padded entries for SIMD
* builds).  Every other step (rescaling, sum of squares, weights, tree build) is
* executed by the reference's own tree_create().
*
* Usage: ref_driver_{box,svy} OUT.drv [FCFC command line options...]
*******************************************************************************/
#include "define.h"
#include "load_conf.h"
#include "eval_cf.h"
#include "build_tree.h"
#include "count_func.h"
#include "read_file.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <time.h>

#ifndef FCFC_NUM_REAL
  #define DRV_PAD 0
#else
  #define DRV_PAD FCFC_NUM_REAL
#endif

int __real_read_ascii_data(const char *fname, const size_t skip,
    const char comment, const char *fmtr, char *const *rcol_ids,
    const int nrcol, const char *sel, real ***rout, size_t *num,
    const int verb);

/* Binary container: "FCFCBIN1", uint64 n, uint32 ncol, uint32 0, then ncol
 * column-major double arrays of length n (x, y, z[, w]). */
int __wrap_read_ascii_data(const char *fname, const size_t skip,
    const char comment, const char *fmtr, char *const *rcol_ids,
    const int nrcol, const char *sel, real ***rout, size_t *num,
    const int verb) {
  size_t len = strlen(fname);
  if (len < 5 || strcmp(fname + len - 5, ".fbin"))
    return __real_read_ascii_data(fname, skip, comment, fmtr, rcol_ids, nrcol,
        sel, rout, num, verb);

  FILE *fp = fopen(fname, "rb");
  if (!fp) { P_ERR("cannot open `%s'\n", fname); return FCFC_ERR_FILE; }
  char magic[8]; uint64_t n; uint32_t ncol, pad;
  if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "FCFCBIN1", 8) ||
      fread(&n, 8, 1, fp) != 1 || fread(&ncol, 4, 1, fp) != 1 ||
      fread(&pad, 4, 1, fp) != 1 || (int) ncol < nrcol) {
    P_ERR("invalid binary catalogue `%s'\n", fname); fclose(fp);
    return FCFC_ERR_FILE;
  }
  real **res = malloc(sizeof(real *) * nrcol);
  double *buf = malloc(sizeof(double) * (n ? n : 1));
  if (!res || !buf) return FCFC_ERR_MEMORY;
  for (int c = 0; c < nrcol; c++) {
    if (!(res[c] = calloc(n + DRV_PAD + 1, sizeof(real)))) return FCFC_ERR_MEMORY;
    if (fread(buf, sizeof(double), n, fp) != n) {
      P_ERR("truncated binary catalogue `%s'\n", fname); fclose(fp);
      return FCFC_ERR_FILE;
    }
    for (size_t i = 0; i < n; i++) res[c][i] = (real) buf[i];
  }
  free(buf);
  fclose(fp);
  if (verb) printf("  %zu objects read from the binary catalogue\n", (size_t) n);
  *num = n;
  *rout = res;
  return 0;
}

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void wr(FILE *fp, const void *p, size_t sz) {
  if (sz && fwrite(p, 1, sz, fp) != sz) { perror("fwrite"); exit(3); }
}
static void wr_i32(FILE *fp, int32_t v) { wr(fp, &v, 4); }
static void wr_u64(FILE *fp, uint64_t v) { wr(fp, &v, 8); }
static void wr_f64(FILE *fp, double v) { wr(fp, &v, 8); }
static void wr_reals(FILE *fp, const real *x, size_t n) {
  wr_u64(fp, x ? n : 0);
  if (x) for (size_t i = 0; i < n; i++) wr_f64(fp, (double) x[i]);
}
/* Number of entries of a lookup table, as in src/util/create_lut.c:62-64,105-107. */
static size_t tab_len(const real *bins, int num, int tabtype) {
  long min = bins[0];
  long max = (tabtype == FCFC_LOOKUP_TYPE_INT) ? (long) bins[num]
      : (long) ceil(bins[num]);
  return (size_t) (max - min);
}

int main(int argc, char *argv[]) {
  if (argc < 2) {
    fprintf(stderr, "Usage: %s OUT.drv [FCFC options]\n", argv[0]);
    return 2;
  }
  const char *fout = argv[1];
  PARA para;
  para_init(&para);

  CONF *conf = load_conf(argc - 1, argv + 1, &para);
  if (!conf) { P_EXT("failed to load configuration parameters\n"); return 1; }
  CF *cf = cf_setup(conf, &para);
  if (!cf) { P_EXT("failed to initialise the calculator\n"); return 1; }

  FILE *fp = fopen(fout, "wb");
  if (!fp) { perror(fout); return 1; }
  wr(fp, "FCFCDRV1", 8);
#ifdef SINGLE_PREC
  wr_i32(fp, 1);
#else
  wr_i32(fp, 0);
#endif
#ifdef DRV_BOX
  wr_i32(fp, 1);
#else
  wr_i32(fp, 0);
#endif
  wr_i32(fp, cf->bintype); wr_i32(fp, cf->tabtype);
  wr_i32(fp, cf->ns); wr_i32(fp, cf->np); wr_i32(fp, cf->nmu);
  wr_i32(fp, cf->swidth); wr_i32(fp, cf->pwidth);
  wr_i32(fp, cf->npc); wr_i32(fp, cf->ncat); wr_i32(fp, para.nthread);
  wr_u64(fp, cf->ntot);
  wr_f64(fp, (double) cf->rescale);
#ifdef DRV_BOX
  for (int i = 0; i < 3; i++) wr_f64(fp, (double) cf->bsize[i]);
  const real *pedge = cf->pbin;
#else
  for (int i = 0; i < 3; i++) wr_f64(fp, 0);
  const real *pedge = cf->p2bin;
#endif
  wr_reals(fp, cf->s2bin, cf->ns + 1);
  const int spi = (cf->bintype == FCFC_BIN_SPI);
  wr_reals(fp, spi ? pedge : NULL, cf->np + 1);
  size_t nst = tab_len(cf->s2bin, cf->ns, cf->tabtype);
  size_t esz = (cf->swidth == FCFC_LOOKUP_TABLE_W8) ? 1 : 2;
  wr_u64(fp, nst); wr(fp, cf->stab, nst * esz);
  if (spi) {
    size_t npt = tab_len(pedge, cf->np, cf->tabtype);
    esz = (cf->pwidth == FCFC_LOOKUP_TABLE_W8) ? 1 : 2;
    wr_u64(fp, npt); wr(fp, cf->ptab, npt * esz);
  }
  else wr_u64(fp, 0);
  if (cf->bintype == FCFC_BIN_SMU) {
    wr_u64(fp, (uint64_t) cf->nmu * cf->nmu);
    wr(fp, cf->mutab, (size_t) cf->nmu * cf->nmu);
  }
  else wr_u64(fp, 0);

  void **tree = calloc(cf->ncat, sizeof(void *));
  for (int i = 0; i < cf->npc; i++) {
    if (!cf->comp_pc[i]) { P_EXT("driver needs every pair count computed\n"); return 1; }
    int cat[2] = {cf->pc_idx[0][i], cf->pc_idx[1][i]};
    double t_tree = 0;
    for (int j = 0; j < 2; j++) {
      if (!tree[cat[j]]) {
        double t0 = now();
        if (!(tree[cat[j]] = tree_create(conf, cf, cat[j]))) {
          P_EXT("tree_create failed\n"); return 1;
        }
        t_tree += now() - t0;
      }
    }
    const bool isauto = (cat[0] == cat[1]);
    printf("Counting %c%c pairs ...\n", cf->label[cat[0]], cf->label[cat[1]]);
    fflush(stdout);
    double t0 = now();
    count_pairs(tree[cat[0]], tree[cat[1]], cf, cf->cnt[i], isauto, cf->wt[i]);
    double t_count = now() - t0;
    printf("  count_pairs: %.6f s (tree build/IO %.6f s)\n", t_count, t_tree);

    char lab[2] = {cf->label[cat[0]], cf->label[cat[1]]};
    wr(fp, lab, 2);
    wr_i32(fp, isauto); wr_i32(fp, cf->wt[i]);
    wr_f64(fp, t_tree); wr_f64(fp, t_count);
    wr_u64(fp, cf->data[cat[0]].n); wr_u64(fp, cf->data[cat[1]].n);
    wr_f64(fp, cf->data[cat[0]].wt); wr_f64(fp, cf->data[cat[1]].wt);
    wr(fp, cf->cnt[i], sizeof(COUNT) * cf->ntot);
  }
  fclose(fp);
  /* The process exits here; the OS reclaims the catalogues and trees. */
  return 0;
}
