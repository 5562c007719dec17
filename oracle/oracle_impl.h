/*******************************************************************************
* oracle/oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
*
* Type-generic body of the CPU restatement; included twice by fcfc_oracle.c with
*   REAL = double, SFX(x) = x##_d      (reference default build)
*   REAL = float,  SFX(x) = x##_f      (reference -DSINGLE_PREC build)
* Every function cites the reference file:line it restates (paths relative to
* /root/reference/src).  The counting loop is a plain O(n1*n2) brute force: it
* shares nothing with the GPU engine's cell list, tiling or scheduling.
*******************************************************************************/

#ifndef REAL
  #error include from fcfc_oracle.c
#endif

typedef struct {
  int bintype;          /* 0 iso, 1 (s,mu), 2 (s_perp,pi): fcfc/2pt_box/define.h:45-47 */
  int periodic;         /* 1: FCFC_2PT_BOX metric; 0: FCFC_2PT (survey) metric          */
  int tabtype;          /* 0 integer table, 1 hybrid table: define.h:49-50              */
  int ns, np, nmu;
  int swidth, pwidth;   /* 0: uint8_t entries, 1: uint16_t entries: define.h:75-76      */
  int with_mu_one;      /* -DWITH_MU_ONE                                                */
  int arith;            /* 0: scalar formulas (metric_common.c scalar functions),
                           1: FMA order of the AVX-512 vector functions                 */
  REAL rescale;
  REAL bsize[3];        /* rescaled box (periodic only)                                 */
  REAL *s2bin;          /* ns+1 squared s (s_perp) edges, rescaled                      */
  REAL *pbin;           /* np+1 edges: pi (box) or pi^2 (survey), rescaled              */
  void *stab; size_t nstab;
  void *ptab; size_t nptab;
  uint8_t *mutab;       /* nmu*nmu                                                      */
} SFX(oracle_bins);

/* util/create_lut.c:58-91 -- integer-edge lookup table. */
static void *SFX(lut_int)(const REAL *bins, int num, int *width, size_t *len) {
  const long min = (long) bins[0];
  const long max = (long) bins[num];
  const long ntab = max - min;
  *width = (num <= UINT8_MAX) ? 0 : 1;          /* fcfc/2pt_box/setup_cf.c:241-251 */
  if (num > UINT16_MAX || ntab < 0) return NULL;
  uint8_t *t8 = NULL; uint16_t *t16 = NULL;
  if (*width == 0) t8 = calloc(ntab + 8, 1); else t16 = calloc(ntab + 8, 2);
  int n = 1;
  for (long i = 0; i < ntab; i++) {
    while (!((REAL) (i + min) < bins[n])) { if (++n > num) { free(t8); free(t16); return NULL; } }
    if (t8) t8[i] = (uint8_t) (n - 1); else t16[i] = (uint16_t) (n - 1);
  }
  *len = (size_t) ntab;
  return t8 ? (void *) t8 : (void *) t16;
}

/* util/create_lut.c:102-140 -- hybrid table for arbitrary edges: entries < num are
 * final, entries >= num encode "start at bin (entry - num) and walk down". */
static void *SFX(lut_hybrid)(const REAL *bins, int num, int *width, size_t *len) {
  const long min = (long) bins[0];
  const long max = (long) ceil((double) bins[num]);
  const long ntab = max - min;
  *width = (num <= UINT8_MAX / 2) ? 0 : 1;      /* setup_cf.c:263-273 */
  if (num > UINT16_MAX / 2 || ntab < 0) return NULL;
  uint8_t *t8 = NULL; uint16_t *t16 = NULL;
  if (*width == 0) t8 = calloc(ntab + 8, 1); else t16 = calloc(ntab + 8, 2);
#define ORC_SET(i, v) do { if (t8) t8[i] = (uint8_t) (v); else t16[i] = (uint16_t) (v); } while (0)
  int n = 1;
  long edge = (long) bins[n];
  for (long i = 0; i < ntab; i++) {
    if (i + min < edge) ORC_SET(i, n - 1);
    else {
      while (++n <= num) {
        edge = (long) bins[n];
        if (i + min < edge) { ORC_SET(i, num + n - 1); break; }
      }
      if (n > num) { for (; i < ntab; i++) ORC_SET(i, num * 2); break; }
    }
  }
#undef ORC_SET
  *len = (size_t) ntab;
  return t8 ? (void *) t8 : (void *) t16;
}

static size_t SFX(gcd)(size_t a, size_t b) { while (b) { size_t t = b; b = a % b; a = t; } return a; }

/* setup_cf.c:308-333 (least_fac2) and :347-376 (least_fac4): smallest factor turning the
 * arguments into integers (tested with up to 4 decimal digits). */
static REAL SFX(least_fac)(const REAL *v, int nv) {
  REAL a[4];
  for (int k = 0; k < nv; k++) { a[k] = v[k]; if (a[k] < 0 || a[k] > ORC_MAXINT) return 0; }
  if (nv == 2) {
    if (a[0] == 0) return 1 / a[1];
    if (a[1] == 0) return 1 / a[0];
  }
  size_t i = 0, ifac = 1;
  do {
    int ok = 1;
    for (int k = 0; k < nv; k++) if (!(ORC_ABS(ORC_ROUND(a[k]) - a[k]) < ORC_TOL)) ok = 0;
    if (ok) break;
    for (int k = 0; k < nv; k++) a[k] *= 10;
    ifac *= 10;
    i++;
  } while (i <= 4);
  if (i == 5) return 0;
  for (int k = 0; k < nv; k++) if (a[k] > ORC_MAXINT) return 0;
  size_t ig;
  if (nv == 2) ig = SFX(gcd)((size_t) ORC_ROUND(a[0]), (size_t) ORC_ROUND(a[1]));
  else {
    /* the reference rounds only the first non-zero operand (setup_cf.c:365-372) */
    ig = 0;
    for (int k = 0; k < nv; k++) {
      if (a[k] != 0) ig = (ig == 0) ? (size_t) ORC_ROUND(a[k])
          : SFX(gcd)(ig, (k == 0) ? (size_t) ORC_ROUND(a[k]) : (size_t) a[k]);
    }
    if (ig == 0) ig = 1;
  }
  return ifac / (REAL) ig;
}

/* fcfc/2pt_box/setup_cf.c:385-531,596-679 and fcfc/2pt/setup_cf.c:432-503: bin edges ->
 * rescale factor, rescaled (squared) edges, lookup tables.  `sedge`/`pedge` are the
 * unrescaled edges (ns+1 / np+1 values); `lin` tells whether they came from
 * SEP_BIN_MIN/SIZE (and PI_BIN_MIN/SIZE), in which case smin/ds/pmin/dpi are used for
 * the common-factor search exactly as the reference does. */
SFX(oracle_bins) *SFX(oracle_setup)(int periodic, int bintype, int lin,
    double smin_d, double ds_d, double pmin_d, double dpi_d,
    const double *sedge, int ns, const double *pedge, int np, int nmu,
    const double *box, int with_mu_one, int arith) {
  SFX(oracle_bins) *b = calloc(1, sizeof *b);
  b->bintype = bintype; b->periodic = periodic; b->ns = ns; b->np = np; b->nmu = nmu;
  b->with_mu_one = with_mu_one; b->arith = arith;
  REAL *sbin = malloc(sizeof(REAL) * (ns + 1));
  b->s2bin = malloc(sizeof(REAL) * (ns + 1));
  const REAL smin = smin_d, ds = ds_d, pmin = pmin_d, dpi = dpi_d;
  /* setup_cf.c:612: edges are formed in double from the double-typed configuration values */
  for (int i = 0; i <= ns; i++) sbin[i] = lin ? (REAL) (smin_d + ds_d * i) : (REAL) sedge[i];
  REAL *pb = NULL;
  if (bintype == 2) {
    pb = malloc(sizeof(REAL) * (np + 1));
    b->pbin = malloc(sizeof(REAL) * (np + 1));
    for (int i = 0; i <= np; i++) pb[i] = lin ? (REAL) (pmin_d + dpi_d * i) : (REAL) pedge[i];
  }
  int done = 0;
  if (bintype != 2) {           /* create_tab_sbin */
    if (lin) {
      REAL v[2] = {smin, ds};
      const REAL fac = SFX(least_fac)(v, 2);
      if (fac != 0) {
        REAL s1 = sbin[0] * fac; s1 = ORC_ROUND(s1 * s1);
        REAL s2 = sbin[ns] * fac; s2 = ORC_ROUND(s2 * s2);
        if (s1 <= ORC_MAXINT && s2 <= ORC_MAXINT && (size_t) s2 - (size_t) s1 <= 40960) {
          b->rescale = fac;
          for (int i = 0; i <= ns; i++) { sbin[i] = ORC_ROUND(sbin[i] * fac); b->s2bin[i] = sbin[i] * sbin[i]; }
          b->tabtype = 0;
          b->stab = SFX(lut_int)(b->s2bin, ns, &b->swidth, &b->nstab);
          done = 1;
        }
      }
    }
    if (!done) {
      REAL smax = sbin[ns]; smax *= smax;
      REAL fac = 32768 / smax;
      fac = ORC_POW(2, ORC_LOGB(fac));
      while (fac * smax > 8192 * 2) fac *= (REAL) 0.5;
      b->rescale = fac;
      for (int i = 0; i <= ns; i++) { sbin[i] *= fac; b->s2bin[i] = sbin[i] * sbin[i]; }
      b->tabtype = 1;
      b->stab = SFX(lut_hybrid)(b->s2bin, ns, &b->swidth, &b->nstab);
    }
    if (bintype == 1) {         /* create_tab_mu, setup_cf.c:513-531 */
      b->mutab = malloc((size_t) nmu * nmu);
      int n = 1;
      for (int i = 0; i < nmu * nmu; i++) { while (!(i < n * n)) n++; b->mutab[i] = (uint8_t) (n - 1); }
    }
  }
  else {                        /* create_tab_sp_pi */
    if (lin) {
      REAL v[4] = {smin, ds, pmin, dpi};
      const REAL fac = SFX(least_fac)(v, 4);
      if (fac != 0) {
        REAL s1 = sbin[0] * fac; s1 = ORC_ROUND(s1 * s1);
        REAL s2 = sbin[ns] * fac; s2 = ORC_ROUND(s2 * s2);
        if (s1 <= ORC_MAXINT && s2 <= ORC_MAXINT) {
          REAL p1 = pb[0] * fac, p2 = pb[np] * fac;
          if ((size_t) s2 - (size_t) s1 <= 40960 && (size_t) p2 - (size_t) p1 <= 40960) {
            b->rescale = fac;
            for (int i = 0; i <= ns; i++) { sbin[i] = ORC_ROUND(sbin[i] * fac); b->s2bin[i] = sbin[i] * sbin[i]; }
            for (int i = 0; i <= np; i++) {
              pb[i] = ORC_ROUND(pb[i] * fac);
              b->pbin[i] = periodic ? pb[i] : pb[i] * pb[i];    /* 2pt/setup_cf.c:455-458 */
            }
            b->tabtype = 0;
            b->stab = SFX(lut_int)(b->s2bin, ns, &b->swidth, &b->nstab);
            b->ptab = SFX(lut_int)(b->pbin, np, &b->pwidth, &b->nptab);
            done = 1;
          }
        }
      }
    }
    if (!done) {
      REAL smax = sbin[ns]; smax *= smax;
      REAL pmax = pb[np]; if (!periodic) pmax *= pmax;          /* 2pt/setup_cf.c:472-473 */
      REAL larger = (smax >= pmax) ? smax : pmax, smaller = (smax >= pmax) ? pmax : smax;
      REAL fac = 32768 / larger;
      fac = ORC_POW(2, ORC_LOGB(fac));
      while (fac * smaller > 8192 * 2) fac *= (REAL) 0.5;
      b->rescale = fac;
      for (int i = 0; i <= ns; i++) { sbin[i] *= fac; b->s2bin[i] = sbin[i] * sbin[i]; }
      for (int i = 0; i <= np; i++) { pb[i] *= fac; b->pbin[i] = periodic ? pb[i] : pb[i] * pb[i]; }
      b->tabtype = 1;
      b->stab = SFX(lut_hybrid)(b->s2bin, ns, &b->swidth, &b->nstab);
      b->ptab = SFX(lut_hybrid)(b->pbin, np, &b->pwidth, &b->nptab);
    }
  }
  if (periodic) for (int i = 0; i < 3; i++) b->bsize[i] = (REAL) box[i] * b->rescale;  /* setup_cf.c:679 */
  free(sbin); free(pb);
  if (!b->stab || (bintype == 2 && !b->ptab)) { free(b); return NULL; }
  return b;
}

void SFX(oracle_free)(SFX(oracle_bins) *b) {
  if (!b) return;
  free(b->s2bin); free(b->pbin); free(b->stab); free(b->ptab); free(b->mutab); free(b);
}

/* Catalogue pre-processing of tree_create: x *= rescale in `real` precision
 * (fcfc/2pt_box/build_tree.c:121-131) and, survey non-ISO only, x[3] = x^2+y^2+z^2
 * (fcfc/2pt/build_tree.c:55-60 scalar order; :75-82 FMA order for SIMD builds). */
void SFX(oracle_preprocess)(REAL *x, REAL *y, REAL *z, REAL *s, size_t n, REAL rescale, int arith) {
  for (size_t i = 0; i < n; i++) {
    if (rescale != 1) { x[i] *= rescale; y[i] *= rescale; z[i] *= rescale; }
    if (s) {
      if (arith == 0) { REAL t = x[i] * x[i] + y[i] * y[i]; s[i] = t + z[i] * z[i]; }
      else s[i] = ORC_FMA(z[i], z[i], ORC_FMA(y[i], y[i], x[i] * x[i]));
    }
  }
}

/* Division rounded toward zero for non-negative operands (what AVX-512's
 * _mm512_maskz_div_round_p[sd](.., _MM_FROUND_TO_ZERO) returns, metric_common.c:485-486):
 * the residual a - q*b of the round-to-nearest quotient is exact in an FMA. */
static inline REAL SFX(div_rz)(REAL a, REAL b) {
  REAL q = a / b;
  if (ORC_FMA(-q, b, a) < 0) q = ORC_NEXTDOWN(q);
  return q;
}

/* Table lookup incl. the hybrid walk-down: metric_common.c:56-64 (FCFC_LOOKUP_HYBRID). */
static inline int SFX(lookup)(const void *tab, int width, int tabtype, long idx, int nbin,
    REAL val, const REAL *edges) {
  int v = width ? ((const uint16_t *) tab)[idx] : ((const uint8_t *) tab)[idx];
  if (tabtype == 1 && v >= nbin) {
    v -= nbin;
    while (v != 0 && val < edges[v]) v--;
  }
  return v;
}

/* One pair.  Returns the histogram index or -1.
 * Box:    metric_common.c:140-235 (scalar), :377-534 (vector/FMA order), with the
 *         periodic image chosen as in :266-275 but applied as a +L shift of the
 *         lower point *before* the subtraction (the node-shift form of :998-1000).
 * Survey: fcfc/2pt/metric_common.c:142-259 (scalar), :283-460 (vector/FMA order). */
static inline long SFX(pair_bin)(const SFX(oracle_bins) *b,
    REAL x1, REAL y1, REAL z1, REAL s1, REAL x2, REAL y2, REAL z2, REAL s2) {
  const int fma_ = b->arith;
  const REAL s2min = b->s2bin[0], s2max = b->s2bin[b->ns];
  /* The reference instantiates variants without the lower-bound tests when the first edge
   * is zero (count_func.c:4870 `smin0`, `pmin0`): a slightly negative survey s_perp^2 is then
   * *counted* in the first bin ((int) maps (-1,0) to 0), not rejected. */
  const int smin0 = (s2min == 0);
  const int pmin0 = (b->bintype == 2) ? (b->pbin[0] == 0) : 1;
  const REAL nmu2 = (REAL) (b->nmu * b->nmu);
  REAL dist, pi = 0, mu_num = 0;
  if (b->periodic) {
    REAL a[3] = {x1, y1, z1}, c[3] = {x2, y2, z2}, d[3];
    for (int k = 0; k < 3; k++) {
      const REAL L = b->bsize[k], h = L * (REAL) 0.5;
      REAL t = a[k] - c[k];
      if (t > h) t = a[k] - (c[k] + L);
      else if (t < -h) t = (a[k] + L) - c[k];
      d[k] = t;
    }
    if (b->bintype == 2) {
      pi = ORC_ABS(d[2]);
      if (pi >= b->pbin[b->np] || (!pmin0 && pi < b->pbin[0])) return -1;
      dist = fma_ ? ORC_FMA(d[1], d[1], d[0] * d[0]) : d[0] * d[0] + d[1] * d[1];
    }
    else {
      REAL dz2 = d[2] * d[2];
      if (fma_) dist = ORC_FMA(d[1], d[1], ORC_FMA(d[0], d[0], dz2));
      else { REAL t = d[0] * d[0] + d[1] * d[1]; dist = t + dz2; }
      mu_num = dz2;
    }
    if (dist >= s2max || (!smin0 && dist < s2min)) return -1;
  }
  else {
    if (b->bintype == 0) {
      REAL dx = x1 - x2, dy = y1 - y2, dz = z1 - z2;
      if (fma_) dist = ORC_FMA(dz, dz, ORC_FMA(dy, dy, dx * dx));
      else { REAL t = dx * dx + dy * dy; dist = t + dz * dz; }
      if (dist >= s2max || (!smin0 && dist < s2min)) return -1;
    }
    else {
      REAL t;
      if (fma_) t = ORC_FMA(z1, z2, ORC_FMA(y1, y2, x1 * x2)) * 2;
      else { REAL u = x1 * x2 + y1 * y2; t = (u + z1 * z2) * 2; }
      REAL s = s1 + s2;
      REAL d = s1 - s2;
      if (b->bintype == 1) {
        dist = s - t;
        if (dist >= s2max || (!smin0 && dist < s2min)) return -1;
        pi = d * d / (s + t);
        mu_num = pi;
      }
      else {
        pi = d * d / (s + t);
        if (pi >= b->pbin[b->np] || (!pmin0 && pi < b->pbin[0])) return -1;
        dist = s - t - pi;
        if (dist >= s2max || (!smin0 && dist < s2min)) return -1;
      }
    }
  }
  long pidx = 0;
  if (b->bintype == 1) {
    int m;
    if (!fma_) m = (dist < ORC_EPS) ? 0 : (int) ((mu_num / dist) * nmu2);
    else {      /* AVX-512: (num * nmu2) / dist rounded toward zero, 0 if dist < EPS */
      REAL q = mu_num * nmu2;
      q = (dist >= ORC_EPS) ? SFX(div_rz)(q, dist) : 0;
      if (q >= nmu2) m = b->nmu * b->nmu; else m = (int) q;
    }
    if (m >= b->nmu * b->nmu) { if (b->with_mu_one) m = b->nmu * b->nmu - 1; else return -1; }
    pidx = b->mutab[m];
  }
  long sidx = (long) (int) dist - (long) (int) s2min;
  if (sidx < 0) sidx = 0;       /* survey s_perp^2 may be slightly negative: (int) maps (-1,0) to 0 */
  sidx = SFX(lookup)(b->stab, b->swidth, b->tabtype, sidx, b->ns, dist, b->s2bin);
  if (b->bintype == 2) {
    long k = (long) (int) pi - (long) (int) b->pbin[0];
    pidx = SFX(lookup)(b->ptab, b->pwidth, b->tabtype, k, b->np, pi, b->pbin);
  }
  return sidx + pidx * b->ns;
}

/* Brute-force count_pairs (fcfc/2pt_box/count_func.c:4847): auto counts visit every
 * unordered pair i<j once (metric_common.c:2017-2018), the caller doubles; weights are
 * multiplied in `real` and accumulated in double (metric_common.c:216-231). */
int SFX(oracle_count)(const SFX(oracle_bins) *b,
    const REAL *x1, const REAL *y1, const REAL *z1, const REAL *s1, const REAL *w1, size_t n1,
    const REAL *x2, const REAL *y2, const REAL *z2, const REAL *s2, const REAL *w2, size_t n2,
    int isauto, int withwt, int64_t *cnt_i, double *cnt_d) {
  const size_t ntot = (size_t) b->ns * (b->bintype == 0 ? 1 : (b->bintype == 1 ? b->nmu : b->np));
  int err = 0;
#pragma omp parallel
  {
    int64_t *ci = calloc(ntot, sizeof(int64_t));
    double *cd = calloc(ntot, sizeof(double));
    if (!ci || !cd) {
#pragma omp atomic write
      err = 1;
    }
    else {
#pragma omp for schedule(dynamic, 16)
      for (size_t i = 0; i < n1; i++) {
        const REAL si = s1 ? s1[i] : 0;
        for (size_t j = isauto ? i + 1 : 0; j < n2; j++) {
          long k = SFX(pair_bin)(b, x1[i], y1[i], z1[i], si, x2[j], y2[j], z2[j], s2 ? s2[j] : 0);
          if (k < 0) continue;
          if (withwt) { REAL w = w1[i] * w2[j]; cd[k] += w; }
          else ci[k] += 1;
        }
      }
#pragma omp critical
      for (size_t k = 0; k < ntot; k++) { if (withwt) cnt_d[k] += cd[k]; else cnt_i[k] += ci[k]; }
    }
    free(ci); free(cd);
  }
  return err;
}
