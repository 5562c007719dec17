// fcfc_b200/csrc/cnvt.cu -- (ra, dec, z) -> comoving Cartesian coordinates on the device.
//
// Replaces cnvt_coord_integr of the survey host (src/fcfc/2pt/cnvt_coord.c:321-337) together with the quadrature it
// calls (cnvt_legauss :227-249, cnvt_integrand :201-211): one thread per object, Legendre-Gauss integration of
// c / (100 E(z)) from 0 to z with the order the host selected (its convergence test stays on the host: 129 sample
// redshifts), then d cos(dec) cos(ra), d cos(dec) sin(ra), d sin(dec).  Every operation of the distance is the
// reference's, in its order, unfused (the host is compiled as ISO C: no contraction), with IEEE sqrt and division: the
// comoving DISTANCE is bit-identical to the host's.  sin / cos come from the CUDA math library (<= 2 ulp) instead of the
// host's libm (<= 1 ulp), so a coordinate (two such factors) can differ from the host's by a few ulp -- 5 at most and 15 % of the coordinates in the GPU test: the conversion is therefore
// opt-in (FCFC_GPU_CNVT=1 in the shim), and the host's own conversion remains the default and the parity path.
// Only w = -1 dark energy (no pow()) is supported; the shim falls back to the host otherwise.
#include "../../include/fcfc_gpu.h"
#include <cuda_runtime.h>
#include <cstdio>

namespace {
constexpr int kMaxNodes = 17;           // orders up to 32 (+ the centre weight of odd orders)
struct Quad { double x[kMaxNodes], w[kMaxNodes]; int npair, odd; };

__device__ __forceinline__ double integrand(double om, double ol, double ok, double z) {        // cnvt_coord.c:201-211
  const double z1 = __dadd_rn(z, 1.0);
  const double z2 = __dmul_rn(z1, z1);
  double d = __dmul_rn(__dmul_rn(om, z2), z1);
  if (ok != 0.0) d = __dadd_rn(d, __dmul_rn(ok, z2));
  d = __dadd_rn(d, ol);
  return __ddiv_rn(299792.458 * 0.01, __dsqrt_rn(d));
}

template <class T>
__global__ void cnvt_kernel(T *x, T *y, T *z, size_t n, double om, double ol, double ok, const __grid_constant__ Quad q) {
  const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ra = __dmul_rn((double) x[i], 0x1.1df46a2529d39p-6);       // DEGREE_2_RAD, define_comm.h:137
  const double dec = __dmul_rn((double) y[i], 0x1.1df46a2529d39p-6);
  const double zz = (double) z[i];
  const double zp = __dmul_rn(zz, 0.5);                                    // cnvt_coord.c:227-249
  double sum = 0;
  for (int k = 0; k < q.npair; k++) {
    const double za = __dmul_rn(zp, __dadd_rn(1.0, q.x[k])), zb = __dmul_rn(zp, __dsub_rn(1.0, q.x[k]));
    sum = __dadd_rn(sum, __dmul_rn(q.w[k], __dadd_rn(integrand(om, ol, ok, za), integrand(om, ol, ok, zb))));
  }
  if (q.odd) sum = __dadd_rn(sum, __dmul_rn(q.w[q.npair], integrand(om, ol, ok, zp)));
  const double dist = __dmul_rn(sum, zp);
  const double cd = cos(dec), dc = __dmul_rn(dist, cd);
  x[i] = (T) __dmul_rn(dc, cos(ra));                                       // cnvt_coord.c:333-335
  y[i] = (T) __dmul_rn(dc, sin(ra));
  z[i] = (T) __dmul_rn(dist, sin(dec));
}

template <class T>
int run(void *x, void *y, void *z, size_t n, double om, double ol, double ok, const Quad &q) {
  if (!n) return 0;
  T *d[3] = {nullptr, nullptr, nullptr};
  void *h[3] = {x, y, z};
  int rc = 0;
  for (int k = 0; k < 3 && !rc; k++) {
    if (cudaMalloc(&d[k], n * sizeof(T)) != cudaSuccess) rc = FCFC_GPU_ERR_MEMORY;
    else if (cudaMemcpy(d[k], h[k], n * sizeof(T), cudaMemcpyDefault) != cudaSuccess) rc = FCFC_GPU_ERR_CUDA;
  }
  if (!rc) {
    cnvt_kernel<T><<<(unsigned) ((n + 255) / 256), 256>>>(d[0], d[1], d[2], n, om, ol, ok, q);
    for (int k = 0; k < 3 && !rc; k++)
      if (cudaMemcpy(h[k], d[k], n * sizeof(T), cudaMemcpyDefault) != cudaSuccess) rc = FCFC_GPU_ERR_CUDA;
  }
  for (int k = 0; k < 3; k++) cudaFree(d[k]);
  if (rc) cudaGetLastError();
  return rc;
}
}  // namespace

extern "C" int fcfc_gpu_cnvt_coord(void *x, void *y, void *z, size_t n, int is_float, double omega_m, double omega_l,
                                   double omega_k, int order, const double *gl_x, const double *gl_w) {
  if (order < 1 || order > 32 || !gl_x || !gl_w || (n && (!x || !y || !z))) return FCFC_GPU_ERR_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return FCFC_GPU_ERR_CUDA; }     // no CPU fallback
  Quad q;
  q.npair = order >> 1; q.odd = order & 1;
  for (int k = 0; k < q.npair + q.odd; k++) { q.x[k] = gl_x[k]; q.w[k] = gl_w[k]; }
  return is_float ? run<float>(x, y, z, n, omega_m, omega_l, omega_k, q) : run<double>(x, y, z, n, omega_m, omega_l, omega_k, q);
}
