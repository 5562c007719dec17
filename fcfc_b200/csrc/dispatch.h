// fcfc_b200/csrc/dispatch.h -- maps the runtime description of a count onto one of the compiled
// kernel variants.  The reference selects one of 384 macro-generated functions at this point
// (src/fcfc/2pt_box/count_func.c:4871-7141); here the lookup-table width, table type and the
// non-zero lower bounds are handled at run time inside a "generic" variant, which leaves
// precision x binning x metric x weights x arithmetic order x {fast, generic, global-histogram}.
#pragma once
#include "count_kernel.cuh"
#include "count_kernel_pf.cuh"
#include "count_kernel_df.cuh"
#include "count_kernel_cl.cuh"

namespace fcfc {

struct Variant {
  bool is_float, box, wt, generic, smem_hist;
  int bintype, arith;
};

constexpr int kR = 4;   // primary points per lane (a tile holds up to 32 kR points of a cell; count_kernel_cl: kClR)

// Defined (explicitly instantiated) in the generated inst_*.cu files.
template <class T, int BIN, bool BOX, bool WT, int ARITH, bool GENERIC, bool SMEMHIST>
cudaError_t launch_variant(const CountParams<T> &P, int nblocks, int smem_bytes);

template <class T, int BIN, bool BOX, bool WT, int ARITH>
static cudaError_t launch_l3(const Variant &v, const CountParams<T> &P, int nb, int sm) {
  if (!v.smem_hist) return launch_variant<T, BIN, BOX, WT, ARITH, true, false>(P, nb, sm);
  if (v.generic) return launch_variant<T, BIN, BOX, WT, ARITH, true, true>(P, nb, sm);
  return launch_variant<T, BIN, BOX, WT, ARITH, false, true>(P, nb, sm);
}
template <class T, int BIN, bool BOX>
static cudaError_t launch_l2(const Variant &v, const CountParams<T> &P, int nb, int sm) {
  if (v.wt) return v.arith ? launch_l3<T, BIN, BOX, true, 1>(v, P, nb, sm) : launch_l3<T, BIN, BOX, true, 0>(v, P, nb, sm);
  return v.arith ? launch_l3<T, BIN, BOX, false, 1>(v, P, nb, sm) : launch_l3<T, BIN, BOX, false, 0>(v, P, nb, sm);
}
template <class T>
static cudaError_t launch_count(const Variant &v, const CountParams<T> &P, int nb, int sm) {
  switch (v.bintype) {
    case BIN_ISO: return v.box ? launch_l2<T, BIN_ISO, true>(v, P, nb, sm) : launch_l2<T, BIN_ISO, false>(v, P, nb, sm);
    case BIN_SMU: return v.box ? launch_l2<T, BIN_SMU, true>(v, P, nb, sm) : launch_l2<T, BIN_SMU, false>(v, P, nb, sm);
    default: return v.box ? launch_l2<T, BIN_SPI, true>(v, P, nb, sm) : launch_l2<T, BIN_SPI, false>(v, P, nb, sm);
  }
}

// Double precision through the float pre-filter (count_kernel_pf.cuh): shared-memory histogram variants only.
template <int BIN, bool BOX, bool WT, int ARITH, bool GENERIC>
cudaError_t launch_variant_pf(const CountParams<double> &P, int nblocks, int smem_bytes);

template <int BIN, bool BOX>
static cudaError_t launch_pf_l2(const Variant &v, const CountParams<double> &P, int nb, int sm) {
#define FCFC_PF_PICK(WT, AR) (v.generic ? launch_variant_pf<BIN, BOX, WT, AR, true>(P, nb, sm) : launch_variant_pf<BIN, BOX, WT, AR, false>(P, nb, sm))
  if (v.wt) return v.arith ? FCFC_PF_PICK(true, 1) : FCFC_PF_PICK(true, 0);
  return v.arith ? FCFC_PF_PICK(false, 1) : FCFC_PF_PICK(false, 0);
#undef FCFC_PF_PICK
}
static inline cudaError_t launch_count_pf(const Variant &v, const CountParams<double> &P, int nb, int sm) {
  switch (v.bintype) {
    case BIN_ISO: return v.box ? launch_pf_l2<BIN_ISO, true>(v, P, nb, sm) : launch_pf_l2<BIN_ISO, false>(v, P, nb, sm);
    case BIN_SMU: return v.box ? launch_pf_l2<BIN_SMU, true>(v, P, nb, sm) : launch_pf_l2<BIN_SMU, false>(v, P, nb, sm);
    default: return v.box ? launch_pf_l2<BIN_SPI, true>(v, P, nb, sm) : launch_pf_l2<BIN_SPI, false>(v, P, nb, sm);
  }
}

#define FCFC_DEFINE_VARIANT_PF(BIN, BOX, WT, ARITH, GENERIC)                                             \
  template <> cudaError_t launch_variant_pf<BIN, BOX, WT, ARITH, GENERIC>(                               \
      const CountParams<double> &P, int nblocks, int smem_bytes) {                                       \
    auto kern = count_kernel_pf<BIN, BOX, WT, ARITH, GENERIC, true, kR>;                                 \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
    if (e != cudaSuccess) return e;                                                                      \
    kern<<<nblocks, pf_warps(BIN, BOX) * 32, smem_bytes>>>(P);                                                        \
    return cudaGetLastError();                                                                           \
  }

// Double precision at float speed (count_kernel_df.cuh): box (s,mu) / isotropic and survey isotropic, computed bins.
template <int BIN, bool BOX, bool WT, int ARITH>
cudaError_t launch_variant_df(const CountParams<double> &P, int nblocks, int smem_bytes);

template <int LAZY = 0>      // (a template, so that the specialisations of the generated files precede its instantiation)
static cudaError_t launch_count_df(const Variant &v, const CountParams<double> &P, int nb, int sm) {
#define FCFC_DF_PICK(BIN, BOX) (v.wt ? (v.arith ? launch_variant_df<BIN, BOX, true, 1>(P, nb, sm) : launch_variant_df<BIN, BOX, true, 0>(P, nb, sm)) \
                                     : (v.arith ? launch_variant_df<BIN, BOX, false, 1>(P, nb, sm) : launch_variant_df<BIN, BOX, false, 0>(P, nb, sm)))
  if (v.bintype == BIN_SMU) return FCFC_DF_PICK(BIN_SMU, true);
  return v.box ? FCFC_DF_PICK(BIN_ISO, true) : FCFC_DF_PICK(BIN_ISO, false);
#undef FCFC_DF_PICK
}

#define FCFC_DEFINE_VARIANT_DF(BIN, BOX, WT, ARITH)                                                      \
  template <> cudaError_t launch_variant_df<BIN, BOX, WT, ARITH>(                                        \
      const CountParams<double> &P, int nblocks, int smem_bytes) {                                       \
    auto kern = count_kernel_df<BIN, BOX, WT, ARITH, kR>;                                                \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
    if (e != cudaSuccess) return e;                                                                      \
    kern<<<nblocks, kDfThreads, smem_bytes>>>(P);                                                        \
    return cudaGetLastError();                                                                           \
  }

// Single precision with classified staging (count_kernel_cl.cuh): box (s,mu) / isotropic and survey isotropic, computed bins.
template <int BIN, bool BOX, int ARITH>
cudaError_t launch_variant_cl(const CountParams<float> &P, int nblocks, int smem_bytes);

template <int LAZY = 0>
static cudaError_t launch_count_cl(const Variant &v, const CountParams<float> &P, int nb, int sm) {
#define FCFC_CL_PICK(BIN, BOX) (v.arith ? launch_variant_cl<BIN, BOX, 1>(P, nb, sm) : launch_variant_cl<BIN, BOX, 0>(P, nb, sm))
  if (v.bintype == BIN_SMU) return FCFC_CL_PICK(BIN_SMU, true);
  return v.box ? FCFC_CL_PICK(BIN_ISO, true) : FCFC_CL_PICK(BIN_ISO, false);
#undef FCFC_CL_PICK
}

#define FCFC_DEFINE_VARIANT_CL(BIN, BOX, ARITH)                                                          \
  template <> cudaError_t launch_variant_cl<BIN, BOX, ARITH>(                                            \
      const CountParams<float> &P, int nblocks, int smem_bytes) {                                        \
    auto kern = count_kernel_cl<BIN, BOX, ARITH, kClR>;                                                    \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
    if (e != cudaSuccess) return e;                                                                      \
    kern<<<nblocks, kClThreads, smem_bytes>>>(P);                                                        \
    return cudaGetLastError();                                                                           \
  }

// Body used by the generated instantiation files.
#define FCFC_DEFINE_VARIANT(T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST)                                   \
  template <> cudaError_t launch_variant<T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST>(                     \
      const CountParams<T> &P, int nblocks, int smem_bytes) {                                            \
    auto kern = count_kernel<T, BIN, BOX, WT, ARITH, GENERIC, SMEMHIST, kR>;                             \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
    if (e != cudaSuccess) return e;                                                                      \
    kern<<<nblocks, BlockShape<T>::kThreads, smem_bytes>>>(P);                                                          \
    return cudaGetLastError();                                                                           \
  }

}  // namespace fcfc
