// fcfc_b200/csrc/peaks.cu -- measured FP32 issue peak: the denominator of the pair-evaluation roofline.
// MEASURED_PEAKS.json (driver-written) holds only HBM and BF16 tensor figures; the counting kernels
// are bound by CUDA-core FP32 instruction issue, so the engine measures that
// itself: independent FFMA chains, 8 per thread, all SMs, timed with CUDA events.
#include "../../include/fcfc_gpu.h"
#include <cuda_runtime.h>

namespace {
constexpr int kIters = 8192;
__global__ void __launch_bounds__(1024) ffma_kernel(float *out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(1024) dfma_kernel(double *out, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < kIters / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Shared-memory histogram increments: every accepted pair of a count costs one (red.shared.add.u32 on one of ntot 32-bit
// counters).  Pseudo-random bins of a 4800-bin histogram (the bench workload's 40 x 120), 8 independent increments per
// step, two blocks of 1024 threads per SM.
constexpr int kHistBins = 4800, kAtomIters = 2048;
__global__ void __launch_bounds__(1024) atoms_kernel(unsigned int *out) {
  __shared__ unsigned int hist[kHistBins];
  for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  unsigned int x = (blockIdx.x * 1024u + threadIdx.x) * 2654435761u + 12345u;
  const unsigned int base = (unsigned int) __cvta_generic_to_shared(hist);
  for (int it = 0; it < kAtomIters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x = x * 1664525u + 1013904223u;
      const unsigned int a = base + 4u * __umulhi(x, (unsigned int) kHistBins);
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(1u) : "memory");
    }
  }
  __syncthreads();
  unsigned int s = 0;
  for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) s += hist[i];
  if (s == 0xffffffffu) out[blockIdx.x] = s;      // (keeps the histogram alive)
}
}  // namespace

// Shared-memory atomic increments per second (lane-increments, whole device): what the histogram update alone allows.
extern "C" double fcfc_gpu_measure_smem_atomic_peak(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  const int grid = p.multiProcessorCount * 2;
  unsigned int *out = nullptr;
  if (cudaMalloc(&out, sizeof(unsigned int) * grid) != cudaSuccess) return 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    atoms_kernel<<<grid, 1024>>>(out);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = 0; break; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double) grid * 1024 * kAtomIters * 8 / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  cudaGetLastError();
  return best;
}

// FP64 issue peak (DFMA stream): the denominator for the double-precision kernels.
extern "C" double fcfc_gpu_measure_fp64_peak(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  const int grid = p.multiProcessorCount * 2;
  double *out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * grid * 1024) != cudaSuccess) return 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<grid, 1024>>>(out, 1.0001, 0.5);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = 0; break; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double) grid * 1024 * (kIters / 4) * 8 / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  return best;
}

extern "C" double fcfc_gpu_measure_fp32_peak(double *sm_clock_mhz_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  const int grid = p.multiProcessorCount * 2;
  float *out = nullptr;
  if (cudaMalloc(&out, sizeof(float) * grid * 1024) != cudaSuccess) return 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    ffma_kernel<<<grid, 1024>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = 0; break; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double) grid * 1024 * kIters * 8 / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  if (sm_clock_mhz_out) *sm_clock_mhz_out = best / (128.0 * p.multiProcessorCount) * 1e-6;  // implied clock at 128 lanes/SM
  return best;
}
