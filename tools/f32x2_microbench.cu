#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void k(float *out, float s, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  if (MODE == 0) {
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) { a0 = fmaf(a0, s, s); a1 = fmaf(a1, s, s); a2 = fmaf(a2, s, s); a3 = fmaf(a3, s, s); a4 = fmaf(a4, s, s); a5 = fmaf(a5, s, s); a6 = fmaf(a6, s, s); a7 = fmaf(a7, s, s); }
    }
  } else {
    unsigned long long p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7), ss = pk(s, s);
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (MODE == 1) { p0 = fma2(p0, ss, ss); p1 = fma2(p1, ss, ss); p2 = fma2(p2, ss, ss); p3 = fma2(p3, ss, ss); }
        if (MODE == 2) { p0 = add2(p0, ss); p1 = add2(p1, ss); p2 = add2(p2, ss); p3 = add2(p3, ss); }
        if (MODE == 3) { p0 = mul2(p0, ss); p1 = mul2(p1, ss); p2 = mul2(p2, ss); p3 = mul2(p3, ss); }
      }
    }
    upk(p0, a0, a1); upk(p1, a2, a3); upk(p2, a4, a5); upk(p3, a6, a7);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <int MODE> void run(const char *name, float *d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 4096, blocks = 148 * 4, thr = 512;
  k<MODE><<<blocks, thr>>>(d, 0.999f, 16);
  cudaEventRecord(e0); k<MODE><<<blocks, thr>>>(d, 0.999f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops_lane_ops = (double) blocks * thr * iters * 64.0;   // scalar-equivalent FP32 ops (fma counted once)
  printf("%s: %.3f ms, %.2f T scalar-equivalent lane-ops/s\n", name, ms, flops_lane_ops / ms * 1e-9);
}
int main() { float *d; cudaMalloc(&d, 148 * 4 * 512 * 4); run<0>("ffma scalar", d); run<1>("fma.f32x2", d); run<2>("add.f32x2", d); run<3>("mul.f32x2", d); return 0; }
