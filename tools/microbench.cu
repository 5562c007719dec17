// Micro-benchmarks that size the design of the pair-count kernels on B200 (sm_100a).
// Measures issue rates of the instruction classes the hot loop is made of:
//   FP32 FFMA/FADD/FMUL (roofline denominator R_eval), FP64, shared-memory atomics
//   (u32 / u64 / f64, random bins), random LDS table lookups, F2I, MUFU.SQRT, IEEE div.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

constexpr int ITERS = 4096;

template <int MODE> __global__ void __launch_bounds__(1024) k_fp32(float *out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) x[i] = fmaf(x[i], a, b);          // FFMA 3 distinct regs
      else if (MODE == 1) x[i] = __fadd_rn(x[i], a);   // FADD
      else if (MODE == 2) x[i] = __fmul_rn(x[i], a);   // FMUL
      else if (MODE == 3) x[i] = fmaf(x[i], x[i], b);  // FFMA (x*x+b), 2 distinct regs
      else if (MODE == 4) x[i] = fmaf(x[i], 1.0001f, b); // FFMA imm
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The pair-evaluation sequence: 3 sub, mul, 2 fma, compare, predicated count. R primary per thread.
template <int R, bool FMA> __global__ void __launch_bounds__(256) k_eval(const float4 *sec, int nsec, unsigned long long *out, float s2max) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < nsec; i += blockDim.x) sm[i] = sec[i];
  __syncthreads();
  float px[R], py[R], pz[R];
  for (int r = 0; r < R; r++) { px[r] = threadIdx.x * 0.37f + r; py[r] = threadIdx.x * 0.11f - r; pz[r] = blockIdx.x * 0.01f + r * 0.5f; }
  unsigned cnt = 0;
  for (int rep = 0; rep < 16; rep++) {
#pragma unroll 4
    for (int j = 0; j < nsec; j++) {
      float4 q = sm[j];
#pragma unroll
      for (int r = 0; r < R; r++) {
        float dx = px[r] - q.x, dy = py[r] - q.y, dz = pz[r] - q.z;
        float d2;
        if (FMA) d2 = fmaf(dy, dy, fmaf(dx, dx, dz * dz));
        else d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        cnt += (d2 < s2max) ? 1u : 0u;
      }
    }
    s2max += 1.f;
  }
  atomicAdd(out, (unsigned long long)cnt);
}

template <int MODE> __global__ void __launch_bounds__(1024) k_fp64(double *out, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) x[i] = fma(x[i], a, b);
      else if (MODE == 1) x[i] = __dadd_rn(x[i], a);
      else if (MODE == 2) x[i] = __dmul_rn(x[i], a);
    }
  }
  double s = 0; for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// Shared-memory atomics on a histogram of H bins, random bins, all lanes active.
// MODE 0: u32 add, 1: u64 add, 2: f64 add, 3: u32 add with 25% lanes active, 4: u32 inc on lane-striped copies
template <int MODE> __global__ void __launch_bounds__(512) k_atoms(int H, unsigned long long *out, int iters) {
  extern __shared__ unsigned long long smem64[];
  unsigned *h32 = (unsigned *)smem64; double *hd = (double *)smem64;
  int nwords = (MODE == 1 || MODE == 2) ? H * 2 : H;
  if (MODE == 4) nwords = H * 32;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) h32[i] = 0;
  __syncthreads();
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1;
  for (int it = 0; it < iters; it++) {
    unsigned r = lcg(s);
    unsigned bin = (unsigned)(((unsigned long long)r * (unsigned)H) >> 24);
    if (MODE == 0) atomicAdd(&h32[bin], 1u);
    else if (MODE == 1) atomicAdd(&smem64[bin], 1ull);
    else if (MODE == 2) atomicAdd(&hd[bin], 1.0);
    else if (MODE == 3) { if ((r & 3) == 0) atomicAdd(&h32[bin], 1u); }
    else if (MODE == 4) atomicAdd(&h32[bin * 32 + (threadIdx.x & 31)], 1u);
  }
  __syncthreads();
  unsigned long long t = 0;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) t += h32[i];
  atomicAdd(out, t);
}

// Random byte-table lookups from shared memory (LUT) - H entries.
__global__ void __launch_bounds__(512) k_lds(int H, unsigned long long *out, int iters) {
  extern __shared__ unsigned char tab[];
  for (int i = threadIdx.x; i < H; i += blockDim.x) tab[i] = i & 0xff;
  __syncthreads();
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1, acc = 0;
  for (int it = 0; it < iters; it++) {
    unsigned r = lcg(s);
    unsigned idx = (unsigned)(((unsigned long long)r * (unsigned)H) >> 24);
    acc += tab[idx];
  }
  atomicAdd(out, (unsigned long long)acc);
}

// misc per-accepted-pair ops. MODE 0: F2I trunc, 1: MUFU.SQRT, 2: __fdiv_rn, 3: __ddiv_rn, 4: magic-add trunc
template <int MODE> __global__ void __launch_bounds__(512) k_misc(float *out, float a) {
  float x[4]; double xd[4];
  for (int i = 0; i < 4; i++) { x[i] = threadIdx.x * 0.5f + i + 1.f; xd[i] = x[i]; }
  int acc = 0;
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (MODE == 0) { acc += (int)x[i]; x[i] += a; }
      else if (MODE == 1) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r + a; }
      else if (MODE == 2) { x[i] = __fdiv_rn(a, x[i]) + 1.5f; }
      else if (MODE == 3) { xd[i] = __ddiv_rn((double)a, xd[i]) + 1.5; }
      else if (MODE == 4) { acc += __float_as_int(__fadd_rz(x[i], 8388608.f)) & 0x7fffff; x[i] += a; }
    }
  }
  float s = acc; for (int i = 0; i < 4; i++) s += x[i] + (float)xd[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < reps; i++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, nsm, clk);
  float *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  unsigned long long *cnt; CK(cudaMalloc(&cnt, 8)); cudaMemset(cnt, 0, 8);
  int grid = nsm * 4;
  const char *names[] = {"ffma_3reg", "fadd", "fmul", "ffma_xx", "ffma_imm"};
  auto rep32 = [&](const char *n, float ms) { double ops = (double)grid * 1024 * ITERS * 8; printf("{\"test\":\"%s\",\"ms\":%.4f,\"Tinstr_lanes_per_s\":%.3f,\"lanes_per_clk_per_sm_at_max\":%.2f}\n", n, ms, ops / ms * 1e-9, ops / (ms * 1e-3) / nsm / (clk * 1e3)); };
  rep32(names[0], timeit([&] { k_fp32<0><<<grid, 1024>>>(out, 1.0001f, 0.5f); }));
  rep32(names[1], timeit([&] { k_fp32<1><<<grid, 1024>>>(out, 1.0001f, 0.5f); }));
  rep32(names[2], timeit([&] { k_fp32<2><<<grid, 1024>>>(out, 1.0001f, 0.5f); }));
  rep32(names[3], timeit([&] { k_fp32<3><<<grid, 1024>>>(out, 1.0001f, 0.5f); }));
  rep32(names[4], timeit([&] { k_fp32<4><<<grid, 1024>>>(out, 1.0001f, 0.5f); }));
  const char *n64[] = {"dfma", "dadd", "dmul"};
  auto rep64 = [&](const char *n, float ms) { double ops = (double)grid * 1024 * (ITERS / 4) * 8; printf("{\"test\":\"%s\",\"ms\":%.4f,\"Tinstr_lanes_per_s\":%.3f,\"lanes_per_clk_per_sm_at_max\":%.2f}\n", n, ms, ops / ms * 1e-9, ops / (ms * 1e-3) / nsm / (clk * 1e3)); };
  rep64(n64[0], timeit([&] { k_fp64<0><<<grid, 1024>>>((double *)out, 1.0001, 0.5); }));
  rep64(n64[1], timeit([&] { k_fp64<1><<<grid, 1024>>>((double *)out, 1.0001, 0.5); }));
  rep64(n64[2], timeit([&] { k_fp64<2><<<grid, 1024>>>((double *)out, 1.0001, 0.5); }));
  // eval sequence
  {
    int nsec = 2048; float4 *sec; CK(cudaMalloc(&sec, nsec * 16)); cudaMemset(sec, 0, nsec * 16);
    auto rep = [&](const char *n, int R, float ms, int g) { double ev = (double)g * 256 * R * nsec * 16; printf("{\"test\":\"%s\",\"ms\":%.4f,\"Tevals_per_s\":%.3f,\"evals_per_clk_per_sm_at_max\":%.2f}\n", n, ms, ev / ms * 1e-9, ev / (ms * 1e-3) / nsm / (clk * 1e3)); };
    int g = nsm * 8;
    rep("eval_fma_R1", 1, timeit([&] { k_eval<1, true><<<g, 256, nsec * 16>>>(sec, nsec, cnt, 1.f); }), g);
    rep("eval_fma_R2", 2, timeit([&] { k_eval<2, true><<<g, 256, nsec * 16>>>(sec, nsec, cnt, 1.f); }), g);
    rep("eval_fma_R4", 4, timeit([&] { k_eval<4, true><<<g, 256, nsec * 16>>>(sec, nsec, cnt, 1.f); }), g);
    rep("eval_fma_R8", 8, timeit([&] { k_eval<8, true><<<g, 256, nsec * 16>>>(sec, nsec, cnt, 1.f); }), g);
    rep("eval_unfused_R4", 4, timeit([&] { k_eval<4, false><<<g, 256, nsec * 16>>>(sec, nsec, cnt, 1.f); }), g);
    rep("eval_unfused_R8", 8, timeit([&] { k_eval<8, false><<<g, 256, nsec * 16>>>(sec, nsec, cnt, 1.f); }), g);
  }
  // atomics
  {
    int iters = 4096; int g = nsm * 4;
    auto rep = [&](const char *n, int H, float ms, double frac) { double ops = (double)g * 512 * iters * frac; printf("{\"test\":\"%s\",\"H\":%d,\"ms\":%.4f,\"Gatomics_per_s\":%.2f,\"atomic_lanes_per_clk_per_sm_at_max\":%.3f}\n", n, H, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / nsm / (clk * 1e3)); };
    int Hs[] = {40, 1600, 4800};
    for (int H : Hs) {
      rep("atoms_u32", H, timeit([&] { k_atoms<0><<<g, 512, H * 8>>>(H, cnt, iters); }), 1);
      rep("atoms_u64", H, timeit([&] { k_atoms<1><<<g, 512, H * 8>>>(H, cnt, iters); }), 1);
      rep("atoms_f64", H, timeit([&] { k_atoms<2><<<g, 512, H * 8>>>(H, cnt, iters); }), 1);
      rep("atoms_u32_quarter_lanes", H, timeit([&] { k_atoms<3><<<g, 512, H * 8>>>(H, cnt, iters); }), 0.25);
    }
    rep("atoms_u32_lane_striped", 40, timeit([&] { k_atoms<4><<<g, 512, 40 * 32 * 4>>>(40, cnt, iters); }), 1);
    auto repl = [&](const char *n, int H, float ms) { double ops = (double)g * 512 * iters; printf("{\"test\":\"%s\",\"H\":%d,\"ms\":%.4f,\"Glookups_per_s\":%.2f,\"lanes_per_clk_per_sm_at_max\":%.3f}\n", n, H, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / nsm / (clk * 1e3)); };
    repl("lds_u8_random", 1600, timeit([&] { k_lds<<<g, 512, 1600>>>(1600, cnt, iters); }));
    repl("lds_u8_random", 14400, timeit([&] { k_lds<<<g, 512, 14400>>>(14400, cnt, iters); }));
  }
  {
    const char *nm[] = {"f2i_trunc", "mufu_sqrt", "fdiv_rn", "ddiv_rn", "magic_trunc"};
    int g = nsm * 4;
    auto rep = [&](const char *n, float ms) { double ops = (double)g * 512 * (ITERS / 4) * 4; printf("{\"test\":\"%s\",\"ms\":%.4f,\"Gops_per_s\":%.2f,\"lanes_per_clk_per_sm_at_max\":%.3f}\n", n, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / nsm / (clk * 1e3)); };
    rep(nm[0], timeit([&] { k_misc<0><<<g, 512>>>(out, 1.25f); }));
    rep(nm[1], timeit([&] { k_misc<1><<<g, 512>>>(out, 1.25f); }));
    rep(nm[2], timeit([&] { k_misc<2><<<g, 512>>>(out, 1.25f); }));
    rep(nm[3], timeit([&] { k_misc<3><<<g, 512>>>(out, 1.25f); }));
    rep(nm[4], timeit([&] { k_misc<4><<<g, 512>>>(out, 1.25f); }));
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
