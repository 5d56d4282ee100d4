// Micro-benchmark: FFMA vs FFMA2 (packed fp32x2) issue/throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16]; unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2*i]) << 32) | __float_as_uint(a[2*i+1]);
  unsigned long long sp = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], sp, sp);
    }
  }
  float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int iters = 20000;
  for (int warps_per_sm : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 2; ++mode) {
      int threads = 128, blocks = 148 * (warps_per_sm / 4);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<blocks, threads>>>(out, iters, 0.999f); else k<1><<<blocks, threads>>>(out, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)blocks * threads * iters * 16;  // scalar FMAs (FFMA2 = 2 each)
      printf("warps/SM=%2d %s: %.3f ms  %.1f GFMA/s  = %.1f FMA/clk/SM @1.9GHz\n", warps_per_sm, mode ? "FFMA2" : "FFMA ", ms,
             fma / ms / 1e6, fma / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
