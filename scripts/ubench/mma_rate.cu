// mma_rate.cu -- tcgen05.mma issue-rate micro-benchmark (sm_100a), cta_group::1.
//
// One CTA, one issuing warp.  A burst of `reps` kind::f16 MMAs (M = 128, K = 16, N = 64 / 128 / 256; A from shared
// memory "SS" or from tensor memory "TS"; `nacc` accumulators used round-robin) is issued either under
// `if (lane == 0)` (elect = 0) or by a converged warp under elect.sync (elect = 1), committed to an mbarrier and
// waited for; cycles per MMA = (clock after the wait - clock before the first issue) / reps.  `bg` = 1 adds seven
// warps of dependent FMA work on the same SM.  Operand contents are irrelevant (zero-filled shared memory).
// This is the tool behind the "elect.sync issue" and "TS form" decisions of fused_ts.cu; the committed output
// profiles/r02_ubench_mma_issue.txt was produced by the first version of this file (which also covered
// cta_group::2 and was lost with a replaced container before it was committed; this is a re-creation of the
// cta_group::1 part).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../baybe_b200/csrc -I../../include \
//        mma_rate.cu -o mma_rate        run: ./mma_rate
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace bb;

__device__ __forceinline__ bool elect_one_ub() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

template <int N, bool TS, bool ELECT>
__global__ void __launch_bounds__(256, 1) k_rate(int reps, int nacc, int bg, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < (16384 + 32768) / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 0) {
    const uint64_t a_desc = make_sw128_desc(smem_u32(smem));            // A: 128 rows x 64 k (only 16 k used per MMA)
    const uint64_t b_desc = make_sw128_desc(smem_u32(smem + 16384));    // B: up to 256 rows x 64 k
    const uint32_t idesc = make_idesc_f16(128, N);
    const int cols_per_acc = N;                                         // accumulators side by side, A operand at column 448
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {  // the last repetition is reported
      __syncwarp();
      t0 = clock64();
      if (ELECT) {
        if (elect_one_ub()) {
#pragma unroll 8
          for (int i = 0; i < reps; ++i) {
            const uint32_t d = tm + (uint32_t)((i % nacc) * cols_per_acc) % 448u;
            if (TS) umma_ts(d, tm + 448u, b_desc + (uint64_t)((i & 3) * 2), idesc, 1u);
            else umma_f16(d, a_desc + (uint64_t)((i & 3) * 2), b_desc + (uint64_t)((i & 3) * 2), idesc, 1u);
          }
          umma_commit(&bar);
        }
        __syncwarp();
      } else {
        if (lane == 0) {
#pragma unroll 8
          for (int i = 0; i < reps; ++i) {
            const uint32_t d = tm + (uint32_t)((i % nacc) * cols_per_acc) % 448u;
            if (TS) umma_ts(d, tm + 448u, b_desc + (uint64_t)((i & 3) * 2), idesc, 1u);
            else umma_f16(d, a_desc + (uint64_t)((i & 3) * 2), b_desc + (uint64_t)((i & 3) * 2), idesc, 1u);
          }
          umma_commit(&bar);
        }
        __syncwarp();
      }
      mbar_wait(&bar, (uint32_t)(rep & 1));
      tc_fence_after();
      t1 = clock64();
    }
    if (lane == 0) out[0] = t1 - t0;
  } else if (bg) {
    float x = (float)threadIdx.x;
    for (int i = 0; i < 200000; ++i) x = fmaf(x, 1.0000001f, 0.5f);
    if (x == 123.0f) sink[0] = x;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, bool TS, bool ELECT>
static void run(int reps, int nacc, int bg, long long* d_out, float* d_sink) {
  auto kern = k_rate<N, TS, ELECT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  double best = 1e30, sum = 0;
  for (int it = 0; it < 5; ++it) {
    kern<<<1, 256, 16384 + 32768 + 1024, 0>>>(reps, nacc, bg, d_out, d_sink);
    long long c = 0;
    cudaMemcpy(&c, d_out, sizeof(c), cudaMemcpyDeviceToHost);
    const double per = (double)c / reps;
    best = per < best ? per : best;
    sum += per;
  }
  const double floor_c = N / 2.0;  // 128 x N x 16 MMA at 8192 flop/clk/SM
  printf("cg=1 M=128 %s N=%3d nacc=%d bg=%d reps=%3d elect=%d : %7.1f cyc/MMA (min %7.1f)  floor %5.1f  eff %.2f\n",
         TS ? "TS" : "SS", N, nacc, bg, reps, ELECT ? 1 : 0, sum / 5, best, floor_c, floor_c / best);
}

int main() {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 64);
  cudaMalloc(&d_sink, 64);
  for (int bg = 0; bg < 2; ++bg) {
    run<64, false, false>(64, 1, bg, d_out, d_sink);  run<128, false, false>(64, 1, bg, d_out, d_sink);
    run<256, false, false>(64, 1, bg, d_out, d_sink); run<64, true, false>(64, 1, bg, d_out, d_sink);
    run<128, true, false>(64, 1, bg, d_out, d_sink);  run<256, true, false>(64, 1, bg, d_out, d_sink);
  }
  for (int bg = 0; bg < 2; ++bg) {
    run<64, false, true>(64, 1, bg, d_out, d_sink);  run<64, false, true>(64, 4, bg, d_out, d_sink);
    run<128, false, true>(64, 1, bg, d_out, d_sink); run<256, false, true>(64, 1, bg, d_out, d_sink);
    run<64, true, true>(64, 1, bg, d_out, d_sink);   run<64, true, true>(64, 4, bg, d_out, d_sink);
    run<128, true, true>(64, 1, bg, d_out, d_sink);  run<256, true, true>(64, 1, bg, d_out, d_sink);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", e == cudaSuccess ? "MMA_RATE_DONE" : cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}
