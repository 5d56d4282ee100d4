// Micro-benchmark: cycles per tcgen05.mma (kind::f16, fp32 accumulate) on sm_100a as a function of
//   operand form (SS: A and B from shared memory; TS: A from tensor memory), N, accumulator reuse,
//   concurrent shared-memory traffic from 16 other warps, and cta_group::1 (M=128) vs ::2 (M=256).
// Every CTA of a 148-CTA grid runs the same loop; CTA 0 reports clock64 deltas.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../baybe_b200/csrc/common.cuh"

namespace bb {
void set_error(const char*, ...) {}
}
using namespace bb;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

struct Cfg {
  int form;      // 0 SS, 1 TS
  int n;         // UMMA N
  int nacc;      // accumulators cycled through (1: all into one)
  int bg;        // background shared-memory traffic from the 16 other warps (0 none, 1 LDS+STS 128-bit)
  int reps;      // MMAs per timed batch
  int batches;   // timed batches (commit + wait between)
  int elect;     // 0: the issuing code runs under `if (lane == 0)`; 1: whole warp converged, issue under elect.sync
};

// smem: [A tile 128 x 64 fp16 SW128 = 16 KB] [B tile 256 x 64 fp16 SW128 = 32 KB] [bg scratch 64 KB]
constexpr int kThreads = 18 * 32;
constexpr size_t kSmem = 16384 + 32768 + 65536 + 1024;

template <int CG, int ELECT>
__global__ void __launch_bounds__(kThreads, 1) k_mma(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  uint8_t* sm = smem_raw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < (int)((16384 + 32768 + 65536) / 16); e += kThreads)
    reinterpret_cast<uint4*>(sm)[e] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);  // fp16 1.0
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  if (warp == 16) {
    if (CG == 1) {
      tmem_alloc(&tmem_ptr, 512);
      tmem_relinquish();
    } else {
      tmem_alloc2(&tmem_ptr, 512);
      tmem_relinquish2();
    }
  }
  fence_proxy_async();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  const bool leader = (CG == 1) || cluster_ctarank() == 0;

  if (warp == 17) {
    if (leader && (ELECT || lane == 0)) {
      const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm + 16384);
      const uint64_t a_desc = make_sw128_desc(a_addr), b_desc = make_sw128_desc(b_addr);
      const uint32_t idesc = make_idesc_f16(CG == 1 ? 128 : 256, c.n);
      uint32_t ph = 0;
      long long tot = 0, tmin = 1ll << 60;
      for (int bt = 0; bt < c.batches; ++bt) {
        const long long t0 = clock64();
        const uint32_t dmask = (uint32_t)c.nacc - 1u;  // nacc is a power of two
#pragma unroll 1
        for (int r = 0; r < c.reps; r += 4) {
          const uint32_t d = tmem_base + (((uint32_t)(r >> 2) & dmask) * (uint32_t)c.n);  // accumulators in columns 0..255
          const bool issue = ELECT ? elect_one() : true;
          if (issue) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint64_t ko = (uint64_t)(u * 2);
              const uint32_t a_t = tmem_base + 256u + (uint32_t)(u * 8);  // TS: A in columns 256..
              if (CG == 1) {
                if (c.form == 0) umma_f16(d, a_desc + ko, b_desc + ko, idesc, 1u);
                else umma_f16_ts(d, a_t, b_desc + ko, idesc, 1u);
              } else {
                if (c.form == 0) umma2_f16_ss(d, a_desc + ko, b_desc + ko, idesc, 1u);
                else umma2_f16_ts(d, a_t, b_desc + ko, idesc, 1u);
              }
            }
          }
          if (ELECT) __syncwarp();
        }
        if (ELECT) {
          if (elect_one()) { if (CG == 1) umma_commit(&bar); else umma2_commit(&bar); }
          __syncwarp();
        } else {
          if (CG == 1) umma_commit(&bar); else umma2_commit(&bar);
        }
        mbar_wait(&bar, ph);
        ph ^= 1u;
        const long long dt = clock64() - t0;
        if (bt > 0) {
          tot += dt;
          tmin = dt < tmin ? dt : tmin;
        }
      }
      if (blockIdx.x == 0 && lane == 0) {
        out[0] = tot / (c.batches - 1);
        out[1] = tmin;
      }
      if (lane == 0) stop = 1;
    } else if (lane == 0) {
      stop = 1;  // non-leader CTA of a pair: nothing to issue
    }
  } else if (warp < 16 && c.bg) {
    // 128-bit loads and stores on a private 64 KB region until the issuer is done
    uint4* scratch = reinterpret_cast<uint4*>(sm + 16384 + 32768);
    uint4 acc = make_uint4(0, 0, 0, 0);
    int guard = 0;
    while (!stop && guard < (1 << 22)) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 v = scratch[(tid + u * 512) & 4095];
        acc.x ^= v.x;
        acc.y += v.y;
      }
      scratch[(tid * 7 + guard) & 4095] = acc;
      ++guard;
    }
    if (acc.x == 0x12345u) out[7] = acc.y;
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 16) {
    if (CG == 1) tmem_dealloc(tmem_base, 512); else tmem_dealloc2(tmem_base, 512);
  }
}

static void run(int cg, Cfg c, long long* d_out) {
  cudaMemset(d_out, 0, 64);
  cudaError_t e;
  void (*kern)(Cfg, long long*) = cg == 1 ? (c.elect ? k_mma<1, 1> : k_mma<1, 0>) : (c.elect ? k_mma<2, 1> : k_mma<2, 0>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cg;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, c, d_out);
  e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  const int m = cg == 1 ? 128 : 256;
  const double cyc = (double)h[0] / c.reps, cmin = (double)h[1] / c.reps;
  const double floor_c = (double)m * c.n / (256.0 * cg) ;  // guide: max(M,128)*N/(256*cta_group) per K=16 atom
  printf("cg=%d M=%d %s N=%3d nacc=%d bg=%d reps=%3d elect=%d : %7.1f cyc/MMA (min %7.1f)  floor %5.1f  eff %.2f  %s\n", cg, m,
         c.form ? "TS" : "SS", c.n, c.nacc, c.bg, c.reps, c.elect, cyc, cmin, floor_c, floor_c / cmin,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int cg = argc > 1 ? atoi(argv[1]) : 1;
  long long* d_out;
  cudaMalloc(&d_out, 64);
  for (int elect = 0; elect < 2; ++elect)
    for (int bg = 0; bg < 2; ++bg)
      for (int form = 0; form < 2; ++form)
        for (int n : {64, 128, 256}) {
          run(cg, Cfg{form, n, 1, bg, 64, 5, elect}, d_out);
          if (n == 64) run(cg, Cfg{form, n, 4, bg, 64, 5, elect}, d_out);
        }
  return 0;
}
