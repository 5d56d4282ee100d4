// fused_common.cuh -- parameters and helpers shared by the two fused scoring kernels
//   fused.cu     k_fused     distances on CUDA cores (FFMA2), any n_pad <= 512, any d
//   fused_tc.cu  k_fused_tc  distances on tcgen05 as well (fp16 hi/mid/lo split GEMM), n_pad <= 256, d_pad <= 64
#pragma once

#include "acq_math.cuh"
#include "common.cuh"

namespace bb {

constexpr int kComputeWarps = 16;
constexpr int kComputeThreads = kComputeWarps * 32;  // 512
constexpr int kFusedThreads = kComputeThreads + 64;  // + producer warp + MMA warp
constexpr int kWarpProducer = 16;
constexpr int kWarpMma = 17;
constexpr int kMaxSlotsA = 4;
constexpr int kMaxStagesB = 8;
constexpr uint32_t kSlotABytes = 32768;  // [hi 16 KB | lo 16 KB], each 128 rows x 64 fp16, SW128
constexpr uint32_t kStageBBytes = 16384; // [hi 8 KB | lo 8 KB],  each  64 rows x 64 fp16, SW128
constexpr int kMaxTasks = 16;
constexpr int kMaxSamples = 1024;

struct FusedParams {
  // candidates
  const void* x;
  int layout;
  int64_t N, ldx;
  int num_tiles;
  // model
  const float *cand_scale, *cand_shift, *train_m2, *train_sq, *alpha, *task_covar, *mean_const;
  const int32_t* train_task;
  const uint8_t* rimg;
  const uint8_t* bimg;           // distance-GEMM B operand (fused_tc only)
  float dist_scale_a, inv_dist_scale;  // a is scaled by dist_scale_a; D2 * inv_dist_scale = -2 a.b
  int family;
  int n_pad, d, d_pad, n_chunks, task_col, n_tasks;
  float y_mean, y_std, prior_scale, inv_r_scale2;
  int scaled;  // task kernel or output scale present
  // ring sizes
  int slots_a, stages_b;
  uint32_t tmem_cols;
  // acquisition (has_acq == 0: posterior only)
  int has_acq;
  bb_acq_spec acq;
  const float* z;
  int S;
  // outputs (nullable)
  float *mu, *var, *score;
  const uint8_t* keep;
  long long* best_key;
  int64_t index_offset;
};

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }


// Spin with back-off: used by the two single-lane helper warps so that their polling does not
// eat issue slots of the compute warps sharing their scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}

int launch_fused_tc(FusedParams& p, int grid, cudaStream_t stream);
bool fused_tc_supported(FusedParams& p, int max_smem);

}  // namespace bb
