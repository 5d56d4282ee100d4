// fused_common.cuh -- parameters and helpers shared by the scoring kernels
//   fused_tc.cu  k_fused_tc  distances AND posterior contraction on tcgen05 (n_pad <= 256, d_pad <= 64): headline
//   fused.cu     k_fused     distances on CUDA cores (FFMA2), n_pad <= 512, training rows resident in shared
//                            memory; its PRE variant reads a K* block instead (wide-feature path)
//   wide.cu      k_kmat_tc   K-looped tcgen05 distance GEMM for wide / bit-packed feature spaces and n_pad > 512
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
  const uint8_t* rimg2;          // pair-grouped L^-1 image (fused_tc only)
  float dist_scale_a, inv_dist_scale;  // a is scaled by dist_scale_a; D2 * inv_dist_scale = -2 a.b
  int family;
  int dist_k;                    // 32 or 64: K extent of the distance-GEMM tiles (0: none)
  int n_pad, d, d_pad, n_chunks, task_col, n_tasks;
  float y_mean, y_std, prior_scale, inv_r_scale2;
  int scaled;  // task kernel or output scale present
  // ring sizes
  int slots_a, stages_b;
  uint32_t stage_b_bytes;  // bytes of one L^-1 ring stage: GMAX * 16 KB
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
  // column panel of V handled by this launch (TMEM holds 512 columns): sub-blocks [sb_lo, sb_hi) of 64 columns,
  // fed by the first c_count K* chunks; a model with n_pad <= 512 is one panel (0, n_chunks, n_chunks)
  int sb_lo, sb_hi, c_count;
  const float* vacc_in;   // |V|^2 partial of the earlier panels (null: none)
  float* vacc_out;        // non-null: this is not the last panel -- store the partial and stop
  const float* mc_table;  // K*-reading variant: qLogEI table built once per call by k_mc_table (null: per-sample loop)
  const float* kpre;  // wide-feature path: K* block already materialised by k_kmat_tc (else null)
  int64_t ldk;
  long long* trace;  // test-only event trace (bb_debug_set_trace); null in normal operation
  int trace_cap;
  // fused_ts.cu (A operand of the V contraction in tensor memory): operand images and folded constants
  const uint8_t* timg_l;   // L^-1 image: [hi tiles, chunk c = 0..C-1][lo tiles], tile c = (n_pad - 64c) rows x 64 k, SW128
  const uint8_t* timg_b;   // augmented training-row image: 3 splits x n_pad rows x 32 k (|b|^2 and 1 in k = 31, 30), SW64
  const float* ts_alpha;   // alpha / ts_kscale
  float ts_sa;             // power of two folded into the candidate rows of the A2 image
  float ts_aug_sq;         // A2 column 30 = |a|^2 * ts_aug_sq
  float ts_aug_one;        // A2 column 31 = ts_aug_one
  float ts_g;              // D2 * ts_g = scaled squared distance t (family constant folded in)
  float ts_kscale;         // power of two folded into K* before the fp16 hi/lo split
  // single-launch end-to-end pass (bb_score_fused_overlapped): rows arrive from the host WHILE the kernel runs
  const unsigned* ready_rows;  // rows [0, *ready_rows) of x have landed (published by the copy stream); null: all
  int32_t* gate_status;        // set to 1 if a tile's rows were not published within the time-out
  const float* code_table;     // level-coded layouts: value table [d][code_table_ld]
  int code_table_ld;
};

// candidate layouts beyond bb_layout, internal to the overlapped host pass: level codes expanded in the staging step
constexpr int kLayoutCodes4 = 16, kLayoutCodes8 = 17;

// optional request threaded through launch_fused by the overlapped host pass (null: none)
struct StreamGate {
  const unsigned* ready_rows;
  int32_t* status;
  const float* code_table;
  int code_table_ld;
  int layout;  // kLayoutCodes4 / kLayoutCodes8 / BB_ROW_MAJOR_F32
};

// test-only: (event id, SM clock) pairs of CTA 0 for a few tiles, to reconstruct the pipeline timeline
__device__ __forceinline__ void trace_ev(const FusedParams& p, int it, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && it >= 6 && it < 9) {
    const long long c = clock64();
    const unsigned long long i = atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull);
    if ((long long)i < p.trace_cap) {
      p.trace[1 + 2 * i] = (long long)it * 1000 + ev;
      p.trace[2 + 2 * i] = c;
    }
  }
}

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }


// Wait used by the single-lane helper warps: mbarrier.try_wait with a suspend-time hint parks the
// thread in hardware until the phase completes (wake-up ~60 cycles after the arrive) instead of
// polling, so the helpers neither burn issue slots nor add sleep-granularity latency.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(1000000u)
        : "memory");
  }
}

// 16 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// fp16 hi/mid/lo split of four fp32 values -> three 8-byte packets.
__device__ __forceinline__ void split3_quad(const float (&x)[4], uint2& hi, uint2& mid, uint2& lo) {
  __half2 h01 = __floats2half2_rn(x[0], x[1]), h23 = __floats2half2_rn(x[2], x[3]);
  float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const float r0 = x[0] - f01.x, r1 = x[1] - f01.y, r2 = x[2] - f23.x, r3 = x[3] - f23.y;
  __half2 m01 = __floats2half2_rn(r0, r1), m23 = __floats2half2_rn(r2, r3);
  float2 g01 = __half22float2(m01), g23 = __half22float2(m23);
  __half2 l01 = __floats2half2_rn(r0 - g01.x, r1 - g01.y), l23 = __floats2half2_rn(r2 - g23.x, r3 - g23.y);
  hi = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
  mid = make_uint2(*reinterpret_cast<uint32_t*>(&m01), *reinterpret_cast<uint32_t*>(&m23));
  lo = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
}

int launch_fused_tc(FusedParams& p, int grid, cudaStream_t stream);
int launch_fused_ts(FusedParams& p, int grid, cudaStream_t stream);
bool fused_ts_supported(const FusedParams& p, int max_smem);
bool kmat_ts_supported(const FusedParams& p, const float* d_k, int64_t ldk, int max_smem);
int launch_kmat_ts(FusedParams& p, float* d_k, int64_t ldk, int n_cols, int grid, cudaStream_t stream);
int try_kmat_ts(const bb_model* m, const void* d_x, int32_t layout, int64_t N, int64_t ldx, float* d_k, int64_t ldk,
                cudaStream_t stream, bool* handled);
// elect.sync: true in exactly one lane of a converged warp.  tcgen05.mma / cp.async.bulk / tcgen05.commit issued
// under this predicate compile to straight uniform-datapath code; the same instructions under `if (lane == 0)` are
// wrapped by ptxas in an ELECT / R2UR / BRA.U.ANY loop that costs ~106 cycles per MMA (scripts/ubench/mma_rate.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
bool fused_tc_supported(FusedParams& p, int max_smem);
int launch_pend_images(const bb_model* m, int32_t layout, const float* d_pend_x, int32_t P, cudaStream_t stream);
int launch_cross_wide(const bb_model* m, const void* d_x, int32_t layout, int64_t nb, int64_t ldx,
                      const float* d_pend_beta, int32_t P, float* d_cross_blk, cudaStream_t stream);
// pending-point request threaded through launch_fused on the wide path (null members: none)
struct WideCross {
  const float* pend_x;
  const float* pend_beta;
  int32_t P;
  float* cross;
};
// true when the single-launch gated pass can run this model (headline kernel envelope)
bool fused_gate_supported(const bb_model* m, const bb_acq_spec* acq, int32_t S);
int launch_kmat_wide(const bb_model* m, const void* d_x, int32_t layout, int64_t N, int64_t ldx,
                     float* d_out, int64_t ldk, int64_t out_rows, int out_cols, cudaStream_t stream);

}  // namespace bb
