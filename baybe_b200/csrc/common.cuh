// common.cuh -- shared helpers: error plumbing, PTX wrappers for sm_100a (mbarrier, bulk copy
// (TMA engine), tcgen05 MMA / TMEM), fp16 split, packed arg-max keys.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/baybe_b200.h"

namespace bb {

// ------------------------------------------------------------------------------------------
// error plumbing (thread-local message, integer status)
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define BB_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      bb::set_error(__VA_ARGS__);      \
      return BB_ERR_INVALID;           \
    }                                  \
  } while (0)

#define BB_CHECK_SUPPORTED(cond, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      bb::set_error(__VA_ARGS__);      \
      return BB_ERR_UNSUPPORTED;       \
    }                                  \
  } while (0)

#define BB_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      bb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                  \
                    cudaGetErrorString(e_));                                              \
      return BB_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define BB_LAUNCH_CHECK() BB_CUDA(cudaGetLastError())

// SM count and opt-in shared-memory limit of the current device, queried once per device (not per launch).
static inline int device_limits(int* sms, int* max_smem) {
  constexpr int kMaxDev = 64;
  static int c_sms[kMaxDev], c_smem[kMaxDev];
  static bool c_ok[kMaxDev];
  int dev = 0;
  BB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev || !c_ok[dev]) {
    int a = 0, b = 0;
    BB_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    BB_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 0 || dev >= kMaxDev) {
      *sms = a;
      *max_smem = b;
      return BB_OK;
    }
    c_sms[dev] = a;
    c_smem[dev] = b;
    c_ok[dev] = true;
  }
  *sms = c_sms[dev];
  *max_smem = c_smem[dev];
  return BB_OK;
}

// Opt a kernel into the full dynamic shared-memory carve-out once per device (function-local static per
// instantiation site), instead of one cudaFuncSetAttribute per launch.
#define BB_SMEM_OPTIN_ONCE(kernel)                                                                         \
  do {                                                                                                     \
    static unsigned long long done_mask_ = 0ull;                                                           \
    int dev_ = 0;                                                                                          \
    BB_CUDA(cudaGetDevice(&dev_));                                                                         \
    if (dev_ < 0 || dev_ >= 64 || !((done_mask_ >> dev_) & 1ull)) {                                        \
      BB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));      \
      if (dev_ >= 0 && dev_ < 64) done_mask_ |= 1ull << dev_;                                              \
    }                                                                                                      \
  } while (0)

constexpr int kTileM = 128;     // candidates per tile = UMMA_M
constexpr int kChunk = 64;      // training points per K chunk = one 128-byte swizzle row of fp16
constexpr int kSMs = 148;

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------
// candidate loads for the four layouts of bb_layout
// ------------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ __forceinline__ float load_x(const void* __restrict__ x, int64_t row, int col,
                                        int64_t ld) {
  if constexpr (LAYOUT == BB_ROW_MAJOR_F32) {
    return __ldg(reinterpret_cast<const float*>(x) + row * ld + col);
  } else if constexpr (LAYOUT == BB_COL_MAJOR_F32) {
    return __ldg(reinterpret_cast<const float*>(x) + (int64_t)col * ld + row);
  } else if constexpr (LAYOUT == BB_ROW_MAJOR_F64) {
    return (float)__ldg(reinterpret_cast<const double*>(x) + row * ld + col);
  } else {
    return (float)__ldg(reinterpret_cast<const double*>(x) + (int64_t)col * ld + row);
  }
}

// ------------------------------------------------------------------------------------------
// kernel epilogues k(r^2) -- the scaled squared distance t already carries the family's
// constant (5 r^2 for Matern-5/2, 3 r^2 for 3/2, r^2 for 1/2, r^2 log2(e)/2 for RBF), folded
// into the lengthscale by bb_model_build.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <int FAMILY>
__device__ __forceinline__ float kernel_from_t(float t) {
  t = fmaxf(t, 0.0f);
  if constexpr (FAMILY == BB_KERNEL_RBF) {
    return fast_ex2(-t);
  } else if constexpr (FAMILY == BB_KERNEL_MATERN12) {
    float s = fast_sqrt(t);
    return fast_ex2(-kLog2e * s);
  } else if constexpr (FAMILY == BB_KERNEL_MATERN32) {
    float s = fast_sqrt(t);
    return (1.0f + s) * fast_ex2(-kLog2e * s);
  } else {
    float s = fast_sqrt(t);
    float poly = fmaf(t, (1.0f / 3.0f), s) + 1.0f;
    return poly * fast_ex2(-kLog2e * s);
  }
}

// ------------------------------------------------------------------------------------------
// packed (score, lowest-index) keys: signed-int64 max == (max score, then min index)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t pack_key(float score, uint32_t local_idx) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(score);
#else
  memcpy(&u, &score, 4);
#endif
  int32_t s = (int32_t)u;
  s ^= (s >> 31) & 0x7fffffff;  // total order as signed int
  return (int64_t)(((uint64_t)(uint32_t)s << 32) | (uint64_t)(0xffffffffu - local_idx));
}
constexpr int64_t kEmptyKey = INT64_MIN;

// ------------------------------------------------------------------------------------------
// PTX: shared-memory addresses, mbarrier, fences, bulk copies (TMA engine)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared through the TMA engine (UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------
// PTX: tcgen05 (5th-gen tensor cores) + tensor memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 operands, fp32 accumulate), one CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32 from TMEM: thread i gets row (lane base + i), 32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 64 fp16 (128 B),
// 8-row groups 1024 B apart (SBO), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);  // start address, 16-byte units
  d |= (uint64_t)0 << 16;                       // leading byte offset (unused, one K atom)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                       // version = 1
  d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: A,B = fp16 (format 0), D = fp32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4)                    // c_format = F32
         | (0u << 7) | (0u << 10)     // a_format = b_format = F16
         | (0u << 15) | (0u << 16)    // a_major = b_major = K
         | ((uint32_t)(N >> 3) << 17) // n_dim
         | ((uint32_t)(M >> 4) << 24);// m_dim
}
// Byte offset of 16-byte chunk `c16` (0..7) of row `r` inside a 128B-swizzled tile whose rows are
// 128 bytes (64 fp16): Swizzle<3,4,3> -- XOR the chunk index with (row mod 8).
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c16) {
  return r * 128u + ((c16 ^ (r & 7u)) << 4);
}

// K-major tiles whose rows hold K2 fp16 (K2 = 64 -> 128-byte rows / SWIZZLE_128B, K2 = 32 ->
// 64-byte rows / SWIZZLE_64B).  8-row groups are contiguous (SBO = 8 * row bytes).
template <int K2>
__host__ __device__ __forceinline__ uint32_t swk_offset(uint32_t r, uint32_t c16) {
  if constexpr (K2 == 64) return r * 128u + ((c16 ^ (r & 7u)) << 4);
  else return r * 64u + ((c16 ^ ((r >> 1) & 3u)) << 4);  // Swizzle<2,4,3>: bits[5:4] ^= bits[8:7]
}
template <int K2>
__device__ __forceinline__ uint64_t make_swk_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)((K2 == 64 ? 1024 : 512) >> 4) << 32;  // stride between 8-row groups
  d |= (uint64_t)1 << 46;                                // descriptor version 1
  d |= (uint64_t)(K2 == 64 ? 2 : 4) << 61;               // SWIZZLE_128B : SWIZZLE_64B
  return d;
}

// fp16 hi/lo split of a non-negative-or-signed fp32 value: x ~= hi + lo, relative error 2^-22.
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(x0, x1);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace bb
