// fused_ts.cu -- fused scoring kernel, round-2 headline (n_pad <= 256, d <= 30, S <= 512, Matern-3/2, -5/2, RBF).
//
// Same mathematics as fused_tc.cu (GEMM-form distances like gpytorch's Distance._sq_dist, closed-form kernel,
// V = K* L^-T on tcgen05, sigma^2 = k** - |V|^2, q = 1 acquisition, packed-key arg-max); what changed is how the
// tensor pipe and the CUDA cores are fed -- every change follows a measurement (scripts/ubench/mma_rate.cu,
// profiles/r02_*):
//   * MMAs are issued by a CONVERGED warp under elect.sync.  Issued under `if (lane == 0)` every tcgen05.mma is
//     wrapped by ptxas in an ELECT/R2UR/BRA.U.ANY loop: ~106 cycles per instruction whatever its shape, which is
//     what bounded fused_tc.cu (146-182 cycles per MMA against 32-128 of math).
//   * The A operand of the V contraction (the converted K* chunk, fp16 hi/lo) lives in TENSOR MEMORY (TS form):
//     each thread overwrites the 16 D2 columns it just read with 8 packed hi + 8 packed lo columns.  No A ring in
//     shared memory, no per-MMA re-read of A from shared memory, no "slot empty" wait in the conversion loop.
//   * One MMA spans every output column a K step can reach (N = n_pad - 64c - 16kk <= 256): 68 MMAs per 128-row
//     tile instead of 84, and the triangular skip at 16-column granularity (8.5/16 of n^2 instead of 10/16).
//   * |a|^2 + |b|^2 ride in two spare K columns of the distance GEMM (A2 = [a, |a|^2 P, P1], Bt = [-2b, Q1, |b|^2 Q]),
//     so the accumulator holds the scaled squared distance itself and the per-value epilogue is
//     FMNMX, MUFU.SQRT, MUFU.EX2 and packed f32x2 arithmetic: 8 instructions per kernel value (19 before).
//   * The hi image of L^-1 and the training rows are RESIDENT in shared memory; only the lo image (80 KB per tile at
//     n_pad = 256) streams from L2 through a 3-stage bulk-copy ring.
//   * FOUR EPILOGUE WARPS (one per TMEM lane quarter, one candidate row per thread) read the V accumulator, form
//     the moments and evaluate the acquisition function while the 16 conversion warps are already converting the
//     next tile: with that work on the conversion warps a tile took 11.4 k cycles, 5.1 k of them |V|^2 +
//     acquisition.
//   * The MMA issue loop is instantiated with a compile-time chunk count for n_pad = 256: the descriptor arithmetic
//     in front of every MMA batch (~250 cycles on the one issuing thread, twelve batches per tile) becomes immediates.
//   * The qLogEI table is built once per call (k_mc_table_grid, overlapped with this kernel's prologue by programmatic
//     dependent launch) instead of once per persistent CTA (~23 us of every launch).
//   * Gated end-to-end pass (bb_score_fused_overlapped): the kernel can be launched over rows that are still on
//     their way from the host; one thread per CTA watches the copy stream's publication counter.
// Tried on top of this and REJECTED by same-box A/B timing (profiles/r02_ab_kernel_variants.txt): serving the next
// tile's accumulator between whole-warp rows (+7 %), an incremental accumulator reader with FMA-pipe reciprocals
// (+5 %), per-K-step commits of the last sub-block, lowest warp ids for the single-warp roles (no effect), a
// two-stage accumulator release with a column-split first chunk (+5 %): each shortens one wait in the event trace and
// lengthens the tensor pipe's own time (more, smaller MMAs; more tensor-memory traffic) by more.
// 704 threads: warps 0-15 conversion, 16-19 epilogue, 20 bulk-copy producer, 21 MMA issuer (80 registers each: the
// allocation unit is 512 registers per warp, so 22 warps cap a thread at 80).
//
// Tensor memory: columns [0, n_pad) V accumulator; [256, 256 + n_pad) D2, overwritten in place by the A operand.
// Reference path replaced: see fused.cu.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "assemble.cuh"
#include "fused_common.cuh"

namespace bb {

constexpr int kTsK2 = 32;                    // K extent of the augmented distance GEMM: d data columns + 2
constexpr int kTsColSq = 30, kTsColOne = 31; // A2 columns holding |a|^2 * P and P1 (Bt holds Q1 and |b|^2 * Q there)
constexpr uint32_t kTsLoStage = 16384;       // one streamed piece of the lo image: <= 128 rows x 64 k fp16
constexpr int kTsLoStages = 3;
constexpr uint32_t kTsD2Col0 = 256;
constexpr uint32_t kTsA2Split = 128u * kTsK2 * 2u;  // one fp16 panel of the candidate tile: 8 KB
// Warp roles of k_fused_ts: 16 conversion warps (stage candidate rows, D2 -> K* operand), 4 epilogue warps (one per
// TMEM lane quarter: |V|^2, moments, acquisition, arg-max), the bulk-copy producer and the MMA issuer.
constexpr int kTsEpiWarp0 = kComputeWarps, kTsEpiWarps = 4;
constexpr int kTsWarpProducer = kTsEpiWarp0 + kTsEpiWarps, kTsWarpMma = kTsWarpProducer + 1;
constexpr int kTsThreads = (kTsWarpMma + 1) * 32;  // 704
constexpr int kTsTaskSlots = 4;                    // candidate task ids of the tiles in flight

struct TsSmem {
  uint8_t *lh, *ring, *bt, *a2;
  float *alpha_s, *z_s, *mc_tab, *tcov, *meanc, *cscale_s, *cshift_s, *an_part, *mean_part, *var_part;
  int32_t *ttask, *cand_task;
  uint64_t *a_full, *vsub_full, *r_full, *r_empty, *d2_full, *a2_full, *v_empty, *res_full, *mean_full;
  long long* best_red;
  uint32_t* tmem_ptr;
  float* zstat;
  volatile unsigned* ready_cache;
};

__host__ __device__ inline uint32_t ts_hi_bytes(int n_pad) {  // sum over chunks of (n_pad - 64c) rows x 128 B
  uint32_t b = 0;
  for (int r = n_pad; r > 0; r -= kChunk) b += (uint32_t)r * 128u;
  return b;
}

__host__ __device__ inline size_t ts_carve(uint8_t* base, int n_pad, int n_tasks, TsSmem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) {
    off = (off + align - 1) / align * align;
    size_t o = off;
    off += bytes;
    return o;
  };
  const size_t o_lh = take(ts_hi_bytes(n_pad), 1024);
  const size_t o_ring = take((size_t)kTsLoStages * kTsLoStage, 1024);
  const size_t o_bt = take((size_t)3 * n_pad * kTsK2 * 2, 1024);
  const size_t o_a2 = take((size_t)3 * kTsA2Split, 1024);
  const size_t o_al = take((size_t)n_pad * 4, 16), o_tt = take((size_t)n_pad * 4, 16);
  const size_t o_z = take(512 * 4, 16), o_mc = take(1024 * 4, 16);
  const size_t o_tc = take((size_t)kMaxTasks * kMaxTasks * 4, 16), o_mcn = take(kMaxTasks * 4, 16);
  const size_t o_cs = take(32 * 4, 16), o_sh = take(32 * 4, 16);
  const size_t o_an = take(4 * kTileM * 4, 16);
  const size_t o_mp = take(2 * 4 * kTileM * 4, 16), o_vp = take(16, 16);
  const size_t o_ct = take(kTsTaskSlots * kTileM * 4, 16);
  const size_t o_bar = take(32 * 8, 16);
  const size_t o_best = take(16 * 8, 16), o_misc = take(32, 16);
  (void)n_tasks;
  if (s) {
    s->lh = base + o_lh;
    s->ring = base + o_ring;
    s->bt = base + o_bt;
    s->a2 = base + o_a2;
    s->alpha_s = reinterpret_cast<float*>(base + o_al);
    s->ttask = reinterpret_cast<int32_t*>(base + o_tt);
    s->z_s = reinterpret_cast<float*>(base + o_z);
    s->mc_tab = reinterpret_cast<float*>(base + o_mc);
    s->tcov = reinterpret_cast<float*>(base + o_tc);
    s->meanc = reinterpret_cast<float*>(base + o_mcn);
    s->cscale_s = reinterpret_cast<float*>(base + o_cs);
    s->cshift_s = reinterpret_cast<float*>(base + o_sh);
    s->an_part = reinterpret_cast<float*>(base + o_an);
    s->mean_part = reinterpret_cast<float*>(base + o_mp);
    s->var_part = reinterpret_cast<float*>(base + o_vp);
    s->cand_task = reinterpret_cast<int32_t*>(base + o_ct);
    uint64_t* b = reinterpret_cast<uint64_t*>(base + o_bar);
    s->a_full = b;          // [4]
    s->vsub_full = b + 4;   // [4]
    s->r_full = b + 8;      // [kTsLoStages]
    s->r_empty = b + 12;    // [kTsLoStages]
    s->d2_full = b + 16;    // [2]: one per 128-column D2 slab
    s->a2_full = b + 18;
    s->v_empty = b + 19;
    s->res_full = b + 20;
    s->mean_full = b + 21;  // [2]
    s->best_red = reinterpret_cast<long long*>(base + o_best);
    s->tmem_ptr = reinterpret_cast<uint32_t*>(base + o_misc);
    s->zstat = reinterpret_cast<float*>(base + o_misc + 8);
    s->ready_cache = reinterpret_cast<volatile unsigned*>(base + o_misc + 16);
  }
  return off;
}

// ---- PTX used only here ------------------------------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]^T: A = 128 lanes x 8 columns of packed fp16 pairs (16 K values) per instruction.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {  // cvt.rn.f16x2.f32: first source -> upper half
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float h_lo_f32(uint32_t h2) {
  float f;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(f) : "r"(h2));
  return f;
}
__device__ __forceinline__ float h_hi_f32(uint32_t h2) {
  float f;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}" : "=f"(f) : "r"(h2));
  return f;
}
__device__ __forceinline__ void bar_quarter_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_quarter_arrive(int id) { asm volatile("bar.arrive %0, 128;" ::"r"(id) : "memory"); }

// Test-only pipeline trace without atomics (the atomicAdd of trace_ev costs ~700 cycles per event and distorts what
// it measures): the compute thread 0 and the MMA lane 0 of CTA 0 each append (code, clock) pairs to their own half
// of the buffer with plain stores; slot 0 of the buffer is set to the capacity at the end (unused pairs stay zero).
template <bool TRACE>
__device__ __forceinline__ void ts_trace(const FusedParams& p, int half, int& n, int it, int ev) {
  if constexpr (!TRACE) return;
  if (p.trace != nullptr && blockIdx.x == 0 && it >= 6 && it < 9) {
    const int cap2 = p.trace_cap / 3;  // three writers: conversion thread 0, MMA lane 0, epilogue thread 0
    if (n < cap2) {
      const int i = half * cap2 + n;
      p.trace[1 + 2 * i] = (long long)it * 1000 + ev;
      p.trace[2 + 2 * i] = clock64();
      ++n;
    }
  }
}

struct TsStageRegs {
  float4 v[2];
};

// Rows that arrive while the kernel runs (gated pass) are read with ld.global.cg: coherent at L2, where the copy
// engine's writes land; ld.global.nc (__ldg) presumes data that is constant for the kernel's lifetime.
__device__ __forceinline__ float4 ts_load_quad(const FusedParams& p, int64_t row, int jq) {
  const int j0 = jq * 4;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= p.N || j0 >= p.d) return q;
  switch (p.layout) {
    case kLayoutCodes4: {  // two columns per byte, low nibble = even column (bb_decode_codes)
      const uint8_t* c = reinterpret_cast<const uint8_t*>(p.x) + row * p.ldx + (j0 >> 1);
      const uint32_t b0 = __ldcg(c), b1 = (j0 + 2 < p.d) ? (uint32_t)__ldcg(c + 1) : 0u;
      const float* t = p.code_table + (size_t)j0 * p.code_table_ld;
      q.x = __ldg(t + (b0 & 15u));
      if (j0 + 1 < p.d) q.y = __ldg(t + p.code_table_ld + (b0 >> 4));
      if (j0 + 2 < p.d) q.z = __ldg(t + 2 * p.code_table_ld + (b1 & 15u));
      if (j0 + 3 < p.d) q.w = __ldg(t + 3 * p.code_table_ld + (b1 >> 4));
      return q;
    }
    case kLayoutCodes8: {
      const uint8_t* c = reinterpret_cast<const uint8_t*>(p.x) + row * p.ldx + j0;
      const float* t = p.code_table + (size_t)j0 * p.code_table_ld;
      q.x = __ldg(t + __ldcg(c));
      if (j0 + 1 < p.d) q.y = __ldg(t + p.code_table_ld + __ldcg(c + 1));
      if (j0 + 2 < p.d) q.z = __ldg(t + 2 * p.code_table_ld + __ldcg(c + 2));
      if (j0 + 3 < p.d) q.w = __ldg(t + 3 * p.code_table_ld + __ldcg(c + 3));
      return q;
    }
    case BB_ROW_MAJOR_F32: {
      const float* ptr = reinterpret_cast<const float*>(p.x) + row * p.ldx + j0;
      if (p.ready_rows != nullptr) {
        if (j0 + 3 < p.d && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)) return __ldcg(reinterpret_cast<const float4*>(ptr));
        q.x = __ldcg(ptr);
        if (j0 + 1 < p.d) q.y = __ldcg(ptr + 1);
        if (j0 + 2 < p.d) q.z = __ldcg(ptr + 2);
        if (j0 + 3 < p.d) q.w = __ldcg(ptr + 3);
        return q;
      }
      if (j0 + 3 < p.d && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)) return __ldg(reinterpret_cast<const float4*>(ptr));
      q.x = __ldg(ptr);
      if (j0 + 1 < p.d) q.y = __ldg(ptr + 1);
      if (j0 + 2 < p.d) q.z = __ldg(ptr + 2);
      if (j0 + 3 < p.d) q.w = __ldg(ptr + 3);
      return q;
    }
    case BB_COL_MAJOR_F32:
      q.x = load_x<BB_COL_MAJOR_F32>(p.x, row, j0, p.ldx);
      if (j0 + 1 < p.d) q.y = load_x<BB_COL_MAJOR_F32>(p.x, row, j0 + 1, p.ldx);
      if (j0 + 2 < p.d) q.z = load_x<BB_COL_MAJOR_F32>(p.x, row, j0 + 2, p.ldx);
      if (j0 + 3 < p.d) q.w = load_x<BB_COL_MAJOR_F32>(p.x, row, j0 + 3, p.ldx);
      return q;
    case BB_ROW_MAJOR_F64:
      q.x = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0, p.ldx);
      if (j0 + 1 < p.d) q.y = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0 + 1, p.ldx);
      if (j0 + 2 < p.d) q.z = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0 + 2, p.ldx);
      if (j0 + 3 < p.d) q.w = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0 + 3, p.ldx);
      return q;
    default:
      q.x = load_x<BB_COL_MAJOR_F64>(p.x, row, j0, p.ldx);
      if (j0 + 1 < p.d) q.y = load_x<BB_COL_MAJOR_F64>(p.x, row, j0 + 1, p.ldx);
      if (j0 + 2 < p.d) q.z = load_x<BB_COL_MAJOR_F64>(p.x, row, j0 + 2, p.ldx);
      if (j0 + 3 < p.d) q.w = load_x<BB_COL_MAJOR_F64>(p.x, row, j0 + 3, p.ldx);
      return q;
  }
}

// Gated pass: block until the copy stream has published the rows of `tile` (acquire at system scope: the data
// was written by the copy engine before the counter).  Bounded: after ~2 s the status word is raised and the
// kernel carries on (the host reports the pass as failed) -- a missing publication must not hang the GPU.
// ONE thread per CTA polls: every thread of every CTA reading the same word made that L2 line the bottleneck of
// the whole pass (75,000 requests per tile step).  The value it sees is handed to the other conversion threads
// through shared memory + the conversion warps' barrier, and kept in a register: publications run far ahead of
// the tiles, so most tiles need no load at all.  Called by all 512 conversion threads (uniform `have`).
__device__ __forceinline__ unsigned ts_wait_rows(const FusedParams& p, volatile unsigned* cache_s, int tile, unsigned have) {
  const long long last = (long long)(tile + 1) * kTileM;
  const unsigned need = (unsigned)(last < p.N ? last : p.N);
  if (have >= need) return have;
  if (threadIdx.x == 0) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ready_rows) : "memory");
    if (v < need) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      while (true) {
        __nanosleep(256);
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ready_rows) : "memory");
        if (v >= need) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) {
          if (p.gate_status != nullptr) *p.gate_status = 1;
          v = 0xffffffffu;  // give up waiting for good: the pass is reported as failed
          break;
        }
      }
    }
    *cache_s = v;
  }
  bar_compute();  // the 512 conversion threads; also orders their loads of the rows behind thread 0's acquire
  const unsigned v = *cache_s;
  bar_compute();  // the word may be rewritten by the next call
  return v;
}

// Sixteen kernel values (scaled by ts_kscale) from sixteen accumulator values D = t / g, two at a time in packed
// f32x2 arithmetic.  Matern-5/2: k = (1 + s + t/3) e^-s, s = sqrt(t) = sqrt(g) sqrt(D); Matern-3/2: (1 + s) e^-s;
// RBF: 2^-t (the family constants 5 / 3 / log2(e)/2 are folded into the lengthscales by bb_model_build).
struct TsConsts {
  unsigned long long k0, c1, c2, c3;  // packed pairs: ks, ks*sqrt(g), ks*g/3, -log2(e)*sqrt(g) (RBF: c3 = -g)
};

template <int FAMILY>
__device__ __forceinline__ unsigned long long ts_kernel_pair(float d0, float d1, const TsConsts& cst) {
  d0 = fmaxf(d0, 0.f);
  d1 = fmaxf(d1, 0.f);
  if constexpr (FAMILY == BB_KERNEL_RBF) {
    const unsigned long long ea = mul2(pack2(d0, d1), cst.c3);
    return mul2(pack2(fast_ex2(lo_of(ea)), fast_ex2(hi_of(ea))), cst.k0);
  } else {
    const unsigned long long q = pack2(fast_sqrt(d0), fast_sqrt(d1));
    const unsigned long long ea = mul2(q, cst.c3);
    unsigned long long poly = fma2(q, cst.c1, cst.k0);
    if constexpr (FAMILY == BB_KERNEL_MATERN52) poly = fma2(pack2(d0, d1), cst.c2, poly);
    return mul2(poly, pack2(fast_ex2(lo_of(ea)), fast_ex2(hi_of(ea))));
  }
}

// ---- MMA issuer ----------------------------------------------------------------------------------------------
// The whole warp runs the loop (converged), one elected lane issues.  Order on the tensor pipe:
//   DIST(0,s0) DIST(0,s1) | V(t,c0) V(t,c1) DIST(t+1,s0) V(t,c2) V(t,c3) DIST(t+1,s1) | ...
// (a slab of D2 is overwritten by the next tile's distance GEMM right behind the V MMAs that read the A operand
// stored in it -- the pipe executes in issue order).
// CC > 0 fixes the number of 64-row K chunks at compile time (n_pad = 64 CC): every column count, instruction
// descriptor and operand offset below becomes an immediate.  With n_pad read from the parameter block the
// descriptor arithmetic in front of each batch of MMAs ran ~250 cycles on the single issuing thread, twelve batches
// per tile -- a quarter of the tile time with the tensor pipe idle (profiles/r02_pipeline_trace_fused_ts_v3.txt).
template <int CC, bool TRACE>
__device__ __forceinline__ void ts_mma_role(const FusedParams& p, const TsSmem& s, uint32_t tmem_base, int lane) {
  const int n_pad = CC > 0 ? CC * kChunk : p.n_pad;
  const int C = CC > 0 ? CC : p.n_chunks;
  const int C0 = C < 2 ? C : 2;
  const uint32_t bt_split = (uint32_t)n_pad * kTsK2 * 2u;
  const uint32_t a2_addr = smem_u32(s.a2), bt_addr = smem_u32(s.bt), lh_addr = smem_u32(s.lh);
  const uint32_t ring_addr = smem_u32(s.ring);
  const uint64_t a2_h = make_swk_desc<kTsK2>(a2_addr), a2_m = make_swk_desc<kTsK2>(a2_addr + kTsA2Split),
                 a2_l = make_swk_desc<kTsK2>(a2_addr + 2 * kTsA2Split);
  // descriptors differ from these bases only in the 14-bit start-address field (16-byte units; the operand images
  // lie well inside the 256 KB window, so the additions below cannot carry out of the field)
  const uint64_t bt_h0 = make_swk_desc<kTsK2>(bt_addr), lh_d0 = make_sw128_desc(lh_addr), ring_d0 = make_sw128_desc(ring_addr);
  const uint32_t d2_addr = tmem_base + kTsD2Col0;
  int mma_it = 0, mma_n = 0;
  // distance GEMM of one slab (training columns [128 slab, ...)): six split products (2^-33)
  auto issue_distance = [&](int slab) {
    const int ncols = (n_pad - 128 * slab) < 128 ? (n_pad - 128 * slab) : 128;
    const uint32_t idesc = make_idesc_f16(kTileM, ncols);
    const uint64_t b_h = bt_h0 + (uint64_t)(((uint32_t)slab * 128u * (kTsK2 * 2u)) >> 4);
    const uint64_t b_m = b_h + (uint64_t)(bt_split >> 4), b_l = b_m + (uint64_t)(bt_split >> 4);
    const uint32_t d_addr = d2_addr + (uint32_t)(128 * slab);
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < kTsK2 / 16; ++kk) {
        const uint64_t ko = (uint64_t)(kk * 2);
        umma_f16(d_addr, a2_h + ko, b_h + ko, idesc, kk > 0 ? 1u : 0u);
        umma_f16(d_addr, a2_h + ko, b_m + ko, idesc, 1u);
        umma_f16(d_addr, a2_m + ko, b_h + ko, idesc, 1u);
        umma_f16(d_addr, a2_h + ko, b_l + ko, idesc, 1u);
        umma_f16(d_addr, a2_l + ko, b_h + ko, idesc, 1u);
        umma_f16(d_addr, a2_m + ko, b_m + ko, idesc, 1u);
      }
      umma_commit(&s.d2_full[slab]);
    }
    __syncwarp();
    if (lane == 0) ts_trace<TRACE>(p, 1, mma_n, mma_it, 250 + slab);
  };
  mbar_wait_relaxed(s.res_full, 0u);
  uint32_t rs = 0, rph = 0;
  // V MMAs of K chunk c: resident hi image (K*hi x Lhi, K*lo x Lhi), then the streamed lo pieces (K*hi x Llo)
  auto issue_v_chunk = [&](int c, uint32_t par, uint32_t hi_off) {
    const int rows_c = n_pad - c * kChunk;
    mbar_wait_relaxed(&s.a_full[c], par);
    tc_fence_after();
    if (lane == 0) ts_trace<TRACE>(p, 1, mma_n, mma_it, 210 + c);
    const uint32_t a_col = d2_addr + (uint32_t)(c * kChunk);
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // K step kk only reaches columns >= 64c + 16kk
        const uint32_t n_cols = (uint32_t)(rows_c - 16 * kk);
        const uint64_t bd = lh_d0 + (uint64_t)((hi_off + (uint32_t)kk * 2048u) >> 4) + (uint64_t)(kk * 2);
        const uint32_t d_addr = tmem_base + (uint32_t)(c * kChunk + 16 * kk);
        const uint32_t id = make_idesc_f16(kTileM, (int)n_cols);
        umma_f16_ts(d_addr, a_col + (uint32_t)(16 * kk), bd, id, (c > 0 || kk > 0) ? 1u : 0u);
        umma_f16_ts(d_addr, a_col + (uint32_t)(16 * kk + 8), bd, id, 1u);
      }
    }
    __syncwarp();
#pragma unroll
    for (int r0 = 0; r0 < rows_c; r0 += 128) {
      const int rows_p = rows_c - r0 < 128 ? rows_c - r0 : 128;
      mbar_wait_relaxed(&s.r_full[rs], rph);
      tc_fence_after();
      const uint64_t b_d = ring_d0 + (uint64_t)((rs * kTsLoStage) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int skip = (r0 == 0) ? 16 * kk : 0;  // rows of this piece the K step cannot reach
          const uint32_t n_cols = (uint32_t)(rows_p - skip);
          const uint64_t bd = b_d + (uint64_t)(((uint32_t)skip * 128u) >> 4) + (uint64_t)(kk * 2);
          const uint32_t d_addr = tmem_base + (uint32_t)(c * kChunk + r0 + skip);
          umma_f16_ts(d_addr, a_col + (uint32_t)(16 * kk), bd, make_idesc_f16(kTileM, (int)n_cols), 1u);
        }
        umma_commit(&s.r_empty[rs]);
      }
      __syncwarp();
      if (++rs == (uint32_t)kTsLoStages) {
        rs = 0;
        rph ^= 1u;
      }
    }
    if (elect_one()) umma_commit(&s.vsub_full[c]);  // sub-block c of V has received its last contribution
    __syncwarp();
    if (lane == 0) ts_trace<TRACE>(p, 1, mma_n, mma_it, 220 + c);
  };
  int j = 0;
  int tile = blockIdx.x;
  if (tile < p.num_tiles) {
    mbar_wait_relaxed(s.a2_full, 0u);
    tc_fence_after();
    issue_distance(0);
    if (C > 2) issue_distance(1);
  }
  for (; tile < p.num_tiles; tile += gridDim.x, ++j) {
    const uint32_t par = (uint32_t)(j & 1);
    const bool has_next = tile + (int)gridDim.x < p.num_tiles;
    mma_it = j;
    if (j > 0) mbar_wait_relaxed(s.v_empty, par ^ 1u);  // the previous tile's |V|^2 has been read
    tc_fence_after();
    if (lane == 0) ts_trace<TRACE>(p, 1, mma_n, j, 200);
    uint32_t hi_off = 0;
#pragma unroll
    for (int c = 0; c < (CC > 0 ? (CC < 2 ? CC : 2) : 2); ++c) {
      if (c < C0) {
        issue_v_chunk(c, par, hi_off);
        hi_off += (uint32_t)(n_pad - c * kChunk) * 128u;
      }
    }
    if (has_next) {
      mbar_wait_relaxed(s.a2_full, par ^ 1u);  // the next tile's candidate rows are staged
      tc_fence_after();
      if (lane == 0) ts_trace<TRACE>(p, 1, mma_n, j, 260);
      issue_distance(0);
    }
#pragma unroll
    for (int c = 2; c < (CC > 0 ? CC : 4); ++c) {
      if (c < C) {
        issue_v_chunk(c, par, hi_off);
        hi_off += (uint32_t)(n_pad - c * kChunk) * 128u;
      }
    }
    if (has_next && C > 2) issue_distance(1);
  }
}

template <int FAMILY, bool TASKS, bool TRACE>
__global__ void __launch_bounds__(kTsThreads, 1) k_fused_ts(const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TsSmem s;
  ts_carve(smem_raw, p.n_pad, p.n_tasks, &s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.n_chunks;
  const uint32_t hi_bytes = ts_hi_bytes(p.n_pad);
  const uint32_t bt_split = (uint32_t)p.n_pad * kTsK2 * 2u;
  if (tid == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();

  // ---- one-time setup ----
  if (warp == kTsWarpMma && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s.a_full[i], kComputeWarps);
      mbar_init(&s.vsub_full[i], 1);
    }
    for (int i = 0; i < kTsLoStages; ++i) {
      mbar_init(&s.r_full[i], 1);
      mbar_init(&s.r_empty[i], 1);
    }
    mbar_init(&s.d2_full[0], 1);
    mbar_init(&s.d2_full[1], 1);
    mbar_init(s.a2_full, kComputeWarps);
    mbar_init(s.v_empty, kTsEpiWarps);
    mbar_init(s.res_full, 1);
    mbar_init(&s.mean_full[0], kComputeWarps);
    mbar_init(&s.mean_full[1], kComputeWarps);
    fence_mbar_init();
  }
  if (warp == kTsWarpProducer) {
    tmem_alloc(s.tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int e = tid; e < (int)(3 * kTsA2Split / 16); e += kTsThreads)  // unused K columns stay zero for good
    reinterpret_cast<uint4*>(s.a2)[e] = make_uint4(0u, 0u, 0u, 0u);
  for (int e = tid; e < 32; e += kTsThreads) {
    s.cscale_s[e] = e < p.d_pad ? __ldg(p.cand_scale + e) : 0.f;
    s.cshift_s[e] = e < p.d_pad ? __ldg(p.cand_shift + e) : 0.f;
  }
  for (int e = tid; e < p.n_pad; e += kTsThreads) {
    s.alpha_s[e] = __ldg(p.ts_alpha + e);
    s.ttask[e] = TASKS ? __ldg(p.train_task + e) : 0;
  }
  for (int e = tid; e < p.n_tasks * p.n_tasks; e += kTsThreads) s.tcov[e] = __ldg(p.task_covar + e);
  for (int e = tid; e < p.n_tasks; e += kTsThreads) s.meanc[e] = __ldg(p.mean_const + e);
  for (int e = tid; e < kTsTaskSlots * kTileM; e += kTsThreads) s.cand_task[e] = 0;
  if (p.has_acq && p.z != nullptr)
    for (int e = tid; e < p.S; e += kTsThreads) s.z_s[e] = __ldg(p.z + e);
  fence_proxy_async();  // the zero-filled A2 tile is read by the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;
  if (p.has_acq && warp == 0) {
    float sz = 0.f, sa = 0.f;
    for (int e = lane; e < p.S; e += 32) sz += s.z_s[e];
    for (int o = 16; o > 0; o >>= 1) sz += __shfl_xor_sync(0xffffffffu, sz, o);
    const float zm = sz / (float)p.S;
    for (int e = lane; e < p.S; e += 32) sa += fabsf(s.z_s[e] - zm);  // qUCB: deviations from the SAMPLE mean
    for (int o = 16; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
    if (lane == 0) {
      s.zstat[0] = zm;
      s.zstat[1] = sa / (float)p.S;
    }
  }
  const bool fast_mc = mc_table_applicable(p.has_acq, p.acq, p.S);
  if (fast_mc && p.mc_table != nullptr) {  // built once per call by k_mc_table_grid
    // programmatic dependent launch: everything above ran while the table kernel was still in flight
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int e = tid; e < kMcRows; e += kTsThreads) s.mc_tab[e] = __ldcg(p.mc_table + e);
    __syncthreads();
  } else if (fast_mc) {
    mc_table_setup(s.mc_tab, s.z_s, p.S, p.acq.obj_scale < 0.f ? -1.f : 1.f);  // contains __syncthreads
  } else {
    __syncthreads();  // zstat visible to everyone
  }

  if (warp < kComputeWarps) {
    // =====================================================================================================
    // conversion warps.  Thread = TMEM lane (candidate row_e of the tile) x column group cg (16 of the 64
    // columns of a chunk): candidate rows -> A2 operand panels, D2 -> kernel values -> K* operand in TMEM,
    // mean partials sum_i k_i alpha_i -> shared memory for the epilogue warps.
    // =====================================================================================================
    const int row_e = tid & 127, cg = tid >> 7, quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int dq = p.d_pad >> 2;
    TsConsts cst;
    {
      // without a task kernel the prior scale (ScaleKernel) is one constant: folded into the polynomial
      const float g = p.ts_g, ks = p.ts_kscale * ((!TASKS && p.scaled) ? s.tcov[0] : 1.0f), sg = sqrtf(g);
      cst.k0 = pack2(ks, ks);
      cst.c1 = pack2(ks * sg, ks * sg);
      cst.c2 = pack2(ks * g * (1.0f / 3.0f), ks * g * (1.0f / 3.0f));
      const float c3 = (FAMILY == BB_KERNEL_RBF) ? -g : -kLog2e * sg;
      cst.c3 = pack2(c3, c3);
    }
    TsStageRegs regs;

    unsigned rows_have = 0;  // gated pass: rows known to be published (uniform over the conversion threads)
    // the resident layout (fp32 rows, every quad a full aligned 16 bytes) takes one predicated 128-bit load per
    // quad; the layout decision and the address checks of ts_load_quad ran ~130 instructions per thread and tile
    const bool gated = p.ready_rows != nullptr;
    const bool fast_rows = p.layout == BB_ROW_MAJOR_F32 && (p.ldx & 3) == 0 && (p.d & 3) == 0 &&
                           (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
    auto prefetch = [&](int tile) {
      const int64_t row = (int64_t)tile * kTileM + row_e;
      if (gated) rows_have = ts_wait_rows(p, s.ready_cache, tile, rows_have);
      if (fast_rows) {
        const float4* base = reinterpret_cast<const float4*>(p.x) + row * (p.ldx >> 2) + cg;
        const bool q0 = row < p.N && 4 * cg < p.d, q1 = row < p.N && 4 * (cg + 4) < p.d;
        regs.v[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        regs.v[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gated) {
          if (q0) regs.v[0] = __ldcg(base);
          if (q1) regs.v[1] = __ldcg(base + 4);
        } else {
          if (q0) regs.v[0] = __ldg(base);
          if (q1) regs.v[1] = __ldg(base + 4);
        }
        return;
      }
      regs.v[0] = (cg < dq) ? ts_load_quad(p, row, cg) : make_float4(0.f, 0.f, 0.f, 0.f);
      regs.v[1] = (cg + 4 < dq) ? ts_load_quad(p, row, cg + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };

    int trace_it = 0, trace_n = 0;  // test-only event trace: tile counter the next events are filed under
    // scaled candidate rows -> fp16 hi/mid/lo A2 panels; the thread that owns quad 7 appends |a|^2 P and P1
    auto stage_a2 = [&](int slot) {
      float an = 0.f;
      float a7[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jq = cg + 4 * u;
        if (jq < dq) {
          const int j0 = jq * 4;
          const float4 q = regs.v[u];
          if (TASKS && p.task_col >= j0 && p.task_col < j0 + 4) {
            const float tv = (p.task_col == j0) ? q.x : (p.task_col == j0 + 1) ? q.y : (p.task_col == j0 + 2) ? q.z : q.w;
            s.cand_task[slot * kTileM + row_e] = min(max(__float2int_rn(tv), 0), p.n_tasks - 1);
          }
          float a[4];
          a[0] = fmaf(q.x, s.cscale_s[j0], s.cshift_s[j0]);
          a[1] = fmaf(q.y, s.cscale_s[j0 + 1], s.cshift_s[j0 + 1]);
          a[2] = fmaf(q.z, s.cscale_s[j0 + 2], s.cshift_s[j0 + 2]);
          a[3] = fmaf(q.w, s.cscale_s[j0 + 3], s.cshift_s[j0 + 3]);
          an = fmaf(a[0], a[0], fmaf(a[1], a[1], fmaf(a[2], a[2], fmaf(a[3], a[3], an))));
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] *= p.ts_sa;  // exact: power of two
          if (jq == 7) {
#pragma unroll
            for (int e = 0; e < 4; ++e) a7[e] = a[e];
          } else {
            uint2 hi, mid, lo;
            split3_quad(a, hi, mid, lo);
            const uint32_t off = swk_offset<kTsK2>((uint32_t)row_e, (uint32_t)(jq >> 1)) + (uint32_t)(jq & 1) * 8u;
            *reinterpret_cast<uint2*>(s.a2 + off) = hi;
            *reinterpret_cast<uint2*>(s.a2 + kTsA2Split + off) = mid;
            *reinterpret_cast<uint2*>(s.a2 + 2 * kTsA2Split + off) = lo;
          }
        }
      }
      s.an_part[cg * kTileM + row_e] = an;
      if (tid == 0) ts_trace<TRACE>(p, 0, trace_n, trace_it, 121);
      if (cg != 3) {
        bar_quarter_arrive(2 + quarter);
      } else {
        bar_quarter_sync(2 + quarter);  // the four |a|^2 partials of this thread's row are in shared memory
        const float asq = (s.an_part[row_e] + s.an_part[kTileM + row_e]) +
                          (s.an_part[2 * kTileM + row_e] + s.an_part[3 * kTileM + row_e]);
        a7[kTsColSq - 28] = asq * p.ts_aug_sq;
        a7[kTsColOne - 28] = p.ts_aug_one;
        uint2 hi, mid, lo;
        split3_quad(a7, hi, mid, lo);
        const uint32_t off = swk_offset<kTsK2>((uint32_t)row_e, 3u) + 8u;
        *reinterpret_cast<uint2*>(s.a2 + off) = hi;
        *reinterpret_cast<uint2*>(s.a2 + kTsA2Split + off) = mid;
        *reinterpret_cast<uint2*>(s.a2 + 2 * kTsA2Split + off) = lo;
      }
      if (tid == 0) ts_trace<TRACE>(p, 0, trace_n, trace_it, 122);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.a2_full);
    };

    // One K* chunk: D2 (TMEM) -> kernel values -> fp16 hi/lo written over the same TMEM columns; returns the
    // chunk's contribution to sum_i k_i alpha_i in `mean2`.
    auto convert_chunk = [&](int c, int slot, unsigned long long& mean2) {
      float v[16];
      const uint32_t col = kTsD2Col0 + (uint32_t)(c * kChunk + cg * 16);
      tmem_ld16(tmem_base + lane_base + col, v);
      tmem_ld_wait();
      if (tid == 0) ts_trace<TRACE>(p, 0, trace_n, trace_it, 100 + c);
      const int i0 = c * kChunk + cg * 16;
      const float* tcrow = s.tcov + (TASKS ? s.cand_task[slot * kTileM + row_e] : 0) * p.n_tasks;
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        unsigned long long k2 = ts_kernel_pair<FAMILY>(v[2 * e], v[2 * e + 1], cst);
        if constexpr (TASKS) k2 = mul2(k2, pack2(tcrow[s.ttask[i0 + 2 * e]], tcrow[s.ttask[i0 + 2 * e + 1]]));
        const float2 al = *reinterpret_cast<const float2*>(s.alpha_s + i0 + 2 * e);
        mean2 = fma2(k2, pack2(al.x, al.y), mean2);
        const uint32_t h = pack_h2(lo_of(k2), hi_of(k2));
        const unsigned long long r2 = fma2(pack2(h_lo_f32(h), h_hi_f32(h)), pack2(-1.0f, -1.0f), k2);  // exact
        hi[e] = h;
        lo[e] = pack_h2(lo_of(r2), hi_of(r2));
      }
      tmem_st8(tmem_base + lane_base + col, hi);
      tmem_st8(tmem_base + lane_base + col + 8u, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_full[c]);
      if (tid == 0) ts_trace<TRACE>(p, 0, trace_n, trace_it, 110 + c);
    };
    // Schedule (software-pipelined over tiles; D2 lives in two 128-column slabs = chunks {0,1} and {2,3}):
    //   iteration t:  [slab 1 of tile t ready]   stage A2(t+1), convert chunks 2,3 of tile t, publish the mean partials
    //                 [slab 0 of tile t+1 ready] convert chunks 0,1 of tile t+1      <- runs under V(t) on the tensor pipe
    // so the tensor pipe sees V(t,c0) V(t,c1) DIST(t+1,slab0) V(t,c2) V(t,c3) DIST(t+1,slab1) back to back, and the
    // epilogue warps finish tile t (|V|^2, acquisition) while these warps are already converting tile t+1.
    const int C0 = C < 2 ? C : 2;  // chunks in slab 0
    int it = 0;
    int tile = blockIdx.x;
    unsigned long long mean_cur = 0ull;  // mean partial of the tile whose upper chunks are converted next
    if (tile < p.num_tiles) {
      prefetch(tile);
      stage_a2(0);
      if (tile + (int)gridDim.x < p.num_tiles) prefetch(tile + gridDim.x);
      mbar_wait(&s.d2_full[0], 0u);
      tc_fence_after();
      for (int c = 0; c < C0; ++c) convert_chunk(c, 0, mean_cur);
    }
    for (; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1, slot = it & (kTsTaskSlots - 1), slot_next = (it + 1) & (kTsTaskSlots - 1);
      const uint32_t par = (uint32_t)(it & 1);
      const int next = tile + (int)gridDim.x;
      // ---- upper chunks of this tile; the A2 tile is free once this tile's distance GEMMs are complete ----
      trace_it = it;
      if (C > 2) {
        mbar_wait(&s.d2_full[1], par);
        tc_fence_after();
      }
      if (tid == 0) ts_trace<TRACE>(p, 0, trace_n, it, 161);
      if (next < p.num_tiles) {
        stage_a2(slot_next);
        if (next + (int)gridDim.x < p.num_tiles) prefetch(next + gridDim.x);
      }
      if (tid == 0) ts_trace<TRACE>(p, 0, trace_n, it, 120);
      for (int c = 2; c < C; ++c) convert_chunk(c, slot, mean_cur);
      // mean partials of tile `it`: buffer it & 1 was last read by the epilogue of tile it - 2, which finished
      // reading before it released the V accumulator -- and this tile's D2 could not exist before that
      s.mean_part[(buf * 4 + cg) * kTileM + row_e] = lo_of(mean_cur) + hi_of(mean_cur);
      mean_cur = 0ull;
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.mean_full[buf]);
      // ---- lower chunks of the NEXT tile: their distance GEMM was issued behind V(t, c1) ----
      if (next < p.num_tiles) {
        mbar_wait(&s.d2_full[0], par ^ 1u);
        tc_fence_after();
        if (tid == 0) ts_trace<TRACE>(p, 0, trace_n, it, 160);
        trace_it = it + 1;
        for (int c = 0; c < C0; ++c) convert_chunk(c, slot_next, mean_cur);
        trace_it = it;
      }
    }
    if (TRACE && tid == 0 && p.trace != nullptr && blockIdx.x == 0) p.trace[0] = p.trace_cap;
  } else if (warp < kTsWarpProducer) {
    // =====================================================================================================
    // epilogue warps: warp 16 + q owns TMEM lanes 32q..32q+31, one candidate row per thread.
    //   |V|^2 over the row's n_pad accumulator columns (sub-block sb is final once chunk sb's MMAs completed),
    //   moments, acquisition value, running packed-key arg-max.  Everything after the accumulator has been read
    //   runs under the NEXT tile's MMAs and conversions.
    // =====================================================================================================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float inv_v_scale2 = p.inv_r_scale2 / (p.ts_kscale * p.ts_kscale);
    const bool is_mc = p.has_acq && p.acq.kind <= BB_ACQ_QPI;
    const float log_tau = (p.has_acq && p.acq.kind == BB_ACQ_QLOGEI) ? logf(p.acq.tau_relu) : 0.f;
    long long best = kEmptyKey;
    int trace_n = 0;
    const bool tr = (tid == kTsEpiWarp0 * 32);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1, slot = it & (kTsTaskSlots - 1);
      const uint32_t par = (uint32_t)(it & 1);
      unsigned long long ssa = 0ull, ssb = 0ull;
      for (int sb = 0; sb < C; ++sb) {
        mbar_wait_relaxed(&s.vsub_full[sb], par);
        tc_fence_after();
        if (tr) ts_trace<TRACE>(p, 2, trace_n, it, 130 + sb);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + (uint32_t)(sb * kChunk + h * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const unsigned long long va = pack2(v[4 * e], v[4 * e + 1]), vb = pack2(v[4 * e + 2], v[4 * e + 3]);
            ssa = fma2(va, va, ssa);
            ssb = fma2(vb, vb, ssb);
          }
        }
      }
      // inputs that the conversion warps overwrite two tiles later are read BEFORE the accumulator is released:
      // V(t+1) waits for v_empty(t), DIST(t+2) follows V(t+1), and only then can those buffers be written again
      mbar_wait_relaxed(&s.mean_full[buf], (uint32_t)((it >> 1) & 1));
      const float mp = (s.mean_part[(buf * 4 + 0) * kTileM + r] + s.mean_part[(buf * 4 + 1) * kTileM + r]) +
                       (s.mean_part[(buf * 4 + 2) * kTileM + r] + s.mean_part[(buf * 4 + 3) * kTileM + r]);
      const int ctr = TASKS ? s.cand_task[slot * kTileM + r] : 0;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.v_empty);
      if (tr) ts_trace<TRACE>(p, 2, trace_n, it, 140);

      const float vp = (lo_of(ssa) + hi_of(ssa)) + (lo_of(ssb) + hi_of(ssb));
      const float kss = (TASKS || p.scaled) ? s.tcov[ctr * p.n_tasks + ctr] : 1.0f;
      const float var_t = fmaxf(kss - vp * inv_v_scale2, 1e-10f);
      const float mu = fmaf(p.y_std, s.meanc[ctr] + mp, p.y_mean);
      const float var = p.y_std * p.y_std * var_t;
      const int64_t row = (int64_t)tile * kTileM + r;
      const bool live = row < p.N;
      if (live) {
        if (p.mu) p.mu[row] = mu;
        if (p.var) p.var[row] = var;
      }
      if (p.has_acq) {
        float score = 0.f;
        if (is_mc) {
          float c0, c1, s0 = 0.f, s1 = 0.f;
          mc_coef(p.acq, mu, var, c0, c1);
          if (fast_mc) {
            const bool fast = mc_row_fast(s.mc_tab, c0, c1, s0, s1);
            if (tr) ts_trace<TRACE>(p, 2, trace_n, it, 142);
            // rows outside the tabulated envelope: exact sum over all S samples, the whole warp per row
            unsigned need = __ballot_sync(0xffffffffu, !fast);
            while (need != 0u) {
              const int b = __ffs(need) - 1;
              need &= need - 1u;
              const float e0 = __shfl_sync(0xffffffffu, c0, b), e1 = __shfl_sync(0xffffffffu, c1, b);
              float a0, a1;
              mc_row_exact_warp(s.z_s, p.S, e0, e1, lane, a0, a1);
              if (lane == b) {
                s0 = a0;
                s1 = a1;
              }
            }
          } else if (p.acq.kind != BB_ACQ_QUCB && p.acq.kind != BB_ACQ_QSR) {
            // per-sample kinds without a table: the warp takes its 32 rows one after the other, lane l sums the
            // sample groups l, l + 32, ... (fixed order: a row's value does not depend on where it is evaluated)
            for (int b = 0; b < 32; ++b) {
              const float e0 = __shfl_sync(0xffffffffu, c0, b), e1 = __shfl_sync(0xffffffffu, c1, b);
              float a0, a1;
              mc_row_groups_warp(p.acq.kind, s.z_s, p.S, e0, e1, lane, a0, a1);
              if (lane == b) {
                s0 = a0;
                s1 = a1;
              }
            }
          }
          if (tr) ts_trace<TRACE>(p, 2, trace_n, it, 143);
          score = (p.acq.kind == BB_ACQ_QLOGEI) ? log_tau + logf((s0 + 0.1f * s1) / (float)p.S)
                                                 : mc_finalize(p.acq, mu, var, s0, s1, p.S, s.zstat[0], s.zstat[1]);
        } else {
          score = analytic_value(p.acq, mu, var);
        }
        if (live) {
          if (p.score) p.score[row] = score;
          const bool ok = (p.keep == nullptr || p.keep[row] != 0) && !(score != score);
          if (ok) {
            const long long key = pack_key(score, (uint32_t)(row + p.index_offset));
            best = key > best ? key : best;
          }
        }
      }
      if (tr) ts_trace<TRACE>(p, 2, trace_n, it, 150);
    }
    if (p.best_key != nullptr && p.has_acq) {
      for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
      if (lane == 0) s.best_red[quarter] = best;
      asm volatile("bar.sync 6, 128;" ::: "memory");  // the four epilogue warps
      if (quarter == 0 && lane == 0) {
        long long b = s.best_red[0];
        for (int w = 1; w < kTsEpiWarps; ++w) b = s.best_red[w] > b ? s.best_red[w] : b;
        if (b != kEmptyKey) atomicMax(p.best_key, b);
      }
    }
  } else if (warp == kTsWarpProducer) {
    // =====================================================================================================
    // producer (TMA engine): resident images once, then the lo image of L^-1 piece by piece for every tile
    // =====================================================================================================
    if (elect_one()) {
      uint32_t left = hi_bytes;
      mbar_expect_tx(s.res_full, hi_bytes + 3u * bt_split);
      for (uint32_t o = 0; left > 0;) {
        const uint32_t n = left < 32768u ? left : 32768u;
        bulk_g2s(s.lh + o, p.timg_l + o, n, s.res_full);
        o += n;
        left -= n;
      }
      left = 3u * bt_split;
      for (uint32_t o = 0; left > 0;) {
        const uint32_t n = left < 32768u ? left : 32768u;
        bulk_g2s(s.bt + o, p.timg_b + o, n, s.res_full);
        o += n;
        left -= n;
      }
      uint32_t rs = 0, rph = 0;
      const uint8_t* lo_img = p.timg_l + hi_bytes;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        uint32_t off = 0;
        for (int c = 0; c < C; ++c) {
          const int rows_c = p.n_pad - c * kChunk;
          for (int r0 = 0; r0 < rows_c; r0 += 128) {
            const uint32_t bytes = (uint32_t)(rows_c - r0 < 128 ? rows_c - r0 : 128) * 128u;
            mbar_wait_relaxed(&s.r_empty[rs], rph ^ 1u);
            mbar_expect_tx(&s.r_full[rs], bytes);
            bulk_g2s(s.ring + (size_t)rs * kTsLoStage, lo_img + off + (uint32_t)r0 * 128u, bytes, &s.r_full[rs]);
            if (++rs == (uint32_t)kTsLoStages) {
              rs = 0;
              rph ^= 1u;
            }
          }
          off += (uint32_t)rows_c * 128u;
        }
      }
    }
    __syncwarp();
  } else {
    // MMA issuer (ts_mma_role): n_pad = 256, the headline shape, runs the fully constant-folded instance
    if (p.n_pad == 4 * kChunk) ts_mma_role<4, TRACE>(p, s, tmem_base, lane);
    else ts_mma_role<0, TRACE>(p, s, tmem_base, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTsWarpProducer) tmem_dealloc(tmem_base, 512);
}

// ============================================================================================================
// k_kmat_ts -- stand-alone K(X*, X) (bb_kernel_matrix), the HBM-side kernel of the north star.
// Same front half as k_fused_ts (augmented tcgen05 distance GEMM, packed-f32x2 kernel epilogue); the kernel values
// go to a 128B-swizzled staging tile in shared memory and leave the SM through the TMA engine: one
// cp.async.bulk.tensor.2d store (UTMASTG) per 128 x 32 box, addressed through a tensor map that also clips the
// ragged edges (rows >= N, columns >= n).  D2 is double-buffered in tensor memory (2 x 256 columns) so the distance
// GEMM of tile t+1 runs under the epilogue of tile t.  Algorithmic bytes: 4d read + 4n written per candidate.
// ============================================================================================================
constexpr uint32_t kKmBox = 128u * 128u;   // one TMA box: 128 rows x 32 fp32 = 16 KB
constexpr uint32_t kKmStage = 2u * kKmBox; // one 64-column chunk of a tile

struct KmSmem {
  uint8_t *bt, *a2, *out;
  float *tcov, *cscale_s, *cshift_s, *an_part;
  int32_t *ttask, *cand_task;
  uint64_t *d2_full, *d2_empty, *a2_full, *res_full;
  uint32_t* tmem_ptr;
};

__host__ __device__ inline size_t km_carve(uint8_t* base, int n_pad, KmSmem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) {
    off = (off + align - 1) / align * align;
    size_t o = off;
    off += bytes;
    return o;
  };
  const size_t o_bt = take((size_t)3 * n_pad * kTsK2 * 2, 1024);
  const size_t o_a2 = take((size_t)3 * kTsA2Split, 1024);
  const size_t o_out = take((size_t)2 * kKmStage, 1024);
  const size_t o_tt = take((size_t)n_pad * 4, 16);
  const size_t o_tc = take((size_t)kMaxTasks * kMaxTasks * 4, 16);
  const size_t o_cs = take(32 * 4, 16), o_sh = take(32 * 4, 16);
  const size_t o_an = take(4 * kTileM * 4, 16), o_ct = take(2 * kTileM * 4, 16);
  const size_t o_bar = take(16 * 8, 16), o_misc = take(32, 16);
  if (s) {
    s->bt = base + o_bt;
    s->a2 = base + o_a2;
    s->out = base + o_out;
    s->ttask = reinterpret_cast<int32_t*>(base + o_tt);
    s->tcov = reinterpret_cast<float*>(base + o_tc);
    s->cscale_s = reinterpret_cast<float*>(base + o_cs);
    s->cshift_s = reinterpret_cast<float*>(base + o_sh);
    s->an_part = reinterpret_cast<float*>(base + o_an);
    s->cand_task = reinterpret_cast<int32_t*>(base + o_ct);
    uint64_t* b = reinterpret_cast<uint64_t*>(base + o_bar);
    s->d2_full = b;        // [2]
    s->d2_empty = b + 2;   // [2]
    s->a2_full = b + 4;
    s->res_full = b + 5;
    s->tmem_ptr = reinterpret_cast<uint32_t*>(base + o_misc);
  }
  return off;
}

__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int FAMILY, bool TASKS>
__global__ void __launch_bounds__(kFusedThreads, 1) k_kmat_ts(const FusedParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  KmSmem s;
  km_carve(smem_raw, p.n_pad, &s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.n_chunks;
  const uint32_t bt_split = (uint32_t)p.n_pad * kTsK2 * 2u;
  if (tid == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.d2_full[i], 1);
      mbar_init(&s.d2_empty[i], kComputeWarps);
    }
    mbar_init(s.a2_full, kComputeWarps);
    mbar_init(s.res_full, 1);
    fence_mbar_init();
  }
  if (warp == kWarpProducer) {
    tmem_alloc(s.tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int e = tid; e < (int)(3 * kTsA2Split / 16); e += kFusedThreads)
    reinterpret_cast<uint4*>(s.a2)[e] = make_uint4(0u, 0u, 0u, 0u);
  for (int e = tid; e < 32; e += kFusedThreads) {
    s.cscale_s[e] = e < p.d_pad ? __ldg(p.cand_scale + e) : 0.f;
    s.cshift_s[e] = e < p.d_pad ? __ldg(p.cand_shift + e) : 0.f;
  }
  for (int e = tid; e < p.n_pad; e += kFusedThreads) s.ttask[e] = TASKS ? __ldg(p.train_task + e) : 0;
  for (int e = tid; e < p.n_tasks * p.n_tasks; e += kFusedThreads) s.tcov[e] = __ldg(p.task_covar + e);
  for (int e = tid; e < 2 * kTileM; e += kFusedThreads) s.cand_task[e] = 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;

  if (warp < kComputeWarps) {
    const int row_e = tid & 127, cg = tid >> 7, quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int dq = p.d_pad >> 2;
    TsConsts cst;
    {
      const float g = p.ts_g, ks = (!TASKS && p.scaled) ? s.tcov[0] : 1.0f, sg = sqrtf(g);
      cst.k0 = pack2(ks, ks);
      cst.c1 = pack2(ks * sg, ks * sg);
      cst.c2 = pack2(ks * g * (1.0f / 3.0f), ks * g * (1.0f / 3.0f));
      const float c3 = (FAMILY == BB_KERNEL_RBF) ? -g : -kLog2e * sg;
      cst.c3 = pack2(c3, c3);
    }
    TsStageRegs regs;
    auto prefetch = [&](int tile) {
      const int64_t row = (int64_t)tile * kTileM + row_e;
      regs.v[0] = (cg < dq) ? ts_load_quad(p, row, cg) : make_float4(0.f, 0.f, 0.f, 0.f);
      regs.v[1] = (cg + 4 < dq) ? ts_load_quad(p, row, cg + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto stage_a2 = [&](int buf) {  // as in k_fused_ts
      float an = 0.f;
      float a7[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jq = cg + 4 * u;
        if (jq < dq) {
          const int j0 = jq * 4;
          const float4 q = regs.v[u];
          if (TASKS && p.task_col >= j0 && p.task_col < j0 + 4) {
            const float tv = (p.task_col == j0) ? q.x : (p.task_col == j0 + 1) ? q.y : (p.task_col == j0 + 2) ? q.z : q.w;
            s.cand_task[buf * kTileM + row_e] = min(max(__float2int_rn(tv), 0), p.n_tasks - 1);
          }
          float a[4];
          a[0] = fmaf(q.x, s.cscale_s[j0], s.cshift_s[j0]);
          a[1] = fmaf(q.y, s.cscale_s[j0 + 1], s.cshift_s[j0 + 1]);
          a[2] = fmaf(q.z, s.cscale_s[j0 + 2], s.cshift_s[j0 + 2]);
          a[3] = fmaf(q.w, s.cscale_s[j0 + 3], s.cshift_s[j0 + 3]);
          an = fmaf(a[0], a[0], fmaf(a[1], a[1], fmaf(a[2], a[2], fmaf(a[3], a[3], an))));
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] *= p.ts_sa;
          if (jq == 7) {
#pragma unroll
            for (int e = 0; e < 4; ++e) a7[e] = a[e];
          } else {
            uint2 hi, mid, lo;
            split3_quad(a, hi, mid, lo);
            const uint32_t off = swk_offset<kTsK2>((uint32_t)row_e, (uint32_t)(jq >> 1)) + (uint32_t)(jq & 1) * 8u;
            *reinterpret_cast<uint2*>(s.a2 + off) = hi;
            *reinterpret_cast<uint2*>(s.a2 + kTsA2Split + off) = mid;
            *reinterpret_cast<uint2*>(s.a2 + 2 * kTsA2Split + off) = lo;
          }
        }
      }
      s.an_part[cg * kTileM + row_e] = an;
      if (cg != 3) {
        bar_quarter_arrive(2 + quarter);
      } else {
        bar_quarter_sync(2 + quarter);
        const float asq = (s.an_part[row_e] + s.an_part[kTileM + row_e]) +
                          (s.an_part[2 * kTileM + row_e] + s.an_part[3 * kTileM + row_e]);
        a7[kTsColSq - 28] = asq * p.ts_aug_sq;
        a7[kTsColOne - 28] = p.ts_aug_one;
        uint2 hi, mid, lo;
        split3_quad(a7, hi, mid, lo);
        const uint32_t off = swk_offset<kTsK2>((uint32_t)row_e, 3u) + 8u;
        *reinterpret_cast<uint2*>(s.a2 + off) = hi;
        *reinterpret_cast<uint2*>(s.a2 + kTsA2Split + off) = mid;
        *reinterpret_cast<uint2*>(s.a2 + 2 * kTsA2Split + off) = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.a2_full);
    };

    int it = 0, chunk_no = 0;
    int tile = blockIdx.x;
    if (tile < p.num_tiles) {
      prefetch(tile);
      stage_a2(0);
      if (tile + (int)gridDim.x < p.num_tiles) prefetch(tile + gridDim.x);
    }
    for (; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int slot = it & 1;
      mbar_wait(&s.d2_full[slot], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      // this tile's distance GEMM has consumed A2: stage the next tile so that ITS GEMM runs under this epilogue
      const int next = tile + (int)gridDim.x;
      if (next < p.num_tiles) {
        stage_a2((it + 1) & 1);
        if (next + (int)gridDim.x < p.num_tiles) prefetch(next + gridDim.x);
      }
      const int ct = TASKS ? s.cand_task[(it & 1) * kTileM + row_e] : 0;
      const float* tcrow = s.tcov + ct * p.n_tasks;
      for (int c = 0; c < C; ++c, ++chunk_no) {
        float v[16];
        tmem_ld16(tmem_base + lane_base + (uint32_t)(slot * 256 + c * kChunk + cg * 16), v);
        tmem_ld_wait();
        if (c == C - 1) {  // all of D2 is in registers: the distance GEMM of tile t+2 may overwrite this slot
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.d2_empty[slot]);
        }
        const int i0 = c * kChunk + cg * 16;
        uint8_t* box = s.out + (size_t)(chunk_no & 1) * kKmStage + (size_t)(cg >> 1) * kKmBox + (size_t)row_e * 128u;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          unsigned long long k01 = ts_kernel_pair<FAMILY>(v[4 * q4], v[4 * q4 + 1], cst);
          unsigned long long k23 = ts_kernel_pair<FAMILY>(v[4 * q4 + 2], v[4 * q4 + 3], cst);
          if constexpr (TASKS) {
            k01 = mul2(k01, pack2(tcrow[s.ttask[i0 + 4 * q4]], tcrow[s.ttask[i0 + 4 * q4 + 1]]));
            k23 = mul2(k23, pack2(tcrow[s.ttask[i0 + 4 * q4 + 2]], tcrow[s.ttask[i0 + 4 * q4 + 3]]));
          }
          const uint32_t chunk16 = (uint32_t)(((cg & 1) * 4 + q4) ^ (row_e & 7));  // SWIZZLE_128B
          *reinterpret_cast<float4*>(box + chunk16 * 16u) = make_float4(lo_of(k01), hi_of(k01), lo_of(k23), hi_of(k23));
        }
        fence_proxy_async();            // generic-proxy writes -> visible to the TMA engine
        if (tid == 0) tma_wait_read0(); // the store that last read the OTHER buffer (one chunk ago) has drained it
        bar_compute();
        if (tid == 0) {
          const uint8_t* src = s.out + (size_t)(chunk_no & 1) * kKmStage;
          const int y = tile * kTileM;  // row coordinate; int32 limits N to 2^31 rows
          tma_store_2d(&tmap, src, c * kChunk, y);
          tma_store_2d(&tmap, src + kKmBox, c * kChunk + 32, y);
          tma_commit_group();
        }
      }
    }
    if (tid == 0) tma_wait_all();  // global writes complete before the CTA exits
  } else if (warp == kWarpProducer) {
    if (elect_one()) {
      uint32_t left = 3u * bt_split;
      mbar_expect_tx(s.res_full, left);
      for (uint32_t o = 0; left > 0;) {
        const uint32_t n = left < 32768u ? left : 32768u;
        bulk_g2s(s.bt + o, p.timg_b + o, n, s.res_full);
        o += n;
        left -= n;
      }
    }
    __syncwarp();
  } else {
    const uint32_t idesc_d2 = make_idesc_f16(kTileM, p.n_pad);
    const uint32_t a2_addr = smem_u32(s.a2), bt_addr = smem_u32(s.bt);
    const uint64_t a2_h = make_swk_desc<kTsK2>(a2_addr), a2_m = make_swk_desc<kTsK2>(a2_addr + kTsA2Split),
                   a2_l = make_swk_desc<kTsK2>(a2_addr + 2 * kTsA2Split);
    const uint64_t b_h = make_swk_desc<kTsK2>(bt_addr), b_m = make_swk_desc<kTsK2>(bt_addr + bt_split),
                   b_l = make_swk_desc<kTsK2>(bt_addr + 2 * bt_split);
    mbar_wait_relaxed(s.res_full, 0u);
    int j = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
      const int slot = j & 1;
      mbar_wait_relaxed(s.a2_full, (uint32_t)(j & 1));
      if (j >= 2) mbar_wait_relaxed(&s.d2_empty[slot], (uint32_t)(((j >> 1) - 1) & 1));
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(slot * 256);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kTsK2 / 16; ++kk) {
          const uint64_t ko = (uint64_t)(kk * 2);
          umma_f16(d_addr, a2_h + ko, b_h + ko, idesc_d2, kk > 0 ? 1u : 0u);
          umma_f16(d_addr, a2_h + ko, b_m + ko, idesc_d2, 1u);
          umma_f16(d_addr, a2_m + ko, b_h + ko, idesc_d2, 1u);
          umma_f16(d_addr, a2_h + ko, b_l + ko, idesc_d2, 1u);
          umma_f16(d_addr, a2_l + ko, b_h + ko, idesc_d2, 1u);
          umma_f16(d_addr, a2_m + ko, b_m + ko, idesc_d2, 1u);
        }
        umma_commit(&s.d2_full[slot]);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpProducer) tmem_dealloc(tmem_base, 512);
}

bool kmat_ts_supported(const FusedParams& p, const float* d_k, int64_t ldk, int max_smem) {
  if (p.timg_b == nullptr || p.n_pad > 256 || p.d > kTsColSq || p.family == BB_KERNEL_MATERN12) return false;
  if (p.n_tasks > kMaxTasks || p.N >= (1ll << 31)) return false;
  if ((reinterpret_cast<uintptr_t>(d_k) & 15u) != 0 || (ldk & 3) != 0) return false;  // TMA: 16-byte aligned base / row pitch
  return km_carve(nullptr, p.n_pad, nullptr) + 1024 <= (size_t)max_smem;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int FAMILY, bool TASKS>
static int launch_km_one(FusedParams& p, const CUtensorMap& tm, int grid, size_t smem, cudaStream_t stream) {
  static int configured_for = -1;
  int dev = 0;
  BB_CUDA(cudaGetDevice(&dev));
  if (configured_for != dev) {
    BB_CUDA(cudaFuncSetAttribute(k_kmat_ts<FAMILY, TASKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured_for = dev;
  }
  k_kmat_ts<FAMILY, TASKS><<<grid, kFusedThreads, smem, stream>>>(p, tm);
  BB_LAUNCH_CHECK();
  return BB_OK;
}

// K(X*, X) -> d_k [N, ldk] (columns >= n untouched); returns BB_ERR_UNSUPPORTED if the driver lacks tensor maps.
int launch_kmat_ts(FusedParams& p, float* d_k, int64_t ldk, int n_cols, int grid, cudaStream_t stream) {
  static EncodeTiledFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    BB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    BB_CHECK_SUPPORTED(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)n_cols, (cuuint64_t)p.N};
  const cuuint64_t gstride[1] = {(cuuint64_t)ldk * 4};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_k, gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return BB_ERR_CUDA;
  }
  const size_t smem = km_carve(nullptr, p.n_pad, nullptr) + 1024;
  const bool tasks = p.task_col >= 0;
  switch (p.family) {
    case BB_KERNEL_MATERN32:
      return tasks ? launch_km_one<BB_KERNEL_MATERN32, true>(p, tm, grid, smem, stream)
                   : launch_km_one<BB_KERNEL_MATERN32, false>(p, tm, grid, smem, stream);
    case BB_KERNEL_MATERN52:
      return tasks ? launch_km_one<BB_KERNEL_MATERN52, true>(p, tm, grid, smem, stream)
                   : launch_km_one<BB_KERNEL_MATERN52, false>(p, tm, grid, smem, stream);
    default:
      return tasks ? launch_km_one<BB_KERNEL_RBF, true>(p, tm, grid, smem, stream)
                   : launch_km_one<BB_KERNEL_RBF, false>(p, tm, grid, smem, stream);
  }
}

// bb_kernel_matrix front door: *handled = false when the shape / alignment is outside this kernel's envelope.
int try_kmat_ts(const bb_model* m, const void* d_x, int32_t layout, int64_t N, int64_t ldx, float* d_k, int64_t ldk,
                cudaStream_t stream, bool* handled) {
  *handled = false;
  if (m->d_timg_b == nullptr || layout == BB_BITS_U8) return BB_OK;
  static const bool off = [] {
    const char* e = getenv("BB_FORCE_KERNEL");
    return e != nullptr && e[0] == 't' && e[1] == 'c';
  }();
  if (off) return BB_OK;
  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.x = d_x;
  p.layout = layout;
  p.N = N;
  p.ldx = ldx;
  p.num_tiles = (int)((N + kTileM - 1) / kTileM);
  p.cand_scale = m->d_cand_scale;
  p.cand_shift = m->d_cand_shift;
  p.task_covar = m->d_task_covar;
  p.train_task = m->d_train_task;
  p.family = m->family;
  p.n_pad = m->n_pad;
  p.d = m->d;
  p.d_pad = m->d_pad;
  p.n_chunks = m->n_chunks;
  p.task_col = m->task_col;
  p.n_tasks = m->n_tasks;
  p.scaled = (m->task_col >= 0 || m->prior_scale != 1.0f) ? 1 : 0;
  p.timg_b = reinterpret_cast<const uint8_t*>(m->d_timg_b);
  p.ts_sa = m->ts_sa;
  p.ts_aug_sq = m->ts_aug_sq;
  p.ts_aug_one = m->ts_aug_one;
  p.ts_g = m->ts_g;
  p.ts_kscale = 1.0f;
  static int sms = 0, max_smem = 0, cached_dev = -1;  // device attributes once per device, not per call
  int dev = 0;
  BB_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    BB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    BB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cached_dev = dev;
  }
  if (!kmat_ts_supported(p, d_k, ldk, max_smem)) return BB_OK;
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  const int rc = launch_kmat_ts(p, d_k, ldk, m->n, grid, stream);
  if (rc == BB_OK) *handled = true;
  return rc == BB_ERR_UNSUPPORTED ? BB_OK : rc;
}

// Shape envelope of this kernel; everything else runs fused_tc.cu / fused.cu.
bool fused_ts_supported(const FusedParams& p, int max_smem) {
  if (p.timg_l == nullptr || p.timg_b == nullptr || p.ts_alpha == nullptr) return false;
  if (p.layout > BB_COL_MAJOR_F64 && p.layout != kLayoutCodes4 && p.layout != kLayoutCodes8) return false;
  if (p.n_pad > 256 || p.d > kTsColSq || p.family == BB_KERNEL_MATERN12) return false;
  if (p.has_acq && (p.S > 512 || (p.S & 15) != 0)) return false;
  if (p.n_tasks > kMaxTasks) return false;
  return ts_carve(nullptr, p.n_pad, p.n_tasks, nullptr) + 1024 <= (size_t)max_smem;
}

template <int FAMILY, bool TASKS, bool TRACE>
static int launch_ts_one2(FusedParams& p, int grid, size_t smem, cudaStream_t stream) {
  static int configured_for = -1;  // cudaFuncSetAttribute once per device, not per launch
  int dev = 0;
  BB_CUDA(cudaGetDevice(&dev));
  if (configured_for != dev) {
    BB_CUDA(cudaFuncSetAttribute(k_fused_ts<FAMILY, TASKS, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured_for = dev;
  }
  if (p.mc_table != nullptr) {
    // the table kernel launched just before on this stream releases its dependents at once
    // (griddepcontrol.launch_dependents): this kernel's prologue overlaps it and waits (griddepcontrol.wait) only
    // before it reads the table
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)kTsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BB_CUDA(cudaLaunchKernelEx(&cfg, k_fused_ts<FAMILY, TASKS, TRACE>, p));
    return BB_OK;
  }
  k_fused_ts<FAMILY, TASKS, TRACE><<<grid, kTsThreads, smem, stream>>>(p);
  BB_LAUNCH_CHECK();
  return BB_OK;
}
template <int FAMILY, bool TASKS>
static int launch_ts_one(FusedParams& p, int grid, size_t smem, cudaStream_t stream) {
  // the instrumented variant only exists for the headline family without tasks (test-only pipeline timeline)
  if constexpr (FAMILY == BB_KERNEL_MATERN52 && !TASKS)
    if (p.trace != nullptr) return launch_ts_one2<FAMILY, TASKS, true>(p, grid, smem, stream);
  return launch_ts_one2<FAMILY, TASKS, false>(p, grid, smem, stream);
}

int launch_fused_ts(FusedParams& p, int grid, cudaStream_t stream) {
  const size_t smem = ts_carve(nullptr, p.n_pad, p.n_tasks, nullptr) + 1024;
  const bool tasks = p.task_col >= 0;
  switch (p.family) {
    case BB_KERNEL_MATERN32:
      return tasks ? launch_ts_one<BB_KERNEL_MATERN32, true>(p, grid, smem, stream)
                   : launch_ts_one<BB_KERNEL_MATERN32, false>(p, grid, smem, stream);
    case BB_KERNEL_MATERN52:
      return tasks ? launch_ts_one<BB_KERNEL_MATERN52, true>(p, grid, smem, stream)
                   : launch_ts_one<BB_KERNEL_MATERN52, false>(p, grid, smem, stream);
    default:
      return tasks ? launch_ts_one<BB_KERNEL_RBF, true>(p, grid, smem, stream)
                   : launch_ts_one<BB_KERNEL_RBF, false>(p, grid, smem, stream);
  }
}

}  // namespace bb
