// fused_tc.cu -- fused scoring kernel with BOTH contractions on the tensor cores.
//
// Per tile of 128 candidates:
//   MMA-2   D2 = A2 * Bt^T          tcgen05, A2 = scaled candidate rows, Bt = -2 x scaled training rows,
//                                   fp16 hi/mid/lo split of both (6 products, error 2^-33), fp32 in TMEM
//   assemble t = D2 + |a|^2 + |b|^2 -> Matern/RBF -> K* chunk (fp32), mean partial, fp16 hi/lo split
//                                   -> A operand ring in shared memory                (CUDA cores)
//   MMA-1   V = K* L^-T             tcgen05, fp16 hi/lo x3, triangular skip, fp32 in TMEM (as fused.cu)
//   epilogue |V|^2 per 64-column sub-block as soon as that sub-block is final (no MMA tail to wait
//           for), un-standardise, q=1 acquisition, arg-max.
// The GEMM-form distance is the formulation gpytorch uses (Distance._sq_dist), here with the inner
// product on tcgen05.  TMEM: columns 0..255 = V accumulator, 256..511 = four D2 slots of 64 columns.
// Requires n_pad <= 256, d_pad <= 64, S <= 512; other shapes run fused.cu.
//
// Reference path replaced: see fused.cu.
#include "fused_common.cuh"

namespace bb {

constexpr int kTcSlotsA = 2;
constexpr int kTcMaxStagesR = 4;  // runtime p.stages_b in [2, 4]
constexpr uint32_t kRStageBytes = 32768;  // one pair block of the L^-1 image: [hi 16 KB | lo 16 KB]
// [hi | mid | lo] split panels; K2 = p.dist_k fp16 per row
__host__ __device__ constexpr uint32_t bt_split_bytes(int k2, int n_pad) { return (uint32_t)n_pad * (uint32_t)k2 * 2u; }
__host__ __device__ constexpr uint32_t a2_split_bytes(int k2) { return 128u * (uint32_t)k2 * 2u; }  // 128 candidates
constexpr int kTcMaxSamples = 512;
constexpr uint32_t kD2Col0 = 256;

struct TcSmem {
  uint8_t *ring_a, *ring_r, *bt, *a2;
  float *tsq, *alpha_s;
  int32_t* ttask;
  float *z_s, *an_part, *mean_part, *var_part, *mc_part, *tcov, *meanc, *cscale_s, *cshift_s;
  int32_t* cand_task;
  uint64_t *a_full, *a_empty, *r_full, *r_empty, *d2_full, *d2_empty, *dsub_full, *d_empty, *a2_full,
      *a2_empty;
  long long* best_red;
  uint32_t* tmem_ptr;
  float* zstat;
};

__host__ __device__ inline size_t tc_carve(uint8_t* base, const FusedParams& p, TcSmem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 15) / 16 * 16;
    return o;
  };
  const size_t o_ra = take((size_t)kTcSlotsA * kSlotABytes);
  const size_t o_rr = take((size_t)p.stages_b * kRStageBytes);
  const size_t o_bt = take((size_t)3 * bt_split_bytes(p.dist_k, p.n_pad));  // resident Bt panels
  const size_t o_a2 = take((size_t)3 * a2_split_bytes(p.dist_k));
  const size_t o_tsq = take(p.n_pad * 4), o_al = take(p.n_pad * 4), o_tt = take(p.scaled ? p.n_pad * 4 : 16);
  const size_t o_z = take(kTcMaxSamples * 4);
  const size_t o_an = take(2 * 4 * kTileM * 4), o_mp = take(4 * kTileM * 4), o_vp = take(4 * kTileM * 4);
  const size_t o_mc = take(4 * kTileM * 2 * 4);
  const size_t o_tc = take(kMaxTasks * kMaxTasks * 4), o_mcn = take(kMaxTasks * 4);
  const size_t o_cs = take(64 * 4), o_sh = take(64 * 4);
  const size_t o_ct = take(2 * kTileM * 4);
  const size_t o_bar = take(48 * 8);
  const size_t o_best = take(32), o_misc = take(32);
  if (s) {
    s->ring_a = base + o_ra;
    s->ring_r = base + o_rr;
    s->bt = base + o_bt;
    s->a2 = base + o_a2;
    s->tsq = reinterpret_cast<float*>(base + o_tsq);
    s->alpha_s = reinterpret_cast<float*>(base + o_al);
    s->ttask = reinterpret_cast<int32_t*>(base + o_tt);
    s->z_s = reinterpret_cast<float*>(base + o_z);
    s->an_part = reinterpret_cast<float*>(base + o_an);
    s->mean_part = reinterpret_cast<float*>(base + o_mp);
    s->var_part = reinterpret_cast<float*>(base + o_vp);
    s->mc_part = reinterpret_cast<float*>(base + o_mc);
    s->tcov = reinterpret_cast<float*>(base + o_tc);
    s->meanc = reinterpret_cast<float*>(base + o_mcn);
    s->cscale_s = reinterpret_cast<float*>(base + o_cs);
    s->cshift_s = reinterpret_cast<float*>(base + o_sh);
    s->cand_task = reinterpret_cast<int32_t*>(base + o_ct);
    uint64_t* b = reinterpret_cast<uint64_t*>(base + o_bar);
    s->a_full = b;            // [<=4]
    s->a_empty = b + 4;       // [<=4]
    s->r_full = b + 8;        // [<=6]
    s->r_empty = b + 14;      // [<=6]
    s->d2_full = b + 24;      // [1]
    s->d2_empty = b + 28;     // [1]
    s->dsub_full = b + 32;    // [4]
    s->d_empty = b + 36;      // [1]
    s->a2_full = b + 37;      // [1]
    s->a2_empty = b + 38;     // [1]
    s->best_red = reinterpret_cast<long long*>(base + o_best);
    s->tmem_ptr = reinterpret_cast<uint32_t*>(base + o_misc);
    s->zstat = reinterpret_cast<float*>(base + o_misc + 8);
  }
  return off;
}

constexpr int kTcStageQuads = 4;  // d_pad <= 64 -> at most 16 quads over 4 thread groups

struct TcStageRegs {
  float4 v[kTcStageQuads];
};

__device__ __forceinline__ float4 tc_load_quad(const FusedParams& p, int64_t row, int jq) {
  const int j0 = jq * 4;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= p.N) return q;
  switch (p.layout) {
    case BB_ROW_MAJOR_F32: {
      const float* ptr = reinterpret_cast<const float*>(p.x) + row * p.ldx + j0;
      if (j0 + 3 < p.d && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)) return __ldg(reinterpret_cast<const float4*>(ptr));
      if (j0 < p.d) q.x = __ldg(ptr);
      if (j0 + 1 < p.d) q.y = __ldg(ptr + 1);
      if (j0 + 2 < p.d) q.z = __ldg(ptr + 2);
      if (j0 + 3 < p.d) q.w = __ldg(ptr + 3);
      return q;
    }
    case BB_COL_MAJOR_F32:
      if (j0 < p.d) q.x = load_x<BB_COL_MAJOR_F32>(p.x, row, j0, p.ldx);
      if (j0 + 1 < p.d) q.y = load_x<BB_COL_MAJOR_F32>(p.x, row, j0 + 1, p.ldx);
      if (j0 + 2 < p.d) q.z = load_x<BB_COL_MAJOR_F32>(p.x, row, j0 + 2, p.ldx);
      if (j0 + 3 < p.d) q.w = load_x<BB_COL_MAJOR_F32>(p.x, row, j0 + 3, p.ldx);
      return q;
    case BB_ROW_MAJOR_F64:
      if (j0 < p.d) q.x = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0, p.ldx);
      if (j0 + 1 < p.d) q.y = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0 + 1, p.ldx);
      if (j0 + 2 < p.d) q.z = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0 + 2, p.ldx);
      if (j0 + 3 < p.d) q.w = load_x<BB_ROW_MAJOR_F64>(p.x, row, j0 + 3, p.ldx);
      return q;
    default:
      if (j0 < p.d) q.x = load_x<BB_COL_MAJOR_F64>(p.x, row, j0, p.ldx);
      if (j0 + 1 < p.d) q.y = load_x<BB_COL_MAJOR_F64>(p.x, row, j0 + 1, p.ldx);
      if (j0 + 2 < p.d) q.z = load_x<BB_COL_MAJOR_F64>(p.x, row, j0 + 2, p.ldx);
      if (j0 + 3 < p.d) q.w = load_x<BB_COL_MAJOR_F64>(p.x, row, j0 + 3, p.ldx);
      return q;
  }
}

template <int FAMILY, int K2>
__global__ void __launch_bounds__(kFusedThreads, 1) k_fused_tc(const FusedParams p) {
  constexpr uint32_t kA2Split = a2_split_bytes(K2);
  constexpr uint32_t kA2Bytes = 3 * kA2Split;
  const uint32_t kBtSplit = bt_split_bytes(K2, p.n_pad);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcSmem s;
  tc_carve(smem_raw, p, &s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.n_chunks;
  const int dq = p.d_pad >> 2;
  const int ksteps = (p.d_pad + 15) >> 4;  // 16 fp16 per tcgen05 K step
  if (tid == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();

  // ---- one-time setup ----
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < kTcSlotsA; ++i) {
      mbar_init(&s.a_full[i], kComputeWarps);
      mbar_init(&s.a_empty[i], 1);
    }
    for (int i = 0; i < p.stages_b; ++i) {
      mbar_init(&s.r_full[i], 1);
      mbar_init(&s.r_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&s.dsub_full[i], 1);
    mbar_init(s.d2_full, 1);
    mbar_init(s.d2_empty, kComputeWarps);
    mbar_init(s.d_empty, kComputeWarps);
    mbar_init(s.a2_full, kComputeWarps);
    mbar_init(s.a2_empty, 1);
    fence_mbar_init();
  }
  if (warp == kWarpProducer) {
    tmem_alloc(s.tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int e = tid; e < (int)(kA2Bytes / 16); e += kFusedThreads)  // dims >= d stay zero for good
    reinterpret_cast<uint4*>(s.a2)[e] = make_uint4(0u, 0u, 0u, 0u);
  for (int e = tid; e < (int)(3 * kBtSplit / 16); e += kFusedThreads)  // resident Bt panels
    reinterpret_cast<uint4*>(s.bt)[e] = __ldg(reinterpret_cast<const uint4*>(p.bimg) + e);
  for (int e = tid; e < 64; e += kFusedThreads) {
    s.cscale_s[e] = e < p.d_pad ? __ldg(p.cand_scale + e) : 0.f;
    s.cshift_s[e] = e < p.d_pad ? __ldg(p.cand_shift + e) : 0.f;
  }
  for (int e = tid; e < p.n_pad; e += kFusedThreads) {
    s.tsq[e] = __ldg(p.train_sq + e);
    s.alpha_s[e] = __ldg(p.alpha + e);
    if (p.scaled) s.ttask[e] = __ldg(p.train_task + e);
  }
  for (int e = tid; e < p.n_tasks * p.n_tasks; e += kFusedThreads) s.tcov[e] = __ldg(p.task_covar + e);
  for (int e = tid; e < p.n_tasks; e += kFusedThreads) s.meanc[e] = __ldg(p.mean_const + e);
  for (int e = tid; e < 2 * kTileM; e += kFusedThreads) s.cand_task[e] = 0;
  if (p.has_acq && p.z != nullptr)
    for (int e = tid; e < p.S; e += kFusedThreads) s.z_s[e] = __ldg(p.z + e);
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

  // ---- qLogEI: tabulated fat-tail sum (acq_math.cuh: mc_table_setup / mc_row_fast / mc_row_exact_warp) ----
  const bool fast_mc = mc_table_applicable(p.has_acq, p.acq, p.S);
  if (fast_mc) mc_table_setup(s.mc_part, s.z_s, p.S, p.acq.obj_scale < 0.f ? -1.f : 1.f);

  if (warp < kComputeWarps) {
    // =====================================================================================
    // compute warps: thread = TMEM lane (candidate row_e) x column group cg (16 of the 64
    // columns of a chunk; quarter of the MC samples)
    // =====================================================================================
    const int row_e = tid & 127, cg = tid >> 7;
    const int jg = cg;  // staging: dimension quads jg, jg+4, ...
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const bool scaled = p.scaled != 0;
    long long best = kEmptyKey;
    TcStageRegs regs;

    auto prefetch = [&](int tile) {
      const int64_t row = (int64_t)tile * kTileM + row_e;
#pragma unroll
      for (int u = 0; u < kTcStageQuads; ++u) {
        const int jq = jg + 4 * u;
        regs.v[u] = (jq < dq) ? tc_load_quad(p, row, jq) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // scaled candidate rows -> fp16 hi/mid/lo A2 tiles, |a|^2 partials, task ids
    auto stage_a2 = [&](int buf) {
      float an = 0.f;
#pragma unroll
      for (int u = 0; u < kTcStageQuads; ++u) {
        const int jq = jg + 4 * u;
        if (jq < dq) {
          const int j0 = jq * 4;
          const float4 q = regs.v[u];
          if (p.task_col >= j0 && p.task_col < j0 + 4) {
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
          for (int e = 0; e < 4; ++e) a[e] *= p.dist_scale_a;  // exact: power of two
          uint2 hi, mid, lo;
          split3_quad(a, hi, mid, lo);
          const uint32_t off = swk_offset<K2>((uint32_t)row_e, (uint32_t)(jq >> 1)) + (uint32_t)(jq & 1) * 8u;
          *reinterpret_cast<uint2*>(s.a2 + off) = hi;
          *reinterpret_cast<uint2*>(s.a2 + kA2Split + off) = mid;
          *reinterpret_cast<uint2*>(s.a2 + 2 * kA2Split + off) = lo;
        }
      }
      s.an_part[(buf * 4 + jg) * kTileM + row_e] = an;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.a2_full);
    };

    // Monte-Carlo work of the PREVIOUS tile is interleaved into the chunk loop of the current one:
    // it fills the time the tensor cores need to drain the A-operand ring and keeps MUFU/FMA busy
    // while this tile's distance tiles are converted.
    const bool is_mc = p.has_acq && p.acq.kind <= BB_ACQ_QPI;
    const int per = is_mc ? p.S / 4 : 0;            // samples of this thread's group
    const float4* zq = reinterpret_cast<const float4*>(s.z_s + cg * per);
    const int n4_total = per >> 2;
    bool have_prev = false;
    float mu_p = 0.f, var_p = 0.f, c0_p = 0.f, c1_p = 0.f, s0_p = 0.f, s1_p = 0.f;
    int64_t row0_p = 0;
    int mc_done = 0;  // float4 groups of the previous tile's samples already accumulated

    float* mc_c0 = s.mc_part + kMcRows;            // exact (s0, s1) of the previous tile's rows outside the
    float* mc_c1 = s.mc_part + kMcRows + kTileM;   // tabulated envelope, handed to the row's cg == 0 thread
    bool slow_p = false;
    // qLogEI, tabulated path: every thread of a row (all four cg) evaluates it redundantly right after the
    // moments are known -- ~170 instructions, no shared memory, no barrier, no idle warps
    auto mc_row = [&]() {
      float s0, s1;
      const bool fast = mc_row_fast(s.mc_part, c0_p, c1_p, s0, s1);
      slow_p = !fast;
      s0_p = s0;
      s1_p = s1;
      if (p.trace != nullptr && cg == 0)  // test-only: rows outside the tabulated envelope / rows seen
        atomicAdd(reinterpret_cast<unsigned long long*>(p.trace) + 1 + 2 * p.trace_cap + (fast ? 1 : 0), 1ull);
    };
    // exact sums of the rows outside the envelope: the four warps that share a row group (same rows, cg = 0..3)
    // see the same ballot and take every fourth such row each; the 32 lanes split the row's samples (fixed
    // assignment and reduction order: a row's value does not depend on its position)
    auto mc_rows_exact = [&](int part, int nparts) {
      unsigned need = __ballot_sync(0xffffffffu, slow_p);
      for (int rank = 0; need != 0u; ++rank) {
        const int b = __ffs(need) - 1;
        need &= need - 1u;
        if ((rank & 3) != cg || ((rank >> 2) % nparts) != part) continue;
        const float c0 = __shfl_sync(0xffffffffu, c0_p, b), c1 = __shfl_sync(0xffffffffu, c1_p, b);
        float a0, a1;
        mc_row_exact_warp(s.z_s, p.S, c0, c1, lane, a0, a1);
        if (lane == b) {
          mc_c0[row_e] = a0;
          mc_c1[row_e] = a1;
        }
      }
    };
    auto mc_slice = [&](int part, int nparts) {  // accumulate slice `part` of `nparts`
      if (!have_prev || !is_mc) return;
      if (fast_mc) {
        mc_rows_exact(part, nparts);
        return;
      }
      const int lo = (n4_total * part) / nparts, hi = (n4_total * (part + 1)) / nparts;
      mc_accumulate(p.acq.kind, c0_p, c1_p, zq + lo, hi - lo, s0_p, s1_p);
      mc_done = hi;
    };
    // combine the four sample groups of the previous tile, score, arg-max
    auto finish_prev = [&]() {
      if (!p.has_acq) return;
      if (is_mc && !fast_mc) *reinterpret_cast<float2*>(s.mc_part + (cg * kTileM + row_e) * 2) = make_float2(s0_p, s1_p);
      bar_compute();
      if (cg == 0 && have_prev) {
        float score;
        if (is_mc) {
          float s0 = 0.f, s1 = 0.f;
          if (fast_mc) {
            s0 = slow_p ? mc_c0[row_e] : s0_p;
            s1 = slow_p ? mc_c1[row_e] : s1_p;
          } else {
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              const float2 pr = *reinterpret_cast<const float2*>(s.mc_part + (gg * kTileM + row_e) * 2);
              s0 += pr.x;
              s1 += pr.y;
            }
          }
          score = mc_finalize(p.acq, mu_p, var_p, s0, s1, p.S, s.zstat[0], s.zstat[1]);
        } else {
          score = analytic_value(p.acq, mu_p, var_p);
        }
        const int64_t row = row0_p + row_e;
        if (row < p.N) {
          if (p.score) p.score[row] = score;
          const bool ok = (p.keep == nullptr || p.keep[row] != 0) && !(score != score);
          if (ok) {
            const long long key = pack_key(score, (uint32_t)(row + p.index_offset));
            best = key > best ? key : best;
          }
        }
      }
    };

    uint32_t slot = 0, ph = 0;
    int it = 0;
    int tile = blockIdx.x;
    if (tile < p.num_tiles) {
      prefetch(tile);
      stage_a2(0);
      if (tile + (int)gridDim.x < p.num_tiles) prefetch(tile + gridDim.x);
    }
    bar_compute();
    for (; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t par = (uint32_t)(it & 1);
      const int64_t row0 = (int64_t)tile * kTileM;
      const float* anp = s.an_part + (buf * 4) * kTileM + row_e;
      const float an = (anp[0] + anp[kTileM]) + (anp[2 * kTileM] + anp[3 * kTileM]);
      const int ct = s.cand_task[buf * kTileM + row_e];
      const float* tcrow = s.tcov + ct * p.n_tasks;
      float mean = 0.f;
      // ---- K* chunk by chunk: D2 (TMEM) -> kernel values -> fp16 hi/lo A operand ----
      for (int c = 0; c < C; ++c) {
        float v[16];
        if (tid == 0) trace_ev(p, it, 100 + c);
        if (c == 0) {
          mbar_wait(s.d2_full, par);
          tc_fence_after();
        }
        tmem_ld16(tmem_base + lane_base + kD2Col0 + (uint32_t)(c * kChunk + cg * 16), v);
        tmem_ld_wait();
        if (c == C - 1) {  // all of D2 is in registers now: the next tile's distance GEMM may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s.d2_empty);
        }
        if (tid == 0) trace_ev(p, it, 110 + c);
        const int i0 = c * kChunk + cg * 16;
        float k[16];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const float4 bn = *reinterpret_cast<const float4*>(s.tsq + i0 + e4 * 4);
          const float4 al = *reinterpret_cast<const float4*>(s.alpha_s + i0 + e4 * 4);
          const float bb4[4] = {bn.x, bn.y, bn.z, bn.w}, al4[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ii = e4 * 4 + e;
            const float t = fmaf(v[ii], p.inv_dist_scale, an + bb4[e]);
            float kv = kernel_from_t<FAMILY>(t);
            if (scaled) kv *= tcrow[s.ttask[i0 + ii]];
            k[ii] = kv;
            mean = fmaf(kv, al4[e], mean);
          }
        }
        uint4 h0, l0, h1, l1;
        split_pair(k[0], k[1], h0.x, l0.x);
        split_pair(k[2], k[3], h0.y, l0.y);
        split_pair(k[4], k[5], h0.z, l0.z);
        split_pair(k[6], k[7], h0.w, l0.w);
        split_pair(k[8], k[9], h1.x, l1.x);
        split_pair(k[10], k[11], h1.y, l1.y);
        split_pair(k[12], k[13], h1.z, l1.z);
        split_pair(k[14], k[15], h1.w, l1.w);
        if (tid == 0) trace_ev(p, it, 120 + c);
        if (!fast_mc) mc_slice(c, C);  // previous tile's MC slice: overlaps the MMA drain of the A ring
        if (tid == 0) trace_ev(p, it, 130 + c);
        mbar_wait(&s.a_empty[slot], ph ^ 1u);
        if (tid == 0) trace_ev(p, it, 140 + c);
        uint8_t* sa = s.ring_a + (size_t)slot * kSlotABytes;
        const uint32_t o0 = sw128_offset((uint32_t)row_e, (uint32_t)(2 * cg));
        const uint32_t o1 = sw128_offset((uint32_t)row_e, (uint32_t)(2 * cg + 1));
        *reinterpret_cast<uint4*>(sa + o0) = h0;
        *reinterpret_cast<uint4*>(sa + o1) = h1;
        *reinterpret_cast<uint4*>(sa + 16384 + o0) = l0;
        *reinterpret_cast<uint4*>(sa + 16384 + o1) = l1;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.a_full[slot]);
        if (fast_mc) mc_slice(c, C);  // the few exact-sum rows: behind the arrive, so a straggler warp does not hold the MMAs
        if (tid == 0) trace_ev(p, it, 150 + c);
        if (++slot == (uint32_t)kTcSlotsA) {
          slot = 0;
          ph ^= 1u;
        }
      }
      s.mean_part[cg * kTileM + row_e] = mean;

      // ---- finish the previous tile (its partial buffers are free again afterwards) ----
      finish_prev();
      if (tid == 0) trace_ev(p, it, 160);

      // ---- stage the next tile's A2 so that its distance GEMM runs under this epilogue ----
      const int next = tile + (int)gridDim.x;
      if (next < p.num_tiles) {
        mbar_wait(s.a2_empty, par);  // distance GEMMs of the current tile have consumed A2
        stage_a2(buf ^ 1);
        if (next + (int)gridDim.x < p.num_tiles) prefetch(next + gridDim.x);
      }

      if (tid == 0) trace_ev(p, it, 161);
      // ---- |V|^2: every 64-column sub-block is final as soon as its diagonal chunk is done ----
      {
        float ss = 0.f;
        for (int sb = 0; sb < C; ++sb) {
          float v[16];
          mbar_wait(&s.dsub_full[sb], par);
          if (tid == 0) trace_ev(p, it, 170 + sb);
          tc_fence_after();
          tmem_ld16(tmem_base + lane_base + (uint32_t)(sb * kChunk + cg * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) ss = fmaf(v[e], v[e], ss);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s.d_empty);
        s.var_part[cg * kTileM + row_e] = ss;
      }
      bar_compute();
      if (tid == 0) trace_ev(p, it, 180);

      // ---- moments of this tile in original units; its acquisition work runs next iteration ----
      float msum = s.meanc[ct];
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) msum += s.mean_part[gg * kTileM + row_e];
      const float vsum = (s.var_part[row_e] + s.var_part[kTileM + row_e]) +
                         (s.var_part[2 * kTileM + row_e] + s.var_part[3 * kTileM + row_e]);
      const float kss = scaled ? s.tcov[ct * p.n_tasks + ct] : 1.0f;
      const float var_t = fmaxf(kss - vsum * p.inv_r_scale2, 1e-10f);
      mu_p = fmaf(p.y_std, msum, p.y_mean);
      var_p = p.y_std * p.y_std * var_t;
      row0_p = row0;
      have_prev = true;
      s0_p = 0.f;
      s1_p = 0.f;
      mc_done = 0;
      if (is_mc) mc_coef(p.acq, mu_p, var_p, c0_p, c1_p);
      if (cg == 0 && row0 + row_e < p.N) {
        if (p.mu) p.mu[row0 + row_e] = mu_p;
        if (p.var) p.var[row0 + row_e] = var_p;
      }
      if (fast_mc) mc_row();
    }
    // ---- drain: acquisition work of the last tile ----
    if (have_prev) {
      for (int c = 0; c < C; ++c) mc_slice(c, C);  // same slices and order as the interleaved path:
      finish_prev();                               // a row's score must not depend on its tile's position
    }
    if (p.best_key != nullptr && p.has_acq) {
      if (cg == 0) {
        for (int o = 16; o > 0; o >>= 1) {
          const long long other = __shfl_xor_sync(0xffffffffu, best, o);
          best = other > best ? other : best;
        }
        if (lane == 0) s.best_red[warp] = best;
      }
      bar_compute();
      if (tid == 0) {
        long long b = s.best_red[0];
        for (int w = 1; w < 4; ++w) b = s.best_red[w] > b ? s.best_red[w] : b;
        if (b != kEmptyKey) atomicMax(p.best_key, b);
      }
    }
  } else if (warp == kWarpProducer) {
    // =====================================================================================
    // producer (TMA engine): streams the L^-1 tiles (same sequence for every candidate tile)
    // =====================================================================================
    if (elect_one()) {  // elect.sync: straight UBLKCP (fused_common.cuh)
      uint32_t rs = 0, rph = 0;
      int pit = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++pit) {
        size_t off = 0;
        int tb = 0;
        for (int c = 0; c < C; ++c) {
          for (int sb = c; sb < C; ++tb) {
            const int g = ((sb & 1) == 0 && sb + 1 < C) ? 2 : 1;
            const uint32_t bytes = (uint32_t)g * 16384u;
            mbar_wait_relaxed(&s.r_empty[rs], rph ^ 1u);
            trace_ev(p, pit, 300 + tb);
            mbar_expect_tx(&s.r_full[rs], bytes);
            bulk_g2s(s.ring_r + (size_t)rs * kRStageBytes, p.rimg2 + off, bytes, &s.r_full[rs]);
            off += bytes;
            sb += g;
            if (++rs == (uint32_t)p.stages_b) {
              rs = 0;
              rph ^= 1u;
            }
          }
        }
      }
    }
  } else {
    // =====================================================================================
    // MMA issuer: the whole warp runs the loop converged, one lane issues under elect.sync (issued under
    // `if (lane == 0)` every tcgen05.mma is wrapped in an ELECT/R2UR/BRA.U.ANY loop, ~106 cycles per instruction)
    // =====================================================================================
    {
      const uint32_t idesc = make_idesc_f16(kTileM, kChunk), idesc128 = make_idesc_f16(kTileM, 2 * kChunk);
      const uint32_t idesc_d2 = make_idesc_f16(kTileM, p.n_pad);  // one MMA spans all training columns
      uint32_t slot = 0, pha = 0, rs = 0, rph = 0;
      const uint32_t a2_addr = smem_u32(s.a2), bt_addr = smem_u32(s.bt);
      const uint64_t a2_h = make_swk_desc<K2>(a2_addr), a2_m = make_swk_desc<K2>(a2_addr + kA2Split),
                     a2_l = make_swk_desc<K2>(a2_addr + 2 * kA2Split);
      const uint64_t b_h = make_swk_desc<K2>(bt_addr), b_m = make_swk_desc<K2>(bt_addr + kBtSplit),
                     b_l = make_swk_desc<K2>(bt_addr + 2 * kBtSplit);
      // distance GEMM of tile number j: D2 = A2 * Bt^T, six split products, N = n_pad per MMA
      auto issue_distance = [&](int j) {
        const uint32_t par = (uint32_t)(j & 1);
        if (lane == 0) trace_ev(p, j, 250);
        mbar_wait_relaxed(s.a2_full, par);
        if (lane == 0) trace_ev(p, j, 251);
        mbar_wait_relaxed(s.d2_empty, par ^ 1u);  // D2 drained by the previous tile's chunk loop
        if (lane == 0) trace_ev(p, j, 252);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + kD2Col0;
        if (elect_one()) {
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t ko = (uint64_t)(kk * 2);
            umma_f16(d_addr, a2_h + ko, b_h + ko, idesc_d2, kk > 0 ? 1u : 0u);
            umma_f16(d_addr, a2_h + ko, b_m + ko, idesc_d2, 1u);
            umma_f16(d_addr, a2_m + ko, b_h + ko, idesc_d2, 1u);
            umma_f16(d_addr, a2_h + ko, b_l + ko, idesc_d2, 1u);
            umma_f16(d_addr, a2_l + ko, b_h + ko, idesc_d2, 1u);
            umma_f16(d_addr, a2_m + ko, b_m + ko, idesc_d2, 1u);
          }
          umma_commit(s.d2_full);
          umma_commit(s.a2_empty);
        }
        __syncwarp();
        if (lane == 0) trace_ev(p, j, 253);
      };
      int j = 0;
      int tile = blockIdx.x;
      if (tile < p.num_tiles) issue_distance(0);
      for (; tile < p.num_tiles; tile += gridDim.x, ++j) {
        mbar_wait_relaxed(s.d_empty, (uint32_t)((j & 1) ^ 1));  // previous epilogue drained V
        if (lane == 0) trace_ev(p, j, 200);
        tc_fence_after();
        for (int c = 0; c < C; ++c) {
          mbar_wait_relaxed(&s.a_full[slot], pha);
          if (lane == 0) trace_ev(p, j, 210 + c);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(s.ring_a + (size_t)slot * kSlotABytes);
          const uint64_t a_hi = make_sw128_desc(a_addr), a_lo = make_sw128_desc(a_addr + 16384);
          // adjacent column sub-blocks are issued as one 128-column MMA wherever the lower-triangular structure
          // allows a pair
          for (int sb = c; sb < C;) {
            const int g = ((sb & 1) == 0 && sb + 1 < C) ? 2 : 1;
            mbar_wait_relaxed(&s.r_full[rs], rph);
            if (lane == 0) trace_ev(p, j, 220 + c * 4 + sb);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s.ring_r + (size_t)rs * kRStageBytes);
            const uint64_t b_hi = make_sw128_desc(b_addr), b_lo = make_sw128_desc(b_addr + (uint32_t)g * 8192u);
            const uint32_t d_addr = tmem_base + (uint32_t)(sb * kChunk);
            const uint32_t id = (g == 2) ? idesc128 : idesc;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ko = (uint64_t)(kk * 2);
                umma_f16(d_addr, a_hi + ko, b_hi + ko, id, (c > 0 || kk > 0) ? 1u : 0u);
                umma_f16(d_addr, a_hi + ko, b_lo + ko, id, 1u);
                umma_f16(d_addr, a_lo + ko, b_hi + ko, id, 1u);
              }
              umma_commit(&s.r_empty[rs]);
            }
            __syncwarp();
            if (++rs == (uint32_t)p.stages_b) {
              rs = 0;
              rph ^= 1u;
            }
            sb += g;
          }
          if (elect_one()) {
            umma_commit(&s.a_empty[slot]);
            umma_commit(&s.dsub_full[c]);  // sub-block c of V has received its last contribution
          }
          __syncwarp();
          if (lane == 0) trace_ev(p, j, 240 + c);
          if (++slot == (uint32_t)kTcSlotsA) {
            slot = 0;
            pha ^= 1u;
          }
        }
        if (tile + (int)gridDim.x < p.num_tiles) issue_distance(j + 1);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpProducer) tmem_dealloc(tmem_base, 512);
}

// Picks the deepest L^-1 ring (p.stages_b) that fits next to the fixed buffers; false if the shape
// is outside this kernel's envelope.
bool fused_tc_supported(FusedParams& p, int max_smem) {
  if (p.n_pad > 256 || p.d_pad > 64 || p.family == BB_KERNEL_MATERN12) return false;
  if (p.has_acq && p.S > kTcMaxSamples) return false;
  if (p.bimg == nullptr || p.rimg2 == nullptr || (p.dist_k != 32 && p.dist_k != 64) || p.d_pad > p.dist_k) return false;
  for (int st = kTcMaxStagesR; st >= 2; --st) {
    p.stages_b = st;
    if (tc_carve(nullptr, p, nullptr) + 2048 <= (size_t)max_smem) return true;
  }
  return false;
}

template <int FAMILY, int K2>
static int launch_tc_one(FusedParams& p, int grid, size_t smem, cudaStream_t stream) {
  BB_SMEM_OPTIN_ONCE((k_fused_tc<FAMILY, K2>));
  k_fused_tc<FAMILY, K2><<<grid, kFusedThreads, smem, stream>>>(p);
  BB_LAUNCH_CHECK();
  return BB_OK;
}

template <int FAMILY>
static int launch_tc_family(FusedParams& p, int grid, size_t smem, cudaStream_t stream) {
  return p.dist_k == 32 ? launch_tc_one<FAMILY, 32>(p, grid, smem, stream)
                        : launch_tc_one<FAMILY, 64>(p, grid, smem, stream);
}

int launch_fused_tc(FusedParams& p, int grid, cudaStream_t stream) {
  const size_t smem = tc_carve(nullptr, p, nullptr);
  switch (p.family) {
    case BB_KERNEL_MATERN32: return launch_tc_family<BB_KERNEL_MATERN32>(p, grid, smem, stream);
    case BB_KERNEL_MATERN52: return launch_tc_family<BB_KERNEL_MATERN52>(p, grid, smem, stream);
    default: return launch_tc_family<BB_KERNEL_RBF>(p, grid, smem, stream);
  }
}

}  // namespace bb
