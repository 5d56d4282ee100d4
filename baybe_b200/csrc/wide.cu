// wide.cu -- K(X*, X) for WIDE feature spaces (d in the hundreds or thousands: substance fingerprints,
// BASELINE config 4) on the tensor cores, and the two-stage scoring path built on it.
//
//   stage 1  k_kmat_tc   t[m][i] = |a_m|^2 + |b_i|^2 - 2 a_m.b_i  with the inner product as a
//                        K-looped tcgen05 GEMM (fp16 split operands, fp32 accumulators in TMEM),
//                        Matern/RBF epilogue, K* block -> L2-resident workspace (or the caller's
//                        matrix for bb_kernel_matrix)
//   stage 2  k_fused<PRE> (fused.cu) posterior GEMM + acquisition + arg-max reading that K* block
//
// Candidate layouts: the four float layouts (operand split hi/mid/lo, six products, error 2^-33)
// and BB_BITS_U8 -- bit-packed binary fingerprints.  For bits x in {0,1}^d the whole scaled squared
// distance is LINEAR in x:  t = sum_j x_j W_ij + c_i,  W_ij = s_j (-2 b_ij + s_j + 2 h_j),
// c_i = |b_i|^2 + sum_j h_j (-2 b_ij + h_j)   (a_j = s_j x_j + h_j), so the A operand is the exact
// 0/1 matrix (one fp16 panel, no split) and only W is split (hi/mid: W takes at most two distinct
// values per column, 2^-22 relative is ample): two products.
//
// Work item = 256 candidates x <=256 training columns: two UMMA_M=128 accumulators share every
// B slice (halves the L2 traffic per flop), 2 x 256 TMEM columns.  Warp roles as in fused.cu.
//
// Reference path replaced: gpytorch Kernel.forward over the comp-rep of a SubstanceParameter space
// (/root/reference/baybe/kernels/base.py:173-178, parameters/substance.py comp_df), reached from
// SingleTaskGP.posterior (surrogates/gaussian_process/core.py:268-269).
#include "fused_common.cuh"

namespace bb {

constexpr int kWTileM = 256;   // candidates per work item
constexpr int kWHalfN = 256;   // training columns per work item
constexpr int kWK = 32;        // fp16 per K stage: 64-byte rows, SWIZZLE_64B
constexpr uint32_t kWPanelA = kWTileM * kWK * 2;  // 16 KB: 256 rows x 64 B
constexpr uint32_t kWPanelB = kWHalfN * kWK * 2;  // 16 KB: 256 rows x 64 B; a stage holds [hi | mid | lo] (floats) or [hi | mid] (bits)

struct WideParams {
  const void* x;
  int layout;
  int64_t N, ldx;
  int d, n, n_pad, n_halves, n_kc;
  const float *cand_scale, *cand_shift;  // [d]
  const uint8_t* wimg;                   // B image: per (half, K stage) [hi|mid|lo] swizzled panels
  const float* wnorm;                    // [n_pad] additive per-training-row term (|b|^2 or c_i)
  float a_scale, inv_scale;
  const int32_t* train_task;
  const float* task_covar;
  int task_col, n_tasks, scaled;
  float* out;
  int64_t ldk, out_rows;
  int out_cols, vec_ok, bits_vec;
  int num_items, stages;
};

struct WideSmem {
  uint8_t* ring;
  float *wnorm_s, *tcov;
  double* an_part;  // [2 buffers][2 K halves][256 rows], float64: |a|^2 sums d terms
  int32_t* ttask;
  uint64_t *full, *empty, *acc_full, *acc_empty;
  uint32_t* tmem_ptr;
};

template <bool BITS>
__host__ __device__ inline size_t wide_carve(uint8_t* base, const WideParams& p, WideSmem* s) {
  constexpr uint32_t kStage = (BITS ? 1 : 3) * kWPanelA + (BITS ? 2 : 3) * kWPanelB;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 15) / 16 * 16;
    return o;
  };
  const size_t o_ring = take((size_t)p.stages * kStage);
  const size_t o_wn = take((size_t)p.n_pad * 4), o_tt = take((size_t)p.n_pad * 4);
  const size_t o_an = take(2 * 2 * kWTileM * 8), o_tc = take(kMaxTasks * kMaxTasks * 4);
  const size_t o_bar = take(16 * 8), o_misc = take(16);
  if (s) {
    s->ring = base + o_ring;
    s->wnorm_s = reinterpret_cast<float*>(base + o_wn);
    s->ttask = reinterpret_cast<int32_t*>(base + o_tt);
    s->an_part = reinterpret_cast<double*>(base + o_an);
    s->tcov = reinterpret_cast<float*>(base + o_tc);
    uint64_t* b = reinterpret_cast<uint64_t*>(base + o_bar);
    s->full = b;           // [<=4]
    s->empty = b + 4;      // [<=4]
    s->acc_full = b + 8;   // [1]
    s->acc_empty = b + 9;  // [1]
    s->tmem_ptr = reinterpret_cast<uint32_t*>(base + o_misc);
  }
  return off;
}

// 16 consecutive features k0..k0+15 of one candidate row as fp32 (0 beyond d / beyond N).
__device__ __forceinline__ void wide_load16(const WideParams& p, int64_t row, int k0, float (&v)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = 0.f;
  if (row >= p.N || k0 >= p.d) return;
  switch (p.layout) {
    case BB_ROW_MAJOR_F32: {
      const float* ptr = reinterpret_cast<const float*>(p.x) + row * p.ldx + k0;
      if (k0 + 16 <= p.d && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(ptr) + q);
          v[4 * q] = f.x;
          v[4 * q + 1] = f.y;
          v[4 * q + 2] = f.z;
          v[4 * q + 3] = f.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (k0 + e < p.d) v[e] = __ldg(ptr + e);
      }
      return;
    }
    case BB_COL_MAJOR_F32:
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (k0 + e < p.d) v[e] = load_x<BB_COL_MAJOR_F32>(p.x, row, k0 + e, p.ldx);
      return;
    case BB_ROW_MAJOR_F64:
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (k0 + e < p.d) v[e] = load_x<BB_ROW_MAJOR_F64>(p.x, row, k0 + e, p.ldx);
      return;
    default:
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (k0 + e < p.d) v[e] = load_x<BB_COL_MAJOR_F64>(p.x, row, k0 + e, p.ldx);
      return;
  }
}

// 16 feature bits k0..k0+15 (k0 a multiple of 16) of one bit-packed candidate row.
__device__ __forceinline__ uint32_t wide_load_bits16(const WideParams& p, int64_t row, int k0) {
  if (row >= p.N || k0 >= p.d) return 0u;
  const uint8_t* ptr = reinterpret_cast<const uint8_t*>(p.x) + row * p.ldx + (k0 >> 3);
  uint32_t b = __ldg(ptr);
  if (k0 + 8 < p.d) b |= (uint32_t)__ldg(ptr + 1) << 8;
  const int valid = p.d - k0;  // bits beyond d are padding of the last byte
  if (valid < 16) b &= (1u << valid) - 1u;
  return b;
}

template <int FAMILY, bool BITS>
__global__ void __launch_bounds__(kFusedThreads, 1) k_kmat_tc(const WideParams p) {
  constexpr int PA = BITS ? 1 : 3;
  constexpr int PB = BITS ? 2 : 3;  // W of the bit-linear form has <= 2 distinct values per column: hi+mid (2^-22) suffices
  constexpr uint32_t kStageA = PA * kWPanelA;
  constexpr uint32_t kStage = kStageA + PB * kWPanelB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WideSmem s;
  wide_carve<BITS>(smem_raw, p, &s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();

  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&s.full[i], kComputeWarps + 1);  // 16 staging warps + the producer's expect_tx
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(s.acc_full, 1);
    mbar_init(s.acc_empty, kComputeWarps);
    fence_mbar_init();
  }
  if (warp == kWarpProducer) {
    tmem_alloc(s.tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int e = tid; e < p.n_pad; e += kFusedThreads) {
    s.wnorm_s[e] = __ldg(p.wnorm + e);
    s.ttask[e] = __ldg(p.train_task + e);
  }
  for (int e = tid; e < p.n_tasks * p.n_tasks; e += kFusedThreads) s.tcov[e] = __ldg(p.task_covar + e);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;

  if (warp < kComputeWarps) {
    // =====================================================================================
    // compute warps: stage the A operand (candidate rows -> fp16 panels), then the epilogue
    // =====================================================================================
    const int r = tid & 255, kh = tid >> 8;        // staging: row of the item, 16-feature half of the stage
    const int row_e = tid & 127, cg = tid >> 7;    // epilogue: TMEM lane, 64-column group
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t st = 0, ph = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int tile = item / p.n_halves, half = item - tile * p.n_halves;
      const int64_t row0 = (int64_t)tile * kWTileM;
      const int64_t row = row0 + r;
      const int ncols = min(kWHalfN, p.n_pad - half * kWHalfN);
      double an = 0.0;
      float cur[16];
      uint32_t curb = 0;
      // bit rows 16-byte aligned: one 128-bit load covers four K stages and is issued four stages ahead
      // (a one-stage-ahead load would put a full L2/HBM latency on every stage of the staging warps)
      uint4 cur4 = make_uint4(0u, 0u, 0u, 0u), nxt4 = cur4;
      const uint4* brow = nullptr;
      if constexpr (BITS) {
        if (p.bits_vec) {
          if (row < p.N) brow = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.x) + row * p.ldx);
          if (brow != nullptr) {
            cur4 = __ldg(brow);
            if (p.n_kc > 4) nxt4 = __ldg(brow + 1);
          }
        } else {
          curb = wide_load_bits16(p, row, kh * 16);
        }
      }
      else wide_load16(p, row, kh * 16, cur);
      for (int kc = 0; kc < p.n_kc; ++kc) {
        const int k0 = kc * kWK + kh * 16;
        uint4 pk[PA][2];
        if constexpr (BITS) {
          if (p.bits_vec) {
            const int sub = kc & 3;
            const uint32_t word = sub == 0 ? cur4.x : sub == 1 ? cur4.y : sub == 2 ? cur4.z : cur4.w;
            curb = (word >> (kh * 16)) & 0xffffu;
            if (sub == 3) {
              cur4 = nxt4;
              if (brow != nullptr && kc + 5 < p.n_kc) nxt4 = __ldg(brow + ((kc + 5) >> 2));
            }
          }
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            w[e] = (((curb >> (2 * e)) & 1u) * 0x3C00u) | (((curb >> (2 * e + 1)) & 1u) * 0x3C000000u);
          pk[0][0] = make_uint4(w[0], w[1], w[2], w[3]);
          pk[0][1] = make_uint4(w[4], w[5], w[6], w[7]);
          if (!p.bits_vec && kc + 1 < p.n_kc) curb = wide_load_bits16(p, row, k0 + kWK);
        } else {
          float a[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int k = k0 + e;
            const float sc = (k < p.d) ? __ldg(p.cand_scale + k) : 0.f;
            const float sh = (k < p.d) ? __ldg(p.cand_shift + k) : 0.f;
            const float v = fmaf(cur[e], sc, sh);
            an = fma((double)v, (double)v, an);
            a[e] = v * p.a_scale;
          }
          if (kc + 1 < p.n_kc) wide_load16(p, row, k0 + kWK, cur);  // next stage's loads fly under the split
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint2 h0, m0, l0, h1, m1, l1;
            const float q0[4] = {a[8 * h], a[8 * h + 1], a[8 * h + 2], a[8 * h + 3]};
            const float q1[4] = {a[8 * h + 4], a[8 * h + 5], a[8 * h + 6], a[8 * h + 7]};
            split3_quad(q0, h0, m0, l0);
            split3_quad(q1, h1, m1, l1);
            pk[0][h] = make_uint4(h0.x, h0.y, h1.x, h1.y);
            pk[PA > 1 ? 1 : 0][h] = make_uint4(m0.x, m0.y, m1.x, m1.y);
            pk[PA > 2 ? 2 : 0][h] = make_uint4(l0.x, l0.y, l1.x, l1.y);
          }
        }
        mbar_wait(&s.empty[st], ph ^ 1u);  // the MMAs that read this stage last time are done
        uint8_t* sa = s.ring + (size_t)st * kStage;
#pragma unroll
        for (int pa = 0; pa < PA; ++pa) {
          *reinterpret_cast<uint4*>(sa + pa * kWPanelA + swk_offset<kWK>((uint32_t)r, (uint32_t)(kh * 2))) = pk[pa][0];
          *reinterpret_cast<uint4*>(sa + pa * kWPanelA + swk_offset<kWK>((uint32_t)r, (uint32_t)(kh * 2 + 1))) = pk[pa][1];
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.full[st]);
        if (++st == (uint32_t)p.stages) {
          st = 0;
          ph ^= 1u;
        }
      }
      double* anp = s.an_part + (it & 1) * 2 * kWTileM;
      anp[kh * kWTileM + r] = an;

      // ---- epilogue: t -> k(t) -> K* block ----
      mbar_wait(s.acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      bar_compute();  // an_part complete
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        const int64_t grow = row0 + m * 128 + row_e;
        const float an_r = (float)(anp[m * 128 + row_e] + anp[kWTileM + m * 128 + row_e]);
        int ct = 0;
        if (p.scaled && p.task_col >= 0 && grow < p.N) {
          float tv;
          switch (p.layout) {
            case BB_ROW_MAJOR_F32: tv = load_x<BB_ROW_MAJOR_F32>(p.x, grow, p.task_col, p.ldx); break;
            case BB_COL_MAJOR_F32: tv = load_x<BB_COL_MAJOR_F32>(p.x, grow, p.task_col, p.ldx); break;
            case BB_ROW_MAJOR_F64: tv = load_x<BB_ROW_MAJOR_F64>(p.x, grow, p.task_col, p.ldx); break;
            default: tv = load_x<BB_COL_MAJOR_F64>(p.x, grow, p.task_col, p.ldx); break;
          }
          ct = min(max(__float2int_rn(tv), 0), p.n_tasks - 1);
        }
        const float* tcrow = s.tcov + ct * p.n_tasks;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int col = cg * 64 + j * 16;
          if (col >= ncols) break;  // warp-uniform: cg and ncols are
          float v[16];
          tmem_ld16(tmem_base + lane_base + (uint32_t)(m * kWHalfN + col), v);
          tmem_ld_wait();
          const int i0 = half * kWHalfN + col;
          float k[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float t = fmaf(v[e], p.inv_scale, an_r + s.wnorm_s[i0 + e]);
            float kv = kernel_from_t<FAMILY>(t);
            if (p.scaled) kv *= tcrow[s.ttask[i0 + e]];
            k[e] = (i0 + e < p.n) ? kv : 0.f;
          }
          if (grow < p.out_rows) {
            float* dst = p.out + grow * p.ldk + i0;
            if (p.vec_ok && i0 + 16 <= p.out_cols) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(k[4 * q], k[4 * q + 1], k[4 * q + 2], k[4 * q + 3]);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (i0 + e < p.out_cols) dst[e] = k[e];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.acc_empty);
    }
  } else if (warp == kWarpProducer) {
    // =====================================================================================
    // producer: one bulk copy (TMA engine) per K stage of the B image
    // =====================================================================================
    if (elect_one()) {  // elect.sync: straight UBLKCP (fused_common.cuh)
      uint32_t st = 0, ph = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int half = item % p.n_halves;
        const int ncols = min(kWHalfN, p.n_pad - half * kWHalfN);
        const uint32_t bytes = (uint32_t)PB * (uint32_t)ncols * (kWK * 2);
        const uint8_t* src = p.wimg + (size_t)half * p.n_kc * (PB * kWPanelB);  // only the last half is narrower
        for (int kc = 0; kc < p.n_kc; ++kc) {
          mbar_wait_relaxed(&s.empty[st], ph ^ 1u);
          mbar_expect_tx(&s.full[st], bytes);
          bulk_g2s(s.ring + (size_t)st * kStage + kStageA, src + (size_t)kc * bytes, bytes, &s.full[st]);
          if (++st == (uint32_t)p.stages) {
            st = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else {
    // =====================================================================================
    // MMA issuer: D[m] (128 x ncols, fp32, TMEM) += A_panel[m] (128 x 32) * B_panel^T
    // converged warp, one lane issues under elect.sync (no ELECT/R2UR/BRA.U.ANY wrapper per MMA, fused_common.cuh)
    // =====================================================================================
    {
      uint32_t st = 0, ph = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int half = item % p.n_halves;
        const int ncols = min(kWHalfN, p.n_pad - half * kWHalfN);
        const uint32_t idesc = make_idesc_f16(kTileM, ncols);
        const uint32_t bsplit = (uint32_t)ncols * (kWK * 2);
        mbar_wait_relaxed(s.acc_empty, (uint32_t)((it & 1) ^ 1));  // previous epilogue drained TMEM
        tc_fence_after();
        for (int kc = 0; kc < p.n_kc; ++kc) {
          mbar_wait_relaxed(&s.full[st], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(s.ring + (size_t)st * kStage), b_base = a_base + kStageA;
          const uint64_t b_h = make_swk_desc<kWK>(b_base), b_m = make_swk_desc<kWK>(b_base + bsplit),
                         b_l = make_swk_desc<kWK>(b_base + (PB > 2 ? 2 : 0) * bsplit);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t ko = (uint64_t)(kk * 2);  // 16 fp16 = 32 bytes
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                const uint32_t d_addr = tmem_base + (uint32_t)(m * kWHalfN);
                const uint32_t a_m0 = a_base + (uint32_t)m * (128u * kWK * 2);
                const uint32_t acc = (kc > 0 || kk > 0) ? 1u : 0u;
                const uint64_t a_h = make_swk_desc<kWK>(a_m0);
                if constexpr (BITS) {
                  umma_f16(d_addr, a_h + ko, b_h + ko, idesc, acc);
                  umma_f16(d_addr, a_h + ko, b_m + ko, idesc, 1u);
                } else {
                  const uint64_t a_md = make_swk_desc<kWK>(a_m0 + kWPanelA), a_l = make_swk_desc<kWK>(a_m0 + 2 * kWPanelA);
                  umma_f16(d_addr, a_h + ko, b_h + ko, idesc, acc);
                  umma_f16(d_addr, a_h + ko, b_m + ko, idesc, 1u);
                  umma_f16(d_addr, a_md + ko, b_h + ko, idesc, 1u);
                  umma_f16(d_addr, a_h + ko, b_l + ko, idesc, 1u);
                  umma_f16(d_addr, a_l + ko, b_h + ko, idesc, 1u);
                  umma_f16(d_addr, a_md + ko, b_m + ko, idesc, 1u);
                }
              }
            }
            umma_commit(&s.empty[st]);
          }
          __syncwarp();
          if (++st == (uint32_t)p.stages) {
            st = 0;
            ph ^= 1u;
          }
        }
        if (elect_one()) umma_commit(s.acc_full);
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpProducer) tmem_dealloc(tmem_base, 512);
}

template <int FAMILY, bool BITS>
static int launch_kmat_one(WideParams& p, int sms, int max_smem, cudaStream_t stream) {
  p.stages = BITS ? 4 : 2;
  const size_t smem = wide_carve<BITS>(nullptr, p, nullptr);
  BB_CHECK_SUPPORTED(smem <= (size_t)max_smem, "wide kernel-matrix path: shared-memory budget exceeded (%zu bytes)", smem);
  BB_SMEM_OPTIN_ONCE((k_kmat_tc<FAMILY, BITS>));
  const int grid = p.num_items < sms ? p.num_items : sms;
  k_kmat_tc<FAMILY, BITS><<<grid, kFusedThreads, smem, stream>>>(p);
  BB_LAUNCH_CHECK();
  return BB_OK;
}

template <int FAMILY>
static int launch_kmat_family(WideParams& p, bool bits, int sms, int max_smem, cudaStream_t stream) {
  return bits ? launch_kmat_one<FAMILY, true>(p, sms, max_smem, stream)
              : launch_kmat_one<FAMILY, false>(p, sms, max_smem, stream);
}

// Column set of a K(X*, .) launch: the training rows (model images) or the pending points (scratch images).
struct WideColumns {
  const uint8_t* wimg;
  const float* wnorm;
  const int32_t* task;
  int n, n_pad;
  float inv_scale;
};

static int launch_kmat_cols(const bb_model* m, const WideColumns& c, const void* d_x, int32_t layout, int64_t N,
                            int64_t ldx, float* d_out, int64_t ldk, int64_t out_rows, int out_cols,
                            cudaStream_t stream) {
  const bool bits = layout == BB_BITS_U8;
  WideParams p;
  memset(&p, 0, sizeof(p));
  p.x = d_x;
  p.layout = layout;
  p.N = N;
  p.ldx = ldx;
  p.d = m->d;
  p.n = c.n;
  p.n_pad = c.n_pad;
  p.n_halves = (c.n_pad + kWHalfN - 1) / kWHalfN;
  p.n_kc = m->d_wide / kWK;
  p.cand_scale = m->d_cand_scale;
  p.cand_shift = m->d_cand_shift;
  p.wimg = c.wimg;
  p.wnorm = c.wnorm;
  p.a_scale = bits ? 1.0f : m->dist_scale_a;
  p.inv_scale = c.inv_scale;
  p.train_task = c.task;
  p.task_covar = m->d_task_covar;
  p.task_col = m->task_col;
  p.n_tasks = m->n_tasks;
  p.scaled = (m->task_col >= 0 || m->prior_scale != 1.0f) ? 1 : 0;
  p.out = d_out;
  p.ldk = ldk;
  p.out_rows = out_rows;
  p.out_cols = out_cols;
  p.vec_ok = ((ldk & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) ? 1 : 0;
  p.num_items = (int)((N + kWTileM - 1) / kWTileM) * p.n_halves;
  p.bits_vec = (bits && (ldx & 15) == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (m->d & 127) == 0) ? 1 : 0;
  int max_smem = 0, sms = 0;
  {
    const int rc_lim = device_limits(&sms, &max_smem);
    if (rc_lim != BB_OK) return rc_lim;
  }
  switch (m->family) {
    case BB_KERNEL_MATERN32: return launch_kmat_family<BB_KERNEL_MATERN32>(p, bits, sms, max_smem, stream);
    case BB_KERNEL_MATERN52: return launch_kmat_family<BB_KERNEL_MATERN52>(p, bits, sms, max_smem, stream);
    default: return launch_kmat_family<BB_KERNEL_RBF>(p, bits, sms, max_smem, stream);
  }
}

static int wide_checks(const bb_model* m, int32_t layout) {
  BB_CHECK_SUPPORTED(m->wide != 0, "model has no wide-feature images");
  BB_CHECK_SUPPORTED(m->family != BB_KERNEL_MATERN12,
                     "Matern-1/2 is not supported on the wide-feature path (GEMM-form distances are "
                     "singular at r = 0)");
  BB_CHECK_SUPPORTED(!(layout == BB_BITS_U8 && m->task_col >= 0), "bit-packed candidates cannot carry a task column");
  return BB_OK;
}

// K(X*[0..N), X) -> d_out[row * ldk + i]; rows < out_rows and columns < out_cols are written.
int launch_kmat_wide(const bb_model* m, const void* d_x, int32_t layout, int64_t N, int64_t ldx,
                     float* d_out, int64_t ldk, int64_t out_rows, int out_cols, cudaStream_t stream) {
  const int rc = wide_checks(m, layout);
  if (rc != BB_OK) return rc;
  const bool bits = layout == BB_BITS_U8;
  WideColumns c;
  c.wimg = reinterpret_cast<const uint8_t*>(bits ? m->d_wimg_bits : m->d_wimg);
  c.wnorm = bits ? m->d_wnorm_bits : m->d_train_sq;
  c.task = m->d_train_task;
  c.n = m->n;
  c.n_pad = m->n_pad;
  c.inv_scale = bits ? 1.0f / m->dist_scale_w : 1.0f / (m->dist_scale_a * m->dist_scale_b);
  return launch_kmat_cols(m, c, d_x, layout, N, ldx, d_out, ldk, out_rows, out_cols, stream);
}

// ------------------------------------------------------------------------------------------
// pending points (sequential greedy, K9 prologue) on the wide path: the <= 31 pending rows become a
// 64-column "training set" of their own -- k(x*, pending) comes out of the same tensor-core kernel --
// and the posterior cross-covariance  cov(x*, p_j) = k(x*, p_j) - K*(x*) . beta_j  is a small CUDA-core
// contraction over the K* block that is already in the workspace.
// ------------------------------------------------------------------------------------------
// One CTA per pending row pp < 64 (rows >= P are zero): operand image (float form: -2 b; bit-linear form:
// W = s (-2 b + s + 2 h)), additive norm term, task id.
__global__ void __launch_bounds__(256) k_build_pend_img(const float* __restrict__ pend_x, int P, int d, int d_wide,
                                                        const float* __restrict__ cscale, const float* __restrict__ cshift,
                                                        int bits, float scale, int task_col, int n_tasks,
                                                        uint8_t* __restrict__ img, float* __restrict__ norm,
                                                        int32_t* __restrict__ task) {
  __shared__ double red[256];
  const int pp = blockIdx.x, panels = bits ? 2 : 3;
  const size_t panel = 64 * 64;  // bytes: 64 rows x 32 fp16
  double acc = 0.0;
  for (int k = threadIdx.x; k < d_wide; k += blockDim.x) {
    float b = 0.f, sk = 0.f, hk = 0.f;
    if (pp < P && k < d) {
      sk = __ldg(cscale + k);
      hk = __ldg(cshift + k);
      b = fmaf(__ldg(pend_x + (size_t)pp * d + k), sk, hk);
    }
    float v;
    if (bits) {
      v = sk * (-2.0f * b + sk + 2.0f * hk);
      acc += (double)b * b + (double)hk * (-2.0 * (double)b + (double)hk);
    } else {
      v = -2.0f * b;
      acc += (double)b * b;
    }
    v *= scale;
    const int kc = k >> 5, kl = k & 31;
    const size_t base = (size_t)kc * panels * panel;
    const uint32_t off = swk_offset<32>((uint32_t)pp, (uint32_t)(kl >> 3)) + (uint32_t)(kl & 7) * 2u;
    const __half h = __float2half_rn(v);
    const float r1 = v - __half2float(h);
    const __half mid = __float2half_rn(r1);
    *reinterpret_cast<__half*>(img + base + off) = h;
    *reinterpret_cast<__half*>(img + base + panel + off) = mid;
    if (panels > 2) *reinterpret_cast<__half*>(img + base + 2 * panel + off) = __float2half_rn(r1 - __half2float(mid));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    norm[pp] = pp < P ? (float)red[0] : 0.f;
    int t = 0;
    if (task_col >= 0 && pp < P)
      t = min(max(__float2int_rn(__ldg(pend_x + (size_t)pp * d + task_col)), 0), n_tasks - 1);
    task[pp] = t;
  }
}

// cross[row][j] = y_std^2 * ( kpend[row][j] - sum_i kstar[row][i] beta[j][i] ): one warp per candidate row.
__global__ void __launch_bounds__(256) k_cross_pre(const float* __restrict__ kstar, const float* __restrict__ kpend,
                                                   const float* __restrict__ beta, int P, int n_pad, int64_t nrows,
                                                   float s2, float* __restrict__ cross) {
  extern __shared__ float beta_s[];  // [P][n_pad]
  for (int e = threadIdx.x; e < P * n_pad; e += blockDim.x) beta_s[e] = __ldg(beta + e);
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * wpb + wib; row < nrows; row += (int64_t)gridDim.x * wpb) {
    float acc[BB_MAX_PENDING];
#pragma unroll
    for (int j = 0; j < BB_MAX_PENDING; ++j) acc[j] = 0.f;
    const float* kr = kstar + row * n_pad;
    for (int i = lane; i < n_pad; i += 32) {
      const float kv = kr[i];
#pragma unroll
      for (int j = 0; j < BB_MAX_PENDING; ++j)
        if (j < P) acc[j] = fmaf(kv, beta_s[j * n_pad + i], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < BB_MAX_PENDING; ++j) {
      if (j < P) {
        float v = acc[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) cross[row * P + j] = s2 * (kpend[row * 64 + j] - v);
      }
    }
  }
}

// Scratch images of the pending rows (once per bb_posterior call).
int launch_pend_images(const bb_model* m, int32_t layout, const float* d_pend_x, int32_t P, cudaStream_t stream) {
  const bool bits = layout == BB_BITS_U8;
  k_build_pend_img<<<64, 256, 0, stream>>>(d_pend_x, P, m->d, m->d_wide, m->d_cand_scale, m->d_cand_shift, bits ? 1 : 0,
                                           bits ? m->dist_scale_wp : m->dist_scale_p, m->task_col, m->n_tasks,
                                           reinterpret_cast<uint8_t*>(m->d_pend_img), m->d_pend_norm, m->d_pend_task);
  BB_LAUNCH_CHECK();
  return BB_OK;
}

// One block of candidates: k(x*, pending) through k_kmat_tc, then the cross-covariances from the K* block
// that launch_kmat_wide left in the workspace.
int launch_cross_wide(const bb_model* m, const void* d_x, int32_t layout, int64_t nb, int64_t ldx,
                      const float* d_pend_beta, int32_t P, float* d_cross_blk, cudaStream_t stream) {
  const bool bits = layout == BB_BITS_U8;
  WideColumns c;
  c.wimg = reinterpret_cast<const uint8_t*>(m->d_pend_img);
  c.wnorm = m->d_pend_norm;
  c.task = m->d_pend_task;
  c.n = P;
  c.n_pad = 64;
  c.inv_scale = bits ? 1.0f / m->dist_scale_wp : 1.0f / (m->dist_scale_a * m->dist_scale_p);
  const int64_t rows_pad = (nb + 255) / 256 * 256;
  int rc = launch_kmat_cols(m, c, d_x, layout, nb, ldx, m->d_kpend_ws, 64, rows_pad, 64, stream);
  if (rc != BB_OK) return rc;
  const size_t smem = (size_t)P * m->n_pad * sizeof(float);
  BB_CUDA(cudaFuncSetAttribute(k_cross_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_cross_pre<<<148 * 4, 256, smem, stream>>>(m->d_wide_ws, m->d_kpend_ws, d_pend_beta, P, m->n_pad, nb,
                                              m->y_std * m->y_std, d_cross_blk);
  BB_LAUNCH_CHECK();
  return BB_OK;
}

}  // namespace bb
