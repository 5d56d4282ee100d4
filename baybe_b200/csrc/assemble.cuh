// assemble.cuh -- the K(X*, X) assembly core shared by the fused scoring kernel (A operand of the
// tcgen05 GEMM), the stand-alone kernel-matrix kernel and the pending cross-covariance kernel.
//
// Restates gpytorch MaternKernel/RBFKernel.forward + Distance._sq_dist (constructed by the
// reference at /root/reference/baybe/kernels/base.py:173-178): scaled squared distance in the
// "GEMM form" |a|^2 + |b|^2 - 2 a.b on mean-centred inputs, clamped at 0, then the family's
// closed form.  Normalisation (botorch Normalize, gaussian_process/core.py:301-305), centring,
// the ARD 1/lengthscale and the family constant (5, 3, 1, log2(e)/2) are folded into one
// per-column scale/shift applied when the candidate tile is staged in shared memory.
#pragma once

#include "common.cuh"

namespace bb {

// Shared-memory resident model data needed by the assembly core.
struct AsmSmem {
  const float4* xt4;     // [n_pad][dq] scaled training rows, pre-multiplied by -2
  const float* tsq;      // [n_pad] squared norms
  const int32_t* ttask;  // [n_pad] task ids of the training rows
  const float* tcov;     // [T*T] task covariance (prior scale folded in)
  float4* a_s;           // [dq][128] scaled candidate tile, transposed (conflict-free LDS.128)
  int32_t* cand_task;    // [128]
  int dq;                // d_pad / 4
  int T;
  bool scaled;           // task covariance / prior scale must be applied
};

// Cooperative staging of one tile of 128 candidates: load (any layout), normalise+scale, store
// transposed.  `nthreads` threads with ids `t` take part; caller synchronises afterwards.
__device__ __forceinline__ void stage_candidates(const void* __restrict__ x, int layout,
                                                 int64_t N, int64_t ldx, int64_t row0, int d,
                                                 int d_pad, int task_col,
                                                 const float* __restrict__ cscale,
                                                 const float* __restrict__ cshift, AsmSmem& sm,
                                                 int t, int nthreads) {
  float* a_flat = reinterpret_cast<float*>(sm.a_s);
  const int total = kTileM * d_pad;
  const bool col_major = (layout == BB_COL_MAJOR_F32 || layout == BB_COL_MAJOR_F64);
  for (int e = t; e < total; e += nthreads) {
    int r, j;
    if (col_major) {
      r = e & (kTileM - 1);
      j = e >> 7;
    } else {
      r = e / d_pad;
      j = e - r * d_pad;
    }
    int64_t row = row0 + r;
    float xv = 0.0f;
    if (j < d && row < N) {
      switch (layout) {
        case BB_ROW_MAJOR_F32: xv = load_x<BB_ROW_MAJOR_F32>(x, row, j, ldx); break;
        case BB_COL_MAJOR_F32: xv = load_x<BB_COL_MAJOR_F32>(x, row, j, ldx); break;
        case BB_ROW_MAJOR_F64: xv = load_x<BB_ROW_MAJOR_F64>(x, row, j, ldx); break;
        default: xv = load_x<BB_COL_MAJOR_F64>(x, row, j, ldx); break;
      }
    }
    float a = (j < d) ? fmaf(xv, cscale[j], cshift[j]) : 0.0f;
    a_flat[((j >> 2) * kTileM + r) * 4 + (j & 3)] = a;
    if (j == task_col) {
      int ct = __float2int_rn(xv);
      sm.cand_task[r] = min(max(ct, 0), sm.T - 1);
    }
  }
}

__device__ __forceinline__ float cand_sqnorm(const AsmSmem& sm, int m) {
  float s = 0.0f;
  for (int jc = 0; jc < sm.dq; ++jc) {
    float4 a = sm.a_s[jc * kTileM + m];
    s = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, s))));
  }
  return s;
}

// Kernel values of candidates (m0, m1) against the 8 training points i0..i0+7.
template <int FAMILY>
__device__ __forceinline__ void assemble_2x8(const AsmSmem& sm, int m0, int m1, float an0,
                                             float an1, int i0, float (&k0)[8], float (&k1)[8]) {
  float acc0[8], acc1[8];
#pragma unroll
  for (int ii = 0; ii < 8; ++ii) {
    float t = sm.tsq[i0 + ii];
    acc0[ii] = t + an0;
    acc1[ii] = t + an1;
  }
  const float4* xt = sm.xt4 + (size_t)i0 * sm.dq;
#pragma unroll 1
  for (int jc = 0; jc < sm.dq; ++jc) {
    const float4 a0 = sm.a_s[jc * kTileM + m0];
    const float4 a1 = sm.a_s[jc * kTileM + m1];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const float4 b = xt[ii * sm.dq + jc];
      acc0[ii] = fmaf(a0.x, b.x, fmaf(a0.y, b.y, fmaf(a0.z, b.z, fmaf(a0.w, b.w, acc0[ii]))));
      acc1[ii] = fmaf(a1.x, b.x, fmaf(a1.y, b.y, fmaf(a1.z, b.z, fmaf(a1.w, b.w, acc1[ii]))));
    }
  }
#pragma unroll
  for (int ii = 0; ii < 8; ++ii) {
    k0[ii] = kernel_from_t<FAMILY>(acc0[ii]);
    k1[ii] = kernel_from_t<FAMILY>(acc1[ii]);
  }
  if (sm.scaled) {
    const float* r0 = sm.tcov + sm.cand_task[m0] * sm.T;
    const float* r1 = sm.tcov + sm.cand_task[m1] * sm.T;
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      int tt = sm.ttask[i0 + ii];
      k0[ii] *= r0[tt];
      k1[ii] *= r1[tt];
    }
  }
}

}  // namespace bb
