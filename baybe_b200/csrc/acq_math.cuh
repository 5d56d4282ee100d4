// acq_math.cuh -- acquisition-function arithmetic shared by the fused kernel and the
// stand-alone scoring kernels.  Restates botorch's q=1 Monte-Carlo and analytic acquisition
// functions (classes selected by name at /root/reference/baybe/acquisition/base.py:162-181,
// arguments assembled at baybe/acquisition/_builder.py:195-265); objective o = a*y + b
// (baybe/objectives/single.py:66-91).
//
// qLogEI identity used here: logmeanexp_s(log fatplus(x_s)) == log(mean_s fatplus(x_s)), so the
// per-sample log/exp pair of the reference formulation is replaced by one sum of
// fatplus(x)/tau = softplus(t) + 0.1/(1+t^2), t = x/tau, followed by a single log.
#pragma once

#include "common.cuh"

namespace bb {

// softplus(t) - max(t,0) = log1p(exp(-|t|)); only matters for |t| < 30 in fp32 sums.
__device__ __forceinline__ float softplus_tail(float abs_t) {
  float e = fast_ex2(-kLog2e * abs_t);
  return log1pf(e);
}

// One Monte-Carlo sample contributes to two running sums; t = c0 + c1 * z_s with
//   qLogEI: t = (o_s - best_f)/tau_relu ; s0 += softplus(t) ; s1 += 1/(1+t^2)
//   qEI   : t = o_s - best_f            ; s0 += relu(t)
//   qPI   : t = (o_s - best_f)/tau_pi   ; s0 += sigmoid(t)
//   qUCB, qSR: closed form from the sample statistics mean(z), mean(|z|); no per-sample work.
__device__ __forceinline__ void mc_coef(const bb_acq_spec& a, float mu, float var, float& c0,
                                        float& c1) {
  const float mo = fmaf(a.obj_scale, mu, a.obj_shift);
  const float so = a.obj_scale * sqrtf(var);  // signed: o_s = mo + so * z_s
  float inv_tau = 1.0f;
  if (a.kind == BB_ACQ_QLOGEI) inv_tau = 1.0f / a.tau_relu;
  if (a.kind == BB_ACQ_QPI) inv_tau = 1.0f / a.tau_pi;
  c0 = (mo - a.best_f) * inv_tau;
  c1 = so * inv_tau;
}

template <int KIND>
__device__ __forceinline__ void mc_term(float t, float& s0, float& s1) {
  if constexpr (KIND == BB_ACQ_QLOGEI) {
    s0 += fmaxf(t, 0.f);
    s1 += fast_rcp(fmaf(t, t, 1.0f));
    if (fabsf(t) < 30.f) s0 += softplus_tail(fabsf(t));
  } else if constexpr (KIND == BB_ACQ_QEI) {
    s0 += fmaxf(t, 0.f);
  } else if constexpr (KIND == BB_ACQ_QPI) {
    s0 += fast_rcp(1.0f + fast_ex2(-kLog2e * t));
  }
}

template <int KIND>
__device__ __forceinline__ void mc_partial_kind(float c0, float c1, const float4* __restrict__ z4,
                                                int n4, float& s0, float& s1) {
  if constexpr (KIND == BB_ACQ_QLOGEI) {
    // fast path for all four samples, one rarely-taken correction branch per group of four
    float s0b = 0.f, s1b = 0.f;
#pragma unroll 2
    for (int i = 0; i < n4; ++i) {
      const float4 z = z4[i];
      const float t0 = fmaf(c1, z.x, c0), t1 = fmaf(c1, z.y, c0);
      const float t2 = fmaf(c1, z.z, c0), t3 = fmaf(c1, z.w, c0);
      s0 += fmaxf(t0, 0.f);
      s0b += fmaxf(t1, 0.f);
      s0 += fmaxf(t2, 0.f);
      s0b += fmaxf(t3, 0.f);
      s1 += fast_rcp(fmaf(t0, t0, 1.0f));
      s1b += fast_rcp(fmaf(t1, t1, 1.0f));
      s1 += fast_rcp(fmaf(t2, t2, 1.0f));
      s1b += fast_rcp(fmaf(t3, t3, 1.0f));
      const float tmin = fminf(fminf(fabsf(t0), fabsf(t1)), fminf(fabsf(t2), fabsf(t3)));
      if (tmin < 30.f) {
        if (fabsf(t0) < 30.f) s0 += softplus_tail(fabsf(t0));
        if (fabsf(t1) < 30.f) s0 += softplus_tail(fabsf(t1));
        if (fabsf(t2) < 30.f) s0 += softplus_tail(fabsf(t2));
        if (fabsf(t3) < 30.f) s0 += softplus_tail(fabsf(t3));
      }
    }
    s0 += s0b;
    s1 += s1b;
  } else {
    for (int i = 0; i < n4; ++i) {
      const float4 z = z4[i];
      mc_term<KIND>(fmaf(c1, z.x, c0), s0, s1);
      mc_term<KIND>(fmaf(c1, z.y, c0), s0, s1);
      mc_term<KIND>(fmaf(c1, z.z, c0), s0, s1);
      mc_term<KIND>(fmaf(c1, z.w, c0), s0, s1);
    }
  }
}

// Partial sums over part `part` of `nparts` equal shares of the S base samples in shared memory
// (S / nparts must be a multiple of 4).
__device__ __forceinline__ void mc_partial(const bb_acq_spec& a, float mu, float var,
                                           const float* __restrict__ z_s, int S, int part,
                                           int nparts, float& s0, float& s1) {
  s0 = 0.f;
  s1 = 0.f;
  float c0, c1;
  mc_coef(a, mu, var, c0, c1);
  const int per = S / nparts;
  const float4* z4 = reinterpret_cast<const float4*>(z_s + part * per);
  const int n4 = per >> 2;
  if (a.kind == BB_ACQ_QLOGEI) mc_partial_kind<BB_ACQ_QLOGEI>(c0, c1, z4, n4, s0, s1);
  else if (a.kind == BB_ACQ_QEI) mc_partial_kind<BB_ACQ_QEI>(c0, c1, z4, n4, s0, s1);
  else if (a.kind == BB_ACQ_QPI) mc_partial_kind<BB_ACQ_QPI>(c0, c1, z4, n4, s0, s1);
}

// Accumulate the samples z4[0..n4) (4 per element) into (s0, s1) for pre-computed coefficients;
// lets a caller spread one candidate's Monte-Carlo work over several program phases.
__device__ __forceinline__ void mc_accumulate(int kind, float c0, float c1, const float4* __restrict__ z4,
                                              int n4, float& s0, float& s1) {
  if (kind == BB_ACQ_QLOGEI) mc_partial_kind<BB_ACQ_QLOGEI>(c0, c1, z4, n4, s0, s1);
  else if (kind == BB_ACQ_QEI) mc_partial_kind<BB_ACQ_QEI>(c0, c1, z4, n4, s0, s1);
  else if (kind == BB_ACQ_QPI) mc_partial_kind<BB_ACQ_QPI>(c0, c1, z4, n4, s0, s1);
}

__device__ __forceinline__ float mc_finalize(const bb_acq_spec& a, float mu, float var, float s0,
                                             float s1, int S, float z_mean, float zabs_mean) {
  const float sd = sqrtf(var);
  const float mo = fmaf(a.obj_scale, mu, a.obj_shift);
  const float so = a.obj_scale * sd;
  switch (a.kind) {
    case BB_ACQ_QLOGEI:
      return logf(a.tau_relu) + logf((s0 + 0.1f * s1) / (float)S);
    case BB_ACQ_QEI:
    case BB_ACQ_QPI:
      return s0 / (float)S;
    case BB_ACQ_QSR:
      return fmaf(so, z_mean, mo);
    default:  // qUCB: mean_s( mo + sqrt(beta*pi/2) |o_s - mo| )
      return fmaf(sqrtf(a.beta * 1.5707963267948966f) * fabsf(so), zabs_mean, mo);
  }
}

// Analytic acquisition functions with an affine posterior transform (_builder.py:224-236).
// Evaluated in float64: one evaluation per candidate, far off the critical path.
__device__ __forceinline__ double norm_pdf(double u) {
  return exp(-0.5 * u * u) * 0.3989422804014327;
}
__device__ __forceinline__ double norm_cdf(double u) { return 0.5 * erfc(-u * 0.7071067811865476); }
__device__ __forceinline__ double log1mexp_d(double x) {
  return (x > -0.6931471805599453) ? log(-expm1(x)) : log1p(-exp(x));
}
__device__ __forceinline__ double log_ei_helper(double u) {
  // botorch _log_ei_helper: log(phi(u) + u Phi(u)), stable in the left tail
  if (u > -1.0) return log(norm_pdf(u) + u * norm_cdf(u));
  const double log_phi = -0.5 * (u * u + 1.8378770664093453);
  if (u > -1e6) {
    double w = log(erfcx(-u * 0.7071067811865476) * fabs(u)) + 0.22579135264472744;
    return log_phi + log1mexp_d(w);
  }
  return log_phi - 2.0 * log(fabs(u));
}

__device__ __forceinline__ float analytic_value(const bb_acq_spec& a, float mu, float var) {
  const double m = (double)a.obj_scale * (double)mu + (double)a.obj_shift;
  const double s = fabs((double)a.obj_scale) * sqrt((double)var);
  switch (a.kind) {
    case BB_ACQ_PM:
      return (float)m;
    case BB_ACQ_PSTD:
      return (float)(a.maximize ? s : -s);
    case BB_ACQ_UCB:
      return (float)(m + sqrt((double)a.beta) * s);
    default:
      break;
  }
  const double u = (m - (double)a.best_f) / s;
  if (a.kind == BB_ACQ_EI) return (float)(s * (norm_pdf(u) + u * norm_cdf(u)));
  if (a.kind == BB_ACQ_LOGEI) return (float)(log_ei_helper(u) + log(s));
  return (float)norm_cdf(u);  // PI
}

}  // namespace bb
