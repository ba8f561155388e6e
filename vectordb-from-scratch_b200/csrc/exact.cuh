// exact.cuh -- the reference's arithmetic, in ONE place: every distance that leaves a search is produced by these
// two functions (K3 rerank and the fused tail of the scan kernel; K6 pair scoring shares exact_step and keeps a
// per-pair status instead of the call-wide flags).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace gfi {

// One step of the reference's sequential sums (src/distance.rs:37-44,67-73): separately rounded
// subtract / multiply / add, never contracted to FMA.
template <int METRIC>
__device__ __forceinline__ float exact_term(float a, float b) {  // the term one step adds: (a-b)^2 or a*b
  if (METRIC == kMetricL2) {
    const float t = __fsub_rn(a, b);
    return __fmul_rn(t, t);
  }
  return __fmul_rn(a, b);
}
template <int METRIC>
__device__ __forceinline__ float exact_step(float acc, float a, float b) {
  return __fadd_rn(acc, exact_term<METRIC>(a, b));
}

// From the finished sum to DistanceMetric::distance's value (src/distance.rs:27-33,47-64): sqrt for L2, negation
// for the dot product, dot / (|q| |x|) clamped to [-1, 1] and subtracted from 1 for cosine.  A zero norm (cosine)
// or a NaN distance raises the call's flag (the reference returns InvalidVector / panics) and yields 0.
template <int METRIC>
__device__ __forceinline__ float exact_finish(float acc, float xn, float qn, uint32_t* flags) {
  float dist;
  if (METRIC == kMetricL2) {
    dist = __fsqrt_rn(acc);
  } else if (METRIC == kMetricDot) {
    dist = -acc;
  } else {
    if (xn == 0.f || qn == 0.f) {
      atomicOr(flags, kFlagZeroNorm);
      dist = 0.f;
    } else {
      float sim = __fdiv_rn(acc, __fmul_rn(qn, xn));
      if (sim < -1.0f) sim = -1.0f;
      else if (sim > 1.0f) sim = 1.0f;
      dist = __fsub_rn(1.0f, sim);
    }
  }
  if (dist != dist) {
    atomicOr(flags, kFlagNaN);
    dist = 0.f;
  }
  return dist;
}

// Certification of the fp32 scan path.  The scan kernel ranks rows by an APPROXIMATE score -- the same f32 products
// as the reference, but summed by FMA chains and a tree instead of the reference's left-to-right chain
// (src/distance.rs:37-44,67-73) -- and only the K best rows by that score are re-scored exactly.  Any d-term f32 sum
// of the same terms, in any order, with or without FMA, is within gamma = (d + 8) * 2^-24 (relative to the sum of the
// terms' magnitudes) of the real-number sum, so the approximate score a and the reference's value r of one row obey
//   L2 (a = sum (q-x)^2, all terms >= 0):   r >= sqrt(a (1 - 2 gamma)) (1 - 2^-24)
//   dot (a = -sum q x):                      r >= a - 2 gamma |q| |x|
//   cosine (a = -sum q x / |x|):             r >= 1 + a / |q| - 2 gamma - O(2^-23)
// Returns a value every row with approximate score >= a_s is guaranteed to reach or exceed (slack included for the
// f32 evaluation of the bound itself).  If it is strictly above the exact k-th distance of the re-scored rows, no
// dropped row can belong to the top-k; otherwise the host proves the query by paging (api.cu prove_query).
__host__ __device__ inline float scan_lower_bound(int metric, float a_s, float qn, float xnorm_max, int d) {
  const float gamma = ((float)d + 8.f) * 5.9604645e-08f;
  if (metric == kMetricL2) {
    const float v = a_s * (1.f - 2.5f * gamma);
    return v > 0.f ? sqrtf(v) * (1.f - 2.3841858e-07f) : 0.f;
  }
  if (metric == kMetricDot)
    return a_s - 2.2f * gamma * qn * xnorm_max - 2.3841858e-07f * (a_s < 0.f ? -a_s : a_s) - 1e-37f;
  return 1.0f + a_s / qn - 2.5f * gamma - 9.5367432e-07f;
}

// Error model of the tensor path (K2 ranks rows by an fp16 dot product accumulated in f32).  With
// |q^.x^ - q.x| <= eps_rel |q||x| (+ the flushed-element term eta_q), a row whose approximate score is `a` has a
// reference-arithmetic distance r inside [lb(a), ub(a)].  `lb` is what the certification of rerank_finalize_kernel
// rests on (every row that was not re-scored has approximate score >= a_s, hence r >= lb(a_s)); `ub` is only used to
// decide how many candidates are worth re-scoring (select_kernel's rerank cut), so it affects how often a query falls
// back, never what is returned.
struct TensorBoundIn {
  float qn, qs;         // |q| and sum q^2 (reference-exact sequential sums)
  float eps_rel, qmax;  // relative fp16 error bound, batch max |q|
  float xnorm_max;      // max |x| over the index
  int d;
};
__device__ __forceinline__ void tensor_bounds(int metric, float a, const TensorBoundIn& t, float* lb, float* ub) {
  const float dd = (float)t.d;
  const float gamma = (dd + 8.f) * 5.9604645e-08f;  // (d+8) * 2^-24: sequential-sum rounding
  const float eta_q = 3.7252903e-09f * t.qmax;      // 2^-28 * max|q|: flushed fp16 query elements
  const float e_dot = (t.eps_rel * t.qn + eta_q * sqrtf(dd)) * t.xnorm_max;
  if (metric == kMetricDot) {
    const float e = e_dot + gamma * t.qn * t.xnorm_max;
    *lb = a - e_dot - gamma * t.qn * t.xnorm_max;
    *ub = a + e;
  } else if (metric == kMetricCos) {
    const float e_s = t.eps_rel * t.qn + eta_q * sqrtf(dd) + 9.5367432e-07f * t.qn;
    *lb = 1.0f + (a - e_s) / t.qn - 3.f * gamma - 9.5367432e-07f;
    *ub = 1.0f + (a + e_s) / t.qn + 3.f * gamma + 9.5367432e-07f;
  } else {
    const float xs = t.xnorm_max * t.xnorm_max;
    // `a` uses the PRECOMPUTED sequential sums sum x^2 and sum q^2, each off by up to gamma relatively
    // (ADVICE r1), plus the epilogue's own fp32 rounding (2^-22 of the magnitudes involved)
    float d2 = a + t.qs - 2.f * e_dot - (gamma + 2.3841858e-07f) * (xs + t.qs);
    d2 = fmaxf(d2, 0.f);
    *lb = sqrtf(d2) * (1.f - gamma) - 1e-30f;
    float u2 = a + t.qs + 2.f * e_dot + (gamma + 2.3841858e-07f) * (xs + t.qs);
    u2 = fmaxf(u2, 0.f);
    *ub = sqrtf(u2) * (1.f + gamma) + 1e-30f;
  }
}

}  // namespace gfi
