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
__device__ __forceinline__ float exact_step(float acc, float a, float b) {
  if (METRIC == kMetricL2) {
    const float t = __fsub_rn(a, b);
    return __fadd_rn(acc, __fmul_rn(t, t));
  }
  return __fadd_rn(acc, __fmul_rn(a, b));
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

}  // namespace gfi
