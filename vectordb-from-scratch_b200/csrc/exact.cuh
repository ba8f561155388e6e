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

}  // namespace gfi
