// select_rerank.cu -- K3 `select_rerank` and K5 `merge_topk`.
//
// K3 finishes a query: from the candidate keys produced by the scan kernel (K1) or the
// tcgen05 kernel (K2) it selects the KP best by approximate score (exact radix select),
// re-scores those KP rows with the REFERENCE'S EXACT ARITHMETIC -- sequential f32 sums with
// separately rounded multiply and add, src/distance.rs:37-73 -- sorts by (distance, slot)
// (slot order == internal id order), and emits the first k: the `sort_by` + `truncate(k)` of
// src/flat_index.rs:62-63.  Distances that leave this kernel are therefore bit-identical to
// the reference's.  On the tensor path it also *certifies* the result: using a rigorous bound
// on the fp16 dot-product error it proves that no row outside the reranked set can belong to
// the top-k; queries that cannot be certified are appended to a fallback list and re-run on
// the exact scan kernel.
#include "common.cuh"
#include "kernels.h"

namespace gfi {

namespace {

constexpr int kSelThreads = 256;
constexpr int kMaxKP = 1024;

__device__ __forceinline__ void block_bitonic_sort(uint64_t* arr, int N, int tid) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < N / 2; t += kSelThreads) {
        const int i = ((t / j) * 2 * j) + (t % j);
        const int l = i + j;
        const bool up = ((i & k) == 0);
        const uint64_t a = arr[i], b = arr[l];
        if ((a > b) == up) {
          arr[i] = b;
          arr[l] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Reference-exact distance of query q (d floats) to row x; *score_out = pre-sqrt / raw value.
template <int METRIC>
__device__ __forceinline__ float exact_step(float acc, float a, float b) {
  if (METRIC == kMetricL2) {
    const float t = __fsub_rn(a, b);
    return __fadd_rn(acc, __fmul_rn(t, t));
  }
  return __fadd_rn(acc, __fmul_rn(a, b));
}

template <int METRIC>
__device__ __forceinline__ float exact_distance(const float* __restrict__ q, const float* __restrict__ x,
                                                int d, float qnorm, float xnorm) {
  // Exactly d terms, in order (the zero padding beyond d is never touched).
  float acc = -0.0f;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int nv = d >> 2;
#pragma unroll 4
  for (int i = 0; i < nv; ++i) {
    const float4 a = __ldg(q4 + i);
    const float4 b = __ldg(x4 + i);
    acc = exact_step<METRIC>(acc, a.x, b.x);
    acc = exact_step<METRIC>(acc, a.y, b.y);
    acc = exact_step<METRIC>(acc, a.z, b.z);
    acc = exact_step<METRIC>(acc, a.w, b.w);
  }
  for (int i = nv * 4; i < d; ++i) acc = exact_step<METRIC>(acc, __ldg(q + i), __ldg(x + i));
  if (METRIC == kMetricL2) return __fsqrt_rn(acc);
  if (METRIC == kMetricDot) return -acc;
  float sim = __fdiv_rn(acc, __fmul_rn(qnorm, xnorm));
  if (sim < -1.0f) sim = -1.0f;
  else if (sim > 1.0f) sim = 1.0f;
  return __fsub_rn(1.0f, sim);
}

template <int METRIC>
__global__ void __launch_bounds__(kSelThreads) select_rerank_kernel(const SelectParams p) {
  __shared__ uint64_t sel[kMaxKP];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_count, s_nvalid, s_bucket, s_need;
  const int tid = threadIdx.x;
  const IndexView& iv = p.iv;
  const int nq = p.nq_dev ? (int)*p.nq_dev : p.nq;

  for (int qq = blockIdx.x; qq < nq; qq += gridDim.x) {
    const uint32_t qg = p.qlist ? p.qlist[qq] : (uint32_t)qq;
    const uint64_t* cand = p.cand + (size_t)qg * p.cand_stride;
    const uint32_t cnt_raw = p.cand_cnt[qg];
    const bool overflow = cnt_raw > (uint32_t)p.cand_stride;
    const uint32_t cnt = overflow ? (uint32_t)p.cand_stride : cnt_raw;
    const uint32_t k = p.ks[qg];
    const int KP = p.KP;

    // ---- count valid keys ----
    if (tid == 0) { s_count = 0; s_nvalid = 0; }
    __syncthreads();
    uint32_t local = 0;
    for (uint32_t i = tid; i < cnt; i += kSelThreads) local += (cand[i] != kKeySentinel);
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0 && local) atomicAdd(&s_nvalid, local);
    __syncthreads();
    const uint32_t nvalid = s_nvalid;
    const uint32_t kpeff = min((uint32_t)KP, nvalid);

    // ---- exact radix select of the kpeff-th smallest key (only when nvalid > KP) ----
    uint64_t pivot = kKeySentinel - 1;  // everything valid is <= pivot
    if (nvalid > (uint32_t)KP) {
      uint64_t prefix = 0, mask = 0;
      uint32_t need = KP;
      for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < cnt; i += kSelThreads) {
          const uint64_t key = cand[i];
          if (key != kKeySentinel && (key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid < 32) {
          // each lane owns 8 consecutive buckets
          uint32_t h[8], sum = 0;
#pragma unroll
          for (int b = 0; b < 8; ++b) { h[b] = hist[tid * 8 + b]; sum += h[b]; }
          uint32_t incl = sum;
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
          }
          uint32_t before = incl - sum;
          if (before < need && need <= incl) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
              if (need <= before + h[b]) { s_bucket = tid * 8 + b; s_need = need - before; break; }
              before += h[b];
            }
          }
        }
        __syncthreads();
        prefix |= (uint64_t)s_bucket << shift;
        mask |= 0xffull << shift;
        need = s_need;
        __syncthreads();
      }
      pivot = prefix;
    }

    // ---- gather the selected keys ----
    for (int i = tid; i < kMaxKP; i += kSelThreads) sel[i] = kKeySentinel;
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += kSelThreads) {
      const uint64_t key = cand[i];
      if (key != kKeySentinel && key <= pivot) {
        const uint32_t pos = atomicAdd(&s_count, 1u);
        if (pos < (uint32_t)kMaxKP) sel[pos] = key;
      }
    }
    __syncthreads();

    // ---- reference-exact rerank ----
    const float qn = p.qnorm[qg];
    const float* qv = p.q32 + (size_t)qg * iv.dpad;
    if (METRIC == kMetricCos && nvalid > 0 && qn == 0.f && tid == 0) atomicOr(p.flags, kFlagZeroNorm);
    for (uint32_t t = tid; t < kpeff; t += kSelThreads) {
      const uint32_t slot = (uint32_t)(sel[t] & 0xffffffffu);
      const float xn = (METRIC == kMetricCos) ? iv.norm[slot] : 1.f;
      float dist;
      if (METRIC == kMetricCos && (xn == 0.f || qn == 0.f)) {
        atomicOr(p.flags, kFlagZeroNorm);
        dist = 0.f;
      } else {
        dist = exact_distance<METRIC>(qv, iv.x32 + (size_t)slot * iv.dpad, iv.d, qn, xn);
        if (dist != dist) {
          atomicOr(p.flags, kFlagNaN);
          dist = 0.f;
        }
      }
      sel[t] = pack_key(dist, slot);
    }
    __syncthreads();
    int N = 32;
    while (N < (int)kpeff) N <<= 1;
    block_bitonic_sort(sel, N, tid);

    // ---- emit ----
    const uint32_t kq = min(k, kpeff);
    for (uint32_t t = tid; t < kq; t += kSelThreads) {
      const uint64_t key = sel[t];
      const uint32_t slot = (uint32_t)(key & 0xffffffffu);
      p.out_ids[(size_t)qg * p.kstride + t] = iv.ids[slot];
      p.out_dist[(size_t)qg * p.kstride + t] = key_f32((uint32_t)(key >> 32));
    }
    if (tid == 0) p.out_counts[qg] = kq;

    // ---- certification (tensor path) ----
    if (p.certify && tid == 0 && k > 0) {
      bool ok = !overflow && kpeff >= k;
      if (ok) {
        const float tau = key_f32((uint32_t)(sel[k - 1] >> 32));  // exact k-th distance
        // every row that was not reranked has approximate score >= a_s
        const float a_s = (nvalid > (uint32_t)KP) ? key_f32((uint32_t)(pivot >> 32)) : p.thresh[qg];
        const float dd = (float)iv.d;
        const float gamma = (dd + 8.f) * 5.9604645e-08f;  // (d+8) * 2^-24: sequential-sum rounding
        const float qmax = *p.qmaxabs;
        const float eta_q = 3.7252903e-09f * qmax;         // 2^-28 * max|q|: flushed fp16 query elements
        const float e_dot = (p.eps_rel * qn + eta_q * sqrtf(dd)) * p.xnorm_max;
        float lb;  // lower bound on the reference-arithmetic distance of any non-reranked row
        if (METRIC == kMetricDot) {
          lb = a_s - e_dot - gamma * qn * p.xnorm_max;
        } else if (METRIC == kMetricCos) {
          const float e_s = p.eps_rel * qn + eta_q * sqrtf(dd) + 9.5367432e-07f * qn;
          lb = 1.0f + (a_s - e_s) / qn - 3.f * gamma - 9.5367432e-07f;
        } else {
          const float qs = p.qsumsq[qg];
          const float xs = p.xnorm_max * p.xnorm_max;
          float d2 = a_s + qs - 2.f * e_dot - 4.7683716e-07f * (xs + qs);
          d2 = fmaxf(d2, 0.f);
          lb = sqrtf(d2) * (1.f - gamma) - 1e-30f;
        }
        ok = lb > tau;  // strict: a tie could hide a row with a lower id
        if (!(a_s == a_s)) ok = false;
      }
      if (!ok) {
        const uint32_t pos = atomicAdd(p.fb_count, 1u);
        p.fb_list[pos] = qg;
        if (p.uncertified) atomicAdd(p.uncertified, 1u);
      }
    }
    __syncthreads();
  }
}

// K5: one warp per query merges G sorted lists (lane g owns list g; G <= 32).
__global__ void merge_topk_kernel(const uint64_t* ids, const float* dist, const uint32_t* counts, int G,
                                  int64_t q, int64_t kstride, const uint32_t* ks, uint64_t* out_ids,
                                  float* out_dist, uint32_t* out_counts, int64_t out_kstride) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= q) return;
  const uint32_t k = ks[w];
  uint32_t len = 0, pos = 0;
  const uint64_t* my_ids = nullptr;
  const float* my_dist = nullptr;
  if (lane < G) {
    len = counts[(size_t)lane * q + w];
    my_ids = ids + ((size_t)lane * q + w) * kstride;
    my_dist = dist + ((size_t)lane * q + w) * kstride;
  }
  uint32_t produced = 0;
  while (produced < k) {
    uint32_t dk = 0xffffffffu;
    uint64_t id = ~0ull;
    if (pos < len) { dk = f32_key(my_dist[pos]); id = my_ids[pos]; }
    const unsigned have = __ballot_sync(0xffffffffu, pos < len);
    if (!have) break;
    uint32_t mk = dk;
    for (int o = 16; o > 0; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
    uint64_t cid = (pos < len && dk == mk) ? id : ~0ull;
    uint64_t mid = cid;
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, mid, o);
      mid = other < mid ? other : mid;
    }
    const unsigned win = __ballot_sync(0xffffffffu, pos < len && dk == mk && id == mid);
    const int wl = __ffs(win) - 1;
    if (lane == wl) {
      out_ids[w * out_kstride + produced] = id;
      out_dist[w * out_kstride + produced] = my_dist[pos];
      ++pos;
    }
    ++produced;
  }
  if (lane == 0) out_counts[w] = produced;
}

}  // namespace

cudaError_t launch_select_rerank(const SelectParams& p, int grid, cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  switch (p.iv.metric) {
    case kMetricL2: select_rerank_kernel<kMetricL2><<<grid, kSelThreads, 0, st>>>(p); break;
    case kMetricCos: select_rerank_kernel<kMetricCos><<<grid, kSelThreads, 0, st>>>(p); break;
    case kMetricDot: select_rerank_kernel<kMetricDot><<<grid, kSelThreads, 0, st>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_merge(const uint64_t* ids, const float* dist, const uint32_t* counts, int G, int64_t q,
                         int64_t kstride, const uint32_t* ks, uint64_t* out_ids, float* out_dist,
                         uint32_t* out_counts, int64_t out_kstride, cudaStream_t st) {
  if (q <= 0) return cudaSuccess;
  if (G > 32) return cudaErrorInvalidValue;
  const int64_t blocks = (q * 32 + 255) / 256;
  merge_topk_kernel<<<(unsigned)blocks, 256, 0, st>>>(ids, dist, counts, G, q, kstride, ks, out_ids, out_dist,
                                                      out_counts, out_kstride);
  return cudaGetLastError();
}

}  // namespace gfi
