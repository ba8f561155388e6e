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

constexpr int kSelCap = 2048;     // valid candidate keys kept in shared memory (else: passes over global memory)
constexpr int kTileFloats = 4096;  // rerank tile: GC candidates x CW floats, GC * CW = 4096

// Exact 64-bit radix select: returns the `need`-th smallest (1-based) of the valid keys in src[0..cnt).
__device__ __forceinline__ uint64_t radix_select(const uint64_t* src, uint32_t cnt, uint32_t need, uint64_t bound,
                                                 uint32_t* hist, uint32_t* s_bucket, uint32_t* s_need, int tid) {
  uint64_t prefix = 0, mask = 0;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += kSelThreads) {
      const uint64_t key = src[i];
      if (key <= bound && (key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
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
          if (need <= before + h[b]) { *s_bucket = tid * 8 + b; *s_need = need - before; break; }
          before += h[b];
        }
      }
    }
    __syncthreads();
    prefix |= (uint64_t)(*s_bucket) << shift;
    mask |= 0xffull << shift;
    need = *s_need;
    __syncthreads();
  }
  return prefix;
}

template <int METRIC>
__global__ void __launch_bounds__(kSelThreads) select_rerank_kernel(const SelectParams p) {
  // keys[] (compacted valid candidates) is dead once the selection is made; the rerank tiles reuse it
  __shared__ __align__(16) unsigned char s_union[2 * (kTileFloats + 256) * 4];
  __shared__ uint64_t sel[kMaxKP];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bslot[256];
  __shared__ uint32_t s_count, s_nvalid, s_bucket, s_need, s_bound_hi, s_bound_lo;
  static_assert(sizeof(s_union) >= kSelCap * 8, "key buffer must fit the union");
  uint64_t* keys = reinterpret_cast<uint64_t*>(s_union);
  float* tiles = reinterpret_cast<float*>(s_union);
  const int tid = threadIdx.x, lane = tid & 31;
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

    // ---- one pass over the candidate buffer: compact the useful keys into shared memory ----
    // Scan-path input is `cnt / list_len` ascending lists: the smallest of the lists' last keys is an
    // upper bound on the K-th best key overall, so only keys <= that bound can matter.
    if (tid == 0) { s_count = 0; s_nvalid = 0; s_bound_hi = 0xffffffffu; s_bound_lo = 0xffffffffu; }
    __syncthreads();
    uint64_t bound = kKeySentinel - 1;
    if (p.list_len > 0 && p.KP <= p.list_len) {
      uint64_t mine = kKeySentinel;
      for (uint32_t l = tid; l * (uint32_t)p.list_len < cnt; l += kSelThreads)
        mine = min(mine, cand[(size_t)l * p.list_len + p.list_len - 1]);
      for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
      if (lane == 0) atomicMin(&s_bound_hi, (uint32_t)(mine >> 32));
      __syncthreads();
      if (lane == 0 && (uint32_t)(mine >> 32) == s_bound_hi) atomicMin(&s_bound_lo, (uint32_t)mine);
      __syncthreads();
      bound = ((uint64_t)s_bound_hi << 32) | s_bound_lo;
      if (bound == kKeySentinel) bound = kKeySentinel - 1;
    }
    for (uint32_t i0 = 0; i0 < cnt; i0 += kSelThreads * 4) {
      uint64_t kk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // loads first: four round trips overlap
        const uint32_t i = i0 + u * kSelThreads + tid;
        kk[u] = i < cnt ? cand[i] : kKeySentinel;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool valid = kk[u] <= bound;  // the sentinel is above every bound
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (m) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(&s_nvalid, (uint32_t)__popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
          if (valid && pos < (uint32_t)kSelCap) keys[pos] = kk[u];
        }
      }
    }
    __syncthreads();
    const uint32_t nvalid = s_nvalid;
    const uint32_t kpeff = min((uint32_t)KP, nvalid);
    const bool in_smem = nvalid <= (uint32_t)kSelCap;
    const uint64_t* src = in_smem ? keys : cand;
    const uint32_t src_n = in_smem ? nvalid : cnt;

    // ---- exact radix select of the KP-th smallest key (only when there are more than KP) ----
    uint64_t pivot = kKeySentinel - 1;  // everything valid is <= pivot
    if (nvalid > (uint32_t)KP) pivot = radix_select(src, src_n, (uint32_t)KP, bound, hist, &s_bucket, &s_need, tid);
    for (int i = tid; i < kMaxKP; i += kSelThreads) sel[i] = kKeySentinel;
    __syncthreads();
    for (uint32_t i = tid; i < src_n; i += kSelThreads) {
      const uint64_t key = src[i];
      if (key <= bound && key <= pivot) {
        const uint32_t pos = atomicAdd(&s_count, 1u);
        if (pos < (uint32_t)kMaxKP) sel[pos] = key;
      }
    }
    __syncthreads();  // keys[] is dead from here on

    // ---- reference-exact rerank: candidate rows staged through shared-memory tiles with coalesced
    //      loads, one thread per candidate walks its row in order (sequential f32 chain) ----
    const float qn = p.qnorm[qg];
    const float* qv = p.q32 + (size_t)qg * iv.dpad;
    if (METRIC == kMetricCos && nvalid > 0 && qn == 0.f && tid == 0) atomicOr(p.flags, kFlagZeroNorm);
    const int GC = kpeff <= 64 ? 64 : (kpeff <= 128 ? 128 : 256);
    const int CW = kTileFloats / GC;        // floats per row chunk
    const int cw4 = CW >> 2;
    const int tstride = CW + 1;             // odd stride: conflict-free column walks
    const int nchunk = (iv.d + CW - 1) / CW;
    for (uint32_t b0 = 0; b0 < kpeff; b0 += GC) {
      const int nb = (int)min((uint32_t)GC, kpeff - b0);
      __syncthreads();
      if (tid < GC) bslot[tid] = tid < nb ? (uint32_t)(sel[b0 + tid] & 0xffffffffu) : 0xffffffffu;
      __syncthreads();
      float acc = -0.0f;
      float4 stage[4];
      auto load_chunk = [&](int c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = tid + kSelThreads * i;
          const int row = idx / cw4, c4 = idx - row * cw4;
          const int col = c * CW + 4 * c4;
          const uint32_t slot = bslot[row];
          stage[i] = (slot != 0xffffffffu && col < iv.dpad)
                         ? __ldg(reinterpret_cast<const float4*>(iv.x32 + (size_t)slot * iv.dpad + col))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto store_chunk = [&](float* tile) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = tid + kSelThreads * i;
          const int row = idx / cw4, c4 = idx - row * cw4;
          float* dst = tile + row * tstride + 4 * c4;
          dst[0] = stage[i].x; dst[1] = stage[i].y; dst[2] = stage[i].z; dst[3] = stage[i].w;
        }
      };
      load_chunk(0);
      store_chunk(tiles);
      __syncthreads();
      for (int c = 0; c < nchunk; ++c) {
        float* cur = tiles + (c & 1) * (kTileFloats + 256);
        float* nxt = tiles + ((c + 1) & 1) * (kTileFloats + 256);
        if (c + 1 < nchunk) load_chunk(c + 1);  // global loads in flight during the chain below
        if (tid < nb) {
          const float* row = cur + tid * tstride;
          const int c0 = c * CW;
          const int lim = min(CW, iv.d - c0);
#pragma unroll 8
          for (int i = 0; i < lim; ++i) acc = exact_step<METRIC>(acc, __ldg(qv + c0 + i), row[i]);
        }
        if (c + 1 < nchunk) store_chunk(nxt);
        __syncthreads();
      }
      if (tid < nb) {
        const uint32_t slot = bslot[tid];
        float dist;
        if (METRIC == kMetricL2) {
          dist = __fsqrt_rn(acc);
        } else if (METRIC == kMetricDot) {
          dist = -acc;
        } else {
          const float xn = iv.norm[slot];
          if (xn == 0.f || qn == 0.f) {
            atomicOr(p.flags, kFlagZeroNorm);
            dist = 0.f;
          } else {
            float sim = __fdiv_rn(acc, __fmul_rn(qn, xn));
            if (sim < -1.0f) sim = -1.0f;
            else if (sim > 1.0f) sim = 1.0f;
            dist = __fsub_rn(1.0f, sim);
          }
        }
        if (dist != dist) {
          atomicOr(p.flags, kFlagNaN);
          dist = 0.f;
        }
        sel[b0 + tid] = pack_key(dist, slot);
      }
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
