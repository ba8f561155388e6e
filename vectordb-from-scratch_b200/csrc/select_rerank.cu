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
#include <algorithm>

#include "common.cuh"
#include "exact.cuh"
#include "kernels.h"

namespace gfi {

namespace {

constexpr int kSelThreads = 256;
constexpr int kMaxKP = 1024;

// (valid candidate keys kept in shared memory: SelectParams::sel_cap, default kSelectStageKeys; else passes over
// global memory)
constexpr int kTileFloats = 4096;  // rerank tile: GC candidates x CW floats, GC * CW = 4096

// Exact 64-bit radix select: returns the `need`-th smallest (1-based) of the valid keys in src[0..cnt).
// After the four passes over the score half of the keys the remaining group usually holds ONE key (scores tie only
// among duplicate rows): that key is the answer, and the four passes over the slot half are skipped.
__device__ __forceinline__ uint64_t radix_select(const uint64_t* src, uint32_t cnt, uint32_t need, uint64_t bound,
                                                 uint32_t* hist, uint32_t* s_bucket, uint32_t* s_need, int tid) {
  __shared__ uint32_t s_group;
  __shared__ unsigned long long s_only;
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
          if (need <= before + h[b]) { *s_bucket = tid * 8 + b; *s_need = need - before; s_group = h[b]; break; }
          before += h[b];
        }
      }
    }
    __syncthreads();
    prefix |= (uint64_t)(*s_bucket) << shift;
    mask |= 0xffull << shift;
    need = *s_need;
    const bool single = pass == 3 && s_group == 1;  // block-uniform
    __syncthreads();
    if (single) {
      for (uint32_t i = tid; i < cnt; i += kSelThreads) {
        const uint64_t key = src[i];
        if (key <= bound && (key & mask) == prefix) s_only = key;
      }
      __syncthreads();
      const uint64_t only = s_only;
      __syncthreads();
      return only;
    }
  }
  return prefix;
}

// ---- K3a: per query, pick the KP best candidate keys (by approximate score) ----------------------
__global__ void __launch_bounds__(kSelThreads) select_kernel(const SelectParams p) {
  griddep_wait();
  extern __shared__ __align__(16) uint64_t s_dyn[];  // sel[KP] | keys[kSelCap] | (sorted-list path only) keys2[kSelCap]
  const uint32_t kSelCap = p.sel_cap > 0 ? (uint32_t)p.sel_cap : (uint32_t)kSelectStageKeys;
  uint64_t* sel = s_dyn;
  uint64_t* keys = s_dyn + p.KP;
  uint64_t* keys2 = keys + kSelCap;  // prefix keys of the sorted-list path (keys[] holds the heads meanwhile)
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_count, s_nvalid, s_bucket, s_need, s_over;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nq = p.nq_dev ? (int)*p.nq_dev : p.nq;
  auto sel_stage = [&](uint64_t*, uint32_t pos, uint64_t key) { keys2[pos] = key; };

  for (int qq = blockIdx.x; qq < nq; qq += gridDim.x) {
    const uint32_t qg = p.qlist ? p.qlist[qq] : (uint32_t)qq;
    const uint64_t* cand = p.cand + (size_t)qg * p.cand_stride;
    const uint32_t cnt_raw = p.slice_cnt ? (uint32_t)p.cand_stride : p.cand_cnt[qg];
    bool overflow = cnt_raw > (uint32_t)p.cand_stride;
    const uint32_t cnt = overflow ? (uint32_t)p.cand_stride : cnt_raw;
    const int KP = p.KP;

    // ---- compact the useful keys into shared memory ----
    if (tid == 0) { s_count = 0; s_nvalid = 0; s_over = 0; }
    __syncthreads();
    uint64_t bound = kKeySentinel - 1;
    const uint32_t L = p.list_len > 0 ? cnt / (uint32_t)p.list_len : 0;  // number of ascending lists
    const uint32_t m_heads = L ? min((uint32_t)p.list_len, ((uint32_t)KP + L - 1) / L + 1) : 0;
    if (L && L * m_heads <= kSelCap) {
      // Scan-path input: L ascending lists.  (a) The KP-th smallest of the lists' first m_heads keys
      // bounds the KP-th smallest overall from above; (b) the keys <= that bound are a prefix of each
      // list, so one thread per list walks its prefix (batches of 4 independent loads).
      for (uint32_t l = tid; l < L; l += kSelThreads) {
        for (uint32_t j0 = 0; j0 < m_heads; j0 += 4) {
          uint64_t kk[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) kk[u] = (j0 + u < m_heads) ? cand[(size_t)l * p.list_len + j0 + u] : kKeySentinel;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j0 + u < m_heads) keys[l * m_heads + j0 + u] = kk[u];
        }
      }
      __syncthreads();
      const uint32_t nh = L * m_heads;
      uint32_t nh_valid = 0;
      for (uint32_t i = tid; i < nh; i += kSelThreads) nh_valid += keys[i] != kKeySentinel;
      for (int o = 16; o > 0; o >>= 1) nh_valid += __shfl_xor_sync(0xffffffffu, nh_valid, o);
      if (lane == 0 && nh_valid) atomicAdd(&s_count, nh_valid);
      __syncthreads();
      const uint32_t heads_valid = s_count;
      __syncthreads();
      if (tid == 0) s_count = 0;
      if (heads_valid >= (uint32_t)KP)
        bound = radix_select(keys, nh, (uint32_t)KP, kKeySentinel - 1, hist, &s_bucket, &s_need, tid);
      __syncthreads();
      for (uint32_t l = tid; l < L; l += kSelThreads) {
        const uint64_t* lst = cand + (size_t)l * p.list_len;
        for (uint32_t j0 = 0; j0 < (uint32_t)p.list_len; j0 += 4) {
          uint64_t kk[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) kk[u] = (j0 + u < (uint32_t)p.list_len) ? lst[j0 + u] : kKeySentinel;
          bool more = true;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (more && kk[u] <= bound) {
              const uint32_t pos = atomicAdd(&s_nvalid, 1u);
              if (pos < kSelCap) sel_stage(keys, pos, kk[u]);
            } else {
              more = false;
            }
          }
          if (!more) break;
        }
      }
    } else if (p.slice_cnt && p.slice_gather) {
      // Tensor-path input: one thread per slice reads the slice's count, then its (few) valid keys.
      for (int sl = tid; sl < p.nslices; sl += kSelThreads) {
        uint32_t n = p.slice_cnt[(size_t)sl * p.slice_q + qg];
        if (n > p.slice_cap) { s_over = 1; n = p.slice_cap; }
        if (n) {
          const uint64_t* src = cand + (size_t)sl * p.slice_cap;
          const uint32_t base = atomicAdd(&s_nvalid, n);
          for (uint32_t j = 0; j < n; ++j)
            if (base + j < kSelCap) keys[base + j] = src[j];
        }
      }
    } else {
      if (p.slice_cnt)  // sentinel-filled tensor-path input (large KP): only the overflow check uses the counts
        for (int sl = tid; sl < p.nslices; sl += kSelThreads)
          if (p.slice_cnt[(size_t)sl * p.slice_q + qg] > p.slice_cap) s_over = 1;
      for (uint32_t i0 = 0; i0 < cnt; i0 += kSelThreads * 4) {
        uint64_t kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // loads first: four round trips overlap
          const uint32_t i = i0 + u * kSelThreads + tid;
          kk[u] = i < cnt ? cand[i] : kKeySentinel;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool valid = kk[u] != kKeySentinel;
          const unsigned m = __ballot_sync(0xffffffffu, valid);
          if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_nvalid, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (valid && pos < kSelCap) keys[pos] = kk[u];
          }
        }
      }
    }
    __syncthreads();
    const uint32_t nvalid = s_nvalid;
    const uint32_t kpeff = min((uint32_t)KP, nvalid);
    // slices are not contiguous in global memory, so a query whose keys do not fit the staging area is
    // treated as overflowed (it falls back to the exact scan)
    if (p.slice_cnt && (s_over || (p.slice_gather && nvalid > kSelCap))) overflow = true;
    const bool in_smem = nvalid <= kSelCap || (p.slice_cnt && p.slice_gather);
    const bool sorted_path = L && L * m_heads <= kSelCap;
    const uint64_t* src = in_smem ? (sorted_path ? keys2 : keys) : cand;
    const uint32_t src_n = in_smem ? min(nvalid, kSelCap) : cnt;

    // ---- exact radix select of the KP-th smallest key (only when there are more than KP) ----
    uint64_t pivot = kKeySentinel - 1;  // everything valid is <= pivot
    // (nvalid == KP: the pivot is the largest selected key -- the scan-path certification needs it, exact.cuh)
    if (nvalid >= (uint32_t)KP) pivot = radix_select(src, src_n, (uint32_t)KP, bound, hist, &s_bucket, &s_need, tid);
    for (int i = tid; i < KP; i += kSelThreads) sel[i] = kKeySentinel;
    __syncthreads();
    for (uint32_t i = tid; i < src_n; i += kSelThreads) {
      const uint64_t key = src[i];
      if (key <= bound && key <= pivot) {
        const uint32_t pos = atomicAdd(&s_count, 1u);
        if (pos < (uint32_t)KP) sel[pos] = key;
      }
    }
    __syncthreads();
    // ---- rerank cut (tensor path): which of the selected rows are worth re-scoring? ----
    // The k best approximate scores are k real rows, so the exact k-th distance is at most ub(a_k) (exact.cuh
    // tensor_bounds); a candidate with lb(a) > ub(a_k) cannot belong to the top-k.  At the benchmark sizes the fp16
    // error interval is a tenth of the gap between the k-th and the KP-th score: 12-15 of 64 candidates remain, and
    // the rerank -- random 128-byte pieces of HBM rows, its bytes are its time -- shrinks with them.  The keys are put
    // in ascending order by counting ranks (they are unique: each carries its slot), the re-scored set is a prefix,
    // and the FIRST key behind it replaces the pivot in the certification, which still compares against the exact
    // k-th distance: a cut that is too tight makes the query fall back, it cannot change an answer.
    uint32_t kp_out = kpeff, cut_key = 0, has_cut = 0;
    const uint32_t kq = p.ks[qg];
    if (p.certify == 1 && p.rerank_cut && (uint32_t)KP <= kSelCap && kq > 0 && kpeff > kq && !overflow) {
      uint64_t* sorted = keys;  // the staging area is free again
      for (int t = tid; t < KP; t += kSelThreads) {
        const uint64_t key = sel[t];
        if (key == kKeySentinel) continue;
        uint32_t rank = 0;
        for (int j = 0; j < KP; ++j) rank += sel[j] < key;
        sorted[rank] = key;
      }
      // (s_need is free here) first index that need not be re-scored; rerank_cut == 2 is the test mode "cut right
      // behind the k-th key", which the certification must answer by falling back
      if (tid == 0) s_need = p.rerank_cut == 2 ? kq : kpeff;
      __syncthreads();
      TensorBoundIn tb{p.qnorm[qg], p.qsumsq[qg], p.eps_rel, *p.qmaxabs, p.xnorm_max, p.iv.d};
      float lb_k, ub_k;
      tensor_bounds(p.iv.metric, key_f32((uint32_t)(sorted[kq - 1] >> 32)), tb, &lb_k, &ub_k);
      for (uint32_t t = kq + tid; t < kpeff && p.rerank_cut != 2; t += kSelThreads) {
        float lb_t, ub_t;
        tensor_bounds(p.iv.metric, key_f32((uint32_t)(sorted[t] >> 32)), tb, &lb_t, &ub_t);
        if (lb_t > ub_k) { atomicMin(&s_need, t); break; }  // (ascending: the thread's later keys are behind it too)
      }
      __syncthreads();
      kp_out = s_need;
      if (kp_out < kpeff) { has_cut = 1; cut_key = (uint32_t)(sorted[kp_out] >> 32); }
      for (int i = tid; i < KP; i += kSelThreads)
        p.sel_keys[(size_t)qg * p.KP + i] = (uint32_t)i < kpeff ? sorted[i] : kKeySentinel;
    } else {
      for (int i = tid; i < KP; i += kSelThreads) p.sel_keys[(size_t)qg * p.KP + i] = sel[i];
    }
    if (tid == 0) {
      SelInfo info;
      info.pivot = pivot;
      info.nvalid = nvalid;
      info.kpeff = kp_out;
      info.overflow = overflow ? 1u : 0u;
      info.done = 0;
      info.has_cut = has_cut;
      info.cut_key = cut_key;
      p.sel_info[qg] = info;
    }
    __syncthreads();
  }
}

// ---- K3b: reference-exact rerank (bulk: 8 candidates per warp; latency mode: one) + finalize by the last warp ----
constexpr int kRrWarps = 4;                       // warps per block
constexpr int kRrCW = 32;                         // floats per row chunk (one 128-byte line)
constexpr int kRrRPW = 8;                         // bulk mode: candidate rows per warp (4 lanes per row)
constexpr int kRrDepth = 8;                       // bulk mode: row chunks in flight per lane
constexpr int kRrDepthQ = 2;                      // bulk mode: query chunks in flight per lane (kRrDepth % kRrDepthQ == 0)
static_assert(kRrDepth % kRrDepthQ == 0, "query slots are addressed by chunk index modulo kRrDepthQ");
constexpr int kRrTile = 32 * (kRrCW + 1);         // per-warp scratch: the latency mode's row + query, the final sort
static_assert(2 * kRrTile * 4 >= kMaxKP * 8, "the finalizing warp sorts in its tile buffers");

__device__ __forceinline__ void warp_bitonic_sort(uint64_t* arr, int N, int lane) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < N / 2; t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // j is a power of two: (t / j) * 2j + t % j
        const int l = i + j;
        const bool up = ((i & k) == 0);
        const uint64_t a = arr[i], b = arr[l];
        if ((a > b) == up) {
          arr[i] = b;
          arr[l] = a;
        }
      }
      __syncwarp();
    }
  }
}

template <int METRIC>
__global__ void __launch_bounds__(kRrWarps * 32) rerank_finalize_kernel(const SelectParams p) {
  griddep_wait();
  __shared__ __align__(16) float s_tiles[kRrWarps][2 * kRrTile];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const IndexView& iv = p.iv;
  const int nq = p.nq_dev ? (int)*p.nq_dev : p.nq;
  const int rpw = p.warp_per_candidate ? 1 : kRrRPW;  // candidate rows per warp
  const int groups = p.KP / rpw;                       // warps per query
  const int64_t gw = (int64_t)blockIdx.x * kRrWarps + wib;
  const int64_t total = (int64_t)nq * groups;
  float* tile = s_tiles[wib];

  for (int64_t wq = gw; wq < total; wq += (int64_t)gridDim.x * kRrWarps) {
    // group-major: the first nq warps take group 0 of every query, the next nq group 1, ...  Warps that have rows
    // sit next to each other, so whole blocks of the late groups exit at once and the live ones fit one wave (query-
    // major, a block held two live and two idle warps and the kernel ran in two waves at a quarter of the occupancy)
    const int gi = (int)(wq / nq), qq = (int)(wq - (int64_t)gi * nq);
    const uint32_t qg = p.qlist ? p.qlist[qq] : (uint32_t)qq;
    SelInfo* info = p.sel_info + qg;
    const uint32_t kpeff = info->kpeff;
    // warps that have rows to re-score (at least one, which finalizes a query without candidates); the others have
    // nothing to do with this query at all -- with the rerank cut that is most of them
    const int ga = max(1, ((int)kpeff + rpw - 1) / rpw);
    if (gi >= ga) continue;
    uint64_t* skeys = p.sel_keys + (size_t)qg * p.KP;
    const float qn = p.qnorm[qg];
    const float* qv = p.q32 + (size_t)qg * iv.dpad;
    auto finish_distance = [&](float acc, uint32_t slot) -> float {
      return exact_finish<METRIC>(acc, METRIC == kMetricCos ? iv.norm[slot] : 1.f, qn, p.flags);
    };

    if (p.warp_per_candidate) {
      // latency mode (few candidates in total): the warp pulls one whole row and the query into shared
      // memory with coalesced loads, then lane 0 walks them in order
      if ((uint32_t)gi < kpeff) {
        const uint32_t slot = (uint32_t)(skeys[gi] & 0xffffffffu);
        float* xr = tile;
        float* qr = tile + iv.dpad;
        const float4* x4 = reinterpret_cast<const float4*>(iv.x32 + (size_t)slot * iv.dpad);
        const float4* q4 = reinterpret_cast<const float4*>(qv);
        for (int t = lane; t < (iv.dpad >> 2); t += 32) {
          reinterpret_cast<float4*>(xr)[t] = __ldg(x4 + t);
          reinterpret_cast<float4*>(qr)[t] = __ldg(q4 + t);
        }
        __syncwarp();
        if (lane == 0) {
          float acc = -0.0f;
#pragma unroll 8
          for (int i = 0; i < iv.d; ++i) acc = exact_step<METRIC>(acc, qr[i], xr[i]);
          skeys[gi] = pack_key(finish_distance(acc, slot), slot);
        }
        __syncwarp();
      }
    } else if (gi < ga && kpeff > 0) {
      // Bulk mode.  Four lanes per candidate row, eight rows per warp.  A chunk is one 128-byte line of each row:
      // lane (r, c8) loads ITS 8 consecutive floats of row r straight into registers (two 128-bit loads: the four
      // lanes of a row cover the line, no shared-memory transpose) together with the matching 8 query floats, and
      // forms the eight terms q*x (or (q-x)^2) -- separately rounded, independent of each other.  The reference's
      // left-to-right sum (src/distance.rs:37-44,67-73) is the only sequential part: the running sum travels as a
      // token through the four lanes of the row (one shuffle per hand-over), each adding its eight terms in order,
      // and wraps around to the first lane for the next chunk.  kRrDepth chunks of loads are in flight per lane.
      // Round 2's first form staged 32 rows per warp through shared memory and let every lane walk one row: ~250
      // instructions per chunk on ONE warp's critical path; with the rerank cut leaving 10-20 rows per query that
      // single warp set the kernel's time (ncu: 58 us for 48 MB, 9 % of the warp slots occupied).
      const int r = lane >> 2, c8 = lane & 3;
      const int ci = gi * kRrRPW + r;
      const bool have = (uint32_t)ci < kpeff;
      uint32_t slot = have ? (uint32_t)(skeys[ci] & 0xffffffffu) : 0u;
      const uint32_t slot0 = __shfl_sync(0xffffffffu, slot, 0);
      if (!have) slot = slot0;  // absent rows alias the warp's first row: same lines, results never stored
      const float* xrow = iv.x32 + (size_t)slot * iv.dpad;
      const int nchunk = (iv.d + kRrCW - 1) / kRrCW;
      const int col_last = iv.dpad - 4;
      // row pieces come from HBM (random rows: ~2000 cycles under load) and are fetched kRrDepth chunks ahead; the
      // query pieces are shared by the warp's rows and by every warp of the query (L1/L2 hits): kRrDepthQ ahead
      float4 xa[kRrDepth], xb[kRrDepth], qa[kRrDepthQ], qb[kRrDepthQ];
      auto issue_x = [&](int j, int c) {
        const int col = c * kRrCW + 8 * c8;
        xa[j] = __ldg(reinterpret_cast<const float4*>(xrow + min(col, col_last)));
        xb[j] = __ldg(reinterpret_cast<const float4*>(xrow + min(col + 4, col_last)));
      };
      auto issue_q = [&](int j, int c) {
        const int col = c * kRrCW + 8 * c8;
        qa[j] = __ldg(reinterpret_cast<const float4*>(qv + min(col, col_last)));
        qb[j] = __ldg(reinterpret_cast<const float4*>(qv + min(col + 4, col_last)));
      };
#pragma unroll
      for (int j = 0; j < kRrDepth; ++j) {
        xa[j] = xb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < nchunk) issue_x(j, j);
      }
#pragma unroll
      for (int j = 0; j < kRrDepthQ; ++j) {
        qa[j] = qb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < nchunk) issue_q(j, j);
      }
      const int src_lane = (lane & ~3) | ((lane + 3) & 3);  // previous lane of the row (the first takes from the last)
      float acc = -0.0f;
      for (int c0 = 0; c0 < nchunk; c0 += kRrDepth) {
#pragma unroll
        for (int j = 0; j < kRrDepth; ++j) {
          const int c = c0 + j;
          if (c < nchunk) {  // warp-uniform
            const float4 x0 = xa[j], x1 = xb[j], q0 = qa[j % kRrDepthQ], q1 = qb[j % kRrDepthQ];
            if (c + kRrDepth < nchunk) issue_x(j, c + kRrDepth);
            if (c + kRrDepthQ < nchunk) issue_q(j % kRrDepthQ, c + kRrDepthQ);
            float t[8];
            t[0] = exact_term<METRIC>(q0.x, x0.x); t[1] = exact_term<METRIC>(q0.y, x0.y);
            t[2] = exact_term<METRIC>(q0.z, x0.z); t[3] = exact_term<METRIC>(q0.w, x0.w);
            t[4] = exact_term<METRIC>(q1.x, x1.x); t[5] = exact_term<METRIC>(q1.y, x1.y);
            t[6] = exact_term<METRIC>(q1.z, x1.z); t[7] = exact_term<METRIC>(q1.w, x1.w);
            // terms behind position d do not exist (the sum stops at d): they become the neutral element -0.0, which
            // leaves every value it is added to unchanged, bit for bit (including -0.0, infinities and NaN)
            const int nv = iv.d - (c * kRrCW + 8 * c8);
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] = e < nv ? t[e] : -0.0f;
            // branch-free hand-over: every lane adds its terms to the incoming token, only the lane whose turn it is
            // keeps the result (a divergent `if (c8 == ph)` cost a BSSY/BSYNC pair per phase on the critical path)
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
              float v = __shfl_sync(0xffffffffu, acc, src_lane);
#pragma unroll
              for (int e = 0; e < 8; ++e) v = __fadd_rn(v, t[e]);
              acc = c8 == ph ? v : acc;
            }
          }
        }
      }
      // the finished sum sits in the row's last lane (lanes behind position d passed the token on unchanged)
      if (have && c8 == 3) skeys[ci] = pack_key(finish_distance(acc, slot), slot);
    }
    // ---- the last warp of a query to get here sorts, emits and certifies ----
    __threadfence();
    __syncwarp();
    uint32_t prev = 0;
    if (lane == 0) prev = atomicAdd(&info->done, 1u);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev != (uint32_t)ga - 1) continue;
    __threadfence();
    if (METRIC == kMetricCos && info->nvalid > 0 && qn == 0.f && lane == 0) atomicOr(p.flags, kFlagZeroNorm);
    uint64_t* sk = reinterpret_cast<uint64_t*>(tile);
    int N = 32;
    while (N < (int)kpeff) N <<= 1;
    for (int i = lane; i < N; i += 32) sk[i] = (uint32_t)i < kpeff ? __ldcg(skeys + i) : kKeySentinel;
    __syncwarp();
    warp_bitonic_sort(sk, N, lane);
    const uint32_t k = p.ks[qg];
    const uint32_t kq = min(k, kpeff);
    for (uint32_t t = lane; t < kq; t += 32) {
      const uint64_t key = sk[t];
      const uint32_t sl = (uint32_t)(key & 0xffffffffu);
      p.out_ids[(size_t)qg * p.kstride + t] = iv.ids[sl];
      p.out_dist[(size_t)qg * p.kstride + t] = key_f32((uint32_t)(key >> 32));
    }
    if (lane == 0) p.out_counts[qg] = kq;

    // ---- certification (scan path): rows were dropped only if some list was full, i.e. >= KP keys existed; every
    // dropped row then has an approximate key >= the KP-th smallest one overall, the pivot (exact.cuh) ----
    if (p.certify == 2 && lane == 0 && k > 0 && info->nvalid >= (uint32_t)p.KP) {
      bool ok = kq == k;
      if (ok) {
        const float tau = key_f32((uint32_t)(sk[k - 1] >> 32));  // exact k-th distance
        const float a_s = key_f32((uint32_t)(info->pivot >> 32));
        const float lb = scan_lower_bound(METRIC, a_s, qn, p.xnorm_max, iv.d);
        ok = (a_s == a_s) && lb > tau;  // strict: a tie could hide a row with a lower id
      }
      if (!ok) {
        const uint32_t pos = atomicAdd(p.up_count, 1u);
        if (pos < p.up_cap) p.up_list[pos] = qg + p.up_base;
      }
    }
    // ---- certification (tensor path) ----
    if (p.certify == 1 && lane == 0 && k > 0) {
      bool ok = !info->overflow && kpeff >= k;
      if (ok) {
        const float tau = key_f32((uint32_t)(sk[k - 1] >> 32));  // exact k-th distance
        // every row that was not reranked has approximate score >= a_s: the first selected key behind the rerank
        // cut if select_kernel shortened the set, else the selection's pivot, else the tensor pass's threshold
        const float a_s = info->has_cut ? key_f32(info->cut_key)
                          : (info->nvalid > (uint32_t)p.KP) ? key_f32((uint32_t)(info->pivot >> 32)) : p.thresh[qg];
        const TensorBoundIn tb{qn, METRIC == kMetricL2 ? p.qsumsq[qg] : 0.f, p.eps_rel, *p.qmaxabs, p.xnorm_max, iv.d};
        float lb, ub;  // lb: lower bound on the reference-arithmetic distance of any non-reranked row
        tensor_bounds(METRIC, a_s, tb, &lb, &ub);
        ok = lb > tau;  // strict: a tie could hide a row with a lower id
        if (!(a_s == a_s)) ok = false;
      }
      if (!ok) {
        const uint32_t pos = atomicAdd(p.fb_count, 1u);
        p.fb_list[pos] = qg;
        if (p.uncertified) atomicAdd(p.uncertified, 1u);
      }
    }
    __syncwarp();
  }
}

// K5: one warp per query merges G sorted lists (lane g owns list g; G <= 32).
// `gstride` = bytes between consecutive shards' blocks of each array (0: the arrays are dense [G][q][kstride]);
// a non-zero stride lets all three arrays live in ONE packed per-shard block, i.e. one all-gather per search.
__global__ void merge_topk_kernel(const uint64_t* ids, const float* dist, const uint32_t* counts, int G,
                                  int64_t q, int64_t kstride, int64_t gstride, const uint32_t* ks, uint64_t* out_ids,
                                  float* out_dist, uint32_t* out_counts, int64_t out_kstride) {
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= q) return;
  const uint32_t k = ks[w];
  uint32_t len = 0, pos = 0;
  const uint64_t* my_ids = nullptr;
  const float* my_dist = nullptr;
  if (lane < G) {
    const size_t gi = gstride ? (size_t)lane * (size_t)(gstride / 8) : (size_t)lane * q * kstride;
    const size_t gd = gstride ? (size_t)lane * (size_t)(gstride / 4) : (size_t)lane * q * kstride;
    const size_t gc = gstride ? (size_t)lane * (size_t)(gstride / 4) : (size_t)lane * q;
    len = counts[gc + w];
    my_ids = ids + gi + (size_t)w * kstride;
    my_dist = dist + gd + (size_t)w * kstride;
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

// K6: exact distance of explicit (query, row id) pairs -- DistanceMetric::distance (src/distance.rs:20-33) as the
// HNSW code calls it for candidate lists (src/hnsw/graph.rs:221-232, search_layer).  One warp per pair: lane 0 finds
// the row by binary search over the id column (slot order == id order), the warp stages row and query chunks in
// shared memory with coalesced loads and lane 0 walks them with the reference's sequential f32 chain.
constexpr int kSpWarps = 4, kSpChunk = 1024;
template <int METRIC>
__global__ void __launch_bounds__(kSpWarps * 32) score_pairs_kernel(const ScorePairsParams p) {
  griddep_wait();
  __shared__ __align__(16) float s_x[kSpWarps][kSpChunk];
  __shared__ __align__(16) float s_qv[kSpWarps][kSpChunk];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const IndexView& iv = p.iv;
  const int64_t total = p.q * p.m;
  for (int64_t w = (int64_t)blockIdx.x * kSpWarps + wib; w < total; w += (int64_t)gridDim.x * kSpWarps) {
    const int64_t qi = w / p.m;
    int64_t slot = -1;
    if (lane == 0) {
      const uint64_t id = p.cand_ids[w];
      int64_t lo = 0, hi = iv.n_slots;  // lower bound of id in the ascending id column
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (iv.ids[mid] < id) lo = mid + 1; else hi = mid;
      }
      if (lo < iv.n_slots && iv.ids[lo] == id && ((iv.live[lo >> 5] >> (lo & 31)) & 1u)) slot = lo;
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot < 0) {
      if (lane == 0) {
        p.out_dist[w] = __int_as_float(0x7f800000);
        p.out_status[w] = 1;
      }
      continue;
    }
    const float* xr = iv.x32 + (size_t)slot * iv.dpad;
    const float* qr = p.q32 + (size_t)qi * iv.dpad;
    float acc = -0.0f;
    for (int c0 = 0; c0 < iv.d; c0 += kSpChunk) {
      const int nf = min(kSpChunk, iv.dpad - c0);  // multiple of 4
      for (int t = lane * 4; t < nf; t += 128) {
        *reinterpret_cast<float4*>(&s_x[wib][t]) = __ldg(reinterpret_cast<const float4*>(xr + c0 + t));
        *reinterpret_cast<float4*>(&s_qv[wib][t]) = __ldg(reinterpret_cast<const float4*>(qr + c0 + t));
      }
      __syncwarp();
      if (lane == 0) {
        const int lim = min(kSpChunk, iv.d - c0);
#pragma unroll 8
        for (int i = 0; i < lim; ++i) acc = exact_step<METRIC>(acc, s_qv[wib][i], s_x[wib][i]);
      }
      __syncwarp();
    }
    if (lane == 0) {
      float dist;
      uint8_t status = 0;
      if (METRIC == kMetricL2) {
        dist = __fsqrt_rn(acc);
      } else if (METRIC == kMetricDot) {
        dist = -acc;
      } else {
        const float xn = iv.norm[slot], qn = p.qnorm[qi];
        if (xn == 0.f || qn == 0.f) {
          dist = __int_as_float(0x7f800000);
          status = 2;
        } else {
          float sim = __fdiv_rn(acc, __fmul_rn(qn, xn));
          if (sim < -1.0f) sim = -1.0f;
          else if (sim > 1.0f) sim = 1.0f;
          dist = __fsub_rn(1.0f, sim);
        }
      }
      if (dist == 0.f) dist = 0.f;  // -0.0 -> +0.0, as everywhere in this library
      p.out_dist[w] = dist;
      p.out_status[w] = status;
    }
  }
}

}  // namespace

cudaError_t launch_select_rerank(const SelectParams& p, int grid, cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  if (p.KP < 32 || p.KP > kMaxKP || (p.KP & (p.KP - 1))) return cudaErrorInvalidValue;
  const size_t sel_cap = p.sel_cap > 0 ? (size_t)p.sel_cap : (size_t)kSelectStageKeys;
  const size_t sel_smem = ((size_t)p.KP + sel_cap * (p.list_len > 0 ? 2 : 1)) * 8;
  if (sel_smem > 48 * 1024) {
    if (sel_smem > 200 * 1024) return cudaErrorInvalidValue;
    static unsigned long long big_smem_set = 0;  // bit per device (function attributes are per device); a racing
    int dev = 0;                                 // duplicate cudaFuncSetAttribute call is harmless
    cudaGetDevice(&dev);
    if (!((big_smem_set >> (dev & 63)) & 1ull)) {
      cudaError_t ea = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (ea != cudaSuccess) return ea;
      big_smem_set |= 1ull << (dev & 63);
    }
  }
  cudaError_t e = launch_pdl(select_kernel, dim3(grid), dim3(kSelThreads), sel_smem, st, p);
  if (e != cudaSuccess) return e;
  SelectParams pp = p;
  const int nq_max = p.nq_dev ? p.nq_max : p.nq;
  pp.few_candidates = (int64_t)nq_max * p.KP <= 4096 ? 1 : 0;
  pp.warp_per_candidate = ((int64_t)nq_max * p.KP <= 2048 && 2 * p.iv.dpad <= 2 * kRrTile) ? 1 : 0;
  const int64_t warps = (int64_t)nq_max * (pp.warp_per_candidate ? p.KP : p.KP / kRrRPW);
  const int blocks = (int)std::min<int64_t>((warps + kRrWarps - 1) / kRrWarps, 148 * 16);
  switch (p.iv.metric) {
    case kMetricL2: return launch_pdl(rerank_finalize_kernel<kMetricL2>, dim3(blocks), dim3(kRrWarps * 32), 0, st, pp);
    case kMetricCos: return launch_pdl(rerank_finalize_kernel<kMetricCos>, dim3(blocks), dim3(kRrWarps * 32), 0, st, pp);
    case kMetricDot: return launch_pdl(rerank_finalize_kernel<kMetricDot>, dim3(blocks), dim3(kRrWarps * 32), 0, st, pp);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_score_pairs(const ScorePairsParams& p, cudaStream_t st) {
  const int64_t total = p.q * p.m;
  if (total <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((total + kSpWarps - 1) / kSpWarps, 148 * 8);
  switch (p.iv.metric) {
    case kMetricL2: return launch_pdl(score_pairs_kernel<kMetricL2>, dim3(blocks), dim3(kSpWarps * 32), 0, st, p);
    case kMetricCos: return launch_pdl(score_pairs_kernel<kMetricCos>, dim3(blocks), dim3(kSpWarps * 32), 0, st, p);
    case kMetricDot: return launch_pdl(score_pairs_kernel<kMetricDot>, dim3(blocks), dim3(kSpWarps * 32), 0, st, p);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_merge(const uint64_t* ids, const float* dist, const uint32_t* counts, int G, int64_t q,
                         int64_t kstride, int64_t gstride, const uint32_t* ks, uint64_t* out_ids, float* out_dist,
                         uint32_t* out_counts, int64_t out_kstride, cudaStream_t st) {
  if (q <= 0) return cudaSuccess;
  if (G > 32 || gstride < 0 || (gstride & 7)) return cudaErrorInvalidValue;
  const int64_t blocks = (q * 32 + 255) / 256;
  return launch_pdl(merge_topk_kernel, dim3((unsigned)blocks), dim3(256), 0, st, ids, dist, counts, G, q, kstride, gstride,
                    ks, out_ids, out_dist, out_counts, out_kstride);
}

}  // namespace gfi
