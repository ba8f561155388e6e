// scan.cu -- K1 `flat_scan_topk`: bandwidth-bound streaming scan of the fp32 rows with a
// fused distance epilogue, eligibility (tombstone + metadata-filter bitmask) test and per-warp
// top-K lists.  Replaces the score-all + full-sort of the reference's FlatIndex::search
// (src/flat_index.rs:53-63) and the per-pair metric loops (src/distance.rs:37-73) for small
// query batches (q <= 8 queries share one pass over the database).
//
// Shape of the work: no reuse of the database bytes, so the kernel is HBM-bound.  One
// persistent CTA per SM; a producer warp streams row blocks into a 4 x 32 KB shared-memory
// ring with 1-D bulk async copies (TMA engine, SASS UBLKCP) completing on mbarriers; eight
// consumer warps read the ring with conflict-free 128-bit loads (8 lanes per row), keep QT
// accumulators per lane and insert into per-warp sorted lists only when a row beats the
// list's current K-th key.  The scores computed here only rank candidates; the K survivors
// per CTA are re-scored with the reference's exact arithmetic in select_rerank.cu.
#include "common.cuh"
#include "kernels.h"

namespace gfi {

namespace {

__device__ __forceinline__ void warp_list_insert(uint64_t* list, int K, uint64_t key, int lane) {
  // list ascending, key < list[K-1]; the last entry falls off.
  bool placed = false;
  for (int c = K / 32 - 1; c >= 0; --c) {
    const int i = c * 32 + lane;
    const uint64_t e = list[i];
    const bool gt = e > key;
    const unsigned m = __ballot_sync(0xffffffffu, gt);
    if (gt && i + 1 < K) list[i + 1] = e;
    if (m != 0xffffffffu) {
      const int pos = c * 32 + __popc(~m);
      __syncwarp();
      if (lane == 0) list[pos] = key;
      placed = true;
      break;
    }
    __syncwarp();
  }
  if (!placed) {
    __syncwarp();
    if (lane == 0) list[0] = key;
  }
  __syncwarp();
}

// In-place ascending bitonic sort of arr[0..N) (N a power of two) by the 256 consumer threads.
__device__ __forceinline__ void consumers_bitonic_sort(uint64_t* arr, int N, int tid) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < N / 2; t += kScanConsumerWarps * 32) {
        const int i = ((t / j) * 2 * j) + (t % j);
        const int l = i + j;
        const bool up = ((i & k) == 0);
        const uint64_t a = arr[i], b = arr[l];
        if ((a > b) == up) {
          arr[i] = b;
          arr[l] = a;
        }
      }
      named_bar_sync(1, kScanConsumerWarps * 32);
    }
  }
}

template <int METRIC, int QT, int LPR>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(const ScanParams p) {
  constexpr int G = 32 / LPR;  // rows processed concurrently by one warp (LPR lanes per row)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const IndexView& iv = p.iv;
  const int dpad = iv.dpad;
  const int K = p.K;
  float* stages = reinterpret_cast<float*>(smem_raw);
  const int NS = p.nstages;
  float* qs = stages + (size_t)NS * kScanStageFloats;
  uint64_t* lists = reinterpret_cast<uint64_t*>(qs + QT * dpad);  // [QT][8 warps][K]
  uint64_t* full = lists + (size_t)QT * kScanConsumerWarps * K;
  uint64_t* empty = full + kScanMaxStages;

  const int nq = p.nq_dev ? (int)*p.nq_dev : p.nq;
  if (nq <= 0) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kScanConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int64_t n = iv.n_slots;
  const int R = p.rows_per_stage, nseg = p.nseg, segw = p.seg_floats;
  const int64_t nblocks = (n + R - 1) / R;
  const int rstride = (nseg == 1) ? dpad : segw;
  uint32_t it = 0;  // ring iteration; producer and consumers walk the same sequence

  for (int q0 = 0; q0 < nq; q0 += QT) {
    const int nqp = min(QT, nq - q0);
    for (int i = tid; i < QT * dpad; i += kScanThreads) {
      const int qi = i / dpad, c = i - qi * dpad;
      float v = 0.f;
      if (qi < nqp) {
        const uint32_t qg = p.qlist ? p.qlist[q0 + qi] : (uint32_t)(q0 + qi);
        v = p.q32[(size_t)qg * dpad + c];
      }
      qs[i] = v;
    }
    for (int i = tid; i < QT * kScanConsumerWarps * K; i += kScanThreads) lists[i] = kKeySentinel;
    __syncthreads();

    if (warp == kScanConsumerWarps) {
      // ---------------- producer: bulk async copies into the ring ----------------
      for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t row0 = blk * R;
        const int rows = (int)min((int64_t)R, n - row0);
        for (int seg = 0; seg < nseg; ++seg, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          float* dst = stages + (size_t)s * kScanStageFloats;
          if (nseg == 1) {
            if (lane == 0) {
              const uint32_t bytes = (uint32_t)rows * dpad * 4u;
              mbar_arrive_expect_tx(&full[s], bytes);
              bulk_g2s(dst, iv.x32 + row0 * dpad, bytes, &full[s]);
            }
          } else {
            const int segf = min(segw, dpad - seg * segw);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)rows * segf * 4u);
            __syncwarp();
            if (lane < rows)
              bulk_g2s(dst + lane * segw, iv.x32 + (row0 + lane) * dpad + (size_t)seg * segw, segf * 4u,
                       &full[s]);
          }
          __syncwarp();
        }
      }
    } else {
      // ---------------- consumers ----------------
      const int g = lane / LPR, j = lane % LPR;
      uint64_t T[QT];
      uint64_t floor64[QT];
      float qn[QT];
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) {
        T[qi] = kKeySentinel;
        floor64[qi] = 0;
        qn[qi] = 1.f;
        if (qi < nqp) {
          const uint32_t qg = p.qlist ? p.qlist[q0 + qi] : (uint32_t)(q0 + qi);
          if (p.floor64) floor64[qi] = p.floor64[qg];
        }
      }
      float acc[QT];
      for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t row0 = blk * R;
        const int rows = (int)min((int64_t)R, n - row0);
        for (int seg = 0; seg < nseg; ++seg, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1u;
          mbar_wait(&full[s], ph);
          const float* sb = stages + (size_t)s * kScanStageFloats;
          const int segf = (nseg == 1) ? dpad : min(segw, dpad - seg * segw);
          const int nf4 = segf >> 2;
          const float4* qb = reinterpret_cast<const float4*>(qs + (size_t)seg * segw);
          for (int i0 = warp * G; i0 < rows; i0 += kScanConsumerWarps * G) {
            const int rl = i0 + g;
            const int64_t slot = row0 + rl;
            // eligibility + per-row scalars are fetched early so the loads overlap the FMA loop
            bool elig = false;
            float rnorm = 1.f;
            if (seg == nseg - 1 && rl < rows) {
              elig = (iv.live[slot >> 5] >> (slot & 31)) & 1u;
              if (elig && p.mask.bits) {
                const uint64_t id = iv.ids_identity ? (uint64_t)slot : iv.ids[slot];
                elig = (id < (uint64_t)p.mask.nbits) && ((p.mask.bits[id >> 6] >> (id & 63)) & 1ull);
              }
              if (METRIC == kMetricCos) rnorm = iv.norm[slot];
            }
            if (nseg == 1 || seg == 0) {
#pragma unroll
              for (int qi = 0; qi < QT; ++qi) acc[qi] = 0.f;
            }
            const float4* xr = reinterpret_cast<const float4*>(sb + (size_t)rl * rstride);
#pragma unroll 4
            for (int t = j; t < nf4; t += LPR) {
              const float4 x = xr[t];
#pragma unroll
              for (int qi = 0; qi < QT; ++qi) {
                const float4 qv = qb[qi * (dpad >> 2) + t];
                if (METRIC == kMetricL2) {
                  const float a = qv.x - x.x, b = qv.y - x.y, c = qv.z - x.z, e = qv.w - x.w;
                  acc[qi] = fmaf(a, a, acc[qi]);
                  acc[qi] = fmaf(b, b, acc[qi]);
                  acc[qi] = fmaf(c, c, acc[qi]);
                  acc[qi] = fmaf(e, e, acc[qi]);
                } else {
                  acc[qi] = fmaf(qv.x, x.x, acc[qi]);
                  acc[qi] = fmaf(qv.y, x.y, acc[qi]);
                  acc[qi] = fmaf(qv.z, x.z, acc[qi]);
                  acc[qi] = fmaf(qv.w, x.w, acc[qi]);
                }
              }
            }
            if (seg == nseg - 1) {
              if (METRIC == kMetricCos && elig && j == 0 && rnorm == 0.f) atomicOr(p.flags, kFlagZeroNorm);
              const float inv = (METRIC == kMetricCos) ? (1.0f / rnorm) : 1.f;
#pragma unroll
              for (int qi = 0; qi < QT; ++qi) {
                float v = acc[qi];
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (qi < nqp) {
                  const float score = (METRIC == kMetricL2) ? v : (METRIC == kMetricCos ? -v * inv : -v);
                  const bool mine = elig && j == 0;
                  if (mine && score != score && !(METRIC == kMetricCos && rnorm == 0.f))
                    atomicOr(p.flags, kFlagNaN);
                  const uint64_t key = pack_key(score, (uint32_t)slot);
                  const bool hit = mine && (score == score) && key < T[qi] && key > floor64[qi];
                  unsigned m = __ballot_sync(0xffffffffu, hit);
                  while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint64_t k64 = __shfl_sync(0xffffffffu, key, src);
                    if (k64 < T[qi]) {
                      uint64_t* list = lists + ((size_t)qi * kScanConsumerWarps + warp) * K;
                      warp_list_insert(list, K, k64, lane);
                      T[qi] = list[K - 1];
                    }
                  }
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
      }
      (void)qn;
    }
    __syncthreads();

    // ---------------- CTA merge: 8 sorted lists -> top K per query, written to cand ----------------
    if (warp < kScanConsumerWarps) {
      for (int qi = 0; qi < nqp; ++qi) {
        uint64_t* arr = lists + (size_t)qi * kScanConsumerWarps * K;
        consumers_bitonic_sort(arr, kScanConsumerWarps * K, tid);
        const uint32_t qg = p.qlist ? p.qlist[q0 + qi] : (uint32_t)(q0 + qi);
        uint64_t* out = p.cand + (size_t)qg * p.cand_stride + (size_t)blockIdx.x * K;
        for (int i = tid; i < K; i += kScanConsumerWarps * 32) out[i] = arr[i];
        if (blockIdx.x == 0 && tid == 0) p.cand_cnt[qg] = gridDim.x * (uint32_t)K;
      }
    }
    __syncthreads();
  }
}

template <int METRIC, int LPR>
cudaError_t launch_scan_metric(const ScanParams& p, int QT, int grid, size_t smem, cudaStream_t st) {
#define GFI_SCAN_CASE(Q)                                                                                  \
  case Q: {                                                                                               \
    auto kern = scan_topk_kernel<METRIC, Q, LPR>;                                                              \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e != cudaSuccess) return e;                                                                       \
    kern<<<grid, kScanThreads, smem, st>>>(p);                                                            \
    return cudaGetLastError();                                                                            \
  }
  switch (QT) {
    GFI_SCAN_CASE(1)
    GFI_SCAN_CASE(2)
    GFI_SCAN_CASE(4)
    GFI_SCAN_CASE(8)
  }
#undef GFI_SCAN_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

size_t scan_smem_bytes(int QT, int dpad, int K, int nstages) {
  return (size_t)nstages * kScanStageFloats * 4 + (size_t)QT * dpad * 4 +
         (size_t)QT * kScanConsumerWarps * K * 8 + 2 * kScanMaxStages * 8 + 16;
}

cudaError_t launch_scan(const ScanParams& p, int QT, int grid, cudaStream_t st) {
  const size_t smem = scan_smem_bytes(QT, p.iv.dpad, p.K, p.nstages);
  if (p.lanes_per_row == 8) {
    switch (p.iv.metric) {
      case kMetricL2: return launch_scan_metric<kMetricL2, 8>(p, QT, grid, smem, st);
      case kMetricCos: return launch_scan_metric<kMetricCos, 8>(p, QT, grid, smem, st);
      case kMetricDot: return launch_scan_metric<kMetricDot, 8>(p, QT, grid, smem, st);
    }
  } else if (p.lanes_per_row == 32) {
    switch (p.iv.metric) {
      case kMetricL2: return launch_scan_metric<kMetricL2, 32>(p, QT, grid, smem, st);
      case kMetricCos: return launch_scan_metric<kMetricCos, 32>(p, QT, grid, smem, st);
      case kMetricDot: return launch_scan_metric<kMetricDot, 32>(p, QT, grid, smem, st);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace gfi
