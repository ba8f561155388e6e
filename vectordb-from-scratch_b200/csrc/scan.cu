// scan.cu -- K1 `flat_scan_topk`: bandwidth-bound streaming scan of the fp32 rows with a
// fused distance epilogue, eligibility (tombstone + metadata-filter bitmask) test and per-warp
// top-K lists.  Replaces the score-all + full-sort of the reference's FlatIndex::search
// (src/flat_index.rs:53-63) and the per-pair metric loops (src/distance.rs:37-73) for small
// query batches (QT <= 4 queries share one pass over the database).
//
// Shape of the work: no reuse of the database bytes, so the kernel is HBM-bound and the design
// is about keeping bytes in flight and the per-row dependency chain short:
//   * one persistent CTA per SM; a producer warp streams row blocks into a shared-memory ring
//     with 1-D bulk async copies (TMA engine, SASS UBLKCP) that complete on mbarriers.  A stage
//     is ONE contiguous copy of whole rows (24-64 KB) whenever rows fit; a bandwidth probe
//     (scripts/probes/bw_probe.cu) shows this path streams 7.3 TB/s on B200 with >= 16 KB copies;
//   * sixteen consumer warps read the ring with conflict-free 128-bit loads: 8 lanes per row
//     (4 rows per warp at a time) for rows <= 1 KB, one warp per row for longer rows;
//   * tombstone / filter bits and the row norm for the NEXT stage are fetched while the current
//     stage is processed, so no global-memory latency sits on a warp's per-row critical path;
//   * a row touches the per-warp sorted list only when it beats the list's current K-th key.
// The scores computed here only rank candidates; the K survivors per CTA are re-scored with the
// reference's exact arithmetic in select_rerank.cu -- or, for small single-pass searches, by this
// kernel's own last CTA (scan_fused_tail below), which then also emits the results.
#include <algorithm>

#include "common.cuh"
#include "exact.cuh"
#include "kernels.h"

namespace gfi {

namespace {

constexpr int NW = kScanConsumerWarps;
constexpr int kMaxRounds = 4;  // row rounds per stage (rows_per_stage = rounds * NW * rows-per-warp)

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

// In-place ascending bitonic sort of arr[0..N) (N a power of two) by the consumer threads.
__device__ __forceinline__ void consumers_bitonic_sort(uint64_t* arr, int N, int tid) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < N / 2; t += NW * 32) {
        const int i = ((t / j) * 2 * j) + (t % j);
        const int l = i + j;
        const bool up = ((i & k) == 0);
        const uint64_t a = arr[i], b = arr[l];
        if ((a > b) == up) {
          arr[i] = b;
          arr[l] = a;
        }
      }
      named_bar_sync(1, NW * 32);
    }
  }
}

// Row metadata fetched one block ahead as RAW words (nothing depends on the loads until the next
// iteration, so their latency is hidden behind the current block's work).
struct RowMeta {
  uint32_t slot[kMaxRounds];
  uint32_t live[kMaxRounds];
  uint64_t mask[kMaxRounds];
  float norm[kMaxRounds];
};

// Fused tail, run by the last CTA of a small single-pass search (ScanParams::fused): K3 in place.  Per query:
// (A) all CTAs' K-lists into shared memory; (B) the K-th smallest of the lists' first few keys bounds the K-th
// smallest overall, the keys below the bound are ranked by counting (keys are unique: they carry the slot) and the
// K smallest land in ascending order; (C) a warp per candidate pulls the row, lane 0 walks it with the reference's
// sequential arithmetic (exact.cuh); (D) the exact keys are ranked again and the first k emitted -- the same
// select -> rerank -> sort -> truncate as select_kernel + rerank_finalize_kernel (src/flat_index.rs:53-63).
template <int METRIC>
__device__ __noinline__ void scan_fused_tail(const ScanParams& p, int nq, const float* qs, unsigned char* scratch,
                                             uint64_t* ctl, int tid) {
  constexpr int NT = kScanThreads, NWARP = NT / 32;
  // (control words in the dynamic allocation: a static __shared__ here would push the kernel past the 227 KB limit)
  uint64_t& s_bound = ctl[0];
  uint32_t& s_n = reinterpret_cast<uint32_t*>(ctl + 1)[0];
  uint32_t& s_m = reinterpret_cast<uint32_t*>(ctl + 1)[1];
  float& s_tau = reinterpret_cast<float*>(ctl + 3)[0];  // exact k-th distance of the query being finished
  const IndexView& iv = p.iv;
  const int K = p.K, L = (int)gridDim.x, dpad = iv.dpad, LK = L * K;
  const int lane = tid & 31, warp = tid >> 5;
  uint64_t* sel = reinterpret_cast<uint64_t*>(scratch);  // [K] best approximate keys, ascending
  uint64_t* ex = sel + K;                                // [K] exact keys
  uint64_t* keys = ex + K;                               // [L*K] every CTA's list
  uint64_t* cands = keys + LK;                           // [L*K] heads, then the keys below the bound
  float* tiles = reinterpret_cast<float*>(keys);         // rerank: one row per warp (keys/cands are dead by then)
  for (int qi = 0; qi < nq; ++qi) {
    const uint32_t qg = p.qlist ? p.qlist[qi] : (uint32_t)qi;
    const uint64_t* cand = p.cand + (size_t)qg * p.cand_stride;
    if (tid == 0) { s_n = 0; s_m = 0; s_bound = kKeySentinel - 1; }
    __syncthreads();
    uint32_t nv = 0;
    for (int i0 = 0; i0 < LK; i0 += NT * 4) {  // four independent L2 round trips in flight per thread
      uint64_t kk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * NT + tid;
        kk[u] = i < LK ? __ldcg(cand + i) : kKeySentinel;  // written by the other CTAs: read at L2
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * NT + tid;
        if (i < LK) keys[i] = kk[u];
        nv += kk[u] != kKeySentinel;
      }
    }
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    if (lane == 0 && nv) atomicAdd(&s_n, nv);
    __syncthreads();
    const uint32_t nvalid = s_n, kpeff = min((uint32_t)K, nvalid);
    if (nvalid > (uint32_t)K) {
      const int m = min(K, (K + L - 1) / L + 1), nh = L * m;  // nh <= K + 2L
      for (int h = tid; h < nh; h += NT) cands[h] = keys[(h / m) * K + (h % m)];
      __syncthreads();
      for (int h = tid; h < nh; h += NT) {
        const uint64_t key = cands[h];
        if (key != kKeySentinel) {
          uint32_t below = 0;
          for (int j = 0; j < nh; ++j) below += cands[j] < key;
          if (below == (uint32_t)K - 1) s_bound = key;  // (with fewer than K valid heads the bound stays open)
        }
      }
      __syncthreads();
    }
    const uint64_t bound = s_bound;
    for (int i = tid; i < LK; i += NT) {
      const uint64_t k64 = keys[i];
      if (k64 <= bound) cands[atomicAdd(&s_m, 1u)] = k64;  // the sentinel is above any bound
    }
    __syncthreads();
    const int M = (int)s_m;  // >= kpeff
    for (int i = tid; i < M; i += NT) {
      const uint64_t key = cands[i];
      uint32_t below = 0;
      for (int j = 0; j < M; ++j) below += cands[j] < key;
      if (below < (uint32_t)K) sel[below] = key;
    }
    __syncthreads();
    const float qn = p.qnorm[qg];
    const float* qv = qs + (size_t)qi * dpad;
    for (int c = warp; c < (int)kpeff; c += NWARP) {
      const uint32_t slot = (uint32_t)(sel[c] & 0xffffffffu);
      float* xr = tiles + (size_t)warp * dpad;
      const float4* x4 = reinterpret_cast<const float4*>(iv.x32 + (size_t)slot * dpad);
      for (int t = lane; t < (dpad >> 2); t += 32) reinterpret_cast<float4*>(xr)[t] = __ldg(x4 + t);
      __syncwarp();
      if (lane == 0) {
        float acc = -0.0f;
#pragma unroll 8
        for (int i = 0; i < iv.d; ++i) acc = exact_step<METRIC>(acc, qv[i], xr[i]);
        ex[c] = pack_key(exact_finish<METRIC>(acc, METRIC == kMetricCos ? iv.norm[slot] : 1.f, qn, p.flags), slot);
      }
      __syncwarp();
    }
    __syncthreads();
    if (METRIC == kMetricCos && nvalid > 0 && qn == 0.f && tid == 0) atomicOr(p.flags, kFlagZeroNorm);
    const uint32_t kq = min(p.ks[qg], kpeff);
    for (int i = tid; i < (int)kpeff; i += NT) {
      const uint64_t key = ex[i];
      uint32_t below = 0;
      for (int j = 0; j < (int)kpeff; ++j) below += ex[j] < key;
      if (below < kq) {
        p.out_ids[(size_t)qg * p.kstride + below] = iv.ids[(uint32_t)(key & 0xffffffffu)];
        p.out_dist[(size_t)qg * p.kstride + below] = key_f32((uint32_t)(key >> 32));
        if (p.h_ctrl) {
          p.h_out_ids[(size_t)qg * p.kstride + below] = iv.ids[(uint32_t)(key & 0xffffffffu)];
          p.h_out_dist[(size_t)qg * p.kstride + below] = key_f32((uint32_t)(key >> 32));
        }
        if (below == kq - 1) s_tau = key_f32((uint32_t)(key >> 32));
      }
    }
    __syncthreads();
    if (tid == 0) {
      p.out_counts[qg] = kq;
      if (p.h_ctrl) p.h_out_counts[qg] = kq;
      // certification, as in rerank_finalize_kernel: rows were dropped only if K or more keys existed, and then
      // every dropped row's approximate key is >= sel[K-1] (exact.cuh, scan_lower_bound)
      const uint32_t k = p.ks[qg];
      if (p.up_list && k > 0 && nvalid >= (uint32_t)K) {
        bool ok = kq == k;
        if (ok) {
          const float a_s = key_f32((uint32_t)(sel[K - 1] >> 32));
          ok = (a_s == a_s) && scan_lower_bound(METRIC, a_s, qn, p.xnorm_max, iv.d) > s_tau;
        }
        if (!ok) {
          const uint32_t pos = atomicAdd(p.up_count, 1u);
          if (pos < p.up_cap) p.up_list[pos] = qg + p.up_base;
        }
      }
    }
    __syncthreads();
  }
  if (p.h_ctrl) {
    // latency mode: results are in host memory; publish the control block, then the word the host polls
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const volatile uint32_t* dc = p.d_ctrl;  // flags raised by the other CTAs precede their tickets
      for (int w = 0; w < kCtrlWords; ++w) p.h_ctrl[w] = dc[w];
      __threadfence_system();
      *reinterpret_cast<volatile uint32_t*>(p.h_ctrl + kCtrlDoneWord) = p.done_seq;
    }
  }
}

template <int METRIC, int QT, int LPR, bool SEG>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(const __grid_constant__ ScanParams p) {
  griddep_wait();
  constexpr int G = 32 / LPR;              // rows processed concurrently by one warp (LPR lanes per row)
  constexpr bool FAST = (QT == 1) && !SEG;  // single query, whole rows per stage: query lives in registers
  constexpr int kQRegs = 8;                 // float4 per lane per row on the FAST path (dpad <= 256 * G)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const IndexView& iv = p.iv;
  const int dpad = iv.dpad;
  const int K = p.K;
  const int NS = p.nstages;
  const int stage_floats = p.stage_floats;
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* qs = stages + (size_t)NS * stage_floats;
  uint64_t* lists = reinterpret_cast<uint64_t*>(qs + QT * dpad);  // [QT][NW][K]
  uint64_t* full = lists + (size_t)QT * NW * K;
  uint64_t* empty = full + kScanMaxStages;

  const int nq = p.nq_dev ? (int)*p.nq_dev : p.nq;
  if (nq <= 0) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // gather mode (filtered / tombstoned scans): rows are taken from a compacted list of eligible slots,
  // one bulk copy per run of adjacent slots, so ineligible rows cost no bandwidth at all
  const uint32_t* glist = p.gather_list;
  const uint32_t n = glist ? *p.gather_count : (uint32_t)iv.n_slots;
  if (glist && p.elig_out && blockIdx.x == 0 && tid == 0) *p.elig_out = n + 1u;
  const int R = p.rows_per_stage, nseg = SEG ? p.nseg : 1, segw = p.seg_floats;
  const uint32_t nblocks = (n + R - 1) / R;
  const int rstride = SEG ? segw : dpad;
  const int rounds = R / (NW * G);
  int stage = 0;       // ring position; producer and consumers walk the same sequence
  uint32_t phase = 0;  // parity of the ring pass

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
    for (int i = tid; i < QT * NW * K; i += kScanThreads) lists[i] = kKeySentinel;
    __syncthreads();

    if (warp == NW) {
      // ---------------- producer: bulk async copies into the ring ----------------
      for (uint32_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const uint32_t row0 = blk * R;
        const int rows = (int)min((uint32_t)R, n - row0);
        for (int seg = 0; seg < nseg; ++seg) {
          mbar_wait(&empty[stage], phase ^ 1u);
          float* dst = stages + (size_t)stage * stage_floats;
          if (!SEG && glist) {
            if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)rows * dpad * 4u);
            __syncwarp();
            // Listed slots that are adjacent in the slot array are adjacent in HBM AND in the stage, so a run of
            // them is ONE copy (measured at 50 % selectivity, mean run 2 rows of 1.5 KB: 4.7 -> 5.3 TB/s of touched
            // bytes; longer runs from a better ordered list change nothing more -- half-used DRAM pages, not the
            // number of copies, bound the gather from there).  Runs are cut at 32-entry chunk boundaries (one
            // ballot per chunk, no cross-chunk state).
            for (int base = 0; base < rows; base += 32) {
              const int r = base + lane;
              const bool in = r < rows;
              const uint32_t s = in ? __ldg(glist + row0 + r) : 0u;
              const uint32_t sprev = __shfl_up_sync(0xffffffffu, s, 1);
              const bool head = in && (lane == 0 || sprev + 1u != s);
              const uint32_t heads = __ballot_sync(0xffffffffu, head);
              if (head) {
                const uint32_t above = (lane == 31) ? 0u : (heads >> (lane + 1));
                const int end = above ? lane + __ffs(above) : min(32, rows - base);
                bulk_g2s(dst + (size_t)r * dpad, iv.x32 + (size_t)s * dpad, (uint32_t)(end - lane) * dpad * 4u,
                         &full[stage]);
              }
            }
          } else if (!SEG) {
            if (lane == 0) {
              const uint32_t bytes = (uint32_t)rows * dpad * 4u;
              mbar_arrive_expect_tx(&full[stage], bytes);
              bulk_g2s(dst, iv.x32 + (size_t)row0 * dpad, bytes, &full[stage]);
            }
          } else {
            // rows longer than a segment: one copy per row and column segment (rows <= 32 here)
            const int segf = min(segw, dpad - seg * segw);
            if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)rows * segf * 4u);
            __syncwarp();
            if (lane < rows)
              bulk_g2s(dst + lane * segw,
                       iv.x32 + (size_t)(glist ? glist[row0 + lane] : row0 + lane) * dpad + (size_t)seg * segw,
                       segf * 4u, &full[stage]);
          }
          __syncwarp();
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    } else {
      // ---------------- consumers ----------------
      const int g = lane / LPR, j = lane % LPR;
      const int rsub = warp * G + g;  // this lane's row within a round
      uint64_t T[QT];
      uint64_t floor64[QT];
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) {
        T[qi] = kKeySentinel;
        floor64[qi] = 0;
        if (qi < nqp) {
          const uint32_t qg = p.qlist ? p.qlist[q0 + qi] : (uint32_t)(q0 + qi);
          if (p.floor64) floor64[qi] = p.floor64[qg];
        }
      }
      const bool has_mask = p.mask.bits != nullptr;
      const bool mask_by_slot = has_mask && iv.ids_identity;
      auto fetch_meta = [&](uint32_t blk, RowMeta& m) {
        const uint32_t row0 = blk * R;
#pragma unroll
        for (int r = 0; r < kMaxRounds; ++r) {
          m.live[r] = 0;
          m.mask[r] = ~0ull;
          m.norm[r] = 1.f;
          const uint32_t pos = row0 + r * (NW * G) + rsub;
          m.slot[r] = pos;
          if (r < rounds && pos < n) {
            if (glist) {
              // every listed row is eligible; its slot comes from the list (one more dependent load for
              // the cosine norm, still a whole block ahead of its use)
              const uint32_t s = __ldg(glist + pos);
              m.slot[r] = s;
              m.live[r] = ~0u;
              if (METRIC == kMetricCos) m.norm[r] = __ldg(iv.norm + s);
            } else {
              m.live[r] = __ldg(iv.live + (pos >> 5));
              if (mask_by_slot) m.mask[r] = ((int64_t)pos < p.mask.nbits) ? __ldg(p.mask.bits + (pos >> 6)) : 0ull;
              if (METRIC == kMetricCos) m.norm[r] = __ldg(iv.norm + pos);
            }
          }
        }
      };
      RowMeta meta_next;
      if (blockIdx.x < nblocks) fetch_meta(blockIdx.x, meta_next);

      // FAST path: this lane's slice of the (single) query stays in registers for the whole pass
      float4 qreg[FAST ? kQRegs : 1];
      if (FAST) {
        const float4* q4 = reinterpret_cast<const float4*>(qs);
#pragma unroll
        for (int i = 0; i < kQRegs; ++i) {
          const int t = j + i * LPR;
          qreg[FAST ? i : 0] = (t < (dpad >> 2)) ? q4[t] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }

      float acc[QT];
      for (uint32_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const uint32_t row0 = blk * R;
        const int rows = (int)min((uint32_t)R, n - row0);
        const RowMeta meta = meta_next;
        if (blk + gridDim.x < nblocks) fetch_meta(blk + gridDim.x, meta_next);

        for (int seg = 0; seg < nseg; ++seg) {
          mbar_wait(&full[stage], phase);
          const float* sb = stages + (size_t)stage * stage_floats;
          const int segf = SEG ? min(segw, dpad - seg * segw) : dpad;
          const int nf4 = segf >> 2;
#pragma unroll
          for (int r = 0; r < kMaxRounds; ++r) {
            if (r * (NW * G) + warp * G >= rows) break;  // warp-uniform
            const int rl = r * (NW * G) + rsub;
            const uint32_t slot = meta.slot[r];
            const float4* xr = reinterpret_cast<const float4*>(sb + (size_t)rl * rstride);
            if (FAST) {
              float a0 = 0.f, a1 = 0.f;
#pragma unroll
              for (int i = 0; i < kQRegs; ++i) {
                const int t = j + i * LPR;
                if (t < nf4) {
                  const float4 x = xr[t];
                  const float4 qv = qreg[FAST ? i : 0];
                  float& a = (i & 1) ? a1 : a0;
                  if (METRIC == kMetricL2) {
                    const float d0 = qv.x - x.x, d1 = qv.y - x.y, d2 = qv.z - x.z, d3 = qv.w - x.w;
                    a = fmaf(d0, d0, a); a = fmaf(d1, d1, a); a = fmaf(d2, d2, a); a = fmaf(d3, d3, a);
                  } else {
                    a = fmaf(qv.x, x.x, a); a = fmaf(qv.y, x.y, a); a = fmaf(qv.z, x.z, a); a = fmaf(qv.w, x.w, a);
                  }
                }
              }
              acc[0] = a0 + a1;
            } else {
              if (!SEG || seg == 0) {
#pragma unroll
                for (int qi = 0; qi < QT; ++qi) acc[qi] = 0.f;
              }
              const float4* qb = reinterpret_cast<const float4*>(qs + (size_t)seg * segw);
#pragma unroll 4
              for (int t = j; t < nf4; t += LPR) {
                const float4 x = xr[t];
#pragma unroll
                for (int qi = 0; qi < QT; ++qi) {
                  const float4 qv = qb[qi * (dpad >> 2) + t];
                  if (METRIC == kMetricL2) {
                    const float d0 = qv.x - x.x, d1 = qv.y - x.y, d2 = qv.z - x.z, d3 = qv.w - x.w;
                    acc[qi] = fmaf(d0, d0, acc[qi]);
                    acc[qi] = fmaf(d1, d1, acc[qi]);
                    acc[qi] = fmaf(d2, d2, acc[qi]);
                    acc[qi] = fmaf(d3, d3, acc[qi]);
                  } else {
                    acc[qi] = fmaf(qv.x, x.x, acc[qi]);
                    acc[qi] = fmaf(qv.y, x.y, acc[qi]);
                    acc[qi] = fmaf(qv.z, x.z, acc[qi]);
                    acc[qi] = fmaf(qv.w, x.w, acc[qi]);
                  }
                }
              }
            }
            if (!SEG || seg == nseg - 1) {
              // eligibility: tombstone bit, then the filter bit (by slot when ids are the identity)
              bool elig = ((meta.live[r] >> (slot & 31)) & 1u) && ((meta.mask[r] >> (slot & 63)) & 1ull);
              if (elig && has_mask && !mask_by_slot && !glist) {
                const uint64_t id = iv.ids[slot];
                elig = (id < (uint64_t)p.mask.nbits) && ((p.mask.bits[id >> 6] >> (id & 63)) & 1ull);
              }
              const float rnorm = meta.norm[r];
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
                  const uint64_t key = pack_key(score, slot);
                  const bool hit = mine && (score == score) && key < T[qi] && key > floor64[qi];
                  unsigned m = __ballot_sync(0xffffffffu, hit);
                  while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint64_t k64 = __shfl_sync(0xffffffffu, key, src);
                    if (k64 < T[qi]) {
                      uint64_t* list = lists + ((size_t)qi * NW + warp) * K;
                      warp_list_insert(list, K, k64, lane);
                      T[qi] = list[K - 1];
                    }
                  }
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncthreads();

    // ---------------- CTA merge: NW sorted lists -> top K per query, written to cand ----------------
    if (warp < NW) {
      for (int qi = 0; qi < nqp; ++qi) {
        uint64_t* arr = lists + (size_t)qi * NW * K;
        consumers_bitonic_sort(arr, NW * K, tid);
        const uint32_t qg = p.qlist ? p.qlist[q0 + qi] : (uint32_t)(q0 + qi);
        uint64_t* out = p.cand + (size_t)qg * p.cand_stride + (size_t)blockIdx.x * K;
        for (int i = tid; i < K; i += NW * 32) out[i] = arr[i];
        if (blockIdx.x == 0 && tid == 0) p.cand_cnt[qg] = gridDim.x * (uint32_t)K;
      }
    }
    __syncthreads();
  }
  if (!SEG && p.fused) {
    uint64_t* ctl = empty + kScanMaxStages;  // 4 words behind the barriers (scan_smem_bytes)
    uint32_t& s_last = reinterpret_cast<uint32_t*>(ctl + 2)[0];
    __threadfence();  // this CTA's lists before its ticket
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(p.done_ctr, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    scan_fused_tail<METRIC>(p, nq, qs, smem_raw, ctl, tid);
  }
}

template <int METRIC, int LPR, bool SEG>
cudaError_t launch_scan_metric(const ScanParams& p, int QT, int grid, size_t smem, cudaStream_t st) {
#define GFI_SCAN_CASE(Q)                                                                                  \
  case Q: {                                                                                               \
    auto kern = scan_topk_kernel<METRIC, Q, LPR, SEG>;                                                    \
    static bool attr_set[16] = {};                                                                        \
    int dev = 0;                                                                                          \
    cudaGetDevice(&dev);                                                                                  \
    if (!attr_set[dev & 15]) { /* once per kernel and device: the driver call is not free */              \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);    \
      if (e != cudaSuccess) return e;                                                                     \
      attr_set[dev & 15] = true;                                                                          \
    }                                                                                                     \
    return launch_pdl(kern, dim3(grid), dim3(kScanThreads), smem, st, p);                                 \
  }
  switch (QT) {
    GFI_SCAN_CASE(1)
    GFI_SCAN_CASE(2)
    GFI_SCAN_CASE(4)
  }
#undef GFI_SCAN_CASE
  return cudaErrorInvalidValue;
}

template <int METRIC>
cudaError_t launch_scan_shape(const ScanParams& p, int QT, int grid, size_t smem, cudaStream_t st) {
  if (p.lanes_per_row == 8 && p.nseg == 1) return launch_scan_metric<METRIC, 8, false>(p, QT, grid, smem, st);
  if (p.lanes_per_row == 16 && p.nseg == 1) return launch_scan_metric<METRIC, 16, false>(p, QT, grid, smem, st);
  if (p.lanes_per_row == 32 && p.nseg == 1) return launch_scan_metric<METRIC, 32, false>(p, QT, grid, smem, st);
  if (p.lanes_per_row == 32) return launch_scan_metric<METRIC, 32, true>(p, QT, grid, smem, st);
  return cudaErrorInvalidValue;
}

// Compacts the eligible (live and unmasked) slots into a list.  A block takes tiles of 8192 slots: every thread
// owns one 32-slot eligibility word (read straight from the live / mask words when ids are the identity, built by
// ballots from the id column otherwise), a block-wide exclusive scan places the words, ONE atomic per tile
// reserves the tile's range, and the warps expand their words with coalesced stores.  Inside a tile the list is
// ascending, so adjacent eligible slots are adjacent in the list (the gather scan copies such runs as one bulk
// copy); tiles land in atomic order, which is irrelevant to the result (keys carry the slot).
// (The first version issued one atomic per 32 slots on a single counter: 312k serialised atomics at 10M slots.)
constexpr int kCompactThreads = 256;
template <bool IDENT>
__global__ void __launch_bounds__(kCompactThreads) compact_eligible_kernel(const IndexView iv, const MaskView mask,
                                                                          uint32_t* list, uint32_t* count) {
  griddep_wait();
  __shared__ uint32_t s_warp[kCompactThreads / 32];
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nwords = (iv.n_slots + 31) >> 5;
  const int64_t ntiles = (nwords + kCompactThreads - 1) / kCompactThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t w = tile * kCompactThreads + tid;  // this thread's word: slots [32w, 32w + 32)
    uint32_t e = 0;
    if (w < nwords) {
      e = __ldg(iv.live + w);
      const int64_t left = iv.n_slots - (w << 5);
      if (left < 32) e &= (1u << left) - 1u;
      if (IDENT && mask.bits) {
        const int64_t bleft = mask.nbits - (w << 5);  // ids >= nbits are ineligible
        uint32_t m = 0;
        if (bleft > 0) {
          m = (uint32_t)(__ldg(mask.bits + (w >> 1)) >> ((w & 1) * 32));
          if (bleft < 32) m &= (1u << bleft) - 1u;
        }
        e &= m;
      }
    }
    if (!IDENT && mask.bits) {
      // ids are arbitrary: the warp walks its 32 words together, a lane per slot, and ballots the mask bits
      const int64_t w0 = tile * kCompactThreads + warp * 32;
      uint32_t mine = 0;
      for (int j = 0; j < 32; ++j) {
        const uint32_t ej = __shfl_sync(0xffffffffu, e, j);
        bool bit = false;
        if ((ej >> lane) & 1u) {
          const uint64_t id = __ldg(iv.ids + ((w0 + j) << 5) + lane);
          bit = (id < (uint64_t)mask.nbits) && ((__ldg(mask.bits + (id >> 6)) >> (id & 63)) & 1ull);
        }
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if (lane == j) mine = word;
      }
      e = mine;
    }
    // block-wide exclusive scan of the words' populations
    const uint32_t cnt = (uint32_t)__popc(e);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kCompactThreads / 32; ++i) {
      const uint32_t t = s_warp[i];
      if (i < warp) woff += t;
      total += t;
    }
    if (tid == 0) s_base = total ? atomicAdd(count, total) : 0u;
    __syncthreads();
    const uint32_t off = s_base + woff + incl - cnt;
    if (total) {
      const uint32_t slot0 = (uint32_t)((tile * kCompactThreads + warp * 32) << 5);
      for (int j = 0; j < 32; ++j) {
        const uint32_t ej = __shfl_sync(0xffffffffu, e, j);
        const uint32_t oj = __shfl_sync(0xffffffffu, off, j);
        if ((ej >> lane) & 1u) list[oj + __popc(ej & ((1u << lane) - 1u))] = slot0 + (uint32_t)(j << 5) + lane;
      }
    }
    __syncthreads();  // s_warp / s_base are reused by the next tile
  }
}

__global__ void route_kernel(const RouteParams p) {
  griddep_wait();
  const double scan_s = gather_scan_seconds((double)*p.elig_count, p.row_bytes, p.passes);
  const bool scan = !(p.tensor_s < 0.9 * scan_s);
  for (int i = threadIdx.x; i < p.q; i += blockDim.x)
    if (scan) p.fb_list[i] = (uint32_t)i;
  if (threadIdx.x == 0) {
    *p.elig_out = *p.elig_count + 1u;
    *p.skip = scan ? 1u : 0u;
    *p.tensor_nq = scan ? 0u : (uint32_t)p.q;
    if (scan) {
      *p.fb_count = (uint32_t)p.q;
      *p.routed_scan = (uint32_t)p.q;
    }
  }
}

}  // namespace

cudaError_t launch_route(const RouteParams& p, cudaStream_t st) {
  return launch_pdl(route_kernel, dim3(1), dim3(32), 0, st, p);
}

cudaError_t launch_compact_eligible(const IndexView& iv, const MaskView& mask, uint32_t* list, uint32_t* count,
                                    cudaStream_t st) {
  if (iv.n_slots == 0) return cudaSuccess;
  const int64_t tiles = (iv.n_slots + 32 * kCompactThreads - 1) / (32 * kCompactThreads);
  const int blocks = (int)std::min<int64_t>(tiles, 148 * 8);
  if (iv.ids_identity || !mask.bits)
    return launch_pdl(compact_eligible_kernel<true>, dim3(blocks), dim3(kCompactThreads), 0, st, iv, mask, list, count);
  return launch_pdl(compact_eligible_kernel<false>, dim3(blocks), dim3(kCompactThreads), 0, st, iv, mask, list, count);
}

size_t scan_fused_tail_bytes(int grid, int K, int dpad) {
  return 2 * (size_t)K * 8 + std::max<size_t>(2 * (size_t)grid * K * 8, (size_t)(kScanThreads / 32) * dpad * 4);
}

size_t scan_smem_bytes(int QT, int dpad, int K, int nstages, int stage_floats) {
  return (size_t)nstages * stage_floats * 4 + (size_t)QT * dpad * 4 + (size_t)QT * NW * K * 8 +
         2 * kScanMaxStages * 8 + 4 * 8;  // barriers + the fused tail's control words
}

cudaError_t launch_scan(const ScanParams& p, int QT, int grid, cudaStream_t st) {
  const size_t smem = scan_smem_bytes(QT, p.iv.dpad, p.K, p.nstages, p.stage_floats);
  if (p.rows_per_stage > kMaxRounds * NW * (32 / p.lanes_per_row)) return cudaErrorInvalidValue;
  if (p.iv.n_slots >= 0xFFFFFFF0ll) return cudaErrorInvalidValue;
  switch (p.iv.metric) {
    case kMetricL2: return launch_scan_shape<kMetricL2>(p, QT, grid, smem, st);
    case kMetricCos: return launch_scan_shape<kMetricCos>(p, QT, grid, smem, st);
    case kMetricDot: return launch_scan_shape<kMetricDot>(p, QT, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace gfi
