// gemm_topk.cu -- K2 `flat_gemm_topk`: the batched-search path as a real Q.X^T contraction on
// the 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM), with
// the distance epilogue, the eligibility mask and the candidate filter fused so that no
// n x q distance matrix ever reaches HBM.
//
// Replaces, for large query batches, the q independent passes of VectorStore::search_batch
// (src/storage.rs:302-310) over FlatIndex::search (src/flat_index.rs:52-65): each database
// tile is read from HBM once and reused for every query of the batch.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: 2-D tiled bulk loads (SWIZZLE_128B) of a 128x64 query block and a
//               256x64 row block per k-step into a 4-stage shared-memory ring (mbarrier full/empty)
//   warp 1      MMA issuer: one lane issues 4 x tcgen05.mma (M=128,N=256,K=16) per k-step into one
//               of two 256-column TMEM accumulators; tcgen05.commit releases ring slots and
//               publishes finished accumulators
//   warps 2-5   epilogue: thread == query.  tcgen05.ld 32 columns at a time, one FMA per value
//               (score = acc * a[row] + b[row], which covers L2 / cosine / dot, per-row fp16 scale,
//               tombstones and the filter mask via b = +inf), compare against the query's
//               threshold, and append the rare survivors to the query's candidate list.
// The thresholds come from a seed pass of the same kernel over an evenly strided sample of
// tiles (seed_mode = 1).  Scores are approximate (fp16 inputs); select_rerank.cu re-scores the
// best candidates with the reference's exact arithmetic and certifies the result.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace gfi {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2;  // 16 KB
constexpr int kBBytes = BN * BK * 2;  // 32 KB
constexpr int kGemmThreads = 192;
constexpr int kEpiThreads = 128;
constexpr uint32_t kTmemCols = 512;

constexpr size_t kOffA = 0;
constexpr size_t kOffB = kOffA + (size_t)kStages * kABytes;
constexpr size_t kOffCoef = kOffB + (size_t)kStages * kBBytes;
constexpr size_t kOffBar = kOffCoef + 2 * BN * sizeof(float2);
constexpr size_t kOffCnt = kOffBar + 16 * 8 + 16;  // u8 hit counters, one per query of the batch
constexpr size_t kSmemUsed = kOffCnt + 2 * kGemmMaxQueries;
constexpr size_t kSmemBytes = kSmemUsed + 1024;  // slack for manual 1024-byte alignment

// UMMA instruction descriptor, kind::f16: D=f32, A=B=f16, both K-major, N=256, M=128.
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) |
                            ((uint32_t)(BM >> 4) << 24);

// UMMA shared-memory descriptor for a K-major SWIZZLE_128B tile (rows of 128 bytes, 8-row groups
// 1024 bytes apart): start>>4 | LBO(16B, unused)<<16 | SBO(1024B)<<32 | version 1<<46 | SW128 (2)<<61.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

__device__ __forceinline__ float pow2_scale_inv(float maxabs) {
  // inverse of the power-of-two scale chosen in ingest.cu / convert_queries16_kernel
  if (!(maxabs > 0.f) || !(maxabs <= 3.4028234664e38f)) return 1.f;
  int e = (int)((__float_as_uint(maxabs) >> 23) & 0xffu);
  e = e == 0 ? -126 : e - 127;
  const int se = min(max(14 - e, -100), 100);
  return __uint_as_float((uint32_t)(127 - se) << 23);
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmq,
                 const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  unsigned char* sA = smem + kOffA;
  unsigned char* sB = smem + kOffB;
  float2* sCoef = reinterpret_cast<float2*>(smem + kOffCoef);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  // u16 hit counter per query, private to this CTA: no atomics on the hot path.  Items map to
  // (row tile, query tile) with a rotation so that every CTA meets every query tile and a query's
  // candidates spread evenly over all CTAs' slices.
  unsigned short* hitcnt = reinterpret_cast<unsigned short*>(smem + kOffCnt);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const IndexView& iv = p.iv;
  const int num_kb = (iv.dpad16 + BK - 1) / BK;
  const int64_t n_items = (p.seed_mode == 1 ? p.seed_tiles : p.num_n_tiles) * p.num_m_tiles;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiThreads);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < p.num_m_tiles * BM; i += kGemmThreads) hitcnt[i] = 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmq);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    uint32_t it = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int64_t nt_idx = w / p.num_m_tiles;
      const int m_tile = (int)((w - nt_idx * p.num_m_tiles + nt_idx) % p.num_m_tiles);  // rotated: see item_tiles note
      const int64_t n_tile = p.seed_mode == 1 ? nt_idx * p.seed_stride : nt_idx;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        if (lane == 0) {
          // p.debug (timing experiments only, results invalid): bit0 / bit1 stop re-loading A / B once the
          // ring has been filled once, to separate the load path from the MMA and epilogue cost.
          const bool ldA = !((p.debug & 1) && it >= (uint32_t)kStages);
          const bool ldB = !((p.debug & 2) && it >= (uint32_t)kStages);
          if (ldA || ldB) mbar_arrive_expect_tx(&full[s], (ldA ? kABytes : 0) + (ldB ? kBBytes : 0));
          else mbar_arrive(&full[s]);
          if (ldA) tma_load_2d(sA + (size_t)s * kABytes, &tmq, kb * BK, m_tile * BM, &full[s]);
          if (ldB) tma_load_2d(sB + (size_t)s * kBBytes, &tmx, kb * BK, (int)(n_tile * BN), &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    uint32_t it = 0, ai = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      mbar_wait(&tempty[as], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(sA + (size_t)s * kABytes);
          const uint32_t b0 = smem_u32(sB + (size_t)s * kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_f16(d_tmem, make_sw128_desc(a0 + k * 32), make_sw128_desc(b0 + k * 32), kIdesc,
                     (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);                       // ring slot free once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull[as]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue: thread == query ------------------------------
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int et = (warp - 2) * 32 + lane;    // 0..127 among epilogue threads
    const int mrow = quarter * 32 + lane;     // row of the 128-query tile owned by this thread
    const float inv_sq = pow2_scale_inv(*p.qmaxabs);
    const float kInf = __int_as_float(0x7f800000);

    // per-row epilogue coefficients (a, b) of rows et and et+128 of an item's tile; ineligible rows
    // (beyond the index, tombstoned, filtered out) get (0, +inf) so they can never pass a threshold
    auto load_coef = [&](int64_t w, float2 (&c)[2]) {
      const int64_t nt_idx = w / p.num_m_tiles;
      const int64_t n_tile = p.seed_mode == 1 ? nt_idx * p.seed_stride : nt_idx;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int64_t slot = n_tile * BN + et + rr * kEpiThreads;
        c[rr] = make_float2(0.f, kInf);
        if (slot < iv.n_slots) {
          bool elig = (iv.live[slot >> 5] >> (slot & 31)) & 1u;
          if (elig && p.mask.bits) {
            const uint64_t id = iv.ids_identity ? (uint64_t)slot : iv.ids[slot];
            elig = (id < (uint64_t)p.mask.nbits) && ((p.mask.bits[id >> 6] >> (id & 63)) & 1ull);
          }
          if (elig) {
            c[rr] = iv.coef[slot];
            c[rr].x *= inv_sq;
          }
        }
      }
    };

    float2 cnext[2];
    if ((int64_t)blockIdx.x < n_items) load_coef(blockIdx.x, cnext);
    uint32_t ai = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      const int64_t nt_idx = w / p.num_m_tiles;
      const int m_tile = (int)((w - nt_idx * p.num_m_tiles + nt_idx) % p.num_m_tiles);  // rotated: see item_tiles note
      const int64_t n_tile = p.seed_mode == 1 ? nt_idx * p.seed_stride : nt_idx;
      const int64_t n0 = n_tile * BN;
      const int qidx = m_tile * BM + mrow;
      float2* cs = sCoef + as * BN;
      cs[et] = cnext[0];
      cs[et + kEpiThreads] = cnext[1];
      named_bar_sync(2, kEpiThreads);
      // the next item's coefficient loads stay in flight while this item is processed
      if (w + gridDim.x < n_items) load_coef(w + gridDim.x, cnext);

      float thr = __int_as_float(0xff800000);  // -inf: padding queries never match
      if (p.seed_mode == 0 && qidx < p.q) thr = p.thresh[qidx];
      float sd[kSeedR];
#pragma unroll
      for (int i = 0; i < kSeedR; ++i) sd[i] = kInf;

      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < ((p.debug & 4) ? 0 : BN); c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        // 32 (a,b) pairs as 16 broadcast 128-bit shared loads, issued while the TMEM load is in flight
        float4 cf[16];
        const float4* c4 = reinterpret_cast<const float4*>(cs + c0);
#pragma unroll
        for (int i = 0; i < 16; ++i) cf[i] = c4[i];
        tmem_ld_wait();
        float sc[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          sc[2 * i] = fmaf(__uint_as_float(r[2 * i]), cf[i].x, cf[i].y);
          sc[2 * i + 1] = fmaf(__uint_as_float(r[2 * i + 1]), cf[i].z, cf[i].w);
        }
        if (p.seed_mode == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (sc[j] < sd[kSeedR - 1]) {
              sd[kSeedR - 1] = sc[j];
#pragma unroll
              for (int i = kSeedR - 1; i > 0; --i) {
                const float lo = fminf(sd[i - 1], sd[i]), hi = fmaxf(sd[i - 1], sd[i]);
                sd[i - 1] = lo;
                sd[i] = hi;
              }
            }
          }
        } else if (p.seed_mode == 2) {
          // debug dump of every approximate score (tests only; small inputs)
          if (qidx < p.q) {
#pragma unroll
            for (int j = 0; j < 32; ++j) p.seeds[(size_t)qidx * p.seed_stride + (size_t)(n0 + c0 + j)] = sc[j];
          }
        } else {
          // common case: no value of the block beats the threshold -> one min tree + one compare.
          // (fminf drops NaN operands; a NaN query makes every score NaN, which still reaches the slow path.)
          float m[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) m[i] = fminf(sc[2 * i], sc[2 * i + 1]);
#pragma unroll
          for (int w2 = 8; w2 > 0; w2 >>= 1) {
#pragma unroll
            for (int i = 0; i < w2; ++i) m[i] = fminf(m[i], m[i + w2]);
          }
          if (!(m[0] >= thr) && qidx < p.q && !(p.debug & 8)) {
            // rare: append to this (query, CTA)'s private slice of the candidate buffer -- plain
            // stores, no atomics, nothing to wait for
            uint32_t cnt = hitcnt[qidx];
            uint64_t* mine = p.cand + (size_t)qidx * p.cand_stride + (size_t)blockIdx.x * p.cand_cap;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float score = sc[j];
              if (!(score >= thr)) {
                if (score != score) {
                  atomicOr(p.flags, kFlagNaN);
                } else {
                  if (cnt < p.cand_cap) mine[cnt] = pack_key(score, (uint32_t)(n0 + c0 + j));
                  else p.cand_cnt[qidx] = 0xffffffffu;  // overflow marker: this query falls back to the scan
                  ++cnt;
                }
              }
            }
            hitcnt[qidx] = (unsigned short)min(cnt, 65535u);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
      if (p.seed_mode == 1 && qidx < p.q) {
        float* out = p.seeds + ((size_t)qidx * p.seed_tiles + nt_idx) * kSeedR;
#pragma unroll
        for (int i = 0; i < kSeedR; ++i) out[i] = sd[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// One warp per query: the rank-th smallest of the query's seed scores becomes its threshold.
__global__ void seed_finalize_kernel(const SeedFinalizeParams p) {
  const int lane = threadIdx.x & 31;
  const int qi = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (qi >= p.q) return;
  const int64_t total = p.seed_tiles * kSeedR;
  const float* s = p.seeds + (size_t)qi * total;
  uint64_t prev = 0;
  uint64_t cur = ~0ull;
  for (int r = 0; r < p.rank; ++r) {
    uint64_t best = ~0ull;
    for (int64_t i = lane; i < total; i += 32) {
      const float v = s[i];
      if (v != v) continue;
      const uint64_t key = ((uint64_t)f32_key(v) << 32) | (uint32_t)(i + 1);
      if (key > prev && key < best) best = key;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    cur = best;
    if (best == ~0ull) break;
    prev = best;
  }
  if (lane == 0) p.thresh[qi] = (cur == ~0ull) ? __int_as_float(0x7f800000) : key_f32((uint32_t)(cur >> 32));
}

}  // namespace

size_t gemm_smem_bytes() { return kSmemBytes; }

cudaError_t launch_gemm_topk(const GemmParams& p, const void* tmap_x_host, const void* tmap_q_host, int grid,
                             cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  cudaError_t e =
      cudaFuncSetAttribute(gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return e;
  const CUtensorMap* tx = reinterpret_cast<const CUtensorMap*>(tmap_x_host);
  const CUtensorMap* tq = reinterpret_cast<const CUtensorMap*>(tmap_q_host);
  gemm_topk_kernel<<<grid, kGemmThreads, kSmemBytes, st>>>(*tx, *tq, p);
  return cudaGetLastError();
}

cudaError_t launch_seed_finalize(const SeedFinalizeParams& p, cudaStream_t st) {
  if (p.q <= 0) return cudaSuccess;
  const int blocks = (p.q * 32 + 255) / 256;
  seed_finalize_kernel<<<blocks, 256, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace gfi
