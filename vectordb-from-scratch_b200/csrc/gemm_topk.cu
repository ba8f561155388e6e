// gemm_topk.cu -- K2 `flat_gemm_topk`: the batched-search path as a real Q.X^T contraction on
// the 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM), with
// the distance epilogue, the eligibility mask and the candidate filter fused so that no
// n x q distance matrix ever reaches HBM.
//
// Replaces, for large query batches, the q independent passes of VectorStore::search_batch
// (src/storage.rs:302-310) over FlatIndex::search (src/flat_index.rs:52-65): each database
// tile is read from HBM once and reused for every query of the batch.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: 2-D tiled bulk loads (SWIZZLE_128B) of a 128x64 query block and a
//               256x64 row block per k-step into a 4-stage shared-memory ring (mbarrier full/empty)
//   warp 1      MMA issuer: one lane issues 4 x tcgen05.mma (M=128,N=256,K=16) per k-step into one
//               of two 256-column TMEM accumulators; tcgen05.commit releases ring slots and
//               publishes finished accumulators
//   warp 2      coefficient stager: per row of the tile the epilogue pair (a, b) -- metric, per-row
//               fp16 scale, tombstone and filter bit (b = +inf) -- loaded one item ahead and
//               published through an mbarrier, so the epilogue never waits on global memory
//   warps 4-11  epilogue, thread == (query, half of the tile's columns): tcgen05.ld 32 columns at
//               a time, one FMA per value (score = acc * a[row] + b[row]), a min tree and ONE
//               compare per 32 values against the query's threshold; the rare survivors are
//               appended with plain stores to a slice of the candidate buffer that is private to
//               this (query, CTA, half) -- no atomics.
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
constexpr int kGemmThreads = 384;
constexpr int kEpiWarp0 = 4;          // first epilogue warp
constexpr int kEpiThreads = 256;
constexpr int kEpiWarps = kEpiThreads / 32;
constexpr uint32_t kTmemCols = 512;

constexpr size_t kOffA = 0;
constexpr size_t kOffB = kOffA + (size_t)kStages * kABytes;
constexpr size_t kOffCoef = kOffB + (size_t)kStages * kBBytes;
constexpr size_t kOffBar = kOffCoef + 2 * BN * sizeof(float2);
constexpr size_t kOffCnt = kOffBar + 16 * 8 + 16;  // u16 hit counters [2 halves][queries]
constexpr size_t kSmemUsed = kOffCnt + 2 * 2 * kGemmMaxQueries;
constexpr size_t kSmemBytes = kSmemUsed + 1024;  // slack for manual 1024-byte alignment
static_assert(kSmemBytes <= 232448, "shared memory budget");

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

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// item -> (row tile, query tile).  The query tile is rotated by the row-tile index so that every
// CTA meets every query tile: a query's candidates then spread evenly over all CTAs' slices.
template <int MODE>
__device__ __forceinline__ void item_tiles(const GemmParams& p, int64_t w, int64_t& nt_idx, int64_t& n_tile,
                                           int& m_tile) {
  nt_idx = w / p.num_m_tiles;
  m_tile = (int)((w - nt_idx * p.num_m_tiles + nt_idx) % p.num_m_tiles);
  n_tile = MODE == 1 ? nt_idx * p.seed_stride : nt_idx;
}

// MODE 0: main pass (threshold filter, per-row coefficients), 1: seed pass (per-thread smallest scores),
// 2: debug dump of all scores, 3: main pass with the raw epilogue (cosine, no mask).
// One instance per mode keeps the hot instance's code small: the epilogue is sensitive to instruction-cache
// misses (a 4x larger unrolled epilogue ran 3x slower).
template <int MODE>
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
  uint64_t* tfull = empty + kStages;   // accumulator stage complete (MMA -> epilogue)
  uint64_t* tempty = tfull + 2;        // accumulator + coefficient stage drained (epilogue -> MMA, stager)
  uint64_t* cfull = tempty + 2;        // coefficient stage published (stager -> epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cfull + 2);
  unsigned short* hitcnt = reinterpret_cast<unsigned short*>(smem + kOffCnt);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const IndexView& iv = p.iv;
  const int num_kb = (iv.dpad16 + BK - 1) / BK;
  const int64_t n_items = (MODE == 1 ? p.seed_tiles : p.num_n_tiles) * p.num_m_tiles;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiWarps);
      mbar_init(&cfull[s], 1);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < 2 * kGemmMaxQueries; i += kGemmThreads) hitcnt[i] = 0;
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
    int s = 0;
    uint32_t ph = 0, it = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE>(p, w, nt_idx, n_tile, m_tile);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
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
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // Convergent code: all lanes run the loop with identical (uniform) operands and one elected lane issues
    // inside the asm, so the descriptors live in uniform registers and a k-step is a few dozen instructions.
    // Descriptor low words: (addr >> 4) | LBO<<16; a stage is 16 KB (A) / 32 KB (B) further, a K step 32 bytes.
    const uint32_t a_lo_base = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo_base = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version 1 | SWIZZLE_128B
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull);
    int s = 0;
    uint32_t ph = 0, ai = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      mbar_wait(&tempty[as], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo_base + (uint32_t)s * (kABytes >> 4);
        const uint32_t b_lo = b_lo_base + (uint32_t)s * (kBBytes >> 4);
        if (!(p.debug & 16)) {  // bit4 (timing experiments): barrier handshakes only, no MMA
          umma_f16_elect(d_tmem, a_lo, b_lo, desc_hi, kIdesc, kb != 0 ? 1u : 0u);
          umma_f16_elect(d_tmem, a_lo + 2, b_lo + 2, desc_hi, kIdesc, 1u);
          umma_f16_elect(d_tmem, a_lo + 4, b_lo + 4, desc_hi, kIdesc, 1u);
          umma_f16_elect(d_tmem, a_lo + 6, b_lo + 6, desc_hi, kIdesc, 1u);
        }
        umma_commit_elect(empty0 + s * 8);                           // ring slot free once these MMAs retire
        if (kb == num_kb - 1) umma_commit_elect(tfull0 + as * 8);    // accumulator complete
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
    (void)full0;
  } else if (warp == 2 && MODE != 3) {
    // ------------------------------ coefficient stager ------------------------------
    // lane owns rows lane, lane+32, ... of the tile.  Raw loads for the NEXT item are issued before
    // the current item's values are consumed, so nothing here waits on memory in steady state.
    constexpr int RPL = BN / 32;  // rows per lane
    const float inv_sq = pow2_scale_inv(*p.qmaxabs);
    const float kInf = __int_as_float(0x7f800000);
    const bool has_mask = p.mask.bits != nullptr;
    const bool by_slot = has_mask && iv.ids_identity;
    struct Raw { float2 c[RPL]; uint32_t live[RPL]; uint64_t mask[RPL]; };
    auto fetch = [&](int64_t w, Raw& r) {
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE>(p, w, nt_idx, n_tile, m_tile);
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int64_t slot = n_tile * BN + lane + 32 * i;
        r.c[i] = make_float2(0.f, kInf);
        r.live[i] = 0;
        r.mask[i] = ~0ull;
        if (slot < iv.n_slots) {
          r.c[i] = __ldg(iv.coef + slot);
          r.live[i] = __ldg(iv.live + (slot >> 5));
          if (by_slot) r.mask[i] = slot < p.mask.nbits ? __ldg(p.mask.bits + (slot >> 6)) : 0ull;
        }
      }
    };
    Raw nxt;
    if ((int64_t)blockIdx.x < n_items) fetch(blockIdx.x, nxt);
    uint32_t ai = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE>(p, w, nt_idx, n_tile, m_tile);
      const Raw cur = nxt;
      if (w + gridDim.x < n_items) fetch(w + gridDim.x, nxt);
      mbar_wait(&tempty[as], aph ^ 1u);  // the epilogue has drained the previous use of this stage
      float2* cs = sCoef + as * BN;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int64_t slot = n_tile * BN + lane + 32 * i;
        bool elig = ((cur.live[i] >> (slot & 31)) & 1u) && ((cur.mask[i] >> (slot & 63)) & 1ull);
        if (elig && has_mask && !by_slot) {
          const uint64_t id = iv.ids[slot];
          elig = (id < (uint64_t)p.mask.nbits) && ((p.mask.bits[id >> 6] >> (id & 63)) & 1ull);
        }
        cs[lane + 32 * i] = elig ? make_float2(cur.c[i].x * inv_sq, cur.c[i].y) : make_float2(0.f, kInf);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&cfull[as]);
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------ epilogue: thread == (query, column half) ------------------------------
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;     // which 128 columns of the 256-column tile
    const int mrow = quarter * 32 + lane;         // row of the 128-query tile owned by this thread
    const float kInf = __int_as_float(0x7f800000);
    // raw mode: score = acc * c_q, c_q = -(1 / query scale) * 2^-14 (row scale), exact powers of two
    const float c_q = MODE == 3 ? -pow2_scale_inv(*p.qmaxabs) * 6.103515625e-05f : 0.f;
    uint32_t ai = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE>(p, w, nt_idx, n_tile, m_tile);
      const int64_t n0 = n_tile * BN;
      const int qidx = m_tile * BM + mrow;
      float thr = __int_as_float(0xff800000);  // -inf: padding queries never match
      if ((MODE == 0 || MODE == 3) && qidx < p.q) thr = p.thresh[qidx];
      float sd[kSeedR];
#pragma unroll
      for (int i = 0; i < kSeedR; ++i) sd[i] = kInf;

      // one lane polls, the warp follows: 32x fewer try_wait probes competing with the MMA / TMA warps' own
      // barrier traffic (acquire by lane 0 + __syncwarp orders the other lanes' reads)
      if (MODE != 3) mbar_wait(&cfull[as], aph);
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + half * (BN / 2);
      const uint32_t cs_addr = smem_u32(sCoef + as * BN + half * (BN / 2));
      if (MODE == 3) {
        // Raw epilogue (cosine rows are stored pre-normalised, so score = acc * c_q with one per-batch constant):
        // the filter compares accumulators with thr / c_q directly -- no per-row coefficients, 3-input max trees.
        // Two 32-column loads are kept in flight; tombstoned / out-of-range rows are weeded out in the rare path.
        const float thr_raw = __fdiv_rn(thr, c_q);  // exact: c_q is a (negative) power of two
        auto process = [&](const uint32_t (&r)[32], const int c0) {
          float m8[4];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const float a = fmax3(__uint_as_float(r[8 * g8]), __uint_as_float(r[8 * g8 + 1]), __uint_as_float(r[8 * g8 + 2]));
            const float b = fmax3(__uint_as_float(r[8 * g8 + 3]), __uint_as_float(r[8 * g8 + 4]), __uint_as_float(r[8 * g8 + 5]));
            m8[g8] = fmax3(a, b, fmaxf(__uint_as_float(r[8 * g8 + 6]), __uint_as_float(r[8 * g8 + 7])));
          }
          const float mall = fmax3(m8[0], m8[1], fmaxf(m8[2], m8[3]));
          if (__any_sync(0xffffffffu, !(mall <= thr_raw) && qidx < p.q) && !(p.debug & 8)) {
#pragma unroll 1
            for (int g8 = 0; g8 < 4; ++g8) {
              const float mg = g8 == 0 ? m8[0] : (g8 == 1 ? m8[1] : (g8 == 2 ? m8[2] : m8[3]));
              const bool mine = !(mg <= thr_raw) && qidx < p.q;
              if (!__any_sync(0xffffffffu, mine)) continue;
              uint32_t v[8];
              tmem_ld_32x32b_x8(taddr + c0 + 8 * g8, v);
              const int64_t slot0 = n0 + half * (BN / 2) + c0 + 8 * g8;  // multiple of 8: one word of live bits
              uint32_t lv = 0;
              if (mine && slot0 < iv.n_slots) lv = (__ldg(iv.live + (slot0 >> 5)) >> (slot0 & 31)) & 0xffu;
              tmem_ld_wait();
              if (mine) {
                unsigned short* hc = hitcnt + half * kGemmMaxQueries + qidx;
                uint32_t cnt = *hc;
                uint64_t* slice = p.cand + (size_t)qidx * p.cand_stride + (size_t)(blockIdx.x * 2 + half) * p.cand_cap;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const float acc = __uint_as_float(v[jj]);
                  if (!(acc <= thr_raw) && ((lv >> jj) & 1u) && slot0 + jj < iv.n_slots) {
                    if (acc != acc) {
                      atomicOr(p.flags, kFlagNaN);
                    } else {
                      if (cnt < p.cand_cap) slice[cnt] = pack_key(acc * c_q, (uint32_t)(slot0 + jj));
                      else p.cand_cnt[qidx] = 0xffffffffu;
                      ++cnt;
                    }
                  }
                }
                *hc = (unsigned short)min(cnt, 65535u);
              }
            }
          }
        };
        if (!(p.debug & 4)) {
          uint32_t ra[32], rb[32];
          tmem_ld_32x32b_x32(taddr, ra);
          tmem_ld_wait();
          tmem_ld_32x32b_x32(taddr + 32, rb);
          process(ra, 0);
          tmem_ld_wait();
          tmem_ld_32x32b_x32(taddr + 64, ra);
          process(rb, 32);
          tmem_ld_wait();
          tmem_ld_32x32b_x32(taddr + 96, rb);
          process(ra, 64);
          tmem_ld_wait();
          process(rb, 96);
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < ((MODE == 3 || (p.debug & 4)) ? 0 : BN / 2); c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        // 32 (a,b) pairs as 16 broadcast 128-bit shared loads, issued while the TMEM load is in flight
        float4 cf[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cf[i] = lds128(cs_addr + (c0 + 2 * i) * 8);
        tmem_ld_wait();
        float sc[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          sc[2 * i] = fmaf(__uint_as_float(r[2 * i]), cf[i].x, cf[i].y);
          sc[2 * i + 1] = fmaf(__uint_as_float(r[2 * i + 1]), cf[i].z, cf[i].w);
        }
        const int col0 = half * (BN / 2) + c0;
        if (MODE == 1) {
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
        } else if (MODE == 2) {
          // debug dump of every approximate score (tests only; small inputs)
          if (qidx < p.q) {
#pragma unroll
            for (int j = 0; j < 32; ++j) p.seeds[(size_t)qidx * p.seed_stride + (size_t)(n0 + col0 + j)] = sc[j];
          }
        } else {
          // common case: no value of the block beats the threshold -> one min tree + one compare.
          // (fminf drops NaN operands; a NaN query makes every score NaN, which still reaches the slow path.)
          float m8[4];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const float a = fminf(fminf(sc[8 * g8], sc[8 * g8 + 1]), fminf(sc[8 * g8 + 2], sc[8 * g8 + 3]));
            const float b = fminf(fminf(sc[8 * g8 + 4], sc[8 * g8 + 5]), fminf(sc[8 * g8 + 6], sc[8 * g8 + 7]));
            m8[g8] = fminf(a, b);
          }
          const float mall = fminf(fminf(m8[0], m8[1]), fminf(m8[2], m8[3]));
          // Rare path, kept SMALL and warp-uniform: if any lane of the warp has a survivor in this block, the
          // warp re-reads the 8-column groups concerned from TMEM (tcgen05.ld is warp-collective) and each lane
          // appends its own survivors to the candidate slice private to this (query, CTA, half) -- plain
          // stores, no atomics.
          if (__any_sync(0xffffffffu, !(mall >= thr) && qidx < p.q) && !(p.debug & 8)) {
#pragma unroll 1
            for (int g8 = 0; g8 < 4; ++g8) {
              const float mg = g8 == 0 ? m8[0] : (g8 == 1 ? m8[1] : (g8 == 2 ? m8[2] : m8[3]));
              const bool mine = !(mg >= thr) && qidx < p.q;
              if (!__any_sync(0xffffffffu, mine)) continue;
              uint32_t v[8];
              tmem_ld_32x32b_x8(taddr + c0 + 8 * g8, v);
              const float4 k0 = lds128(cs_addr + (c0 + 8 * g8) * 8), k1 = lds128(cs_addr + (c0 + 8 * g8 + 2) * 8);
              const float4 k2 = lds128(cs_addr + (c0 + 8 * g8 + 4) * 8), k3 = lds128(cs_addr + (c0 + 8 * g8 + 6) * 8);
              tmem_ld_wait();
              if (mine) {
                float s8[8];
                s8[0] = fmaf(__uint_as_float(v[0]), k0.x, k0.y); s8[1] = fmaf(__uint_as_float(v[1]), k0.z, k0.w);
                s8[2] = fmaf(__uint_as_float(v[2]), k1.x, k1.y); s8[3] = fmaf(__uint_as_float(v[3]), k1.z, k1.w);
                s8[4] = fmaf(__uint_as_float(v[4]), k2.x, k2.y); s8[5] = fmaf(__uint_as_float(v[5]), k2.z, k2.w);
                s8[6] = fmaf(__uint_as_float(v[6]), k3.x, k3.y); s8[7] = fmaf(__uint_as_float(v[7]), k3.z, k3.w);
                unsigned short* hc = hitcnt + half * kGemmMaxQueries + qidx;
                uint32_t cnt = *hc;
                uint64_t* slice = p.cand + (size_t)qidx * p.cand_stride + (size_t)(blockIdx.x * 2 + half) * p.cand_cap;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const float score = s8[jj];
                  if (!(score >= thr)) {
                    if (score != score) {
                      atomicOr(p.flags, kFlagNaN);
                    } else {
                      if (cnt < p.cand_cap) slice[cnt] = pack_key(score, (uint32_t)(n0 + col0 + 8 * g8 + jj));
                      else p.cand_cnt[qidx] = 0xffffffffu;  // overflow marker: the query falls back to the scan
                      ++cnt;
                    }
                  }
                }
                *hc = (unsigned short)min(cnt, 65535u);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);  // one arrival per warp
      if (MODE == 1 && qidx < p.q) {
        float* out = p.seeds + (((size_t)qidx * p.seed_tiles + nt_idx) * 2 + half) * kSeedR;
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
// Each lane keeps the kSeedR smallest of its strided share in registers (single pass), then the warp
// pops the global minimum `rank` times.
__global__ void seed_finalize_kernel(const SeedFinalizeParams p) {
  const int lane = threadIdx.x & 31;
  const int qi = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (qi >= p.q) return;
  const float kInf = __int_as_float(0x7f800000);
  const int64_t total = p.seed_tiles * 2 * kSeedR;
  const float* s = p.seeds + (size_t)qi * total;
  float sd[kSeedR];
#pragma unroll
  for (int i = 0; i < kSeedR; ++i) sd[i] = kInf;
  for (int64_t i = lane; i < total; i += 32) {
    const float v = s[i];
    if (v < sd[kSeedR - 1]) {  // false for NaN
      sd[kSeedR - 1] = v;
#pragma unroll
      for (int j = kSeedR - 1; j > 0; --j) {
        const float lo = fminf(sd[j - 1], sd[j]), hi = fmaxf(sd[j - 1], sd[j]);
        sd[j - 1] = lo;
        sd[j] = hi;
      }
    }
  }
  float cur = kInf;
  for (int r = 0; r < p.rank; ++r) {
    float m = sd[0];
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    cur = m;
    const unsigned owners = __ballot_sync(0xffffffffu, sd[0] == m);
    if (owners && lane == __ffs(owners) - 1) {  // pop this lane's head
#pragma unroll
      for (int j = 0; j < kSeedR - 1; ++j) sd[j] = sd[j + 1];
      sd[kSeedR - 1] = kInf;
    }
  }
  if (lane == 0) p.thresh[qi] = cur;
}

}  // namespace

cudaError_t launch_gemm_topk(const GemmParams& p, const void* tmap_x_host, const void* tmap_q_host, int grid,
                             cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  if (p.num_m_tiles * BM > kGemmMaxQueries) return cudaErrorInvalidValue;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 15]) {  // once per device: the driver call is not free
    cudaError_t e = cudaFuncSetAttribute(gemm_topk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev & 15] = true;
  }
  const CUtensorMap* tx = reinterpret_cast<const CUtensorMap*>(tmap_x_host);
  const CUtensorMap* tq = reinterpret_cast<const CUtensorMap*>(tmap_q_host);
  if (p.seed_mode == 0) gemm_topk_kernel<0><<<grid, kGemmThreads, kSmemBytes, st>>>(*tx, *tq, p);
  else if (p.seed_mode == 1) gemm_topk_kernel<1><<<grid, kGemmThreads, kSmemBytes, st>>>(*tx, *tq, p);
  else if (p.seed_mode == 3) gemm_topk_kernel<3><<<grid, kGemmThreads, kSmemBytes, st>>>(*tx, *tq, p);
  else gemm_topk_kernel<2><<<grid, kGemmThreads, kSmemBytes, st>>>(*tx, *tq, p);
  return cudaGetLastError();
}

cudaError_t launch_seed_finalize(const SeedFinalizeParams& p, cudaStream_t st) {
  if (p.q <= 0) return cudaSuccess;
  const int blocks = (p.q * 32 + 255) / 256;
  seed_finalize_kernel<<<blocks, 256, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace gfi
