// gemm_topk.cu -- K2 `flat_gemm_topk`: the batched-search path as a real Q.X^T contraction on
// the 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM), with
// the distance epilogue, the eligibility mask and the candidate filter fused so that no
// n x q distance matrix ever reaches HBM.
//
// Replaces, for large query batches, the q independent passes of VectorStore::search_batch
// (src/storage.rs:302-310) over FlatIndex::search (src/flat_index.rs:52-65): each database
// tile is read from HBM once and reused for every query of the batch.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised; one kernel instance per MODE):
//   warp 0      TMA producer: 2-D tiled bulk loads (SWIZZLE_128B) of a 128x64 query block and a
//               256x64 row block per k-step into a 4-stage shared-memory ring (mbarrier full/empty)
//   warp 1      MMA issuer: one elected lane issues 4 x tcgen05.mma (M=128,N=256,K=16) per k-step into
//               one of two 256-column TMEM accumulators; tcgen05.commit releases ring slots and
//               publishes finished accumulators
//   warps 2-3   coefficient stagers (coefficient epilogue only): per row of the tile the pair (a, b) --
//               metric, per-row fp16 scale, tombstone and filter bit (b = +inf) -- loaded one item ahead
//               and published through an mbarrier, so the epilogue never waits on global memory
//   warps 4-11  epilogue, thread == (query, half of the tile's columns): two tcgen05.ld of 32 columns in
//               flight; coefficient mode: one FMA per value (score = acc * a[row] + b[row]) and 3-input
//               min trees; raw mode (cosine, rows stored normalised): 3-input max trees on the raw
//               accumulators; ONE compare per 32 values against the query's threshold.  The rare
//               survivors are appended with plain stores to a slice of the candidate buffer that is
//               private to this (query, unit, half) -- no atomics; per-slice counts are published at the
//               end of the kernel.
// A CTA-pair instance (tcgen05.mma.cta_group::2, clusters of two CTAs) exists behind an option; see Geo<>.
// The thresholds come from a seed pass of the same kernel over an evenly strided sample of
// tiles (seed_mode = 1).  Scores are approximate (fp16 inputs); select_rerank.cu re-scores the
// best candidates with the reference's exact arithmetic and certifies the result.
#include <cuda.h>
#include <cstdio>

#include "common.cuh"
#include "kernels.h"

namespace gfi {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kABytes = BM * BK * 2;  // 16 KB
constexpr int kGemmThreads = 384;
constexpr int kEpiWarp0 = 4;          // first epilogue warp
constexpr int kEpiThreads = 256;
constexpr int kEpiWarps = kEpiThreads / 32;
constexpr uint32_t kTmemCols = 512;

// CG = CTAs per MMA group.  CG == 1: one SM computes a 128 (queries) x 256 (rows) tile from a 16 KB query block
// and a 32 KB row block per k-step.  CG == 2: a CTA pair (two SMs of one TPC) issues tcgen05.mma.cta_group::2,
// M = 256: each CTA holds its own 128 queries and HALF of the row block (16 KB), the tensor cores fetch the other
// half from the peer's shared memory.  Per SM that cuts the bytes taken in through the L2->SM port from 48 to
// 32 KB per k-step and leaves room for 6 ring stages.  Measured on B200: 96.5% tensor-pipe
// utilisation per cycle without the epilogue (87.7% for CG == 1), but the board is power-limited and the SM clock
// drops accordingly -- same wall time -- so the launcher uses CG == 2 only on request.
template <int CG> struct Geo {
  static constexpr int kStages = CG == 2 ? 6 : 4;
  static constexpr int kBRows = BN / CG;               // rows of the row block held by one CTA
  static constexpr int kBBytes = kBRows * BK * 2;      // 32 KB / 16 KB
  static constexpr size_t kOffA = 0;
  static constexpr size_t kOffB = kOffA + (size_t)kStages * kABytes;
  static constexpr size_t kOffCoef = kOffB + (size_t)kStages * kBBytes;
  static constexpr size_t kOffBar = kOffCoef + 2 * BN * sizeof(float2);
  static constexpr size_t kOffCnt = kOffBar + 24 * 8 + 16;  // u16 hit counters [2 halves][queries]
  static constexpr size_t kSmemUsed = kOffCnt + 2 * 2 * kGemmMaxQueries;
  static constexpr size_t kSmemBytes = kSmemUsed + 1024;  // slack for manual 1024-byte alignment
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  // UMMA instruction descriptor, kind::f16: D=f32, A=B=f16, both K-major, N=256, M=128*CG.
  static constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                     ((uint32_t)((BM * CG) >> 4) << 24);
};

// UMMA shared-memory descriptors (K-major SWIZZLE_128B tiles: rows of 128 bytes, 8-row groups 1024 bytes apart) are
// built in the MMA issuer: start>>4 | LBO(16 B, unused)<<16 in the low word, SBO(1024 B)>>4 | version 1<<14 |
// SWIZZLE_128B (2)<<29 in the high word.

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

// Main-pass filter of one accumulator half (128 columns) for the calling thread's query.
//
// Hot loop (every item): 32 columns at a time with two TMEM loads in flight; per 32 values ONE compare against the
// query's threshold, whose outcome is one bit of a 4-bit block mask -- nothing else happens per block.
//  MODE 0: score = acc * a[row] + b[row] (coefficients as broadcast 128-bit shared loads), 3-input MIN trees.
//  MODE 3 (cosine, rows stored pre-normalised: score = acc * c_q with one per-batch constant): raw accumulators
//          against thr_raw = thr / c_q -- no coefficients, 3-input MAX trees; tombstoned / out-of-range rows are
//          weeded out in the survivor section with the live words `lv` the caller loaded once per row tile.
// Survivor section (epi_survivors, ONE copy of the code behind ONE warp vote per item): the blocks some lane flagged
// are re-read from TMEM and the flagged lanes append their survivors with plain stores to the candidate slice
// private to this (query, unit, half) -- no atomics.
// Round-2 profile of the previous form (a survivor path inlined after each of the four blocks, a vote per block, a
// global load of the live word and an 8-way branch ladder per hit group; profiles/r02_ncu_gemm_sk_c5_*.txt): a
// survivor-path entry cost ~1700 cycles and its four copies sat between the hot blocks; at C5's hit rate (one block
// in eight) that was a third of the epilogue's time and most of the spread between the eight epilogue warps, which is
// what the MMA issuer waits for.  Shared by the k-ring kernel and the short-K row-stationary kernel.
template <int MODE>
__device__ __noinline__ bool epi_survivors(uint64_t* slice, const uint32_t cap, uint32_t* flags, const uint32_t taddr,
                                           const uint32_t cs_addr, const float thr, const float thr_raw,
                                           const float c_q, const uint32_t slot_base, unsigned short* hc,
                                           const uint32_t bm, const uint32_t lv0, const uint32_t lv1,
                                           const uint32_t lv2, const uint32_t lv3, const uint32_t release_bar) {
  // (plain scalars only: a reference to the kernel's parameter block would force a copy of it into local memory)
  // The eight epilogue warps of a CTA release an accumulator together, so the LATENCY of this section on the one warp
  // that takes it sets the pace of the whole item: group maxima first, then only the 8-column groups that hold a hit.
  uint32_t wm = __reduce_or_sync(0xffffffffu, bm);
  uint32_t cnt = bm ? *hc : 0u;
  bool nan = false, released = false;
#pragma unroll 1
  while (wm) {
    const int blk = __ffs(wm) - 1;
    wm &= wm - 1;
    const int c0 = blk * 32;
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + c0, v);  // warp-collective: every lane takes part, flagged or not
    tmem_ld_wait();
    if (wm == 0 && release_bar) {
      // the last flagged block is in registers: this warp is done with the accumulator.  Releasing it HERE keeps
      // the scan of the block's values and the candidate stores off the path the MMA issuer waits on.
      tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(release_bar) : "memory");
      released = true;
    }
    if ((bm >> blk) & 1u) {
      const uint32_t lvw = blk == 0 ? lv0 : (blk == 1 ? lv1 : (blk == 2 ? lv2 : lv3));
      float s[32];
      if (MODE == 3) {
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float4 k = lds128(cs_addr + (c0 + j) * 8);
          s[j] = fmaf(__uint_as_float(v[j]), k.x, k.y);
          s[j + 1] = fmaf(__uint_as_float(v[j + 1]), k.z, k.w);
        }
      }
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const float* g = s + 8 * g8;
        // raw: larger accumulator = better; coefficient mode: smaller score = better (min/max drop NaN operands, as in
        // the hot loop: a NaN query makes every value NaN, and that does compare as a hit)
        const float mg = MODE == 3 ? fmax3(fmax3(g[0], g[1], g[2]), fmax3(g[3], g[4], g[5]), fmaxf(g[6], g[7]))
                                   : fmin3(fmin3(g[0], g[1], g[2]), fmin3(g[3], g[4], g[5]), fminf(g[6], g[7]));
        const bool ghit = MODE == 3 ? !(mg <= thr_raw) : !(mg >= thr);  // (an all-NaN group compares as a hit)
        if (ghit) {
          const uint32_t lvb = MODE == 3 ? (lvw >> (8 * g8)) & 0xffu : 0xffu;  // tombstones, slots beyond the index
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float a = g[jj];
            const float score = MODE == 3 ? a * c_q : a;
            const bool live = (lvb >> jj) & 1u;
            const bool take = live && (MODE == 3 ? !(a <= thr_raw) : true) && !(score >= thr);
            const bool is_nan = score != score;
            nan = nan || (take && is_nan);
            const bool keep = take && !is_nan;
            if (keep && cnt < cap) slice[cnt] = pack_key(score, slot_base + (uint32_t)(c0 + 8 * g8 + jj));
            cnt += keep ? 1u : 0u;  // beyond the capacity only counted: select_kernel sees the overflow and falls back
          }
        }
      }
    }
  }
  if (bm) *hc = (unsigned short)min(cnt, 65535u);
  if (nan) atomicOr(flags, kFlagNaN);
  return released;
}

// One 32-column block of the hot loop: does any of this thread's 32 scores beat the query's threshold?
//  UNI (coefficient mode, short-K kernel): every eligible row of the row tile has the same first coefficient `a_u`
//  (rows of one magnitude share their power-of-two fp16 scale -- the usual case), so only b is read from shared
//  memory: 8 broadcast LDS.128 per block instead of 16.  The coefficient loads are what bounds this loop: a
//  warp-wide LDS.128 of one address still returns 512 bytes to the register file (~2 wavefronts), 1100 cycles of the
//  shared-memory pipe per item with (a, b) pairs against 1024 cycles of MMAs (ncu, profiles/r02_ncu_gemm_sk_c5_*).
template <int MODE, bool UNI>
__device__ __forceinline__ bool epi_block_hit(const uint32_t (&r)[32], const int c0, const uint32_t cs_addr,
                                              const uint32_t bs_addr, const float a_u, const float thr,
                                              const float thr_raw) {
  float m8[4];
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    float v0 = __uint_as_float(r[8 * g8]), v1 = __uint_as_float(r[8 * g8 + 1]);
    float v2 = __uint_as_float(r[8 * g8 + 2]), v3 = __uint_as_float(r[8 * g8 + 3]);
    float v4 = __uint_as_float(r[8 * g8 + 4]), v5 = __uint_as_float(r[8 * g8 + 5]);
    float v6 = __uint_as_float(r[8 * g8 + 6]), v7 = __uint_as_float(r[8 * g8 + 7]);
    if (MODE == 3) {
      m8[g8] = fmax3(fmax3(v0, v1, v2), fmax3(v3, v4, v5), fmaxf(v6, v7));
    } else {
      if (UNI) {
        const float4 b0 = lds128(bs_addr + (c0 + 8 * g8) * 4), b1 = lds128(bs_addr + (c0 + 8 * g8 + 4) * 4);
        v0 = fmaf(v0, a_u, b0.x); v1 = fmaf(v1, a_u, b0.y); v2 = fmaf(v2, a_u, b0.z); v3 = fmaf(v3, a_u, b0.w);
        v4 = fmaf(v4, a_u, b1.x); v5 = fmaf(v5, a_u, b1.y); v6 = fmaf(v6, a_u, b1.z); v7 = fmaf(v7, a_u, b1.w);
      } else {
        const float4 k0 = lds128(cs_addr + (c0 + 8 * g8) * 8), k1 = lds128(cs_addr + (c0 + 8 * g8 + 2) * 8);
        const float4 k2 = lds128(cs_addr + (c0 + 8 * g8 + 4) * 8), k3 = lds128(cs_addr + (c0 + 8 * g8 + 6) * 8);
        v0 = fmaf(v0, k0.x, k0.y); v1 = fmaf(v1, k0.z, k0.w);
        v2 = fmaf(v2, k1.x, k1.y); v3 = fmaf(v3, k1.z, k1.w);
        v4 = fmaf(v4, k2.x, k2.y); v5 = fmaf(v5, k2.z, k2.w);
        v6 = fmaf(v6, k3.x, k3.y); v7 = fmaf(v7, k3.z, k3.w);
      }
      // (min drops NaN operands; a NaN query makes every score NaN, which still reaches the survivor section)
      m8[g8] = fmin3(fmin3(v0, v1, v2), fmin3(v3, v4, v5), fminf(v6, v7));
    }
  }
  const float mall = MODE == 3 ? fmax3(m8[0], m8[1], fmaxf(m8[2], m8[3])) : fmin3(m8[0], m8[1], fminf(m8[2], m8[3]));
  return MODE == 3 ? !(mall <= thr_raw) : !(mall >= thr);
}

template <int MODE, bool UNI>
__device__ __forceinline__ uint32_t epi_block_mask(const uint32_t taddr, const uint32_t cs_addr, const uint32_t bs_addr,
                                                   const float a_u, const float thr, const float thr_raw) {
  uint32_t bm = 0;
  uint32_t ra[32], rb[32];
  tmem_ld_32x32b_x32(taddr, ra);
  tmem_ld_wait();
  tmem_ld_32x32b_x32(taddr + 32, rb);
  bm |= epi_block_hit<MODE, UNI>(ra, 0, cs_addr, bs_addr, a_u, thr, thr_raw) ? 1u : 0u;
  tmem_ld_wait();
  tmem_ld_32x32b_x32(taddr + 64, ra);
  bm |= epi_block_hit<MODE, UNI>(rb, 32, cs_addr, bs_addr, a_u, thr, thr_raw) ? 2u : 0u;
  tmem_ld_wait();
  tmem_ld_32x32b_x32(taddr + 96, rb);
  bm |= epi_block_hit<MODE, UNI>(ra, 64, cs_addr, bs_addr, a_u, thr, thr_raw) ? 4u : 0u;
  tmem_ld_wait();
  bm |= epi_block_hit<MODE, UNI>(rb, 96, cs_addr, bs_addr, a_u, thr, thr_raw) ? 8u : 0u;
  return bm;
}

// `uni`: the row tile's eligible rows share the coefficient a_u (warp-uniform; always false for the k-ring kernel,
// which stages coefficients per item).  bs_addr: this half's 128 b values as a plain float array (uni only).
// Returns true when the survivor section has already arrived on `release_bar` (the accumulator stage's "drained"
// barrier, a shared::cta address; 0: the caller releases the stage itself).
template <int MODE>
__device__ __forceinline__ bool epi_filter_half(const GemmParams& p, const uint32_t taddr, const uint32_t cs_addr,
                                                const float thr, const float thr_raw, const float c_q, const int qidx,
                                                const int64_t n0, const int half, const int64_t unit,
                                                unsigned short* hitcnt, const uint32_t (&lv)[4], const bool uni = false,
                                                const uint32_t bs_addr = 0, const float a_u = 0.f,
                                                const uint32_t release_bar = 0, const int dbg = -1) {
  const int debug = dbg < 0 ? p.debug : dbg;  // (the short-K kernel's production instance passes 0: no per-item loads)
  if (debug & 4) return false;
  uint32_t bm;
  if (MODE == 0 && uni) bm = epi_block_mask<MODE, true>(taddr, cs_addr, bs_addr, a_u, thr, thr_raw);
  else bm = epi_block_mask<MODE, false>(taddr, cs_addr, bs_addr, a_u, thr, thr_raw);
  if (qidx >= p.q) bm = 0;  // padding queries (threshold -inf) never match; NaN accumulators cannot fake a hit either
  if (!__any_sync(0xffffffffu, bm != 0) || (debug & 8)) return false;
  // (qidx < p.q <= kGemmMaxQueries whenever bm != 0; other lanes only take part in the warp-collective loads)
  const int qi = bm ? qidx : 0;
  return epi_survivors<MODE>(p.cand + (size_t)qi * p.cand_stride + (size_t)(unit * 2 + half) * p.cand_cap, p.cand_cap,
                             p.flags, taddr, cs_addr, thr, thr_raw, c_q, (uint32_t)(n0 + half * (BN / 2)),
                             hitcnt + half * kGemmMaxQueries + qi, bm, lv[0], lv[1], lv[2], lv[3], release_bar);
}

// Seed pass of one accumulator half for the calling thread's query (short-K kernel): the kSeedR smallest 32-row
// block minima of the 128 scores, ascending.  Seed statistic as in the k-ring kernel's MODE 1: the rank-th smallest
// block minimum is >= the rank-th smallest score, with equality unless two of the sample's `rank` best rows share a
// block -- a slightly looser threshold then, never a wrong one.
__device__ __forceinline__ void epi_seed_half(const uint32_t taddr, const uint32_t cs_addr, float (&sd)[kSeedR],
                                              const bool skip) {
  const float kInf = __int_as_float(0x7f800000);
#pragma unroll
  for (int i = 0; i < kSeedR; ++i) sd[i] = kInf;
  if (skip) return;
  auto block_min = [&](const uint32_t (&r)[32], const int c0) {
    float m8[4];
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      const float4 k0 = lds128(cs_addr + (c0 + 8 * g8) * 8), k1 = lds128(cs_addr + (c0 + 8 * g8 + 2) * 8);
      const float4 k2 = lds128(cs_addr + (c0 + 8 * g8 + 4) * 8), k3 = lds128(cs_addr + (c0 + 8 * g8 + 6) * 8);
      const float v0 = fmaf(__uint_as_float(r[8 * g8]), k0.x, k0.y), v1 = fmaf(__uint_as_float(r[8 * g8 + 1]), k0.z, k0.w);
      const float v2 = fmaf(__uint_as_float(r[8 * g8 + 2]), k1.x, k1.y), v3 = fmaf(__uint_as_float(r[8 * g8 + 3]), k1.z, k1.w);
      const float v4 = fmaf(__uint_as_float(r[8 * g8 + 4]), k2.x, k2.y), v5 = fmaf(__uint_as_float(r[8 * g8 + 5]), k2.z, k2.w);
      const float v6 = fmaf(__uint_as_float(r[8 * g8 + 6]), k3.x, k3.y), v7 = fmaf(__uint_as_float(r[8 * g8 + 7]), k3.z, k3.w);
      m8[g8] = fmin3(fmin3(v0, v1, v2), fmin3(v3, v4, v5), fminf(v6, v7));
    }
    const float m = fmin3(m8[0], m8[1], fminf(m8[2], m8[3]));
    if (m < sd[kSeedR - 1]) {
      sd[kSeedR - 1] = m;
#pragma unroll
      for (int i = kSeedR - 1; i > 0; --i) {
        const float lo = fminf(sd[i - 1], sd[i]), hi = fmaxf(sd[i - 1], sd[i]);
        sd[i - 1] = lo;
        sd[i] = hi;
      }
    }
  };
  uint32_t ra[32], rb[32];
  tmem_ld_32x32b_x32(taddr, ra);
  tmem_ld_wait();
  tmem_ld_32x32b_x32(taddr + 32, rb);
  block_min(ra, 0);
  tmem_ld_wait();
  tmem_ld_32x32b_x32(taddr + 64, ra);
  block_min(rb, 32);
  tmem_ld_wait();
  tmem_ld_32x32b_x32(taddr + 96, rb);
  block_min(ra, 64);
  tmem_ld_wait();
  block_min(rb, 96);
}

// The live words of the 128 rows [slot0, slot0 + 128) (slot0 a multiple of 128), bits of slots beyond the index cleared
// (raw epilogue only: the coefficient epilogue gets tombstones and masks through b = +inf).
__device__ __forceinline__ void load_live4(const IndexView& iv, const int64_t slot0, uint32_t (&lv)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t s = slot0 + 32 * i;
    uint32_t w = 0;
    if (s < iv.n_slots) {
      w = __ldg(iv.live + (s >> 5));
      if (s + 32 > iv.n_slots) w &= (1u << (uint32_t)(iv.n_slots - s)) - 1u;
    }
    lv[i] = w;
  }
}

// item -> (row tile, query tile).  The query tile is rotated by the row-tile index so that every
// CTA meets every query tile: a query's candidates then spread evenly over all CTAs' slices.
template <int MODE, int CG>
__device__ __forceinline__ void item_tiles(const GemmParams& p, int64_t w, uint32_t rank, int64_t& nt_idx,
                                           int64_t& n_tile, int& m_tile) {
  // 32-bit arithmetic: items < 2^31 (launch_gemm_topk checks), and every role pays this once per item
  const uint32_t mg = (uint32_t)p.num_m_tiles / CG;  // query-tile groups: a CTA pair takes two adjacent tiles
  const uint32_t wi = (uint32_t)w;
  const uint32_t nt = wi / mg;
  const uint32_t r = wi - nt * mg + nt % mg;
  nt_idx = nt;
  m_tile = (int)((r >= mg ? r - mg : r) * CG + rank);
  n_tile = MODE == 1 ? (int64_t)nt * p.seed_stride : (int64_t)nt;
}

// MODE 0: main pass (threshold filter, per-row coefficients), 1: seed pass (per-thread smallest scores),
// 2: debug dump of all scores, 3: main pass with the raw epilogue (cosine, no mask).
// One instance per mode keeps the hot instance's code small: the epilogue is sensitive to instruction-cache
// misses (a 4x larger unrolled epilogue ran 3x slower).
template <int MODE, int CG, bool DIAG>
__device__ __forceinline__ void gemm_topk_body(const CUtensorMap& tmx, const CUtensorMap& tmq, const GemmParams& p) {
  const int debug = DIAG ? p.debug : 0;  // timing-experiment switches exist in the DIAG instantiation only
  using G = Geo<CG>;
  constexpr int kStages = G::kStages;
  constexpr int kBBytes = G::kBBytes;
  constexpr uint32_t kIdesc = G::kIdesc;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  unsigned char* sA = smem + G::kOffA;
  unsigned char* sB = smem + G::kOffB;
  float2* sCoef = reinterpret_cast<float2*>(smem + G::kOffCoef);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::kOffBar);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;   // accumulator stage complete (MMA -> epilogue)
  uint64_t* tempty = tfull + 2;        // accumulator + coefficient stage drained (epilogue -> MMA, stager)
  uint64_t* cfull = tempty + 2;        // coefficient stage published (stager -> epilogue)
  uint64_t* cempty = cfull + 2;        // coefficient stage drained (epilogue -> stager, CTA-local)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + 2);
  unsigned short* hitcnt = reinterpret_cast<unsigned short*>(smem + G::kOffCnt);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const IndexView& iv = p.iv;
  const int num_kb = (iv.dpad16 + BK - 1) / BK;
  // work units: CTAs (CG == 1) or CTA pairs (CG == 2); both CTAs of a pair walk the same item sequence
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int64_t unit = CG == 2 ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t nunits = (int64_t)gridDim.x / CG;
  const int64_t n_items = (MODE == 1 ? p.seed_tiles : p.num_n_tiles) * (p.num_m_tiles / CG);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiWarps * CG);  // pair: both CTAs' epilogue warps arrive on the leader's barrier
      mbar_init(&cfull[s], 2);
      mbar_init(&cempty[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < 2 * kGemmMaxQueries; i += kGemmThreads) hitcnt[i] = 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmq);
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long dbg_c0 = 0, dbg_t0 = 0;
  if ((debug & 32) && tid == 0 && blockIdx.x == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // pair: each CTA loads its own query block and its half of the row block; all bytes of a stage (both CTAs')
    // are counted on the LEADER's full barrier, which is the one the MMA issuer waits on.
    int s = 0;
    uint32_t ph = 0, it = 0;
    const uint32_t full_lead0 = CG == 2 ? mapa_u32(smem_u32(full), 0) : 0u;
    for (int64_t w = unit; w < n_items; w += nunits) {
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE, CG>(p, w, rank, nt_idx, n_tile, m_tile);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        mbar_wait(&empty[s], ph ^ 1u);
        if (lane == 0) {
          // debug (timing experiments only, results invalid): bit0 / bit1 stop re-loading A / B once the
          // ring has been filled once, to separate the load path from the MMA and epilogue cost.
          const bool ldA = !((debug & 1) && it >= (uint32_t)kStages);
          const bool ldB = !((debug & 2) && it >= (uint32_t)kStages);
          const uint32_t bytes = (ldA ? kABytes : 0) + (ldB ? kBBytes : 0);
          if (CG == 2) {
            if (rank == 0) {
              if (bytes) mbar_arrive_expect_tx(&full[s], bytes * 2);
              else mbar_arrive(&full[s]);
            }
            const uint32_t bar = full_lead0 + (uint32_t)s * 8u;
            if (ldA) tma_load_2d_pair(sA + (size_t)s * kABytes, &tmq, kb * BK, m_tile * BM, bar);
            if (ldB) tma_load_2d_pair(sB + (size_t)s * kBBytes, &tmx, kb * BK, (int)(n_tile * BN) + (int)rank * G::kBRows, bar);
          } else {
            if (bytes) mbar_arrive_expect_tx(&full[s], bytes);
            else mbar_arrive(&full[s]);
            if (ldA) tma_load_2d(sA + (size_t)s * kABytes, &tmq, kb * BK, m_tile * BM, &full[s]);
            if (ldB) tma_load_2d(sB + (size_t)s * kBBytes, &tmx, kb * BK, (int)(n_tile * BN), &full[s]);
          }
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------ MMA issuer (pair: leader CTA only) ------------------------------
    // Convergent code: all lanes run the loop with identical (uniform) operands and one elected lane issues
    // inside the asm, so the descriptors live in uniform registers and a k-step is a few dozen instructions.
    // Descriptor low words: (addr >> 4) | LBO<<16; a stage is 16 KB (A) / 32|16 KB (B) further, a K step 32 bytes.
    const uint32_t a_lo_base = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo_base = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version 1 | SWIZZLE_128B
    const uint32_t empty0 = smem_u32(empty), tfull0 = smem_u32(tfull);
    int s = 0;
    uint32_t ph = 0, ai = 0;
    for (int64_t w = unit; w < n_items; w += nunits, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      mbar_wait(&tempty[as], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo_base + (uint32_t)s * (kABytes >> 4);
        const uint32_t b_lo = b_lo_base + (uint32_t)s * (kBBytes >> 4);
        if (!(debug & 16)) {  // bit4 (timing experiments): barrier handshakes only, no MMA
          if (CG == 2) {
            umma_f16_elect_pair(d_tmem, a_lo, b_lo, desc_hi, kIdesc, kb != 0 ? 1u : 0u);
            umma_f16_elect_pair(d_tmem, a_lo + 2, b_lo + 2, desc_hi, kIdesc, 1u);
            umma_f16_elect_pair(d_tmem, a_lo + 4, b_lo + 4, desc_hi, kIdesc, 1u);
            umma_f16_elect_pair(d_tmem, a_lo + 6, b_lo + 6, desc_hi, kIdesc, 1u);
          } else {
            umma_f16_elect(d_tmem, a_lo, b_lo, desc_hi, kIdesc, kb != 0 ? 1u : 0u);
            umma_f16_elect(d_tmem, a_lo + 2, b_lo + 2, desc_hi, kIdesc, 1u);
            umma_f16_elect(d_tmem, a_lo + 4, b_lo + 4, desc_hi, kIdesc, 1u);
            umma_f16_elect(d_tmem, a_lo + 6, b_lo + 6, desc_hi, kIdesc, 1u);
          }
        }
        // ring slot free / accumulator complete once these MMAs retire (pair: signalled in both CTAs)
        if (CG == 2) {
          umma_commit_elect_pair(empty0 + s * 8, 3);
          if (kb == num_kb - 1) umma_commit_elect_pair(tfull0 + as * 8, 3);
        } else {
          umma_commit_elect(empty0 + s * 8);
          if (kb == num_kb - 1) umma_commit_elect(tfull0 + as * 8);
        }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if ((warp == 2 || warp == 3) && MODE != 3) {
    // ------------------------------ coefficient stager ------------------------------
    // lane owns rows lane, lane+32, ... of the tile.  Raw loads for the NEXT item are issued before
    // the current item's values are consumed, so nothing here waits on memory in steady state.
    constexpr int RPL = BN / 64;  // rows per lane: warps 2 and 3 stage 128 rows each
    const int r0 = (warp - 2) * (BN / 2);
    const float inv_sq = pow2_scale_inv(*p.qmaxabs);
    const float kInf = __int_as_float(0x7f800000);
    const bool has_mask = p.mask.bits != nullptr;
    const bool by_slot = has_mask && iv.ids_identity;
    struct Raw { float2 c[RPL]; uint32_t live[RPL]; uint64_t mask[RPL]; };
    auto fetch = [&](int64_t w, Raw& r) {
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE, CG>(p, w, rank, nt_idx, n_tile, m_tile);
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int64_t slot = n_tile * BN + r0 + lane + 32 * i;
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
    if (unit < n_items) fetch(unit, nxt);
    uint32_t ai = 0;
    for (int64_t w = unit; w < n_items; w += nunits, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE, CG>(p, w, rank, nt_idx, n_tile, m_tile);
      const Raw cur = nxt;
      if (w + nunits < n_items) fetch(w + nunits, nxt);
      mbar_wait(&cempty[as], aph ^ 1u);  // the epilogue has drained the previous use of this stage
      float2* cs = sCoef + as * BN;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int64_t slot = n_tile * BN + r0 + lane + 32 * i;
        bool elig = ((cur.live[i] >> (slot & 31)) & 1u) && ((cur.mask[i] >> (slot & 63)) & 1ull);
        if (elig && has_mask && !by_slot) {
          const uint64_t id = iv.ids[slot];
          elig = (id < (uint64_t)p.mask.nbits) && ((p.mask.bits[id >> 6] >> (id & 63)) & 1ull);
        }
        cs[r0 + lane + 32 * i] = elig ? make_float2(cur.c[i].x * inv_sq, cur.c[i].y) : make_float2(0.f, kInf);
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
    const uint32_t tempty_lead0 = CG == 2 ? mapa_u32(smem_u32(tempty), 0) : 0u;
    const float c_q = MODE == 3 ? -pow2_scale_inv(*p.qmaxabs) * 6.103515625e-05f : 0.f;
    const float inv_cq = MODE == 3 ? 1.0f / c_q : 0.f;  // exact: c_q is a (negative) power of two
    uint32_t ai = 0;
    for (int64_t w = unit; w < n_items; w += nunits, ++ai) {
      const uint32_t as = ai & 1u, aph = (ai >> 1) & 1u;
      int64_t nt_idx, n_tile;
      int m_tile;
      item_tiles<MODE, CG>(p, w, rank, nt_idx, n_tile, m_tile);
      const int64_t n0 = n_tile * BN;
      const int qidx = m_tile * BM + mrow;
      float thr = __int_as_float(0xff800000);  // -inf: padding queries never match
      if ((MODE == 0 || MODE == 3) && qidx < p.q) thr = p.thresh[qidx];
      uint32_t lv[4] = {0u, 0u, 0u, 0u};
      if (MODE == 3) load_live4(iv, n0 + half * (BN / 2), lv);  // in flight while the accumulator is awaited
      float sd[kSeedR];
#pragma unroll
      for (int i = 0; i < kSeedR; ++i) sd[i] = kInf;

      // (a warp-wide try_wait is one instruction: polling from a single lane + __syncwarp measured slower)
      if (MODE != 3) mbar_wait(&cfull[as], aph);
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + half * (BN / 2);
      const uint32_t cs_addr = smem_u32(sCoef + as * BN + half * (BN / 2));
      if (MODE == 0 || MODE == 3) {
        epi_filter_half<MODE>(p, taddr, cs_addr, thr, thr * inv_cq, c_q, qidx, n0, half, unit, hitcnt, lv, false, 0, 0.f, 0, debug);
      } else {
        // seed pass / debug dump: one 32-column block at a time
#pragma unroll 1
        for (int c0 = 0; c0 < ((debug & 4) ? 0 : BN / 2); c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c0, r);
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
            // Seed statistic: the minimum of each 32-row block (one insertion per block instead of 32 tests).
            // The rank-th smallest block minimum is >= the rank-th smallest score, with equality unless two of
            // the sample's `rank` best rows share a block -- a slightly looser threshold then, never a wrong one.
            float m = fminf(sc[0], sc[1]);
#pragma unroll
            for (int j = 2; j < 32; j += 2) m = fmin3(m, sc[j], sc[j + 1]);
            if (m < sd[kSeedR - 1]) {
              sd[kSeedR - 1] = m;
#pragma unroll
              for (int i = kSeedR - 1; i > 0; --i) {
                const float lo = fminf(sd[i - 1], sd[i]), hi = fmaxf(sd[i - 1], sd[i]);
                sd[i - 1] = lo;
                sd[i] = hi;
              }
            }
          } else if (qidx < p.q) {
            // debug dump of every approximate score (tests only; small inputs)
#pragma unroll
            for (int j = 0; j < 32; ++j) p.seeds[(size_t)qidx * p.seed_stride + (size_t)(n0 + col0 + j)] = sc[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // one arrival per warp
        if (CG == 2) mbar_arrive_cluster(tempty_lead0 + as * 8u);
        else mbar_arrive(&tempty[as]);
        if (MODE != 3) mbar_arrive(&cempty[as]);
      }
      if (MODE == 1 && qidx < p.q) {
        float* out = p.seeds + (((size_t)qidx * p.seed_tiles + nt_idx) * 2 + half) * kSeedR;
#pragma unroll
        for (int i = 0; i < kSeedR; ++i) out[i] = sd[i];
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (MODE == 0 || MODE == 3) {
    // publish this unit's per-query candidate counts (every entry, zeros included: nothing is pre-filled).
    // A pair's CTAs own the query tiles of their own parity.
    for (int i = tid; i < 2 * p.q; i += kGemmThreads) {
      const int hf = i >= p.q ? 1 : 0, qi = i - hf * p.q;
      if (CG == 1 || (uint32_t)((qi >> 7) & 1) == rank)
        p.slice_cnt[(size_t)(unit * 2 + hf) * p.q + qi] = hitcnt[hf * kGemmMaxQueries + qi];
    }
  }
  if ((debug & 32) && tid == 0 && blockIdx.x == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    const long long c1 = clock64();
    printf("[gemm_topk] cycles %lld ns %lld -> %.1f MHz\n", c1 - dbg_c0, t1 - dbg_t0, 1e3 * (double)(c1 - dbg_c0) / (double)(t1 - dbg_t0));
  }
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int MODE, bool DIAG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmq,
                 const GemmParams p) {
  griddep_wait();
  if (p.skip && *p.skip) return;  // device-side route: the scan answers this batch
  gemm_topk_body<MODE, 1, DIAG>(tmx, tmq, p);
}

// =====================================================================================================
// Short-K, ROW-TILE-STATIONARY instance (dpad16 <= 128, i.e. the whole K extent fits ONE shared-memory stage).
//
// Why a second main-pass kernel: at d = 128 an item of the k-ring kernel above is only 8 MMAs (1024 clk) but
//   * 3 tcgen05.commits of ~430 clk each (one per 64-column k-step + the accumulator hand-over): the fixed
//     hand-shake per item was as large as the math (DESIGN.md, round-1 in-kernel clocks);
//   * 96 KB of operands through the L2->SM port (a 32 KB query block AND a 64 KB row block per item), although
//     consecutive items of a CTA could share either.
// Here a CTA keeps ONE 256-row tile of the database (<= 64 KB) in shared memory and streams every 128-query
// block of the batch past it (A ring of kNA stages, <= 32 KB each): the row tile is read from L2 once per
// num_m_tiles items (34 KB of operand traffic per item instead of 96), and an item costs exactly ONE
// tcgen05.commit on a barrier `done[s]` that both the producer (A stage s is free) and the epilogue
// (accumulator s & 1 is complete) wait on.  The per-row coefficients (MODE 0) are staged once per row tile
// instead of once per item.  Epilogue, TMEM layout, candidate slices and certification are unchanged.
namespace sk {
constexpr int kNA = 4;                                  // A-ring stages (items of prefetch)
constexpr int kAStageBytes = 2 * kABytes;               // 128 queries x 128 fp16 (two SWIZZLE_128B chunks)
constexpr int kBChunkBytes = BN * BK * 2;               // 256 rows x 64 fp16 = 32 KB
constexpr size_t kOffB = 0;
constexpr size_t kOffA = kOffB + 2 * (size_t)kBChunkBytes;
constexpr size_t kOffCoef = kOffA + (size_t)kNA * kAStageBytes;
constexpr size_t kOffBOnly = kOffCoef + 2 * BN * sizeof(float2);   // [2 stages][BN] b alone (uniform-a hot loop)
constexpr size_t kOffUni = kOffBOnly + 2 * BN * sizeof(float);     // [2 stages][2 stager warps] shared a, or NaN
constexpr size_t kOffBar = kOffUni + 16;
constexpr size_t kOffCnt = kOffBar + 24 * 8 + 16;
constexpr size_t kSmemUsed = kOffCnt + 2 * 2 * kGemmMaxQueries;
constexpr size_t kSmemBytes = kSmemUsed + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget");
}  // namespace sk

// waits like mbar_wait; with `diag` the cycles spent waiting are added to `acc` (in-kernel diagnostics, debug bit 5)
__device__ __forceinline__ void mbar_wait_d(uint64_t* bar, uint32_t parity, bool diag, long long& acc) {
  if (!diag) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

template <int MODE, bool DIAG>
__device__ __forceinline__ void gemm_topk_sk_body(const CUtensorMap& tmx, const CUtensorMap& tmq, const GemmParams& p) {
  static_assert(MODE == 0 || MODE == 1 || MODE == 3, "main pass (0: coefficients, 3: raw) or seed pass (1)");
  constexpr int kNA = sk::kNA;
  constexpr uint32_t kIdesc = Geo<1>::kIdesc;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  unsigned char* sB = smem + sk::kOffB;
  unsigned char* sA = smem + sk::kOffA;
  float2* sCoef = reinterpret_cast<float2*>(smem + sk::kOffCoef);
  float* sBOnly = reinterpret_cast<float*>(smem + sk::kOffBOnly);
  float* sUni = reinterpret_cast<float*>(smem + sk::kOffUni);
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + sk::kOffBar);  // TMA -> MMA: A stage loaded
  uint64_t* done = afull + kNA;    // MMA -> producer + epilogue: the item's MMAs have completed (ONE commit per item)
  uint64_t* bfull = done + kNA;    // TMA -> MMA: row tile loaded
  uint64_t* tempty = bfull + 1;    // epilogue -> MMA: accumulator drained            [2]
  uint64_t* cfull = tempty + 2;    // stager -> epilogue: coefficients of a row tile  [2]
  uint64_t* cempty = cfull + 2;    // epilogue -> stager                              [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + 2);
  unsigned short* hitcnt = reinterpret_cast<unsigned short*>(smem + sk::kOffCnt);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const IndexView& iv = p.iv;
  const int KS = (iv.dpad16 + BK - 1) / BK;         // 64-column chunks: 1 or 2
  const int nk16 = (iv.dpad16 + 15) / 16;           // K = 16 MMA steps per item (<= 8)
  const int64_t unit = blockIdx.x, nunits = gridDim.x;
  const int M = p.num_m_tiles;
  // this CTA's row tiles: unit, unit + nunits, ... of the index -- or, for the seed pass, of the evenly strided sample
  // (sample tile t is row tile t * seed_stride)
  const int64_t tiles_all = MODE == 1 ? p.seed_tiles : p.num_n_tiles;
  const int64_t tstride = MODE == 1 ? p.seed_stride : 1;
  const int64_t ntiles = unit < tiles_all ? (tiles_all - unit + nunits - 1) / nunits : 0;
  const int64_t total = ntiles * M;                 // items of this CTA
  // (in-kernel timers live in their own instantiation: even predicated off they cost the item loops a branch or two)
  const bool diag = DIAG && (p.debug & 32) && blockIdx.x == 0;
  const int debug = DIAG ? p.debug : 0;  // timing-experiment switches exist in the DIAG instantiation only
  long long w_a = 0, w_b = 0, w_c = 0, w_x = 0;     // per-role wait cycles; w_x: MMA issue / epilogue filter cycles (diag)

  if (tid == 0) {
    for (int s = 0; s < kNA; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&done[s], 1);
    }
    mbar_init(bfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tempty[s], kEpiWarps);
      mbar_init(&cfull[s], 2);
      mbar_init(&cempty[s], kEpiWarps);
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
  long long dbg_c0 = 0, dbg_t0 = 0;
  if (diag && tid == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // In-order stream of loads, each issued as soon as its precondition holds: A(it) needs the MMAs of item it - kNA
    // (its stage) to have completed, B(j) those of the last item of row tile j - 1 (the single row-tile buffer).
    auto load_b = [&](int64_t j) {
      if (lane == 0) {
        const int n_row0 = (int)((unit + j * nunits) * tstride * BN);
        if (debug & 2) { mbar_arrive(bfull); return; }
        mbar_arrive_expect_tx(bfull, (uint32_t)KS * sk::kBChunkBytes);
        for (int kc = 0; kc < KS; ++kc) tma_load_2d(sB + (size_t)kc * sk::kBChunkBytes, &tmx, kc * BK, n_row0, bfull);
      }
    };
    int64_t next_b = 1;
    if (total > 0) load_b(0);
    int m = 0;
    for (int64_t it = 0; it < total; ++it) {
      const int s = (int)(it & (kNA - 1));
      const uint32_t u = (uint32_t)(it / kNA);
      if (it >= kNA) mbar_wait_d(&done[s], (u - 1u) & 1u, diag, w_a);
      if (lane == 0) {
        if ((debug & 1) && it >= kNA) {
          mbar_arrive(&afull[s]);
        } else {
          mbar_arrive_expect_tx(&afull[s], (uint32_t)KS * kABytes);
          for (int kc = 0; kc < KS; ++kc)
            tma_load_2d(sA + (size_t)s * sk::kAStageBytes + (size_t)kc * kABytes, &tmq, kc * BK, m * BM, &afull[s]);
        }
      }
      if (++m == M) m = 0;
      // the wait just passed covers B(next_b)'s precondition when item next_b*M - 1 is not younger than it - kNA
      if (next_b < ntiles && it - kNA >= next_b * M - 1) load_b(next_b++);
      __syncwarp();
    }
    while (next_b < ntiles) {  // (fewer query tiles than ring stages: the triggers above lie beyond the last item)
      const int64_t il = next_b * M - 1;
      mbar_wait_d(&done[il & (kNA - 1)], (uint32_t)(il / kNA) & 1u, diag, w_b);
      load_b(next_b++);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t a_lo_base = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo_base = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version 1 | SWIZZLE_128B
    const uint32_t done0 = smem_u32(done);
    int m = 0;
    uint32_t j = 0;
    for (int64_t it = 0; it < total; ++it) {
      const int s = (int)(it & (kNA - 1));
      const uint32_t u = (uint32_t)(it / kNA);
      const uint32_t as = (uint32_t)it & 1u, aph = (uint32_t)(it >> 1) & 1u;
      if (m == 0) mbar_wait_d(bfull, j & 1u, diag, w_b);
      mbar_wait_d(&tempty[as], aph ^ 1u, diag, w_c);
      mbar_wait_d(&afull[s], u & 1u, diag, w_a);
      tc_fence_after();
      const long long ti0 = diag ? clock64() : 0;
      const uint32_t d_tmem = tmem_base + as * BN;
      const uint32_t a_lo = a_lo_base + (uint32_t)s * (sk::kAStageBytes >> 4);
      // (Issuing the item as two N = 128 halves, each released by its own four epilogue warps, was measured in round
      // 2: 13.4 -> 14.4 ms on C5 -- the N = 128 MMAs and the second wait in the middle of the item cost more than the
      // decoupling of the halves gained.)
      if (!(debug & 16)) {
#pragma unroll 1
        for (int k = 0; k < nk16; ++k) {
          // chunk k >> 2 (16 / 32 KB further), K step k & 3 (32 bytes further)
          const uint32_t ao = (uint32_t)(k >> 2) * (kABytes >> 4) + (uint32_t)(k & 3) * 2u;
          const uint32_t bo = (uint32_t)(k >> 2) * (sk::kBChunkBytes >> 4) + (uint32_t)(k & 3) * 2u;
          umma_f16_elect(d_tmem, a_lo + ao, b_lo_base + bo, desc_hi, kIdesc, k != 0 ? 1u : 0u);
        }
      }
      umma_commit_elect(done0 + s * 8);
      if (diag) w_x += clock64() - ti0;
      if (++m == M) { m = 0; ++j; }
    }
  } else if ((warp == 2 || warp == 3) && MODE != 3) {
    // ------------------------------ coefficient stager: once per ROW TILE ------------------------------
    constexpr int RPL = BN / 64;
    const int r0 = (warp - 2) * (BN / 2);
    const float inv_sq = pow2_scale_inv(*p.qmaxabs);
    const float kInf = __int_as_float(0x7f800000);
    const bool has_mask = p.mask.bits != nullptr;
    const bool by_slot = has_mask && iv.ids_identity;
    for (int64_t j = 0; j < ntiles; ++j) {
      const uint32_t cb = (uint32_t)j & 1u, cph = (uint32_t)(j >> 1) & 1u;
      const int64_t n0 = (unit + j * nunits) * tstride * BN;
      float2 cv[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int64_t slot = n0 + r0 + lane + 32 * i;
        cv[i] = make_float2(0.f, kInf);
        if (slot < iv.n_slots) {
          bool elig = (__ldg(iv.live + (slot >> 5)) >> (slot & 31)) & 1u;
          if (elig && has_mask) {
            const uint64_t id = by_slot ? (uint64_t)slot : iv.ids[slot];
            elig = (id < (uint64_t)p.mask.nbits) && ((__ldg(p.mask.bits + (id >> 6)) >> (id & 63)) & 1ull);
          }
          if (elig) {
            const float2 c = __ldg(iv.coef + slot);
            cv[i] = make_float2(c.x * inv_sq, c.y);
          }
        }
      }
      // Do this warp's eligible rows share one first coefficient?  (0 = no eligible row here, NaN = they differ)
      uint32_t abits = 0;
      bool same = true;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        if (cv[i].y < kInf || cv[i].y != cv[i].y) {  // eligible (b finite or NaN); ineligible rows carry b = +inf
          const uint32_t ab = __float_as_uint(cv[i].x);
          same = same && (abits == 0 || abits == ab) && ab != 0;
          abits = ab;
        }
      }
      const uint32_t ref = __reduce_max_sync(0xffffffffu, abits);
      same = __all_sync(0xffffffffu, same && (abits == 0 || abits == ref));
      mbar_wait(&cempty[cb], cph ^ 1u);
      float2* cs = sCoef + cb * BN;
      float* bs = sBOnly + cb * BN;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        cs[r0 + lane + 32 * i] = cv[i];
        bs[r0 + lane + 32 * i] = cv[i].y;
      }
      if (lane == 0) sUni[cb * 2 + (warp - 2)] = same ? __uint_as_float(ref) : __int_as_float(0x7fc00000);
      __syncwarp();
      if (lane == 0) mbar_arrive(&cfull[cb]);
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------ epilogue: thread == (query, column half) ------------------------------
    const int quarter = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const int mrow = quarter * 32 + lane;
    const float c_q = MODE == 3 ? -pow2_scale_inv(*p.qmaxabs) * 6.103515625e-05f : 0.f;
    const float inv_cq = MODE == 3 ? 1.0f / c_q : 0.f;  // exact: c_q is a (negative) power of two
    uint32_t lv[4] = {0u, 0u, 0u, 0u};
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * (BN / 2);
    uint32_t it = 0;  // item counter of this CTA (< 2^31: launch_gemm_topk checks the item count)
    // (row tile, then query tile: everything that depends on the row tile only is set up once per 32 items)
    for (uint32_t j = 0; j < (uint32_t)ntiles; ++j) {
      const uint32_t cb = j & 1u;
      const int64_t nt_idx = unit + (int64_t)j * nunits;  // (sample) tile index
      const int64_t n0 = nt_idx * tstride * BN;
      const uint32_t cs_addr = smem_u32(sCoef + cb * BN + half * (BN / 2));
      const uint32_t bs_addr = smem_u32(sBOnly + cb * BN + half * (BN / 2));
      bool uni = false;
      float a_u = 0.f;
      if (MODE == 3) load_live4(iv, n0 + half * (BN / 2), lv);
      if (MODE != 3) {
        mbar_wait_d(&cfull[cb], (j >> 1) & 1u, diag, w_c);
        const float u0 = sUni[cb * 2], u1 = sUni[cb * 2 + 1];  // one word per stager warp (128 rows each)
        uni = u0 == u0 && u1 == u1 && (u0 == u1 || u0 == 0.f || u1 == 0.f) && !(debug & 64);
        a_u = u0 != 0.f ? u0 : u1;
      }
      int qidx = mrow;
#pragma unroll 1
      for (int m = 0; m < M; ++m, ++it, qidx += BM) {
        const uint32_t s = it & (uint32_t)(kNA - 1), as = it & 1u;
        float thr = __int_as_float(0xff800000);  // -inf: padding queries never match
        if (MODE != 1 && qidx < p.q) thr = p.thresh[qidx];
        mbar_wait_d(&done[s], (it / kNA) & 1u, diag, w_a);
        tc_fence_after();
        const long long te0 = diag ? clock64() : 0;
        bool released = false;
        float sd[kSeedR];
        if (MODE == 1) {
          epi_seed_half(lane_addr + as * BN, cs_addr, sd, (debug & 4) != 0);
        } else {
          released = epi_filter_half<MODE == 1 ? 0 : MODE>(p, lane_addr + as * BN, cs_addr, thr, thr * inv_cq, c_q, qidx,
                                                           n0, half, unit, hitcnt, lv, uni, bs_addr, a_u,
                                                           smem_u32(&tempty[as]), debug);
        }
        if (diag) w_x += clock64() - te0;
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && !released) mbar_arrive(&tempty[as]);
        if (MODE == 1 && qidx < p.q) {  // the thread's kSeedR smallest 32-row block minima of this (query, tile, half)
          float4* out = reinterpret_cast<float4*>(p.seeds + (((size_t)qidx * p.seed_tiles + nt_idx) * 2 + half) * kSeedR);
          static_assert(kSeedR == 8, "two 16-byte stores");
          out[0] = make_float4(sd[0], sd[1], sd[2], sd[3]);
          out[1] = make_float4(sd[4], sd[5], sd[6], sd[7]);
        }
      }
      if (MODE != 3 && lane == 0) mbar_arrive(&cempty[cb]);  // (after the tile's last __syncwarp)
    }
  }

  if (diag && (tid == 0 || tid == 32 || (warp >= kEpiWarp0 && lane == 0)))
    printf("[gemm_topk_sk] warp %d items %lld waits: A/done %lld, B %lld, tmem/coef %lld clk; issue/filter %lld clk\n", warp,
           (long long)total, w_a, w_b, w_c, w_x);
  tc_fence_before();
  __syncthreads();
  if (MODE != 1)
    for (int i = tid; i < 2 * p.q; i += kGemmThreads) {
      const int hf = i >= p.q ? 1 : 0, qi = i - hf * p.q;
      p.slice_cnt[(size_t)(unit * 2 + hf) * p.q + qi] = hitcnt[hf * kGemmMaxQueries + qi];
    }
  if (diag && tid == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    const long long c1 = clock64();
    printf("[gemm_topk_sk] cycles %lld ns %lld -> %.1f MHz, %.0f clk/item\n", c1 - dbg_c0, t1 - dbg_t0,
           1e3 * (double)(c1 - dbg_c0) / (double)(t1 - dbg_t0), total ? (double)(c1 - dbg_c0) / (double)total : 0.0);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int MODE, bool DIAG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_sk_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmq,
                    const GemmParams p) {
  griddep_wait();
  if (p.skip && *p.skip) return;
  gemm_topk_sk_body<MODE, DIAG>(tmx, tmq, p);
}

// CTA-pair instance: clusters of two CTAs (the two SMs of a TPC), tcgen05.mma.cta_group::2.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_topk_pair_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmq,
                      const GemmParams p) {
  griddep_wait();
  if (p.skip && *p.skip) return;
  gemm_topk_body<MODE, 2, true>(tmx, tmq, p);
}

// One warp per query: the rank-th smallest of the query's seed scores becomes its threshold.
// Each lane keeps the kSeedR smallest of its strided share in registers (single pass), then the warp
// pops the global minimum `rank` times.
__global__ void seed_finalize_kernel(const SeedFinalizeParams p) {
  griddep_wait();
  if (p.skip && *p.skip) return;
  const int lane = threadIdx.x & 31;
  const int qi = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (qi >= p.q) return;
  const float kInf = __int_as_float(0x7f800000);
  const int64_t total = p.seed_tiles * 2 * kSeedR;
  const float* s = p.seeds + (size_t)qi * total;
  float sd[kSeedR];
#pragma unroll
  for (int i = 0; i < kSeedR; ++i) sd[i] = kInf;
  for (int64_t i0 = 0; i0 < total; i0 += 32 * 8) {  // 8 independent loads in flight per lane
    float vv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t i = i0 + 32 * u + lane;
      vv[u] = i < total ? __ldg(s + i) : kInf;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float v = vv[u];
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

// The same with a whole block per query, for small batches: a single query's seed list is up to 64K scores
// (10M rows, k = 10), which one warp walks in ~30 us.  Eight warps take strided shares, their per-lane lists are
// merged lane-wise by warp 0, which then pops as above.
__global__ void __launch_bounds__(256) seed_finalize_block_kernel(const SeedFinalizeParams p) {
  griddep_wait();
  if (p.skip && *p.skip) return;
  __shared__ float s_sd[8][kSeedR][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qi = blockIdx.x;
  const float kInf = __int_as_float(0x7f800000);
  const int64_t total = p.seed_tiles * 2 * kSeedR;
  const float* s = p.seeds + (size_t)qi * total;
  float sd[kSeedR];
#pragma unroll
  for (int i = 0; i < kSeedR; ++i) sd[i] = kInf;
  auto insert = [&](float v) {
    if (v < sd[kSeedR - 1]) {  // false for NaN
      sd[kSeedR - 1] = v;
#pragma unroll
      for (int j = kSeedR - 1; j > 0; --j) {
        const float lo = fminf(sd[j - 1], sd[j]), hi = fmaxf(sd[j - 1], sd[j]);
        sd[j - 1] = lo;
        sd[j] = hi;
      }
    }
  };
  for (int64_t i0 = (int64_t)warp * 256; i0 < total; i0 += 8 * 256) {
    float vv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t i = i0 + 32 * u + lane;
      vv[u] = i < total ? __ldg(s + i) : kInf;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) insert(vv[u]);
  }
#pragma unroll
  for (int j = 0; j < kSeedR; ++j) s_sd[warp][j][lane] = sd[j];
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < 8; ++w)
#pragma unroll
    for (int j = 0; j < kSeedR; ++j) insert(s_sd[w][j][lane]);
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
  if (p.num_n_tiles * p.num_m_tiles >= (1ll << 31)) return cudaErrorInvalidValue;
  const bool pair = p.pair != 0;
  if (pair && ((grid & 1) || (p.num_m_tiles & 1) || (p.seed_mode != 0 && p.seed_mode != 3))) return cudaErrorInvalidValue;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 15]) {  // once per device: the driver call is not free
    const int a = (int)Geo<1>::kSmemBytes, b = (int)Geo<2>::kSmemBytes;
    cudaError_t e = cudaFuncSetAttribute(gemm_topk_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    const int c = (int)sk::kSmemBytes;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_sk_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_sk_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_sk_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_sk_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_sk_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_topk_sk_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c);
    if (e != cudaSuccess) return e;
    attr_set[dev & 15] = true;
  }
  const CUtensorMap* tx = reinterpret_cast<const CUtensorMap*>(tmap_x_host);
  const CUtensorMap* tq = reinterpret_cast<const CUtensorMap*>(tmap_q_host);
  const size_t sm1 = Geo<1>::kSmemBytes, sm2 = Geo<2>::kSmemBytes;
  const dim3 g(grid), b(kGemmThreads);
  if (p.short_k) {
    if (pair || p.iv.dpad16 > 2 * BK || p.seed_mode == 2) return cudaErrorInvalidValue;
    const bool diag = p.debug != 0;  // any timing-experiment switch: the instrumented instantiation
    if (p.seed_mode == 1)
      return diag ? launch_pdl(gemm_topk_sk_kernel<1, true>, g, b, sk::kSmemBytes, st, *tx, *tq, p)
                  : launch_pdl(gemm_topk_sk_kernel<1, false>, g, b, sk::kSmemBytes, st, *tx, *tq, p);
    if (p.seed_mode == 0)
      return diag ? launch_pdl(gemm_topk_sk_kernel<0, true>, g, b, sk::kSmemBytes, st, *tx, *tq, p)
                  : launch_pdl(gemm_topk_sk_kernel<0, false>, g, b, sk::kSmemBytes, st, *tx, *tq, p);
    return diag ? launch_pdl(gemm_topk_sk_kernel<3, true>, g, b, sk::kSmemBytes, st, *tx, *tq, p)
                : launch_pdl(gemm_topk_sk_kernel<3, false>, g, b, sk::kSmemBytes, st, *tx, *tq, p);
  }
  if (pair && p.seed_mode == 0) return launch_pdl(gemm_topk_pair_kernel<0>, g, b, sm2, st, *tx, *tq, p);
  if (pair) return launch_pdl(gemm_topk_pair_kernel<3>, g, b, sm2, st, *tx, *tq, p);
  const bool diag = p.debug != 0;  // any timing-experiment switch: the instrumented instantiation
  if (p.seed_mode == 0)
    return diag ? launch_pdl(gemm_topk_kernel<0, true>, g, b, sm1, st, *tx, *tq, p)
                : launch_pdl(gemm_topk_kernel<0, false>, g, b, sm1, st, *tx, *tq, p);
  if (p.seed_mode == 1)
    return diag ? launch_pdl(gemm_topk_kernel<1, true>, g, b, sm1, st, *tx, *tq, p)
                : launch_pdl(gemm_topk_kernel<1, false>, g, b, sm1, st, *tx, *tq, p);
  if (p.seed_mode == 3)
    return diag ? launch_pdl(gemm_topk_kernel<3, true>, g, b, sm1, st, *tx, *tq, p)
                : launch_pdl(gemm_topk_kernel<3, false>, g, b, sm1, st, *tx, *tq, p);
  return launch_pdl(gemm_topk_kernel<2, false>, g, b, sm1, st, *tx, *tq, p);
}

cudaError_t launch_seed_finalize(const SeedFinalizeParams& p, cudaStream_t st) {
  if (p.q <= 0) return cudaSuccess;
  if (p.q <= 64 && p.seed_tiles * 2 * kSeedR >= 4096)  // few queries with long seed lists: a block per query
    return launch_pdl(seed_finalize_block_kernel, dim3(p.q), dim3(256), 0, st, p);
  const int blocks = (p.q * 32 + 255) / 256;
  return launch_pdl(seed_finalize_kernel, dim3(blocks), dim3(256), 0, st, p);
}

}  // namespace gfi
