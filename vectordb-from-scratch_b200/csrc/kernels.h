// kernels.h -- host-visible launch wrappers and shared parameter structs of libgfi.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace gfi {

constexpr int kMetricL2 = 0, kMetricCos = 1, kMetricDot = 2;

// ---- error word shared by all kernels of one search call ---------------------------
constexpr uint32_t kFlagNaN = 1u;        // a distance is NaN (reference: panic at flat_index.rs:62)
constexpr uint32_t kFlagZeroNorm = 2u;   // cosine with a zero-norm eligible row / query (distance.rs:51-55)
constexpr uint32_t kFlagInternal = 8u;   // watchdog / internal inconsistency

// Read-only device view of the index, passed by value to kernels.
struct IndexView {
  const float* x32;      // [n_slots][dpad]   fp32 rows, zero padded to dpad (multiple of 4)
  const __half* x16;     // [n_slots][dpad16] per-row power-of-two scaled fp16 shadow (tensor path) or null
  const uint64_t* ids;   // [n_slots]         internal id of each slot (slot order == id order)
  const float* norm;     // [n_slots]         reference-exact ||x||   (src/vector.rs:35-37)
  const float* sumsq;    // [n_slots]         reference-exact sum x^2 (pre-sqrt)
  const float2* coef;    // [n_slots]         tensor epilogue: score = acc * coef.x * inv_qscale + coef.y
  const uint32_t* live;  // [ceil(n_slots/32)] 1 = live, 0 = tombstone
  int64_t n_slots;
  int d, dpad, dpad16;
  int metric;
  int ids_identity;      // ids[s] == s for every slot (mask can be indexed by slot)
};

struct MaskView {
  const uint64_t* bits;  // eligibility by INTERNAL ID, or null
  int64_t nbits;
};

// ---- K1: streaming scan + per-warp top-K ---------------------------------------------
constexpr int kScanConsumerWarps = 16;
constexpr int kScanThreads = (kScanConsumerWarps + 1) * 32;
constexpr int kScanMaxStages = 8;

struct ScanParams {
  IndexView iv;
  MaskView mask;
  const float* q32;         // [*, dpad] padded queries
  const float* qnorm;       // exact ||q|| per query (cosine zero check)
  const uint32_t* qlist;    // indices of the queries to process, or null (= 0..nq-1)
  const uint32_t* nq_dev;   // device-side query count (predicated fallback launch), or null
  int nq;                   // host-side query count (used when nq_dev == null)
  int K;                    // per-query list length: power of two, 32..1024
  const uint64_t* floor64;  // per-query exclusive lower bound on keys (paged large-k), or null
  uint64_t* cand;           // [q][cand_stride] packed keys out (gridDim.x * K per query)
  uint32_t* cand_cnt;       // [q]
  int64_t cand_stride;
  uint32_t* flags;
  const uint32_t* gather_list;   // compacted eligible slots (filtered / tombstoned scans), or null
  const uint32_t* gather_count;  // device-side length of gather_list
  uint32_t* elig_out;            // or null: receives *gather_count + 1 (the host learns the mask's population)
  int rows_per_stage, seg_floats, nseg;
  int lanes_per_row;        // 8: 4 rows per warp at a time (dpad <= 256); 32: one row per warp (longer rows)
  int nstages;              // ring depth (<= kScanMaxStages)
  int stage_floats;         // floats per ring stage (rows_per_stage * row stride in the stage)
  // Fused tail (small single-pass searches, K <= 64): the LAST CTA to finish selects the K best keys of all CTAs,
  // re-scores them with the reference's arithmetic, sorts and emits -- the whole of K3 without two more launches
  // (a 10k x 128 single-query search is pure latency: three dependent kernels of ~10 us each).
  int fused;                // 1 = on (nq <= QT, nseg == 1, the ring is large enough: see scan_fused_tail_bytes)
  uint32_t* done_ctr;       // zeroed per call
  const uint32_t* ks;       // per-query k
  uint64_t* out_ids; float* out_dist; uint32_t* out_counts; int64_t kstride;
  // latency mode: the tail also stores the results, a copy of the call's control block and finally `done_seq`
  // into PINNED HOST memory (same [kstride] layout), so the host neither copies nor synchronises: it polls the word
  uint32_t* h_ctrl;         // host copy of the control block (kCtrlWords words) + the done word, or null
  const uint32_t* d_ctrl;   // the device control block
  uint64_t* h_out_ids; float* h_out_dist; uint32_t* h_out_counts;
  uint32_t done_seq;
  // certification of the fused tail's answers (exact.cuh, scan_lower_bound): unproven queries are listed
  uint32_t* up_count; uint32_t* up_list; uint32_t up_base, up_cap; float xnorm_max;
};
constexpr int kCtrlWords = 12, kCtrlDoneWord = 15;  // layout of the 64-byte control area of a host result block
size_t scan_fused_tail_bytes(int grid, int K, int dpad);
// eligible (live and unmasked) slots -> list[0..*count)
cudaError_t launch_compact_eligible(const IndexView& iv, const MaskView& mask, uint32_t* list, uint32_t* count,
                                    cudaStream_t st);
// QT = queries sharing one pass over the database (1,2,4).
cudaError_t launch_scan(const ScanParams& p, int QT, int grid, cudaStream_t st);
size_t scan_smem_bytes(int QT, int dpad, int K, int nstages, int stage_floats);

// ---- ingest / query preparation ------------------------------------------------------
struct IngestParams {
  float* x32; __half* x16; float* norm; float* sumsq; float2* coef; uint32_t* zero_flags;
  int64_t first_slot, n;
  int d, dpad, dpad16, metric;
  // generated rows (x32 is written by the kernel) when gen != 0
  int gen; uint32_t seed; uint64_t first_row; int kind;
};
cudaError_t launch_ingest(const IngestParams& p, cudaStream_t st);
cudaError_t launch_fill_ids(uint64_t* p, uint64_t first, int64_t n, cudaStream_t st);  // p[i] = first + i

struct PrepQueriesParams {
  const float* q_in;   // [q][d] unpadded
  float* q32;          // [q][dpad]
  __half* q16;         // [qpad][dpad16] (rows >= q zeroed) or null
  float* qnorm; float* qsumsq;
  float* qmaxabs;      // [1] max |q| over the batch (for the common fp16 scale)
  int q, qpad, d, dpad, dpad16;
  int use_smem;        // set by launch_prep_queries
  // latency mode (small host searches): q_in and ks_in are PINNED HOST memory read over PCIe by this kernel, which
  // also forwards the per-query k to device memory -- no copy-engine operation precedes the search
  const uint32_t* ks_in; uint32_t* ks_out;  // or null
};
cudaError_t launch_prep_queries(const PrepQueriesParams& p, cudaStream_t st);

// ---- K3: select + reference-exact rerank + certification -----------------------------
struct SelInfo {  // per query, written by select_kernel, consumed by rerank_finalize_kernel
  uint64_t pivot;   // KP-th best approximate key (cut-off of the reranked set)
  uint32_t nvalid;  // useful candidate keys seen
  uint32_t kpeff;   // candidates selected for rerank = min(KP, nvalid)
  uint32_t overflow;
  uint32_t done;    // warps of this query that finished their rerank share
  uint32_t has_cut; // tensor path: kpeff was shortened by the rerank cut (select_kernel); cut_key is then the
  uint32_t cut_key; // orderable approximate score of the best candidate that is NOT re-scored
};
struct SelectParams {
  IndexView iv;
  const float* q32; const float* qnorm; const float* qsumsq;
  const uint32_t* ks;         // per-query k
  const uint32_t* qlist;      // null = all
  const uint32_t* nq_dev; int nq;
  int nq_max;                 // upper bound on *nq_dev (grid sizing of the predicated launches)
  int few_candidates, warp_per_candidate;  // set by launch_select_rerank (latency modes for small batches)
  uint64_t* sel_keys;         // [q][KP] scratch: selected approximate keys, then exact keys
  SelInfo* sel_info;          // [q]
  const uint64_t* cand; const uint32_t* cand_cnt; int64_t cand_stride;
  int KP;                     // candidates reranked per query (power of two <= 1024)
  int list_len;               // > 0: the input is ascending lists of this length (scan path); 0: unordered slices
  // tensor path: cand[q] is nslices slices of slice_cap keys; slice_cnt[slice][q] = keys the tensor pass produced
  // for that (slice, query) (may exceed slice_cap: overflow).  Only the valid prefix of each slice is read.
  const unsigned short* slice_cnt; int nslices; uint32_t slice_cap; int64_t slice_q;
  int slice_gather;           // 1: read the slices' valid prefixes; 0: cand is sentinel-filled, scan it whole
  int sel_cap;                // candidate keys staged in shared memory per query (0 = kSelectStageKeys; <= 16384)
  // certification (tensor path): every row that is not a candidate has approx score >= cutoff
  int certify;                // 0 = none, 1 = tensor path (fp16 error bound; failures -> fb_list, re-run on the scan),
                              // 2 = scan path (fp32 summation-order bound; failures -> up_list, proven by the host)
  const float* thresh;        // per-query score threshold used by the tensor kernel
  int rerank_cut;             // 1: re-score only candidates whose error interval reaches the k-th best one's
  float eps_rel;              // relative error bound of the approximate dot product
  const float* qmaxabs;       // batch max |q| (fp16 common scale), tensor path
  float xnorm_max;            // max ||x|| over rows ever inserted
  uint32_t* fb_count; uint32_t* fb_list;  // uncertified queries are appended here
  uint64_t* out_ids; float* out_dist; uint32_t* out_counts; int64_t kstride;
  uint32_t* flags;
  uint32_t* uncertified;      // counter (stats)
  uint32_t* up_count; uint32_t* up_list; uint32_t up_base, up_cap;  // certify == 2: unproven queries (index + up_base)
};
cudaError_t launch_select_rerank(const SelectParams& p, int grid, cudaStream_t st);

// ---- K6: exact distances of explicit (query, row id) pairs (HNSW candidate evaluation) ------------
struct ScorePairsParams {
  IndexView iv;
  const float* q32; const float* qnorm;   // prepared queries [q][dpad], reference-exact norms
  const uint64_t* cand_ids;               // [q][m]
  int64_t q, m;
  float* out_dist;                        // [q][m]; +inf where status != 0
  uint8_t* out_status;                    // [q][m]: 0 ok, 1 id not in the index, 2 zero-norm cosine operand
  uint32_t* flags;
};
cudaError_t launch_score_pairs(const ScorePairsParams& p, cudaStream_t st);

// ---- K5: merge of per-shard results ---------------------------------------------------
cudaError_t launch_merge(const uint64_t* ids, const float* dist, const uint32_t* counts, int G, int64_t q,
                         int64_t kstride, int64_t gstride, const uint32_t* ks, uint64_t* out_ids, float* out_dist,
                         uint32_t* out_counts, int64_t out_kstride, cudaStream_t st);

// ---- K2: tcgen05 GEMM + fused threshold filter ----------------------------------------
struct GemmParams {
  IndexView iv;
  MaskView mask;
  const void* tmap_x;   // CUtensorMap (device-visible copy lives in kernel param space)
  const void* tmap_q;
  const float* qmaxabs; // batch max |q|
  const float* qsumsq;
  int q, num_m_tiles;
  int64_t num_n_tiles;
  // seed mode: per (sample tile, query) the R smallest scores; main mode: threshold filter
  int seed_mode; int64_t seed_tiles, seed_stride;
  float* seeds;         // [q][seed_tiles][2 column halves][R]
  const float* thresh;  // [q]
  // main mode: unit b (CTA or CTA pair), column half h appends to the slice [(2b+h)*cand_cap, (2b+h+1)*cand_cap)
  // of cand[q][cand_stride]; the number of keys of every (slice, query) goes to slice_cnt (no pre-fill needed)
  uint64_t* cand; uint32_t* cand_cnt; int64_t cand_stride; uint32_t cand_cap;
  unsigned short* slice_cnt;  // [slices][q]: written for every (slice, query) at the end of the main pass
  uint32_t* flags;
  const uint32_t* skip; // device word or null: non-zero = the device-side route chose the scan, exit at once
  int debug;            // timing experiments only (see gemm_topk.cu)
  int short_k;          // 1: short-K row-tile-stationary main pass (dpad16 <= 128; gemm_topk.cu, namespace sk): the
                        // grid's CTA b owns row tiles b, b + grid, ... and meets every query tile for each of them
  int pair;             // 1: CTA-pair kernel (cta_group::2); needs an even grid and an even number of query tiles;
                        // the row tensor map then has a 128-row box and slices are per PAIR: (pair*2+half)
};
constexpr int kSeedR = 8;
constexpr int kSelectStageKeys = 2048;  // candidate keys select_kernel stages in shared memory per query
constexpr int kGemmMaxQueries = 4096;   // per launch (u16 hit counters [2 column halves][query] in shared memory)
cudaError_t launch_gemm_topk(const GemmParams& p, const void* tmap_x_host, const void* tmap_q_host, int grid,
                             cudaStream_t st);
struct SeedFinalizeParams {
  const float* seeds; int q; int64_t seed_tiles; int rank; float* thresh;
  const uint32_t* skip;  // as GemmParams::skip
};
cudaError_t launch_seed_finalize(const SeedFinalizeParams& p, cudaStream_t st);

// ---- device-side route of a masked small batch (scan.cu) --------------------------------
// Seconds of a gather scan over `pop` eligible rows of `row_bytes` (measured on B200, 10M x 384: 5.3 TB/s of touched
// bytes at 50 % selectivity -- partially used DRAM pages -- plus ~0.06 ms of compaction, select and rerank).
__host__ __device__ inline double gather_scan_seconds(double pop, double row_bytes, double passes) {
  return passes * (pop * row_bytes / 5.3e12 + 0.06e-3);
}
// With a mask that lives on the device (caller's device pointer, device-evaluated MetadataFilter) the host does not
// know how many rows are eligible, so the choice between the gather scan (touches eligible fp32 rows only) and the
// masked tensor pass (streams every fp16 row) is made ON the device from the compaction's count: either the
// tensor kernels run as enqueued, or they exit at once and every query goes to the predicated scan that already
// follows them as the certification fallback.
struct RouteParams {
  const uint32_t* elig_count;  // written by compact_eligible_kernel
  int q;
  double row_bytes, passes, tensor_s;
  uint32_t* skip;        // out: 1 = scan route
  uint32_t* tensor_nq;   // out: queries the tensor path's select/rerank handle (q or 0)
  uint32_t* fb_count; uint32_t* fb_list;  // scan route: all q queries
  uint32_t* routed_scan; // stats
  uint32_t* elig_out;    // population + 1, read back with the call's control block
};
cudaError_t launch_route(const RouteParams& p, cudaStream_t st);

// ---- device-side MetadataFilter evaluation (filter.cu) ---------------------------------
constexpr int kFilterEq = 0, kFilterNe = 1, kFilterExists = 2, kFilterAnd = 3, kFilterOr = 4;
constexpr int kFilterMaxOps = 64, kFilterMaxDepth = 24;
struct FilterOp { int kind; int field; uint32_t code; };  // And/Or: code = number of children
struct FilterProgram { int n; FilterOp ops[kFilterMaxOps]; };  // postfix
// cols_len = slots covered by the synced columns; slots beyond it read as "field absent" (code 0)
cudaError_t launch_eval_filter(const FilterProgram& prog, const uint32_t* const* d_cols, int64_t cols_len,
                               int64_t n_slots, uint64_t* mask_words, cudaStream_t st);

// misc
cudaError_t launch_fill_u32(uint32_t* p, uint32_t v, int64_t n, cudaStream_t st);
cudaError_t launch_copy_words(uint32_t* dst, const uint32_t* src, int n, cudaStream_t st);
cudaError_t launch_gather_rows(const IndexView& src, const uint32_t* perm, int64_t n_out, float* x32,
                               __half* x16, uint64_t* ids, float* norm, float* sumsq, float2* coef,
                               cudaStream_t st);

}  // namespace gfi
