// host.h -- host-side data structures of libgfi shared by api.cu (single-device index, C ABI) and sharded.cu
// (one index over several GPUs of a box).  Nothing here crosses the C boundary (include/gfi.h does).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gfi.h"
#include "kernels.h"

namespace gfi {

extern thread_local std::string tl_error;
extern thread_local int64_t tl_expected, tl_actual;

int32_t fail(int32_t code, const std::string& msg);

#define CU_TRY(expr)                                                                             \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return ::gfi::fail(GFI_ERR_INDEX, std::string(#expr) + ": " + cudaGetErrorString(e__));    \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, need); want = need; }
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMallocHost(&p, need + need / 4 + 256);
    if (e == cudaSuccess) bytes = need + need / 4 + 256;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// control block of one search call (device memory, zeroed per call)
struct Ctrl {
  uint32_t flags;
  uint32_t unproven;    // queries whose scan-path answer could not be proven exact on the device (listed in up_list);
                        // like `flags`, sticky across uncollected device searches
  uint32_t fb_count;
  uint32_t uncertified;
  uint32_t qmaxabs_bits;
  // device-side route of a small batch with a device-resident mask (kernels.h, RouteParams)
  uint32_t skip_tensor;
  uint32_t tensor_nq;
  uint32_t routed_scan;
  uint32_t elig_count;  // population of the mask + 1 when a gather scan or the route kernel saw it, else 0
  uint32_t scan_done;   // CTAs of a fused-tail scan that have finished (ScanParams::done_ctr)
  uint32_t pad_[2];
};
static_assert(sizeof(Ctrl) == kCtrlWords * 4, "host result blocks reserve 64 bytes: Ctrl + the done word");

struct SearchCtx {
  cudaStream_t stream = nullptr;
  DevBuf q_in, q32, q16, qnorm, qsumsq, ks, mask, cand, cand_cnt, slice_cnt, cand_fb, cand_fb_cnt, thresh, seeds, ctrl,
      fb_list, up_list, floor, out_ids, out_dist, out_counts, sel_keys, sel_info, gather;
  PinBuf h_q, h_ks, h_ids, h_dist, h_counts, h_ctrl, h_out;
  bool zc_pending = false;   // the search enqueued last publishes its results in h_out itself (latency mode)
  uint32_t zc_seq = 0;
  DevBuf out_blk;            // [Ctrl | counts | dist | ids] of a host search: one D2H copy brings it all back
  void* ctrl_dev = nullptr;  // control block of the search being enqueued (inside out_blk or `ctrl`)
  bool pending_status = false;  // device search issued, status not yet collected
  cudaStream_t last_stream = nullptr;  // stream of the uncollected device search(es)
  cudaEvent_t order_ev = nullptr;      // orders a device search on a new stream behind the uncollected ones
  int64_t auto_tensor_q = 0;     // queries of the search being enqueued that took the tensor path by the cost model
  // CUDA-event pairs around the dominant kernel of each enqueued search (option "profile")
  struct EvPair { cudaEvent_t a, b; int kind; };
  std::vector<EvPair> evs;
  size_t ev_used = 0;
  void release() {
    for (DevBuf* b : {&q_in, &q32, &q16, &qnorm, &qsumsq, &ks, &mask, &cand, &cand_cnt, &slice_cnt, &cand_fb, &cand_fb_cnt,
                      &thresh, &seeds, &ctrl, &fb_list, &up_list, &floor, &out_ids, &out_dist, &out_counts, &sel_keys, &sel_info, &gather})
      b->release();
    for (PinBuf* b : {&h_q, &h_ks, &h_ids, &h_dist, &h_counts, &h_ctrl, &h_out}) b->release();
    out_blk.release();
    for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    evs.clear();
    ev_used = 0;
    if (order_ev) cudaEventDestroy(order_ev);
    order_ev = nullptr;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
};

struct Run {  // ids[slot0 + i] == id0 + i for i < n
  uint32_t slot0;
  uint32_t n;
};

struct ShardSet;  // sharded.cu: the shards, workers and exchange buffers of a multi-GPU index

// One shard's part of a search over a sharded index (sharded.cu), enqueued by that shard's worker thread on the
// shard's own GPU.  Inputs are either pinned host memory (every GPU of the box reads it by DMA) or device memory of
// the root GPU (read through NVLink); the three result arrays and the copy of the control block are DEVICE POINTERS
// INTO THE ROOT GPU'S MEMORY: the shard's finalize kernel stores its k candidates straight into the merging GPU's
// gather block over NVLink (peer stores), so the exchange needs no collective and no separate copy.
struct ShardSearch {
  const float* queries = nullptr;  // [q][dim]
  const uint32_t* ks = nullptr;    // [q]
  bool on_device = false;          // queries / ks / d_mask are device memory (root GPU), else pinned host memory
  int64_t q = 0, dim = 0;
  uint32_t kmax = 0;
  const uint64_t* mask = nullptr;  // eligibility by internal id (host memory unless on_device), or null
  int64_t mask_bits = 0;
  double mask_density = -1.0;      // eligible fraction of a host mask when known (cost model), else < 0
  const char* filter_json = nullptr;
  uint64_t* out_ids = nullptr;     // [q][kstride]
  float* out_dist = nullptr;
  uint32_t* out_counts = nullptr;
  int64_t kstride = 0;
  void* out_ctrl = nullptr;        // 64 bytes: receives the shard's control block (flags, fallback count, ...)
  cudaEvent_t wait_a = nullptr, wait_b = nullptr;  // the shard's stream first waits for these (inputs ready; gather
                                                   // block free again: the merge that read it last has finished)
  cudaEvent_t done = nullptr;      // recorded on the shard's stream behind everything above
  bool keep_flags = false;         // an earlier device search of this context is uncollected: flags stay sticky
};
// Enqueues it (no host synchronisation).  *idle is set when nothing was enqueued on failure.
int32_t shard_enqueue(gfi_index* shard, SearchCtx* c, const ShardSearch& s);
// Bookkeeping of a finished shard search from the host copy of its control block; event timings collected.
void shard_account(gfi_index* shard, SearchCtx* c, const Ctrl& hc);

SearchCtx* host_acquire_ctx(gfi_index* h);
void host_release_ctx(gfi_index* h, SearchCtx* c);
int32_t host_flags_to_status(uint32_t flags);
bool host_needs_flush(gfi_index* h, bool with_metadata);     // staged rows, dirty tombstones, pending reorder, ...
int32_t host_ensure_flushed(gfi_index* h, bool with_metadata);
int64_t host_row_bytes(const gfi_index* h);                  // fp32 bytes of the stored rows

// ---- sharded.cu: one index over several GPUs (entry points of the C ABI dispatch here when h->shards != null) ----
int32_t sharded_create(gfi_index** out, int32_t metric, int64_t dim, const int32_t* devices, int32_t n_devices,
                       uint32_t flags);
void sharded_destroy(gfi_index* h);
int32_t sharded_add(gfi_index* h, const uint64_t* ids, const float* rows, int64_t n, int64_t dim);
int32_t sharded_add_generated(gfi_index* h, uint32_t seed, uint64_t first_row, int64_t n, int32_t kind,
                              uint64_t first_id);
int32_t sharded_remove(gfi_index* h, uint64_t id);
int64_t sharded_len(const gfi_index* h);
int64_t sharded_dim(const gfi_index* h);
int32_t sharded_get_vector(gfi_index* h, uint64_t id, float* out, int64_t cap, int64_t* out_dim);
int32_t sharded_flush(gfi_index* h, bool compact);
int32_t sharded_reserve(gfi_index* h, int64_t n_rows);
int32_t sharded_search(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                       const uint64_t* mask, int64_t mask_bits, const char* filter_json, uint64_t* out_ids,
                       float* out_dist, uint32_t* out_counts, int64_t kstride);
int32_t sharded_search_big_k(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                             const uint64_t* mask, int64_t mask_bits, const char* filter_json, uint64_t* out_ids,
                             float* out_dist, uint32_t* out_counts, int64_t kstride);
int32_t sharded_search_device(gfi_index* h, const float* d_queries, int64_t q, const uint32_t* d_ks, uint32_t kmax,
                              const uint64_t* d_mask, int64_t mask_bits, uint64_t* d_out_ids, float* d_out_dist,
                              uint32_t* d_out_counts, int64_t kstride, void* stream);
int32_t sharded_search_status(gfi_index* h);
int32_t sharded_set_metadata(gfi_index* h, uint64_t id, int32_t n_fields, const char* const* keys,
                             const char* const* values);
int32_t sharded_set_metadata_column(gfi_index* h, const char* key, const uint64_t* ids, int64_t n,
                                    const char* const* values, int32_t n_values, const uint32_t* codes);
int32_t sharded_distances(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint64_t* cand_ids,
                          int64_t m, float* out_dist, uint8_t* out_status);
int32_t sharded_get_stats(gfi_index* h, gfi_stats* out);
int32_t sharded_set_option(gfi_index* h, const char* name, int64_t value);
int64_t sharded_row_bytes(const gfi_index* h);

}  // namespace gfi

// (the C ABI's opaque handle lives at global scope; its members are the gfi:: types above)
using gfi::DevBuf;
using gfi::PinBuf;
using gfi::Ctrl;
using gfi::SearchCtx;
using gfi::Run;
using gfi::IndexView;

struct gfi_index {
  gfi::ShardSet* shards = nullptr;  // non-null: this handle is a router over one sub-index per GPU (sharded.cu)
  int metric = 0;
  int64_t dim = 0;
  int dpad = 0, dpad16 = 0;
  int device = 0;
  uint32_t flags = 0;
  int sm_count = 148;

  mutable std::shared_mutex mu;  // writers: add/remove/flush/compact; readers: search
  std::mutex pool_mu;
  std::vector<std::unique_ptr<SearchCtx>> pool;

  // device storage
  int64_t cap = 0, n_slots = 0, n_live = 0;
  DevBuf x32, x16, ids, norm, sumsq, coef, live, rowflags, counters;
  cudaStream_t ingest_stream = nullptr;
  bool use_x16 = true;
  int64_t zero_rows_ever = 0, unsafe_rows_ever = 0;
  float xnorm_max = 0.f;

  // host bookkeeping
  std::map<uint64_t, Run> runs;          // id0 -> run (disjoint id ranges)
  std::vector<uint32_t> h_live;          // host mirror of the live bitmap
  int64_t live_dirty_lo = -1, live_dirty_hi = -1;
  bool ids_identity = true;
  bool needs_reorder = false;
  uint64_t max_id_seen = 0;
  bool any_id = false;
  std::unordered_map<uint64_t, int64_t> odd_dim_rows;  // id -> dim of rows whose dim != index dim

  // metadata (device-side filter evaluation): dictionary-encoded u32 column per field, code 0 = absent
  std::map<std::string, int> meta_fields;
  std::vector<std::map<std::string, uint32_t>> meta_values;
  std::vector<std::vector<uint32_t>> meta_cols;  // host mirror, one entry per slot
  std::vector<DevBuf> meta_dcols;
  DevBuf meta_dptrs;
  bool meta_dirty = false;
  int64_t meta_synced_slots = 0;  // slots the device columns cover (rows beyond read as "no metadata")
  std::unordered_map<uint64_t, std::vector<std::pair<int, uint32_t>>> meta_pending;

  // staging (pinned)
  PinBuf st_rows, st_ids;
  int64_t st_n = 0, st_cap = 0;

  // stats / options
  std::atomic<int64_t> n_search{0}, n_queries{0}, n_scan_q{0}, n_tensor_q{0}, n_fallback_q{0}, n_launch{0}, n_paged_q{0};
  int opt_tensor_min_q = 16;
  int opt_kp = 0;          // 0 = auto
  int opt_hits = 0;        // 0 = auto
  int opt_scan_qt = 0;     // 0 = auto
  int opt_fused_tail = 1;  // small single-pass scans finish inside the scan kernel (experiments: 0 = separate K3)
  int opt_zero_copy = 1;   // small host searches: pinned-host inputs/outputs, no copies, no stream sync (0 = off)
  int opt_grid = 0;        // 0 = sm_count
  int opt_tensor_min_rows = 8192;
  int opt_seed_rank = 8;
  int opt_pair = 0;          // 1: CTA-pair (cta_group::2) kernel for even query-tile counts.  Measured on B200: no
                             // gain -- the pass is power-limited either way -- so off by default
  int opt_short_k = 1;       // dpad16 <= 128: row-tile-stationary main pass (0 = the k-ring kernel, for A/B timing)
  int opt_short_k_min_tiles = 4;  // ... for batches of at least this many 128-query tiles
  int opt_short_k_seed = 1;  // the seed pass of a short-K search runs on the row-tile-stationary kernel too
  int opt_rerank_cut = 1;    // tensor path: re-score only candidates whose fp16 error interval reaches the k-th one's
  int opt_scan_certify = 1;  // 0: scan-path answers are not certified (A/B timing, tests of the proof itself)
  int opt_tensor_auto = 1;   // small batches on large indexes take the tensor path when the cost model says so
  std::atomic<int64_t> last_mask_pop{-1};     // population of the last device-resident mask searched (route predictor)
  std::atomic<bool> auto_tensor_off{false};   // set when such searches keep falling back (uncertifiable data)
  std::atomic<int64_t> auto_q{0}, auto_fb{0};
  int opt_raw_epilogue = 1;  // 0: force the per-row-coefficient epilogue for cosine (A/B timing, tests)
  std::atomic<uint64_t> layout_gen{0};  // bumped whenever rows change slots (compaction)
  int opt_profile = 0;
  // Group commit of concurrent plain searches (micro-batching of concurrent Index::search
  // calls).  While one batch runs, arriving calls queue up; the next leader takes everything queued as ONE batch
  // (one scan pass serves up to 4 queries, >= 16 go to the tensor path).  A call that finds the index idle runs at
  // once, alone: no timers, no added latency.
  int opt_coalesce = 1;
  struct CoReq {
    const float* queries; int64_t q, dim; const uint32_t* ks;
    uint64_t* out_ids; float* out_dist; uint32_t* out_counts; int64_t kstride;
    int32_t rc = 0; std::string err; int64_t exp = 0, act = 0;
    bool done = false, lead = false, answered = false;
    std::vector<CoReq*> batch;  // filled for the request promoted to leader
    std::condition_variable cv;  // one per request: the leader wakes exactly the requests it finished and its successor
  };
  std::mutex co_mu;
  bool co_busy = false;
  std::vector<CoReq*> co_pending;
  std::atomic<int64_t> n_co_batches{0}, n_co_requests{0};
  int opt_gemm_debug = 0;
  int opt_scan_stages = 0;
  std::atomic<int64_t> prof_ns[2] = {{0}, {0}}, prof_cnt[2] = {{0}, {0}};

  IndexView view() const {
    IndexView v;
    v.x32 = x32.as<float>();
    v.x16 = use_x16 ? x16.as<__half>() : nullptr;
    v.ids = ids.as<uint64_t>();
    v.norm = norm.as<float>();
    v.sumsq = sumsq.as<float>();
    v.coef = use_x16 ? coef.as<float2>() : nullptr;
    v.live = live.as<uint32_t>();
    v.n_slots = n_slots;
    v.d = (int)dim;
    v.dpad = dpad;
    v.dpad16 = dpad16;
    v.metric = metric;
    v.ids_identity = ids_identity ? 1 : 0;
    return v;
  }
};
