// api.cu -- host runtime of libgfi.so and its C ABI (include/gfi.h).
//
// GpuFlatIndex keeps the database resident in HBM as a dense, append-ordered slot array
// (slot order == internal-id order, tombstones for removed ids):
//     x32  [slots][dpad]    fp32 rows (exact scan + reference-exact rerank)
//     x16  [slots][dpad16]  per-row power-of-two scaled fp16 shadow (tcgen05 candidate pass)
//     ids / norm / sumsq / coef / live bitmap
// Adds are staged in pinned host memory and flushed lazily; searches borrow a context
// (stream + workspace) from a pool so concurrent `&self` callers never share state.
#include <cuda.h>
#include <cuda_runtime.h>
#include <limits>
#include <chrono>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <sys/stat.h>
#include <unistd.h>

#include "host.h"
#include "exact.cuh"  // scan_lower_bound (host side of the scan-path certification)

using namespace gfi;

namespace gfi {
thread_local std::string tl_error;
thread_local int64_t tl_expected = 0, tl_actual = 0;
int32_t fail(int32_t code, const std::string& msg) {
  tl_error = msg;
  return code;
}
}  // namespace gfi

namespace {

thread_local bool tl_mask_by_slot = false;  // search_impl: the host mask passed in is indexed by slot (big-k passes)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace


namespace {

int32_t set_device(const gfi_index* h) {
  CU_TRY(cudaSetDevice(h->device));
  return GFI_OK;
}

void latch_dim(gfi_index* h, int64_t dim) {
  h->dim = dim;
  h->dpad = (int)((dim + 3) / 4 * 4);
  h->dpad16 = (int)((dim + 7) / 8 * 8);
}

// grow device storage to hold at least `need` rows (keeps contents)
int32_t grow(gfi_index* h, int64_t need) {
  if (need <= h->cap) return GFI_OK;
  int64_t ncap = std::max<int64_t>(need, h->cap + h->cap / 2);
  ncap = std::max<int64_t>(ncap, 1024);
  ncap = (ncap + 255) / 256 * 256;
  if (ncap > 0xFFFFFFF0ll) return fail(GFI_ERR_INDEX, "too many rows for 32-bit slots");
  struct Item { DevBuf* b; size_t per_row; };
  std::vector<Item> items = {{&h->x32, (size_t)h->dpad * 4}, {&h->ids, 8}, {&h->norm, 4}, {&h->sumsq, 4},
                             {&h->rowflags, 4}};
  if (h->use_x16) {
    items.push_back({&h->x16, (size_t)h->dpad16 * 2});
    items.push_back({&h->coef, 8});
  }
  for (auto& it : items) {
    void* np = nullptr;
    size_t nbytes = it.per_row * (size_t)ncap + 256;
    CU_TRY(cudaMalloc(&np, nbytes));
    if (h->n_slots > 0 && it.b->p)
      CU_TRY(cudaMemcpyAsync(np, it.b->p, it.per_row * (size_t)h->n_slots, cudaMemcpyDeviceToDevice,
                             h->ingest_stream));
    CU_TRY(cudaStreamSynchronize(h->ingest_stream));
    if (it.b->p) cudaFree(it.b->p);
    it.b->p = np;
    it.b->bytes = nbytes;
  }
  {
    size_t words = (size_t)(ncap + 31) / 32;
    void* np = nullptr;
    CU_TRY(cudaMalloc(&np, words * 4 + 256));
    CU_TRY(cudaMemsetAsync(np, 0, words * 4 + 256, h->ingest_stream));
    if (h->live.p && h->n_slots > 0)
      CU_TRY(cudaMemcpyAsync(np, h->live.p, ((size_t)h->n_slots + 31) / 32 * 4, cudaMemcpyDeviceToDevice,
                             h->ingest_stream));
    CU_TRY(cudaStreamSynchronize(h->ingest_stream));
    if (h->live.p) cudaFree(h->live.p);
    h->live.p = np;
    h->live.bytes = words * 4 + 256;
    h->h_live.resize(words, 0u);
  }
  h->cap = ncap;
  return GFI_OK;
}

bool lookup_slot(const gfi_index* h, uint64_t id, uint32_t* slot) {
  auto it = h->runs.upper_bound(id);
  if (it == h->runs.begin()) return false;
  --it;
  const uint64_t off = id - it->first;
  if (off >= it->second.n) return false;
  const uint32_t s = it->second.slot0 + (uint32_t)off;
  if (!((h->h_live[s >> 5] >> (s & 31)) & 1u)) return false;
  *slot = s;
  return true;
}

void mark_live_dirty(gfi_index* h, int64_t word) {
  if (h->live_dirty_lo < 0 || word < h->live_dirty_lo) h->live_dirty_lo = word;
  if (word > h->live_dirty_hi) h->live_dirty_hi = word;
}

void tombstone(gfi_index* h, uint32_t slot) {
  h->h_live[slot >> 5] &= ~(1u << (slot & 31));
  mark_live_dirty(h, slot >> 5);
  --h->n_live;
}

// registers ids for slots [slot0, slot0+n): consecutive ids extend the last run
void register_ids(gfi_index* h, const uint64_t* ids, uint64_t first_id, int64_t n, uint32_t slot0) {
  int64_t i = 0;
  while (i < n) {
    const uint64_t id = ids ? ids[i] : first_id + (uint64_t)i;
    int64_t len = 1;
    if (ids) {
      while (i + len < n && ids[i + len] == id + (uint64_t)len) ++len;
    } else {
      len = n - i;
    }
    if (h->any_id && id <= h->max_id_seen) h->needs_reorder = true;  // out of order / overwrite
    if (id != (uint64_t)(slot0 + i)) h->ids_identity = false;
    // extend the previous run if contiguous in both id and slot space
    bool merged = false;
    if (!h->runs.empty()) {
      auto last = std::prev(h->runs.end());
      if (last->first + last->second.n == id && last->second.slot0 + last->second.n == slot0 + (uint32_t)i) {
        last->second.n += (uint32_t)len;
        merged = true;
      }
    }
    if (!merged) h->runs[id] = Run{slot0 + (uint32_t)i, (uint32_t)len};
    h->max_id_seen = std::max(h->max_id_seen, id + (uint64_t)len - 1);
    h->any_id = true;
    for (int64_t j = 0; j < len; ++j) {
      const uint32_t s = slot0 + (uint32_t)(i + j);
      h->h_live[s >> 5] |= 1u << (s & 31);
    }
    mark_live_dirty(h, (slot0 + i) >> 5);
    mark_live_dirty(h, (slot0 + i + len - 1) >> 5);
    h->n_live += len;
    i += len;
  }
}

// removes `id` from the run map (splitting its run) and tombstones its slot; false if absent
bool erase_id(gfi_index* h, uint64_t id) {
  auto it = h->runs.upper_bound(id);
  if (it == h->runs.begin()) return false;
  --it;
  const uint64_t id0 = it->first;
  const Run r = it->second;
  const uint64_t off = id - id0;
  if (off >= r.n) return false;
  const uint32_t s = r.slot0 + (uint32_t)off;
  const bool was_live = (h->h_live[s >> 5] >> (s & 31)) & 1u;
  h->runs.erase(it);
  if (off > 0) h->runs[id0] = Run{r.slot0, (uint32_t)off};
  if (off + 1 < r.n) h->runs[id + 1] = Run{s + 1, (uint32_t)(r.n - off - 1)};
  if (was_live) tombstone(h, s);
  return was_live;
}

int32_t upload_live(gfi_index* h) {
  if (h->live_dirty_lo < 0) return GFI_OK;
  const int64_t lo = h->live_dirty_lo, hi = h->live_dirty_hi;
  CU_TRY(cudaMemcpyAsync(h->live.as<uint32_t>() + lo, h->h_live.data() + lo, (size_t)(hi - lo + 1) * 4,
                         cudaMemcpyHostToDevice, h->ingest_stream));
  CU_TRY(cudaStreamSynchronize(h->ingest_stream));
  h->live_dirty_lo = h->live_dirty_hi = -1;
  return GFI_OK;
}

int32_t read_counters(gfi_index* h) {
  struct { uint32_t zero, unsafe, maxnorm_bits, pad; } c;
  CU_TRY(cudaMemcpyAsync(&c, h->counters.p, sizeof(c), cudaMemcpyDeviceToHost, h->ingest_stream));
  CU_TRY(cudaStreamSynchronize(h->ingest_stream));
  h->zero_rows_ever = c.zero;
  h->unsafe_rows_ever = c.unsafe;
  float m;
  memcpy(&m, &c.maxnorm_bits, 4);
  h->xnorm_max = m;
  return GFI_OK;
}

}  // namespace

// implemented in ingest.cu (kept here to avoid widening kernels.h): folds per-row flags into counters
namespace gfi {
bool compile_filter(const char* json, const std::map<std::string, int>& fields,
                    const std::vector<std::map<std::string, uint32_t>>& values, FilterProgram* out, std::string* err);
cudaError_t launch_fold_rowflags(const uint32_t* rowflags, const float* norm, int64_t first, int64_t n,
                                 uint32_t* counters, cudaStream_t st);
}

namespace {

int32_t ingest_slots(gfi_index* h, int64_t first_slot, int64_t n, bool gen, uint32_t seed, uint64_t first_row,
                     int kind) {
  IngestParams p{};
  p.x32 = h->x32.as<float>();
  p.x16 = h->use_x16 ? h->x16.as<__half>() : nullptr;
  p.norm = h->norm.as<float>();
  p.sumsq = h->sumsq.as<float>();
  p.coef = h->use_x16 ? h->coef.as<float2>() : nullptr;
  p.zero_flags = h->rowflags.as<uint32_t>();
  p.first_slot = first_slot;
  p.n = n;
  p.d = (int)h->dim;
  p.dpad = h->dpad;
  p.dpad16 = h->dpad16;
  p.metric = h->metric;
  p.gen = gen ? 1 : 0;
  p.seed = seed;
  p.first_row = first_row;
  p.kind = kind;
  CU_TRY(launch_ingest(p, h->ingest_stream));
  CU_TRY(launch_fold_rowflags(h->rowflags.as<uint32_t>(), h->norm.as<float>(), first_slot, n,
                              h->counters.as<uint32_t>(), h->ingest_stream));
  h->n_launch += gen ? 3 : 2;
  return GFI_OK;
}

// flush staged rows: H2D + row_stats.  Caller holds the unique lock.
int32_t flush_locked(gfi_index* h) {
  int32_t rc;
  if ((rc = set_device(h)) != GFI_OK) return rc;
  if (h->st_n > 0) {
    const int64_t n = h->st_n;
    if ((rc = grow(h, h->n_slots + n)) != GFI_OK) return rc;
    const uint32_t slot0 = (uint32_t)h->n_slots;
    CU_TRY(cudaMemcpyAsync(h->x32.as<float>() + (size_t)slot0 * h->dpad, h->st_rows.p,
                           (size_t)n * h->dpad * 4, cudaMemcpyHostToDevice, h->ingest_stream));
    CU_TRY(cudaMemcpyAsync(h->ids.as<uint64_t>() + slot0, h->st_ids.p, (size_t)n * 8, cudaMemcpyHostToDevice,
                           h->ingest_stream));
    // overwrite semantics of FlatIndex::add (HashMap::insert): drop any previous row with the same id
    const uint64_t* sid = h->st_ids.as<uint64_t>();
    for (int64_t i = 0; i < n; ++i) {
      if (h->any_id && sid[i] <= h->max_id_seen) erase_id(h, sid[i]);
      h->odd_dim_rows.erase(sid[i]);
    }
    h->n_slots += n;
    register_ids(h, sid, 0, n, slot0);
    if ((rc = ingest_slots(h, slot0, n, false, 0, 0, 0)) != GFI_OK) return rc;
    CU_TRY(cudaStreamSynchronize(h->ingest_stream));
    h->st_n = 0;
    if (!h->meta_values.empty()) h->meta_dirty = true;  // the per-field columns must grow with the slot array
    if ((rc = read_counters(h)) != GFI_OK) return rc;
  }
  if ((rc = upload_live(h)) != GFI_OK) return rc;
  return GFI_OK;
}

int32_t compact_locked(gfi_index* h);

SearchCtx* acquire_ctx(gfi_index* h) {
  std::lock_guard<std::mutex> g(h->pool_mu);
  if (!h->pool.empty()) {
    SearchCtx* c = h->pool.back().release();
    h->pool.pop_back();
    return c;
  }
  SearchCtx* c = new SearchCtx();
  // (the stream belongs to the index's GPU whatever device the calling thread has current)
  if (cudaSetDevice(h->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return nullptr;
  }
  return c;
}

void release_ctx(gfi_index* h, SearchCtx* c) {
  std::lock_guard<std::mutex> g(h->pool_mu);
  h->pool.emplace_back(c);
}

int pow2_at_least(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

bool make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_pitch_bytes,
                  uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

void prof_begin(gfi_index* h, SearchCtx* c, int kind, cudaStream_t st) {
  if (!h->opt_profile) return;
  if (c->ev_used == c->evs.size()) {
    if (c->evs.size() >= 4096) return;  // collect with gfi_search_status before issuing more
    SearchCtx::EvPair e{nullptr, nullptr, kind};
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    c->evs.push_back(e);
  }
  c->evs[c->ev_used].kind = kind;
  cudaEventRecord(c->evs[c->ev_used].a, st);
}
void prof_end(gfi_index* h, SearchCtx* c, cudaStream_t st) {
  if (!h->opt_profile || c->ev_used >= c->evs.size()) return;
  cudaEventRecord(c->evs[c->ev_used].b, st);
  ++c->ev_used;
}
// call after the stream has been synchronised
void prof_collect(gfi_index* h, SearchCtx* c) {
  for (size_t i = 0; i < c->ev_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->evs[i].a, c->evs[i].b) == cudaSuccess) {
      h->prof_ns[c->evs[i].kind] += (int64_t)((double)ms * 1e6);
      h->prof_cnt[c->evs[i].kind] += 1;
    }
  }
  c->ev_used = 0;
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

struct SearchArgs {
  const float* d_queries;  // device, [q][dim] unpadded
  int64_t q;
  const uint32_t* d_ks;
  uint32_t kmax;
  const uint64_t* d_mask;
  int64_t mask_bits;
  uint64_t* d_out_ids;
  float* d_out_dist;
  uint32_t* d_out_counts;
  int64_t kstride;
  // latency mode (small host searches, see ScanParams::h_ctrl): pinned host views of ks and of the result block
  const uint32_t* h_ks_in = nullptr;
  uint32_t* h_ctrl = nullptr;
  uint64_t* h_out_ids = nullptr;
  float* h_out_dist = nullptr;
  uint32_t* h_out_counts = nullptr;
  uint32_t done_seq = 0;
  bool mask_by_slot = false;  // the mask is indexed by slot (device-evaluated filter), not by internal id
  int64_t mask_popcount = -1;  // eligible bits of a host mask when known (cost model), else -1
  bool keep_flags = false;     // an earlier search of this context has not been collected yet: its error flags stay
  int64_t q_base = 0, q_total = 0;  // chunked batches: index of this chunk's first query, queries of the whole call
};

// Ring / list geometry of the scan kernel for per-query lists of K keys and a batch of q queries.
struct ScanGeom { int R, segf, nseg, lpr, stage_floats, QT, nstages; };
int32_t scan_geometry(const gfi_index* h, int K, int q, ScanGeom* g) {
  // A stage is one contiguous bulk copy of whole rows whenever rows fit (<= 4 KB): rows <= 1 KB use 8 lanes per
  // row (64 rows per round of the 16 consumer warps), longer rows one warp per row (16 rows per round); rows
  // beyond 4 KB are cut into 4 KB column segments.
  int R, segf, nseg, lpr;
  const int kTargetStageBytes = 48 * 1024;
  if (h->dpad <= 256) {
    lpr = 8;
    segf = h->dpad;
    nseg = 1;
    R = 64 * std::max(1, std::min(4, kTargetStageBytes / (64 * h->dpad * 4)));
  } else if (h->dpad <= 512) {
    lpr = 16;  // two rows per warp at a time: halves the per-row reduction/bookkeeping cost of mid-size rows
    segf = h->dpad;
    nseg = 1;
    R = 32 * std::max(1, std::min(4, kTargetStageBytes / (32 * h->dpad * 4)));
  } else {
    lpr = 32;
    segf = std::min(h->dpad, 1024);
    nseg = (h->dpad + segf - 1) / segf;
    R = nseg == 1 ? 16 * std::max(1, std::min(4, kTargetStageBytes / (16 * h->dpad * 4))) : 16;
  }
  const int stage_floats = R * (nseg == 1 ? h->dpad : segf);
  // queries per pass: as many as shared memory allows next to a ring of at least 3 stages
  int QT = 4, nstages = 0;
  for (;; QT >>= 1) {
    const size_t fixed = (size_t)QT * h->dpad * 4 + (size_t)QT * kScanConsumerWarps * K * 8 + 1024;
    const size_t budget = 227 * 1024;
    nstages = fixed < budget ? (int)std::min<size_t>(kScanMaxStages, (budget - fixed) / ((size_t)stage_floats * 4)) : 0;
    if (nstages >= 3 || QT == 1) break;
  }
  if (nstages < 2) return fail(GFI_ERR_INDEX, "k and dimension too large for the scan kernel's shared memory");
  if (h->opt_scan_qt > 0) QT = std::min(QT, pow2_at_least(h->opt_scan_qt));
  while (QT > 1 && QT / 2 >= q) QT >>= 1;  // (also on the tensor path: its fallback never has more than q queries)
  {
    const size_t fixed = (size_t)QT * h->dpad * 4 + (size_t)QT * kScanConsumerWarps * K * 8 + 1024;
    nstages = (int)std::min<size_t>(kScanMaxStages, (227 * 1024 - fixed) / ((size_t)stage_floats * 4));
    if (h->opt_scan_stages > 0) nstages = std::max(2, std::min(nstages, h->opt_scan_stages));
  }
  *g = ScanGeom{R, segf, nseg, lpr, stage_floats, QT, nstages};
  return GFI_OK;
}

// Enqueues the whole device pipeline of one search batch on `st`.  No host synchronisation.
int32_t enqueue_chunk(gfi_index* h, SearchCtx* c, const SearchArgs& a, cudaStream_t st, bool first_chunk, bool last_chunk) {
  const int q = (int)a.q;
  const int grid_sm = h->opt_grid > 0 ? h->opt_grid : h->sm_count;
  IndexView iv = h->view();
  if (a.mask_by_slot) iv.ids_identity = 1;  // only mask indexing looks at this flag
  MaskView mv{a.d_mask, a.mask_bits};

  // Small batches on LARGE indexes also take the tensor path: both paths are then HBM-bound, and the fp16 shadow is
  // half the bytes of the fp32 rows (measured, q = 1: 10M x 768 4.21 -> 2.40 ms, 10M x 384 2.20 -> 1.16 ms).  The
  // cost model compares streaming times; the tensor path's fixed passes (seed, select, rerank) are ~0.25 ms.
  // With a caller mask the scan gathers only the eligible rows (adjacent rows merged into one copy) while the tensor pass
  // still streams every fp16 row, so the comparison needs the mask's population (known for host masks; a
  // device-evaluated filter keeps the scan).  Only while the index's small-batch searches keep being certified (an
  // uncertified query pays for both paths).
  bool small_q_tensor = false;
  bool device_route = false;  // the population of the mask is only known on the device: route there
  if (h->opt_tensor_auto && !h->auto_tensor_off && q < h->opt_tensor_min_q && h->opt_tensor_min_q <= kGemmMaxQueries) {
    // streaming rates and fixed costs as measured on B200: fp32 scan 7.2 TB/s + 0.05 ms,
    // gather scan 5.3 TB/s of touched bytes + 0.06 ms, tensor pass over fp16 rows 7.0 TB/s + 0.12 ms
    const double rows = (double)h->n_slots;
    const double passes = std::ceil(q / 4.0);
    const double tensor_s = rows * h->dpad16 * 2.0 / 7.0e12 + 0.12e-3;
    if (a.d_mask != nullptr && a.mask_popcount < 0) {
      // worth a route kernel only where the tensor pass could win at all (every row eligible)
      device_route = tensor_s < 0.9 * gather_scan_seconds(rows, h->dpad * 4.0, passes);
      // ... and not when the previous device-resident mask was so sparse that the scan won by a wide margin: the
      // tensor kernels are then not even enqueued (~30 us of launches that would exit at once).  A wrong guess
      // costs time only, and every masked search reports its population for the next one.
      const int64_t last = h->last_mask_pop.load();
      if (last >= 0 && !(tensor_s < 1.2 * gather_scan_seconds((double)last, h->dpad * 4.0, passes))) device_route = false;
      small_q_tensor = device_route;
    } else {
      const double pop = (double)std::min<int64_t>(std::max<int64_t>(a.mask_popcount, 0), h->n_slots);
      const double scan_s = a.d_mask == nullptr ? passes * (rows * h->dpad * 4.0 / 7.2e12 + 0.05e-3)
                                                : gather_scan_seconds(pop, h->dpad * 4.0, passes);
      small_q_tensor = tensor_s < 0.9 * scan_s;
    }
  }
  const bool tensor_ok = h->use_x16 && !(h->flags & GFI_FLAG_NO_TENSOR) && h->unsafe_rows_ever == 0 &&
                         !(h->metric == GFI_METRIC_COSINE && h->zero_rows_ever > 0) &&
                         (q >= h->opt_tensor_min_q || small_q_tensor) && q <= kGemmMaxQueries &&
                         h->n_slots >= h->opt_tensor_min_rows && a.kmax <= 256 && h->dpad16 >= 64 &&
                         get_encode_fn() != nullptr;
  c->auto_tensor_q = (tensor_ok && q < h->opt_tensor_min_q) ? q : 0;
  device_route = device_route && tensor_ok && q < h->opt_tensor_min_q;

  // ---- workspace ----
  // query tiles of 128; with the CTA-pair kernel enabled, an even number of them (the last may be all padding)
  const int qpad = (h->opt_pair && q > 128) ? (q + 255) / 256 * 256 : (q + 127) / 128 * 128;
  CU_TRY(c->q32.ensure((size_t)q * h->dpad * 4));
  CU_TRY(c->qnorm.ensure((size_t)q * 4));
  CU_TRY(c->qsumsq.ensure((size_t)q * 4));
  if (!c->ctrl_dev) {
    CU_TRY(c->ctrl.ensure(sizeof(Ctrl)));
    c->ctrl_dev = c->ctrl.p;
  }
  CU_TRY(c->fb_list.ensure((size_t)q * 4));
  CU_TRY(c->up_list.ensure((size_t)std::max<int64_t>(a.q_total, q) * 4));
  CU_TRY(c->sel_keys.ensure((size_t)q * 1024 * 8));
  CU_TRY(c->sel_info.ensure((size_t)q * sizeof(SelInfo)));
  if (tensor_ok) CU_TRY(c->q16.ensure((size_t)qpad * h->dpad16 * 2));
  if (first_chunk) {
    // (the flags word is the first of the block; it is sticky across uncollected device searches)
    const size_t skip = a.keep_flags ? 2 * sizeof(uint32_t) : 0;  // flags, unproven
    CU_TRY(cudaMemsetAsync(static_cast<char*>(c->ctrl_dev) + skip, 0, sizeof(Ctrl) - skip, st));
  }
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(c->ctrl_dev);

  PrepQueriesParams pq{};
  pq.q_in = a.d_queries;
  pq.q32 = c->q32.as<float>();
  pq.q16 = tensor_ok ? c->q16.as<__half>() : nullptr;
  pq.qnorm = c->qnorm.as<float>();
  pq.qsumsq = c->qsumsq.as<float>();
  pq.qmaxabs = reinterpret_cast<float*>(&ctrl->qmaxabs_bits);
  pq.q = q;
  pq.qpad = qpad;
  pq.d = (int)h->dim;
  pq.dpad = h->dpad;
  pq.dpad16 = h->dpad16;
  pq.ks_in = a.h_ks_in;
  pq.ks_out = a.h_ks_in ? const_cast<uint32_t*>(a.d_ks) : nullptr;
  CU_TRY(launch_prep_queries(pq, st));
  h->n_launch += tensor_ok ? 2 : 1;

  // ---- scan geometry (also used by the tensor path's fallback) ----
  const int K = std::min(1024, pow2_at_least(std::max<int>(32, (int)a.kmax + 8)));
  if ((int)a.kmax > K) return fail(GFI_ERR_INDEX, "k too large: the scan path supports k <= 1024");
  if (h->dpad > 16384) return fail(GFI_ERR_INDEX, "dimension too large (max 16384)");
  ScanGeom geo;
  {
    int32_t grc = scan_geometry(h, K, q, &geo);
    if (grc != GFI_OK) return grc;
  }
  const int R = geo.R, segf = geo.segf, nseg = geo.nseg, lpr = geo.lpr, stage_floats = geo.stage_floats;
  const int QT = geo.QT, nstages = geo.nstages;
  // filtered or heavily tombstoned scans gather only the eligible rows (one bulk copy per run of adjacent eligible rows)
  const bool gather = a.d_mask != nullptr || h->n_live * 10 < h->n_slots * 9;
  const int64_t nblocks = (h->n_slots + R - 1) / R;
  if (gather) {
    CU_TRY(c->gather.ensure(((size_t)h->n_slots + 8) * 4));
    CU_TRY(cudaMemsetAsync(c->gather.p, 0, 16, st));
    CU_TRY(launch_compact_eligible(iv, mv, c->gather.as<uint32_t>() + 4, c->gather.as<uint32_t>(), st));
    ++h->n_launch;
  }
  if (device_route) {
    RouteParams rp{};
    rp.elig_count = c->gather.as<uint32_t>();
    rp.q = q;
    rp.row_bytes = h->dpad * 4.0;
    rp.passes = std::ceil(q / 4.0);
    rp.tensor_s = (double)h->n_slots * h->dpad16 * 2.0 / 7.0e12 + 0.12e-3;
    rp.skip = &ctrl->skip_tensor;
    rp.tensor_nq = &ctrl->tensor_nq;
    rp.fb_count = &ctrl->fb_count;
    rp.fb_list = c->fb_list.as<uint32_t>();
    rp.routed_scan = &ctrl->routed_scan;
    rp.elig_out = &ctrl->elig_count;
    CU_TRY(launch_route(rp, st));
    ++h->n_launch;
  }
  const int scan_grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid_sm, nblocks));
  const int64_t scan_stride = (int64_t)scan_grid * K;

  auto fill_scan = [&](ScanParams& sp, DevBuf& cand, DevBuf& cnt) {
    sp.iv = iv;
    sp.mask = mv;
    sp.q32 = c->q32.as<float>();
    sp.qnorm = c->qnorm.as<float>();
    sp.K = K;
    sp.floor64 = nullptr;
    sp.cand = cand.as<uint64_t>();
    sp.cand_cnt = cnt.as<uint32_t>();
    sp.cand_stride = scan_stride;
    sp.flags = &ctrl->flags;
    sp.gather_list = gather ? c->gather.as<uint32_t>() + 4 : nullptr;
    sp.gather_count = gather ? c->gather.as<uint32_t>() : nullptr;
    sp.elig_out = (gather && a.d_mask != nullptr) ? &ctrl->elig_count : nullptr;
    sp.rows_per_stage = R;
    sp.seg_floats = segf;
    sp.nseg = nseg;
    sp.lanes_per_row = lpr;
    sp.nstages = nstages;
    sp.stage_floats = stage_floats;
    if (h->opt_scan_certify) {
      sp.up_count = &ctrl->unproven;
      sp.up_list = c->up_list.as<uint32_t>();
      sp.up_base = (uint32_t)a.q_base;
      sp.up_cap = (uint32_t)std::max<int64_t>(a.q_total, q);
      sp.xnorm_max = h->xnorm_max;
    }
  };
  auto fill_select = [&](SelectParams& s, DevBuf& cand, DevBuf& cnt, int64_t stride, int KP) {
    s.iv = iv;
    s.sel_keys = c->sel_keys.as<uint64_t>();
    s.sel_info = c->sel_info.as<SelInfo>();
    s.nq_max = q;
    s.q32 = c->q32.as<float>();
    s.qnorm = c->qnorm.as<float>();
    s.qsumsq = c->qsumsq.as<float>();
    s.ks = a.d_ks;
    s.cand = cand.as<uint64_t>();
    s.cand_cnt = cnt.as<uint32_t>();
    s.cand_stride = stride;
    s.KP = KP;
    s.qmaxabs = reinterpret_cast<const float*>(&ctrl->qmaxabs_bits);
    s.xnorm_max = h->xnorm_max;
    s.fb_count = &ctrl->fb_count;
    s.fb_list = c->fb_list.as<uint32_t>();
    s.out_ids = a.d_out_ids;
    s.out_dist = a.d_out_dist;
    s.out_counts = a.d_out_counts;
    s.kstride = a.kstride;
    s.flags = &ctrl->flags;
    s.uncertified = &ctrl->uncertified;
    s.up_count = &ctrl->unproven;
    s.up_list = c->up_list.as<uint32_t>();
    s.up_base = (uint32_t)a.q_base;
    s.up_cap = (uint32_t)std::max<int64_t>(a.q_total, q);
  };

  if (!tensor_ok) {
    CU_TRY(c->cand.ensure((size_t)q * scan_stride * 8));
    CU_TRY(c->cand_cnt.ensure((size_t)q * 4));
    ScanParams sp{};
    fill_scan(sp, c->cand, c->cand_cnt);
    sp.qlist = nullptr;
    sp.nq_dev = nullptr;
    sp.nq = q;
    // small single-pass searches: the scan's last CTA does the select / rerank / emit itself (scan.cu)
    const bool fused = h->opt_fused_tail && q <= QT && nseg == 1 && K <= 64 &&
                       scan_fused_tail_bytes(scan_grid, K, h->dpad) <= (size_t)nstages * stage_floats * 4;
    if (fused) {
      sp.fused = 1;
      sp.done_ctr = &ctrl->scan_done;
      sp.ks = a.d_ks;
      sp.out_ids = a.d_out_ids;
      sp.out_dist = a.d_out_dist;
      sp.out_counts = a.d_out_counts;
      sp.kstride = a.kstride;
      if (a.h_ctrl) {
        sp.h_ctrl = a.h_ctrl;
        sp.d_ctrl = reinterpret_cast<const uint32_t*>(ctrl);
        sp.h_out_ids = a.h_out_ids;
        sp.h_out_dist = a.h_out_dist;
        sp.h_out_counts = a.h_out_counts;
        sp.done_seq = a.done_seq;
        c->zc_pending = true;
      }
    }
    prof_begin(h, c, 0, st);
    CU_TRY(launch_scan(sp, QT, scan_grid, st));
    prof_end(h, c, st);
    if (fused) {
      h->n_launch += 1;
      h->n_scan_q += q;
      return GFI_OK;
    }
    SelectParams s{};
    fill_select(s, c->cand, c->cand_cnt, scan_stride, K);
    s.qlist = nullptr;
    s.nq_dev = nullptr;
    s.nq = q;
    s.certify = h->opt_scan_certify ? 2 : 0;
    s.list_len = K;
    CU_TRY(launch_select_rerank(s, std::min(q, grid_sm * 8), st));
    h->n_launch += 3;
    h->n_scan_q += q;
    return GFI_OK;
  }

  // ---- tensor path: seed thresholds -> tcgen05 filter -> select/rerank/certify -> scan fallback ----
  const int KP = h->opt_kp > 0 ? std::min(1024, pow2_at_least(h->opt_kp))
                               : std::min(1024, pow2_at_least(std::max<int>(64, 4 * (int)a.kmax)));
  const int hits = h->opt_hits > 0 ? h->opt_hits : 8 * KP;
  int main_grid = (int)std::min<int64_t>(grid_sm, ((h->n_slots + 255) / 256) * (qpad / 128));
  // rows of at most 128 fp16 columns and a batch of several query tiles: the row-tile-stationary main pass
  // (gemm_topk.cu, namespace sk) -- a CTA owns whole row tiles and meets every query tile for each of them
  const bool short_k = h->opt_short_k && !h->opt_pair && h->dpad16 <= 128 && qpad / 128 >= h->opt_short_k_min_tiles;
  if (short_k) main_grid = (int)std::min<int64_t>(grid_sm, (h->n_slots + 255) / 256);
  // CTA pairs (tcgen05 cta_group::2) when the batch has an even number of 128-query tiles: each SM then takes in
  // a third less data per k-step through its L2 port, which is what bounds the single-CTA kernel
  const bool pair = h->opt_pair && ((qpad / 128) % 2 == 0) && main_grid >= 2;
  if (pair) main_grid &= ~1;
  const int units = pair ? main_grid / 2 : main_grid;  // candidate slices are per (query, unit, column half)
  const int64_t num_n_tiles = (h->n_slots + 255) / 256;
  const int rank = std::max(1, std::min(kSeedR, h->opt_seed_rank));
  // sample size S (rows) so that the expected number of rows below the rank-th sample score is `hits`; the
  // sample is capped so that the per-(query, tile) seed lists stay below 256 MiB, and the expectation that results
  // from the cap (hits_eff >= hits) is what sizes the candidate slices and picks the select path below.  (The cap was
  // 64 MiB in round 1: a 12.5M-row shard with 4096 queries then got a third of the sample it asked for, three times
  // the survivors -- 1500 per query -- and with them three times the entries into the main pass's survivor section.)
  int64_t seed_tiles = (int64_t)std::ceil((double)rank * (double)h->n_slots / (double)hits / 256.0);
  const int64_t seed_tiles_max = std::max<int64_t>(64, std::min<int64_t>(4096, (256ll << 20) / ((int64_t)q * 2 * kSeedR * 4)));
  seed_tiles = std::max<int64_t>(1, std::min<int64_t>(seed_tiles, std::min<int64_t>(num_n_tiles, seed_tiles_max)));
  {
    // whole waves: trim the sample (by at most 20%) so the seed pass does not end in a mostly idle round
    const int64_t mt = qpad / 128, items = seed_tiles * mt, waves = (items + grid_sm - 1) / grid_sm;
    const int64_t trimmed = (waves - 1) * grid_sm / mt;
    if (waves > 1 && items % grid_sm != 0 && trimmed * 5 >= seed_tiles * 4) seed_tiles = trimmed;
  }
  const int64_t seed_stride = std::max<int64_t>(1, num_n_tiles / seed_tiles);
  const int hits_eff = (int)std::min<double>(1 << 20, std::max<double>(hits, (double)rank * (double)h->n_slots / ((double)seed_tiles * 256.0)));
  // candidate slots per slice: 4x the expected hits of a slice (+ slack), power of two
  const uint32_t cap = (uint32_t)std::min(4096, pow2_at_least(std::max(16, 4 * hits_eff / std::max(2 * units, 1) + 8)));
  const int64_t cand_stride = (int64_t)units * 2 * cap;  // per query
  const int num_m_tiles = qpad / 128;

  CU_TRY(c->cand.ensure((size_t)q * cand_stride * 8));
  CU_TRY(c->cand_cnt.ensure((size_t)q * 4));
  CU_TRY(c->cand_fb.ensure((size_t)q * scan_stride * 8));
  CU_TRY(c->cand_fb_cnt.ensure((size_t)q * 4));
  CU_TRY(c->thresh.ensure((size_t)q * 4));
  CU_TRY(c->seeds.ensure((size_t)q * seed_tiles * 2 * kSeedR * 4));
  CU_TRY(c->slice_cnt.ensure((size_t)units * 2 * q * 2));  // written in full by the main pass: no pre-fill
  // Usual case (expected candidates per query well below select_kernel's staging area): the select pass reads
  // only the valid prefix of every slice.  Large k: pre-fill with sentinels and let it scan the whole block.
  // (the staging area grows with the expected candidate count up to 16384 keys = 128 KB of shared memory: k = 100
  // means KP = 512 and ~4096 expected candidates, which the sentinel path answered with 8 radix passes over
  // 150 KB of global memory per query -- 156 us on C3a)
  const int sel_cap = std::max(kSelectStageKeys, std::min(16384, pow2_at_least(3 * hits_eff)));
  const bool slice_gather = 3 * hits_eff <= sel_cap;
  if (!slice_gather) CU_TRY(cudaMemsetAsync(c->cand.p, 0xff, (size_t)q * cand_stride * 8, st));

  CUtensorMap tmx, tmq, tmx_half;
  if (pair && !make_tmap_2d(&tmx_half, iv.x16, (uint64_t)h->dpad16, (uint64_t)h->n_slots, (uint64_t)h->dpad16 * 2, 64, 128))
    return fail(GFI_ERR_INDEX, "cuTensorMapEncodeTiled failed");
  if (!make_tmap_2d(&tmx, iv.x16, (uint64_t)h->dpad16, (uint64_t)h->n_slots, (uint64_t)h->dpad16 * 2, 64, 256) ||
      !make_tmap_2d(&tmq, c->q16.p, (uint64_t)h->dpad16, (uint64_t)qpad, (uint64_t)h->dpad16 * 2, 64, 128))
    return fail(GFI_ERR_INDEX, "cuTensorMapEncodeTiled failed");

  GemmParams gp{};
  gp.iv = iv;
  gp.mask = mv;
  gp.qmaxabs = reinterpret_cast<const float*>(&ctrl->qmaxabs_bits);
  gp.qsumsq = c->qsumsq.as<float>();
  gp.q = q;
  gp.num_m_tiles = num_m_tiles;
  gp.num_n_tiles = num_n_tiles;
  gp.seeds = c->seeds.as<float>();
  gp.thresh = c->thresh.as<float>();
  gp.cand = c->cand.as<uint64_t>();
  gp.cand_cnt = c->cand_cnt.as<uint32_t>();
  gp.cand_stride = cand_stride;
  gp.cand_cap = cap;
  gp.slice_cnt = c->slice_cnt.as<unsigned short>();
  gp.flags = &ctrl->flags;
  gp.seed_tiles = seed_tiles;
  gp.seed_stride = seed_stride;
  gp.skip = device_route ? &ctrl->skip_tensor : nullptr;
  gp.debug = h->opt_gemm_debug;
  // seed pass (short rows: the row-tile-stationary instance, a CTA per sample tile)
  gp.seed_mode = 1;
  gp.short_k = (short_k && h->opt_short_k_seed) ? 1 : 0;
  CU_TRY(launch_gemm_topk(gp, &tmx, &tmq,
                          (int)std::min<int64_t>(grid_sm, gp.short_k ? seed_tiles : seed_tiles * num_m_tiles), st));
  SeedFinalizeParams sf{c->seeds.as<float>(), q, seed_tiles, rank, c->thresh.as<float>(), gp.skip};
  CU_TRY(launch_seed_finalize(sf, st));
  // main pass
  // cosine without a caller mask: raw-accumulator epilogue (rows are stored pre-normalised, one coefficient)
  gp.seed_mode = (h->metric == kMetricCos && mv.bits == nullptr && h->opt_raw_epilogue) ? 3 : 0;
  gp.pair = pair ? 1 : 0;
  gp.short_k = short_k ? 1 : 0;
  prof_begin(h, c, 1, st);
  CU_TRY(launch_gemm_topk(gp, pair ? &tmx_half : &tmx, &tmq, main_grid, st));
  prof_end(h, c, st);
  // select + exact rerank + certification
  SelectParams s{};
  fill_select(s, c->cand, c->cand_cnt, cand_stride, KP);
  s.slice_cnt = c->slice_cnt.as<unsigned short>();
  s.nslices = units * 2;
  s.slice_cap = cap;
  s.slice_q = q;
  s.slice_gather = slice_gather ? 1 : 0;
  s.sel_cap = sel_cap;
  s.qlist = nullptr;
  s.nq_dev = device_route ? &ctrl->tensor_nq : nullptr;  // 0 when the device-side route chose the scan
  s.nq = device_route ? 0 : q;
  s.certify = 1;
  s.rerank_cut = h->opt_rerank_cut;
  s.thresh = c->thresh.as<float>();
  const float dd = (float)h->dim;
  s.eps_rel = 9.765625e-4f + 9.5367432e-07f + 2.f * std::sqrt(dd) * 1.4901161e-08f + (dd + 8.f) * 1.1920929e-07f;
  CU_TRY(launch_select_rerank(s, std::min(q, grid_sm * 8), st));
  // predicated exact fallback for uncertified queries (device-side count; no host sync)
  ScanParams sp{};
  fill_scan(sp, c->cand_fb, c->cand_fb_cnt);
  sp.qlist = c->fb_list.as<uint32_t>();
  sp.nq_dev = &ctrl->fb_count;
  sp.nq = 0;
  if (device_route) prof_begin(h, c, 0, st);  // this launch is the search itself when the device picks the scan
  CU_TRY(launch_scan(sp, QT, scan_grid, st));
  if (device_route) prof_end(h, c, st);
  SelectParams s2{};
  fill_select(s2, c->cand_fb, c->cand_fb_cnt, scan_stride, K);
  s2.qlist = c->fb_list.as<uint32_t>();
  s2.nq_dev = &ctrl->fb_count;
  s2.nq = 0;
  s2.certify = h->opt_scan_certify ? 2 : 0;
  s2.list_len = K;
  CU_TRY(launch_select_rerank(s2, std::min(q, grid_sm * 2), st));
  if (!last_chunk) CU_TRY(cudaMemsetAsync(&ctrl->fb_count, 0, 4, st));  // the next chunk starts an empty fallback list
  h->n_launch += 8;
  h->n_tensor_q += q;
  return GFI_OK;
}

// Enqueues one search batch; batches beyond the tensor kernel's per-launch query limit are cut into chunks.
int32_t enqueue_search(gfi_index* h, SearchCtx* c, const SearchArgs& a, cudaStream_t st) {
  const int64_t lim = kGemmMaxQueries;
  if (a.q <= lim) return enqueue_chunk(h, c, a, st, true, true);
  for (int64_t q0 = 0; q0 < a.q; q0 += lim) {
    SearchArgs s = a;
    s.q_base = q0;
    s.q_total = a.q;
    s.q = std::min(lim, a.q - q0);
    s.d_queries = a.d_queries + q0 * h->dim;
    s.d_ks = a.d_ks + q0;
    if (a.h_ks_in) s.h_ks_in = a.h_ks_in + q0;
    s.d_out_ids = a.d_out_ids + q0 * a.kstride;
    s.d_out_dist = a.d_out_dist + q0 * a.kstride;
    s.d_out_counts = a.d_out_counts + q0;
    int32_t rc = enqueue_chunk(h, c, s, st, q0 == 0, q0 + lim >= a.q);
    if (rc != GFI_OK) return rc;
  }
  return GFI_OK;
}

// Host-driven proof of ONE query whose scan-path answer the device could not certify (exact.cuh, scan_lower_bound):
// more rows than the kernel's K-entry lists hold lie within the summation-order error band of the k-th distance
// (clusters of near-duplicate rows).  The query is answered in PAGES of up to 1024 candidates in ascending order of the
// approximate key -- page j scans the rows whose key lies above page j-1's last key (ScanParams::floor64), re-scores
// its rows with the reference's arithmetic and returns its own exact top-k -- until the bound of a page's last
// key clears the k-th exact distance seen so far (or the rows run out).  The union of the pages' top-k lists then
// contains the reference's top-k, whatever the size of the cluster: FlatIndex::search's answer (flat_index.rs:52-65).
// The rare path: one scan per page, synchronous.  `res` receives (distance, id) ascending.
int32_t prove_query(gfi_index* h, SearchCtx* c, cudaStream_t st, const float* query_host, uint32_t k,
                    const uint64_t* d_mask, int64_t mask_bits, bool mask_by_slot,
                    std::vector<std::pair<float, uint64_t>>* res) {
  res->clear();
  if (k == 0) return GFI_OK;
  if (k > 1024u) return fail(GFI_ERR_INDEX, "k too large: the scan path supports k <= 1024");
  IndexView iv = h->view();
  if (mask_by_slot) iv.ids_identity = 1;
  MaskView mv{d_mask, mask_bits};
  // page size: the longest list the scan kernel's shared memory takes next to a ring of this row length
  // (1024 keys for short rows, 512 at d = 768), never below the search's own list length
  ScanGeom geo;
  int KPG = 1024;
  int32_t rc;
  while ((rc = scan_geometry(h, KPG, 1, &geo)) != GFI_OK && KPG / 2 >= std::max(32, pow2_at_least((int)k))) KPG /= 2;
  if (rc != GFI_OK) return rc;
  const int grid_sm = h->opt_grid > 0 ? h->opt_grid : h->sm_count;
  const int64_t nblocks = (h->n_slots + geo.R - 1) / geo.R;
  const int scan_grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid_sm, nblocks));
  const int64_t stride = (int64_t)scan_grid * KPG;
  CU_TRY(c->q_in.ensure((size_t)h->dim * 4));
  CU_TRY(c->q32.ensure((size_t)h->dpad * 4));
  CU_TRY(c->qnorm.ensure(4));
  CU_TRY(c->qsumsq.ensure(4));
  CU_TRY(c->ks.ensure(4));
  CU_TRY(c->ctrl.ensure(sizeof(Ctrl)));
  CU_TRY(c->floor.ensure(8));
  CU_TRY(c->cand_fb.ensure((size_t)stride * 8));
  CU_TRY(c->cand_fb_cnt.ensure(4));
  CU_TRY(c->sel_keys.ensure((size_t)KPG * 8));
  CU_TRY(c->sel_info.ensure(sizeof(SelInfo)));
  CU_TRY(c->out_ids.ensure((size_t)KPG * 8));
  CU_TRY(c->out_dist.ensure((size_t)KPG * 4));
  CU_TRY(c->out_counts.ensure(4));
  Ctrl* ctrl = c->ctrl.as<Ctrl>();
  CU_TRY(cudaMemsetAsync(ctrl, 0, sizeof(Ctrl), st));
  CU_TRY(cudaMemcpyAsync(c->q_in.p, query_host, (size_t)h->dim * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(c->ks.p, &k, 4, cudaMemcpyHostToDevice, st));
  PrepQueriesParams pq{};
  pq.q_in = c->q_in.as<float>();
  pq.q32 = c->q32.as<float>();
  pq.qnorm = c->qnorm.as<float>();
  pq.qsumsq = c->qsumsq.as<float>();
  pq.qmaxabs = reinterpret_cast<float*>(&ctrl->qmaxabs_bits);
  pq.q = 1;
  pq.qpad = 128;
  pq.d = (int)h->dim;
  pq.dpad = h->dpad;
  pq.dpad16 = h->dpad16;
  CU_TRY(launch_prep_queries(pq, st));
  const bool gather = d_mask != nullptr || h->n_live * 10 < h->n_slots * 9;
  if (gather) {
    CU_TRY(c->gather.ensure(((size_t)h->n_slots + 8) * 4));
    CU_TRY(cudaMemsetAsync(c->gather.p, 0, 16, st));
    CU_TRY(launch_compact_eligible(iv, mv, c->gather.as<uint32_t>() + 4, c->gather.as<uint32_t>(), st));
  }
  float qn = 0.f;
  CU_TRY(cudaMemcpyAsync(&qn, c->qnorm.p, 4, cudaMemcpyDeviceToHost, st));
  h->n_launch += gather ? 2 : 1;
  std::vector<uint64_t> pids(k);
  std::vector<float> pdist(k);
  uint64_t floor_key = 0;
  const int64_t max_pages = h->n_slots / KPG + 2;
  for (int64_t page = 0; page < max_pages; ++page) {
    if (page > 0) CU_TRY(cudaMemcpyAsync(c->floor.p, &floor_key, 8, cudaMemcpyHostToDevice, st));
    ScanParams sp{};
    sp.iv = iv;
    sp.mask = mv;
    sp.q32 = c->q32.as<float>();
    sp.qnorm = c->qnorm.as<float>();
    sp.nq = 1;
    sp.K = KPG;
    sp.floor64 = page > 0 ? c->floor.as<uint64_t>() : nullptr;
    sp.cand = c->cand_fb.as<uint64_t>();
    sp.cand_cnt = c->cand_fb_cnt.as<uint32_t>();
    sp.cand_stride = stride;
    sp.flags = &ctrl->flags;
    sp.gather_list = gather ? c->gather.as<uint32_t>() + 4 : nullptr;
    sp.gather_count = gather ? c->gather.as<uint32_t>() : nullptr;
    sp.rows_per_stage = geo.R;
    sp.seg_floats = geo.segf;
    sp.nseg = geo.nseg;
    sp.lanes_per_row = geo.lpr;
    sp.nstages = geo.nstages;
    sp.stage_floats = geo.stage_floats;
    CU_TRY(launch_scan(sp, geo.QT, scan_grid, st));
    SelectParams sl{};
    sl.iv = iv;
    sl.sel_keys = c->sel_keys.as<uint64_t>();
    sl.sel_info = c->sel_info.as<SelInfo>();
    sl.nq_max = 1;
    sl.nq = 1;
    sl.q32 = c->q32.as<float>();
    sl.qnorm = c->qnorm.as<float>();
    sl.qsumsq = c->qsumsq.as<float>();
    sl.ks = c->ks.as<uint32_t>();
    sl.cand = c->cand_fb.as<uint64_t>();
    sl.cand_cnt = c->cand_fb_cnt.as<uint32_t>();
    sl.cand_stride = stride;
    sl.KP = KPG;
    sl.list_len = KPG;
    sl.xnorm_max = h->xnorm_max;
    sl.out_ids = c->out_ids.as<uint64_t>();
    sl.out_dist = c->out_dist.as<float>();
    sl.out_counts = c->out_counts.as<uint32_t>();
    sl.kstride = KPG;
    sl.flags = &ctrl->flags;
    CU_TRY(launch_select_rerank(sl, 1, st));
    h->n_launch += 3;
    SelInfo info{};
    uint32_t cnt = 0;
    CU_TRY(cudaMemcpyAsync(&info, c->sel_info.p, sizeof(SelInfo), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&cnt, c->out_counts.p, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pids.data(), c->out_ids.p, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pdist.data(), c->out_dist.p, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (uint32_t j = 0; j < std::min(cnt, k); ++j) res->emplace_back(pdist[j] + 0.0f, pids[j]);
    std::sort(res->begin(), res->end());  // (distance, id): the reference's order with "lower id wins"
    if (res->size() > k) res->resize(k);
    if (info.kpeff < (uint32_t)KPG) return GFI_OK;  // the rows have run out
    uint32_t hi = (uint32_t)(info.pivot >> 32);
    hi = (hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi;  // key_f32 (common.cuh)
    float a_s;
    memcpy(&a_s, &hi, 4);
    if (res->size() >= k && a_s == a_s &&
        scan_lower_bound(h->metric, a_s, qn, h->xnorm_max, (int)h->dim) > (*res)[k - 1].first)
      return GFI_OK;
    floor_key = info.pivot;
  }
  return fail(GFI_ERR_INDEX, "internal: paging did not terminate");
}

int32_t flags_to_status(uint32_t fl) {
  if (fl & kFlagZeroNorm)
    return fail(GFI_ERR_INVALID_VECTOR, "Cannot compute cosine distance with zero vector");
  if (fl & kFlagNaN) return fail(GFI_ERR_NAN, "a distance is NaN (the reference panics in sort_by)");
  if (fl & kFlagInternal) return fail(GFI_ERR_INDEX, "internal kernel error");
  return GFI_OK;
}

// validation shared by host and device searches; *empty is set when the answer is trivially empty
int32_t precheck(gfi_index* h, int64_t qdim, bool* empty) {
  *empty = false;
  if (h->n_live + (int64_t)h->odd_dim_rows.size() == 0) {  // FlatIndex::search on an empty map: Ok(vec![])
    *empty = true;
    return GFI_OK;
  }
  if (!h->odd_dim_rows.empty()) {
    tl_expected = qdim;
    tl_actual = h->odd_dim_rows.begin()->second;
    for (auto& kv : h->odd_dim_rows)
      if (kv.second != qdim) { tl_actual = kv.second; break; }
    if (tl_actual != qdim || (h->n_live > 0 && h->dim != qdim)) {
      if (tl_actual == qdim) tl_actual = h->dim;
      return fail(GFI_ERR_DIMENSION_MISMATCH, "Dimension mismatch");
    }
    // every stored row has the query's dimension only if the index itself is empty: fall through
  }
  if (h->n_live > 0 && qdim != h->dim) {
    tl_expected = qdim;
    tl_actual = h->dim;
    return fail(GFI_ERR_DIMENSION_MISMATCH, "Dimension mismatch");
  }
  if (h->n_live == 0) {
    // Rows of a second dimension were added while rows of the latched dimension were live (they are recorded, not
    // stored), and the latter have all been removed since: FlatIndex::search would now return the former.  That
    // state cannot be served from the GPU copy -- say so instead of answering with an empty Ok.
    if (!h->odd_dim_rows.empty())
      return fail(GFI_ERR_INDEX, "rows whose dimension differs from the index dimension are not stored on the GPU; "
                                 "re-add them now that the index holds no other rows");
    *empty = true;
  }
  return GFI_OK;
}

}  // namespace

// ======================================= C ABI =======================================
extern "C" {

int32_t gfi_version(void) { return 101; }

const char* gfi_last_error(void) { return tl_error.c_str(); }

void gfi_last_mismatch(int64_t* expected, int64_t* actual) {
  if (expected) *expected = tl_expected;
  if (actual) *actual = tl_actual;
}

int32_t gfi_create(gfi_index** out, int32_t metric, int64_t dim, int32_t device, uint32_t flags) {
  if (!out) return fail(GFI_ERR_INDEX, "null out pointer");
  *out = nullptr;
  if (metric < 0 || metric > 2) return fail(GFI_ERR_INDEX, "unknown metric");
  if (dim < 0) return fail(GFI_ERR_INDEX, "negative dimension");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(GFI_ERR_INDEX, "no CUDA device: libgfi has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(GFI_ERR_INDEX, "bad device ordinal");
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(GFI_ERR_INDEX, "libgfi is built for sm_100a (B200) only");
  std::unique_ptr<gfi_index> h(new gfi_index());
  h->metric = metric;
  h->device = device;
  h->flags = flags;
  h->sm_count = prop.multiProcessorCount;
  h->use_x16 = !(flags & GFI_FLAG_NO_TENSOR);
  if (dim > 0) latch_dim(h.get(), dim);
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaStreamCreateWithFlags(&h->ingest_stream, cudaStreamNonBlocking));
  CU_TRY(h->counters.ensure(16));
  CU_TRY(cudaMemset(h->counters.p, 0, 16));
  *out = h.release();
  // experiment aid for native probes that cannot call gfi_set_option: GFI_OPTS="name=value,name=value"
  if (const char* env = getenv("GFI_OPTS")) {
    std::string e(env);
    size_t at = 0;
    while (at < e.size()) {
      size_t end = e.find(',', at);
      if (end == std::string::npos) end = e.size();
      const std::string kv = e.substr(at, end - at);
      const size_t eq = kv.find('=');
      if (eq != std::string::npos) gfi_set_option(*out, kv.substr(0, eq).c_str(), atoll(kv.c_str() + eq + 1));
      at = end + 1;
    }
  }
  return GFI_OK;
}

int32_t gfi_create_sharded(gfi_index** out, int32_t metric, int64_t dim, const int32_t* devices, int32_t n_devices,
                           uint32_t flags) {
  return sharded_create(out, metric, dim, devices, n_devices, flags);
}

int32_t gfi_destroy(gfi_index* h) {
  if (!h) return GFI_OK;
  if (h->shards) { sharded_destroy(h); return GFI_OK; }
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& c : h->pool) c->release();
  h->pool.clear();
  for (DevBuf* b : {&h->x32, &h->x16, &h->ids, &h->norm, &h->sumsq, &h->coef, &h->live, &h->rowflags, &h->counters})
    b->release();
  for (auto& b : h->meta_dcols) b.release();
  h->meta_dptrs.release();
  h->st_rows.release();
  h->st_ids.release();
  if (h->ingest_stream) cudaStreamDestroy(h->ingest_stream);
  delete h;
  return GFI_OK;
}

int64_t gfi_len(const gfi_index* h) {
  if (!h) return 0;
  if (h->shards) return sharded_len(h);
  std::shared_lock<std::shared_mutex> g(h->mu);
  // staged rows that overwrite an id already stored replace it at flush (HashMap::insert): count them once
  int64_t staged_new = 0;
  const uint64_t* sid = h->st_ids.as<uint64_t>();
  for (int64_t i = 0; i < h->st_n; ++i) {
    uint32_t slot;
    if (!(h->any_id && sid[i] <= h->max_id_seen && lookup_slot(h, sid[i], &slot)) && !h->odd_dim_rows.count(sid[i]))
      ++staged_new;
  }
  return h->n_live + staged_new + (int64_t)h->odd_dim_rows.size();
}
int32_t gfi_metric(const gfi_index* h) { return h ? h->metric : -1; }
int64_t gfi_dim(const gfi_index* h) { return !h ? 0 : h->shards ? sharded_dim(h) : h->dim; }

int32_t gfi_reserve(gfi_index* h, int64_t n_rows) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (h->shards) return sharded_reserve(h, n_rows);
  std::unique_lock<std::shared_mutex> g(h->mu);
  if (h->dim == 0) return fail(GFI_ERR_INDEX, "reserve before the dimension is known");
  int32_t rc;
  if ((rc = set_device(h)) != GFI_OK) return rc;
  return grow(h, n_rows);
}

static int32_t add_rows_locked(gfi_index* h, const uint64_t* ids, const float* rows, int64_t n, int64_t dim) {
  if (h->dim == 0 && n > 0) {
    if (dim <= 0) return fail(GFI_ERR_INDEX, "zero-dimensional vectors are not supported");
    latch_dim(h, dim);
  }
  bool odd_covered = true;  // every recorded odd-dimension row is re-added by this very call
  if (dim != h->dim && !h->odd_dim_rows.empty()) {
    std::unordered_map<uint64_t, bool> incoming;
    for (int64_t i = 0; i < n; ++i) incoming[ids[i]] = true;
    for (auto& kv : h->odd_dim_rows) odd_covered = odd_covered && incoming.count(kv.first);
  }
  if (dim != h->dim && n > 0 && dim > 0 && h->n_live == 0 && odd_covered) {
    // nothing is stored any more (every row was removed): the index takes the new dimension, as an empty FlatIndex
    // would (it has no dimension of its own)
    int32_t rc = flush_locked(h);
    if (rc != GFI_OK) return rc;
    if (h->n_live == 0) {
      if ((rc = set_device(h)) != GFI_OK) return rc;
      CU_TRY(cudaDeviceSynchronize());  // no search may still be reading the old arrays
      for (DevBuf* b : {&h->x32, &h->x16, &h->ids, &h->norm, &h->sumsq, &h->coef, &h->live, &h->rowflags}) b->release();
      h->cap = h->n_slots = 0;
      h->runs.clear();
      h->h_live.clear();
      h->live_dirty_lo = h->live_dirty_hi = -1;
      h->ids_identity = true;
      h->needs_reorder = false;
      h->any_id = false;
      h->max_id_seen = 0;
      h->st_rows.release();
      h->st_ids.release();
      h->st_cap = 0;
      for (auto& col : h->meta_cols) col.clear();
      h->meta_synced_slots = 0;
      if (!h->meta_values.empty()) h->meta_dirty = true;
      ++h->layout_gen;
      CU_TRY(cudaMemset(h->counters.p, 0, 16));
      h->zero_rows_ever = h->unsafe_rows_ever = 0;
      h->xnorm_max = 0.f;
      latch_dim(h, dim);
    }
  }
  if (dim != h->dim) {
    // FlatIndex::add never checks dimensions (flat_index.rs:38-41); remember the row so that searches
    // fail with DimensionMismatch exactly as the per-pair check would (distance.rs:21-26).
    int32_t rc = flush_locked(h);
    if (rc != GFI_OK) return rc;
    for (int64_t i = 0; i < n; ++i) {
      erase_id(h, ids[i]);
      h->odd_dim_rows[ids[i]] = dim;
    }
    return GFI_OK;
  }
  int32_t rc;
  if ((rc = set_device(h)) != GFI_OK) return rc;
  const int64_t chunk_rows = std::max<int64_t>(1, (64ll << 20) / ((int64_t)h->dpad * 4));
  int64_t done = 0;
  while (done < n) {
    if (h->st_cap == 0) {
      CU_TRY(h->st_rows.ensure((size_t)chunk_rows * h->dpad * 4));
      CU_TRY(h->st_ids.ensure((size_t)chunk_rows * 8));
      h->st_cap = chunk_rows;
    }
    const int64_t take = std::min(n - done, h->st_cap - h->st_n);
    float* dst = h->st_rows.as<float>() + (size_t)h->st_n * h->dpad;
    if (h->dpad == h->dim) {
      memcpy(dst, rows + (size_t)done * dim, (size_t)take * dim * 4);
    } else {
      for (int64_t i = 0; i < take; ++i) {
        memcpy(dst + (size_t)i * h->dpad, rows + (size_t)(done + i) * dim, (size_t)dim * 4);
        memset(dst + (size_t)i * h->dpad + dim, 0, (size_t)(h->dpad - dim) * 4);
      }
    }
    memcpy(h->st_ids.as<uint64_t>() + h->st_n, ids + done, (size_t)take * 8);
    // an id staged twice in one batch must keep only the last copy: flush between duplicates is
    // handled by erase_id at flush time only for already-flushed ids, so flush eagerly when the
    // staging buffer is full or ids are not strictly increasing within the staged batch.
    bool increasing = true;
    const uint64_t* sid = h->st_ids.as<uint64_t>();
    for (int64_t i = std::max<int64_t>(1, h->st_n); i < h->st_n + take; ++i)
      if (sid[i] <= sid[i - 1]) { increasing = false; break; }
    h->st_n += take;
    done += take;
    if (!increasing) {
      // flush row by row segments so duplicates inside the staged batch resolve in order
      const int64_t total = h->st_n;
      std::vector<uint64_t> ids_copy(sid, sid + total);
      std::vector<float> rows_copy(h->st_rows.as<float>(), h->st_rows.as<float>() + (size_t)total * h->dpad);
      h->st_n = 0;
      int64_t s0 = 0;
      while (s0 < total) {
        int64_t s1 = s0 + 1;
        while (s1 < total && ids_copy[s1] > ids_copy[s1 - 1]) ++s1;
        memcpy(h->st_rows.p, rows_copy.data() + (size_t)s0 * h->dpad, (size_t)(s1 - s0) * h->dpad * 4);
        memcpy(h->st_ids.p, ids_copy.data() + s0, (size_t)(s1 - s0) * 8);
        h->st_n = s1 - s0;
        if ((rc = flush_locked(h)) != GFI_OK) return rc;
        s0 = s1;
      }
    } else if (h->st_n == h->st_cap) {
      if ((rc = flush_locked(h)) != GFI_OK) return rc;
    }
  }
  return GFI_OK;
}

int32_t gfi_add(gfi_index* h, const uint64_t* ids, const float* rows, int64_t n, int64_t dim) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (n < 0 || (n > 0 && (!ids || (!rows && dim > 0)))) return fail(GFI_ERR_INDEX, "bad arguments");
  if (h->shards) return sharded_add(h, ids, rows, n, dim);
  std::unique_lock<std::shared_mutex> g(h->mu);
  return add_rows_locked(h, ids, rows, n, dim);
}

int32_t gfi_add_from_file(gfi_index* h, const char* path, uint64_t first_id, int64_t* out_rows) {
  if (!h || !path) return fail(GFI_ERR_INDEX, "null argument");
  if (out_rows) *out_rows = 0;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(GFI_ERR_INDEX, std::string("cannot open ") + path);
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
  unsigned char hdr[8];
  if (fread(hdr, 1, 8, f) != 8) return fail(GFI_ERR_INDEX, "File too small for header");
  const uint32_t dim = (uint32_t)hdr[0] | ((uint32_t)hdr[1] << 8) | ((uint32_t)hdr[2] << 16) | ((uint32_t)hdr[3] << 24);
  const uint32_t count = (uint32_t)hdr[4] | ((uint32_t)hdr[5] << 8) | ((uint32_t)hdr[6] << 16) | ((uint32_t)hdr[7] << 24);
  if (dim == 0) return fail(GFI_ERR_INDEX, "flat file has dimension 0");
  std::unique_lock<std::shared_mutex> g(h->mu, std::defer_lock);
  if (!h->shards) g.lock();  // (a sharded index locks per chunk inside sharded_add)
  const int64_t have_dim = h->shards ? sharded_dim(h) : h->dim;
  if (have_dim != 0 && (int64_t)dim != have_dim) {
    tl_expected = have_dim;
    tl_actual = dim;
    return fail(GFI_ERR_DIMENSION_MISMATCH, "Dimension mismatch");
  }
  if (!h->shards && count > 0 && !(h->any_id && first_id <= h->max_id_seen)) {
    // Bulk path (fresh, ascending ids -- what a load of the reference's flat file is): the file is read straight
    // into two pinned buffers in turn, and each chunk's H2D copy, id fill and row_stats pass run on the ingest stream
    // while the next chunk is being read; one synchronisation at the end.
    int32_t rc;
    {
      // a file shorter than its header says must add nothing: checked before the first chunk reaches the GPU (the
      // per-chunk row statistics feed index-wide counters that cannot be taken back)
      struct stat sb;
      if (fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode) &&
          (uint64_t)sb.st_size < 8ull + (uint64_t)count * (uint64_t)dim * 4ull)
        return fail(GFI_ERR_INDEX, "flat file is shorter than its header says");
    }
    if (h->dim == 0) latch_dim(h, dim);
    if ((rc = flush_locked(h)) != GFI_OK) return rc;
    if ((rc = grow(h, h->n_slots + (int64_t)count)) != GFI_OK) return rc;
    const uint32_t slot0 = (uint32_t)h->n_slots;
    const int64_t rows_per = std::max<int64_t>(1, (64ll << 20) / ((int64_t)dim * 4));
    PinBuf bufs[2];
    cudaEvent_t ev[2] = {nullptr, nullptr};
    struct Cleanup {
      PinBuf* b; cudaEvent_t* e; cudaStream_t st;
      ~Cleanup() { cudaStreamSynchronize(st); for (int i = 0; i < 2; ++i) { b[i].release(); if (e[i]) cudaEventDestroy(e[i]); } }
    } cleanup{bufs, ev, h->ingest_stream};
    for (int i = 0; i < 2; ++i) {
      CU_TRY(bufs[i].ensure((size_t)rows_per * dim * 4));
      CU_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    const int fd = fileno(f);
    const unsigned hw = std::thread::hardware_concurrency();
    const int n_readers = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 4u));
    int64_t done = 0;
    for (int64_t ci = 0; done < (int64_t)count; ++ci) {
      const int b = (int)(ci & 1);
      const int64_t take = std::min<int64_t>(rows_per, (int64_t)count - done);
      if (ci >= 2) CU_TRY(cudaEventSynchronize(ev[b]));  // the copy that read this buffer two chunks ago
      // the chunk is read by several threads at once (pread on disjoint slices): one thread copying out of the page
      // cache moves ~5 GB/s, a tenth of what the PCIe copy behind it can take
      {
        const size_t bytes = (size_t)take * dim * 4;
        const off_t base = (off_t)8 + (off_t)done * (off_t)dim * 4;
        const int nthr = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_readers, bytes >> 20));
        const size_t slice = ((bytes + (size_t)nthr - 1) / (size_t)nthr + 4095) & ~(size_t)4095;
        std::atomic<bool> short_read{false};
        auto read_slice = [&](int t) {
          size_t lo = (size_t)t * slice, hi = std::min(bytes, lo + slice);
          char* dst = static_cast<char*>(bufs[b].p);
          while (lo < hi) {
            const ssize_t got = pread(fd, dst + lo, hi - lo, base + (off_t)lo);
            if (got <= 0) { short_read = true; return; }
            lo += (size_t)got;
          }
        };
        std::vector<std::thread> readers;
        for (int t = 1; t < nthr; ++t) readers.emplace_back(read_slice, t);
        read_slice(0);
        for (auto& th : readers) th.join();
        if (short_read) return fail(GFI_ERR_INDEX, "flat file is shorter than its header says");
      }
      float* dst = h->x32.as<float>() + (size_t)(slot0 + done) * h->dpad;
      if (h->dpad == (int)dim) {
        CU_TRY(cudaMemcpyAsync(dst, bufs[b].p, (size_t)take * dim * 4, cudaMemcpyHostToDevice, h->ingest_stream));
      } else {  // rows are padded to a multiple of 4 floats on the device: zero the block, copy row by row pitch
        CU_TRY(cudaMemsetAsync(dst, 0, (size_t)take * h->dpad * 4, h->ingest_stream));
        CU_TRY(cudaMemcpy2DAsync(dst, (size_t)h->dpad * 4, bufs[b].p, (size_t)dim * 4, (size_t)dim * 4, (size_t)take,
                                 cudaMemcpyHostToDevice, h->ingest_stream));
      }
      CU_TRY(cudaEventRecord(ev[b], h->ingest_stream));
      CU_TRY(launch_fill_ids(h->ids.as<uint64_t>() + slot0 + done, first_id + (uint64_t)done, take, h->ingest_stream));
      if ((rc = ingest_slots(h, (int64_t)slot0 + done, take, false, 0, 0, 0)) != GFI_OK) return rc;
      done += take;
    }
    CU_TRY(cudaStreamSynchronize(h->ingest_stream));
    h->n_slots += (int64_t)count;
    if (!h->odd_dim_rows.empty())  // ids recorded earlier with another dimension are replaced, as FlatIndex::add would
      for (uint64_t id = first_id; id < first_id + (uint64_t)count; ++id) h->odd_dim_rows.erase(id);
    register_ids(h, nullptr, first_id, (int64_t)count, slot0);
    if (!h->meta_values.empty()) h->meta_dirty = true;
    if ((rc = read_counters(h)) != GFI_OK) return rc;
    if ((rc = upload_live(h)) != GFI_OK) return rc;
    if (out_rows) *out_rows = (int64_t)count;
    return GFI_OK;
  }
  const int64_t chunk = std::max<int64_t>(1, (32ll << 20) / ((int64_t)dim * 4));
  std::vector<float> buf((size_t)chunk * dim);
  std::vector<uint64_t> ids((size_t)chunk);
  int64_t done = 0;
  while (done < (int64_t)count) {
    const int64_t take = std::min<int64_t>(chunk, (int64_t)count - done);
    if (fread(buf.data(), 4, (size_t)take * dim, f) != (size_t)take * dim)
      return fail(GFI_ERR_INDEX, "flat file is shorter than its header says");
    for (int64_t i = 0; i < take; ++i) ids[(size_t)i] = first_id + (uint64_t)(done + i);
    int32_t rc = h->shards ? sharded_add(h, ids.data(), buf.data(), take, dim)
                           : add_rows_locked(h, ids.data(), buf.data(), take, dim);
    if (rc != GFI_OK) return rc;
    done += take;
    if (out_rows) *out_rows = done;
  }
  return GFI_OK;
}

int32_t gfi_add_generated(gfi_index* h, uint32_t seed, uint64_t first_row, int64_t n, int32_t kind,
                          uint64_t first_id) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (n <= 0) return GFI_OK;
  if (h->shards) return sharded_add_generated(h, seed, first_row, n, kind, first_id);
  std::unique_lock<std::shared_mutex> g(h->mu);
  if (h->dim == 0) return fail(GFI_ERR_INDEX, "gfi_add_generated needs an index created with a dimension");
  int32_t rc;
  if ((rc = flush_locked(h)) != GFI_OK) return rc;
  if ((rc = grow(h, h->n_slots + n)) != GFI_OK) return rc;
  const uint32_t slot0 = (uint32_t)h->n_slots;
  // ids: first_id + i, written on device by a tiny fill via host chunks
  {
    const int64_t chunk = 1 << 20;
    std::vector<uint64_t> tmp((size_t)std::min(chunk, n));
    for (int64_t o = 0; o < n; o += chunk) {
      const int64_t m = std::min(chunk, n - o);
      for (int64_t i = 0; i < m; ++i) tmp[(size_t)i] = first_id + (uint64_t)(o + i);
      CU_TRY(cudaMemcpyAsync(h->ids.as<uint64_t>() + slot0 + o, tmp.data(), (size_t)m * 8,
                             cudaMemcpyHostToDevice, h->ingest_stream));
      CU_TRY(cudaStreamSynchronize(h->ingest_stream));
    }
  }
  if (h->any_id && first_id <= h->max_id_seen)
    return fail(GFI_ERR_INDEX, "gfi_add_generated: ids must be above every id already stored");
  h->n_slots += n;
  register_ids(h, nullptr, first_id, n, slot0);
  if ((rc = ingest_slots(h, slot0, n, true, seed, first_row, kind)) != GFI_OK) return rc;
  CU_TRY(cudaStreamSynchronize(h->ingest_stream));
  if (!h->meta_values.empty()) h->meta_dirty = true;  // rows without metadata still need (absent) column entries
  if ((rc = read_counters(h)) != GFI_OK) return rc;
  return upload_live(h);
}

int32_t gfi_remove(gfi_index* h, uint64_t id) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (h->shards) return sharded_remove(h, id);
  std::unique_lock<std::shared_mutex> g(h->mu);
  if (h->odd_dim_rows.erase(id)) return GFI_OK;
  if (h->st_n > 0) {
    int32_t rc = flush_locked(h);
    if (rc != GFI_OK) return rc;
  }
  erase_id(h, id);  // idempotent: FlatIndex::remove of a missing id is Ok(())
  return GFI_OK;
}

int32_t gfi_flush(gfi_index* h) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (h->shards) return sharded_flush(h, false);
  std::unique_lock<std::shared_mutex> g(h->mu);
  int32_t rc = flush_locked(h);
  if (rc != GFI_OK) return rc;
  if (h->needs_reorder) return compact_locked(h);
  return GFI_OK;
}

int32_t gfi_compact(gfi_index* h) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (h->shards) return sharded_flush(h, true);
  std::unique_lock<std::shared_mutex> g(h->mu);
  int32_t rc = flush_locked(h);
  if (rc != GFI_OK) return rc;
  return compact_locked(h);
}

int32_t gfi_get_vector(gfi_index* h, uint64_t id, float* out, int64_t cap, int64_t* out_dim) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (h->shards) return sharded_get_vector(h, id, out, cap, out_dim);
  std::unique_lock<std::shared_mutex> g(h->mu);
  int32_t rc = flush_locked(h);
  if (rc != GFI_OK) return rc;
  uint32_t slot;
  if (!lookup_slot(h, id, &slot)) return fail(GFI_ERR_INDEX, "vector not found");
  if (out_dim) *out_dim = h->dim;
  if (cap < h->dim) return fail(GFI_ERR_INDEX, "output buffer too small");
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaMemcpy(out, h->x32.as<float>() + (size_t)slot * h->dpad, (size_t)h->dim * 4, cudaMemcpyDeviceToHost));
  return GFI_OK;
}

// applies pending metadata to the per-field columns and uploads dirty columns.  Unique lock held.
static int32_t sync_metadata_locked(gfi_index* h) {
  if (!h->meta_dirty) return GFI_OK;
  int32_t rc;
  if ((rc = set_device(h)) != GFI_OK) return rc;
  const size_t nf = h->meta_values.size();
  h->meta_cols.resize(nf);
  h->meta_dcols.resize(nf);
  for (auto& c : h->meta_cols) c.resize((size_t)h->n_slots, 0u);
  for (auto it = h->meta_pending.begin(); it != h->meta_pending.end();) {
    uint32_t slot;
    if (lookup_slot(h, it->first, &slot)) {
      for (auto& c : h->meta_cols) c[slot] = 0u;  // Metadata is replaced wholesale at insert (storage.rs:169)
      for (auto& fv : it->second) h->meta_cols[(size_t)fv.first][slot] = fv.second;
      it = h->meta_pending.erase(it);
    } else {
      ++it;  // row not stored (yet): keep for later
    }
  }
  std::vector<const uint32_t*> ptrs(nf);
  for (size_t f = 0; f < nf; ++f) {
    CU_TRY(h->meta_dcols[f].ensure(std::max<size_t>(h->meta_cols[f].size(), 1) * 4));
    if (!h->meta_cols[f].empty())
      CU_TRY(cudaMemcpyAsync(h->meta_dcols[f].p, h->meta_cols[f].data(), h->meta_cols[f].size() * 4,
                             cudaMemcpyHostToDevice, h->ingest_stream));
    ptrs[f] = h->meta_dcols[f].as<uint32_t>();
  }
  CU_TRY(h->meta_dptrs.ensure(std::max<size_t>(nf, 1) * sizeof(void*)));
  if (nf) CU_TRY(cudaMemcpyAsync(h->meta_dptrs.p, ptrs.data(), nf * sizeof(void*), cudaMemcpyHostToDevice, h->ingest_stream));
  CU_TRY(cudaStreamSynchronize(h->ingest_stream));
  h->meta_dirty = false;
  h->meta_synced_slots = h->n_slots;
  return GFI_OK;
}

static int32_t ensure_flushed(gfi_index* h, bool with_metadata = false) {
  bool need;
  {
    std::shared_lock<std::shared_mutex> g(h->mu);
    need = h->st_n > 0 || h->live_dirty_lo >= 0 || h->needs_reorder || (with_metadata && h->meta_dirty);
  }
  if (!need) return GFI_OK;
  std::unique_lock<std::shared_mutex> g(h->mu);
  int32_t rc = flush_locked(h);
  if (rc != GFI_OK) return rc;
  if (h->needs_reorder && (rc = compact_locked(h)) != GFI_OK) return rc;
  // (after a compaction, which moves rows and re-marks the columns dirty, never before)
  if (with_metadata) return sync_metadata_locked(h);
  return GFI_OK;
}

static int32_t search_impl(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                           const uint64_t* mask, int64_t mask_bits, const char* filter_json, uint64_t* out_ids,
                           float* out_dist, uint32_t* out_counts, int64_t kstride) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (q < 0 || (q > 0 && (!queries || !ks || !out_counts))) return fail(GFI_ERR_INDEX, "bad arguments");
  if (q == 0) return GFI_OK;
  if (h->shards)  // one index over several GPUs: fan out, exchange over NVLink, merge (sharded.cu)
    return sharded_search(h, queries, q, dim, ks, mask, mask_bits, filter_json, out_ids, out_dist, out_counts, kstride);
  int32_t rc;
  std::shared_lock<std::shared_mutex> g(h->mu, std::defer_lock);
  for (;;) {
    if ((rc = ensure_flushed(h, filter_json != nullptr)) != GFI_OK) return rc;
    g.lock();
    // a writer may have slipped in between the two locks; a filtered search needs columns that match the layout
    if (!(filter_json && h->meta_dirty)) break;
    g.unlock();
  }
  ++h->n_search;
  h->n_queries += q;
  bool empty;
  if ((rc = precheck(h, dim, &empty)) != GFI_OK) return rc;
  FilterProgram prog{};
  if (filter_json) {
    std::string err;
    if (!compile_filter(filter_json, h->meta_fields, h->meta_values, &prog, &err))
      return fail(GFI_ERR_INDEX, "bad filter: " + err);
    // a field created after the last metadata sync has no column yet: treat as absent
    for (int i = 0; i < prog.n; ++i)
      if (prog.ops[i].field >= (int)h->meta_dcols.size()) prog.ops[i].field = -1;
  }
  uint32_t kmax = 0;
  for (int64_t i = 0; i < q; ++i) kmax = std::max(kmax, ks[i]);
  if (empty || kmax == 0) {
    for (int64_t i = 0; i < q; ++i) out_counts[i] = 0;
    return GFI_OK;
  }
  if ((int64_t)kmax > kstride) return fail(GFI_ERR_INDEX, "kstride smaller than max k");
  if (!out_ids || !out_dist) return fail(GFI_ERR_INDEX, "null output buffers");
  if ((rc = set_device(h)) != GFI_OK) return rc;
  static const bool host_trace = getenv("GFI_HOST_TRACE") != nullptr;  // investigation aid: phase times on stderr
  auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = host_trace ? now_us() : 0.0;
  SearchCtx* c = acquire_ctx(h);
  if (!c) return fail(GFI_ERR_INDEX, "cannot create a CUDA stream");
  // (a context goes back to the pool only with an idle stream: every early error return below drains it first)
  struct Releaser {
    gfi_index* h; SearchCtx* c; bool in_flight;
    ~Releaser() { if (in_flight) cudaStreamSynchronize(c->stream); release_ctx(h, c); }
  } rel{h, c, false};
  cudaStream_t st = c->stream;
  const uint32_t kout = std::min<uint32_t>(kmax, (uint32_t)std::min<int64_t>(kstride, 1 << 20));
  if (filter_json) { mask = nullptr; mask_bits = h->n_slots; }
  const size_t mask_words = (mask || filter_json) ? (size_t)((mask_bits + 63) / 64) : 0;
  CU_TRY(c->q_in.ensure((size_t)q * dim * 4));
  CU_TRY(c->ks.ensure((size_t)q * 4));
  // one device block [Ctrl | counts | dist | ids] so a single D2H copy returns everything
  const size_t off_cnt = 64, off_dist = off_cnt + (((size_t)q * 4 + 63) & ~(size_t)63);
  const size_t off_ids = off_dist + (((size_t)q * kout * 4 + 63) & ~(size_t)63);
  const size_t blk_bytes = off_ids + (size_t)q * kout * 8;
  CU_TRY(c->out_blk.ensure(blk_bytes));
  CU_TRY(c->h_out.ensure(blk_bytes));
  CU_TRY(c->h_q.ensure((size_t)q * dim * 4));
  CU_TRY(c->h_ks.ensure((size_t)q * 4));
  char* blk = c->out_blk.as<char>();
  c->ctrl_dev = blk;
  if (mask || filter_json) CU_TRY(c->mask.ensure(mask_words * 8 + 8));
  // Latency mode (small plain searches): the kernels read queries and ks straight from this context's pinned
  // staging buffers over PCIe, and a fused-tail scan writes the results and a completion word back into pinned host
  // memory, so the call issues no copy-engine operation and no stream synchronisation (each costs a few
  // microseconds of engine hand-over, which is most of a 10k-row search).  Not with per-kernel event timing on.
  const bool zero_copy = h->opt_zero_copy && !h->opt_profile && !mask && !filter_json && (size_t)q * dim * 4 <= 16384;
  c->zc_pending = false;
  const float* d_queries = c->q_in.as<float>();
  memcpy(c->h_ks.p, ks, (size_t)q * 4);
  rel.in_flight = true;
  if (zero_copy) {
    memcpy(c->h_q.p, queries, (size_t)q * dim * 4);
    d_queries = c->h_q.as<float>();  // cudaMallocHost memory: device-accessible under unified addressing
  } else {
    // queries already in pinned memory are copied straight from the caller's buffer
    const void* q_src = queries;
    if (!is_pinned_host(queries)) {
      memcpy(c->h_q.p, queries, (size_t)q * dim * 4);
      q_src = c->h_q.p;
    }
    CU_TRY(cudaMemcpyAsync(c->q_in.p, q_src, (size_t)q * dim * 4, cudaMemcpyHostToDevice, st));
    // (the per-query k is read by prep_queries straight from the pinned staging block: one copy-engine op fewer)
  }
  if (mask) CU_TRY(cudaMemcpyAsync(c->mask.p, mask, mask_words * 8, cudaMemcpyHostToDevice, st));
  if (filter_json) {
    // the filter is evaluated on the device into a bitmask over slots: host metadata is never walked
    CU_TRY(cudaMemsetAsync(c->mask.p, 0, mask_words * 8 + 8, st));
    CU_TRY(launch_eval_filter(prog, h->meta_dptrs.as<const uint32_t*>(), h->meta_synced_slots, h->n_slots,
                                 c->mask.as<uint64_t>(), st));
    ++h->n_launch;
  }
  SearchArgs a{d_queries, q, c->ks.as<uint32_t>(), kmax,
               (mask || filter_json) ? c->mask.as<uint64_t>() : nullptr, mask_bits,
               reinterpret_cast<uint64_t*>(blk + off_ids), reinterpret_cast<float*>(blk + off_dist),
               reinterpret_cast<uint32_t*>(blk + off_cnt), (int64_t)kout};
  a.mask_by_slot = filter_json != nullptr || tl_mask_by_slot;
  if (mask && !tl_mask_by_slot && h->n_slots * (int64_t)h->dpad * 4 >= (1ll << 30)) {  // cost model input (large indexes)
    // an estimate is enough: every 64th word of the mask (one word of every 8th cache line; a full pass over 10M
    // bits costs more than the model saves)
    int64_t pc = 0, seen = 0;
    const size_t full = (size_t)(mask_bits / 64);
    for (size_t w = 0; w < full; w += 64, ++seen) pc += __builtin_popcountll(mask[w]);
    a.mask_popcount = seen ? (int64_t)((double)pc / (double)(seen * 64) * (double)mask_bits) : 0;
  }
  a.h_ks_in = c->h_ks.as<uint32_t>();
  if (zero_copy) {
    char* hb0 = c->h_out.as<char>();
    a.h_ctrl = reinterpret_cast<uint32_t*>(hb0);
    a.h_out_counts = reinterpret_cast<uint32_t*>(hb0 + off_cnt);
    a.h_out_dist = reinterpret_cast<float*>(hb0 + off_dist);
    a.h_out_ids = reinterpret_cast<uint64_t*>(hb0 + off_ids);
    a.done_seq = ++c->zc_seq ? c->zc_seq : ++c->zc_seq;  // never 0
    reinterpret_cast<volatile uint32_t*>(hb0)[kCtrlDoneWord] = 0;
  }
  rc = enqueue_search(h, c, a, st);
  c->ctrl_dev = nullptr;
  if (rc != GFI_OK) { cudaStreamSynchronize(st); return rc; }
  bool published = false;
  double t_enq = host_trace ? now_us() : 0.0;
  if (c->zc_pending) {
    // the scan's fused tail publishes the results itself: poll its completion word (bounded; a search that takes
    // longer than that is not latency-bound, and the ordinary copy below returns the same block)
    const volatile uint32_t* done = reinterpret_cast<const volatile uint32_t*>(c->h_out.p) + kCtrlDoneWord;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t it = 1;; ++it) {
      if (*done == a.done_seq) { published = true; break; }
      if ((it & 1023u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(2000)) break;
      if (it >= 4096 && (it & 255u) == 0) std::this_thread::yield();  // more clients than cores: let them enqueue
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  if (!published) {
    CU_TRY(cudaMemcpyAsync(c->h_out.p, blk, blk_bytes, cudaMemcpyDeviceToHost, st));
    if (host_trace) t_enq = now_us();
    CU_TRY(cudaStreamSynchronize(st));
  }
  rel.in_flight = false;
  const double t_sync = host_trace ? now_us() : 0.0;
  prof_collect(h, c);
  const char* hb = c->h_out.as<char>();
  const Ctrl* hc = reinterpret_cast<const Ctrl*>(hb);
  h->n_fallback_q += hc->uncertified;
  h->n_tensor_q -= hc->routed_scan;  // device-side route: counted as tensor queries at enqueue
  h->n_scan_q += hc->routed_scan;
  if (hc->elig_count) h->last_mask_pop = (int64_t)hc->elig_count - 1;
  if (c->auto_tensor_q > 0) {  // keep the cost-model route only while it pays: < 1/4 of its queries falling back
    const int64_t aq = (h->auto_q += c->auto_tensor_q), af = (h->auto_fb += hc->uncertified);
    if (aq >= 32 && af * 4 > aq) h->auto_tensor_off = true;
  }
  if ((rc = flags_to_status(hc->flags)) != GFI_OK) return rc;
  const uint32_t* hcnt = reinterpret_cast<const uint32_t*>(hb + off_cnt);
  const float* hdist = reinterpret_cast<const float*>(hb + off_dist);
  const uint64_t* hids = reinterpret_cast<const uint64_t*>(hb + off_ids);
  for (int64_t i = 0; i < q; ++i) {
    out_counts[i] = hcnt[i];
    memcpy(out_ids + i * kstride, hids + (size_t)i * kout, (size_t)hcnt[i] * 8);
    memcpy(out_dist + i * kstride, hdist + (size_t)i * kout, (size_t)hcnt[i] * 4);
  }
  if (hc->unproven > 0) {
    // scan-path answers the device could not prove exact (near-duplicate clusters): proven here, query by query
    const uint32_t n_up = std::min<uint32_t>(hc->unproven, (uint32_t)q);
    std::vector<uint32_t> up(n_up);
    rel.in_flight = true;
    CU_TRY(cudaMemcpyAsync(up.data(), c->up_list.p, (size_t)n_up * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    std::vector<std::pair<float, uint64_t>> res;
    for (uint32_t qi : up) {
      if (qi >= (uint32_t)q) continue;
      if ((rc = prove_query(h, c, st, queries + (size_t)qi * dim, ks[qi], a.d_mask, a.mask_bits, a.mask_by_slot,
                            &res)) != GFI_OK) {
        cudaStreamSynchronize(st);
        return rc;
      }
      out_counts[qi] = (uint32_t)res.size();
      for (size_t j = 0; j < res.size(); ++j) {
        out_dist[qi * kstride + (int64_t)j] = res[j].first;
        out_ids[qi * kstride + (int64_t)j] = res[j].second;
      }
    }
    rel.in_flight = false;
    h->n_paged_q += n_up;
  }
  if (host_trace)
    fprintf(stderr, "[gfi trace] q=%lld enqueue %.1f us, wait %.1f us, unpack %.1f us\n", (long long)q, t_enq - t_begin,
            t_sync - t_enq, now_us() - t_sync);
  return GFI_OK;
}

}  // extern "C"

// ---- hooks for sharded.cu (namespace gfi, declared in host.h) ----
namespace gfi {

SearchCtx* host_acquire_ctx(gfi_index* h) { return acquire_ctx(h); }
void host_release_ctx(gfi_index* h, SearchCtx* c) { release_ctx(h, c); }
int32_t host_flags_to_status(uint32_t flags) { return flags_to_status(flags); }
int32_t host_ensure_flushed(gfi_index* h, bool with_metadata) { return ensure_flushed(h, with_metadata); }
bool host_needs_flush(gfi_index* h, bool with_metadata) {
  std::shared_lock<std::shared_mutex> g(h->mu);
  return h->st_n > 0 || h->live_dirty_lo >= 0 || h->needs_reorder || (with_metadata && h->meta_dirty);
}
int64_t host_row_bytes(const gfi_index* h) {
  if (h->shards) return sharded_row_bytes(h);
  std::shared_lock<std::shared_mutex> g(h->mu);  // (writers mutate these under the unique lock)
  return h->n_slots * (int64_t)h->dpad * 4;
}

void shard_account(gfi_index* h, SearchCtx* c, const Ctrl& hc) {
  prof_collect(h, c);
  h->n_fallback_q += hc.uncertified;
  h->n_tensor_q -= hc.routed_scan;
  h->n_scan_q += hc.routed_scan;
  if (hc.elig_count) h->last_mask_pop = (int64_t)hc.elig_count - 1;
  if (c->auto_tensor_q > 0) {
    const int64_t aq = (h->auto_q += c->auto_tensor_q), af = (h->auto_fb += hc.uncertified);
    if (aq >= 32 && af * 4 > aq) h->auto_tensor_off = true;
  }
}

// The shard-local half of a sharded search: same pipeline as search_impl, but the inputs come from pinned host
// memory shared by all shards (or from the root GPU), and the results go to wherever the caller points -- for a
// shard on another GPU that is peer memory of the root GPU, written by the finalize kernels themselves.
int32_t shard_enqueue(gfi_index* h, SearchCtx* c, const ShardSearch& s) {
  int32_t rc;
  std::shared_lock<std::shared_mutex> g(h->mu);  // (the sharded handle's own lock keeps writers away until the end)
  ++h->n_search;
  h->n_queries += s.q;
  bool empty;
  if ((rc = precheck(h, s.dim, &empty)) != GFI_OK) return rc;
  FilterProgram prog{};
  if (s.filter_json) {
    std::string err;
    if (!compile_filter(s.filter_json, h->meta_fields, h->meta_values, &prog, &err))
      return fail(GFI_ERR_INDEX, "bad filter: " + err);
    for (int i = 0; i < prog.n; ++i)
      if (prog.ops[i].field >= (int)h->meta_dcols.size()) prog.ops[i].field = -1;
  }
  if ((rc = set_device(h)) != GFI_OK) return rc;
  cudaStream_t st = c->stream;
  if (s.wait_a) CU_TRY(cudaStreamWaitEvent(st, s.wait_a, 0));
  if (s.wait_b) CU_TRY(cudaStreamWaitEvent(st, s.wait_b, 0));
  CU_TRY(c->ctrl.ensure(sizeof(Ctrl)));
  auto finish = [&]() -> int32_t {
    // (a one-warp kernel in the PDL chain: a 64-byte peer copy through the copy engine costs a stream ~15 us)
    CU_TRY(launch_copy_words(reinterpret_cast<uint32_t*>(s.out_ctrl), c->ctrl.as<uint32_t>(), (int)(sizeof(Ctrl) / 4), st));
    if (s.done) CU_TRY(cudaEventRecord(s.done, st));
    return GFI_OK;
  };
  if (empty || s.kmax == 0) {
    CU_TRY(cudaMemsetAsync(s.out_counts, 0, (size_t)s.q * 4, st));
    if (!s.keep_flags) CU_TRY(cudaMemsetAsync(c->ctrl.p, 0, sizeof(Ctrl), st));
    return finish();
  }
  const int64_t q = s.q, dim = s.dim;
  const bool masked = s.mask != nullptr || s.filter_json != nullptr;
  const int64_t mask_bits = s.filter_json ? h->n_slots : s.mask_bits;
  const size_t mask_words = masked ? (size_t)((mask_bits + 63) / 64) : 0;
  CU_TRY(c->q_in.ensure((size_t)q * dim * 4));
  CU_TRY(c->ks.ensure((size_t)q * 4));
  if (masked) CU_TRY(c->mask.ensure(mask_words * 8 + 8));
  // Host inputs (pinned) come in through the copy engine.  Inputs that live on a GPU -- the root's memory, a peer of
  // every shard -- are read in place by prep_queries over NVLink (it writes the padded fp32 / fp16 copies and the
  // per-query k locally): no copy-engine operation sits between the kernels of a device-resident search.
  // (The per-query k -- a few KB in the sharded handle's pinned block or on the root GPU -- is always read in place.)
  const bool direct = s.on_device;
  if (!direct) CU_TRY(cudaMemcpyAsync(c->q_in.p, s.queries, (size_t)q * dim * 4, cudaMemcpyDefault, st));
  if (s.mask) CU_TRY(cudaMemcpyAsync(c->mask.p, s.mask, mask_words * 8, cudaMemcpyDefault, st));
  if (s.filter_json) {
    CU_TRY(cudaMemsetAsync(c->mask.p, 0, mask_words * 8 + 8, st));
    CU_TRY(launch_eval_filter(prog, h->meta_dptrs.as<const uint32_t*>(), h->meta_synced_slots, h->n_slots,
                              c->mask.as<uint64_t>(), st));
    ++h->n_launch;
  }
  SearchArgs a{direct ? s.queries : c->q_in.as<float>(), q, c->ks.as<uint32_t>(), s.kmax,
               masked ? c->mask.as<uint64_t>() : nullptr, mask_bits, s.out_ids, s.out_dist, s.out_counts, s.kstride};
  a.h_ks_in = s.ks;  // prep_queries copies the per-query k into c->ks
  a.mask_by_slot = s.filter_json != nullptr;
  if (s.mask && s.mask_density >= 0.0) a.mask_popcount = (int64_t)(s.mask_density * (double)h->n_slots);
  a.keep_flags = s.keep_flags;
  c->zc_pending = false;
  c->ctrl_dev = c->ctrl.p;
  rc = enqueue_search(h, c, a, st);
  c->ctrl_dev = nullptr;
  if (rc != GFI_OK) { cudaStreamSynchronize(st); return rc; }
  return finish();
}

}  // namespace gfi

extern "C" {

// k beyond the kernels' list capacity (FlatIndex::search accepts any k, src/flat_index.rs:63).  Queries with
// k <= kMaxListK run as one ordinary batch; each larger one is answered in passes of kMaxListK results: a pass is an
// exact search over the rows not returned yet (slot-indexed eligibility mask), so the concatenation of the passes
// is the exact ascending top-k.  ceil(k / kMaxListK) scans per such query -- the rare path, kept simple.
constexpr uint32_t kMaxListK = 1016;
// The metadata filter on the host, for the large-k passes only: same postfix program and truth table as
// eval_filter_kernel (filter.cu), over the host mirror of the field columns.
static bool eval_filter_host(const FilterProgram& prog, const std::vector<std::vector<uint32_t>>& cols, int64_t s) {
  bool stack[kFilterMaxDepth];
  int sp = 0;
  for (int i = 0; i < prog.n; ++i) {
    const FilterOp& op = prog.ops[i];
    if (op.kind <= kFilterExists) {
      uint32_t code = 0u;  // no column (yet) or a slot beyond it: the field is absent
      if (op.field >= 0 && (size_t)op.field < cols.size() && (size_t)s < cols[(size_t)op.field].size())
        code = cols[(size_t)op.field][(size_t)s];
      stack[sp++] = op.kind == kFilterEq ? (code != 0u && code == op.code)
                    : op.kind == kFilterNe ? !(code != 0u && code == op.code) : code != 0u;
    } else {
      bool acc = op.kind == kFilterAnd;
      for (uint32_t c = 0; c < op.code; ++c) {
        const bool v = stack[--sp];
        acc = op.kind == kFilterAnd ? (acc && v) : (acc || v);
      }
      stack[sp++] = acc;
    }
  }
  return sp > 0 ? stack[sp - 1] : true;
}

static int32_t search_big_k(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                            const uint64_t* cmask, int64_t cmask_bits, const char* filter_json, uint64_t* out_ids,
                            float* out_dist, uint32_t* out_counts, int64_t kstride) {
  // cmask: the caller's eligibility mask by internal id (or null).  The passes work on a SLOT-indexed mask, so the
  // caller's bits are carried over slot by slot through the id runs before the first pass.  filter_json (instead
  // of a mask): the filter is evaluated once per query over the host mirror of the metadata columns.
  if (!queries || !ks || !out_counts || !out_ids || !out_dist) return fail(GFI_ERR_INDEX, "bad arguments");
  std::vector<int64_t> small, big;
  for (int64_t i = 0; i < q; ++i) {
    if ((int64_t)ks[i] > kstride) return fail(GFI_ERR_INDEX, "kstride smaller than max k");
    (ks[i] > kMaxListK ? big : small).push_back(i);
  }
  int32_t rc;
  if (!small.empty()) {
    uint32_t km = 1;
    for (int64_t i : small) km = std::max(km, ks[i]);
    std::vector<float> qs(small.size() * (size_t)dim);
    std::vector<uint32_t> k2(small.size()), cnt(small.size());
    std::vector<uint64_t> ids(small.size() * (size_t)km);
    std::vector<float> dist(small.size() * (size_t)km);
    for (size_t j = 0; j < small.size(); ++j) {
      memcpy(qs.data() + j * dim, queries + small[j] * dim, (size_t)dim * 4);
      k2[j] = ks[small[j]];
    }
    rc = search_impl(h, qs.data(), (int64_t)small.size(), dim, k2.data(), cmask, cmask_bits, filter_json, ids.data(),
                     dist.data(), cnt.data(), km);
    if (rc != GFI_OK) return rc;
    for (size_t j = 0; j < small.size(); ++j) {
      const int64_t i = small[j];
      out_counts[i] = cnt[j];
      memcpy(out_ids + i * kstride, ids.data() + j * km, (size_t)cnt[j] * 8);
      memcpy(out_dist + i * kstride, dist.data() + j * km, (size_t)cnt[j] * 4);
    }
  }
  std::vector<uint64_t> tmp_ids(kMaxListK);
  std::vector<float> tmp_dist(kMaxListK);
  for (int64_t i : big) {
    for (int attempt = 0;; ++attempt) {
      if ((rc = ensure_flushed(h, filter_json != nullptr)) != GFI_OK) return rc;
      const uint64_t gen = h->layout_gen.load();
      int64_t n_slots;
      {
        std::shared_lock<std::shared_mutex> g(h->mu);
        n_slots = h->n_slots;
      }
      std::vector<uint64_t> mask((size_t)(n_slots + 63) / 64 + 1, (cmask || filter_json) ? 0ull : ~0ull);
      bool meta_stale = false;
      if (filter_json) {
        std::shared_lock<std::shared_mutex> g(h->mu);
        if (h->meta_dirty || h->n_slots != n_slots) {
          meta_stale = true;  // a writer slipped in between the flush and this lock: flush again
        } else {
          FilterProgram prog{};
          std::string err;
          if (!compile_filter(filter_json, h->meta_fields, h->meta_values, &prog, &err))
            return fail(GFI_ERR_INDEX, "bad filter: " + err);
          for (int64_t sl = 0; sl < n_slots; ++sl)
            if (eval_filter_host(prog, h->meta_cols, sl)) mask[(size_t)sl >> 6] |= 1ull << (sl & 63);
        }
      }
      if (meta_stale) {
        if (attempt >= 3) return fail(GFI_ERR_INDEX, "index kept changing during a large-k filtered search");
        continue;
      }
      if (cmask) {
        std::shared_lock<std::shared_mutex> g(h->mu);
        for (const auto& kv : h->runs) {
          uint64_t id = kv.first;
          for (uint32_t j = 0; j < kv.second.n; ++j, ++id) {
            const uint32_t sl = kv.second.slot0 + j;
            if ((int64_t)sl < n_slots && id < (uint64_t)cmask_bits && ((cmask[id >> 6] >> (id & 63)) & 1ull))
              mask[sl >> 6] |= 1ull << (sl & 63);
          }
        }
      }
      uint32_t got = 0;
      bool stale = false;
      while (got < ks[i]) {
        const uint32_t kp = std::min<uint32_t>(kMaxListK, ks[i] - got);
        uint32_t cnt = 0;
        tl_mask_by_slot = n_slots > 0;
        rc = search_impl(h, queries + i * dim, 1, dim, &kp, n_slots > 0 ? mask.data() : nullptr, n_slots, nullptr,
                         tmp_ids.data(), tmp_dist.data(), &cnt, kMaxListK);
        tl_mask_by_slot = false;
        if (rc != GFI_OK) return rc;
        {
          std::shared_lock<std::shared_mutex> g(h->mu);
          if (h->layout_gen.load() != gen) { stale = true; break; }  // rows moved (compaction): start over
          for (uint32_t j = 0; j < cnt; ++j) {
            uint32_t slot;
            if (lookup_slot(h, tmp_ids[j], &slot) && (int64_t)slot < n_slots) mask[slot >> 6] &= ~(1ull << (slot & 63));
          }
        }
        memcpy(out_ids + i * kstride + got, tmp_ids.data(), (size_t)cnt * 8);
        memcpy(out_dist + i * kstride + got, tmp_dist.data(), (size_t)cnt * 4);
        got += cnt;
        if (cnt < kp) break;  // fewer rows than asked for: FlatIndex::search returns what exists
      }
      if (!stale) { out_counts[i] = got; break; }
      if (attempt >= 3) return fail(GFI_ERR_INDEX, "index kept being compacted during a large-k search");
    }
  }
  return GFI_OK;
}

// Runs a batch of queued plain searches as one search and hands every request its own rows, status and error
// payload.  Any failure of the combined search (one zero-norm or NaN query fails a whole batch, as it would in
// VectorStore::search_batch) is resolved by running the requests one by one, so a caller only ever sees the
// outcome of its own call.
static void run_coalesced(gfi_index* h, std::vector<gfi_index::CoReq*>& batch) {
  auto run_one = [&](gfi_index::CoReq* r) {
    r->rc = search_impl(h, r->queries, r->q, r->dim, r->ks, nullptr, 0, nullptr, r->out_ids, r->out_dist,
                        r->out_counts, r->kstride);
    if (r->rc != GFI_OK) { r->err = tl_error; r->exp = tl_expected; r->act = tl_actual; }
    r->answered = true;
  };
  bool combine = batch.size() > 1;
  int64_t total = 0;
  uint32_t kmax = 1;
  for (auto* r : batch) {
    if (r->dim != batch[0]->dim || r->q <= 0) combine = false;
    total += r->q;
    for (int64_t i = 0; i < r->q; ++i) kmax = std::max(kmax, r->ks[i]);
  }
  if (combine) try {
    const int64_t dim = batch[0]->dim;
    std::vector<float> qs((size_t)total * dim);
    std::vector<uint32_t> ks((size_t)total), cnt((size_t)total);
    std::vector<uint64_t> ids((size_t)total * kmax);
    std::vector<float> dist((size_t)total * kmax);
    int64_t at = 0;
    for (auto* r : batch) {
      memcpy(qs.data() + (size_t)at * dim, r->queries, (size_t)r->q * dim * 4);
      memcpy(ks.data() + at, r->ks, (size_t)r->q * 4);
      at += r->q;
    }
    const int32_t rc = search_impl(h, qs.data(), total, dim, ks.data(), nullptr, 0, nullptr, ids.data(), dist.data(),
                                   cnt.data(), kmax);
    if (rc == GFI_OK) {
      at = 0;
      for (auto* r : batch) {
        for (int64_t i = 0; i < r->q; ++i, ++at) {
          const uint32_t c = cnt[at];  // <= ks[i] <= the request's own kstride (validated before queueing)
          r->out_counts[i] = c;
          memcpy(r->out_ids + i * r->kstride, ids.data() + (size_t)at * kmax, (size_t)c * 8);
          memcpy(r->out_dist + i * r->kstride, dist.data() + (size_t)at * kmax, (size_t)c * 4);
        }
        r->rc = GFI_OK;
        r->answered = true;
      }
      ++h->n_co_batches;
      h->n_co_requests += (int64_t)batch.size();
      return;
    }
  } catch (const std::exception&) {  // staging allocation failed: serve the requests one by one
  }
  for (auto* r : batch) run_one(r);
}

int32_t gfi_search(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                   const uint64_t* mask, int64_t mask_bits, uint64_t* out_ids, float* out_dist,
                   uint32_t* out_counts, int64_t kstride) {
  if (h && q > 0 && ks) {
    bool bigk = false;
    for (int64_t i = 0; i < q; ++i) bigk = bigk || ks[i] > kMaxListK;
    if (bigk)
      return h->shards ? sharded_search_big_k(h, queries, q, dim, ks, mask, mask_bits, nullptr, out_ids, out_dist, out_counts, kstride)
                       : search_big_k(h, queries, q, dim, ks, mask, mask_bits, nullptr, out_ids, out_dist, out_counts, kstride);
  }
  // masked searches, large batches and malformed calls take the direct path
  // (indexes below 32 MB are launch-latency-bound: independent calls on separate streams overlap on the GPU and
  // beat a serialised batch -- measured 68-75k vs 53k q/s at 10k x 128 -- so only large ones are coalesced)
  bool plain = h && !mask && q > 0 && q <= 256 && queries && ks && out_counts && out_ids && out_dist;
  if (plain) {
    plain = h->opt_coalesce && host_row_bytes(h) >= (32ll << 20);
  }
  if (plain)
    for (int64_t i = 0; i < q; ++i)
      if ((int64_t)ks[i] > kstride) { plain = false; break; }
  if (!plain)
    return search_impl(h, queries, q, dim, ks, mask, mask_bits, nullptr, out_ids, out_dist, out_counts, kstride);
  gfi_index::CoReq me{queries, q, dim, ks, out_ids, out_dist, out_counts, kstride};
  {
    std::unique_lock<std::mutex> lk(h->co_mu);
    if (h->co_busy) {
      h->co_pending.push_back(&me);
      me.cv.wait(lk, [&] { return me.done || me.lead; });
      if (me.done) {
        if (me.rc != GFI_OK) { tl_error = me.err; tl_expected = me.exp; tl_actual = me.act; }
        return me.rc;
      }
    } else {
      h->co_busy = true;
      me.batch.push_back(&me);
    }
  }
  // leader: run my batch, then hand the baton to the first queued request (with everything queued so far)
  std::vector<gfi_index::CoReq*> batch = std::move(me.batch);
  try {
    run_coalesced(h, batch);
  } catch (...) {
    // nothing may escape with co_busy set (every later plain search would wait forever): whatever was not answered
    // fails with an index error, and the baton is handed over below as usual
    for (auto* r : batch)
      if (r->rc == GFI_OK && !r->answered) { r->rc = GFI_ERR_INDEX; r->err = "search failed (out of memory?)"; }
  }
  {
    std::lock_guard<std::mutex> lk(h->co_mu);
    // (notified while the lock is held: a woken request cannot return -- and destroy its CoReq -- before the lock is
    // released, and nobody else is woken: with one shared condition variable every hand-over woke all waiting clients)
    for (auto* r : batch)
      if (r != &me) { r->done = true; r->cv.notify_one(); }
    if (h->co_pending.empty()) {
      h->co_busy = false;
    } else {
      gfi_index::CoReq* next = h->co_pending.front();
      int64_t total = 0;
      size_t take = 0;
      while (take < h->co_pending.size() && (take == 0 || total + h->co_pending[take]->q <= 4096)) total += h->co_pending[take++]->q;
      next->batch.assign(h->co_pending.begin(), h->co_pending.begin() + (long)take);
      h->co_pending.erase(h->co_pending.begin(), h->co_pending.begin() + (long)take);
      next->lead = true;
      next->cv.notify_one();
    }
  }
  if (me.rc != GFI_OK) { tl_error = me.err; tl_expected = me.exp; tl_actual = me.act; }
  return me.rc;
}

int32_t gfi_search_filtered(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                            const char* filter_json, uint64_t* out_ids, float* out_dist, uint32_t* out_counts,
                            int64_t kstride) {
  if (!filter_json) return fail(GFI_ERR_INDEX, "null filter");
  if (h && q > 0 && ks) {  // k beyond the kernels' list capacity: exact passes over the rows not returned yet
    bool bigk = false;
    for (int64_t i = 0; i < q; ++i) bigk = bigk || ks[i] > kMaxListK;
    if (bigk)
      return h->shards ? sharded_search_big_k(h, queries, q, dim, ks, nullptr, 0, filter_json, out_ids, out_dist, out_counts, kstride)
                       : search_big_k(h, queries, q, dim, ks, nullptr, 0, filter_json, out_ids, out_dist, out_counts, kstride);
  }
  return search_impl(h, queries, q, dim, ks, nullptr, 0, filter_json, out_ids, out_dist, out_counts, kstride);
}

int32_t gfi_set_metadata(gfi_index* h, uint64_t id, int32_t n_fields, const char* const* keys,
                         const char* const* values) {
  if (!h || n_fields < 0 || (n_fields > 0 && (!keys || !values))) return fail(GFI_ERR_INDEX, "bad arguments");
  if (h->shards) return sharded_set_metadata(h, id, n_fields, keys, values);
  std::unique_lock<std::shared_mutex> g(h->mu);
  std::vector<std::pair<int, uint32_t>> enc;
  enc.reserve((size_t)n_fields);
  for (int i = 0; i < n_fields; ++i) {
    auto it = h->meta_fields.find(keys[i]);
    int f;
    if (it == h->meta_fields.end()) {
      f = (int)h->meta_values.size();
      h->meta_fields[keys[i]] = f;
      h->meta_values.emplace_back();
    } else {
      f = it->second;
    }
    auto& dict = h->meta_values[(size_t)f];
    auto vt = dict.find(values[i]);
    uint32_t code;
    if (vt == dict.end()) {
      code = (uint32_t)dict.size() + 1u;
      dict[values[i]] = code;
    } else {
      code = vt->second;
    }
    enc.emplace_back(f, code);
  }
  h->meta_pending[id] = std::move(enc);
  h->meta_dirty = true;
  return GFI_OK;
}

// Bulk form for ONE field over many rows (loading the reference's `HashMap<usize, Metadata>`, src/storage.rs:90,
// at ingest speed instead of one call per row).  Stored rows are written straight into the field's column; ids that
// are not stored yet go through the pending map like gfi_set_metadata's.
int32_t gfi_set_metadata_column(gfi_index* h, const char* key, const uint64_t* ids, int64_t n,
                                const char* const* values, int32_t n_values, const uint32_t* codes) {
  if (!h || !key || n < 0 || n_values < 0 || (n > 0 && (!ids || !codes)) || (n_values > 0 && !values))
    return fail(GFI_ERR_INDEX, "bad arguments");
  if (h->shards) return sharded_set_metadata_column(h, key, ids, n, values, n_values, codes);
  std::unique_lock<std::shared_mutex> g(h->mu);
  int32_t rc = flush_locked(h);
  if (rc != GFI_OK) return rc;
  if (h->needs_reorder && (rc = compact_locked(h)) != GFI_OK) return rc;
  int f;
  auto it = h->meta_fields.find(key);
  if (it == h->meta_fields.end()) {
    f = (int)h->meta_values.size();
    h->meta_fields[key] = f;
    h->meta_values.emplace_back();
  } else {
    f = it->second;
  }
  auto& dict = h->meta_values[(size_t)f];
  std::vector<uint32_t> code_of((size_t)n_values);
  for (int32_t v = 0; v < n_values; ++v) {
    auto vt = dict.find(values[v]);
    if (vt == dict.end()) {
      code_of[(size_t)v] = (uint32_t)dict.size() + 1u;
      dict[values[v]] = code_of[(size_t)v];
    } else {
      code_of[(size_t)v] = vt->second;
    }
  }
  h->meta_cols.resize(h->meta_values.size());
  for (auto& c : h->meta_cols) c.resize((size_t)h->n_slots, 0u);
  auto& col = h->meta_cols[(size_t)f];
  for (int64_t i = 0; i < n; ++i) {
    if (codes[i] != UINT32_MAX && codes[i] >= (uint32_t)n_values) return fail(GFI_ERR_INDEX, "metadata code out of range");
    const uint32_t code = codes[i] == UINT32_MAX ? 0u : code_of[codes[i]];
    uint32_t slot;
    if (lookup_slot(h, ids[i], &slot)) {
      col[slot] = code;
    } else if (code) {  // not stored yet: merged into whatever is pending for the id
      auto& pend = h->meta_pending[ids[i]];
      bool found = false;
      for (auto& fv : pend)
        if (fv.first == f) { fv.second = code; found = true; }
      if (!found) pend.emplace_back(f, code);
    }
  }
  h->meta_dirty = true;
  return GFI_OK;
}

static thread_local SearchCtx* tl_dev_ctx = nullptr;
static thread_local gfi_index* tl_dev_owner = nullptr;

int32_t gfi_search_device(gfi_index* h, const float* d_queries, int64_t q, const uint32_t* d_ks, uint32_t kmax,
                          const uint64_t* d_mask, int64_t mask_bits, uint64_t* d_out_ids, float* d_out_dist,
                          uint32_t* d_out_counts, int64_t kstride, void* stream) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (q <= 0) return GFI_OK;
  if (h->shards)
    return sharded_search_device(h, d_queries, q, d_ks, kmax, d_mask, mask_bits, d_out_ids, d_out_dist, d_out_counts,
                                 kstride, stream);
  int32_t rc = ensure_flushed(h);
  if (rc != GFI_OK) return rc;
  std::shared_lock<std::shared_mutex> g(h->mu);
  ++h->n_search;
  h->n_queries += q;
  if ((rc = set_device(h)) != GFI_OK) return rc;
  if ((int64_t)kmax > kstride) return fail(GFI_ERR_INDEX, "kstride smaller than max k");
  if (tl_dev_ctx && tl_dev_owner != h) return fail(GFI_ERR_INDEX, "collect gfi_search_status first");
  if (!tl_dev_ctx) {
    tl_dev_ctx = acquire_ctx(h);
    tl_dev_owner = h;
    if (!tl_dev_ctx) return fail(GFI_ERR_INDEX, "cannot create a CUDA stream");
  }
  SearchCtx* c = tl_dev_ctx;
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  // The context's workspace and control block are shared by every search issued before the next
  // gfi_search_status: they are stream-ordered as long as the stream stays the same; a search on ANOTHER stream is
  // ordered behind the uncollected ones with an event, and the error flags are sticky until collected.
  if (c->pending_status && c->last_stream != st) {
    if (!c->order_ev) CU_TRY(cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(c->order_ev, c->last_stream));
    CU_TRY(cudaStreamWaitEvent(st, c->order_ev, 0));
  }
  const bool had_pending = c->pending_status;
  c->last_stream = st;
  CU_TRY(c->ctrl.ensure(sizeof(Ctrl)));
  if (h->n_live == 0 || kmax == 0) {
    CU_TRY(cudaMemsetAsync(d_out_counts, 0, (size_t)q * 4, st));
    if (!had_pending) CU_TRY(cudaMemsetAsync(c->ctrl.p, 0, sizeof(Ctrl), st));
    c->pending_status = true;
    return GFI_OK;
  }
  SearchArgs a{d_queries, q, d_ks, kmax, d_mask, mask_bits, d_out_ids, d_out_dist, d_out_counts, kstride};
  a.keep_flags = had_pending;
  c->ctrl_dev = c->ctrl.p;
  rc = enqueue_search(h, c, a, st);
  c->ctrl_dev = nullptr;
  c->pending_status = true;
  return rc;
}

int32_t gfi_search_status(gfi_index* h) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (h->shards) return sharded_search_status(h);
  SearchCtx* c = tl_dev_ctx;
  if (!c || tl_dev_owner != h) return GFI_OK;
  tl_dev_ctx = nullptr;
  tl_dev_owner = nullptr;
  struct Releaser { gfi_index* h; SearchCtx* c; ~Releaser() { release_ctx(h, c); } } rel{h, c};
  int32_t rc;
  if ((rc = set_device(h)) != GFI_OK) return rc;
  cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
  c->pending_status = false;
  c->last_stream = nullptr;
  if (!c->ctrl.p) return GFI_OK;
  CU_TRY(c->h_ctrl.ensure(sizeof(Ctrl)));
  CU_TRY(cudaMemcpyAsync(c->h_ctrl.p, c->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  prof_collect(h, c);
  const Ctrl* hc = c->h_ctrl.as<Ctrl>();
  h->n_fallback_q += hc->uncertified;
  h->n_tensor_q -= hc->routed_scan;
  h->n_scan_q += hc->routed_scan;
  if (hc->elig_count) h->last_mask_pop = (int64_t)hc->elig_count - 1;
  if (c->auto_tensor_q > 0) {  // same bookkeeping as the host path (the block describes the last search issued)
    const int64_t aq = (h->auto_q += c->auto_tensor_q), af = (h->auto_fb += hc->uncertified);
    if (aq >= 32 && af * 4 > aq) h->auto_tensor_off = true;
  }
  if ((rc = flags_to_status(hc->flags)) != GFI_OK) return rc;
  if (hc->unproven)
    return fail(GFI_ERR_UNPROVEN, std::to_string(hc->unproven) + " quer(ies) since the last status could not be proven "
                "exact on the device (near-duplicate rows around the k-th distance): re-run them through gfi_search");
  return GFI_OK;
}

int32_t gfi_merge_topk_device(const uint64_t* d_ids, const float* d_dist, const uint32_t* d_counts, int32_t G,
                              int64_t q, int64_t kstride, const uint32_t* d_ks, uint64_t* d_out_ids,
                              float* d_out_dist, uint32_t* d_out_counts, int64_t out_kstride, void* stream) {
  CU_TRY(launch_merge(d_ids, d_dist, d_counts, G, q, kstride, 0, d_ks, d_out_ids, d_out_dist, d_out_counts,
                      out_kstride, (cudaStream_t)stream));
  return GFI_OK;
}

int32_t gfi_merge_topk_device_strided(const uint64_t* d_ids, const float* d_dist, const uint32_t* d_counts,
                                      int32_t G, int64_t q, int64_t kstride, int64_t shard_stride_bytes,
                                      const uint32_t* d_ks, uint64_t* d_out_ids, float* d_out_dist,
                                      uint32_t* d_out_counts, int64_t out_kstride, void* stream) {
  if (shard_stride_bytes <= 0 || (shard_stride_bytes & 7)) return fail(GFI_ERR_INDEX, "shard stride must be a positive multiple of 8 bytes");
  CU_TRY(launch_merge(d_ids, d_dist, d_counts, G, q, kstride, shard_stride_bytes, d_ks, d_out_ids, d_out_dist,
                      d_out_counts, out_kstride, (cudaStream_t)stream));
  return GFI_OK;
}

int32_t gfi_distances(gfi_index* h, const float* queries, int64_t q, int64_t dim, const uint64_t* cand_ids, int64_t m,
                      float* out_dist, uint8_t* out_status) {
  if (!h) return fail(GFI_ERR_INDEX, "null handle");
  if (q < 0 || m < 0 || (q > 0 && m > 0 && (!queries || !cand_ids || !out_dist))) return fail(GFI_ERR_INDEX, "bad arguments");
  if (q == 0 || m == 0) return GFI_OK;
  if (h->shards) return sharded_distances(h, queries, q, dim, cand_ids, m, out_dist, out_status);
  int32_t rc = ensure_flushed(h);
  if (rc != GFI_OK) return rc;
  std::shared_lock<std::shared_mutex> g(h->mu);
  bool empty;
  if ((rc = precheck(h, dim, &empty)) != GFI_OK) return rc;
  const float kInf = std::numeric_limits<float>::infinity();
  if (empty) {  // nothing stored: every id is absent
    for (int64_t i = 0; i < q * m; ++i) { out_dist[i] = kInf; if (out_status) out_status[i] = 1; }
    return GFI_OK;
  }
  if ((rc = set_device(h)) != GFI_OK) return rc;
  SearchCtx* c = acquire_ctx(h);
  if (!c) return fail(GFI_ERR_INDEX, "cannot create a CUDA stream");
  struct Releaser { gfi_index* h; SearchCtx* c; ~Releaser() { release_ctx(h, c); } } rel{h, c};
  cudaStream_t st = c->stream;
  const size_t nq = (size_t)q, np = (size_t)q * (size_t)m;
  CU_TRY(c->q_in.ensure(nq * dim * 4));
  CU_TRY(c->q32.ensure(nq * h->dpad * 4));
  CU_TRY(c->qnorm.ensure(nq * 4));
  CU_TRY(c->qsumsq.ensure(nq * 4));
  CU_TRY(c->ctrl.ensure(sizeof(Ctrl)));
  CU_TRY(c->cand.ensure(np * 8));                 // candidate ids
  CU_TRY(c->out_dist.ensure(np * 4));
  CU_TRY(c->out_counts.ensure(np));               // status bytes
  CU_TRY(c->h_dist.ensure(np * 4));
  CU_TRY(c->h_counts.ensure(np));
  CU_TRY(cudaMemsetAsync(c->ctrl.p, 0, sizeof(Ctrl), st));
  CU_TRY(cudaMemcpyAsync(c->q_in.p, queries, nq * dim * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(c->cand.p, cand_ids, np * 8, cudaMemcpyHostToDevice, st));
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(c->ctrl.p);
  PrepQueriesParams pq{};
  pq.q_in = c->q_in.as<float>();
  pq.q32 = c->q32.as<float>();
  pq.q16 = nullptr;
  pq.qnorm = c->qnorm.as<float>();
  pq.qsumsq = c->qsumsq.as<float>();
  pq.qmaxabs = reinterpret_cast<float*>(&ctrl->qmaxabs_bits);
  pq.q = (int)q;
  pq.qpad = (int)q;
  pq.d = (int)h->dim;
  pq.dpad = h->dpad;
  pq.dpad16 = h->dpad16;
  CU_TRY(launch_prep_queries(pq, st));
  ScorePairsParams sp{};
  sp.iv = h->view();
  sp.q32 = c->q32.as<float>();
  sp.qnorm = c->qnorm.as<float>();
  sp.cand_ids = c->cand.as<uint64_t>();
  sp.q = q;
  sp.m = m;
  sp.out_dist = c->out_dist.as<float>();
  sp.out_status = c->out_counts.as<uint8_t>();
  sp.flags = &ctrl->flags;
  CU_TRY(launch_score_pairs(sp, st));
  h->n_launch += 2;
  CU_TRY(cudaMemcpyAsync(c->h_dist.p, c->out_dist.p, np * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(c->h_counts.p, c->out_counts.p, np, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  memcpy(out_dist, c->h_dist.p, np * 4);
  const uint8_t* hs = reinterpret_cast<const uint8_t*>(c->h_counts.p);
  if (out_status) memcpy(out_status, hs, np);
  else
    for (size_t i = 0; i < np; ++i)  // no per-pair status wanted: the first error fails the call, as `?` would
      if (hs[i] == 2) return fail(GFI_ERR_INVALID_VECTOR, "Cannot compute cosine distance with zero vector");
  return GFI_OK;
}

int32_t gfi_debug_tensor_scores(gfi_index* h, const float* queries, int64_t q, float* out, int64_t out_stride) {
  if (!h || !queries || !out || q <= 0) return fail(GFI_ERR_INDEX, "bad arguments");
  if (h->shards) return fail(GFI_ERR_INDEX, "gfi_debug_tensor_scores: single-GPU handles only");
  int32_t rc = ensure_flushed(h);
  if (rc != GFI_OK) return rc;
  std::shared_lock<std::shared_mutex> g(h->mu);
  if (!h->use_x16 || h->n_slots == 0 || get_encode_fn() == nullptr) return fail(GFI_ERR_INDEX, "tensor path unavailable");
  const int64_t npad = (h->n_slots + 255) / 256 * 256;
  if (out_stride < npad) return fail(GFI_ERR_INDEX, "out_stride must be >= n_slots rounded up to 256");
  if ((rc = set_device(h)) != GFI_OK) return rc;
  SearchCtx* c = acquire_ctx(h);
  if (!c) return fail(GFI_ERR_INDEX, "cannot create a CUDA stream");
  struct Releaser { gfi_index* h; SearchCtx* c; ~Releaser() { release_ctx(h, c); } } rel{h, c};
  cudaStream_t st = c->stream;
  const int qi = (int)q, qpad = (qi + 127) / 128 * 128;
  CU_TRY(c->q_in.ensure((size_t)q * h->dim * 4));
  CU_TRY(c->q32.ensure((size_t)q * h->dpad * 4));
  CU_TRY(c->q16.ensure((size_t)qpad * h->dpad16 * 2));
  CU_TRY(c->qnorm.ensure((size_t)q * 4));
  CU_TRY(c->qsumsq.ensure((size_t)q * 4));
  CU_TRY(c->ctrl.ensure(sizeof(Ctrl)));
  CU_TRY(c->seeds.ensure((size_t)q * out_stride * 4));
  CU_TRY(cudaMemsetAsync(c->ctrl.p, 0, sizeof(Ctrl), st));
  CU_TRY(cudaMemcpyAsync(c->q_in.p, queries, (size_t)q * h->dim * 4, cudaMemcpyHostToDevice, st));
  Ctrl* ctrl = c->ctrl.as<Ctrl>();
  PrepQueriesParams pq{};
  pq.q_in = c->q_in.as<float>(); pq.q32 = c->q32.as<float>(); pq.q16 = c->q16.as<__half>();
  pq.qnorm = c->qnorm.as<float>(); pq.qsumsq = c->qsumsq.as<float>();
  pq.qmaxabs = reinterpret_cast<float*>(&ctrl->qmaxabs_bits);
  pq.q = qi; pq.qpad = qpad; pq.d = (int)h->dim; pq.dpad = h->dpad; pq.dpad16 = h->dpad16;
  CU_TRY(launch_prep_queries(pq, st));
  IndexView iv = h->view();
  CUtensorMap tmx, tmq;
  if (!make_tmap_2d(&tmx, iv.x16, (uint64_t)h->dpad16, (uint64_t)h->n_slots, (uint64_t)h->dpad16 * 2, 64, 256) ||
      !make_tmap_2d(&tmq, c->q16.p, (uint64_t)h->dpad16, (uint64_t)qpad, (uint64_t)h->dpad16 * 2, 64, 128))
    return fail(GFI_ERR_INDEX, "cuTensorMapEncodeTiled failed");
  GemmParams gp{};
  gp.iv = iv;
  gp.mask = MaskView{nullptr, 0};
  gp.qmaxabs = reinterpret_cast<const float*>(&ctrl->qmaxabs_bits);
  gp.qsumsq = c->qsumsq.as<float>();
  gp.q = qi;
  gp.num_m_tiles = qpad / 128;
  gp.num_n_tiles = npad / 256;
  gp.seed_mode = 2;
  gp.seed_stride = out_stride;
  gp.seeds = c->seeds.as<float>();
  gp.flags = &ctrl->flags;
  CU_TRY(launch_gemm_topk(gp, &tmx, &tmq, (int)std::min<int64_t>(h->sm_count, gp.num_n_tiles * gp.num_m_tiles), st));
  CU_TRY(cudaMemcpyAsync(out, c->seeds.p, (size_t)q * out_stride * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  return GFI_OK;
}

int32_t gfi_get_stats(gfi_index* h, gfi_stats* out) {
  if (!h || !out) return fail(GFI_ERR_INDEX, "null argument");
  memset(out, 0, sizeof(*out));
  out->shards = 1;
  if (h->shards) return sharded_get_stats(h, out);
  std::shared_lock<std::shared_mutex> g(h->mu);
  out->n_slots = h->n_slots;
  out->n_live = h->n_live;
  out->searches = h->n_search;
  out->queries = h->n_queries;
  out->scan_queries = h->n_scan_q;
  out->tensor_queries = h->n_tensor_q;
  out->fallback_queries = h->n_fallback_q;
  out->kernel_launches = h->n_launch;
  out->bytes_fp32 = h->n_slots * (int64_t)h->dpad * 4;
  out->bytes_fp16 = h->use_x16 ? h->n_slots * (int64_t)h->dpad16 * 2 : 0;
  out->scan_kernel_ns = h->prof_ns[0];
  out->scan_kernel_count = h->prof_cnt[0];
  out->tensor_kernel_ns = h->prof_ns[1];
  out->tensor_kernel_count = h->prof_cnt[1];
  out->coalesced_batches = h->n_co_batches;
  out->coalesced_requests = h->n_co_requests;
  out->paged_queries = h->n_paged_q;
  return GFI_OK;
}

int32_t gfi_set_option(gfi_index* h, const char* name, int64_t value) {
  if (!h || !name) return fail(GFI_ERR_INDEX, "null argument");
  if (h->shards) return sharded_set_option(h, name, value);
  std::unique_lock<std::shared_mutex> g(h->mu);
  const std::string n(name);
  if (n == "tensor_min_q") h->opt_tensor_min_q = (int)value;
  else if (n == "tensor_min_rows") h->opt_tensor_min_rows = (int)value;
  else if (n == "kp") h->opt_kp = (int)value;
  else if (n == "hits") h->opt_hits = (int)value;
  else if (n == "scan_qt") h->opt_scan_qt = (int)value;
  else if (n == "fused_tail") h->opt_fused_tail = (int)value;
  else if (n == "zero_copy") h->opt_zero_copy = (int)value;
  else if (n == "grid") h->opt_grid = (int)value;
  else if (n == "seed_rank") h->opt_seed_rank = (int)value;
  else if (n == "raw_epilogue") h->opt_raw_epilogue = (int)value;
  else if (n == "pair") h->opt_pair = (int)value;
  else if (n == "short_k") h->opt_short_k = (int)value;
  else if (n == "short_k_min_tiles") h->opt_short_k_min_tiles = (int)value;
  else if (n == "short_k_seed") h->opt_short_k_seed = (int)value;
  else if (n == "tensor_auto") { h->opt_tensor_auto = (int)value; h->auto_tensor_off = false; h->auto_q = 0; h->auto_fb = 0; }
  else if (n == "profile") h->opt_profile = (int)value;
  else if (n == "coalesce") h->opt_coalesce = (int)value;
  else if (n == "gemm_debug") h->opt_gemm_debug = (int)value;
  else if (n == "scan_stages") h->opt_scan_stages = (int)value;
  else if (n == "scan_certify") h->opt_scan_certify = (int)value;
  else if (n == "rerank_cut") h->opt_rerank_cut = (int)value;
  else return fail(GFI_ERR_INDEX, "unknown option: " + n);
  return GFI_OK;
}

}  // extern "C"

namespace {

// Rebuilds the slot array: live rows only, ordered by internal id.  Caller holds the unique lock.
int32_t compact_locked(gfi_index* h) {
  int32_t rc;
  ++h->layout_gen;
  if ((rc = set_device(h)) != GFI_OK) return rc;
  if (h->n_slots == 0) { h->needs_reorder = false; return GFI_OK; }
  // permutation: runs are keyed by id0 and disjoint, so walking the map yields ids in order
  std::vector<uint32_t> perm;
  perm.reserve((size_t)h->n_live);
  std::map<uint64_t, Run> nruns;
  bool identity = true;
  for (auto& kv : h->runs) {
    uint64_t id = kv.first;
    for (uint32_t i = 0; i < kv.second.n; ++i, ++id) {
      const uint32_t s = kv.second.slot0 + i;
      if (!((h->h_live[s >> 5] >> (s & 31)) & 1u)) continue;
      const uint32_t ns = (uint32_t)perm.size();
      perm.push_back(s);
      if (id != ns) identity = false;
      if (!nruns.empty()) {
        auto last = std::prev(nruns.end());
        if (last->first + last->second.n == id && last->second.slot0 + last->second.n == ns) {
          ++last->second.n;
          continue;
        }
      }
      nruns[id] = Run{ns, 1};
    }
  }
  const int64_t n_out = (int64_t)perm.size();
  const int64_t ncap = std::max<int64_t>(1024, (n_out + 255) / 256 * 256);
  DevBuf d_perm, nx32, nx16, nids, nnorm, nsumsq, ncoef, nflags, nlive;
  CU_TRY(d_perm.ensure((size_t)std::max<int64_t>(n_out, 1) * 4));
  CU_TRY(nx32.ensure((size_t)ncap * h->dpad * 4));
  CU_TRY(nids.ensure((size_t)ncap * 8));
  CU_TRY(nnorm.ensure((size_t)ncap * 4));
  CU_TRY(nsumsq.ensure((size_t)ncap * 4));
  CU_TRY(nflags.ensure((size_t)ncap * 4));
  if (h->use_x16) {
    CU_TRY(nx16.ensure((size_t)ncap * h->dpad16 * 2));
    CU_TRY(ncoef.ensure((size_t)ncap * 8));
  }
  const size_t words = (size_t)(ncap + 31) / 32;
  CU_TRY(nlive.ensure(words * 4));
  if (n_out > 0) {
    CU_TRY(cudaMemcpyAsync(d_perm.p, perm.data(), (size_t)n_out * 4, cudaMemcpyHostToDevice, h->ingest_stream));
    CU_TRY(launch_gather_rows(h->view(), d_perm.as<uint32_t>(), n_out, nx32.as<float>(),
                              h->use_x16 ? nx16.as<__half>() : nullptr, nids.as<uint64_t>(), nnorm.as<float>(),
                              nsumsq.as<float>(), h->use_x16 ? ncoef.as<float2>() : nullptr, h->ingest_stream));
  }
  std::vector<uint32_t> nh_live(words, 0u);
  for (int64_t s = 0; s < n_out; ++s) nh_live[(size_t)s >> 5] |= 1u << (s & 31);
  CU_TRY(cudaMemcpyAsync(nlive.p, nh_live.data(), words * 4, cudaMemcpyHostToDevice, h->ingest_stream));
  CU_TRY(cudaMemsetAsync(nflags.p, 0, (size_t)ncap * 4, h->ingest_stream));
  CU_TRY(cudaStreamSynchronize(h->ingest_stream));
  d_perm.release();
  auto swap_in = [](DevBuf& dst, DevBuf& src) { dst.release(); dst = src; src.p = nullptr; src.bytes = 0; };
  swap_in(h->x32, nx32);
  swap_in(h->ids, nids);
  swap_in(h->norm, nnorm);
  swap_in(h->sumsq, nsumsq);
  swap_in(h->rowflags, nflags);
  swap_in(h->live, nlive);
  if (h->use_x16) { swap_in(h->x16, nx16); swap_in(h->coef, ncoef); }
  h->cap = ncap;
  h->n_slots = n_out;
  h->n_live = n_out;
  for (auto& col : h->meta_cols) {
    std::vector<uint32_t> nc((size_t)n_out, 0u);
    for (int64_t j = 0; j < n_out; ++j)
      if ((size_t)perm[(size_t)j] < col.size()) nc[(size_t)j] = col[perm[(size_t)j]];
    col.swap(nc);
  }
  if (!h->meta_cols.empty()) { h->meta_dirty = true; h->meta_synced_slots = 0; }
  h->h_live.swap(nh_live);
  h->runs.swap(nruns);
  h->ids_identity = identity;
  h->needs_reorder = false;
  h->live_dirty_lo = h->live_dirty_hi = -1;
  ++h->n_launch;
  return GFI_OK;
}

}  // namespace
