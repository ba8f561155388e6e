// sharded.cu -- one GpuFlatIndex over several GPUs of a box, in one process, behind the same C ABI handle.
//
// The reference server owns ONE index behind one RwLock (reference src/server/mod.rs:13-16, src/main.rs:152-198), so
// a drop-in `impl Index for GpuFlatIndex` (src/index.rs:11-35) has to reach every GPU from a single handle.  A
// sharded handle is a router over one ordinary single-GPU index ("shard") per device:
//   rows     routed by internal id in blocks: shard = (id / block) % G  (contiguous id ranges after gfi_reserve)
//   workers  one persistent host thread per shard, bound to its GPU: a search posts one job per shard, so the
//            eight pipelines are enqueued concurrently instead of 8 x ~10 launches from one thread
//   exchange none as a separate step: every shard's finalize kernel writes its k candidates (ids, distances, counts)
//            through peer pointers straight into the ROOT GPU's gather block -- posted stores over NVLink/NVSwitch,
//            fused into the kernel that produces them.  The root's stream waits on one event per shard, merges the
//            G sorted lists per query by (distance, id) with merge_topk_kernel and returns ONE block to the host.
//   overlap  device-resident searches (gfi_search_device) cycle through kGather gather blocks: shard g may write the
//            candidates of batches i+1 .. i+3 while the root still merges batch i, so exchange + merge hide behind the
//            next main passes.  No copy-engine operation sits between a shard's kernels: inputs on the root GPU are
//            read in place over NVLink, the control block goes back through a one-warp kernel.
// Everything that defines results (scoring, exact re-scoring, certification, tie order) is the single-GPU code; the
// merge orders by (distance, lower internal id), the same total order every shard already uses.
#include <algorithm>
#include <chrono>
#include <deque>
#include <functional>
#include <thread>

#include "host.h"

namespace gfi {

namespace {

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

// Outcome of a job that ran on a worker thread (error text and mismatch payload are thread-local there).
struct JobResult {
  int32_t rc = GFI_OK;
  std::string err;
  int64_t exp = 0, act = 0;
  void capture(int32_t code) {
    rc = code;
    if (code != GFI_OK) { err = tl_error; exp = tl_expected; act = tl_actual; }
  }
  int32_t publish() const {
    if (rc != GFI_OK) { tl_error = err; tl_expected = exp; tl_actual = act; }
    return rc;
  }
};

struct Latch {
  std::atomic<int> left;
  explicit Latch(int n) : left(n) {}
  void arrive() { left.fetch_sub(1, std::memory_order_release); }
  void wait() {
    for (uint32_t it = 1; left.load(std::memory_order_acquire) > 0; ++it) {
      if (it < 20000) cpu_relax();
      else if (it < 40000) std::this_thread::yield();
      else std::this_thread::sleep_for(std::chrono::microseconds(50));  // bulk loads, flushes: do not burn a core
    }
  }
};

// One host thread per shard, bound to the shard's GPU for its whole life.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::function<void()>> jobs;
  std::atomic<int> queued{0};
  bool stop = false;

  void start(int device) {
    th = std::thread([this, device] {
      cudaSetDevice(device);
      for (;;) {
        // searches arrive in bursts a few hundred microseconds apart: spin briefly before sleeping on the condition
        for (int i = 0; i < 2000 && queued.load(std::memory_order_acquire) == 0; ++i) cpu_relax();
        std::function<void()> job;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return stop || !jobs.empty(); });
          if (jobs.empty()) return;
          job = std::move(jobs.front());
          jobs.pop_front();
          queued.fetch_sub(1, std::memory_order_relaxed);
        }
        job();
      }
    });
  }
  void post(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(mu);
      jobs.push_back(std::move(f));
      queued.fetch_add(1, std::memory_order_release);
    }
    cv.notify_one();
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_one();
    if (th.joinable()) th.join();
  }
};

struct EvPair { cudaEvent_t a = nullptr, b = nullptr; };

// Gather blocks (and their event sets) a device-resident caller cycles through.  Batch i + kGather writes into the
// block batch i used, so it has to wait for merge i -- and on the root GPU a merge only gets SMs between that GPU's own
// persistent kernels.  With two blocks every shard waited for the root's merge each step (N = 8: 1.38 ms per step
// against 1.25 ms on one GPU); with four the shards run three batches ahead of the slowest merge.
constexpr int kGather = 4;

// Buffers of one in-flight search over all shards (pooled; a device-resident caller keeps one per thread).
struct ShardedCtx {
  cudaStream_t root_stream = nullptr;
  PinBuf h_in, h_out;
  DevBuf gather[kGather];  // root GPU: G per-shard blocks [counts | dist | ids]
  DevBuf result;     // root GPU: [G x 64-byte control blocks | counts | dist | ids] of the merged answer
  DevBuf ks_root;    // per-query k for the merge kernel
  std::vector<SearchCtx*> sub;            // one single-GPU search context per shard
  std::vector<cudaEvent_t> done[kGather];  // per shard: its candidates have landed in gather[b]
  cudaEvent_t merge_done[kGather] = {};
  bool merge_recorded[kGather] = {};
  cudaEvent_t in_ready = nullptr;
  uint64_t seq = 0;
  bool pending_status = false;
  std::vector<cudaStream_t> streams;  // caller streams used by the uncollected device searches (usually one or two)
  std::vector<EvPair> evs;  // option "profile": merge timing on the root stream
  size_t ev_used = 0;
};

}  // namespace

struct ShardSet {
  int G = 0;
  int root = 0;  // device ordinal of the merging GPU (devices[0])
  std::vector<int> devices;
  std::vector<gfi_index*> sub;
  std::vector<std::unique_ptr<Worker>> workers;
  int64_t block = 65536;
  std::mutex pool_mu;
  std::vector<std::unique_ptr<ShardedCtx>> pool;
  std::atomic<int64_t> n_search{0}, n_queries{0}, n_launch{0}, merge_ns{0}, merge_cnt{0};
  int opt_profile = 0;
  int stats_shard = -1;  // option "stats_shard": gfi_get_stats reports this shard alone (-1: the sum over shards)

  int shard_of(uint64_t id) const { return (int)((id / (uint64_t)block) % (uint64_t)G); }

  // runs fn(g) on every selected shard's worker and waits; returns the first failure in shard order
  int32_t on_shards(const std::vector<int>& which, const std::function<int32_t(int)>& fn) {
    std::vector<JobResult> res(which.size());
    Latch latch((int)which.size());
    for (size_t i = 0; i < which.size(); ++i) {
      const int g = which[i];
      JobResult* r = &res[i];
      workers[(size_t)g]->post([&fn, g, r, &latch] {
        try {
          r->capture(fn(g));
        } catch (const std::exception& e) {
          r->rc = GFI_ERR_INDEX;
          r->err = std::string("shard job failed: ") + e.what();
        }
        latch.arrive();
      });
    }
    latch.wait();
    for (auto& r : res)
      if (r.rc != GFI_OK) return r.publish();
    return GFI_OK;
  }
  std::vector<int> all() const {
    std::vector<int> v((size_t)G);
    for (int g = 0; g < G; ++g) v[(size_t)g] = g;
    return v;
  }
};

namespace {

ShardedCtx* acquire_sctx(ShardSet* S) {
  {
    std::lock_guard<std::mutex> g(S->pool_mu);
    if (!S->pool.empty()) {
      ShardedCtx* c = S->pool.back().release();
      S->pool.pop_back();
      return c;
    }
  }
  std::unique_ptr<ShardedCtx> c(new ShardedCtx());
  if (cudaSetDevice(S->root) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithFlags(&c->root_stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  bool ok = cudaEventCreateWithFlags(&c->in_ready, cudaEventDisableTiming) == cudaSuccess;
  for (int b = 0; b < kGather; ++b) ok = ok && cudaEventCreateWithFlags(&c->merge_done[b], cudaEventDisableTiming) == cudaSuccess;
  c->sub.assign((size_t)S->G, nullptr);
  for (int b = 0; b < kGather; ++b) c->done[b].assign((size_t)S->G, nullptr);
  for (int g = 0; g < S->G && ok; ++g) {
    c->sub[(size_t)g] = host_acquire_ctx(S->sub[(size_t)g]);  // (creates its stream on the shard's GPU)
    ok = c->sub[(size_t)g] != nullptr && cudaSetDevice(S->devices[(size_t)g]) == cudaSuccess;
    for (int b = 0; b < kGather && ok; ++b)
      ok = cudaEventCreateWithFlags(&c->done[b][(size_t)g], cudaEventDisableTiming) == cudaSuccess;
  }
  cudaSetDevice(S->root);
  if (!ok) return nullptr;  // (leaks a half-built context: only on CUDA failure at start-up)
  return c.release();
}

void release_sctx(ShardSet* S, ShardedCtx* c) {
  std::lock_guard<std::mutex> g(S->pool_mu);
  S->pool.emplace_back(c);
}

void destroy_sctx(ShardSet* S, ShardedCtx* c) {
  for (int g = 0; g < S->G; ++g) {
    if (c->sub[(size_t)g]) {
      cudaStreamSynchronize(c->sub[(size_t)g]->stream);
      host_release_ctx(S->sub[(size_t)g], c->sub[(size_t)g]);
    }
    for (int b = 0; b < kGather; ++b)
      if (c->done[b][(size_t)g]) cudaEventDestroy(c->done[b][(size_t)g]);
  }
  cudaSetDevice(S->root);
  if (c->root_stream) { cudaStreamSynchronize(c->root_stream); cudaStreamDestroy(c->root_stream); }
  for (int b = 0; b < kGather; ++b) { if (c->merge_done[b]) cudaEventDestroy(c->merge_done[b]); c->gather[b].release(); }
  if (c->in_ready) cudaEventDestroy(c->in_ready);
  for (auto& e : c->evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  c->result.release();
  c->ks_root.release();
  c->h_in.release();
  c->h_out.release();
  delete c;
}

void drain(ShardSet* S, ShardedCtx* c) {
  for (int g = 0; g < S->G; ++g)
    if (c->sub[(size_t)g]) cudaStreamSynchronize(c->sub[(size_t)g]->stream);
  cudaStreamSynchronize(c->root_stream);
  for (cudaStream_t st : c->streams) cudaStreamSynchronize(st);
}

void prof_collect_root(ShardSet* S, ShardedCtx* c) {
  for (size_t i = 0; i < c->ev_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->evs[i].a, c->evs[i].b) == cudaSuccess) {
      S->merge_ns += (int64_t)((double)ms * 1e6);
      S->merge_cnt += 1;
    }
  }
  c->ev_used = 0;
}

bool any_needs_flush(ShardSet* S, bool with_metadata) {
  for (gfi_index* s : S->sub)
    if (host_needs_flush(s, with_metadata)) return true;
  return false;
}

// staged rows reach the GPUs under the handle's UNIQUE lock: no search over any shard is in flight then
int32_t ensure_flushed_all(gfi_index* H, bool with_metadata) {
  ShardSet* S = H->shards;
  {
    std::shared_lock<std::shared_mutex> g(H->mu);
    if (!any_needs_flush(S, with_metadata)) return GFI_OK;
  }
  std::unique_lock<std::shared_mutex> g(H->mu);
  std::vector<int> which;
  for (int s = 0; s < S->G; ++s)
    if (host_needs_flush(S->sub[(size_t)s], with_metadata)) which.push_back(s);
  if (which.empty()) return GFI_OK;
  return S->on_shards(which, [&](int s) { return host_ensure_flushed(S->sub[(size_t)s], with_metadata); });
}

bool pinned_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

struct BlockLayout {
  size_t off_cnt, off_dist, off_ids, bytes;
  BlockLayout(size_t head, int64_t q, uint32_t kout) {
    auto up = [](size_t v) { return (v + 63) & ~(size_t)63; };
    off_cnt = up(head);
    off_dist = off_cnt + up((size_t)q * 4);
    off_ids = off_dist + up((size_t)q * kout * 4);
    bytes = off_ids + up((size_t)q * kout * 8);
  }
};

// Fans one search out to the shards and enqueues the merge on `rs`.  Results land at out_* (root GPU memory).
// `b` selects the gather buffer and event set.
int32_t fan_out(gfi_index* H, ShardedCtx* c, ShardSearch proto, int b, cudaStream_t rs, uint64_t* r_ids, float* r_dist,
                uint32_t* r_counts, int64_t r_kstride, char* r_ctrl, const uint32_t* r_ks) {
  ShardSet* S = H->shards;
  const int G = S->G;
  const int64_t q = proto.q;
  const uint32_t kout = (uint32_t)r_kstride;
  const BlockLayout gl(0, q, kout);
  char* gbase = nullptr;
  if (G > 1) {
    if (c->gather[b].bytes < gl.bytes * (size_t)G && (c->pending_status || c->merge_recorded[b])) drain(S, c);
    CU_TRY(cudaSetDevice(S->root));
    CU_TRY(c->gather[b].ensure(gl.bytes * (size_t)G));
    gbase = c->gather[b].as<char>();
  }
  std::vector<JobResult> res((size_t)G);
  Latch latch(G);
  for (int g = 0; g < G; ++g) {
    ShardSearch s = proto;
    if (G > 1) {
      char* blk = gbase + (size_t)g * gl.bytes;
      s.out_counts = reinterpret_cast<uint32_t*>(blk + gl.off_cnt);
      s.out_dist = reinterpret_cast<float*>(blk + gl.off_dist);
      s.out_ids = reinterpret_cast<uint64_t*>(blk + gl.off_ids);
      s.kstride = kout;
    } else {  // a single shard answers in place
      s.out_counts = r_counts;
      s.out_dist = r_dist;
      s.out_ids = r_ids;
      s.kstride = r_kstride;
    }
    s.out_ctrl = r_ctrl + (size_t)g * 64;
    s.wait_b = (G > 1 && c->merge_recorded[b]) ? c->merge_done[b] : nullptr;
    s.done = c->done[b][(size_t)g];
    gfi_index* shard = S->sub[(size_t)g];
    SearchCtx* sc = c->sub[(size_t)g];
    JobResult* r = &res[(size_t)g];
    S->workers[(size_t)g]->post([shard, sc, s, r, &latch] {
      try {
        r->capture(shard_enqueue(shard, sc, s));
      } catch (const std::exception& e) {
        r->rc = GFI_ERR_INDEX;
        r->err = std::string("shard search failed: ") + e.what();
      }
      latch.arrive();
    });
  }
  latch.wait();
  for (auto& r : res)
    if (r.rc != GFI_OK) {
      drain(S, c);
      return r.publish();
    }
  CU_TRY(cudaSetDevice(S->root));
  for (int g = 0; g < G; ++g) CU_TRY(cudaStreamWaitEvent(rs, c->done[b][(size_t)g], 0));
  if (G > 1) {
    const bool prof = S->opt_profile && c->evs.size() < 4096;
    if (prof) {
      if (c->ev_used == c->evs.size()) {
        EvPair e;
        cudaEventCreate(&e.a);
        cudaEventCreate(&e.b);
        c->evs.push_back(e);
      }
      cudaEventRecord(c->evs[c->ev_used].a, rs);
    }
    CU_TRY(launch_merge(reinterpret_cast<const uint64_t*>(gbase + gl.off_ids),
                        reinterpret_cast<const float*>(gbase + gl.off_dist),
                        reinterpret_cast<const uint32_t*>(gbase + gl.off_cnt), G, q, kout, (int64_t)gl.bytes, r_ks, r_ids,
                        r_dist, r_counts, r_kstride, rs));
    ++S->n_launch;
    if (prof) {
      cudaEventRecord(c->evs[c->ev_used].b, rs);
      ++c->ev_used;
    }
    CU_TRY(cudaEventRecord(c->merge_done[b], rs));
    c->merge_recorded[b] = true;
  }
  return GFI_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------

int32_t sharded_create(gfi_index** out, int32_t metric, int64_t dim, const int32_t* devices, int32_t n_devices,
                       uint32_t flags) {
  if (!out) return fail(GFI_ERR_INDEX, "null out pointer");
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 32) return fail(GFI_ERR_INDEX, "a sharded index takes 1..32 devices");
  if (metric < 0 || metric > 2) return fail(GFI_ERR_INDEX, "unknown metric");
  if (dim < 0) return fail(GFI_ERR_INDEX, "negative dimension");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(GFI_ERR_INDEX, "no CUDA device: libgfi has no CPU fallback");
  for (int g = 0; g < n_devices; ++g)
    if (devices[g] < 0 || devices[g] >= ndev) return fail(GFI_ERR_INDEX, "bad device ordinal");
  std::unique_ptr<gfi_index> H(new gfi_index());
  std::unique_ptr<ShardSet> S(new ShardSet());
  S->G = n_devices;
  S->root = devices[0];
  S->devices.assign(devices, devices + n_devices);
  auto cleanup = [&] {
    for (gfi_index* s : S->sub) gfi_destroy(s);
  };
  // every shard's kernels store their results into the root GPU's memory: peer access towards the root (and from the
  // root towards the shards, for device-resident queries read over NVLink)
  for (int g = 0; g < n_devices; ++g) {
    const int dev = devices[g];
    if (dev == S->root) continue;
    for (int dir = 0; dir < 2; ++dir) {
      const int from = dir ? S->root : dev, to = dir ? dev : S->root;
      int can = 0;
      CU_TRY(cudaDeviceCanAccessPeer(&can, from, to));
      if (!can) return fail(GFI_ERR_INDEX, "sharded index: GPU " + std::to_string(from) + " cannot access GPU " +
                                               std::to_string(to) + " as a peer (NVLink/NVSwitch required)");
      CU_TRY(cudaSetDevice(from));
      cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(GFI_ERR_INDEX, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    }
  }
  for (int g = 0; g < n_devices; ++g) {
    gfi_index* s = nullptr;
    int32_t rc = gfi_create(&s, metric, dim, devices[g], flags);
    if (rc != GFI_OK) { cleanup(); return rc; }
    S->sub.push_back(s);
  }
  for (int g = 0; g < n_devices; ++g) {
    S->workers.emplace_back(new Worker());
    S->workers.back()->start(devices[g]);
  }
  H->metric = metric;
  H->dim = dim;
  H->device = S->root;
  H->flags = flags;
  H->shards = S.release();
  CU_TRY(cudaSetDevice(devices[0]));
  *out = H.release();
  return GFI_OK;
}

void sharded_destroy(gfi_index* H) {
  ShardSet* S = H->shards;
  for (auto& w : S->workers) w->shutdown();
  for (auto& c : S->pool) destroy_sctx(S, c.release());
  S->pool.clear();
  for (gfi_index* s : S->sub) gfi_destroy(s);
  delete S;
  H->shards = nullptr;
  delete H;
}

int64_t sharded_len(const gfi_index* H) {
  int64_t n = 0;
  for (gfi_index* s : H->shards->sub) n += gfi_len(s);
  return n;
}

int64_t sharded_dim(const gfi_index* H) {
  // FlatIndex has no dimension of its own; report the latched one, or that of the first shard that holds rows
  for (gfi_index* s : H->shards->sub)
    if (gfi_len(s) > 0 && gfi_dim(s) > 0) return gfi_dim(s);
  return H->dim;
}

int64_t sharded_row_bytes(const gfi_index* H) {
  int64_t b = 0;
  for (gfi_index* s : H->shards->sub) b += host_row_bytes(s);
  return b;
}

int32_t sharded_add(gfi_index* H, const uint64_t* ids, const float* rows, int64_t n, int64_t dim) {
  ShardSet* S = H->shards;
  if (n == 0) return GFI_OK;
  std::unique_lock<std::shared_mutex> lk(H->mu);
  if (H->dim == 0 && dim > 0) H->dim = dim;
  // maximal runs of consecutive input rows that go to the same shard (ids usually ascend: runs are whole blocks)
  struct Piece { int64_t at, len; };
  std::vector<std::vector<Piece>> pieces((size_t)S->G);
  for (int64_t i = 0; i < n;) {
    const int g = S->shard_of(ids[i]);
    int64_t j = i + 1;
    while (j < n && S->shard_of(ids[j]) == g) ++j;
    pieces[(size_t)g].push_back({i, j - i});
    i = j;
  }
  auto add_to = [&](int g) -> int32_t {
    for (const Piece& p : pieces[(size_t)g]) {
      int32_t rc = gfi_add(S->sub[(size_t)g], ids + p.at, rows ? rows + (size_t)p.at * (size_t)dim : nullptr, p.len, dim);
      if (rc != GFI_OK) return rc;
    }
    return GFI_OK;
  };
  std::vector<int> which;
  for (int g = 0; g < S->G; ++g)
    if (!pieces[(size_t)g].empty()) which.push_back(g);
  // Index::add of a few rows: stage them from the calling thread; bulk loads: all shards stage and flush in parallel
  if ((size_t)n * (size_t)std::max<int64_t>(dim, 1) * 4 < (1u << 20) || which.size() == 1) {
    for (int g : which) {
      int32_t rc = add_to(g);
      if (rc != GFI_OK) return rc;
    }
    return GFI_OK;
  }
  return S->on_shards(which, add_to);
}

int32_t sharded_add_generated(gfi_index* H, uint32_t seed, uint64_t first_row, int64_t n, int32_t kind,
                              uint64_t first_id) {
  ShardSet* S = H->shards;
  std::unique_lock<std::shared_mutex> lk(H->mu);
  if (H->dim == 0) return fail(GFI_ERR_INDEX, "gfi_add_generated needs an index created with a dimension");
  struct Piece { uint64_t off; int64_t len; };
  std::vector<std::vector<Piece>> pieces((size_t)S->G);
  for (uint64_t off = 0; off < (uint64_t)n;) {
    const uint64_t id = first_id + off;
    const uint64_t to_boundary = (uint64_t)S->block - id % (uint64_t)S->block;
    const int64_t len = (int64_t)std::min<uint64_t>(to_boundary, (uint64_t)n - off);
    pieces[(size_t)S->shard_of(id)].push_back({off, len});
    off += (uint64_t)len;
  }
  std::vector<int> which;
  for (int g = 0; g < S->G; ++g)
    if (!pieces[(size_t)g].empty()) which.push_back(g);
  return S->on_shards(which, [&](int g) -> int32_t {
    for (const Piece& p : pieces[(size_t)g]) {
      int32_t rc = gfi_add_generated(S->sub[(size_t)g], seed, first_row + p.off, p.len, kind, first_id + p.off);
      if (rc != GFI_OK) return rc;
    }
    return GFI_OK;
  });
}

int32_t sharded_remove(gfi_index* H, uint64_t id) {
  std::unique_lock<std::shared_mutex> lk(H->mu);
  return gfi_remove(H->shards->sub[(size_t)H->shards->shard_of(id)], id);
}

int32_t sharded_get_vector(gfi_index* H, uint64_t id, float* out, int64_t cap, int64_t* out_dim) {
  std::unique_lock<std::shared_mutex> lk(H->mu);
  return gfi_get_vector(H->shards->sub[(size_t)H->shards->shard_of(id)], id, out, cap, out_dim);
}

int32_t sharded_flush(gfi_index* H, bool compact) {
  ShardSet* S = H->shards;
  std::unique_lock<std::shared_mutex> lk(H->mu);
  return S->on_shards(S->all(), [&](int g) { return compact ? gfi_compact(S->sub[(size_t)g]) : gfi_flush(S->sub[(size_t)g]); });
}

int32_t sharded_reserve(gfi_index* H, int64_t n_rows) {
  ShardSet* S = H->shards;
  std::unique_lock<std::shared_mutex> lk(H->mu);
  if (H->dim == 0) return fail(GFI_ERR_INDEX, "reserve before the dimension is known");
  if (n_rows <= 0) return GFI_OK;
  if (sharded_len(H) == 0) {  // nothing routed yet: contiguous id ranges, one per shard
    const int64_t per = (n_rows + S->G - 1) / S->G;
    S->block = std::max<int64_t>(256, (per + 255) / 256 * 256);
  }
  // rows of ids [0, n_rows) each shard will hold under the routing rule
  std::vector<int64_t> share((size_t)S->G, 0);
  const int64_t nblocks = (n_rows + S->block - 1) / S->block;
  for (int64_t bkt = 0; bkt < nblocks; ++bkt)
    share[(size_t)(bkt % S->G)] += std::min<int64_t>(S->block, n_rows - bkt * S->block);
  std::vector<int> which;
  for (int g = 0; g < S->G; ++g)
    if (share[(size_t)g] > 0 && gfi_dim(S->sub[(size_t)g]) > 0) which.push_back(g);
  return S->on_shards(which, [&](int g) { return gfi_reserve(S->sub[(size_t)g], share[(size_t)g]); });
}

int32_t sharded_set_metadata(gfi_index* H, uint64_t id, int32_t n_fields, const char* const* keys,
                             const char* const* values) {
  std::unique_lock<std::shared_mutex> lk(H->mu);
  return gfi_set_metadata(H->shards->sub[(size_t)H->shards->shard_of(id)], id, n_fields, keys, values);
}

int32_t sharded_set_metadata_column(gfi_index* H, const char* key, const uint64_t* ids, int64_t n,
                                    const char* const* values, int32_t n_values, const uint32_t* codes) {
  ShardSet* S = H->shards;
  std::unique_lock<std::shared_mutex> lk(H->mu);
  struct Piece { int64_t at, len; };
  std::vector<std::vector<Piece>> pieces((size_t)S->G);
  for (int64_t i = 0; i < n;) {
    const int g = S->shard_of(ids[i]);
    int64_t j = i + 1;
    while (j < n && S->shard_of(ids[j]) == g) ++j;
    pieces[(size_t)g].push_back({i, j - i});
    i = j;
  }
  // every shard learns the field and the dictionary (same codes everywhere is not required: filters compile per shard)
  return S->on_shards(S->all(), [&](int g) -> int32_t {
    if (pieces[(size_t)g].empty())
      return gfi_set_metadata_column(S->sub[(size_t)g], key, nullptr, 0, values, n_values, nullptr);
    for (const Piece& p : pieces[(size_t)g]) {
      int32_t rc = gfi_set_metadata_column(S->sub[(size_t)g], key, ids + p.at, p.len, values, n_values, codes + p.at);
      if (rc != GFI_OK) return rc;
    }
    return GFI_OK;
  });
}

static int32_t per_shard_host_search(gfi_index* H, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                                     const uint64_t* mask, int64_t mask_bits, const char* filter_json,
                                     uint64_t* out_ids, float* out_dist, uint32_t* out_counts, int64_t kstride);

int32_t sharded_search(gfi_index* H, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                       const uint64_t* mask, int64_t mask_bits, const char* filter_json, uint64_t* out_ids,
                       float* out_dist, uint32_t* out_counts, int64_t kstride) {
  ShardSet* S = H->shards;
  const int G = S->G;
  int32_t rc;
  std::shared_lock<std::shared_mutex> lk(H->mu, std::defer_lock);
  for (;;) {
    if ((rc = ensure_flushed_all(H, filter_json != nullptr)) != GFI_OK) return rc;
    lk.lock();
    if (!any_needs_flush(S, filter_json != nullptr)) break;  // a writer slipped in between the two locks
    lk.unlock();
  }
  ++S->n_search;
  S->n_queries += q;
  uint32_t kmax = 0;
  for (int64_t i = 0; i < q; ++i) kmax = std::max(kmax, ks[i]);
  if (kmax > 0 && (int64_t)kmax > kstride) return fail(GFI_ERR_INDEX, "kstride smaller than max k");
  if (kmax > 0 && (!out_ids || !out_dist)) return fail(GFI_ERR_INDEX, "null output buffers");
  const uint32_t kout = std::max<uint32_t>(1, std::min<uint32_t>(kmax, (uint32_t)std::min<int64_t>(kstride, 1 << 20)));

  ShardedCtx* c = acquire_sctx(S);
  if (!c) return fail(GFI_ERR_INDEX, "cannot create the CUDA streams of a sharded search");
  struct Releaser { ShardSet* S; ShardedCtx* c; ~Releaser() { release_sctx(S, c); } } rel{S, c};
  c->pending_status = false;
  cudaStream_t rs = c->root_stream;
  CU_TRY(cudaSetDevice(S->root));
  // inputs: ONE pinned block every shard's copy engine reads (queries already in pinned memory are used in place)
  const size_t qbytes = (size_t)q * (size_t)dim * 4;
  CU_TRY(c->h_in.ensure(qbytes + (size_t)q * 4 + 64));
  uint32_t* h_ks = reinterpret_cast<uint32_t*>(c->h_in.as<char>() + ((qbytes + 63) & ~(size_t)63));
  memcpy(h_ks, ks, (size_t)q * 4);
  const float* q_src = queries;
  if (!pinned_host(queries)) {
    memcpy(c->h_in.p, queries, qbytes);
    q_src = c->h_in.as<float>();
  }
  const BlockLayout rl((size_t)G * 64, q, kout);
  CU_TRY(c->result.ensure(rl.bytes));
  CU_TRY(c->h_out.ensure(rl.bytes));
  char* rb = c->result.as<char>();
  if (G > 1) {
    CU_TRY(c->ks_root.ensure((size_t)q * 4));
    CU_TRY(cudaMemcpyAsync(c->ks_root.p, h_ks, (size_t)q * 4, cudaMemcpyHostToDevice, rs));
  }
  ShardSearch proto;
  proto.queries = q_src;
  proto.ks = h_ks;
  proto.on_device = false;
  proto.q = q;
  proto.dim = dim;
  proto.kmax = kmax;
  proto.mask = mask;
  proto.mask_bits = mask_bits;
  proto.filter_json = filter_json;
  if (mask && sharded_row_bytes(H) >= (1ll << 30)) {  // cost-model input, as in the single-GPU path: a sampled estimate
    int64_t pc = 0, seen = 0;
    const size_t full = (size_t)(mask_bits / 64);
    for (size_t w = 0; w < full; w += 64, ++seen) pc += __builtin_popcountll(mask[w]);
    proto.mask_density = seen ? (double)pc / (double)(seen * 64) : 0.0;
  }
  c->seq = 0;
  for (bool& r : c->merge_recorded) r = false;  // host searches end synchronised: nothing to wait for
  rc = fan_out(H, c, proto, 0, rs, reinterpret_cast<uint64_t*>(rb + rl.off_ids), reinterpret_cast<float*>(rb + rl.off_dist),
               reinterpret_cast<uint32_t*>(rb + rl.off_cnt), kout, rb, c->ks_root.as<uint32_t>());
  if (rc != GFI_OK) return rc;
  CU_TRY(cudaMemcpyAsync(c->h_out.p, rb, rl.bytes, cudaMemcpyDeviceToHost, rs));
  cudaError_t se = cudaStreamSynchronize(rs);
  if (se != cudaSuccess) { drain(S, c); return fail(GFI_ERR_INDEX, std::string("sharded search: ") + cudaGetErrorString(se)); }
  c->merge_recorded[0] = false;
  prof_collect_root(S, c);
  const char* hb = c->h_out.as<char>();
  uint32_t flags = 0, unproven = 0;
  for (int g = 0; g < G; ++g) {
    const Ctrl* hc = reinterpret_cast<const Ctrl*>(hb + (size_t)g * 64);
    flags |= hc->flags;
    unproven += hc->unproven;
    shard_account(S->sub[(size_t)g], c->sub[(size_t)g], *hc);
  }
  if ((rc = host_flags_to_status(flags)) != GFI_OK) return rc;
  if (unproven)  // some shard could not prove its scan-path answer on the device: its host path does (paging)
    return per_shard_host_search(H, queries, q, dim, ks, mask, mask_bits, filter_json, out_ids, out_dist, out_counts,
                                 kstride);
  const uint32_t* hcnt = reinterpret_cast<const uint32_t*>(hb + rl.off_cnt);
  const float* hdist = reinterpret_cast<const float*>(hb + rl.off_dist);
  const uint64_t* hids = reinterpret_cast<const uint64_t*>(hb + rl.off_ids);
  for (int64_t i = 0; i < q; ++i) {
    const uint32_t cnt = std::min(hcnt[i], ks[i]);
    out_counts[i] = cnt;
    if (cnt) {
      memcpy(out_ids + i * kstride, hids + (size_t)i * kout, (size_t)cnt * 8);
      memcpy(out_dist + i * kstride, hdist + (size_t)i * kout, (size_t)cnt * 4);
    }
  }
  return GFI_OK;
}

// The slow path: every shard answers with its own exact top-k through its HOST entry point and the G sorted lists
// are merged here.  Used (a) for k beyond the kernels' list capacity (passes of 1016, api.cu search_big_k) and
// (b) when a shard could not prove its scan-path answer on the device: its host path proves it by paging
// (api.cu prove_query).  Rare, kept simple.  The caller holds the handle's shared lock.
static int32_t per_shard_host_search(gfi_index* H, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                                     const uint64_t* mask, int64_t mask_bits, const char* filter_json,
                                     uint64_t* out_ids, float* out_dist, uint32_t* out_counts, int64_t kstride) {
  ShardSet* S = H->shards;
  uint32_t kmax = 1;
  for (int64_t i = 0; i < q; ++i) kmax = std::max(kmax, ks[i]);
  const int G = S->G;
  std::vector<std::vector<uint64_t>> ids((size_t)G);
  std::vector<std::vector<float>> dist((size_t)G);
  std::vector<std::vector<uint32_t>> cnt((size_t)G);
  int32_t rc = S->on_shards(S->all(), [&](int g) -> int32_t {
    ids[(size_t)g].resize((size_t)q * kmax);
    dist[(size_t)g].resize((size_t)q * kmax);
    cnt[(size_t)g].assign((size_t)q, 0);
    if (filter_json)
      return gfi_search_filtered(S->sub[(size_t)g], queries, q, dim, ks, filter_json, ids[(size_t)g].data(),
                                 dist[(size_t)g].data(), cnt[(size_t)g].data(), kmax);
    return gfi_search(S->sub[(size_t)g], queries, q, dim, ks, mask, mask_bits, ids[(size_t)g].data(),
                      dist[(size_t)g].data(), cnt[(size_t)g].data(), kmax);
  });
  if (rc != GFI_OK) return rc;
  std::vector<uint32_t> pos((size_t)G);
  for (int64_t i = 0; i < q; ++i) {
    std::fill(pos.begin(), pos.end(), 0u);
    uint32_t produced = 0;
    while (produced < ks[i]) {
      int best = -1;
      for (int g = 0; g < G; ++g) {
        if (pos[(size_t)g] >= cnt[(size_t)g][(size_t)i]) continue;
        if (best < 0) { best = g; continue; }
        const size_t a = (size_t)i * kmax + pos[(size_t)g], bb = (size_t)i * kmax + pos[(size_t)best];
        const float da = dist[(size_t)g][a], db = dist[(size_t)best][bb];
        if (da < db || (da == db && ids[(size_t)g][a] < ids[(size_t)best][bb])) best = g;
      }
      if (best < 0) break;
      const size_t at = (size_t)i * kmax + pos[(size_t)best]++;
      out_ids[i * kstride + produced] = ids[(size_t)best][at];
      out_dist[i * kstride + produced] = dist[(size_t)best][at];
      ++produced;
    }
    out_counts[i] = produced;
  }
  return GFI_OK;
}

int32_t sharded_search_big_k(gfi_index* H, const float* queries, int64_t q, int64_t dim, const uint32_t* ks,
                             const uint64_t* mask, int64_t mask_bits, const char* filter_json, uint64_t* out_ids,
                             float* out_dist, uint32_t* out_counts, int64_t kstride) {
  ShardSet* S = H->shards;
  if (!queries || !ks || !out_counts || !out_ids || !out_dist) return fail(GFI_ERR_INDEX, "bad arguments");
  for (int64_t i = 0; i < q; ++i)
    if ((int64_t)ks[i] > kstride) return fail(GFI_ERR_INDEX, "kstride smaller than max k");
  int32_t rc;
  if ((rc = ensure_flushed_all(H, filter_json != nullptr)) != GFI_OK) return rc;
  std::shared_lock<std::shared_mutex> lk(H->mu);
  ++S->n_search;
  S->n_queries += q;
  return per_shard_host_search(H, queries, q, dim, ks, mask, mask_bits, filter_json, out_ids, out_dist, out_counts, kstride);
}

namespace {
thread_local ShardedCtx* tl_sctx = nullptr;
thread_local gfi_index* tl_sowner = nullptr;
}  // namespace

int32_t sharded_search_device(gfi_index* H, const float* d_queries, int64_t q, const uint32_t* d_ks, uint32_t kmax,
                              const uint64_t* d_mask, int64_t mask_bits, uint64_t* d_out_ids, float* d_out_dist,
                              uint32_t* d_out_counts, int64_t kstride, void* stream) {
  ShardSet* S = H->shards;
  int32_t rc;
  if ((rc = ensure_flushed_all(H, false)) != GFI_OK) return rc;
  std::shared_lock<std::shared_mutex> lk(H->mu);
  ++S->n_search;
  S->n_queries += q;
  if ((int64_t)kmax > kstride) return fail(GFI_ERR_INDEX, "kstride smaller than max k");
  if (tl_sctx && tl_sowner != H) return fail(GFI_ERR_INDEX, "collect gfi_search_status first");
  if (!tl_sctx) {
    tl_sctx = acquire_sctx(S);
    tl_sowner = H;
    if (!tl_sctx) return fail(GFI_ERR_INDEX, "cannot create the CUDA streams of a sharded search");
    tl_sctx->seq = 0;
    for (bool& r : tl_sctx->merge_recorded) r = false;
    tl_sctx->pending_status = false;
  }
  ShardedCtx* c = tl_sctx;
  CU_TRY(cudaSetDevice(S->root));
  // Searches on different caller streams are independent of each other on the root GPU (each merge is ordered behind
  // its own shards' events, and a gather block is handed over through merge_done), so batches issued round-robin on
  // a few streams overlap: the exchange + merge of one hides behind the main passes of the next ones.  On ONE stream
  // the next batch's inputs are by definition ordered behind the previous batch's merge.
  cudaStream_t rs = stream ? (cudaStream_t)stream : c->root_stream;
  if (std::find(c->streams.begin(), c->streams.end(), rs) == c->streams.end()) c->streams.push_back(rs);
  CU_TRY(c->result.ensure((size_t)S->G * 64));  // the shards' control blocks
  CU_TRY(cudaEventRecord(c->in_ready, rs));     // queries / ks / mask are produced on the caller's stream
  ShardSearch proto;
  proto.queries = d_queries;
  proto.ks = d_ks;
  proto.on_device = true;
  proto.q = q;
  proto.dim = sharded_dim(H);
  proto.kmax = kmax;
  proto.mask = d_mask;
  proto.mask_bits = mask_bits;
  proto.wait_a = c->in_ready;
  proto.keep_flags = c->pending_status;
  const int b = (int)(c->seq++ % kGather);
  rc = fan_out(H, c, proto, b, rs, d_out_ids, d_out_dist, d_out_counts, kstride, c->result.as<char>(), d_ks);
  c->pending_status = true;
  return rc;
}

int32_t sharded_search_status(gfi_index* H) {
  ShardSet* S = H->shards;
  ShardedCtx* c = tl_sctx;
  if (!c || tl_sowner != H) return GFI_OK;
  tl_sctx = nullptr;
  tl_sowner = nullptr;
  struct Releaser { ShardSet* S; ShardedCtx* c; ~Releaser() { release_sctx(S, c); } } rel{S, c};
  if (!c->pending_status) return GFI_OK;
  c->pending_status = false;
  CU_TRY(cudaSetDevice(S->root));
  CU_TRY(c->h_out.ensure((size_t)S->G * 64));
  // every merge waits for its shards, so once the caller streams are idle so are the shards' pipelines
  for (cudaStream_t st : c->streams) CU_TRY(cudaStreamSynchronize(st));
  c->streams.clear();
  for (int g = 0; g < S->G; ++g) CU_TRY(cudaStreamSynchronize(c->sub[(size_t)g]->stream));
  cudaStream_t rs = c->root_stream;
  CU_TRY(cudaMemcpyAsync(c->h_out.p, c->result.p, (size_t)S->G * 64, cudaMemcpyDeviceToHost, rs));
  CU_TRY(cudaStreamSynchronize(rs));
  for (bool& r : c->merge_recorded) r = false;
  prof_collect_root(S, c);
  uint32_t flags = 0, unproven = 0;
  for (int g = 0; g < S->G; ++g) {
    const Ctrl* hc = reinterpret_cast<const Ctrl*>(c->h_out.as<char>() + (size_t)g * 64);
    flags |= hc->flags;
    unproven += hc->unproven;
    shard_account(S->sub[(size_t)g], c->sub[(size_t)g], *hc);
  }
  int32_t rc = host_flags_to_status(flags);
  if (rc != GFI_OK) return rc;
  if (unproven)
    return fail(GFI_ERR_UNPROVEN, "some queries since the last status could not be proven exact on the device "
                                  "(near-duplicate rows around the k-th distance): re-run them through gfi_search");
  return GFI_OK;
}

int32_t sharded_distances(gfi_index* H, const float* queries, int64_t q, int64_t dim, const uint64_t* cand_ids,
                          int64_t m, float* out_dist, uint8_t* out_status) {
  ShardSet* S = H->shards;
  int32_t rc;
  if ((rc = ensure_flushed_all(H, false)) != GFI_OK) return rc;
  std::shared_lock<std::shared_mutex> lk(H->mu);
  const int G = S->G;
  const size_t total = (size_t)q * (size_t)m;
  std::vector<std::vector<float>> dist((size_t)G);
  std::vector<std::vector<uint8_t>> st((size_t)G);
  // every shard scores the whole list (ids it does not hold come back "absent"); the owner's answer is kept
  rc = S->on_shards(S->all(), [&](int g) -> int32_t {
    dist[(size_t)g].resize(total);
    st[(size_t)g].resize(total);
    return gfi_distances(S->sub[(size_t)g], queries, q, dim, cand_ids, m, dist[(size_t)g].data(), st[(size_t)g].data());
  });
  if (rc != GFI_OK) return rc;
  bool invalid = false;
  for (size_t i = 0; i < total; ++i) {
    const size_t g = (size_t)S->shard_of(cand_ids[i]);
    out_dist[i] = dist[g][i];
    if (out_status) out_status[i] = st[g][i];
    invalid = invalid || st[g][i] == 2;
  }
  if (invalid && !out_status)
    return fail(GFI_ERR_INVALID_VECTOR, "Cannot compute cosine distance with zero vector");
  return GFI_OK;
}

int32_t sharded_get_stats(gfi_index* H, gfi_stats* out) {
  ShardSet* S = H->shards;
  for (int g = 0; g < S->G; ++g) {
    if (S->stats_shard >= 0 && g != S->stats_shard) continue;
    gfi_index* s = S->sub[(size_t)g];
    gfi_stats t;
    int32_t rc = gfi_get_stats(s, &t);
    if (rc != GFI_OK) return rc;
    out->n_slots += t.n_slots;
    out->n_live += t.n_live;
    out->scan_queries += t.scan_queries;
    out->tensor_queries += t.tensor_queries;
    out->fallback_queries += t.fallback_queries;
    out->kernel_launches += t.kernel_launches;
    out->bytes_fp32 += t.bytes_fp32;
    out->bytes_fp16 += t.bytes_fp16;
    out->scan_kernel_ns += t.scan_kernel_ns;
    out->scan_kernel_count += t.scan_kernel_count;
    out->tensor_kernel_ns += t.tensor_kernel_ns;
    out->tensor_kernel_count += t.tensor_kernel_count;
    out->paged_queries += t.paged_queries;
  }
  out->searches = S->n_search;
  out->queries = S->n_queries;
  out->kernel_launches += S->n_launch;
  out->coalesced_batches = H->n_co_batches;
  out->coalesced_requests = H->n_co_requests;
  out->shards = S->G;
  out->merge_ns = S->merge_ns;
  out->merge_count = S->merge_cnt;
  return GFI_OK;
}

int32_t sharded_set_option(gfi_index* H, const char* name, int64_t value) {
  ShardSet* S = H->shards;
  std::unique_lock<std::shared_mutex> lk(H->mu);
  const std::string n(name);
  if (n == "coalesce") { H->opt_coalesce = (int)value; return GFI_OK; }
  if (n == "shard_block") {
    if (sharded_len(H) != 0) return fail(GFI_ERR_INDEX, "shard_block can only change while the index is empty");
    if (value < 1) return fail(GFI_ERR_INDEX, "shard_block must be positive");
    S->block = value;
    return GFI_OK;
  }
  if (n == "stats_shard") {
    if (value < -1 || value >= S->G) return fail(GFI_ERR_INDEX, "stats_shard out of range");
    S->stats_shard = (int)value;
    return GFI_OK;
  }
  if (n == "profile") S->opt_profile = (int)value;
  for (gfi_index* s : S->sub) {
    int32_t rc = gfi_set_option(s, name, value);
    if (rc != GFI_OK) return rc;
  }
  return GFI_OK;
}

}  // namespace gfi
