/*
 * gfi.h -- C ABI of libgfi.so, the B200-native GpuFlatIndex.
 *
 * Drop-in boundary: this is what a Rust `gpu-flat-index-sys` crate binds
 *  so that `GpuFlatIndex: Index` can replace `FlatIndex`
 * behind the reference's unchanged `trait Index` (reference src/index.rs:11-35).
 * Every entry point names the reference interface it replaces.  No torch or C++
 * types cross this boundary: plain pointers, sizes and integer status codes.
 *
 * Conventions
 *   - every function returns an int32 status (GFI_OK = 0) unless stated;
 *     nothing throws or aborts across the boundary; gfi_last_error() returns a
 *     thread-local message for the last non-zero status on the calling thread.
 *   - the caller owns all host buffers; the library owns device memory behind
 *     the opaque handle.
 *   - threading mirrors the reference server (src/server/mod.rs:13-16):
 *     gfi_search* are re-entrant and may run concurrently from any OS thread
 *     (they are `&self` under RwLock::read); gfi_add/remove/flush/compact are
 *     exclusive (`&mut self` under RwLock::write).
 *   - results are ordered by (distance ascending, then lower internal id); the
 *     reference's own tie order is unspecified (HashMap iteration order under a
 *     stable sort, src/flat_index.rs:53-62), so this is one of its legal outputs.
 *   - distances are bit-identical to the reference's arithmetic (sequential f32
 *     sums, separately rounded multiply/add; src/distance.rs:37-73).
 */
#ifndef GFI_H_
#define GFI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfi_index gfi_index;

/* status codes <-> reference src/error.rs:10-31 */
#define GFI_OK 0
#define GFI_ERR_DIMENSION_MISMATCH 1 /* VectorDbError::DimensionMismatch{expected, actual}: see gfi_last_mismatch */
#define GFI_ERR_INVALID_VECTOR 2     /* VectorDbError::InvalidVector (cosine with a zero-norm query or row) */
#define GFI_ERR_INDEX 3              /* VectorDbError::IndexError(String): CUDA failure, bad argument, ... */
#define GFI_ERR_NAN 4                /* a distance is NaN: the reference panics (flat_index.rs:62); we report */
#define GFI_ERR_UNPROVEN 5           /* gfi_search_status only: a device-resident search met more near-ties of the
                                        k-th distance than the kernels' candidate lists hold and could not PROVE its
                                        answer exact; re-run those queries through gfi_search, which proves them by
                                        paging (host searches never return this code) */

/* DistanceMetric, reference src/distance.rs:9-16 */
#define GFI_METRIC_EUCLIDEAN 0
#define GFI_METRIC_COSINE 1
#define GFI_METRIC_DOT 2

/* gfi_create flags */
#define GFI_FLAG_NO_TENSOR 1u /* never use the tcgen05 batched path (scan path only, no fp16 shadow copy) */

/* data kinds for gfi_add_generated (identical generator in oracle/flat_oracle.c) */
#define GFI_GEN_UNIFORM 0
#define GFI_GEN_NORMAL 1

/*
 * FlatIndex::new(metric), reference src/flat_index.rs:19-24.
 * dim = 0 latches the dimension at the first gfi_add (FlatIndex has no dimension
 * of its own); device = CUDA ordinal.
 */
int32_t gfi_create(gfi_index **out, int32_t metric, int64_t dim, int32_t device, uint32_t flags);
int32_t gfi_destroy(gfi_index *h);

/*
 * The same index sharded row-wise over several GPUs of one box, in ONE process -- what the reference server needs:
 * it owns a single `RwLock<VectorStore<I>>` (src/server/mod.rs:13-16, src/main.rs:152-198), so `GpuFlatIndex: Index`
 * must reach all GPUs from behind one handle.  devices[0] is the root GPU (it merges).  Every entry point of this
 * header accepts the handle; differences:
 *   - rows are routed by internal id in blocks: shard = (id / block) % n_devices, block = 65536 ids, or
 *     ceil(n_rows / n_devices) after gfi_reserve(h, n_rows) on an empty index (contiguous id ranges);
 *   - a search copies the query block once into pinned memory, every shard's worker thread enqueues the ordinary
 *     single-GPU pipeline on its own GPU, and each shard's finalize kernel stores its k candidates straight into
 *     the root GPU's gather block over NVLink (peer stores: no collective, no staging copy); the root merges the
 *     per-shard lists by (distance, id) with one kernel and returns one result block;
 *   - gfi_search_device takes queries/results in the ROOT GPU's memory (the shards read them in place over NVLink)
 *     and cycles through four gather blocks: batches issued round-robin on a few streams run back to back on every
 *     shard while the root merges earlier ones;
 *   - gfi_debug_tensor_scores is single-GPU only.
 * A device may be listed more than once (several shards on one GPU: used by the single-GPU tests).
 * All listed GPUs must be able to access the root GPU's memory (NVLink / NVSwitch peers).
 */
int32_t gfi_create_sharded(gfi_index **out, int32_t metric, int64_t dim, const int32_t *devices, int32_t n_devices,
                           uint32_t flags);

/*
 * Index::add, reference src/index.rs:13 / src/flat_index.rs:38-41, for n rows at
 * once (n = 1 from Index::add; large n for bulk load).  An existing id is
 * overwritten.  Rows are staged in pinned host memory and reach the GPU at the
 * next gfi_flush / gfi_search.  Like FlatIndex::add this never fails on a
 * dimension: a row whose `dim` differs from the index dimension is recorded and
 * makes every later search return GFI_ERR_DIMENSION_MISMATCH, exactly as the
 * per-pair check in DistanceMetric::distance does (src/distance.rs:21-26).
 */
int32_t gfi_add(gfi_index *h, const uint64_t *ids, const float *rows, int64_t n, int64_t dim);

/* Bench/bulk helper: rows first_row..first_row+n of the synthetic generator are created
 * directly in HBM (ids first_id..), so 100M-row shards never cross PCIe. */
int32_t gfi_add_generated(gfi_index *h, uint32_t seed, uint64_t first_row, int64_t n, int32_t kind,
                          uint64_t first_id);

/* Bulk ingest of the reference's flat vector file (MmapVectorStorage, src/persistence/mmap.rs:13-15,
 * 161-172: little-endian [dimension u32][count u32] then count x dimension f32).  Row i gets internal id
 * first_id + i.  Reads in 32 MB chunks straight into the staged-add path (no per-row call). */
int32_t gfi_add_from_file(gfi_index *h, const char *path, uint64_t first_id, int64_t *out_rows);

/* Index::remove, src/index.rs:16 / src/flat_index.rs:43-46: idempotent (tombstone). */
int32_t gfi_remove(gfi_index *h, uint64_t id);

/* Index::len / Index::metric, src/index.rs:26-29.  gfi_dim: 0 while unlatched. */
int64_t gfi_len(const gfi_index *h);
int32_t gfi_metric(const gfi_index *h);
int64_t gfi_dim(const gfi_index *h);

/* Index::get_vector, src/index.rs:23: copies the stored row to `out` (cap floats). Returns
 * GFI_OK and *out_dim = row dimension, or GFI_ERR_INDEX if the id is absent (Option::None).
 * (A Rust wrapper that must lend `&Vector` keeps its own host mirror; see INTEGRATION.md.) */
int32_t gfi_get_vector(gfi_index *h, uint64_t id, float *out, int64_t cap, int64_t *out_dim);

/* Push staged rows to the GPU now (searches do it implicitly). */
int32_t gfi_flush(gfi_index *h);
/* Pre-size device storage for n rows (avoids regrowth during bulk load). */
int32_t gfi_reserve(gfi_index *h, int64_t n_rows);
/* Drop tombstoned slots and restore slot order == id order. */
int32_t gfi_compact(gfi_index *h);

/*
 * Index::search (src/index.rs:20, src/flat_index.rs:52-65) for a batch of q queries with
 * per-query k (VectorStore::search_batch, src/storage.rs:302-310; q = 1 is Index::search).
 *   queries  : q x dim, row-major host floats
 *   ks       : q per-query k (k = 0 => no results)
 *   mask     : NULL, or eligibility bits indexed by INTERNAL ID (bit id%64 of word id/64,
 *              mask_bits bits long; ids >= mask_bits are ineligible).  A masked search is
 *              FlatIndex::search over the eligible rows (filter push-down; the reference's
 *              own post-filter, src/storage.rs:249-290, needs no mask: the caller over-fetches).
 *   out_ids / out_dist : q x kstride, out_counts[i] = min(ks[i], eligible rows)
 * The first failing query fails the batch (collect::<Result<_>>(), storage.rs:306-309).
 * Any k is accepted, with or without a mask (k beyond the kernels' list capacity of 1016 is served in exact
 * passes over the rows not returned yet).  Thread-safe and re-entrant; calls that arrive while another plain search
 * (no mask, q <= 256) is running are combined into one batched search (group commit, option "coalesce"), each
 * caller receiving exactly the outcome -- results or error -- of its own call.
 */
int32_t gfi_search(gfi_index *h, const float *queries, int64_t q, int64_t dim, const uint32_t *ks,
                   const uint64_t *mask, int64_t mask_bits, uint64_t *out_ids, float *out_dist,
                   uint32_t *out_counts, int64_t kstride);

/*
 * Device-side metadata filters (the "next" row N2 of SURVEY.md section 8f).  The reference keeps
 * Metadata{HashMap<String,String>} per internal id in VectorStore (src/storage.rs:19-42,90) and walks it on
 * the host per candidate (storage.rs:272-285).  gfi_set_metadata mirrors VectorStore::insert_with_metadata's
 * `self.metadata.insert(internal_id, metadata)` (storage.rs:169): the fields of `id` are REPLACED by the
 * given key/value pairs and kept as dictionary-encoded columns in HBM.  gfi_search_filtered takes the
 * reference's own JSON form of MetadataFilter (serde tag "op": eq / ne / exists / and / or,
 * storage.rs:44-58), evaluates it on the GPU into an eligibility bitmask and runs FlatIndex::search over the
 * matching rows (exact pre-filter; truth table of storage.rs:62-70).  Any k, as for gfi_search: beyond the kernels'
 * list capacity the filter is evaluated once over the host mirror of the columns and the passes run over the
 * matching rows not returned yet.
 */
int32_t gfi_set_metadata(gfi_index *h, uint64_t id, int32_t n_fields, const char *const *keys,
                         const char *const *values);

/*
 * Bulk form for ONE field over many rows: field `key` of ids[i] becomes values[codes[i]] (codes[i] = UINT32_MAX:
 * the field is absent); the other fields of those ids are left as they are.  This loads the reference's
 * `metadata: HashMap<usize, Metadata>` (src/storage.rs:90) at ingest speed instead of one call per row.
 */
int32_t gfi_set_metadata_column(gfi_index *h, const char *key, const uint64_t *ids, int64_t n,
                                const char *const *values, int32_t n_values, const uint32_t *codes);
int32_t gfi_search_filtered(gfi_index *h, const float *queries, int64_t q, int64_t dim, const uint32_t *ks,
                            const char *filter_json, uint64_t *out_ids, float *out_dist, uint32_t *out_counts,
                            int64_t kstride);

/*
 * Same search with every buffer already in device memory (HBM) and asynchronous on
 * `stream` (a cudaStream_t passed as void*; NULL = the index's own stream): no host
 * synchronisation.  d_ks/d_mask/d_* are device pointers.  Call gfi_search_status to
 * collect the status of the last device search issued on this handle by this thread
 * (synchronises the stream).  Used by multi-GPU callers that all-gather the per-shard
 * results and by the HBM-resident benchmark.
 */
int32_t gfi_search_device(gfi_index *h, const float *d_queries, int64_t q, const uint32_t *d_ks,
                          uint32_t kmax, const uint64_t *d_mask, int64_t mask_bits,
                          uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_counts,
                          int64_t kstride, void *stream);
int32_t gfi_search_status(gfi_index *h);

/*
 * Merge G per-shard result lists (each sorted by (distance, id)) into the global top-k
 * per query, on device, asynchronously on `stream`.  Inputs are laid out [G][q][kstride]
 * (the layout an all-gather of per-shard gfi_search_device outputs produces).
 */
int32_t gfi_merge_topk_device(const uint64_t *d_ids, const float *d_dist, const uint32_t *d_counts,
                              int32_t G, int64_t q, int64_t kstride, const uint32_t *d_ks,
                              uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_counts,
                              int64_t out_kstride, void *stream);

/*
 * Same merge over PACKED per-shard blocks: shard g's ids / distances / counts start at the given base pointers
 * plus g * shard_stride_bytes (a multiple of 8).  A shard that lets gfi_search_device write its three outputs
 * into one contiguous block [ids | dist | counts] needs a single all-gather of that block per search.
 */
int32_t gfi_merge_topk_device_strided(const uint64_t *d_ids, const float *d_dist, const uint32_t *d_counts,
                                      int32_t G, int64_t q, int64_t kstride, int64_t shard_stride_bytes,
                                      const uint32_t *d_ks, uint64_t *d_out_ids, float *d_out_dist,
                                      uint32_t *d_out_counts, int64_t out_kstride, void *stream);

/*
 * Exact distances of explicit (query, row id) pairs: DistanceMetric::distance (src/distance.rs:20-33), bit-identical
 * to the reference, for m candidate ids per query (cand_ids [q][m], out_dist [q][m]).  This is what HNSW
 * construction and search evaluate for their candidate lists (src/hnsw/graph.rs:221-232 `metric.distance(&node_vec,
 * &n.vector)`, search_layer), batched onto the GPU copy of the rows.  out_status [q][m] (may be NULL): 0 ok; 1 the
 * id is not in the index (the reference's `nodes.get(nid)` = None; distance = +inf); 2 cosine with a zero-norm
 * operand (the reference's Err(InvalidVector); distance = +inf).  With out_status == NULL a status-2 pair makes the
 * call return GFI_ERR_INVALID_VECTOR.  NaN distances are returned as NaN (distance() does not panic; sort_by does).
 */
int32_t gfi_distances(gfi_index *h, const float *queries, int64_t q, int64_t dim, const uint64_t *cand_ids,
                      int64_t m, float *out_dist, uint8_t *out_status);

/* DimensionMismatch payload of the last GFI_ERR_DIMENSION_MISMATCH on this thread. */
void gfi_last_mismatch(int64_t *expected, int64_t *actual);
/* Thread-local message for the last error on this thread ("" if none). */
const char *gfi_last_error(void);

/* Counters for the benchmark harness (not part of the reference interface). */
typedef struct gfi_stats {
  int64_t n_slots;          /* device rows incl. tombstones */
  int64_t n_live;           /* live rows */
  int64_t searches;         /* search calls */
  int64_t queries;          /* queries served */
  int64_t scan_queries;     /* queries answered by the streaming scan kernel */
  int64_t tensor_queries;   /* queries answered by the tcgen05 kernel */
  int64_t fallback_queries; /* tensor-path queries re-run on the scan kernel (not certified) */
  int64_t kernel_launches;  /* CUDA kernels launched by this handle */
  int64_t bytes_fp32;       /* device bytes of the fp32 rows */
  int64_t bytes_fp16;       /* device bytes of the fp16 shadow rows */
  /* with option "profile"=1: CUDA-event time of the dominant kernel of each search, summed */
  int64_t scan_kernel_ns, scan_kernel_count;     /* K1 flat_scan_topk launches */
  int64_t tensor_kernel_ns, tensor_kernel_count; /* K2 flat_gemm_topk main-pass launches */
  /* group commit of concurrent plain gfi_search calls (option "coalesce", default on) */
  int64_t coalesced_batches, coalesced_requests; /* combined searches run, calls they served */
  /* sharded handles (gfi_create_sharded); shards = 1 otherwise.  With "profile"=1: CUDA-event time on the root GPU
   * from the moment the last shard's candidates have arrived to the end of the merge kernel, summed. */
  int64_t shards;
  int64_t merge_ns, merge_count;
  /* host searches whose fp32-scan answer the device could not certify and that were proven exact by paging */
  int64_t paged_queries;
} gfi_stats;
int32_t gfi_get_stats(gfi_index *h, gfi_stats *out);

/* Test hook: the raw approximate scores of the tcgen05 candidate pass for every (query, slot) pair,
 * out[q][out_stride] with out_stride >= n_slots rounded up to 256 (small inputs only).  L2: ||x||^2 - 2 q.x
 * (without ||q||^2); cosine: -q.x/||x||; dot: -q.x.  Ineligible slots read +inf. */
int32_t gfi_debug_tensor_scores(gfi_index *h, const float *queries, int64_t q, float *out, int64_t out_stride);

/* Tuning knobs for experiments (name/value); returns GFI_ERR_INDEX for unknown names. */
int32_t gfi_set_option(gfi_index *h, const char *name, int64_t value);

/* Library/ABI version (major*100 + minor). */
int32_t gfi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GFI_H_ */
