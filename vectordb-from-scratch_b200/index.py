"""GpuFlatIndex -- ctypes mirror of the reference's `Index` trait implemented by libgfi.

Method names, argument meaning and error behaviour follow `trait Index`
(reference src/index.rs:11-35) and `FlatIndex` (src/flat_index.rs:12-74); the two additive
batch/mask methods are the defaulted trait extensions SURVEY.md section 8(b) B5 describes.
"""
import ctypes
import enum

import numpy as np

from . import native
from .errors import DimensionMismatch, InvalidVector, IndexError_, NaNDistance, Unproven


class DistanceMetric(enum.IntEnum):
    """reference src/distance.rs:9-16"""
    Euclidean = 0
    Cosine = 1
    DotProduct = 2


def _raise(L, rc):
    msg = (L.gfi_last_error() or b"").decode()
    if rc == 1:
        e, a = ctypes.c_int64(), ctypes.c_int64()
        L.gfi_last_mismatch(ctypes.byref(e), ctypes.byref(a))
        raise DimensionMismatch(e.value, a.value)
    if rc == 2:
        raise InvalidVector(msg)
    if rc == 4:
        raise NaNDistance(msg)
    if rc == 5:
        raise Unproven(msg)
    raise IndexError_(msg)


def pack_mask(bits):
    """bool[n] indexed by internal id -> u64 words (bit id%64 of word id//64)."""
    bits = np.asarray(bits, dtype=bool)
    n = bits.shape[0]
    pad = (-n) % 64
    b = np.concatenate([bits, np.zeros(pad, dtype=bool)]).reshape(-1, 64)
    w = (b.astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
    return np.ascontiguousarray(w, dtype=np.uint64), n


class GpuFlatIndex:
    """Drop-in for FlatIndex behind `trait Index`.  All compute runs in libgfi.so on a B200."""

    def __init__(self, metric=DistanceMetric.Euclidean, dim=0, device=0, flags=0, devices=None):
        """`devices`: a list of GPU ordinals -> ONE index sharded row-wise over them inside libgfi
        (gfi_create_sharded; devices[0] merges).  Every method below works on such a handle unchanged."""
        self._L = native.lib()
        self._h = ctypes.c_void_p()
        if devices is not None:
            devs = (ctypes.c_int32 * len(devices))(*[int(x) for x in devices])
            rc = self._L.gfi_create_sharded(ctypes.byref(self._h), int(metric), int(dim), devs, len(devices), int(flags))
        else:
            rc = self._L.gfi_create(ctypes.byref(self._h), int(metric), int(dim), int(device), int(flags))
        if rc:
            self._h = None
            _raise(self._L, rc)

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None):
            self._L.gfi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            _raise(self._L, rc)

    # -- trait Index --
    def add(self, id, vector):
        """Index::add(&mut self, id, vector) (src/index.rs:13)."""
        v = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        ids = np.array([id], dtype=np.uint64)
        self._chk(self._L.gfi_add(self._h, ids.ctypes.data, v.ctypes.data if v.size else None, 1, v.size))

    def add_batch(self, ids, rows):
        """Bulk form of Index::add (one C-ABI call for n rows)."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2:
            raise ValueError("rows must be n x d")
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        assert ids.shape[0] == rows.shape[0]
        self._chk(self._L.gfi_add(self._h, ids.ctypes.data, rows.ctypes.data, rows.shape[0], rows.shape[1]))

    def add_generated(self, seed, first_row, n, kind, first_id):
        self._chk(self._L.gfi_add_generated(self._h, seed, first_row, n, kind, first_id))

    def add_from_file(self, path, first_id=0):
        """Bulk-load the reference's flat vector file (src/persistence/mmap.rs); returns the row count."""
        n = ctypes.c_int64()
        self._chk(self._L.gfi_add_from_file(self._h, str(path).encode(), int(first_id), ctypes.byref(n)))
        return n.value

    def remove(self, id):
        """Index::remove (src/index.rs:16): idempotent."""
        self._chk(self._L.gfi_remove(self._h, int(id)))

    def get_vector(self, id):
        """Index::get_vector (src/index.rs:23): the stored row or None."""
        d = max(int(self._L.gfi_dim(self._h)), 1)
        out = np.empty(d, dtype=np.float32)
        od = ctypes.c_int64()
        rc = self._L.gfi_get_vector(self._h, int(id), out.ctypes.data, d, ctypes.byref(od))
        if rc:
            return None
        return out[:od.value]

    def metric(self):
        return DistanceMetric(self._L.gfi_metric(self._h))

    def len(self):
        return int(self._L.gfi_len(self._h))

    __len__ = len

    def is_empty(self):
        return self.len() == 0

    def dim(self):
        return int(self._L.gfi_dim(self._h))

    def search(self, query, k):
        """Index::search (src/index.rs:20): list of (id, distance), ascending distance."""
        return self.search_batch([(query, k)])[0]

    # -- additive, defaulted extensions --
    def search_batch(self, queries, mask=None):
        """queries: list of (vector, k).  One C-ABI call; per-query k as in
        VectorStore::search_batch (src/storage.rs:302-310)."""
        if len(queries) == 0:
            return []
        dims = {len(np.asarray(q).reshape(-1)) for q, _ in queries}
        if len(dims) != 1:
            # the reference checks each query separately; the first mismatching one fails the batch
            out = []
            for qv, k in queries:
                out.append(self.search_batch([(qv, k)], mask=mask)[0])
            return out
        qs = np.ascontiguousarray(np.stack([np.asarray(q, dtype=np.float32).reshape(-1) for q, _ in queries]))
        ks = np.array([int(k) for _, k in queries], dtype=np.uint32)
        ids, dist, cnt = self.search_arrays(qs, ks, mask=mask)
        return [[(int(ids[i, j]), float(dist[i, j])) for j in range(cnt[i])] for i in range(len(queries))]

    def search_arrays(self, queries, ks, mask=None):
        """Array form: queries [q,d] f32, ks [q] u32 (or int) -> (ids [q,kmax], dist [q,kmax], counts [q])."""
        qs = np.ascontiguousarray(queries, dtype=np.float32)
        q, d = qs.shape
        ks = np.ascontiguousarray(np.broadcast_to(np.asarray(ks, dtype=np.uint32), (q,)))
        kmax = max(int(ks.max()) if q else 0, 1)
        out_ids = np.zeros((q, kmax), dtype=np.uint64)
        out_dist = np.zeros((q, kmax), dtype=np.float32)
        cnt = np.zeros(q, dtype=np.uint32)
        mptr, mbits = None, 0
        if mask is not None:
            # a bool array indexed by internal id, or a pre-packed (u64 words, nbits) pair
            words, mbits = mask if isinstance(mask, tuple) else pack_mask(mask)
            mptr = words.ctypes.data
        self._chk(self._L.gfi_search(self._h, qs.ctypes.data if qs.size else None, q, d, ks.ctypes.data, mptr,
                                     mbits, out_ids.ctypes.data, out_dist.ctypes.data, cnt.ctypes.data, kmax))
        return out_ids, out_dist, cnt

    def distances(self, queries, cand_ids):
        """Exact DistanceMetric::distance (src/distance.rs:20-33) of explicit (query, row id) pairs: queries [q,d],
        cand_ids [q,m] -> (dist [q,m] f32, status [q,m] u8: 0 ok, 1 id absent, 2 zero-norm cosine operand).  The
        batched form of the candidate evaluation in the reference's HNSW (src/hnsw/graph.rs:221-232)."""
        qs = np.ascontiguousarray(queries, dtype=np.float32)
        ids = np.ascontiguousarray(cand_ids, dtype=np.uint64)
        q, d = qs.shape
        assert ids.ndim == 2 and ids.shape[0] == q
        m = ids.shape[1]
        out = np.full((q, m), np.inf, dtype=np.float32)
        status = np.zeros((q, m), dtype=np.uint8)
        self._chk(self._L.gfi_distances(self._h, qs.ctypes.data if qs.size else None, q, d,
                                        ids.ctypes.data if ids.size else None, m, out.ctypes.data, status.ctypes.data))
        return out, status

    def set_metadata(self, id, fields):
        """Replace the metadata of `id` (dict str -> str): VectorStore::insert_with_metadata's
        `self.metadata.insert(internal_id, metadata)` (src/storage.rs:169), kept as columns in HBM."""
        items = list(fields.items())
        n = len(items)
        keys = (ctypes.c_char_p * max(n, 1))(*[k.encode() for k, _ in items])
        vals = (ctypes.c_char_p * max(n, 1))(*[v.encode() for _, v in items])
        self._chk(self._L.gfi_set_metadata(self._h, int(id), n, keys, vals))

    def set_metadata_column(self, key, ids, values, codes):
        """Bulk form: field `key` of ids[i] = values[codes[i]] (code 0xFFFFFFFF: absent); other fields untouched."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        assert ids.shape == codes.shape
        vals = (ctypes.c_char_p * max(len(values), 1))(*[v.encode() for v in values])
        self._chk(self._L.gfi_set_metadata_column(self._h, key.encode(), ids.ctypes.data if ids.size else None,
                                                  ids.size, vals, len(values), codes.ctypes.data if codes.size else None))

    def search_filtered(self, queries, ks, filter_json):
        """FlatIndex::search over the rows matching a MetadataFilter (the reference's JSON form,
        src/storage.rs:44-58), with the filter evaluated on the GPU.  Array form like search_arrays."""
        import json as _json
        if not isinstance(filter_json, str):
            filter_json = _json.dumps(filter_json)
        qs = np.ascontiguousarray(queries, dtype=np.float32)
        q, d = qs.shape
        ks = np.ascontiguousarray(np.broadcast_to(np.asarray(ks, dtype=np.uint32), (q,)))
        kmax = max(int(ks.max()) if q else 0, 1)
        out_ids = np.zeros((q, kmax), dtype=np.uint64)
        out_dist = np.zeros((q, kmax), dtype=np.float32)
        cnt = np.zeros(q, dtype=np.uint32)
        self._chk(self._L.gfi_search_filtered(self._h, qs.ctypes.data, q, d, ks.ctypes.data, filter_json.encode(),
                                              out_ids.ctypes.data, out_dist.ctypes.data, cnt.ctypes.data, kmax))
        return out_ids, out_dist, cnt

    def search_masked(self, query, k, eligible):
        """FlatIndex::search over the rows whose internal id is eligible (filter push-down)."""
        return self.search_batch([(query, k)], mask=eligible)[0]

    # -- device-pointer forms (multi-GPU plumbing, HBM-resident benchmark) --
    def search_device(self, d_queries, q, d_ks, kmax, d_out_ids, d_out_dist, d_out_counts, kstride, stream=0,
                      d_mask=0, mask_bits=0):
        self._chk(self._L.gfi_search_device(self._h, d_queries, q, d_ks, kmax, d_mask or None, mask_bits,
                                            d_out_ids, d_out_dist, d_out_counts, kstride, stream or None))

    def search_status(self):
        self._chk(self._L.gfi_search_status(self._h))

    def merge_topk_device(self, d_ids, d_dist, d_counts, G, q, kstride, d_ks, d_out_ids, d_out_dist, d_out_counts,
                          out_kstride, stream=0, shard_stride_bytes=0):
        if shard_stride_bytes:
            self._chk(self._L.gfi_merge_topk_device_strided(d_ids, d_dist, d_counts, G, q, kstride,
                                                            shard_stride_bytes, d_ks, d_out_ids, d_out_dist,
                                                            d_out_counts, out_kstride, stream or None))
        else:
            self._chk(self._L.gfi_merge_topk_device(d_ids, d_dist, d_counts, G, q, kstride, d_ks, d_out_ids,
                                                    d_out_dist, d_out_counts, out_kstride, stream or None))

    # -- maintenance / introspection --
    def flush(self):
        self._chk(self._L.gfi_flush(self._h))

    def reserve(self, n_rows):
        self._chk(self._L.gfi_reserve(self._h, int(n_rows)))

    def compact(self):
        self._chk(self._L.gfi_compact(self._h))

    def set_option(self, name, value):
        self._chk(self._L.gfi_set_option(self._h, name.encode(), int(value)))

    def debug_tensor_scores(self, queries):
        """Raw approximate scores of the tcgen05 pass, [q, n_slots] (test hook)."""
        qs = np.ascontiguousarray(queries, dtype=np.float32)
        self.flush()
        n = self.stats()["n_slots"]
        npad = (n + 255) // 256 * 256
        out = np.empty((qs.shape[0], npad), dtype=np.float32)
        self._chk(self._L.gfi_debug_tensor_scores(self._h, qs.ctypes.data, qs.shape[0], out.ctypes.data, npad))
        return out[:, :n]

    def stats(self):
        s = native.GfiStats()
        self._chk(self._L.gfi_get_stats(self._h, ctypes.byref(s)))
        return {n: getattr(s, n) for n, _ in native.GfiStats._fields_}
