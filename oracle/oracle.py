"""ctypes binding of oracle/liboracle.so (flat_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
METRICS = {"euclidean": 0, "cosine": 1, "dot": 2, "dotproduct": 2}
_lib = None


class OracleError(Exception):
    """code: 1 DimensionMismatch, 2 InvalidVector (cosine zero norm), 3 NaN (reference panics)."""

    def __init__(self, code):
        super().__init__({1: "DimensionMismatch", 2: "InvalidVector", 3: "NaN"}.get(code, str(code)))
        self.code = code


def build(force=False):
    src = os.path.join(_HERE, "flat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        f32p, u64p, i64p = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint64),
                            ctypes.POINTER(ctypes.c_int64))
        lib.orc_norm.restype = ctypes.c_float
        lib.orc_norm.argtypes = [f32p, ctypes.c_int64]
        lib.orc_distance.restype = ctypes.c_int
        lib.orc_distance.argtypes = [ctypes.c_int, f32p, ctypes.c_int64, f32p, ctypes.c_int64, f32p]
        lib.orc_flat_search.restype = ctypes.c_int
        lib.orc_flat_search.argtypes = [ctypes.c_int, f32p, u64p, ctypes.c_int64, ctypes.c_int64, u64p,
                                        f32p, ctypes.c_int64, ctypes.c_int64, u64p, f32p, i64p]
        lib.orc_search_post_filter.restype = ctypes.c_int
        lib.orc_search_post_filter.argtypes = lib.orc_flat_search.argtypes
        lib.orc_search_batch.restype = ctypes.c_int
        lib.orc_search_batch.argtypes = [ctypes.c_int, f32p, u64p, ctypes.c_int64, ctypes.c_int64, u64p,
                                         f32p, ctypes.c_int64, ctypes.c_int64, i64p, ctypes.c_int64,
                                         u64p, f32p, i64p, ctypes.c_int]
        lib.orc_search_generated.restype = ctypes.c_int
        lib.orc_search_generated.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                             ctypes.c_int64, ctypes.c_int64, ctypes.c_int, u64p, f32p,
                                             ctypes.c_int64, i64p, ctypes.c_int64, u64p, f32p, i64p, ctypes.c_int]
        lib.orc_gen_rows.restype = None
        lib.orc_gen_rows.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int, f32p]
        lib.orc_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def _metric(m):
    return METRICS[m.lower()] if isinstance(m, str) else int(m)


def pack_mask(bits):
    """bool[n] -> u64 words, bit r of word r//64 (the layout gfi_search's mask uses)."""
    bits = np.asarray(bits, dtype=bool)
    n = bits.shape[0]
    pad = (-n) % 64
    b = np.concatenate([bits, np.zeros(pad, dtype=bool)]).reshape(-1, 64)
    w = (b.astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
    return np.ascontiguousarray(w, dtype=np.uint64)


def norm(v):
    v = _f32(v)
    return float(np.float32(_load().orc_norm(_p(v, ctypes.c_float), v.size)))


def distance(metric, a, b):
    a, b = _f32(a), _f32(b)
    out = ctypes.c_float()
    rc = _load().orc_distance(_metric(metric), _p(a, ctypes.c_float), a.size, _p(b, ctypes.c_float),
                              b.size, ctypes.byref(out))
    if rc:
        raise OracleError(rc)
    return np.float32(out.value)


def _rows(rows):
    rows = _f32(rows)
    if rows.ndim == 1:
        rows = rows.reshape(0, 0) if rows.size == 0 else rows.reshape(1, -1)
    return rows


def flat_search(metric, rows, query, k, ids=None, eligible=None, _fn="orc_flat_search"):
    """FlatIndex::search over `rows` (n x d).  Returns (ids u64[c], dist f32[c])."""
    rows, query = _rows(rows), _f32(query)
    n, d = rows.shape
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
    el = None if eligible is None else pack_mask(eligible)
    kk = max(int(k), 0)
    out_ids = np.zeros(max(kk, 1), dtype=np.uint64)
    out_d = np.zeros(max(kk, 1), dtype=np.float32)
    cnt = ctypes.c_int64()
    rc = getattr(_load(), _fn)(_metric(metric), _p(rows, ctypes.c_float), _p(ids_a, ctypes.c_uint64), n, d,
                               _p(el, ctypes.c_uint64), _p(query, ctypes.c_float), query.size, kk,
                               _p(out_ids, ctypes.c_uint64), _p(out_d, ctypes.c_float), ctypes.byref(cnt))
    if rc:
        raise OracleError(rc)
    return out_ids[:cnt.value].copy(), out_d[:cnt.value].copy()


def search_post_filter(metric, rows, query, k, matches, ids=None):
    """VectorStore::search_with_filter (post-filter, fetch_k = min(max(3k,k), n))."""
    return flat_search(metric, rows, query, k, ids=ids, eligible=matches, _fn="orc_search_post_filter")


def search_batch(metric, rows, queries, ks, ids=None, eligible=None, threads=1):
    """VectorStore::search_batch: returns list of (ids, dist) per query."""
    rows, queries = _rows(rows), _rows(queries)
    n, d = rows.shape
    q, dq = queries.shape
    ks = np.ascontiguousarray(np.broadcast_to(np.asarray(ks, dtype=np.int64), (q,)))
    kmax = max(int(ks.max()) if q else 0, 1)
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
    el = None if eligible is None else pack_mask(eligible)
    out_ids = np.zeros((q, kmax), dtype=np.uint64)
    out_d = np.zeros((q, kmax), dtype=np.float32)
    cnt = np.zeros(q, dtype=np.int64)
    rc = _load().orc_search_batch(_metric(metric), _p(rows, ctypes.c_float), _p(ids_a, ctypes.c_uint64), n, d,
                                  _p(el, ctypes.c_uint64), _p(queries, ctypes.c_float), q, dq,
                                  _p(ks, ctypes.c_int64), kmax, _p(out_ids, ctypes.c_uint64),
                                  _p(out_d, ctypes.c_float), _p(cnt, ctypes.c_int64), int(threads))
    if rc:
        raise OracleError(rc)
    return [(out_ids[i, :cnt[i]].copy(), out_d[i, :cnt[i]].copy()) for i in range(q)]


def search_generated(metric, seed, first_row, n, d, kind, queries, ks, eligible=None, first_id=0, threads=None):
    """FlatIndex::search over rows gen_rows(seed, first_row, n, d, kind) produced on the fly (ids first_id + r),
    rows-parallel: the full-size checker.  Returns list of (ids, dist) per query."""
    queries = _rows(queries)
    q = queries.shape[0]
    assert queries.shape[1] == d
    ks = np.ascontiguousarray(np.broadcast_to(np.asarray(ks, dtype=np.int64), (q,)))
    kmax = max(int(ks.max()) if q else 0, 1)
    el = None if eligible is None else pack_mask(eligible)
    out_ids = np.zeros((q, kmax), dtype=np.uint64)
    out_d = np.zeros((q, kmax), dtype=np.float32)
    cnt = np.zeros(q, dtype=np.int64)
    if threads is None:
        import os
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    rc = _load().orc_search_generated(_metric(metric), seed, first_row, first_id, n, d, kind,
                                      _p(el, ctypes.c_uint64), _p(queries, ctypes.c_float), q,
                                      _p(ks, ctypes.c_int64), kmax, _p(out_ids, ctypes.c_uint64),
                                      _p(out_d, ctypes.c_float), _p(cnt, ctypes.c_int64), int(threads))
    if rc:
        raise OracleError(rc)
    return [(out_ids[i, :cnt[i]].copy(), out_d[i, :cnt[i]].copy()) for i in range(q)]


def gen_rows(seed, first_row, n, d, kind):
    """Counter-based synthetic rows; kind 0 = U[0,1), 1 = normal-like (unit variance)."""
    out = np.empty((n, d), dtype=np.float32)
    _load().orc_gen_rows(seed, first_row, n, d, kind, _p(out, ctypes.c_float))
    return out


def max_threads():
    return int(_load().orc_max_threads())
