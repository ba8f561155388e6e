"""Exact ground truth for recall measurements of approximate indexes.

The reference checks its HNSW index against FlatIndex results with `recall_at_k`
(tests/recall_test.rs:18-26): |ids found by the approximate index that are in the exact top-k| / k.
At the sizes BASELINE.json names (100M rows) the exact side is the GPU flat search of this package;
`export_ground_truth` runs it in batches and writes the exact ids to a small binary file that a
Rust-side recall test can read back, `recall_at_k` is the reference's formula.

File layout (little endian): magic b"GFGT", u32 version = 1, u32 q, u32 k, then q*k u64 ids (row major,
ascending distance, ties by lower id; 0xFFFFFFFFFFFFFFFF pads queries that have fewer than k results).
"""
import struct

import numpy as np

MAGIC = b"GFGT"
PAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def export_ground_truth(index, queries, k, path, batch=4096):
    """Exact top-k ids of every query through GpuFlatIndex.search_arrays, written to `path`.
    Returns the [q, k] uint64 id array (PAD where a query has fewer than k results)."""
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    q = queries.shape[0]
    out = np.full((q, k), PAD, dtype=np.uint64)
    for lo in range(0, q, batch):
        hi = min(q, lo + batch)
        ids, _, cnt = index.search_arrays(queries[lo:hi], k)
        for i in range(hi - lo):
            out[lo + i, :cnt[i]] = ids[i, :cnt[i]]
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<III", 1, q, k))
        f.write(out.astype("<u8").tobytes())
    return out


def load_ground_truth(path):
    with open(path, "rb") as f:
        head = f.read(16)
        if head[:4] != MAGIC:
            raise ValueError("not a ground-truth file")
        version, q, k = struct.unpack("<III", head[4:])
        if version != 1:
            raise ValueError(f"unsupported ground-truth version {version}")
        ids = np.frombuffer(f.read(q * k * 8), dtype="<u8").reshape(q, k)
    return ids


def recall_at_k(truth_ids, found_ids):
    """tests/recall_test.rs:18-26: found ids that are in the ground-truth set / size of the ground truth."""
    truth = {int(x) for x in np.asarray(truth_ids, dtype=np.uint64).ravel() if x != PAD}
    if not truth:
        return 1.0
    found = sum(1 for x in np.asarray(found_ids, dtype=np.uint64).ravel() if int(x) in truth)
    return found / len(truth)


def mean_recall(truth, found):
    """Average recall over queries (the reference's test_recall averages per-query recall_at_k)."""
    return float(np.mean([recall_at_k(t, f) for t, f in zip(truth, found)]))
