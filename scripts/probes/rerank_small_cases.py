"""Small-case probe (also usable under compute-sanitizer where that is open): a few small tensor-path searches (select with the rerank cut + bulk rerank, odd row
lengths so the clamped last-chunk loads are exercised), checked against the oracle.
usage: compute-sanitizer --tool memcheck python scripts/probes/rerank_small_cases.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
import vectordb_from_scratch_b200 as gfi

M = {"euclidean": gfi.DistanceMetric.Euclidean, "cosine": gfi.DistanceMetric.Cosine, "dot": gfi.DistanceMetric.DotProduct}
for metric, n, d, q, k in (("euclidean", 3000, 100, 32, 5), ("cosine", 2500, 37, 20, 10), ("dot", 2000, 260, 17, 3), ("cosine", 2500, 37, 20, 40)):
    rows = oracle.gen_rows(11, 0, n, d, 1)
    queries = oracle.gen_rows(12, 0, q, d, 1)
    idx = gfi.GpuFlatIndex(M[metric])
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    idx.set_option("tensor_min_rows", 256)
    for cut in (1, 0):
        idx.set_option("rerank_cut", cut)
        ids, dist, cnt = idx.search_arrays(queries, k)
        exp = oracle.search_batch(metric, rows, queries, k, threads=4)
        for i, (eids, ed) in enumerate(exp):
            assert cnt[i] == len(eids) and np.array_equal(ids[i, :cnt[i]], eids) and np.array_equal(dist[i, :cnt[i]], ed), (metric, cut, i)
    st = idx.stats()  # (small k-heavy cases may be routed to the scan path: parity is the point, the route is printed)
    print("ok", metric, n, d, q, k, {x: st[x] for x in ("tensor_queries", "scan_queries", "fallback_queries")}, flush=True)
