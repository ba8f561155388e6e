"""GPU: exact (query, row id) pair distances (gfi_distances -- the batched form of the HNSW candidate evaluation,
src/hnsw/graph.rs:221-232) against the oracle's DistanceMetric::distance, bit for bit, and the ground-truth export
for recall measurements."""
import numpy as np
import pytest

import oracle
import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200 import DistanceMetric as DM, groundtruth as gt

pytestmark = pytest.mark.gpu
METRICS = {"euclidean": DM.Euclidean, "cosine": DM.Cosine, "dot": DM.DotProduct}


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot"])
@pytest.mark.parametrize("d", [3, 128, 1500])
def test_pair_distances_equal_reference_distance(metric, d):
    n, q, m = 700, 5, 9
    rows = oracle.gen_rows(81 + d, 0, n, d, 1)
    queries = oracle.gen_rows(82 + d, 0, q, d, 1)
    ids = np.arange(n, dtype=np.uint64) * 3 + 7          # sparse, non-contiguous ids
    idx = gfi.GpuFlatIndex(METRICS[metric])
    idx.add_batch(ids, rows)
    rng = np.random.default_rng(5)
    pick = rng.integers(0, n, size=(q, m))
    cand = ids[pick]
    cand[0, 0] = 1          # never inserted
    cand[1, 2] = ids[-1] + 1
    idx.remove(int(ids[pick[2, 3]]))   # tombstoned
    dist, status = idx.distances(queries, cand)
    for i in range(q):
        for j in range(m):
            absent = (i, j) in ((0, 0), (1, 2)) or cand[i, j] == ids[pick[2, 3]]
            if absent:
                assert status[i, j] == 1 and np.isinf(dist[i, j])
            else:
                assert status[i, j] == 0
                exp = oracle.distance(metric, queries[i], rows[pick[i, j]])
                assert dist[i, j].tobytes() == np.float32(exp + 0.0).tobytes(), (i, j, dist[i, j], exp)


def test_pair_distances_error_semantics():
    rows = np.array([[1, 0, 0], [0, 0, 0], [0, 2, 0]], dtype=np.float32)
    idx = gfi.GpuFlatIndex(DM.Cosine)
    idx.add_batch(np.arange(3, dtype=np.uint64), rows)
    q = np.array([[1, 0, 0]], dtype=np.float32)
    dist, status = idx.distances(q, np.array([[0, 1, 2]], dtype=np.uint64))
    assert list(status[0]) == [0, 2, 0]                     # zero-norm row: Err(InvalidVector) of distance.rs:51-55
    assert dist[0, 0] == 0.0 and np.isinf(dist[0, 1]) and dist[0, 2] == 1.0
    with pytest.raises(gfi.DimensionMismatch):
        idx.distances(np.zeros((1, 4), np.float32), np.array([[0]], dtype=np.uint64))
    # without a status array the first InvalidVector fails the call, as `?` would
    L, out = gfi.lib(), np.zeros(3, np.float32)
    ids = np.array([0, 1, 2], dtype=np.uint64)
    rc = L.gfi_distances(idx._h, q.ctypes.data, 1, 3, ids.ctypes.data, 3, out.ctypes.data, None)
    assert rc == 2
    empty = gfi.GpuFlatIndex(DM.Euclidean)
    d2, s2 = empty.distances(q, np.array([[0, 5]], dtype=np.uint64))
    assert np.all(s2 == 1) and np.all(np.isinf(d2))


def test_ground_truth_export_and_recall(tmp_path):
    n, d, q, k = 20000, 64, 300, 10
    rows = oracle.gen_rows(95, 0, n, d, 0)
    queries = oracle.gen_rows(96, 0, q, d, 0)
    idx = gfi.GpuFlatIndex(DM.Euclidean)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    path = str(tmp_path / "truth.gfgt")
    truth = gt.export_ground_truth(idx, queries, k, path, batch=128)
    assert np.array_equal(gt.load_ground_truth(path), truth)
    exp = oracle.search_batch("euclidean", rows, queries[:40], k, threads=8)
    for i, (eids, _) in enumerate(exp):
        assert [int(x) for x in truth[i]] == [int(x) for x in eids]
    # an "approximate index" that misses the 3 farthest of every top-10: recall 0.7 by the reference's formula
    approx = truth.copy()
    approx[:, 7:] = np.uint64(n + 1)
    assert abs(gt.mean_recall(truth, approx) - 0.7) < 1e-12
