"""CPU: the recall formula of the reference (tests/recall_test.rs:18-26) and the ground-truth file written for
HNSW recall measurements, with the exact side injected (the oracle here)."""
import numpy as np

import oracle
from vectordb_from_scratch_b200 import groundtruth as gt


class OracleIndex:
    """search_arrays with the oracle behind it: stands in for GpuFlatIndex on a GPU-less box."""

    def __init__(self, metric, rows):
        self.metric, self.rows = metric, rows

    def search_arrays(self, queries, k):
        res = oracle.search_batch(self.metric, self.rows, queries, k)
        q = len(res)
        ids = np.zeros((q, k), np.uint64)
        dist = np.zeros((q, k), np.float32)
        cnt = np.zeros(q, np.uint32)
        for i, (a, b) in enumerate(res):
            ids[i, :len(a)], dist[i, :len(b)], cnt[i] = a, b, len(a)
        return ids, dist, cnt


def test_recall_formula_matches_reference_definition():
    # recall_test.rs:18-26: |found ∩ truth| / |truth|
    assert gt.recall_at_k([1, 2, 3, 4], [4, 3, 9, 8]) == 0.5
    assert gt.recall_at_k([1, 2, 3, 4], [1, 2, 3, 4]) == 1.0
    assert gt.recall_at_k([1, 2, 3, 4], []) == 0.0
    assert gt.recall_at_k([5, 6, gt.PAD, gt.PAD], [6, 7]) == 0.5   # padded ground truth: 2 true neighbours
    assert gt.mean_recall([[1, 2], [3, 4]], [[1, 2], [9, 9]]) == 0.5


def test_ground_truth_file_round_trip(tmp_path):
    rows = oracle.gen_rows(91, 0, 300, 16, 0)
    queries = oracle.gen_rows(92, 0, 7, 16, 0)
    idx = OracleIndex("euclidean", rows)
    path = tmp_path / "truth.gfgt"
    ids = gt.export_ground_truth(idx, queries, 10, str(path), batch=3)   # ragged last batch
    back = gt.load_ground_truth(str(path))
    assert back.shape == (7, 10) and np.array_equal(back, ids)
    exp = oracle.search_batch("euclidean", rows, queries, 10)
    for i, (eids, _) in enumerate(exp):
        assert [int(x) for x in back[i]] == [int(x) for x in eids]
    # k larger than the index: padded rows
    small = OracleIndex("euclidean", rows[:4])
    ids2 = gt.export_ground_truth(small, queries[:2], 6, str(path))
    assert np.all(ids2[:, 4:] == gt.PAD) and gt.recall_at_k(ids2[0], ids2[0, :4]) == 1.0
