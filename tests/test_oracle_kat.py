"""CPU: pins the oracle (oracle/flat_oracle.c) against every known-answer test the reference
holds for the hot path, against the independent numpy-float32 restatement, and against the
frozen golden cases."""
import os

import numpy as np
import pytest

import oracle
from oracle import pyref
from helpers import KATS, HERE


@pytest.mark.parametrize("kat", KATS["distance"], ids=lambda k: k["src"])
def test_reference_distance_kats(kat):
    a, b = kat["a"], kat["b"]
    if kat["metric"] == "rawdot":  # dot_product() itself (src/distance.rs:67-73); the metric negates it
        got = -float(oracle.distance("dot", a, b))
    else:
        got = float(oracle.distance(kat["metric"], a, b))
    assert abs(got - kat["expect"]) <= kat["tol"] * max(1.0, abs(kat["expect"]))


def test_reference_dimension_mismatch_kat():
    k = KATS["dimension_mismatch"]
    with pytest.raises(oracle.OracleError) as e:
        oracle.distance("euclidean", k["a"], k["b"])
    assert e.value.code == 1


def test_reference_norm_kat():
    for k in KATS["norm"]:
        assert abs(oracle.norm(k["v"]) - k["expect"]) <= k["tol"]


def test_reference_flat_index_basic_kat():
    k = KATS["flat_index_basic"]
    ids = sorted(int(i) for i in k["rows"])
    rows = [k["rows"][str(i)] for i in ids]
    got_ids, got_d = oracle.flat_search(k["metric"], rows, k["query"], k["k"], ids=ids)
    assert len(got_ids) == k["expect_len"]
    assert got_ids[0] == k["expect_first_id"]
    assert got_d[0] < k["expect_first_dist_below"]


def test_reference_metrics_self_match_kat():
    k = KATS["integration_metrics_self_match"]
    for m in k["metrics"]:
        ids, _ = oracle.flat_search(m, [k["rows"]["v1"]], k["query"], k["k"])
        assert list(ids) == [0]


def test_cosine_zero_vector_is_invalid_vector():
    with pytest.raises(oracle.OracleError) as e:
        oracle.flat_search("cosine", [[1, 0], [0, 0]], [1, 1], 1)
    assert e.value.code == 2
    with pytest.raises(oracle.OracleError) as e:
        oracle.flat_search("cosine", [[1, 0]], [0, 0], 1)
    assert e.value.code == 2


def test_semantics_k_and_empty():
    rows = [[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]]
    ids, d = oracle.flat_search("euclidean", rows, [0.1, 0], 10)  # k > n => n results
    assert list(ids) == [0, 1, 2]
    ids, d = oracle.flat_search("euclidean", rows, [0.1, 0], 0)   # k = 0 => empty
    assert len(ids) == 0
    ids, d = oracle.flat_search("euclidean", np.zeros((0, 2), np.float32), [0.1, 0], 3)
    assert len(ids) == 0


def test_tie_break_lower_id_and_negative_zero():
    rows = [[1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 0.0]]
    ids, _ = oracle.flat_search("euclidean", rows, [1, 0], 3, ids=[7, 3, 9, 5])
    assert list(ids) == [3, 5, 7]
    # -dot = -0.0 must tie with +0.0 (partial_cmp treats them equal) and fall back to the id
    rows = [[0.0, 1.0], [0.0, -1.0]]
    ids, d = oracle.flat_search("dot", rows, [1.0, 0.0], 2, ids=[2, 1])
    assert list(ids) == [1, 2]


def test_nan_is_reported():
    with pytest.raises(oracle.OracleError) as e:
        oracle.flat_search("euclidean", [[np.nan, 0], [1, 0]], [0, 0], 1)
    assert e.value.code == 3


def test_post_filter_semantics():
    # storage.rs:269: fetch_k = min(max(3k, k), len); rows beyond the top-3k are never seen
    rows = np.arange(20, dtype=np.float32).reshape(-1, 1)
    matches = np.zeros(20, dtype=bool)
    matches[[1, 7, 15]] = True
    ids, _ = oracle.search_post_filter("euclidean", rows, [0.0], 2, matches)
    assert list(ids) == [1]  # fetch_k = 6 => only rows 0..5 are candidates
    ids, _ = oracle.flat_search("euclidean", rows, [0.0], 2, eligible=matches)
    assert list(ids) == [1, 7]  # exact pre-filter


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot"])
def test_c_oracle_is_bit_identical_to_numpy_float32_restatement(metric):
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((40, 19)).astype(np.float32)
    q = rng.standard_normal(19).astype(np.float32)
    exp = pyref.flat_search(metric, rows, q, 12)
    ids, d = oracle.flat_search(metric, rows, q, 12)
    assert [int(i) for i in ids] == [i for _, i in exp]
    assert np.array_equal(d, np.array([x for x, _ in exp], dtype=np.float32))


def test_generator_matches_python_restatement():
    for kind in (0, 1):
        rows = oracle.gen_rows(9, (1 << 33) + 5, 3, 7, kind)
        for r in range(3):
            for c in range(7):
                assert rows[r, c] == pyref.gen_elem(9, (1 << 33) + 5 + r, c, kind)
    u = oracle.gen_rows(3, 0, 2000, 32, 0)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    g = oracle.gen_rows(3, 0, 2000, 32, 1)
    assert abs(g.mean()) < 0.02 and abs(g.std() - 1.0) < 0.02


def test_batch_threads_equal_sequential():
    rows = oracle.gen_rows(21, 0, 500, 24, 1)
    qs = oracle.gen_rows(22, 0, 9, 24, 1)
    a = oracle.search_batch("cosine", rows, qs, [3, 1, 4, 1, 5, 9, 2, 6, 5])
    b = oracle.search_batch("cosine", rows, qs, [3, 1, 4, 1, 5, 9, 2, 6, 5], threads=4)
    for (ia, da), (ib, db) in zip(a, b):
        assert np.array_equal(ia, ib) and np.array_equal(da, db)


def test_frozen_golden_cases():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    z = np.load(os.path.join(HERE, "golden", "oracle_cases.npz"))
    for name, metric, n, d, kind, seed, q, k in make_golden.CASES:
        rows = oracle.gen_rows(seed, 0, n, d, kind)
        queries = oracle.gen_rows(seed + 1000, 0, q, d, kind)
        assert np.array_equal(rows[:2, :4], z[name + "/rows_head"])
        res = oracle.search_batch(metric, rows, queries, k)
        assert np.array_equal(np.stack([r[0] for r in res]), z[name + "/ids"])
        assert np.array_equal(np.stack([r[1] for r in res]), z[name + "/dist"])


def test_generated_rows_parallel_mode_equals_the_array_mode():
    """oracle.search_generated (rows produced on the fly, rows-parallel, per-thread top-k + merge) is the full-size
    checker; it must equal the array mode, which the reference's own KATs pin above."""
    for metric, kind in (("euclidean", 0), ("cosine", 1), ("dot", 1)):
        n, d, q = 6000, 40, 5
        rows = oracle.gen_rows(11, 100, n, d, kind)
        qs = oracle.gen_rows(12, 0, q, d, kind)
        elig = (np.arange(n) * 7 % 5) != 0
        ks = [10, 1, 0, 100, 7]
        a = oracle.search_batch(metric, rows, qs, ks, ids=np.arange(n, dtype=np.uint64) + 5000, eligible=elig, threads=2)
        for threads in (1, 3):
            b = oracle.search_generated(metric, 11, 100, n, d, kind, qs, ks, eligible=elig, first_id=5000, threads=threads)
            for (ai, ad), (bi, bd) in zip(a, b):
                assert np.array_equal(ai, bi) and np.array_equal(ad, bd)
    # duplicates: ties resolve by lower id whatever the partition
    rows = np.tile(oracle.gen_rows(13, 0, 1, 16, 0), (50, 1))
    a = oracle.search_batch("euclidean", rows, rows[:1], 7)[0]
    assert list(a[0]) == list(range(7))
