"""Round-2 regression tests for the host runtime of libgfi (ADVICE r1 / VERDICT r1 items), through the C ABI."""
import numpy as np
import pytest

import oracle
import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200 import DistanceMetric as DM
from helpers import assert_topk_matches

pytestmark = pytest.mark.gpu


def test_rows_added_after_a_metadata_sync_read_as_no_metadata():
    """gfi_add / gfi_add_generated after the last metadata sync grew the slot array but not the
    device columns; a filtered search then read past them.  Rows without metadata must behave like the reference's
    `metadata.get(id) = None` rows: eq/exists false, ne true (storage.rs:62-70)."""
    n0, n1, d, k = 3000, 5000, 24, 8
    rows = oracle.gen_rows(41, 0, n0 + n1, d, 1)
    idx = gfi.GpuFlatIndex(DM.Euclidean)
    idx.add_batch(np.arange(n0, dtype=np.uint64), rows[:n0])
    for i in range(n0):
        idx.set_metadata(i, {"color": "red" if i % 2 else "blue"})
    q = oracle.gen_rows(42, 0, 2, d, 1)
    ids, dist, cnt = idx.search_filtered(q, k, {"op": "eq", "field": "color", "value": "red"})  # syncs the columns
    assert all(int(i) % 2 == 1 for i in ids[0, :cnt[0]])
    # plain adds and generated rows: no gfi_set_metadata call for them
    idx.add_batch(np.arange(n0, n0 + 1000, dtype=np.uint64), rows[n0:n0 + 1000])
    idx.add_generated(41, n0 + 1000, n1 - 1000, 1, n0 + 1000)
    has_md = np.arange(n0 + n1) < n0
    red = has_md & (np.arange(n0 + n1) % 2 == 1)
    cases = [({"op": "eq", "field": "color", "value": "red"}, red),
             ({"op": "ne", "field": "color", "value": "red"}, ~red),
             ({"op": "exists", "field": "color"}, has_md),
             ({"op": "and", "filters": [{"op": "ne", "field": "color", "value": "blue"},
                                        {"op": "ne", "field": "color", "value": "red"}]}, ~has_md)]
    for flt, elig in cases:
        ids, dist, cnt = idx.search_filtered(q, k, flt)
        exp = oracle.search_batch("euclidean", rows, q, k, eligible=elig, threads=4)
        for i, (eids, ed) in enumerate(exp):
            assert cnt[i] == len(eids), (flt, cnt[i], len(eids))
            assert_topk_matches(ids[i, :cnt[i]], dist[i, :cnt[i]], eids, ed, ctx=str(flt))


def test_an_emptied_index_takes_a_new_dimension():
    """latch dim 3, remove every row, add dim-5 rows, search with a dim-5 query: FlatIndex::search
    returns those rows (it has no dimension of its own, flat_index.rs:38-41,52-65)."""
    idx = gfi.GpuFlatIndex(DM.Euclidean)
    idx.add(0, [1.0, 2.0, 3.0])
    idx.add(1, [0.0, 0.0, 1.0])
    assert idx.search([1.0, 2.0, 3.0], 1)[0] == (0, 0.0)
    idx.remove(0)
    idx.remove(1)
    assert idx.len() == 0
    rows = oracle.gen_rows(7, 0, 300, 5, 0)
    idx.add_batch(np.arange(10, 310, dtype=np.uint64), rows)
    assert idx.dim() == 5 and idx.len() == 300
    got = idx.search(rows[17], 3)
    exp_ids, exp_d = oracle.search_batch("euclidean", rows, rows[17:18], 3, ids=np.arange(10, 310, dtype=np.uint64))[0]
    assert [g[0] for g in got] == list(exp_ids) and got[0] == (27, 0.0)
    assert np.array_equal(np.float32([g[1] for g in got]), exp_d)
    assert np.array_equal(idx.get_vector(27), rows[17]) and idx.get_vector(0) is None
    # rows of a second dimension recorded while others were live, which then disappear: loud, not an empty Ok
    idx2 = gfi.GpuFlatIndex(DM.Euclidean)
    idx2.add(0, [1.0, 2.0])
    idx2.add(1, [1.0, 2.0, 3.0])
    idx2.remove(0)
    with pytest.raises(gfi.IndexError_):
        idx2.search([1.0, 2.0, 3.0], 1)
    idx2.add(1, [1.0, 2.0, 3.0])  # re-added now that nothing else is stored: the index takes its dimension
    assert idx2.search([1.0, 2.0, 3.0], 1)[0] == (1, 0.0) and idx2.len() == 1


def test_len_counts_a_staged_overwrite_once():
    """FlatIndex::len after add(id) of an existing id stays the same (HashMap::insert)."""
    idx = gfi.GpuFlatIndex(DM.DotProduct)
    idx.add(5, [1.0, 0.0])
    idx.add(6, [0.0, 1.0])
    idx.flush()
    idx.add(5, [2.0, 0.0])      # staged overwrite of a flushed id
    idx.add(9, [3.0, 3.0])      # staged new id
    assert idx.len() == 3
    assert idx.search([1.0, 0.0], 1)[0] == (9, -3.0)
    assert idx.len() == 3


def test_device_search_flags_are_sticky_until_collected():
    """several gfi_search_device calls before one gfi_search_status: an error raised by ANY of them is
    reported (the control block used to be cleared by every call, so only the last search's flags survived)."""
    torch = pytest.importorskip("torch")
    n, d, k = 20000, 64, 5
    idx = gfi.GpuFlatIndex(DM.Cosine, dim=d)
    idx.add_generated(3, 0, n, 1, 0)
    dev = torch.device("cuda", 0)
    good = torch.from_numpy(oracle.gen_rows(4, 0, 2, d, 1)).to(dev)
    bad = good.clone()
    bad[1] = 0.0  # zero-norm query: InvalidVector
    ks = torch.full((2,), k, dtype=torch.int32, device=dev)
    oi = torch.zeros((2, k), dtype=torch.int64, device=dev)
    od = torch.zeros((2, k), dtype=torch.float32, device=dev)
    oc = torch.zeros((2,), dtype=torch.int32, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    idx.search_device(bad.data_ptr(), 2, ks.data_ptr(), k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), k,
                      stream=s1.cuda_stream)
    # a second search on ANOTHER stream is ordered behind the first (shared workspace) and must not clear its error
    idx.search_device(good.data_ptr(), 2, ks.data_ptr(), k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), k,
                      stream=s2.cuda_stream)
    with pytest.raises(gfi.InvalidVector):
        idx.search_status()
    idx.search_device(good.data_ptr(), 2, ks.data_ptr(), k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), k,
                      stream=s2.cuda_stream)
    idx.search_status()  # collected and clean again
    rows = oracle.gen_rows(3, 0, n, d, 1)
    exp = oracle.search_batch("cosine", rows, good.cpu().numpy(), k, threads=4)
    for i, (eids, ed) in enumerate(exp):
        assert_topk_matches(oi[i].cpu().numpy().astype(np.uint64), od[i].cpu().numpy(), eids, ed)


# ---------------------------------------------------------------- short-K row-tile-stationary tcgen05 kernel
SHORT_K_CASES = [  # metric, n, d, kind, q, k, mask?
    ("euclidean", 60000, 128, 0, 520, 10, False),   # C5 scaled down: 5 query tiles (last one partial), 2 k-chunks
    ("cosine", 40000, 128, 1, 512, 10, False),      # raw epilogue
    ("dot", 33000, 100, 1, 640, 20, False),         # d not a multiple of 16: partial second chunk (TMA zero fill)
    ("euclidean", 50000, 64, 0, 512, 10, False),    # one k-chunk
    ("cosine", 30011, 72, 0, 600, 5, True),         # mask -> coefficient epilogue for cosine; n not a tile multiple
    ("euclidean", 300, 96, 1, 512, 10, False),      # fewer row tiles than CTAs (most CTAs idle)
]


@pytest.mark.parametrize("case", SHORT_K_CASES, ids=lambda c: "%s_n%d_d%d_q%d_k%d" % (c[0], c[1], c[2], c[4], c[5]))
def test_short_k_row_stationary_kernel_matches_oracle_and_the_k_ring_kernel(case):
    metric, n, d, kind, q, k, masked = case
    M = {"euclidean": DM.Euclidean, "cosine": DM.Cosine, "dot": DM.DotProduct}
    rows = oracle.gen_rows(500 + d, 0, n, d, kind)
    queries = oracle.gen_rows(600 + d, 0, q, d, kind)
    idx = gfi.GpuFlatIndex(M[metric])
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    idx.set_option("tensor_min_rows", 256)
    for i in range(5, n, 97):
        idx.remove(i)
    elig = np.ones(n, dtype=bool)
    elig[5::97] = False
    mask = None
    if masked:
        mask = (np.arange(n) * 2654435761 >> 4) % 3 != 0
        elig &= mask
    res = {}
    for sk in (1, 0):
        idx.set_option("short_k", sk)
        s0 = idx.stats()
        res[sk] = idx.search_arrays(queries, k, mask=mask)
        s1 = idx.stats()
        assert s1["tensor_queries"] - s0["tensor_queries"] == q, (sk, s0, s1)
        # (U[0,1) rows: a few queries have their k-th and (k+1)-th neighbours inside the fp16 error band and are
        # re-run exactly -- measured 0-1 for the short-K kernel (8 of 512 on the 300-row case); the k-ring kernel's seed order leaves 5 of 520 on the
        # first case and 85 of 600 on the masked all-positive cosine case, where every distance lies within 0.25 +- 0.02)
        assert s1["fallback_queries"] - s0["fallback_queries"] <= (q // 50 if sk else q // 6), (sk, s0, s1)
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    got_ids, got_d, cnt = res[1]
    exp = oracle.search_batch(metric, rows, queries[:48], k, eligible=elig, threads=8)
    for i, (eids, ed) in enumerate(exp):
        assert cnt[i] == len(eids)
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"{case} q{i}")


def test_short_k_uniform_coefficient_loop_selects_the_same_candidates():
    """Row tiles whose eligible rows share their fp16 scale take a hot loop that reads b alone (gemm_topk.cu,
    epi_block_hit<.., UNI>); gemm_debug bit 6 forces the (a, b)-pair loop.  Same scores bit for bit, hence the same
    candidates, fallbacks and answers -- with tombstones inside uniform tiles, a tile mixing two magnitudes, and a
    partial last tile."""
    n, d, q, k = 41_000, 128, 512, 10
    rows = oracle.gen_rows(910, 0, n, d, 0)
    rows[20_000:20_300] *= np.float32(8.0)      # two magnitudes inside row tiles 78 and 79: the general loop there
    idx = gfi.GpuFlatIndex(DM.Euclidean)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    idx.set_option("tensor_min_rows", 256)
    for i in range(3, n, 53):
        idx.remove(i)
    elig = np.ones(n, dtype=bool)
    elig[3::53] = False
    queries = oracle.gen_rows(911, 0, q, d, 0)
    res = {}
    for dbg in (0, 64):
        idx.set_option("gemm_debug", dbg)
        s0 = idx.stats()
        res[dbg] = idx.search_arrays(queries, k)
        s1 = idx.stats()
        res[dbg] += (s1["fallback_queries"] - s0["fallback_queries"],)
    idx.set_option("gemm_debug", 0)
    for a, b in zip(res[0], res[64]):
        assert np.array_equal(a, b)
    exp = oracle.search_batch("euclidean", rows, queries[:32], k, eligible=elig, threads=8)
    for i, (eids, ed) in enumerate(exp):
        assert_topk_matches(res[0][0][i, :res[0][2][i]], res[0][1][i, :res[0][2][i]], eids, ed, ctx=f"uniform loop q{i}")


# ---------------------------------------------------------------- rerank cut (select_kernel)
RERANK_CUT_CASES = [  # metric, n, d, kind, q, k
    ("cosine", 50000, 192, 1, 96, 10),
    ("euclidean", 60000, 128, 0, 512, 10),     # short-K kernel
    ("dot", 40000, 200, 1, 64, 100),           # KP = 512
    ("euclidean", 30000, 96, 1, 48, 1),
    ("cosine", 20000, 64, 0, 40, 33),          # all-positive rows: every distance within a narrow band
    ("euclidean", 30000, 64, 1, 32, 200),      # KP = 1024
]


@pytest.mark.parametrize("case", RERANK_CUT_CASES, ids=lambda c: "%s_n%d_d%d_q%d_k%d" % (c[0], c[1], c[2], c[4], c[5]))
def test_rerank_cut_never_changes_an_answer(case):
    """select_kernel re-scores only the candidates whose fp16 error interval reaches the k-th best one's
    (option rerank_cut, default 1).  Same ids and bit-identical distances with the cut off (0), on, and in the test mode
    that cuts right behind the k-th candidate (2), where the certification has to send almost every query to the
    exact scan; half of the queries are stored rows (k-th distances next to zero, exact duplicates of the query)."""
    metric, n, d, kind, q, k = case
    M = {"euclidean": DM.Euclidean, "cosine": DM.Cosine, "dot": DM.DotProduct}
    rows = oracle.gen_rows(700 + d, 0, n, d, kind)
    rows[1000:1040] = rows[77]                      # 40 exact copies of one row
    queries = oracle.gen_rows(800 + d, 0, q, d, kind)
    queries[::2] = rows[np.arange(0, q, 2) * 37 % n]
    queries[4] = rows[77]
    idx = gfi.GpuFlatIndex(M[metric])
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    idx.set_option("tensor_min_rows", 256)
    res, fb = {}, {}
    for mode in (1, 0, 2):
        idx.set_option("rerank_cut", mode)
        s0 = idx.stats()
        res[mode] = idx.search_arrays(queries, k)
        s1 = idx.stats()
        assert s1["tensor_queries"] - s0["tensor_queries"] == q, (mode, s0, s1)
        fb[mode] = s1["fallback_queries"] - s0["fallback_queries"]
    for mode in (0, 2):
        for a, b in zip(res[1], res[mode]):
            assert np.array_equal(a, b), (mode, fb)
    # a rigorous cut costs no extra fallbacks; the test mode falls back wherever the (k+1)-th row is inside the band
    assert fb[1] <= fb[0] + 1, fb
    assert fb[2] >= fb[1], fb
    got_ids, got_d, cnt = res[1]
    exp = oracle.search_batch(metric, rows, queries[:40], k, threads=8)
    for i, (eids, ed) in enumerate(exp):
        assert cnt[i] == len(eids)
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"{case} q{i}")
