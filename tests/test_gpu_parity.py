"""GPU parity tests: every call goes through the C ABI (libgfi.so via ctypes) and is checked
against the CPU oracle on the same seeded inputs.  Bar: identical ids in (distance, lower id)
order and BIT-IDENTICAL distances (libgfi re-scores survivors with the reference's arithmetic);
id swaps are tolerated only between distances closer than 1e-5 relative (north_star)."""
import numpy as np
import pytest

import oracle
import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200 import DistanceMetric as DM
from helpers import KATS, assert_topk_matches, numpy_merge

pytestmark = pytest.mark.gpu
M = {"euclidean": DM.Euclidean, "cosine": DM.Cosine, "dot": DM.DotProduct}


def build(metric, rows, ids=None, flags=0):
    idx = gfi.GpuFlatIndex(M[metric], flags=flags)
    n = rows.shape[0]
    idx.add_batch(np.arange(n, dtype=np.uint64) if ids is None else ids, rows)
    return idx


def check_batch(idx, metric, rows, queries, ks, ids=None, eligible=None, mask=None, ctx=""):
    got_ids, got_d, cnt = idx.search_arrays(queries, ks, mask=mask)
    exp = oracle.search_batch(metric, rows, queries, ks, ids=ids, eligible=eligible, threads=8)
    for i, (eids, ed) in enumerate(exp):
        assert cnt[i] == len(eids), f"{ctx} q{i}: count {cnt[i]} != {len(eids)}"
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"{ctx} q{i}")


# ---------------------------------------------------------------- reference KATs via the ABI
@pytest.mark.parametrize("kat", KATS["distance"], ids=lambda k: k["src"])
def test_reference_distance_kats_through_abi(kat):
    metric = "dot" if kat["metric"] == "rawdot" else kat["metric"]
    idx = gfi.GpuFlatIndex(M[metric])
    idx.add(0, kat["b"])
    (rid, dist), = idx.search(kat["a"], 1)
    got = -dist if kat["metric"] == "rawdot" else dist
    assert rid == 0 and abs(got - kat["expect"]) <= kat["tol"] * max(1.0, abs(kat["expect"]))
    assert np.float32(dist) == oracle.distance(metric, kat["a"], kat["b"])  # bit-identical


def test_reference_flat_index_kats():
    k = KATS["flat_index_basic"]
    idx = gfi.GpuFlatIndex(DM.Euclidean)
    for i, v in k["rows"].items():
        idx.add(int(i), v)
    res = idx.search(k["query"], k["k"])
    assert len(res) == k["expect_len"] and res[0][0] == k["expect_first_id"] and res[0][1] < 1e-6
    # get_vector hit/miss (flat_index.rs:96-103)
    assert np.array_equal(idx.get_vector(0), np.array(k["rows"]["0"], dtype=np.float32))
    assert idx.get_vector(99) is None
    # remove => len (flat_index.rs:106-114)
    r = KATS["flat_index_remove"]
    idx2 = gfi.GpuFlatIndex(DM.Euclidean)
    for i, v in r["rows"].items():
        idx2.add(int(i), v)
    assert idx2.len() == 2
    idx2.remove(r["remove"])
    assert idx2.len() == r["expect_len"]
    idx2.remove(r["remove"])  # idempotent
    assert idx2.len() == r["expect_len"]
    assert idx2.metric() == DM.Euclidean


def test_reference_store_kats():
    s = KATS["store_search"]
    store = gfi.VectorStore(DM.Euclidean)
    assert store.search([1, 2, 3], 5) == []  # storage.rs:399-404 empty store
    for sid, v in s["rows"].items():
        store.insert(sid, v)
    res = store.search(s["query"], s["k"])
    assert len(res) == 2 and res[0].id == "v1" and res[0].distance == 0.0
    with pytest.raises(gfi.DimensionMismatch):
        store.insert("bad", [1.0, 2.0])  # dimension latch, storage.rs:145-154
    with pytest.raises(gfi.DimensionMismatch):
        store.search([1.0, 2.0], 1)
    w = KATS["integration_workflow"]
    store = gfi.VectorStore(DM.Euclidean)
    for sid, v in w["rows"].items():
        store.insert(sid, v)
    assert store.len() == 3
    res = store.search(w["query"], w["k"])
    assert len(res) == 2 and res[0].id == "v1"
    deleted = store.delete("v2")
    assert np.array_equal(deleted, np.array([0, 1, 0], dtype=np.float32)) and store.len() == 2
    k = KATS["integration_metrics_self_match"]
    for m in k["metrics"]:
        st = gfi.VectorStore(M[m])
        st.insert("v1", k["rows"]["v1"])
        res = st.search(k["query"], 1)
        assert len(res) == 1 and res[0].id == "v1"


@pytest.mark.parametrize("pushdown", [False, True, "device"])
def test_reference_filter_and_batch_kats(pushdown):
    def mk(case):
        st = gfi.VectorStore(DM.Euclidean)
        for sid, (v, md) in case["rows"].items():
            m = gfi.Metadata()
            for a, b in md.items():
                m.insert(a, b)
            st.insert_with_metadata(sid, v, m)
        return st
    c = KATS["filter_matching"]
    res = mk(c).search_with_filter(c["query"], c["k"], gfi.MetadataFilter.from_json(c["filter"]), pushdown=pushdown)
    assert [r.id for r in res] == c["expect_ids"]
    c = KATS["filter_none_matching"]
    assert mk(c).search_with_filter(c["query"], c["k"], gfi.MetadataFilter.from_json(c["filter"]),
                                    pushdown=pushdown) == []
    c = KATS["filter_all_matching"]
    assert len(mk(c).search_with_filter(c["query"], c["k"], gfi.MetadataFilter.from_json(c["filter"]),
                                        pushdown=pushdown)) == 2
    c = KATS["batch_search_with_filter"]
    res = mk(c).search_batch_with_filter([(q, k) for q, k in c["queries"]],
                                         gfi.MetadataFilter.from_json(c["filter"]), pushdown=pushdown)
    assert [[r.id for r in rr] for rr in res] == c["expect_ids"]
    c = KATS["batch_search"]
    st = gfi.VectorStore(DM.Euclidean)
    for sid, v in c["rows"].items():
        st.insert(sid, v)
    for batched in (False, True):
        res = st.search_batch([(q, k) for q, k in c["queries"]], batched=batched)
        assert [r[0].id for r in res] == c["expect_first_ids"]


# ---------------------------------------------------------------- scan path vs oracle
SCAN_CASES = [  # n, d, kind, q, k
    (10000, 128, 0, 3, 10),    # C1 (benches/search_bench.rs:18-33)
    (3000, 768, 1, 2, 100),    # C3 scaled down (column-segmented rows)
    (4000, 384, 0, 5, 10),     # C4 scaled down
    (700, 3, 1, 4, 5),         # tiny d (padding path)
    (900, 17, 0, 9, 33),       # odd d, q > 8 (two passes), k -> K=64
    (1200, 1000, 1, 1, 10),    # d > 256 and not a multiple of 256
    (257, 64, 1, 2, 300),      # k > n
    (600, 1500, 1, 2, 10),     # rows longer than 4 KB: column-segmented ring stages
    (500, 2100, 0, 5, 7),      # three column segments, q > 4 (two passes)
    (3000, 64, 0, 1, 1000),    # k = 1000: the largest list length (K = 1024)
]


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot"])
@pytest.mark.parametrize("case", SCAN_CASES, ids=lambda c: "n%d_d%d_q%d_k%d" % (c[0], c[1], c[3], c[4]))
def test_scan_path_matches_oracle(metric, case):
    n, d, kind, q, k = case
    rows = oracle.gen_rows(100 + d, 0, n, d, kind)
    queries = oracle.gen_rows(200 + d, 0, q, d, kind)
    idx = build(metric, rows, flags=1)  # GFI_FLAG_NO_TENSOR: scan kernel only
    check_batch(idx, metric, rows, queries, k, ctx=f"{metric} {case}")
    st = idx.stats()
    assert st["scan_queries"] == q and st["tensor_queries"] == 0 and st["kernel_launches"] > 0


def test_bench_query_constant_half(tmp_path):
    # benches/search_bench.rs:27: query = 0.5 * ones, x ~ U[0,1)
    rows = oracle.gen_rows(1, 0, 10000, 128, 0)
    idx = build("euclidean", rows)
    check_batch(idx, "euclidean", rows, np.full((1, 128), 0.5, np.float32), 10, ctx="C1 bench query")


def test_per_query_k_and_edge_ks():
    rows = oracle.gen_rows(5, 0, 500, 32, 1)
    queries = oracle.gen_rows(6, 0, 6, 32, 1)
    idx = build("dot", rows)
    check_batch(idx, "dot", rows, queries, [1, 0, 7, 500, 600, 3], ctx="per-query k")


def test_generated_rows_equal_oracle_generator():
    for kind in (0, 1):
        idx = gfi.GpuFlatIndex(DM.Euclidean, dim=40)
        idx.add_generated(77, (1 << 33) + 10, 1000, kind, 5000)
        exp = oracle.gen_rows(77, (1 << 33) + 10, 1000, 40, kind)
        for i in (0, 1, 499, 999):
            assert np.array_equal(idx.get_vector(5000 + i), exp[i])
        q = oracle.gen_rows(78, 0, 2, 40, kind)
        check_batch(idx, "euclidean", exp, q, 10, ids=np.arange(5000, 6000, dtype=np.uint64), ctx="generated")


# ---------------------------------------------------------------- mutation semantics
def test_remove_overwrite_out_of_order_and_compact():
    rng = np.random.default_rng(3)
    rows = rng.standard_normal((400, 24)).astype(np.float32)
    ids = np.arange(400, dtype=np.uint64) * 3 + 10
    idx = build("euclidean", rows, ids=ids)
    q = rng.standard_normal((3, 24)).astype(np.float32)
    live = np.ones(400, dtype=bool)
    for r in (0, 5, 399, 200):
        idx.remove(int(ids[r]))
        live[r] = False
    idx.remove(123456)  # missing id: Ok(())
    assert idx.len() == 396
    check_batch(idx, "euclidean", rows[live], q, 10, ids=ids[live], ctx="after remove")
    # overwrite (HashMap::insert semantics) and an id below the current maximum
    new_row = rng.standard_normal(24).astype(np.float32)
    idx.add(int(ids[7]), new_row)
    rows2 = rows.copy()
    rows2[7] = new_row
    low = rng.standard_normal(24).astype(np.float32)
    idx.add(1, low)
    assert idx.len() == 397
    all_rows = np.concatenate([low[None], rows2[live]])
    all_ids = np.concatenate([[1], ids[live]]).astype(np.uint64)
    check_batch(idx, "euclidean", all_rows, q, 397, ids=all_ids, ctx="after overwrite")
    assert np.array_equal(idx.get_vector(int(ids[7])), new_row) and idx.get_vector(int(ids[0])) is None
    idx.compact()
    assert idx.stats()["n_slots"] == 397
    check_batch(idx, "euclidean", all_rows, q, 20, ids=all_ids, ctx="after compact")


def test_duplicate_rows_tie_break_by_lower_id():
    base = oracle.gen_rows(9, 0, 50, 16, 1)
    rows = np.concatenate([base] * 40)  # 2000 rows, every vector 40 times
    idx = build("cosine", rows)
    check_batch(idx, "cosine", rows, base[:3], 60, ctx="duplicates")
    # signed zeros: -dot = -0.0 ties with +0.0, lower id wins
    idx = gfi.GpuFlatIndex(DM.DotProduct)
    idx.add(2, [0.0, 1.0])
    idx.add(1, [0.0, -1.0])
    assert [i for i, _ in idx.search([1.0, 0.0], 2)] == [1, 2]


def test_error_semantics():
    idx = gfi.GpuFlatIndex(DM.Cosine)
    idx.add(0, [1.0, 0.0])
    idx.add(1, [0.0, 0.0])
    with pytest.raises(gfi.InvalidVector):   # distance.rs:51-55 via flat_index.rs:57-60
        idx.search([1.0, 1.0], 1)
    idx.remove(1)
    assert idx.search([1.0, 1.0], 1)[0][0] == 0
    with pytest.raises(gfi.InvalidVector):
        idx.search([0.0, 0.0], 1)
    nan = gfi.GpuFlatIndex(DM.Euclidean)
    nan.add(0, [float("nan"), 0.0])
    nan.add(1, [1.0, 0.0])
    with pytest.raises(gfi.NaNDistance):
        nan.search([0.0, 0.0], 1)
    mixed = gfi.GpuFlatIndex(DM.Euclidean)
    mixed.add(0, [1.0, 2.0, 3.0])
    with pytest.raises(gfi.DimensionMismatch) as e:  # distance.rs:21-26: expected = query dim
        mixed.search([1.0, 2.0], 1)
    assert (e.value.expected, e.value.actual) == (2, 3)
    mixed.add(1, [1.0, 2.0])  # FlatIndex::add never checks dimensions
    assert mixed.len() == 2
    with pytest.raises(gfi.DimensionMismatch):
        mixed.search([1.0, 2.0, 3.0], 1)
    mixed.remove(1)
    assert mixed.search([1.0, 2.0, 3.0], 1)[0] == (0, 0.0)
    empty = gfi.GpuFlatIndex(DM.Euclidean)
    assert empty.search([1.0, 2.0], 3) == [] and empty.is_empty()


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot"])
def test_small_search_fused_tail_and_latency_mode_equal_the_separate_kernels(metric):
    """Small plain searches finish inside the scan kernel (fused tail) and, from the host, without copies or a
    stream synchronisation (latency mode).  Both must return what the separate select / rerank kernels and the
    copying path return -- and the oracle's answer -- including k > n, per-query k and the error flags."""
    n, d = 7000, 200
    rows = oracle.gen_rows(71, 0, n, d, 1)
    queries = oracle.gen_rows(72, 0, 4, d, 1)
    ks = np.array([10, 1, 56, 33], dtype=np.uint32)
    idx = build(metric, rows, flags=1)
    for i in (0, 17, 6999):
        idx.remove(i)
    live = np.ones(n, dtype=bool)
    live[[0, 17, 6999]] = False
    ids = np.arange(n, dtype=np.uint64)
    exp = oracle.search_batch(metric, rows[live], queries, ks, ids=ids[live], threads=4)
    outs = []
    for fused, zc in ((1, 1), (1, 0), (0, 1), (0, 0)):
        idx.set_option("fused_tail", fused)
        idx.set_option("zero_copy", zc)
        for nq in (4, 1, 3):
            got_ids, got_d, cnt = idx.search_arrays(queries[:nq], ks[:nq])
            for i in range(nq):
                assert cnt[i] == len(exp[i][0])
                assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], exp[i][0], exp[i][1],
                                    ctx=f"{metric} fused={fused} zc={zc} nq={nq} q{i}")
        outs.append(idx.search_arrays(queries, ks))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    # k > n on a tiny index, and the flags raised inside the fused tail
    idx.set_option("fused_tail", 1)
    idx.set_option("zero_copy", 1)
    tiny = build(metric, rows[:5], flags=1)
    res = tiny.search(queries[0], 40)
    e_ids, e_d = oracle.search_batch(metric, rows[:5], queries[:1], 40, threads=1)[0]
    assert [i for i, _ in res] == [int(x) for x in e_ids]
    assert np.array_equal(np.array([x for _, x in res], np.float32), e_d)
    if metric == "cosine":
        with pytest.raises(gfi.InvalidVector):
            idx.search(np.zeros(d, np.float32), 5)
    bad = queries[0].copy()
    bad[3] = np.nan
    with pytest.raises(gfi.NaNDistance):
        idx.search(bad, 5)
    assert idx.search(queries[1], 3)  # the handle is usable after an error


# ---------------------------------------------------------------- filters
@pytest.mark.parametrize("sel", [0.01, 0.5])
def test_mask_pushdown_and_post_filter(sel):
    n, d, k = 20000, 96, 10
    rows = oracle.gen_rows(6, 0, n, d, 0)
    queries = oracle.gen_rows(7, 0, 4, d, 0)
    rng = np.random.default_rng(1)
    elig = rng.random(n) < sel
    idx = build("euclidean", rows, flags=1)
    check_batch(idx, "euclidean", rows, queries, k, eligible=elig, mask=elig, ctx=f"mask {sel}")
    # reference post-filter (fetch_k = 3k) through the unchanged single-query trait
    for qi in range(2):
        fetch = idx.search(queries[qi], 3 * k)
        got = [(i, dd) for i, dd in fetch if elig[i]][:k]
        eids, ed = oracle.search_post_filter("euclidean", rows, queries[qi], k, elig)
        assert [i for i, _ in got] == [int(x) for x in eids]
        assert np.array_equal(np.array([x for _, x in got], np.float32), ed)


def test_mask_by_internal_id_with_gaps_in_the_id_space():
    """The caller's mask is indexed by INTERNAL ID (bit i <=> id i); with gaps in the id space the compaction builds
    each 32-slot eligibility word from the id column (compact_eligible_kernel<false>).  Ids beyond the mask's length
    are ineligible."""
    n, d, k = 9000, 96, 10
    rows = oracle.gen_rows(61, 0, n, d, 1)
    queries = oracle.gen_rows(62, 0, 5, d, 1)
    rng = np.random.default_rng(4)
    ids = np.sort(rng.choice(40000, n, replace=False)).astype(np.uint64)
    for flags in (1, 0):
        idx = build("cosine", rows, ids=ids, flags=flags)
        for r in (3, 500, 8999):
            idx.remove(int(ids[r]))
        live = np.ones(n, dtype=bool)
        live[[3, 500, 8999]] = False
        for nbits in (40000, 25000):
            mask = rng.random(nbits) < 0.3
            elig_rows = np.zeros(n, dtype=bool)
            inside = ids < nbits
            elig_rows[inside] = mask[ids[inside].astype(np.int64)]
            got_ids, got_d, cnt = idx.search_arrays(queries, k, mask=mask)
            exp = oracle.search_batch("cosine", rows[live], queries, k, ids=ids[live], eligible=elig_rows[live], threads=4)
            for i, (eids, ed) in enumerate(exp):
                assert cnt[i] == len(eids)
                assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"id mask {nbits} q{i}")


@pytest.mark.parametrize("metric,d", [("euclidean", 384), ("cosine", 100), ("dot", 768)])
def test_gather_scan_merges_runs_of_adjacent_eligible_rows(metric, d):
    """The gather scan copies a run of adjacent eligible slots as ONE bulk copy (scan.cu producer): runs of every
    length from 1 to 70 (longer than a 32-entry chunk and than a stage), isolated rows, an all-eligible stretch and
    tombstones cutting runs must give the oracle's answer bit for bit."""
    n, k = 40000, 10
    rows = oracle.gen_rows(16, 0, n, d, 1)
    queries = oracle.gen_rows(17, 0, 3, d, 1)
    elig = np.zeros(n, dtype=bool)
    pos, run = 0, 1
    while pos + run + 3 < n // 2:  # run lengths 1, 2, ..., 70, 1, 2, ... separated by gaps of 1-3
        elig[pos:pos + run] = True
        pos += run + 1 + (run % 3)
        run = run % 70 + 1
    elig[n // 2:n // 2 + 5000] = True  # one long stretch
    elig[n // 2 + 6000::7] = True      # isolated rows
    idx = build(metric, rows, flags=1)
    check_batch(idx, metric, rows, queries, k, eligible=elig, mask=elig, ctx="runs")
    removed = np.arange(3, n, 11)
    for i in removed:
        idx.remove(int(i))
    elig2 = elig.copy()
    elig2[removed] = False
    check_batch(idx, metric, rows, queries, k, eligible=elig2, mask=elig, ctx="runs + tombstones")


# ---------------------------------------------------------------- tensor (tcgen05) path vs oracle
TENSOR_CASES = [  # metric, n, d, kind, q, k
    ("cosine", 20000, 768, 1, 64, 10),     # C2 scaled down
    ("euclidean", 30000, 128, 0, 130, 10),  # C5 scaled down, 2 query tiles (one partial)
    ("dot", 16384, 768, 1, 32, 100),       # C3b scaled down
    ("euclidean", 25000, 200, 1, 40, 10),  # d not a multiple of 64 (TMA zero fill along K)
    ("cosine", 9000, 72, 0, 16, 5),        # concentrated cosines (uniform data), n not a tile multiple
]


@pytest.mark.parametrize("case", TENSOR_CASES, ids=lambda c: "%s_n%d_d%d_q%d_k%d" % (c[0], c[1], c[2], c[4], c[5]))
def test_tensor_path_matches_oracle(case):
    metric, n, d, kind, q, k = case
    rows = oracle.gen_rows(300 + d, 0, n, d, kind)
    queries = oracle.gen_rows(400 + d, 0, q, d, kind)
    idx = build(metric, rows)
    check_batch(idx, metric, rows, queries, k, ctx=str(case))
    st = idx.stats()
    assert st["tensor_queries"] == q, st
    # certification normally succeeds: measured 0 fallbacks, except 11 of 130 on the uniform-data L2 case (30000 rows
    # of [0, 1)^128: distances crowd the k-th one more closely than the fp16 bound can separate)
    assert st["fallback_queries"] <= (16 if (metric, n) == ("euclidean", 30000) else 2), st


def test_tensor_path_cosine_coefficient_epilogue_and_tombstones():
    """Cosine normally takes the raw-accumulator epilogue; the per-row-coefficient epilogue must give the same
    answer, and the raw epilogue must drop tombstoned rows (it meets them as ordinary accumulators)."""
    n, d, q, k = 24000, 256, 48, 10
    rows = oracle.gen_rows(51, 0, n, d, 1)
    queries = oracle.gen_rows(52, 0, q, d, 1)
    idx = build("cosine", rows)
    check_batch(idx, "cosine", rows, queries, k, ctx="raw epilogue")
    idx.set_option("raw_epilogue", 0)
    check_batch(idx, "cosine", rows, queries, k, ctx="coefficient epilogue")
    idx.set_option("raw_epilogue", 1)
    # remove the current winners of every query: they must vanish from the next answer
    ids0, _, _ = idx.search_arrays(queries, k)
    live = np.ones(n, dtype=bool)
    for r in np.unique(ids0[:, :3]):
        idx.remove(int(r))
        live[int(r)] = False
    got_ids, got_d, cnt = idx.search_arrays(queries, k)
    ids = np.arange(n, dtype=np.uint64)
    exp = oracle.search_batch("cosine", rows[live], queries, k, ids=ids[live], threads=8)
    for i, (eids, ed) in enumerate(exp):
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"raw tombstones q{i}")
    assert idx.stats()["tensor_queries"] == 4 * q


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_tensor_path_cta_pair_kernel(metric):
    """The cta_group::2 (CTA pair, M = 256) instance of the tensor pass, behind set_option("pair", 1): same
    answers as the single-CTA instance, for the raw (cosine) and the coefficient (euclidean) epilogue."""
    n, d, q, k = 30000, 192, 300, 10      # 3 query tiles -> padded to 4: two tile pairs, one of them half empty
    rows = oracle.gen_rows(71, 0, n, d, 1)
    queries = oracle.gen_rows(72, 0, q, d, 1)
    idx = build(metric, rows)
    idx.set_option("pair", 1)
    check_batch(idx, metric, rows, queries, k, ctx="pair kernel")
    st = idx.stats()
    assert st["tensor_queries"] == q and st["fallback_queries"] <= 2, st  # (measured: 0)


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot"])
def test_tensor_path_rows_of_wildly_different_magnitude(metric):
    """Row norms spread over 24 orders of magnitude (and a batch of queries with its own spread): the fp16 shadow
    uses per-row scales (normalised rows for cosine), the certification bound is relative to the largest row, and
    whatever cannot be certified must come back exact through the fallback."""
    n, d, q, k = 12000, 96, 40, 10
    rng = np.random.default_rng(11)
    rows = oracle.gen_rows(131, 0, n, d, 1) * (10.0 ** rng.uniform(-12, 12, size=(n, 1))).astype(np.float32)
    queries = oracle.gen_rows(132, 0, q, d, 1) * (10.0 ** rng.uniform(-6, 6, size=(q, 1))).astype(np.float32)
    rows, queries = rows.astype(np.float32), queries.astype(np.float32)
    idx = build(metric, rows)
    check_batch(idx, metric, rows, queries, k, ctx=f"magnitudes {metric}")
    assert idx.stats()["tensor_queries"] == q


def test_tensor_path_error_semantics_in_a_batch():
    """One bad query fails the whole batch, as the sequential batch loop's `collect::<Result<_>>` does
    (src/storage.rs:302-310): a zero vector under cosine -> InvalidVector, a NaN element -> the NaN panic's status;
    a NaN / zero ROW does the same for every query that reaches it."""
    n, d, q = 9000, 64, 48
    rows = oracle.gen_rows(141, 0, n, d, 1)
    queries = oracle.gen_rows(142, 0, q, d, 1)
    cos = build("cosine", rows)
    bad = queries.copy()
    bad[17] = 0.0
    with pytest.raises(gfi.InvalidVector):
        cos.search_arrays(bad, 10)
    bad = queries.copy()
    bad[5, 3] = np.nan
    with pytest.raises(gfi.NaNDistance):
        cos.search_arrays(bad, 10)
    cos.search_arrays(queries, 10)                     # the index still answers clean batches
    assert cos.stats()["tensor_queries"] >= q
    l2 = build("euclidean", rows)
    l2.add(n, np.full(d, np.nan, dtype=np.float32))    # a stored NaN row poisons every search (flat_index.rs:62)
    with pytest.raises(gfi.NaNDistance):
        l2.search_arrays(queries, 10)
    l2.remove(n)
    ids, dist, cnt = l2.search_arrays(queries, 10)
    exp = oracle.search_batch("euclidean", rows, queries, 10, threads=4)
    for i, (eids, ed) in enumerate(exp):
        assert_topk_matches(ids[i, :cnt[i]], dist[i, :cnt[i]], eids, ed, ctx=f"after nan row q{i}")
    cz = build("cosine", rows)
    cz.add(n, np.zeros(d, dtype=np.float32))           # zero row under cosine: Err(InvalidVector) for every query
    with pytest.raises(gfi.InvalidVector):
        cz.search_arrays(queries, 10)


def test_tensor_path_with_mask_and_tombstones():
    n, d, q, k = 20000, 256, 48, 10
    rows = oracle.gen_rows(31, 0, n, d, 1)
    queries = oracle.gen_rows(32, 0, q, d, 1)
    idx = build("euclidean", rows)
    rng = np.random.default_rng(2)
    live = np.ones(n, dtype=bool)
    for r in rng.choice(n, 500, replace=False):
        idx.remove(int(r))
        live[r] = False
    elig = rng.random(n) < 0.5
    ids = np.arange(n, dtype=np.uint64)
    got_ids, got_d, cnt = idx.search_arrays(queries, k, mask=elig)
    exp = oracle.search_batch("euclidean", rows[live], queries, k, ids=ids[live], eligible=elig[live], threads=8)
    for i, (eids, ed) in enumerate(exp):
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"tensor mask q{i}")
    assert idx.stats()["tensor_queries"] == q


def test_tensor_path_falls_back_when_not_certifiable():
    # 64 copies of every vector: the top-k boundary is an exact tie, which can never be certified
    base = oracle.gen_rows(41, 0, 256, 128, 1)
    rows = np.concatenate([base] * 64)  # 16384 rows
    queries = oracle.gen_rows(42, 0, 32, 128, 1)
    idx = build("euclidean", rows)
    check_batch(idx, "euclidean", rows, queries, 10, ctx="fallback")
    st = idx.stats()
    assert st["tensor_queries"] == 32 and st["fallback_queries"] > 0, st


def test_repeated_small_searches_stay_exact_across_mutations():
    """Back-to-back small host searches reuse a pooled search context (stream, device and pinned buffers): the
    answers must track new query data, changes of k and batch shape, buffer growth by a large batch in between,
    removals and masked searches."""
    n, d = 5000, 96
    rows = oracle.gen_rows(61, 0, n, d, 1)
    idx = build("euclidean", rows)
    live = np.ones(n, dtype=bool)
    ids = np.arange(n, dtype=np.uint64)

    def check(qseed, q, k, ctx):
        queries = oracle.gen_rows(qseed, 0, q, d, 1)
        got_ids, got_d, cnt = idx.search_arrays(queries, k)
        exp = oracle.search_batch("euclidean", rows[live], queries, k, ids=ids[live], threads=4)
        for i, (eids, ed) in enumerate(exp):
            assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"{ctx} q{i}")

    for rep in range(5):                      # same shape, different queries each time
        check(700 + rep, 1, 10, f"rep{rep}")
    check(710, 1, 5, "other k")
    for rep in range(3):
        check(720 + rep, 4, 10, f"q4 rep{rep}")
    check(730, 300, 10, "tensor batch in between")   # grows the context's buffers
    for rep in range(3):
        check(740 + rep, 1, 10, f"after big rep{rep}")
    for r in (17, 4000, 4999):
        idx.remove(r)
        live[r] = False
    for rep in range(3):
        check(750 + rep, 1, 10, f"after remove rep{rep}")
    elig = np.zeros(n, dtype=bool)
    elig[::7] = True
    q1 = oracle.gen_rows(760, 0, 1, d, 1)
    g_ids, g_d, g_c = idx.search_arrays(q1, 10, mask=elig)
    e_ids, e_d = oracle.search_batch("euclidean", rows[live], q1, 10, ids=ids[live], eligible=elig[live])[0]
    assert_topk_matches(g_ids[0, :g_c[0]], g_d[0, :g_c[0]], e_ids, e_d, ctx="masked")
    for rep in range(3):
        check(770 + rep, 1, 10, f"after mask rep{rep}")


# ---------------------------------------------------------------- device-pointer API + merge kernel
def test_device_search_and_merge_kernel():
    import torch
    n, d, q, k, G = 6000, 64, 5, 10, 3
    rows = oracle.gen_rows(51, 0, n, d, 1)
    queries = oracle.gen_rows(52, 0, q, d, 1)
    bounds = [0, 1500, 4200, n]
    dq = torch.from_numpy(queries).cuda()
    dks = torch.full((q,), k, dtype=torch.int32, device="cuda")
    all_ids = torch.zeros((G, q, k), dtype=torch.int64, device="cuda")
    all_d = torch.zeros((G, q, k), dtype=torch.float32, device="cuda")
    all_c = torch.zeros((G, q), dtype=torch.int32, device="cuda")
    shards = []
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    for g in range(G):
        lo, hi = bounds[g], bounds[g + 1]
        idx = build("euclidean", rows[lo:hi], ids=np.arange(lo, hi, dtype=np.uint64))
        shards.append(idx)
        idx.search_device(dq.data_ptr(), q, dks.data_ptr(), k, all_ids[g].data_ptr(), all_d[g].data_ptr(),
                          all_c[g].data_ptr(), k, stream=stream)
        idx.search_status()
    out_ids = torch.zeros((q, k), dtype=torch.int64, device="cuda")
    out_d = torch.zeros((q, k), dtype=torch.float32, device="cuda")
    out_c = torch.zeros((q,), dtype=torch.int32, device="cuda")
    shards[0].merge_topk_device(all_ids.data_ptr(), all_d.data_ptr(), all_c.data_ptr(), G, q, k, dks.data_ptr(),
                                out_ids.data_ptr(), out_d.data_ptr(), out_c.data_ptr(), k, stream=stream)
    torch.cuda.synchronize()
    exp = oracle.search_batch("euclidean", rows, queries, k, threads=4)
    ref_merge = numpy_merge(all_ids.cpu().numpy(), all_d.cpu().numpy(), all_c.cpu().numpy(), [k] * q)
    for i in range(q):
        assert out_c[i].item() == k
        assert [int(x) for x in out_ids[i].cpu()] == [p[1] for p in ref_merge[i]] == [int(x) for x in exp[i][0]]
        assert np.array_equal(out_d[i].cpu().numpy(), exp[i][1])
    # packed per-shard blocks [ids | dist | counts] (what ONE all-gather per search delivers) + strided merge
    from vectordb_from_scratch_b200.sharded import packed_layout, packed_views
    size = packed_layout(q, k)[2]
    packed = torch.zeros((G, size), dtype=torch.uint8, device="cuda")
    p_ids, p_d, p_c = packed_views(packed, q, k)
    for g in range(G):
        shards[g].search_device(dq.data_ptr(), q, dks.data_ptr(), k, p_ids[g].data_ptr(), p_d[g].data_ptr(),
                                p_c[g].data_ptr(), k, stream=stream)
        shards[g].search_status()
    out2_ids, out2_d, out2_c = torch.zeros_like(out_ids), torch.zeros_like(out_d), torch.zeros_like(out_c)
    shards[0].merge_topk_device(p_ids.data_ptr(), p_d.data_ptr(), p_c.data_ptr(), G, q, k, dks.data_ptr(),
                                out2_ids.data_ptr(), out2_d.data_ptr(), out2_c.data_ptr(), k, stream=stream,
                                shard_stride_bytes=size)
    torch.cuda.synchronize()
    assert torch.equal(out2_ids, out_ids) and torch.equal(out2_d, out_d) and torch.equal(out2_c, out_c)


def test_device_side_route_of_small_batches_with_a_device_mask():
    """A mask that lives on the device has a population the host does not know: the choice between the gather scan
    and the masked tensor pass is made on the device from the compaction's count (kernels.h RouteParams).  Both
    outcomes must give the oracle's answer bit for bit, and the statistics must say which kernel answered."""
    import torch
    from vectordb_from_scratch_b200.index import pack_mask
    n, d, q, k = 600_000, 384, 3, 10  # large enough for the tensor pass to be worth considering
    rows = oracle.gen_rows(6, 0, n, d, 0)
    queries = oracle.gen_rows(7, 0, q, d, 0)
    idx = gfi.GpuFlatIndex(DM.Euclidean, dim=d)
    idx.add_generated(6, 0, n, 0, 0)
    for i in (5, 77, 4097, n - 1):
        idx.remove(i)
    dq = torch.from_numpy(queries).cuda()
    dks = torch.full((q,), k, dtype=torch.int32, device="cuda")
    out_ids = torch.zeros((q, k), dtype=torch.int64, device="cuda")
    out_d = torch.zeros((q, k), dtype=torch.float32, device="cuda")
    out_c = torch.zeros((q,), dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    rng = np.random.default_rng(3)
    # The previous mask's population predicts whether the tensor kernels are worth enqueuing at all, so a dense mask
    # right after a sparse one is still answered by the scan once (a misprediction costs time, never correctness).
    masks = {0.02: rng.random(n) < 0.02, 0.95: rng.random(n) < 0.95}
    for sel, route in ((0.02, "scan"), (0.95, "scan"), (0.95, "tensor"), (0.02, "scan"), (0.02, "scan")):
        elig = masks[sel]
        words, bits = pack_mask(elig)
        dmask = torch.from_numpy(words.view(np.int64)).cuda()
        st0 = idx.stats()
        idx.search_device(dq.data_ptr(), q, dks.data_ptr(), k, out_ids.data_ptr(), out_d.data_ptr(), out_c.data_ptr(),
                          k, stream=ts.cuda_stream, d_mask=dmask.data_ptr(), mask_bits=bits)
        idx.search_status()
        torch.cuda.synchronize()
        st1 = idx.stats()
        live = elig.copy()
        live[[5, 77, 4097, n - 1]] = False
        exp = oracle.search_batch("euclidean", rows, queries, k, eligible=live, threads=8)
        for i, (eids, ed) in enumerate(exp):
            assert out_c[i].item() == k
            assert_topk_matches(out_ids[i].cpu().numpy().astype(np.uint64), out_d[i].cpu().numpy(), eids, ed,
                                ctx=f"route {route} q{i}")
        dscan, dtensor = st1["scan_queries"] - st0["scan_queries"], st1["tensor_queries"] - st0["tensor_queries"]
        assert (dscan, dtensor) == ((q, 0) if route == "scan" else (0, q)), (route, dscan, dtensor)


# ---------------------------------------------------------------- full-size checks (BASELINE.json configs)
def test_full_size_c2_cosine_batch_against_oracle_subset():
    """C2: 1M x 768 cosine, batch 1024, k = 10.  The whole batch runs on the tcgen05 path; the oracle checks 32 of
    the queries at full size, and size-independent properties are checked for all of them."""
    n, d, q, k = 1_000_000, 768, 1024, 10
    idx = gfi.GpuFlatIndex(DM.Cosine, dim=d)
    idx.reserve(n)
    idx.add_generated(3, 0, n, 1, 0)
    queries = oracle.gen_rows(4, 0, q, d, 1)
    ids, dist, cnt = idx.search_arrays(queries, k)
    assert np.all(cnt == k)
    assert np.all(np.diff(dist, axis=1) >= 0) and np.all((dist >= 0) & (dist <= 2))
    st = idx.stats()
    assert st["tensor_queries"] == q and st["fallback_queries"] <= 8, st
    # idempotence and agreement of the two independent GPU paths (tensor vs exact scan) on 8 queries
    ids2, dist2, _ = idx.search_arrays(queries, k)
    assert np.array_equal(ids, ids2) and np.array_equal(dist, dist2)
    idx.set_option("tensor_min_q", 1 << 30)
    ids_s, dist_s, _ = idx.search_arrays(queries[:8], k)
    assert np.array_equal(ids_s, ids[:8]) and np.array_equal(dist_s, dist[:8])
    # oracle parity on 32 queries at full size (rows-parallel oracle over the generated rows)
    pick = np.arange(0, q, q // 32)[:32]
    exp = oracle.search_generated("cosine", 3, 0, n, d, 1, queries[pick], k)
    for j, (eids, ed) in zip(pick, exp):
        assert_topk_matches(ids[j], dist[j], eids, ed, ctx=f"C2 full q{j}")


# ---------------------------------------------------------------- concurrency (B2: &self searches under RwLock::read)
def test_concurrent_searches_from_many_threads():
    """routes.rs:244,342: any tokio worker may call Index::search concurrently.  ctypes releases the GIL
    during the foreign call, so these really overlap inside libgfi (search-context pool, one stream each)."""
    import threading
    n, d, k = 30000, 96, 10
    rows = oracle.gen_rows(81, 0, n, d, 1)
    idx = build("euclidean", rows)
    batches = [oracle.gen_rows(900 + t, 0, 1 if t % 2 else 24, d, 1) for t in range(8)]
    expected = [idx.search_arrays(b, k) for b in batches]  # sequential answers
    results, errors = [None] * len(batches), []

    def worker(t):
        try:
            for _ in range(20):
                results[t] = idx.search_arrays(batches[t], k)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(len(batches))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for (ei, ed, ec), (gi, gd, gc) in zip(expected, results):
        assert np.array_equal(ei, gi) and np.array_equal(ed, gd) and np.array_equal(ec, gc)
    exp0 = oracle.search_batch("euclidean", rows, batches[0], k, threads=4)
    for i, (eids, ed) in enumerate(exp0):
        assert_topk_matches(results[0][0][i], results[0][1][i], eids, ed, ctx=f"concurrent q{i}")


def test_concurrent_single_queries_are_coalesced_and_isolated():
    """Group commit (N1 micro-batching inside the library): single-query calls from many threads are combined into
    shared scans while another search is running; every caller gets exactly its own answer -- or its own error."""
    import threading
    n, d, k, T, per = 150000, 64, 10, 24, 12
    rows = oracle.gen_rows(151, 0, n, d, 1)
    idx = build("cosine", rows)
    queries = oracle.gen_rows(152, 0, T * per, d, 1)
    queries[5 * per + 3] = 0.0                      # one caller's query is a zero vector: InvalidVector for IT only
    got, errs = {}, {}

    def worker(t):
        for j in range(per):
            i = t * per + j
            try:
                got[i] = idx.search(queries[i], k if i % 3 else 3)   # mixed k per call
            except gfi.InvalidVector as e:
                errs[i] = e

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert list(errs) == [5 * per + 3]
    good = [i for i in range(T * per) if i != 5 * per + 3]
    exp = oracle.search_batch("cosine", rows, queries[good], [k if i % 3 else 3 for i in good], threads=8)
    for i, (eids, ed) in zip(good, exp):
        assert_topk_matches([a for a, _ in got[i]], [b for _, b in got[i]], eids, ed, ctx=f"coalesced call {i}")
    st = idx.stats()
    assert st["coalesced_requests"] >= 2 * st["coalesced_batches"] > 0, st   # batches really formed
    idx.set_option("coalesce", 0)
    assert idx.search(queries[0], 3) == got[0]       # call 0 asked for k = 3


def test_large_batch_is_chunked():
    """q above the tensor kernel's per-launch query limit (4096) is cut into chunks transparently."""
    n, d, q, k = 20000, 64, 4500, 5
    rows = oracle.gen_rows(91, 0, n, d, 0)
    queries = oracle.gen_rows(92, 0, q, d, 0)
    idx = build("dot", rows)
    ids, dist, cnt = idx.search_arrays(queries, k)
    assert np.all(cnt == k)
    pick = [0, 1, 4095, 4096, 4499]
    exp = oracle.search_batch("dot", rows, queries[pick], k, threads=4)
    for j, (eids, ed) in zip(pick, exp):
        assert_topk_matches(ids[j], dist[j], eids, ed, ctx=f"chunked q{j}")
    assert idx.stats()["tensor_queries"] == q


def test_bulk_ingest_from_reference_flat_file(tmp_path):
    """src/persistence/mmap.rs:13-15,161-172: [dim u32 LE][count u32 LE][f32 rows]."""
    import struct
    n, d = 5000, 48
    rows = oracle.gen_rows(95, 0, n, d, 1)
    path = tmp_path / "vectors.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<II", d, n))
        f.write(rows.tobytes())
    idx = gfi.GpuFlatIndex(DM.Cosine)
    assert idx.add_from_file(path, first_id=100) == n and idx.len() == n and idx.dim() == d
    assert np.array_equal(idx.get_vector(100 + 4999), rows[4999])
    q = oracle.gen_rows(96, 0, 3, d, 1)
    check_batch(idx, "cosine", rows, q, 10, ids=np.arange(100, 100 + n, dtype=np.uint64), ctx="flat file")
    other = gfi.GpuFlatIndex(DM.Cosine, dim=d + 1)
    with pytest.raises(gfi.DimensionMismatch):
        other.add_from_file(path)


def test_bulk_ingest_many_chunks_padded_rows_appends_overwrites_and_short_files(tmp_path):
    """The pipelined bulk path of gfi_add_from_file (two pinned buffers, parallel preads, one H2D + row_stats per
    64 MB chunk): a dimension that is not a multiple of 4 (rows are padded on the device: pitched copies), several
    chunks, a second file appended behind the first, a file whose ids overlap stored ones (FlatIndex::add overwrites,
    src/flat_index.rs:38-41: the staged path), and a file shorter than its header says."""
    import struct
    d, n1, n2 = 50, 700_000, 1000
    rows1 = oracle.gen_rows(97, 0, n1, d, 1)
    rows2 = oracle.gen_rows(98, 0, n2, d, 1)

    def write(name, rows, count=None):
        path = tmp_path / name
        with open(path, "wb") as f:
            f.write(struct.pack("<II", d, rows.shape[0] if count is None else count))
            f.write(rows.tobytes())
        return path

    idx = gfi.GpuFlatIndex(DM.Euclidean)
    assert idx.add_from_file(write("a.bin", rows1), first_id=0) == n1      # 140 MB: three chunks
    assert idx.add_from_file(write("b.bin", rows2), first_id=n1 + 10) == n2  # appended (ids stay ascending)
    assert idx.len() == n1 + n2 and idx.dim() == d
    for i in (0, 335_543, 335_544, 671_087, 671_088, n1 - 1):                                  # around the chunk boundaries
        assert np.array_equal(idx.get_vector(i), rows1[i])
    assert np.array_equal(idx.get_vector(n1 + 10 + 999), rows2[999])
    ids = np.concatenate([np.arange(n1, dtype=np.uint64), np.arange(n1 + 10, n1 + 10 + n2, dtype=np.uint64)])
    allrows = np.concatenate([rows1, rows2])
    q = oracle.gen_rows(99, 0, 4, d, 1)
    check_batch(idx, "euclidean", allrows, q, 10, ids=ids, ctx="bulk chunks")
    # overlapping ids: rows 5..14 are replaced by the first ten rows of the second file
    assert idx.add_from_file(write("c.bin", rows2[:10]), first_id=5) == 10
    assert idx.len() == n1 + n2
    allrows[5:15] = rows2[:10]
    assert np.array_equal(idx.get_vector(7), rows2[2])
    check_batch(idx, "euclidean", allrows, np.concatenate([q, rows2[3:4]]), 10, ids=ids, ctx="bulk overwrite")
    # a truncated file fails loudly and adds nothing
    with pytest.raises(gfi.IndexError_) as e:
        idx.add_from_file(write("short.bin", rows2[:100], count=5000), first_id=10_000_000)
    assert "shorter" in str(e.value) and idx.len() == n1 + n2
    check_batch(idx, "euclidean", allrows, q[:1], 10, ids=ids, ctx="after a failed load")


# ---------------------------------------------------------------- more full-size checks (BASELINE.json configs)
def test_full_size_c5_shard_l2_batch4096_against_oracle_subset():
    """C5 shard: 12.5M x 128 Euclidean, batch 4096, k = 10 (one GPU's share of the 100M index).  The
    whole batch runs on the tcgen05 path; 32 queries are checked against the oracle at full size, 16 against
    the independent exact-scan path, all for sortedness / idempotence."""
    n, d, q, k = 12_500_000, 128, 4096, 10
    idx = gfi.GpuFlatIndex(DM.Euclidean, dim=d)
    idx.reserve(n)
    idx.add_generated(7, 0, n, 0, 0)
    queries = oracle.gen_rows(8, 0, q, d, 0)
    ids, dist, cnt = idx.search_arrays(queries, k)
    assert np.all(cnt == k) and np.all(np.diff(dist, axis=1) >= 0) and np.all(dist >= 0)
    st = idx.stats()
    assert st["tensor_queries"] == q and st["fallback_queries"] <= 16, st
    idx.set_option("tensor_min_q", 1 << 30)
    ids_s, dist_s, _ = idx.search_arrays(queries[:16], k)
    assert np.array_equal(ids_s, ids[:16]) and np.array_equal(dist_s, dist[:16])
    pick = np.arange(0, q, q // 32)[:32]
    exp = oracle.search_generated("euclidean", 7, 0, n, d, 0, queries[pick], k)
    for j, (eids, ed) in zip(pick, exp):
        assert_topk_matches(ids[j], dist[j], eids, ed, ctx=f"C5 full q{j}")


def test_full_size_c3_c4_properties():
    """C3 (10M x 768 dot, k=100) and C4 (10M x 384 Euclidean + filter, k=10) at full size through
    size-independent properties: a stored row queried against itself comes back first (L2), the
    merge of the answers over a mask and its complement equals the unfiltered answer, sortedness,
    and agreement of the scan and tensor paths."""
    from vectordb_from_scratch_b200 import synth
    # ---- C4
    n, d, k = 10_000_000, 384, 10
    idx = gfi.GpuFlatIndex(DM.Euclidean, dim=d)
    idx.reserve(n)
    idx.add_generated(6, 0, n, 0, 0)
    probe = [0, 1234567, n - 1]
    qs = np.concatenate([synth.gen_rows(6, r, 1, d, 0) for r in probe] + [synth.gen_rows(66, 0, 3, d, 0)])
    ids, dist, cnt = idx.search_arrays(qs, k)
    assert np.all(cnt == k) and np.all(np.diff(dist, axis=1) >= 0)
    for j, r in enumerate(probe):
        assert ids[j, 0] == r and dist[j, 0] == 0.0
    rng = np.random.default_rng(4)
    m = rng.random(n) < 0.5
    a_ids, a_d, a_c = idx.search_arrays(qs, k, mask=m)
    b_ids, b_d, b_c = idx.search_arrays(qs, k, mask=~m)
    assert np.all(m[a_ids.astype(np.int64)]) and not np.any(m[b_ids.astype(np.int64)])
    merged = numpy_merge(np.stack([a_ids, b_ids]), np.stack([a_d, b_d]), np.stack([a_c, b_c]), [k] * len(qs))
    for i in range(len(qs)):
        assert [p[1] for p in merged[i]] == [int(x) for x in ids[i]]
        assert np.array_equal(np.array([p[0] for p in merged[i]], np.float32), dist[i])
    sparse = rng.random(n) < 0.01
    s_ids, s_d, s_c = idx.search_arrays(qs, k, mask=sparse)
    assert np.all(s_c == k) and np.all(sparse[s_ids.astype(np.int64)]) and np.all(np.diff(s_d, axis=1) >= 0)
    assert np.all(s_d[:, 0] >= dist[:, 0])
    idx.close()
    # ---- C3
    n, d, k = 10_000_000, 768, 100
    idx = gfi.GpuFlatIndex(DM.DotProduct, dim=d)
    idx.reserve(n)
    idx.add_generated(5, 0, n, 1, 0)
    qs = synth.gen_rows(55, 0, 64, d, 1)
    t_ids, t_d, t_c = idx.search_arrays(qs, k)            # batch 64: tcgen05 path (C3b)
    assert idx.stats()["tensor_queries"] == 64 and np.all(t_c == k) and np.all(np.diff(t_d, axis=1) >= 0)
    a_ids, a_d, a_c = idx.search_arrays(qs[:2], k)        # small batch on a large index: the cost model picks
    st = idx.stats()                                      # the tensor path too (fp16 rows are half the bytes)
    assert st["tensor_queries"] == 66 and st["scan_queries"] == 0, st
    assert np.array_equal(a_ids, t_ids[:2]) and np.array_equal(a_d, t_d[:2])
    idx.set_option("tensor_auto", 0)
    s_ids, s_d, s_c = idx.search_arrays(qs[:2], k)        # forced onto the fp32 scan path (C3a as specified)
    assert idx.stats()["scan_queries"] == 2
    assert np.array_equal(s_ids, t_ids[:2]) and np.array_equal(s_d, t_d[:2])


def test_full_size_c3_dot_k100_against_the_oracle():
    """C3a / C3b at full size against the ORACLE: 10M x 768 dot, k = 100, bench.py's own seeds.
    Eight queries of the batch-64 answer (tcgen05 path, 16 384-key select staging), the same eight as single queries
    through the cost-model route, and two on the fp32 scan (tensor_auto = 0)."""
    n, d, k = 10_000_000, 768, 100
    idx = gfi.GpuFlatIndex(DM.DotProduct, dim=d)
    idx.reserve(n)
    idx.add_generated(5, 0, n, 1, 0)
    qs = oracle.gen_rows(6, 0, 64, d, 1)
    t_ids, t_d, t_c = idx.search_arrays(qs, k)
    assert idx.stats()["tensor_queries"] == 64 and np.all(t_c == k)
    pick = np.arange(0, 64, 8)
    exp = oracle.search_generated("dot", 5, 0, n, d, 1, qs[pick], k)
    for j, (eids, ed) in zip(pick, exp):
        assert_topk_matches(t_ids[j], t_d[j], eids, ed, ctx=f"C3b q{j}")
    for j, (eids, ed) in zip(pick, exp):  # C3a: one query per call
        a_ids, a_d, a_c = idx.search_arrays(qs[j:j + 1], k)
        assert_topk_matches(a_ids[0], a_d[0], eids, ed, ctx=f"C3a routed q{j}")
    assert idx.stats()["scan_queries"] == 0
    idx.set_option("tensor_auto", 0)
    for j, (eids, ed) in list(zip(pick, exp))[:2]:
        s_ids, s_d, s_c = idx.search_arrays(qs[j:j + 1], k)
        assert_topk_matches(s_ids[0], s_d[0], eids, ed, ctx=f"C3a scan q{j}")
    st = idx.stats()
    assert st["scan_queries"] == 2 and st["fallback_queries"] <= 2 and st["paged_queries"] == 0, st


@pytest.mark.parametrize("pct", [1, 50])
def test_full_size_c4_filtered_against_the_oracle(pct):
    """C4 at full size against the ORACLE over the eligible rows: 10M x 384 Euclidean, eq filter at 1 % / 50 %,
    k = 10, eight queries, through all three reference-facing forms: the eligibility bitmask (filter push-down), the
    device-evaluated metadata filter (gfi_search_filtered over a resident column), and the reference's own
    post-filter (search 3k unfiltered, keep the matching hits: storage.rs:249-290)."""
    import bench
    n, d, k = 10_000_000, 384, 10
    idx = gfi.GpuFlatIndex(DM.Euclidean, dim=d)
    idx.reserve(n)
    idx.add_generated(6, 0, n, 0, 0)
    elig = bench.eligible_rows(0, n, pct)
    qs = oracle.gen_rows(7, 0, 8, d, 0)
    exp = oracle.search_generated("euclidean", 6, 0, n, d, 0, qs, k, eligible=elig)
    m_ids, m_d, m_c = idx.search_arrays(qs, k, mask=elig)              # 8 queries in one call
    for i, (eids, ed) in enumerate(exp):
        assert m_c[i] == k
        assert_topk_matches(m_ids[i], m_d[i], eids, ed, ctx=f"C4 {pct}% mask q{i}")
    idx.set_metadata_column("tag", np.arange(n, dtype=np.uint64), ["hit", "miss"], np.where(elig, 0, 1).astype(np.uint32))
    flt = {"op": "eq", "field": "tag", "value": "hit"}
    for i, (eids, ed) in enumerate(exp):                               # single queries, as the bench issues them
        f_ids, f_d, f_c = idx.search_filtered(qs[i:i + 1], k, flt)
        assert f_c[0] == k
        assert_topk_matches(f_ids[0], f_d[0], eids, ed, ctx=f"C4 {pct}% filter q{i}")
    # reference semantics: unfiltered top-3k, then the filter (may return fewer than k; at 1 % usually none)
    full = oracle.search_generated("euclidean", 6, 0, n, d, 0, qs[:4], 3 * k)
    u_ids, u_d, u_c = idx.search_arrays(qs[:4], 3 * k)
    for i, (eids, ed) in enumerate(full):
        assert_topk_matches(u_ids[i], u_d[i], eids, ed, ctx=f"C4 unfiltered 3k q{i}")
        keep_g = elig[u_ids[i].astype(np.int64)]
        keep_o = elig[eids.astype(np.int64)]
        assert np.array_equal(u_ids[i][keep_g][:k], eids[keep_o][:k])


def test_full_size_cost_model_route_switches_off_on_uncertifiable_data():
    """A 1.6 GB index whose top-k boundary is always an exact tie (64 copies of every vector): single queries first
    take the tensor pass by the cost model, every one of them falls back to the exact scan, and after 32 such
    queries the index stops using the route.  Answers are exact throughout."""
    base_n, copies, d, k = 6250, 64, 1024, 10
    base = oracle.gen_rows(171, 0, base_n, d, 1)
    rows = np.tile(base, (copies, 1))                      # row i is base[i % base_n]
    idx = gfi.GpuFlatIndex(DM.Euclidean, dim=d)
    idx.add_batch(np.arange(rows.shape[0], dtype=np.uint64), rows)
    queries = oracle.gen_rows(172, 0, 48, d, 1)
    exp = oracle.search_batch("euclidean", base, queries, 1)   # nearest distinct vector of every query
    for i in range(48):
        ids, dist, cnt = idx.search_arrays(queries[i:i + 1], k)
        b = int(exp[i][0][0])
        assert [int(x) for x in ids[0]] == [b + j * base_n for j in range(k)]      # ties: lower id first
        assert np.all(dist[0] == exp[i][1][0])
    st = idx.stats()
    assert st["fallback_queries"] >= 32 and st["scan_queries"] >= 8, st            # route taken, then dropped
    assert st["tensor_queries"] < 48, st


# ---------------------------------------------------------------- device-side MetadataFilter evaluation (N2)
def test_device_side_filter_matches_host_truth_table():
    """Filters in the reference's JSON form evaluated on the GPU (csrc/filter.cu) against the host truth
    table (storage.rs:62-70) and the oracle run on the matching subset."""
    n, d, k = 6000, 40, 10
    rows = oracle.gen_rows(97, 0, n, d, 1)
    store = gfi.VectorStore(DM.Euclidean)
    colors = ["red", "green", "blue", "c\"q"]
    for i in range(n):
        md = gfi.Metadata()
        if i % 7 != 0:
            md.insert("color", colors[(i * 2654435761 >> 3) % 4])
        if i % 3 == 0:
            md.insert("size", "s%d" % (i % 2))
        store.insert_with_metadata("v%d" % i, rows[i], md)
    F = gfi.MetadataFilter
    filters = [
        F.eq("color", "red"), F.ne("color", "red"), F.exists("size"), F.eq("color", "nope"), F.eq("nofield", "x"),
        F.ne("nofield", "x"), F.and_([]), F.or_([]), F.eq("color", "c\"q"),
        F.and_([F.eq("color", "blue"), F.eq("size", "s0")]),
        F.or_([F.eq("color", "green"), F.and_([F.exists("size"), F.ne("color", "red")])]),
        F.and_([F.or_([F.eq("size", "s1"), F.eq("size", "s0")]), F.ne("size", "s1"), F.exists("color")]),
    ]
    queries = oracle.gen_rows(98, 0, 3, d, 1)
    for flt in filters:
        host_mask = store.filter_mask(flt)
        ids, dist, cnt = store.index.search_filtered(queries, k, flt.to_json())
        exp = oracle.search_batch("euclidean", rows, queries, k, eligible=host_mask, threads=4)
        for i, (eids, ed) in enumerate(exp):
            assert cnt[i] == len(eids), (flt.to_json(), cnt[i], len(eids))
            assert_topk_matches(ids[i, :cnt[i]], dist[i, :cnt[i]], eids, ed, ctx=str(flt.to_json()))
    # overwrite replaces metadata wholesale; delete removes the row from filtered results
    md = gfi.Metadata()
    md.insert("size", "huge")
    store.insert_with_metadata("v5", rows[5], md)
    store.delete("v10")
    res = store.search_with_filter(rows[5], 3, F.eq("size", "huge"), pushdown="device")
    assert [r.id for r in res] == ["v5"] and res[0].distance == 0.0
    assert store.search_with_filter(rows[5], 3, F.and_([F.eq("size", "huge"), F.exists("color")]), pushdown="device") == []
    with pytest.raises(gfi.IndexError_):
        store.index.search_filtered(queries, k, '{"op": "xor", "filters": []}')


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot"])
def test_k_above_the_kernels_list_capacity(metric):
    """FlatIndex::search accepts any k (src/flat_index.rs:63).  Beyond 1016 the library answers in exact passes over
    the rows not returned yet; the concatenation is the exact ascending top-k."""
    n, d = 5000, 24
    rows = oracle.gen_rows(99, 0, n, d, 1)
    idx = build(metric, rows, ids=np.arange(n, dtype=np.uint64) * 5 + 3)   # sparse ids: the passes mask by slot
    ids = np.arange(n, dtype=np.uint64) * 5 + 3
    queries = oracle.gen_rows(100, 0, 4, d, 1)
    ks = [1500, 10, 2500, 1017]                      # mixed batch: small and large k together
    got_ids, got_d, cnt = idx.search_arrays(queries, np.array(ks, dtype=np.uint32))
    exp = oracle.search_batch(metric, rows, queries, ks, ids=ids)
    for i, (eids, ed) in enumerate(exp):
        assert cnt[i] == ks[i]
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"big k q{i}")
    # k larger than the index: everything, once, in order
    idx.remove(int(ids[7]))
    g2, d2, c2 = idx.search_arrays(queries[:1], 9000)
    live = np.ones(n, dtype=bool)
    live[7] = False
    eids, ed = oracle.search_batch(metric, rows[live], queries[:1], 9000, ids=ids[live])[0]
    assert c2[0] == n - 1
    assert_topk_matches(g2[0, :c2[0]], d2[0, :c2[0]], eids, ed, ctx="k > n")
    # with a caller mask (by internal id) the passes start from the caller's eligible rows
    elig = (np.arange(n) % 3 != 1) & live
    by_id = np.zeros(int(ids[-1]) + 1, dtype=bool)
    by_id[ids[elig]] = True
    g3, d3, c3 = idx.search_arrays(queries[:2], np.array([2200, 12], dtype=np.uint32), mask=by_id)
    exp = oracle.search_batch(metric, rows[elig], queries[:2], [2200, 12], ids=ids[elig])
    for i, (eids, ed) in enumerate(exp):
        assert c3[i] == len(eids)
        assert_topk_matches(g3[i, :c3[i]], d3[i, :c3[i]], eids, ed, ctx=f"big k + mask q{i}")
    # a filter handed down as JSON: evaluated once on the host mirror of the metadata columns, then the same passes
    for i in range(0, n, 2):
        idx.set_metadata(int(ids[i]), {"color": "red" if i % 4 == 0 else "blue"})
    has_color = (np.arange(n) % 2 == 0) & live
    red = (np.arange(n) % 4 == 0) & live
    cases = [('{"op": "exists", "field": "color"}', has_color, [1500, 7]),
             ('{"op": "ne", "field": "color", "value": "red"}', live & ~red, [20, 3000]),
             ('{"op": "eq", "field": "color", "value": "red"}', red, [2000, 1100])]      # 1250 red rows: k > matches
    for flt, elig, kk in cases:
        g4, d4, c4 = idx.search_filtered(queries[:2], np.array(kk, dtype=np.uint32), flt)
        exp = oracle.search_batch(metric, rows[elig], queries[:2], kk, ids=ids[elig])
        for i, (eids, ed) in enumerate(exp):
            assert c4[i] == len(eids), (flt, c4[i], len(eids))
            assert_topk_matches(g4[i, :c4[i]], d4[i, :c4[i]], eids, ed, ctx=f"big k + filter {flt} q{i}")
