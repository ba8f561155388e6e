"""One index sharded row-wise over several GPUs INSIDE libgfi (gfi_create_sharded), through the C ABI, against the
CPU oracle over the WHOLE index.  The reference server owns one index behind one lock (src/server/mod.rs:13-16), so
every `trait Index` method (src/index.rs:11-35) has to work on the sharded handle exactly as on a single-GPU one.

Every case runs twice: several shards on GPU 0 (always; exercises routing, fan-out, gather block, merge kernel) and one
shard per GPU on every GPU of the box (skipped on a 1-GPU box; adds the peer stores over NVLink)."""
import numpy as np
import pytest

import oracle
import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200 import DistanceMetric as DM
from helpers import assert_topk_matches

pytestmark = pytest.mark.gpu
M = {"euclidean": DM.Euclidean, "cosine": DM.Cosine, "dot": DM.DotProduct}


def gpu_count():
    import torch
    return torch.cuda.device_count()


def layouts():
    out = [pytest.param([0, 0, 0], id="3shards-gpu0")]
    out.append(pytest.param("all", id="one-shard-per-gpu"))
    return out


def devices_of(layout):
    if layout == "all":
        n = gpu_count()
        if n < 2:
            pytest.skip("needs 2+ GPUs")
        return list(range(min(n, 8)))
    return layout


def check(idx, metric, rows, queries, ks, ids=None, eligible=None, mask=None, ctx=""):
    got_ids, got_d, cnt = idx.search_arrays(queries, ks, mask=mask)
    exp = oracle.search_batch(metric, rows, queries, ks, ids=ids, eligible=eligible, threads=8)
    for i, (eids, ed) in enumerate(exp):
        assert cnt[i] == len(eids), f"{ctx} q{i}: count {cnt[i]} != {len(eids)}"
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx=f"{ctx} q{i}")


@pytest.mark.parametrize("layout", layouts())
@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot"])
def test_sharded_search_matches_oracle_scan_and_tensor_paths(layout, metric):
    devs = devices_of(layout)
    n, d = 40_000, 96
    rows = oracle.gen_rows(61, 0, n, d, 1)
    idx = gfi.GpuFlatIndex(M[metric], devices=devs)
    idx.set_option("shard_block", 1024)  # ids interleave over the shards in blocks of 1024
    idx.set_option("tensor_min_rows", 256)  # (5000 rows per shard on an 8-GPU box: keep the tcgen05 path in play)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    assert idx.len() == n and idx.dim() == d
    idx.flush()
    st = idx.stats()
    assert st["shards"] == len(devs) and st["n_live"] == n
    for q, k in ((1, 10), (3, 100), (64, 10), (130, 7)):  # scan path (q < 16) and tcgen05 path
        queries = oracle.gen_rows(62 + q, 0, q, d, 1)
        check(idx, metric, rows, queries, k, ctx=f"{metric} q={q} k={k}")
    # per-query k, k = 0, k > n (flat_index.rs:63 truncate)
    queries = oracle.gen_rows(70, 0, 4, d, 1)
    check(idx, metric, rows, queries, np.array([1, 0, 17, 5], dtype=np.uint32), ctx="per-query k")
    st = idx.stats()
    assert st["scan_queries"] > 0 and st["tensor_queries"] > 0


@pytest.mark.parametrize("layout", layouts())
def test_sharded_contiguous_ranges_after_reserve_generated_rows_and_k_above_a_shard(layout):
    devs = devices_of(layout)
    n, d, k = 30_000, 64, 50
    idx = gfi.GpuFlatIndex(DM.Euclidean, dim=d, devices=devs)
    idx.reserve(n)  # empty index: contiguous id ranges, ceil(n / G) rounded up to 256 per shard
    idx.add_generated(9, 0, n, 0, 0)
    rows = oracle.gen_rows(9, 0, n, d, 0)
    assert idx.len() == n
    queries = oracle.gen_rows(10, 0, 20, d, 0)
    check(idx, "euclidean", rows, queries, k, ctx="reserve+generated")
    # a tiny index: most shards hold fewer than k rows, some none at all
    small = gfi.GpuFlatIndex(DM.Euclidean, devices=devs)
    small.set_option("shard_block", 4)
    small.add_batch(np.arange(10, dtype=np.uint64), rows[:10])
    check(small, "euclidean", rows[:10], queries[:3], 8, ctx="tiny")
    check(small, "euclidean", rows[:10], queries[:3], 64, ctx="tiny k>n")
    empty = gfi.GpuFlatIndex(DM.Euclidean, devices=devs)
    ids, dist, cnt = empty.search_arrays(queries[:2], 5)
    assert cnt.tolist() == [0, 0]


@pytest.mark.parametrize("layout", layouts())
def test_sharded_mutations_get_vector_masks_and_filters(layout):
    devs = devices_of(layout)
    n, d, k = 12_000, 48, 12
    rows = oracle.gen_rows(81, 0, n, d, 1).copy()
    ids = np.arange(n, dtype=np.uint64) * 3 + 5  # gaps in the id space
    idx = gfi.GpuFlatIndex(DM.Cosine, devices=devs)
    idx.set_option("shard_block", 512)
    idx.add_batch(ids, rows)
    # remove (idempotent) and overwrite (HashMap::insert semantics, flat_index.rs:38-41)
    live = np.ones(n, dtype=bool)
    for j in (0, 17, 511, 512, 513, 4000, n - 1):
        idx.remove(int(ids[j]))
        idx.remove(int(ids[j]))
        live[j] = False
    newrow = oracle.gen_rows(82, 0, 3, d, 1)
    for t, j in enumerate((3, 2048, 9999)):
        idx.add(int(ids[j]), newrow[t])
        rows[j] = newrow[t]
    assert idx.len() == int(live.sum())
    assert np.array_equal(idx.get_vector(int(ids[2048])), newrow[1])
    assert idx.get_vector(int(ids[17])) is None and idx.get_vector(4) is None
    queries = oracle.gen_rows(83, 0, 6, d, 1)
    check(idx, "cosine", rows[live], queries, k, ids=ids[live], ctx="after mutations")
    big = oracle.gen_rows(84, 0, 40, d, 1)
    check(idx, "cosine", rows[live], big, k, ids=ids[live], ctx="after mutations, tensor path")
    # eligibility bitmask by internal id (filter push-down), 10 % and 60 %
    for pct in (10, 60):
        elig_id = np.zeros(int(ids.max()) + 1, dtype=bool)
        pick = (np.arange(n) * 7919 % 100) < pct
        elig_id[ids[pick]] = True
        check(idx, "cosine", rows[live], queries, k, ids=ids[live], eligible=pick[live], mask=elig_id, ctx=f"mask {pct}%")
    # device-side metadata filter: columns live next to each shard's rows
    for j in range(0, n, 2):
        if live[j]:
            idx.set_metadata(int(ids[j]), {"lang": "en" if j % 4 == 0 else "de"})
    flt = {"op": "eq", "field": "lang", "value": "en"}
    got_ids, got_d, cnt = idx.search_filtered(queries, k, flt)
    elig = live & (np.arange(n) % 4 == 0)
    exp = oracle.search_batch("cosine", rows, queries, k, ids=ids, eligible=elig, threads=4)
    for i, (eids, ed) in enumerate(exp):
        assert cnt[i] == len(eids)
        assert_topk_matches(got_ids[i, :cnt[i]], got_d[i, :cnt[i]], eids, ed, ctx="filter")
    idx.compact()
    check(idx, "cosine", rows[live], queries, k, ids=ids[live], ctx="after compact")
    # exact distances of explicit pairs (HNSW candidate evaluation), ids owned by different shards, one absent
    cand = np.stack([ids[[1, 600, 1100, 5000, 17]] for _ in range(2)])
    dist, status = idx.distances(queries[:2], cand)
    for i in range(2):
        for j, row in enumerate((1, 600, 1100, 5000)):
            assert status[i, j] == 0 and dist[i, j] == oracle.distance("cosine", queries[i], rows[row])
        assert status[i, 4] == 1


@pytest.mark.parametrize("layout", layouts())
def test_sharded_error_semantics(layout):
    devs = devices_of(layout)
    n, d = 5000, 32
    rows = oracle.gen_rows(91, 0, n, d, 1).copy()
    rows[4321] = 0.0  # a zero row in ONE shard fails a cosine search of the whole index (distance.rs:60-64)
    idx = gfi.GpuFlatIndex(DM.Cosine, devices=devs)
    idx.set_option("shard_block", 256)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    queries = oracle.gen_rows(92, 0, 3, d, 1)
    with pytest.raises(gfi.InvalidVector):
        idx.search_arrays(queries, 5)
    with pytest.raises(gfi.InvalidVector):
        idx.search_arrays(oracle.gen_rows(92, 0, 40, d, 1), 5)  # tensor path
    idx.remove(4321)
    check(idx, "cosine", np.delete(rows, 4321, axis=0), queries, 5, ids=np.delete(np.arange(n, dtype=np.uint64), 4321))
    with pytest.raises(gfi.DimensionMismatch) as e:
        idx.search_arrays(np.zeros((1, d + 1), dtype=np.float32), 5)
    assert (e.value.expected, e.value.actual) == (d, d + 1) or (e.value.expected, e.value.actual) == (d + 1, d)
    nanq = queries.copy()
    nanq[1, 3] = np.nan
    with pytest.raises(gfi.NaNDistance):
        idx.search_arrays(nanq, 5)
    check(idx, "cosine", np.delete(rows, 4321, axis=0), queries, 5, ids=np.delete(np.arange(n, dtype=np.uint64), 4321),
          ctx="after errors")


@pytest.mark.parametrize("layout", layouts())
def test_sharded_k_above_the_kernels_list_capacity(layout):
    devs = devices_of(layout)
    n, d, k = 9000, 16, 3000
    rows = oracle.gen_rows(95, 0, n, d, 0)
    idx = gfi.GpuFlatIndex(DM.Euclidean, devices=devs)
    idx.set_option("shard_block", 128)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    queries = oracle.gen_rows(96, 0, 2, d, 0)
    check(idx, "euclidean", rows, queries, np.array([k, 20], dtype=np.uint32), ctx="big k")
    elig = np.arange(n) % 4 != 2   # caller mask by internal id: every shard's passes start from its eligible rows
    check(idx, "euclidean", rows, queries, np.array([k, 20], dtype=np.uint32), eligible=elig, mask=elig, ctx="big k + mask")
    # and with a JSON filter: every shard evaluates it over its own metadata columns before its passes
    for i in range(0, n, 2):
        idx.set_metadata(i, {"tag": "a"})
    tagged = np.arange(n) % 2 == 0
    ks = [k, 20]
    g, dd, c = idx.search_filtered(queries, np.array(ks, dtype=np.uint32), '{"op": "exists", "field": "tag"}')
    exp = oracle.search_batch("euclidean", rows[tagged], queries, ks, ids=np.arange(n, dtype=np.uint64)[tagged])
    for i, (eids, ed) in enumerate(exp):
        assert c[i] == len(eids)
        assert_topk_matches(g[i, :c[i]], dd[i, :c[i]], eids, ed, ctx=f"big k + filter q{i}")


@pytest.mark.parametrize("layout", layouts())
def test_sharded_device_resident_searches_double_buffer_and_keep_flags(layout):
    """gfi_search_device on a sharded handle: queries/results in the ROOT GPU's memory, several searches enqueued
    back to back (the gather block is double-buffered), one status collection at the end."""
    import torch
    devs = devices_of(layout)
    n, d, q, k = 50_000, 64, 48, 10
    rows = oracle.gen_rows(101, 0, n, d, 1)
    idx = gfi.GpuFlatIndex(DM.DotProduct, dim=d, devices=devs)
    idx.reserve(n)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    dev = torch.device("cuda", devs[0])
    torch.cuda.set_device(dev)
    tss = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    outs = []
    dks = torch.full((q,), k, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    for it in range(7):  # batches alternate between two caller streams: merge of one overlaps the next main pass
        ts = tss[it % 2]
        queries = oracle.gen_rows(110 + it, 0, q, d, 1)
        with torch.cuda.stream(ts):
            dq = torch.from_numpy(queries).to(dev, non_blocking=False)
            o_ids = torch.zeros((q, k), dtype=torch.int64, device=dev)
            o_d = torch.zeros((q, k), dtype=torch.float32, device=dev)
            o_c = torch.zeros((q,), dtype=torch.int32, device=dev)
        idx.search_device(dq.data_ptr(), q, dks.data_ptr(), k, o_ids.data_ptr(), o_d.data_ptr(), o_c.data_ptr(), k,
                          stream=ts.cuda_stream)
        outs.append((queries, dq, o_ids, o_d, o_c))
    idx.search_status()
    torch.cuda.synchronize(dev)
    for queries, _, o_ids, o_d, o_c in outs:
        exp = oracle.search_batch("dot", rows, queries, k, threads=8)
        gi, gd, gc = o_ids.cpu().numpy().astype(np.uint64), o_d.cpu().numpy(), o_c.cpu().numpy()
        for i, (eids, ed) in enumerate(exp):
            assert gc[i] == k
            assert_topk_matches(gi[i], gd[i], eids, ed, ctx=f"device q{i}")


@pytest.mark.parametrize("layout", layouts())
def test_sharded_concurrent_searches_from_many_threads(layout):
    import threading
    devs = devices_of(layout)
    n, d, k = 20_000, 40, 9
    rows = oracle.gen_rows(121, 0, n, d, 0)
    idx = gfi.GpuFlatIndex(DM.Euclidean, devices=devs)
    idx.set_option("shard_block", 2048)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    errs = []

    def worker(t):
        try:
            for it in range(6):
                q = 1 + (t + it) % 5 if it % 2 == 0 else 20 + t
                queries = oracle.gen_rows(130 + 10 * t + it, 0, q, d, 0)
                check(idx, "euclidean", rows, queries, k, ctx=f"thread {t} it {it}")
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs[0]
