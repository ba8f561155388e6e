"""CPU, world_size 2, gloo: the multi-GPU plumbing of vectordb-from-scratch_b200/sharded.py --
row-range sharding, the single all-gather of per-rank (ids, distances, counts) and the
[G][q][k] layout handed to the merge -- with the local search and the merge injected (the oracle
and a numpy merge here; libgfi's CUDA kernels on a GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import numpy_merge
    from vectordb_from_scratch_b200.sharded import ShardedSearch, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d, q, k = 3001, 24, 6, 10
    rows = oracle.gen_rows(71, 0, n, d, 1)
    queries = oracle.gen_rows(72, 0, q, d, 1)
    lo, hi = shard_range(n, world, rank)
    ks = [10, 3, 10, 1, 7, 10]

    def local_search(qs, kk):
        res = oracle.search_batch("euclidean", rows[lo:hi], qs.numpy(), kk, ids=np.arange(lo, hi, dtype=np.uint64))
        ids = torch.zeros((q, k), dtype=torch.int64)
        dd = torch.zeros((q, k), dtype=torch.float32)
        cnt = torch.zeros((q,), dtype=torch.int32)
        for i, (a, b) in enumerate(res):
            ids[i, :len(a)] = torch.from_numpy(a.astype(np.int64))
            dd[i, :len(b)] = torch.from_numpy(b)
            cnt[i] = len(a)
        return ids, dd, cnt

    def merge(all_ids, all_d, all_c, kk):
        assert tuple(all_ids.shape) == (world, q, k) and tuple(all_c.shape) == (world, q)
        m = numpy_merge(all_ids.numpy(), all_d.numpy(), all_c.numpy(), kk)
        return m, None, None

    merged, _, _ = ShardedSearch(local_search, merge).search(torch.from_numpy(queries), ks)
    exp = oracle.search_batch("euclidean", rows, queries, ks)

    def matches(m):
        return all([p[1] for p in m[i]] == [int(x) for x in exp[i][0]] and
                   np.array_equal(np.array([p[0] for p in m[i]], np.float32), exp[i][1]) for i in range(q))

    ok = matches(merged)

    # packed exchange: the three outputs are views of one byte block, ONE all-gather, strided views into the merge
    from vectordb_from_scratch_b200.sharded import packed_layout, packed_views

    def local_search_packed(qs, kk):
        ids, dd, cnt = local_search(qs, kk)
        pack = torch.zeros((packed_layout(q, k)[2],), dtype=torch.uint8)
        pi, pd, pc = packed_views(pack, q, k)
        pi.copy_(ids); pd.copy_(dd); pc.copy_(cnt)
        return pi, pd, pc, pack

    def merge_packed(all_ids, all_d, all_c, kk):
        assert tuple(all_ids.shape) == (world, q, k) and tuple(all_c.shape) == (world, q)
        assert all_ids.stride(0) * 8 == packed_layout(q, k)[2] == all_c.stride(0) * 4  # the shard stride
        return numpy_merge(all_ids.numpy(), all_d.numpy(), all_c.numpy(), kk), None, None

    merged2, _, _ = ShardedSearch(local_search_packed, merge_packed, packed=True).search(torch.from_numpy(queries), ks)
    ok = ok and matches(merged2)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_search_world2_gloo():
    from vectordb_from_scratch_b200.sharded import shard_range
    assert [shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_range(3, 8, 7) == (3, 3)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True
