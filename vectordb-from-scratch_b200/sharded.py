"""Row-wise sharding of one index over the GPUs of a box: one process per GPU.

The database shards by contiguous internal-id ranges (shard g owns ids
[g*ceil(n/G), (g+1)*ceil(n/G)) so "lower id wins" stays a pure key comparison); queries are
replicated; every rank searches its shard (no data-path collective), then ONE all-gather of
the per-rank (ids, distances, counts) over NCCL/NVLink and a merge of G sorted lists per
query (CUDA merge kernel, gfi_merge_topk_device).  SURVEY.md section 8(e) G1.

The collective plumbing is backend-agnostic so the same code is covered on CPU with gloo
(world_size 2) in tests/test_sharded_gloo.py, with the local search and the merge injected.
"""
import torch
import torch.distributed as dist


def shard_range(n, world, rank):
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


class ShardedSearch:
    """local_search(queries, ks) -> (ids [q,kmax] int64, dist [q,kmax] f32, counts [q] int32) tensors
    on this rank's device; merge(all_ids [G,q,kmax], all_dist, all_counts [G,q], ks) -> same triple."""

    def __init__(self, local_search, merge, group=None):
        self.local_search = local_search
        self.merge = merge
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def search(self, queries, ks):
        ids, d, cnt = self.local_search(queries, ks)
        if self.world == 1:
            return ids, d, cnt
        G = self.world
        # concatenated along dim 0 (accepted by both NCCL and gloo), viewed as [G, q, ...] afterwards
        def gather(t):
            t = t.contiguous()
            out = torch.empty((G * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t, group=self.group)
            return out.view((G,) + tuple(t.shape))
        all_ids, all_d, all_c = gather(ids), gather(d), gather(cnt)
        return self.merge(all_ids, all_d, all_c, ks)
