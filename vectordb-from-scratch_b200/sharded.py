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


def packed_layout(q, kmax):
    """Byte offsets of one shard's packed result block [ids u64 | dist f32 | counts u32] and its size
    (rounded up to 8 bytes so that shard blocks stay 8-byte aligned after an all-gather)."""
    off_d = q * kmax * 8
    off_c = off_d + q * kmax * 4
    size = (off_c + q * 4 + 7) // 8 * 8
    return off_d, off_c, size


def packed_views(buf, q, kmax):
    """(ids [.., q, kmax] int64, dist [.., q, kmax] f32, counts [.., q] int32) views of a packed block
    (uint8 [size]) or of G gathered blocks (uint8 [G, size]); no copies."""
    off_d, off_c, size = packed_layout(q, kmax)
    lead = tuple(buf.shape[:-1])
    ids = buf[..., :off_d].view(torch.int64).unflatten(-1, (q, kmax))
    d = buf[..., off_d:off_c].view(torch.float32).unflatten(-1, (q, kmax))
    cnt = buf[..., off_c:off_c + q * 4].view(torch.int32)
    assert ids.shape == lead + (q, kmax)
    return ids, d, cnt


class ShardedSearch:
    """local_search(queries, ks) -> (ids [q,kmax] int64, dist [q,kmax] f32, counts [q] int32) tensors
    on this rank's device; merge(all_ids [G,q,kmax], all_dist, all_counts [G,q], ks) -> same triple.

    packed=True: local_search additionally returns, as a 4th element, the uint8 tensor (packed_layout) that its
    three outputs are views of; the exchange is then ONE all-gather per search, and merge receives strided views
    of the gathered [G, size] buffer (gfi_merge_topk_device_strided takes size as the shard stride)."""

    def __init__(self, local_search, merge, group=None, packed=False):
        self.local_search = local_search
        self.merge = merge
        self.group = group
        self.packed = packed
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def search(self, queries, ks):
        res = self.local_search(queries, ks)
        ids, d, cnt = res[0], res[1], res[2]
        if self.world == 1:
            return ids, d, cnt
        G = self.world
        if self.packed:
            pack = res[3]
            out = torch.empty((G * pack.shape[0],), dtype=torch.uint8, device=pack.device)
            dist.all_gather_into_tensor(out, pack, group=self.group)
            all_ids, all_d, all_c = packed_views(out.view(G, pack.shape[0]), ids.shape[0], ids.shape[1])
            return self.merge(all_ids, all_d, all_c, ks)
        # concatenated along dim 0 (accepted by both NCCL and gloo), viewed as [G, q, ...] afterwards
        def gather(t):
            t = t.contiguous()
            out = torch.empty((G * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t, group=self.group)
            return out.view((G,) + tuple(t.shape))
        all_ids, all_d, all_c = gather(ids), gather(d), gather(cnt)
        return self.merge(all_ids, all_d, all_c, ks)
