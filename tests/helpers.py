"""Shared helpers for the parity tests (checker side only)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))

REL_TOL = 1e-5  # north_star: adjacent distances closer than this (relative) may swap


def assert_topk_matches(got_ids, got_dist, exp_ids, exp_dist, rel_tol=REL_TOL, exact_dist=True, ctx=""):
    """The parity rule of BASELINE.json's north_star: identical id order under the (distance,
    lower id) tie-break, except where adjacent distances differ by less than rel_tol; distances
    within rel_tol.  With exact_dist the distances must additionally be bit-identical wherever the
    ids agree (libgfi re-scores survivors with the reference's exact arithmetic)."""
    got_ids, exp_ids = np.asarray(got_ids), np.asarray(exp_ids)
    got_dist, exp_dist = np.asarray(got_dist, dtype=np.float32), np.asarray(exp_dist, dtype=np.float32)
    assert got_ids.shape == exp_ids.shape, f"{ctx}: count {got_ids.shape} != {exp_ids.shape}"
    scale = np.maximum(np.abs(exp_dist), 1e-30)
    assert np.all(np.abs(got_dist - exp_dist) <= rel_tol * scale + 1e-12), f"{ctx}: distances differ"
    same = got_ids == exp_ids
    if exact_dist:
        assert np.array_equal(got_dist[same], exp_dist[same]), f"{ctx}: distances not bit-identical"
    for i in np.nonzero(~same)[0]:
        # a swap is only legal between near-ties
        lo, hi = max(i - 1, 0), min(i + 1, len(exp_dist) - 1)
        near = min(abs(float(exp_dist[i]) - float(exp_dist[lo])) if lo != i else np.inf,
                   abs(float(exp_dist[hi]) - float(exp_dist[i])) if hi != i else np.inf)
        edge = i == len(exp_dist) - 1  # the k-th slot may trade with the unseen (k+1)-th
        assert near <= rel_tol * max(abs(float(exp_dist[i])), 1e-30) or edge, \
            f"{ctx}: id mismatch at rank {i} without a near-tie: got {got_ids[i]} exp {exp_ids[i]}"


def numpy_merge(ids, dist, counts, ks):
    """Reference merge of G sorted per-shard lists: ids/dist [G,q,ks], counts [G,q] -> top-k by
    (distance, id).  Checker for the CUDA merge kernel and the gloo plumbing test."""
    G, q, _ = ids.shape
    out = []
    for i in range(q):
        pairs = []
        for g in range(G):
            c = int(counts[g, i])
            pairs += [(np.float32(dist[g, i, j]) + np.float32(0.0), int(ids[g, i, j])) for j in range(c)]
        pairs.sort(key=lambda t: (t[0], t[1]))
        k = int(ks[i])
        out.append(pairs[:k])
    return out
