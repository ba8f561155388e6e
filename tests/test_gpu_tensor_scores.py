"""GPU: the raw output of the tcgen05 candidate pass (TMA SWIZZLE_128B loads, UMMA descriptors,
TMEM epilogue mapping) against a float64 numpy evaluation of the same fp16-rounded operands, and
the rigorous error bound the certification relies on against the exact scores."""
import numpy as np
import pytest

import oracle
import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200 import DistanceMetric as DM

pytestmark = pytest.mark.gpu


def fp16_scaled(a, per_row):
    """The conversion of ingest.cu / convert_queries16_kernel: power-of-two scale to [2^14, 2^15)."""
    a = np.asarray(a, dtype=np.float32)
    m = np.abs(a).max(axis=1, keepdims=True) if per_row else np.full((a.shape[0], 1), np.abs(a).max())
    e = np.floor(np.log2(np.where(m > 0, m, 1.0)))
    s = np.exp2(14 - e).astype(np.float32)
    v = a * s
    v = np.where(np.abs(v) < 2.0 ** -14, 0.0, v)
    return v.astype(np.float16).astype(np.float64), s.astype(np.float64)


def fp16_normalised(a):
    """ingest.cu for cosine: rows times fl(2^14 / norm), norm = the reference's sequential f32 sum of squares."""
    a = np.asarray(a, dtype=np.float32)
    sumsq = np.cumsum(a * a, axis=1, dtype=np.float32)[:, -1]
    nrm = np.sqrt(sumsq, dtype=np.float32)
    t = (np.float32(16384.0) / nrm).astype(np.float32)
    v = (a * t[:, None]).astype(np.float32)
    v = np.where(np.abs(v) < 2.0 ** -14, np.float32(0.0), v)
    return v.astype(np.float16).astype(np.float64), np.full((a.shape[0], 1), 16384.0)


@pytest.mark.parametrize("metric,n,d,q,kind", [
    ("dot", 512, 64, 128, 1),        # one k-block, exact tiles
    ("dot", 700, 128, 5, 1),         # partial row tile, partial query tile
    ("euclidean", 1000, 200, 130, 0),  # K tail (200 = 3*64 + 8), two query tiles
    ("cosine", 2048, 768, 64, 1),    # 12 k-blocks: ring wraps 3 times
])
def test_tensor_scores_match_fp16_model(metric, n, d, q, kind):
    rows = oracle.gen_rows(61, 0, n, d, kind)
    queries = oracle.gen_rows(62, 0, q, d, kind)
    idx = gfi.GpuFlatIndex({"dot": DM.DotProduct, "euclidean": DM.Euclidean, "cosine": DM.Cosine}[metric])
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    got = idx.debug_tensor_scores(queries).astype(np.float64)
    q16, sq = fp16_scaled(queries, per_row=False)
    r64 = rows.astype(np.float64)
    if metric == "cosine":
        # cosine rows are stored normalised: fp16(x * (2^14 / ||x||)), norm in the reference's f32 arithmetic
        x16, sx = fp16_normalised(rows)
    else:
        x16, sx = fp16_scaled(rows, per_row=True)
    dot = (q16 @ x16.T) / (sq * sx.T)  # [q, n]; for cosine this is already q.x / ||x||
    if metric == "dot":
        model = -dot
    elif metric == "cosine":
        model = -dot
    else:
        model = (r64 ** 2).sum(axis=1)[None, :] - 2.0 * dot
    scale = np.abs(model).max()
    err = np.abs(got - model).max()
    assert err <= 2e-5 * scale, f"tcgen05 scores differ from the fp16 model: max err {err} (scale {scale})"
    # the certification bound: |approx dot - exact dot| <= eps_rel * ||q|| * ||x||  (cosine: both sides / ||x||)
    exact = queries.astype(np.float64) @ r64.T
    qn = np.linalg.norm(queries.astype(np.float64), axis=1)[:, None]
    xn = np.linalg.norm(r64, axis=1)[None, :]
    eps_rel = 2.0 ** -10 + 2.0 ** -20 + 2 * np.sqrt(d) * 2.0 ** -26 + (d + 8) * 2.0 ** -23
    if metric == "cosine":
        assert np.all(np.abs(dot - exact / xn) <= eps_rel * qn)
    else:
        assert np.all(np.abs(dot - exact) <= eps_rel * qn * xn)
