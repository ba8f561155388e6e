"""Regenerates tests/golden/oracle_cases.npz from the CPU oracle (run from the repo root).

The reference is Rust and cannot run here (no cargo/rustc), and it ships no golden files
(its randomised tests are unseeded), so the pinned vectors are (a) the reference's own
known-answer tests, transcribed in reference_kats.json, and (b) these seeded cases whose
expected output is frozen from the oracle after it passed (a) and the independent
numpy-float32 restatement (oracle/pyref.py).  They guard the oracle, the generator and the
CUDA path against regressions.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402

CASES = [  # name, metric, n, d, kind, seed, q, k
    ("l2_u_300x17", "euclidean", 300, 17, 0, 11, 5, 10),
    ("cos_n_257x64", "cosine", 257, 64, 1, 12, 4, 7),
    ("dot_n_500x33", "dot", 500, 33, 1, 13, 3, 100),
    ("l2_u_1000x128", "euclidean", 1000, 128, 0, 1, 4, 10),
]


def main():
    out = {}
    for name, metric, n, d, kind, seed, q, k in CASES:
        rows = oracle.gen_rows(seed, 0, n, d, kind)
        queries = oracle.gen_rows(seed + 1000, 0, q, d, kind)
        res = oracle.search_batch(metric, rows, queries, k)
        out[name + "/ids"] = np.stack([r[0] for r in res])
        out[name + "/dist"] = np.stack([r[1] for r in res])
        out[name + "/rows_head"] = rows[:2, :4]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
