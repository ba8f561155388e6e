#!/usr/bin/env python
"""K2 probe: times the tcgen05 main pass of one search per configuration (CUDA events inside libgfi, option
`profile`) and prints the in-kernel clock diagnostics (gemm_debug bit 5) -- k-ring kernel vs the short-K
row-tile-stationary kernel, with loads / epilogue / MMAs switched off one at a time (results invalid then).
Also checks that both kernels return identical results.  No torch: ctypes + numpy only.

usage: k2_probe.py [rows] [dim] [batch] [metric]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vectordb_from_scratch_b200 as gfi  # noqa: E402
from vectordb_from_scratch_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    q = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    metric = sys.argv[4] if len(sys.argv) > 4 else "euclidean"
    variants = sys.argv[5].split(",") if len(sys.argv) > 5 else None
    mid = {"euclidean": 0, "cosine": 1, "dot": 2}[metric]
    kind = 0 if metric == "euclidean" else 1
    k = 10
    idx = gfi.GpuFlatIndex(mid, dim=d)
    idx.reserve(n)
    idx.add_generated(7, 0, n, kind, 0)
    idx.set_option("profile", 1)
    queries = synth.gen_rows(8, 0, q, d, kind)
    ks = np.full(q, k, dtype=np.uint32)

    def run(name, opts, reps=5, check=None):
        for o, v in opts.items():
            idx.set_option(o, v)
        for _ in range(2):
            res = idx.search_arrays(queries, ks)
        s0 = idx.stats()
        t0 = time.perf_counter()
        for _ in range(reps):
            res = idx.search_arrays(queries, ks)
        wall = (time.perf_counter() - t0) / reps
        s1 = idx.stats()
        cnt = s1["tensor_kernel_count"] - s0["tensor_kernel_count"]
        kms = (s1["tensor_kernel_ns"] - s0["tensor_kernel_ns"]) / max(cnt, 1) / 1e6
        tf = 2.0 * n * d * q / (kms * 1e-3) / 1e12 if kms > 0 else 0.0
        fb = s1["fallback_queries"] - s0["fallback_queries"]
        same = ""
        if check is not None:
            same = " identical=%s" % (np.array_equal(res[0], check[0]) and np.array_equal(res[1], check[1])
                                      and np.array_equal(res[2], check[2]))
        print("%-34s main pass %8.3f ms  %7.1f TFLOP/s  call %8.3f ms  fallbacks/call %.1f%s" %
              (name, kms, tf, wall * 1e3, fb / reps, same), flush=True)
        return res

    print("# %d x %d %s, batch %d, k=%d" % (n, d, metric, q, k), flush=True)
    base = {"gemm_debug": 0, "short_k": 0}
    ref = run("k-ring kernel", base)
    new = run("short-K row-stationary", {"gemm_debug": 0, "short_k": 1}, check=ref)
    if variants is None:
        variants = ["32", "35", "36", "39", "48", "55"]
    for sk in (0, 1):
        for v in variants:
            bits = int(v)
            what = []
            if bits & 3:
                what.append("loads off")
            if bits & 4:
                what.append("epilogue off")
            if bits & 16:
                what.append("mma off")
            run("short_k=%d debug=%d (%s)" % (sk, bits, ", ".join(what) or "all on"), {"gemm_debug": bits, "short_k": sk},
                reps=2)
    idx.set_option("gemm_debug", 0)
    idx.close()


if __name__ == "__main__":
    main()
