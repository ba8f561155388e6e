"""Small driver for ncu / timing experiments: builds one workload and runs a few searches.
usage: python scripts/prof_one.py --workload c2 [--rows N] [--q Q] [--steps S] [--debug-sweep]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import vectordb_from_scratch_b200 as gfi  # noqa: E402
from vectordb_from_scratch_b200 import synth  # noqa: E402
from bench import WORKLOADS, METRIC_ID  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--q", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--debug-sweep", action="store_true")
ap.add_argument("--metric", default="")
ap.add_argument("--opt", action="append", default=[], help="name=value passed to gfi_set_option")
a = ap.parse_args()
metric, n, d, kind, seed, q, k = WORKLOADS[a.workload]
n = a.rows or n
metric = a.metric or metric
q = a.q or q
idx = gfi.GpuFlatIndex(METRIC_ID[metric], dim=d)
idx.reserve(n)
idx.add_generated(seed, 0, n, kind, 0)
idx.set_option("profile", 1)
for o in a.opt:
    name, val = o.split("=")
    idx.set_option(name, int(val))
queries = synth.gen_rows(seed + 1, 0, q, d, kind)
try:  # pinned host queries, as the bench's e2e leg uses
    import torch
    queries = torch.from_numpy(queries).pin_memory().numpy()
except Exception:
    pass
import time
ks = np.full(q, k, dtype=np.uint32)


def timed(steps):
    s0 = idx.stats()
    t0 = time.perf_counter()
    for _ in range(steps):
        idx.search_arrays(queries, ks)
    wall = (time.perf_counter() - t0) / steps
    s1 = idx.stats()
    out = {"e2e_ms": round(wall * 1e3, 4)}
    for kname in ("tensor", "scan"):
        c = s1[f"{kname}_kernel_count"] - s0[f"{kname}_kernel_count"]
        if c:
            out[kname + "_ms"] = (s1[f"{kname}_kernel_ns"] - s0[f"{kname}_kernel_ns"]) / c / 1e6
    out["fallback"] = s1["fallback_queries"] - s0["fallback_queries"]
    return out


if a.debug_sweep:
    idx.search_arrays(queries, ks)
    for dbg in (32, 32 + 4, 32 + 7, 32 + 23, 32 + 19):  # bit5: print cycles / clock of CTA 0; 51 = epilogue only
        idx.set_option("gemm_debug", dbg)
        try:
            r = timed(a.steps)
        except Exception as e:  # garbage results may trip NaN checks: timing is still recorded
            r = {"error": str(e)[:80]}
        print(json.dumps({"workload": a.workload, "n": n, "q": q, "gemm_debug": dbg, **r}))
else:
    idx.search_arrays(queries, ks)
    print(json.dumps({"workload": a.workload, "n": n, "q": q, **timed(a.steps)}))
