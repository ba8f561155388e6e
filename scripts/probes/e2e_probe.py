import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200 import synth
from bench import WORKLOADS, METRIC_ID
wl = sys.argv[1]
metric, n, d, kind, seed, q, k = WORKLOADS[wl]
idx = gfi.GpuFlatIndex(METRIC_ID[metric], dim=d); idx.reserve(n); idx.add_generated(seed, 0, n, kind, 0); idx.set_option("profile", 1)
qh = torch.from_numpy(synth.gen_rows(seed + 1, 0, q, d, kind)).pin_memory().numpy()
ks = np.full(q, k, np.uint32)
for _ in range(3): idx.search_arrays(qh, ks)
torch.cuda.synchronize()
for rep in range(2):
    s0 = idx.stats(); t0 = time.perf_counter()
    N = 10
    for _ in range(N): idx.search_arrays(qh, ks)
    t = (time.perf_counter() - t0) / N; s1 = idx.stats()
    kn = "tensor" if s1["tensor_kernel_count"] > s0["tensor_kernel_count"] else "scan"
    km = (s1[kn + "_kernel_ns"] - s0[kn + "_kernel_ns"]) / N / 1e6
    print(wl, "e2e ms", round(t * 1e3, 4), "dominant kernel ms", round(km, 4))
# raw ctypes call timing without numpy allocation
import ctypes
L = gfi.lib(); h = idx._h
out_ids = np.zeros((q, k), np.uint64); out_d = np.zeros((q, k), np.float32); cnt = np.zeros(q, np.uint32)
t0 = time.perf_counter()
for _ in range(10):
    L.gfi_search(h, qh.ctypes.data, q, d, ks.ctypes.data, None, 0, out_ids.ctypes.data, out_d.ctypes.data, cnt.ctypes.data, k)
print(wl, "raw C call ms", round((time.perf_counter() - t0) / 10 * 1e3, 4))
