import sys, json
sys.path.insert(0, '.')
import numpy as np
import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200 import synth
idx = gfi.GpuFlatIndex(0, dim=128); idx.reserve(2000000); idx.add_generated(7, 0, 2000000, 0, 0); idx.set_option("profile", 1)
q = synth.gen_rows(8, 0, 4096, 128, 0); ks = np.full(4096, 10, np.uint32)
idx.search_arrays(q, ks)
for dbg in (0, 7, 23):
    idx.set_option("gemm_debug", dbg)
    s0 = idx.stats()
    try:
        for _ in range(3): idx.search_arrays(q, ks)
    except Exception as e: pass
    s1 = idx.stats()
    print(dbg, (s1["tensor_kernel_ns"]-s0["tensor_kernel_ns"])/max(1,(s1["tensor_kernel_count"]-s0["tensor_kernel_count"]))/1e6, "ms")
