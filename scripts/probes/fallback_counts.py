"""Prints the certification fallbacks of the tensor-path test cases (what the asserts in tests/test_gpu_parity.py bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import oracle
import vectordb_from_scratch_b200 as gfi
from test_gpu_parity import TENSOR_CASES, build
for metric, n, d, kind, q, k in TENSOR_CASES:
    rows = oracle.gen_rows(300 + d, 0, n, d, kind)
    queries = oracle.gen_rows(400 + d, 0, q, d, kind)
    idx = build(metric, rows)
    idx.search_arrays(queries, k)
    st = idx.stats()
    print("tensor case", metric, n, d, q, k, "fallbacks", st["fallback_queries"], "of", q)
for metric in ("cosine", "euclidean"):
    n, d, q, k = 30000, 192, 300, 10
    rows = oracle.gen_rows(71, 0, n, d, 1)
    queries = oracle.gen_rows(72, 0, q, d, 1)
    idx = build(metric, rows)
    idx.set_option("pair", 1)
    idx.search_arrays(queries, k)
    print("pair case", metric, "fallbacks", idx.stats()["fallback_queries"], "of", q)
