"""CPU oracle for the exact-search hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See flat_oracle.c for what it restates and why.
"""
from .oracle import (  # noqa: F401
    METRICS, OracleError, build, distance, norm, flat_search, search_post_filter,
    search_batch, search_generated, gen_rows, max_threads,
)
