"""vectordb-from-scratch_b200 -- B200-native GpuFlatIndex behind the reference's `Index` trait.

The product is csrc/ (hand-written sm_100a CUDA + the C++ host runtime) built into
libgfi.so, whose C ABI is include/gfi.h.  This Python package is the thin ctypes binding
that stands in for the Rust `gpu-flat-index` crate in this Rust-less environment (see
INTEGRATION.md) plus a mirror of the reference's caller (`VectorStore`) so that the parity
tests read like the reference's own tests.  There is no CPU fallback anywhere: importing
works without a GPU (so symbols can be checked), every compute call needs a B200.
"""
from .errors import (  # noqa: F401
    VectorDbError, DimensionMismatch, InvalidVector, IndexError_, VectorNotFound, NaNDistance, Unproven,
)
from .native import lib, lib_path, build_native, DECLARED_SYMBOLS  # noqa: F401
from .index import GpuFlatIndex, DistanceMetric  # noqa: F401
from .store import VectorStore, Metadata, MetadataFilter, SearchResult, BatchInsertItem  # noqa: F401
