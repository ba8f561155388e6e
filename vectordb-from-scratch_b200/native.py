"""Loader of libgfi.so (the C ABI declared in include/gfi.h).  Fails loudly: no fallback."""
import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "libgfi.so")
_HEADER = os.path.join(_ROOT, "include", "gfi.h")
_lib = None


def lib_path():
    return _SO


def build_native(force=False, jobs=8):
    """Compile every CUDA source for sm_100a into libgfi.so (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-s", "-C", csrc, "clean"])
    subprocess.check_call(["make", "-s", "-C", csrc, f"-j{jobs}"])
    if not os.path.exists(_SO):
        raise RuntimeError("libgfi.so was not produced")
    return _SO


def _declared_symbols():
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gfi_[a-z0-9_]+)\s*\(", text)))


DECLARED_SYMBOLS = _declared_symbols()


class GfiStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in (
        "n_slots", "n_live", "searches", "queries", "scan_queries", "tensor_queries", "fallback_queries",
        "kernel_launches", "bytes_fp32", "bytes_fp16", "scan_kernel_ns", "scan_kernel_count",
        "tensor_kernel_ns", "tensor_kernel_count", "coalesced_batches", "coalesced_requests", "shards", "merge_ns",
        "merge_count", "paged_queries")]


def lib():
    """The loaded library.  Raises if libgfi.so is missing -- there is no CPU path to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError(
            f"{_SO} is missing: build it with __graft_entry__.build() / make -C vectordb-from-scratch_b200/csrc "
            "(libgfi has no CPU fallback)")
    L = ctypes.CDLL(_SO)
    c = ctypes
    vp, i32, i64, u32, u64 = c.c_void_p, c.c_int32, c.c_int64, c.c_uint32, c.c_uint64
    sig = {
        "gfi_version": (i32, []),
        "gfi_last_error": (c.c_char_p, []),
        "gfi_last_mismatch": (None, [c.POINTER(i64), c.POINTER(i64)]),
        "gfi_create": (i32, [c.POINTER(vp), i32, i64, i32, u32]),
        "gfi_create_sharded": (i32, [c.POINTER(vp), i32, i64, c.POINTER(i32), i32, u32]),
        "gfi_destroy": (i32, [vp]),
        "gfi_add": (i32, [vp, vp, vp, i64, i64]),
        "gfi_add_generated": (i32, [vp, u32, u64, i64, i32, u64]),
        "gfi_add_from_file": (i32, [vp, c.c_char_p, u64, c.POINTER(i64)]),
        "gfi_remove": (i32, [vp, u64]),
        "gfi_len": (i64, [vp]),
        "gfi_metric": (i32, [vp]),
        "gfi_dim": (i64, [vp]),
        "gfi_get_vector": (i32, [vp, u64, vp, i64, c.POINTER(i64)]),
        "gfi_flush": (i32, [vp]),
        "gfi_reserve": (i32, [vp, i64]),
        "gfi_compact": (i32, [vp]),
        "gfi_search": (i32, [vp, vp, i64, i64, vp, vp, i64, vp, vp, vp, i64]),
        "gfi_search_filtered": (i32, [vp, vp, i64, i64, vp, c.c_char_p, vp, vp, vp, i64]),
        "gfi_set_metadata": (i32, [vp, u64, i32, c.POINTER(c.c_char_p), c.POINTER(c.c_char_p)]),
        "gfi_set_metadata_column": (i32, [vp, c.c_char_p, vp, i64, c.POINTER(c.c_char_p), i32, vp]),
        "gfi_search_device": (i32, [vp, vp, i64, vp, u32, vp, i64, vp, vp, vp, i64, vp]),
        "gfi_search_status": (i32, [vp]),
        "gfi_merge_topk_device": (i32, [vp, vp, vp, i32, i64, i64, vp, vp, vp, vp, i64, vp]),
        "gfi_distances": (i32, [vp, vp, i64, i64, vp, i64, vp, vp]),
        "gfi_merge_topk_device_strided": (i32, [vp, vp, vp, i32, i64, i64, i64, vp, vp, vp, vp, i64, vp]),
        "gfi_get_stats": (i32, [vp, c.POINTER(GfiStats)]),
        "gfi_set_option": (i32, [vp, c.c_char_p, i64]),
        "gfi_debug_tensor_scores": (i32, [vp, vp, i64, vp, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return _lib
