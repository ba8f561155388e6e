"""CPU: the C-ABI library loads and exports every symbol include/gfi.h declares; compute
entry points fail loudly without a GPU; host-side mirror logic (filters, masks)."""
import ctypes
import os

import numpy as np
import pytest

import vectordb_from_scratch_b200 as gfi
from vectordb_from_scratch_b200.index import pack_mask
from vectordb_from_scratch_b200.store import Metadata, MetadataFilter


def test_library_exports_every_declared_symbol():
    assert os.path.exists(gfi.lib_path()), "build libgfi.so first (__graft_entry__.build())"
    L = ctypes.CDLL(gfi.lib_path())
    assert len(gfi.DECLARED_SYMBOLS) >= 20
    for name in gfi.DECLARED_SYMBOLS:
        assert hasattr(L, name), f"{name} declared in include/gfi.h but not exported"
    assert gfi.lib().gfi_version() >= 100


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_c_client():
    """gcc -Wall -Werror against include/gfi.h, linked against libgfi.so (no compute call happens at build time)."""
    import subprocess
    src = os.path.join(ROOT, "tests", "c_client", "abi_client.c")
    exe = os.path.join(ROOT, "tests", "c_client", "abi_client")
    lib_dir = os.path.dirname(gfi.lib_path())
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", lib_dir,
                           "-lgfi", f"-Wl,-rpath,{lib_dir}", "-lm", "-o", exe])
    return exe


def test_c_client_compiles_and_links_against_the_header():
    assert os.path.exists(build_c_client())


def _c_arg_counts():
    """function name -> number of parameters, from include/gfi.h"""
    import re
    text = open(os.path.join(ROOT, "include", "gfi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(gfi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_rust_sys_crate_declares_every_header_symbol():
    """rust/gpu-flat-index-sys/src/lib.rs is not compiled here (no cargo): keep it in step with the header --
    same function set, same number of arguments, gfi_stats with the header's fields in the header's order."""
    import re
    rs = open(os.path.join(ROOT, "rust", "gpu-flat-index-sys", "src", "lib.rs")).read()
    decl = {}
    for m in re.finditer(r"pub fn (gfi_[a-z0-9_]+)\s*\(([^)]*)\)", rs, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    c = _c_arg_counts()
    assert set(c) == set(gfi.DECLARED_SYMBOLS)
    assert decl == c, {k: (decl.get(k), c.get(k)) for k in set(decl) | set(c) if decl.get(k) != c.get(k)}
    from vectordb_from_scratch_b200 import native
    fields = re.findall(r"pub (\w+): i64", rs)
    assert fields == [n for n, _ in native.GfiStats._fields_]
    hdr = open(os.path.join(ROOT, "include", "gfi.h")).read()
    hdr_stats = hdr[hdr.index("typedef struct gfi_stats"):hdr.index("} gfi_stats;")]
    hdr_stats = re.sub(r"/\*.*?\*/", "", hdr_stats, flags=re.S)
    hdr_fields = [f.strip() for line in re.findall(r"int64_t ([^;]+);", hdr_stats) for f in line.split(",")]
    assert hdr_fields == fields
    for code in re.findall(r"#define (GFI_[A-Z_]+) (\d+)", hdr):
        assert re.search(rf"pub const {code[0]}: [iu]32 = {code[1]};", rs), code


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gfi.IndexError_) as e:
        gfi.GpuFlatIndex(gfi.DistanceMetric.Euclidean)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(os.path.dirname(gfi.lib_path()))
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_pack_mask_layout():
    bits = np.zeros(130, dtype=bool)
    bits[[0, 63, 64, 129]] = True
    words, n = pack_mask(bits)
    assert n == 130 and words.dtype == np.uint64 and len(words) == 3
    assert words[0] == (1 | (1 << 63)) and words[1] == 1 and words[2] == 2


# MetadataFilter truth table, reference src/storage.rs:457-575
def _md(**kw):
    m = Metadata()
    for k, v in kw.items():
        m.insert(k, v)
    return m


def test_filter_eq_ne_exists():
    red = _md(color="red")
    assert MetadataFilter.eq("color", "red").matches(red)
    assert not MetadataFilter.eq("color", "blue").matches(red)
    assert not MetadataFilter.eq("size", "red").matches(red)
    assert MetadataFilter.ne("color", "blue").matches(red)
    assert not MetadataFilter.ne("color", "red").matches(red)
    assert MetadataFilter.ne("size", "x").matches(red)  # absent field: Ne is true (storage.rs:65)
    assert MetadataFilter.exists("color").matches(red)
    assert not MetadataFilter.exists("size").matches(red)


def test_filter_and_or():
    m = _md(color="red", size="large")
    f_and = MetadataFilter.and_([MetadataFilter.eq("color", "red"), MetadataFilter.eq("size", "large")])
    assert f_and.matches(m)
    assert not MetadataFilter.and_([MetadataFilter.eq("color", "red"), MetadataFilter.eq("size", "s")]).matches(m)
    assert MetadataFilter.or_([MetadataFilter.eq("color", "blue"), MetadataFilter.eq("size", "large")]).matches(m)
    assert not MetadataFilter.or_([MetadataFilter.eq("color", "blue")]).matches(m)
    assert MetadataFilter.and_([]).matches(m)      # all() of nothing
    assert not MetadataFilter.or_([]).matches(m)   # any() of nothing
    f = MetadataFilter.from_json({"op": "and", "filters": [{"op": "eq", "field": "color", "value": "red"},
                                                          {"op": "exists", "field": "size"}]})
    assert f.matches(m)
