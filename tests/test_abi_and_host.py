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
