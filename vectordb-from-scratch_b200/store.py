"""Mirror of the reference's caller of the hot path: `VectorStore` (src/storage.rs:83-348).

Only what the parity tests need: String<->internal id mapping, dimension latch, metadata and
`MetadataFilter` truth table (src/storage.rs:47-71), `search`, `search_with_filter`
(reference post-filter with 3x over-fetch, src/storage.rs:249-290), the batch loops
(src/storage.rs:302-322) -- plus the two push-down modes that the additive trait methods
enable (one batched C-ABI call; filter evaluated once into an eligibility bitmask).
In the Rust integration this file does not exist: the reference's own VectorStore is used
unchanged.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .errors import DimensionMismatch, VectorNotFound
from .index import GpuFlatIndex, DistanceMetric


@dataclass
class SearchResult:
    id: str
    distance: float


@dataclass
class Metadata:
    fields: Dict[str, str] = field(default_factory=dict)

    def insert(self, key, value):
        self.fields[key] = value

    def get(self, key):
        return self.fields.get(key)


class MetadataFilter:
    """src/storage.rs:44-71.  JSON form {"op": "eq"|"ne"|"exists"|"and"|"or", ...}."""

    def __init__(self, op, field=None, value=None, filters=None):
        self.op, self.field, self.value, self.filters = op, field, value, filters or []

    @staticmethod
    def eq(field, value): return MetadataFilter("eq", field, value)
    @staticmethod
    def ne(field, value): return MetadataFilter("ne", field, value)
    @staticmethod
    def exists(field): return MetadataFilter("exists", field)
    @staticmethod
    def and_(filters): return MetadataFilter("and", filters=list(filters))
    @staticmethod
    def or_(filters): return MetadataFilter("or", filters=list(filters))

    @staticmethod
    def from_json(obj):
        op = obj["op"]
        if op in ("and", "or"):
            return MetadataFilter(op, filters=[MetadataFilter.from_json(f) for f in obj["filters"]])
        return MetadataFilter(op, obj.get("field"), obj.get("value"))

    def to_json(self):
        if self.op in ("and", "or"):
            return {"op": self.op, "filters": [f.to_json() for f in self.filters]}
        out = {"op": self.op, "field": self.field}
        if self.op != "exists":
            out["value"] = self.value
        return out

    def matches(self, md: Metadata) -> bool:
        if self.op == "eq":
            return md.get(self.field) == self.value and md.get(self.field) is not None
        if self.op == "ne":
            return md.get(self.field) != self.value  # true when the field is absent (storage.rs:65)
        if self.op == "exists":
            return md.get(self.field) is not None
        if self.op == "and":
            return all(f.matches(md) for f in self.filters)  # empty => true
        if self.op == "or":
            return any(f.matches(md) for f in self.filters)  # empty => false
        raise ValueError(self.op)


@dataclass
class BatchInsertItem:
    id: str
    vector: np.ndarray
    metadata: Metadata = field(default_factory=Metadata)


class VectorStore:
    """VectorStore<I: Index> with I = GpuFlatIndex (src/storage.rs:83-127)."""

    def __init__(self, metric=DistanceMetric.Euclidean, index: Optional[GpuFlatIndex] = None, device=0,
                 flags=0):
        self.index = index if index is not None else GpuFlatIndex(metric, device=device, flags=flags)
        self.id_to_internal: Dict[str, int] = {}
        self.internal_to_id: Dict[int, str] = {}
        self.metadata: Dict[int, Metadata] = {}
        self.next_id = 0
        self.dimension: Optional[int] = None
        self._host_rows: Dict[int, np.ndarray] = {}  # host mirror (what a Rust wrapper lends as &Vector)
        self.device_metadata = True  # also keep metadata as dictionary-encoded columns on the GPU

    with_index = classmethod(lambda cls, index: cls(index=index))

    def insert(self, id, vector):
        self.insert_with_metadata(id, vector, Metadata())

    def insert_with_metadata(self, id, vector, metadata):
        """src/storage.rs:135-172"""
        v = np.asarray(vector, dtype=np.float32).reshape(-1)
        if self.dimension is None:
            self.dimension = v.size
        elif v.size != self.dimension:
            raise DimensionMismatch(self.dimension, v.size)
        if id in self.id_to_internal:
            old = self.id_to_internal.pop(id)
            self.index.remove(old)
            self.internal_to_id.pop(old, None)
            self.metadata.pop(old, None)
            self._host_rows.pop(old, None)
        internal = self.next_id
        self.next_id += 1
        self.index.add(internal, v)
        if self.device_metadata:
            self.index.set_metadata(internal, metadata.fields)
        self.id_to_internal[id] = internal
        self.internal_to_id[internal] = id
        self.metadata[internal] = metadata
        self._host_rows[internal] = v.copy()

    def insert_batch(self, items: List[BatchInsertItem]):
        """src/storage.rs:293-298: stops at the first error, earlier items stay."""
        for it in items:
            self.insert_with_metadata(it.id, it.vector, it.metadata)

    def get(self, id):
        internal = self.id_to_internal.get(id)
        if internal is None:
            raise VectorNotFound(id)
        return self._host_rows[internal]

    def get_metadata(self, id):
        internal = self.id_to_internal.get(id)
        return None if internal is None else self.metadata.get(internal)

    def delete(self, id):
        internal = self.id_to_internal.pop(id, None)
        if internal is None:
            raise VectorNotFound(id)
        v = self._host_rows.pop(internal)
        self.index.remove(internal)
        self.internal_to_id.pop(internal, None)
        self.metadata.pop(internal, None)
        return v

    def len(self):
        return self.index.len()

    __len__ = len

    def is_empty(self):
        return self.index.is_empty()

    def list_ids(self):
        return list(self.id_to_internal.keys())

    # ---- search paths ----
    def _check_dim(self, query):
        q = np.asarray(query, dtype=np.float32).reshape(-1)
        if self.dimension is not None and q.size != self.dimension:
            raise DimensionMismatch(self.dimension, q.size)
        return q

    def _to_results(self, pairs):
        return [SearchResult(self.internal_to_id[i], d) for i, d in pairs if i in self.internal_to_id]

    def search(self, query, k):
        """src/storage.rs:217-245"""
        if self.is_empty():
            return []
        q = self._check_dim(query)
        return self._to_results(self.index.search(q, k))

    def search_with_filter(self, query, k, flt: MetadataFilter, pushdown=False):
        """src/storage.rs:249-290 (post-filter).  pushdown=True evaluates the filter once into an
        eligibility bitmask and searches only matching rows (exact pre-filter, SURVEY.md 8(b) B5 ii)."""
        if self.is_empty():
            return []
        q = self._check_dim(query)
        if pushdown == "device":  # filter evaluated on the GPU from the device-side metadata columns
            ids, dist, cnt = self.index.search_filtered(q[None, :], [k], flt.to_json())
            return self._to_results([(int(ids[0, j]), float(dist[0, j])) for j in range(cnt[0])])
        if pushdown:
            return self._to_results(self.index.search_masked(q, k, self.filter_mask(flt)))
        fetch_k = min(max(k * 3, k), self.len())
        out = []
        for internal, dist in self.index.search(q, fetch_k):
            sid = self.internal_to_id.get(internal)
            md = self.metadata.get(internal)
            if sid is None or md is None:
                continue
            if flt.matches(md):
                out.append(SearchResult(sid, dist))
                if len(out) == k:
                    break
        return out

    def filter_mask(self, flt):
        m = np.zeros(self.next_id, dtype=bool)
        for internal, md in self.metadata.items():
            if internal in self.internal_to_id and flt.matches(md):
                m[internal] = True
        return m

    def search_batch(self, queries, batched=True):
        """src/storage.rs:302-310.  batched=False is the reference's sequential loop of single-query
        searches; batched=True is the additive Index::search_batch delegation (one C-ABI call)."""
        if not batched:
            return [self.search(q, k) for q, k in queries]
        if self.is_empty():
            return [[] for _ in queries]
        qs = [(self._check_dim(q), k) for q, k in queries]
        return [self._to_results(r) for r in self.index.search_batch(qs)]

    def search_batch_with_filter(self, queries, flt, pushdown=False):
        """src/storage.rs:313-322"""
        if not pushdown:
            return [self.search_with_filter(q, k, flt) for q, k in queries]
        if self.is_empty():
            return [[] for _ in queries]
        qs = [(self._check_dim(q), k) for q, k in queries]
        if pushdown == "device":
            ids, dist, cnt = self.index.search_filtered(np.stack([q for q, _ in qs]), [k for _, k in qs], flt.to_json())
            return [self._to_results([(int(ids[i, j]), float(dist[i, j])) for j in range(cnt[i])])
                    for i in range(len(qs))]
        mask = self.filter_mask(flt)
        return [self._to_results(r) for r in self.index.search_batch(qs, mask=mask)]
