"""Error types mirroring the reference's `VectorDbError` (src/error.rs:10-31)."""


class VectorDbError(Exception):
    pass


class DimensionMismatch(VectorDbError):
    """VectorDbError::DimensionMismatch { expected, actual } (src/error.rs:11-12)."""

    def __init__(self, expected, actual):
        super().__init__(f"Dimension mismatch: expected {expected}, got {actual}")
        self.expected = expected
        self.actual = actual


class VectorNotFound(VectorDbError):
    """VectorDbError::VectorNotFound { id } (src/error.rs:14-15)."""

    def __init__(self, id_):
        super().__init__(f"Vector not found: {id_}")
        self.id = id_


class InvalidVector(VectorDbError):
    """VectorDbError::InvalidVector { reason } (src/error.rs:17-18)."""


class IndexError_(VectorDbError):
    """VectorDbError::IndexError(String) (src/error.rs:29-30): CUDA failures, bad arguments."""


class NaNDistance(IndexError_):
    """A distance is NaN.  The reference panics (`partial_cmp().unwrap()`, src/flat_index.rs:62);
    libgfi reports it as an error instead of aborting (documented deviation)."""


class Unproven(IndexError_):
    """gfi_search_status only (GFI_ERR_UNPROVEN): a device-resident search could not prove its answer exact
    (more near-ties of the k-th distance than the candidate lists hold); re-run through `search`."""
