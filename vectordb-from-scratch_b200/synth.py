"""Counter-based synthetic data generator, vectorised numpy.

element(seed, row, col) -> u32 -> f32; the same function is implemented in CUDA
(csrc/common.cuh gen_elem, used by gfi_add_generated) and in the test oracle, so any row
can be regenerated anywhere without moving it.  kind 0 = U[0,1); kind 1 = zero-mean,
unit-variance normal-like (Irwin-Hall-4 of 16-bit uniforms).
"""
import numpy as np

U32 = np.uint32


def _mix32(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7feb352d)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846ca68b)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def _gen_u32(seed, rows, cols, lane):
    m = np.uint64(0xFFFFFFFF)
    h = _mix32((np.uint64(seed) * np.uint64(0x9E3779B1) + (rows >> np.uint64(32)) +
                np.uint64(0x7F4A7C15) * np.uint64(lane)) & m)
    h = _mix32(h ^ (rows & m))
    return _mix32((h[:, None] + (cols[None, :] * np.uint64(0x85EBCA77))) & m)


def gen_rows(seed, first_row, n, d, kind):
    rows = np.arange(first_row, first_row + n, dtype=np.uint64)
    cols = np.arange(d, dtype=np.uint64)
    h = _gen_u32(seed, rows, cols, 0)
    if kind == 0:
        return ((h >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
    g = _gen_u32(seed, rows, cols, 1)
    s = ((h & np.uint64(0xFFFF)) + (h >> np.uint64(16)) + (g & np.uint64(0xFFFF)) + (g >> np.uint64(16))).astype(
        np.int64) - 2 * 65535
    scale = np.float32(np.float32(1.7320508) / np.float32(65536.0))
    return (s.astype(np.float32) * scale).astype(np.float32)
