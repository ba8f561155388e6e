"""Independent scalar restatement in numpy.float32 (pure-Python loops; small cases only).

TEST INFRASTRUCTURE ONLY.  Used to cross-check flat_oracle.c: every operation is a
separately rounded binary32 op in the reference's order (src/distance.rs:37-73,
src/vector.rs:35-37), so the two must agree bit for bit.
"""
import numpy as np

f32 = np.float32


def norm(v):
    acc = f32(-0.0)
    for x in np.asarray(v, dtype=f32):
        acc = f32(acc + f32(x * x))
    return f32(np.sqrt(acc))


def dot(a, b):
    acc = f32(-0.0)
    for x, y in zip(np.asarray(a, dtype=f32), np.asarray(b, dtype=f32)):
        acc = f32(acc + f32(x * y))
    return acc


def euclidean(a, b):
    acc = f32(-0.0)
    for x, y in zip(np.asarray(a, dtype=f32), np.asarray(b, dtype=f32)):
        t = f32(x - y)
        acc = f32(acc + f32(t * t))
    return f32(np.sqrt(acc))


def cosine(a, b):
    n1, n2 = norm(a), norm(b)
    if n1 == 0 or n2 == 0:
        raise ValueError("InvalidVector")
    sim = f32(dot(a, b) / f32(n1 * n2))
    sim = f32(-1.0) if sim < -1 else (f32(1.0) if sim > 1 else sim)
    return f32(f32(1.0) - sim)


def distance(metric, q, x):
    if len(q) != len(x):
        raise ValueError("DimensionMismatch")
    if metric == "euclidean":
        return euclidean(q, x)
    if metric == "cosine":
        return cosine(q, x)
    return f32(-dot(q, x))


def flat_search(metric, rows, query, k, ids=None):
    res = []
    for r, row in enumerate(rows):
        res.append((distance(metric, query, row), int(r if ids is None else ids[r])))
    res.sort(key=lambda t: (t[0], t[1]))
    return res[:k]


def mix32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16; x = (x * 0x7feb352d) & 0xFFFFFFFF
    x ^= x >> 15; x = (x * 0x846ca68b) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def gen_u32(seed, row, col, lane):
    h = mix32(seed * 0x9E3779B1 + (row >> 32) + 0x7F4A7C15 * lane)
    h = mix32(h ^ (row & 0xFFFFFFFF))
    return mix32(h + col * 0x85EBCA77)


def gen_elem(seed, row, col, kind):
    h = gen_u32(seed, row, col, 0)
    if kind == 0:
        return f32(f32(h >> 8) * f32(1.0 / 16777216.0))
    g = gen_u32(seed, row, col, 1)
    s = (h & 0xFFFF) + (h >> 16) + (g & 0xFFFF) + (g >> 16) - 2 * 65535
    return f32(f32(s) * f32(f32(1.7320508) / f32(65536.0)))
