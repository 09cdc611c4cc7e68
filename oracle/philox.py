"""Oracle restatement of the counter-based RNG of the CUDA kernels (TEST INFRASTRUCTURE).

Philox4x32-10 with the counter convention of deepbelief_b200/csrc/common.cuh:
  c0 = column (unit) index, c1 = global_row >> 2, c2/c3 = draw id low/high, key = seed low/high,
  output word j belongs to global_row & 3 == j.
Uniforms restate TensorFlow's documented construction [TF, not vendored] (23 mantissa bits,
bits(1.0f) | (x >> 9), minus 1); normals are Box-Muller on word pairs (0,1) and (2,3).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _to_uniform(x):
    bits = (np.uint32(0x3F800000) | (x >> np.uint32(9))).astype(np.uint32)
    return bits.view(np.float32) - np.float32(1.0)


def _words(rows, cols, seed, draw, row_offset):
    r = np.arange(rows, dtype=np.uint64)[:, None] + np.uint64(row_offset)
    c = np.arange(cols, dtype=np.uint64)[None, :]
    words = philox4x32_10(c, r >> np.uint64(2), np.uint64(draw & 0xFFFFFFFF), np.uint64((draw >> 32) & 0xFFFFFFFF),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return r, words


def uniform(rows, cols, seed, draw, row_offset=0):
    """float32 [rows, cols] uniforms in [0,1) exactly as the kernels draw them."""
    r, (x, y, z, w) = _words(rows, cols, seed, draw, row_offset)
    sel = (r & np.uint64(3)) + np.zeros_like(x, dtype=np.uint64)
    word = np.where(sel == 0, x, np.where(sel == 1, y, np.where(sel == 2, z, w)))
    return _to_uniform(word.astype(np.uint32))


def normal(rows, cols, seed, draw, row_offset=0):
    """float32 [rows, cols] standard normals (Box-Muller in fp32; device sincos/log differ by ~1 ulp)."""
    r, (x, y, z, w) = _words(rows, cols, seed, draw, row_offset)
    upper = ((r & np.uint64(2)) != 0) + np.zeros(x.shape, dtype=bool)
    a = np.where(upper, z, x).astype(np.uint32)
    b = np.where(upper, w, y).astype(np.uint32)
    u1 = np.maximum(_to_uniform(a), np.float32(1.0e-7))
    v1 = np.float32(6.28318530717958647692) * _to_uniform(b)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1))
    odd = ((r & np.uint64(1)) != 0) + np.zeros(x.shape, dtype=bool)
    return np.where(odd, rad * np.cos(v1), rad * np.sin(v1)).astype(np.float32)
