"""``jax.random.normal(jax.random.PRNGKey(0), (3, 3))`` without jax.

The reference mixes the joints with ``J = eye(3) + mean * normal(PRNGKey(0))``
(trajectory.py:42).  jax is not a dependency of this package, so the draw is
re-created here: Threefry-2x32 (20 rounds), 23 mantissa bits -> uniform on
(-1, 1) -> sqrt(2) * erfinv.  Two streams exist in the wild:

* ``legacy``        jax < 0.5 (``jax_threefry_partitionable=False``): the 9
                    counters are padded to 10 and hashed as two halves.  This is
                    the stream the reference's checked-in results were made with.
* ``partitionable`` jax >= 0.5 default: one hash per element, ``y0 ^ y1``.
"""
from __future__ import annotations

import numpy as np

_R = (13, 15, 26, 6, 17, 29, 16, 24)
_PARITY = 0x1BD11BDA
_M = 0xFFFFFFFF


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M


def threefry2x32_scalar(k0: int, k1: int, c0: int, c1: int):
    """One Threefry-2x32-20 block on Python ints."""
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    x0, x1 = (c0 + ks[0]) & _M, (c1 + ks[1]) & _M
    for block in range(5):
        rots = _R[:4] if block % 2 == 0 else _R[4:]
        for r in rots:
            x0 = (x0 + x1) & _M
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(block + 1) % 3]) & _M
        x1 = (x1 + ks[(block + 2) % 3] + block + 1) & _M
    return x0, x1


def _random_bits(n: int, stream: str):
    if stream == "legacy":
        half = (n + 1) // 2
        counters = list(range(n)) + [0] * (2 * half - n)
        lo, hi = [], []
        for i in range(half):
            y0, y1 = threefry2x32_scalar(0, 0, counters[i], counters[half + i])
            lo.append(y0)
            hi.append(y1)
        return (lo + hi)[:n]
    if stream == "partitionable":
        out = []
        for i in range(n):
            y0, y1 = threefry2x32_scalar(0, 0, 0, i)
            out.append(y0 ^ y1)
        return out
    raise ValueError(f"unknown jax PRNG stream {stream!r}")


def normal_key0(shape=(3, 3), stream: str = "legacy") -> np.ndarray:
    from scipy.special import erfinv
    n = int(np.prod(shape))
    bits = np.array(_random_bits(n, stream), dtype=np.uint32)
    mant = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    u = np.maximum(lo, mant * (np.float32(1.0) - lo) + lo).astype(np.float32)
    z = np.float32(np.sqrt(2.0)) * erfinv(u.astype(np.float64)).astype(np.float32)
    return z.astype(np.float32).reshape(shape)
