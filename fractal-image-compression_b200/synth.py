"""Integer-only synthetic greyscale images.

No floating point, so numpy, C++ (tools/umma_probe.cu) and a Java port produce the
same bytes.  `noise` is uniform u8; `structured` is three integer triangle waves plus a
little noise (natural-image-like correlation, compressible by fractal coding).
"""
from __future__ import annotations

import numpy as np


def lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _hash(W: int, H: int, seed: int) -> np.ndarray:
    idx = np.arange(W * H, dtype=np.uint64).reshape(H, W)
    with np.errstate(over="ignore"):
        return lowbias32(((idx + np.uint64(seed) * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)).astype(np.uint32))


def noise(W: int, H: int, seed: int = 1) -> np.ndarray:
    return (_hash(W, H, seed) >> np.uint32(24)).astype(np.uint8)


def _tri(v: np.ndarray, period: int) -> np.ndarray:
    p = np.mod(v, period)
    half = period // 2
    t = np.where(p < half, p, period - p)
    return t * 255 // half


def structured(W: int, H: int, seed: int = 1) -> np.ndarray:
    y, x = np.mgrid[0:H, 0:W].astype(np.int64)
    h = _hash(W, H, seed)
    v = (_tri(x + seed, 97) + _tri(3 * y + x, 211) + _tri((x * y) // 64, 151)) // 3
    v = v + (h >> np.uint32(28)).astype(np.int64) - 8
    return np.clip(v, 0, 255).astype(np.uint8)


def grey_to_argb(plane: np.ndarray) -> np.ndarray:
    v = plane.astype(np.uint32)
    return (np.uint32(0xFF000000) | (v << np.uint32(16)) | (v << np.uint32(8)) | v).view(np.int32)
