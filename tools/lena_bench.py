"""Encode / decode times of the bundled Lena images (decoded pixels in tests/golden) at the reference's default
settings and with the full pool: GPU (through the C ABI, host buffers in and out) against the CPU oracle.

    python tools/lena_bench.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fractal_image_compression_b200 as fic  # noqa: E402
from oracle import oracle as O  # noqa: E402  (measurement script: the CPU leg is the oracle)

GOLD = os.path.join(ROOT, "tests", "golden")


def grey(name, n):
    p = np.fromfile(os.path.join(GOLD, name), np.uint8).reshape(n, n).astype(np.uint32)
    return (np.uint32(0xFF000000) | (p << 16) | (p << 8) | p).view(np.int32)


def rgb(name, n):
    a = np.fromfile(os.path.join(GOLD, name), np.uint8).reshape(n, n, 3).astype(np.uint32)
    return (np.uint32(0xFF000000) | (a[..., 0] << 16) | (a[..., 1] << 8) | a[..., 2]).view(np.int32)


def main():
    h = fic.Handle(0)
    cases = [("LenaGrey 256^2 B=8 wk=2 (reference default)", grey("lena_grey_256.u8", 256), 8, 2, False),
             ("LenaGrey 256^2 B=8 wk=16", grey("lena_grey_256.u8", 256), 8, 16, False),
             ("LenaGrey 256^2 B=8 full pool (wk=61)", grey("lena_grey_256.u8", 256), 8, 61, False),
             ("LenaGrey 256^2 B=4 full pool (wk=125)", grey("lena_grey_256.u8", 256), 4, 125, False),
             ("Lena64 64^2 B=8 wk=2", grey("lena64.u8", 64), 8, 2, False),
             ("LenaColored 256^2 B=8 wk=2 (RGB)", rgb("lena_colored_256.rgb", 256), 8, 2, True),
             ("LenaColored 256^2 B=8 full pool (RGB)", rgb("lena_colored_256.rgb", 256), 8, 61, True),
             ("LenaColored 256^2 B=4 full pool (RGB)", rgb("lena_colored_256.rgb", 256), 4, 125, True)]
    print(f"{'case':46s} {'GPU encode':>11s} {'GPU decode':>11s} {'CPU encode':>11s} {'CPU decode':>11s}  evals")
    for name, img, B, wk, is_rgb in cases:
        H, W = img.shape
        for _ in range(3):
            info, q = h.encode(img, B, wk, rgb=is_rgb)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            info, q = h.encode(img, B, wk, rgb=is_rgb)
        t_enc = (time.perf_counter() - t0) / reps
        h.decode(q, W, H, B, wk, is_rgb)
        t0 = time.perf_counter()
        for _ in range(reps):
            dec, avg, it = h.decode(q, W, H, B, wk, is_rgb)
        t_dec = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        oinfo = O.encode(img, B, wk, rgb=is_rgb)
        c_enc = time.perf_counter() - t0
        stream = O.write_data(oinfo, W, H, B, wk, rgb=is_rgb)
        t0 = time.perf_counter()
        O.decode(stream)
        c_dec = time.perf_counter() - t0
        assert fic.stream_write(q, W, H, B, wk, rgb=is_rgb) == stream
        evals = (W // B) * (H // B) * wk * wk
        print(f"{name:46s} {t_enc * 1e3:9.3f}ms {t_dec * 1e3:9.3f}ms {c_enc * 1e3:9.1f}ms {c_dec * 1e3:9.1f}ms  {evals:.3g}")
    h.close()


if __name__ == "__main__":
    main()
