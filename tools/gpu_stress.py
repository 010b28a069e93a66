"""Randomised cross-check of the two search engines on a GPU (no oracle: sizes it could not finish).

    python tools/gpu_stress.py [cases] [seed]

For every case a random image (noise / low contrast / binary / ramps / periodic / mixtures, random size, block size
and mode) is encoded with the full pool by the CUDA-core direct search (oracle-exact on every small case of the
test-suite) and by the tensor-core search with both instruction kinds; every code must agree bit for bit.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractal_image_compression_b200 as fic  # noqa: E402


def plane(rng, W, H, kind):
    if kind == 0:
        return rng.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == 1:
        return (120 + rng.integers(0, 3, (H, W))).astype(np.uint8)
    if kind == 2:
        return (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
    if kind == 3:
        y, x = np.mgrid[0:H, 0:W]
        return ((x * 5 + y * 3 + rng.integers(0, 8, (H, W))) % 256).astype(np.uint8)
    if kind == 4:   # periodic tiles: masses of identical domains (ties, flag-list overflow)
        t = rng.integers(0, 256, (16, 16), dtype=np.uint8)
        p = np.tile(t, (H // 16 + 1, W // 16 + 1))[:H, :W].copy()
        p[::48, ::48] ^= 1
        return p
    if kind == 5:   # black with a few bright pixels: blocks of mean 0 (the kind::i8 sign flip)
        p = np.zeros((H, W), np.uint8)
        m = rng.random((H, W)) < 0.002
        p[m] = rng.integers(200, 256, int(m.sum()), dtype=np.uint8)
        p[: H // 2, : W // 2] = rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8)
        return p
    a, b = plane(rng, W, H, 0), plane(rng, W, H, 3)   # half noise, half ramps
    a[:, W // 2:] = b[:, W // 2:]
    return a


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    h = fic.Handle(0)
    bad = 0
    for n in range(cases):
        B = int(rng.choice([4, 8, 16]))
        r = int(rng.integers(3, {4: 96, 8: 80, 16: 48}[B]))
        W = r * B
        wk = 2 * r - 3
        kind = int(rng.integers(0, 7))
        iso = bool(rng.random() < 0.25)
        rgb = not iso and B != 16 and bool(rng.random() < 0.35)   # RGB tensor path: kind::f16, B = 4 or 8
        mode = fic.FIC_MODE_GREY_ISO if iso else (fic.FIC_MODE_RGB if rgb else fic.FIC_MODE_GREY)
        if rgb:
            # channels equal (largest |gR|: the rows the CUDA-core kernel keeps), independent, or of mixed kinds
            how = int(rng.integers(0, 3))
            first = plane(rng, W, W, kind)
            v = [first if how == 0 else plane(rng, W, W, kind if how == 1 else int(rng.integers(0, 7))) for _ in range(3)]
            v = [a.astype(np.uint32) for a in v]
            img = (np.uint32(0xFF000000) | (v[0] << np.uint32(16)) | (v[1] << np.uint32(8)) | v[2]).view(np.int32)
        else:
            img = fic.synth.grey_to_argb(plane(rng, W, W, kind))
        h.set_engine(fic.FIC_ENGINE_DIRECT)
        i0, q0 = h.encode(img, B, wk, rgb=mode)
        h.set_engine(fic.FIC_ENGINE_UMMA)
        for mma in ((fic.FIC_UMMA_KIND_F16,) if rgb else (fic.FIC_UMMA_KIND_I8, fic.FIC_UMMA_KIND_F16)):
            h.set_umma_kind(mma)
            i1, q1 = h.encode(img, B, wk, rgb=mode)
            same = (q0 == q1).all() and np.array_equal(i0.view(np.uint32)[~np.isnan(i0)], i1.view(np.uint32)[~np.isnan(i1)])
            if not same:
                bad += 1
                print(f"MISMATCH case {n}: W={W} B={B} kind={kind} iso={iso} rgb={rgb} mma={mma} rows={(q0 != q1).any(1).sum()}")
        h.set_umma_kind(fic.FIC_UMMA_KIND_AUTO)
        h.set_engine(fic.FIC_ENGINE_AUTO)
    print(f"STRESS {'PASS' if bad == 0 else 'FAIL'}: {cases} cases, {bad} mismatches")
    h.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
