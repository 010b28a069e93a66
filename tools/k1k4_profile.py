"""Runs the HBM-bound kernels once so that ncu can capture them: the fused windowed encode (reference defaults),
the multi-kernel pool builder + direct search (a window the fused path does not take), and the decoder sweeps.

    python tools/k1k4_profile.py [W]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractal_image_compression_b200 as fic  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = fic.synth.grey_to_argb(fic.synth.structured(W, W, 1))
h = fic.Handle(0)
h.set_engine(fic.FIC_ENGINE_FUSED)
info, q = h.encode(img, 8, 2, rgb=False)         # reference default window: the fused one-launch encode
t = h.timings()
print(f"encode wk=2 (engine {t.engine}): total {t.total_ms:.3f} ms (h2d {t.h2d_ms:.3f}, kernel {t.kernel_ms:.3f}, d2h {t.d2h_ms:.3f})")
h.set_engine(fic.FIC_ENGINE_DIRECT)              # K1 (decimate, stats) + direct search + solve
info, q = h.encode(img, 8, 16, rgb=False)
t = h.timings()
print(f"encode wk=16 direct: total {t.total_ms:.3f} ms (pool {t.pool_ms:.3f}, search {t.search_ms:.3f})")
h.set_engine(fic.FIC_ENGINE_AUTO)
out, avg, it = h.decode_u8(q, W, W, 8, 16, False)
t = h.timings()
print(f"decode: {it} sweeps, avgError {avg}, device {t.total_ms:.3f} ms -> {W * W * it / t.total_ms / 1e3:.1f} Mpixel/s per sweep")
