"""Decoder sweeps at W x W (B = 8 unless given), device-resident codes and image.  Prints the library's device time
per call and per sweep; run it under ncu to capture the sweep kernels (--cache-control none: the planes of a 4096^2
decode stay in L2 from sweep to sweep, which a flushed capture hides).

    python tools/decode_profile.py [W] [B] [grey|rgb] [wk: 2 | full]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractal_image_compression_b200 as fic  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rgb = len(sys.argv) > 3 and sys.argv[3] == "rgb"
wk = 2 * W // B - 3 if len(sys.argv) > 4 and sys.argv[4] == "full" else 2   # full pool: every range points anywhere
h = fic.Handle(0)
planes = np.stack([fic.synth.structured(W, W, s) for s in (1, 2, 3)]) if rgb else fic.synth.structured(W, W, 1)
info, q = h.encode_u8(planes, B, wk)
d_q = torch.from_numpy(q).cuda()
d_out = torch.empty(planes.shape, dtype=torch.uint8, device="cuda")
h.set_stream(None)
best = 1e9
for rep in range(5):
    avg, it = h.decode_planes_dev(d_q.data_ptr(), W, W, B, wk, rgb, d_out.data_ptr())
    best = min(best, h.timings().total_ms)
C = 3 if rgb else 1
bound_us = (3.25 * W * W * C + 12 * (W // B) ** 2) / 6546.6e9 * 1e6
print(f"{it} sweeps, avgError {avg}, device {best:.4f} ms = {best / it * 1e3:.2f} us per sweep "
      f"(whole call / sweeps; 3.25 W H bytes at the measured copy bandwidth: {bound_us:.2f} us -> {bound_us / (best / it * 1e3):.3f})")
ref, ravg, rit = h.decode_u8(q, W, W, B, wk, rgb)
assert (d_out.cpu().numpy() == ref).all() and avg == ravg and it == rit
