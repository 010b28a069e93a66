"""What ONE rank of the N-GPU run executes, as a single process (for ncu, which must not wrap a multi-rank command):
the full pool of the 8192^2 image and the range rows of rank `r` of `n`.

    python tools/rank_shard_profile.py [n] [r] [size]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractal_image_compression_b200 as fic  # noqa: E402
from fractal_image_compression_b200.dist import partition_range_rows  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
B = 8
rpw = W // B
wk = 2 * rpw - 3
plane = fic.synth.structured(W, W, 1)
j0, j1 = partition_range_rows(rpw, rpw, n)[r]
h = fic.Handle(0)
info = np.zeros((rpw * rpw, 3), np.float32)
q = np.zeros((rpw * rpw, 3), np.int32)
for _ in range(2):
    h.encode_u8(plane, B, wk, range_begin=j0, range_end=j1, info=info, q=q)
t = h.timings()
print(f"rank {r}/{n} of {W}x{W}: ranges [{j0},{j1}) total {t.total_ms:.3f} ms (h2d {t.h2d_ms:.3f} pool {t.pool_ms:.3f} search {t.search_ms:.3f} "
      f"kernel {t.kernel_ms:.3f} solve {t.solve_ms:.3f} d2h {t.d2h_ms:.3f})")
