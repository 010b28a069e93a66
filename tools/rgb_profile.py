"""One RGB full-pool encode on each engine at a size ncu can replay quickly (run under `ncu -k regex:...`)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import fractal_image_compression_b200 as fic  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
engine = sys.argv[3] if len(sys.argv) > 3 else "umma"
p = np.stack([fic.synth.structured(size, size, s) for s in (1, 2, 3)], -1).astype(np.uint32)
img = (np.uint32(0xFF000000) | (p[..., 0] << np.uint32(16)) | (p[..., 1] << np.uint32(8)) | p[..., 2]).view(np.int32)
h = fic.Handle(0)
h.set_engine(fic.FIC_ENGINE_UMMA if engine == "umma" else fic.FIC_ENGINE_DIRECT)
for _ in range(2):
    h.encode(img, B, 2 * size // B - 3, rgb=True)
t = h.timings()
print(f"{engine} {size} B={B}: total {t.total_ms:.3f} ms, search {t.search_ms:.3f} ms, kernel {t.kernel_ms:.3f} ms")
h.close()
