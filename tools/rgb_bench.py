"""Time the RGB full-pool encode (tensor-core path) on synthetic images; `direct` adds the CUDA-core engine."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import fractal_image_compression_b200 as fic  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    with_direct = "direct" in sys.argv
    p = np.stack([fic.synth.structured(size, size, s) for s in (1, 2, 3)], -1).astype(np.uint32)
    img = (0xFF000000 | (p[..., 0] << 16) | (p[..., 1] << 8) | p[..., 2]).astype(np.uint32).view(np.int32)
    wk = 2 * size // B - 3
    h = fic.Handle(0)
    res = {}
    for name, eng in [("umma", fic.FIC_ENGINE_UMMA)] + ([("direct", fic.FIC_ENGINE_DIRECT)] if with_direct else []):
        h.set_engine(eng)
        for rep in range(3 if name == "umma" else 1):
            t0 = time.perf_counter()
            info, q = h.encode(img, B, wk, rgb=True)
            dt = time.perf_counter() - t0
            tm = h.timings()
            print(f"{name} {size}x{size} B={B} rep {rep}: host {dt*1e3:.2f} ms, total {tm.total_ms:.2f} ms, search {tm.search_ms:.2f} ms, "
                  f"kernel {tm.kernel_ms:.2f} ms, evals/s {tm.search_evals / (tm.total_ms * 1e-3):.3e}", flush=True)
        res[name] = (info.copy(), q.copy())
    if with_direct:
        same = (res["umma"][1] == res["direct"][1]).all() and (res["umma"][0].view(np.uint32) == res["direct"][0].view(np.uint32)).all()
        print("engines agree:", bool(same))
    h.close()


main()
