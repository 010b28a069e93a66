#!/usr/bin/env python
"""The multi-GPU path behind the C ABI (fic_create_multi / fic_multi_encode_*), one process, N GPUs:

    python tools/multi_abi_bench.py [N] [size] [B]        (defaults: all GPUs, 8192, 8)

Encodes the synthetic structured image through ONE multi handle -- host pixels in (int32 ARGB, the reference's
RasterImage.argb, and 8-bit planes), one upload, NCCL broadcast over NVLink, range rows sharded, code rows straight
into the caller's arrays -- times the call with the host clock (pinned buffers), and checks the result against the
single-device entry on a row sample and against the CPU oracle on random range blocks.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import fractal_image_compression_b200 as fic
    from oracle import oracle as O

    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    rpw = size // B
    wk = 2 * rpw - 3
    NR, ND = rpw * rpw, wk * wk
    plane = fic.synth.structured(size, size, 1)
    argb = fic.synth.grey_to_argb(plane)
    m = fic.MultiHandle(list(range(n)))
    h0 = m.handle(0)
    info = np.zeros((NR, 3), np.float32)
    q = np.zeros((NR, 3), np.int32)
    for a in (plane, argb, info, q):
        h0.pin(a)
    out = {"n_gpus": n, "workload": f"synthetic structured {size}x{size} grey, B={B}, widthKernel={wk} (full pool)", "ranges": NR, "domains": ND}
    for name, px in (("argb", argb), ("u8", plane)):
        for _ in range(2):
            m.encode(px, B, wk, info=info, q=q)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            m.encode(px, B, wk, info=info, q=q)
            ts.append((time.perf_counter() - t0) * 1e3)
        t = m.timings()
        out[f"e2e_ms_{name}"] = min(ts)
        out[f"e2e_gevals_per_s_{name}"] = NR * ND / (min(ts) * 1e-3) / 1e9
        out[f"device_stage_ms_{name}"] = {"h2d": t.h2d_ms, "pool": t.pool_ms, "search": t.search_ms, "kernel": t.kernel_ms, "solve": t.solve_ms,
                                          "d2h": t.d2h_ms, "total_slowest_device": t.total_ms}
    out["slices"] = [m.range_slice(r) for r in range(n)]
    # against the single-device entry on the first and last rows of every slice, and the oracle on random ranges
    single = fic.Handle(0)
    bad = 0
    for a, b in out["slices"]:
        for j0, j1 in ((a, min(a + 2 * rpw, b)), (max(b - 2 * rpw, a), b)):
            i1 = np.zeros_like(info)
            q1 = np.zeros_like(q)
            single.encode(argb, B, wk, rgb=False, range_begin=j0, range_end=j1, info=i1, q=q1)
            bad += int((q1[j0:j1] != q[j0:j1]).any(1).sum()) + int((i1[j0:j1].view(np.uint32) != info[j0:j1].view(np.uint32)).any(1).sum())
    out["rows_differing_from_single_device"] = bad
    rng = np.random.default_rng(5)
    ranges = np.unique(np.concatenate([[0, NR - 1], rng.integers(0, NR, 96)])).astype(np.int64)
    ref = O.encode_list(argb, B, wk, ranges, nthreads=os.cpu_count() or 1)
    got = info[ranges]
    same = ((got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))).all(1)
    out["parity_spot"] = {"ranges": int(len(ranges)), "mismatches": int((~same).sum()), "checker": "oracle"}
    for a in (plane, argb, info, q):
        h0.unpin(a)
    m.close()
    print(json.dumps(out))
    return 0 if bad == 0 and same.all() else 3


if __name__ == "__main__":
    sys.exit(main())
