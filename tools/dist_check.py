"""torchrun check: sharded encode over N GPUs (NCCL broadcast + gather) equals the single-GPU encode.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractal_image_compression_b200 as fic  # noqa: E402
from fractal_image_compression_b200.dist import ShardedEncoder, argb_to_planes  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for W, B, rgb in [(1024, 8, False), (512, 4, False), (256, 8, True), (1024, 8, True), (512, 8, fic.FIC_MODE_GREY_ISO)]:
    if rgb is True:
        planes_np = np.stack([fic.synth.structured(W, W, s) for s in (1, 2, 3)])
        wk = 2 if W == 256 else 2 * W // B - 3   # the reference's default window / the full pool (tensor cores)
    else:
        planes_np = fic.synth.structured(W, W, 3)[None]
        wk = 2 * W // B - 3
    enc = ShardedEncoder()
    planes = torch.from_numpy(planes_np) if rank == 0 else None
    out = enc.encode(planes, rgb, W, W, B, wk, device=dev)
    torch.cuda.synchronize()
    if rank == 0:
        info, q = out
        h = fic.Handle(local)
        if rgb is True:
            a = planes_np.astype(np.uint32)
            argb = (0xFF000000 | (a[0] << 16) | (a[1] << 8) | a[2]).view(np.int32)
        else:
            argb = fic.synth.grey_to_argb(planes_np[0])
        info1, q1 = h.encode(argb, B, wk, rgb=rgb)
        same = (q.cpu().numpy() == q1).all() and np.array_equal(info.cpu().numpy(), info1, equal_nan=True)
        print(f"W={W} B={B} rgb={rgb} world={world}: sharded == single-GPU: {same}")
        ok = ok and bool(same)
        h.close()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
