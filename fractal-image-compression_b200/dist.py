"""Multi-GPU encode: one process per GPU, torch.distributed for the plumbing.

Range blocks are independent (the reference's loop FC:125-159 carries no state but the
output index), so an encode shards by contiguous range rows.  Every rank needs the whole
image because the domain pool spans it: rank 0 broadcasts the 8-bit plane(s) once (NCCL
over NVLink on GPUs), every rank builds the full pool and searches its own rows through
the C ABI (fic_encode_planes_dev), and the codes (12 B / range grey, 20 B RGB) are
gathered to rank 0.  No other exchange exists on this path; results are byte-identical
for every world size.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def partition_range_rows(rph: int, rpw: int, world: int) -> list[tuple[int, int]]:
    """Contiguous split of the rph range rows over `world` ranks -> [(j0, j1)] in range-block units."""
    out = []
    base, extra = divmod(rph, world)
    row = 0
    for r in range(world):
        rows = base + (1 if r < extra else 0)
        out.append((row * rpw, (row + rows) * rpw))
        row += rows
    return out


class ShardedEncoder:
    """Sharded encode over the default (or a given) process group.

    worker(planes, rgb, W, H, B, wk, j0, j1) -> (info[NR,S] float32, q[NR,S] int32) tensors on
    the planes' device with rows [j0, j1) filled.  The default worker is the CUDA library;
    CPU tests inject one (the product never runs a CPU search).
    """

    def __init__(self, group=None, worker: Callable | None = None, handle=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._handle = handle
        self.worker = worker or self._cuda_worker
        self._bufs = {}

    def _cuda_worker(self, planes: torch.Tensor, rgb, W, H, B, wk, j0, j1):
        if not planes.is_cuda:
            raise RuntimeError("libfic_b200 needs CUDA tensors: there is no CPU fallback")
        if self._handle is None:
            from .codec import Handle

            self._handle = Handle(planes.device.index or 0)
        S = (3, 5, 4)[int(rgb)]   # grey, RGB, grey + isometry index (extension)
        NR = (W // B) * (H // B)
        key = (NR, S, planes.device)
        if key not in self._bufs:   # scratch, reused by every encode of this shape; only rows [j0, j1) are meaningful
            self._bufs[key] = (torch.empty((NR, S), dtype=torch.float32, device=planes.device),
                               torch.empty((NR, S), dtype=torch.int32, device=planes.device))
        info, q = self._bufs[key]
        self._handle.set_stream(torch.cuda.current_stream(planes.device).cuda_stream)
        self._handle.encode_planes_dev(planes.data_ptr(), int(rgb), W, H, B, wk, j0, j1, info.data_ptr(), q.data_ptr())
        return info, q

    def encode(self, planes: torch.Tensor | None, rgb, W: int, H: int, B: int, wk: int,
               device: torch.device | str = "cpu"):
        """planes: uint8 [C, H, W] on rank 0 (ignored elsewhere).  Returns (info, q) tensors on rank 0
        (on `device`), None on the other ranks."""
        C = 3 if int(rgb) == 1 else 1
        S = (3, 5, 4)[int(rgb)]
        rpw, rph = W // B, H // B
        NR = rpw * rph
        if self.rank == 0:
            buf = planes.to(device).reshape(C, H, W).contiguous()
        else:
            buf = torch.empty((C, H, W), dtype=torch.uint8, device=device)
        if self.world > 1:
            dist.broadcast(buf, src=0, group=self.group)
        parts = partition_range_rows(rph, rpw, self.world)
        j0, j1 = parts[self.rank]
        info, q = self.worker(buf, rgb, W, H, B, wk, j0, j1)
        if self.world == 1:
            # the default worker's buffers are scratch that the next encode overwrites: hand out copies
            return info.clone(), q.clone()
        # gather equal-sized (padded) row slices to rank 0
        maxrows = max(b - a for a, b in parts)
        send = torch.zeros((maxrows, 2 * S), dtype=torch.int32, device=buf.device)
        send[: j1 - j0, :S] = info[j0:j1].view(torch.int32)
        send[: j1 - j0, S:] = q[j0:j1]
        recv = [torch.empty_like(send) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(send, recv, dst=0, group=self.group)
        if self.rank != 0:
            return None
        out_info = torch.empty((NR, S), dtype=torch.float32, device=buf.device)
        out_q = torch.empty((NR, S), dtype=torch.int32, device=buf.device)
        for r, (a, b) in enumerate(parts):
            out_info[a:b] = recv[r][: b - a, :S].view(torch.float32)
            out_q[a:b] = recv[r][: b - a, S:]
        return out_info, out_q


def argb_to_planes(argb: np.ndarray, rgb: bool) -> np.ndarray:
    """int32 ARGB [H, W] -> uint8 [C, H, W] (grey keeps the red channel, FC:596)."""
    u = np.ascontiguousarray(argb).view(np.uint32)
    if not rgb:
        return ((u >> 16) & 0xFF).astype(np.uint8)[None]
    return np.stack([(u >> 16) & 0xFF, (u >> 8) & 0xFF, u & 0xFF]).astype(np.uint8)
