"""Host-side logic: synthetic generators, RasterImage, sharding, and the world_size-2
(gloo, CPU) run of the multi-GPU plumbing with an injected per-rank worker."""
import hashlib
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_synth_is_deterministic(fic):
    n = fic.synth.noise(64, 32, seed=1)
    s = fic.synth.structured(64, 32, seed=1)
    assert n.shape == (32, 64) and n.dtype == np.uint8
    assert hashlib.sha256(n.tobytes()).hexdigest()[:16] == hashlib.sha256(fic.synth.noise(64, 32, 1).tobytes()).hexdigest()[:16]
    # known answers (integer-only formula; any port must reproduce these bytes)
    assert n[0, :8].tolist() == [1, 159, 7, 15, 106, 167, 221, 63]
    assert s[3, :8].tolist() == [3, 3, 15, 21, 18, 24, 21, 22]
    assert hashlib.sha256(s.tobytes()).hexdigest()[:16] == "6b7d7858bd4bd242"
    assert hashlib.sha256(n.tobytes()).hexdigest()[:16] == "66790abde9810744"
    assert fic.synth.noise(64, 32, 2)[0, 0] != n[0, 0] or fic.synth.noise(64, 32, 2)[0, 1] != n[0, 1]


def test_raster_image(fic):
    im = fic.RasterImage(4, 2)
    assert im.argb.shape == (2, 4) and (im.argb.view(np.uint32) == 0xFFA0A0A0).all()
    g = fic.RasterImage.from_grey(np.arange(8, dtype=np.uint8).reshape(2, 4))
    assert fic.FractalCompression.isGreyScale(g) and (g.red() == np.arange(8).reshape(2, 4)).all()
    c = fic.RasterImage.from_rgb(np.arange(24, dtype=np.uint8).reshape(2, 4, 3))
    assert not fic.FractalCompression.isGreyScale(c) and (c.rgb() == np.arange(24).reshape(2, 4, 3)).all()


def test_partition(fic):
    from fractal_image_compression_b200.dist import partition_range_rows

    for rph, rpw, world in [(32, 32, 1), (32, 32, 2), (32, 32, 8), (7, 5, 4), (3, 9, 8)]:
        parts = partition_range_rows(rph, rpw, world)
        assert parts[0][0] == 0 and parts[-1][1] == rph * rpw
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all((b - a) % rpw == 0 for a, b in parts)
        sizes = [(b - a) // rpw for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import fractal_image_compression_b200 as fic
from fractal_image_compression_b200.dist import ShardedEncoder, argb_to_planes
from oracle import oracle as O
dist.init_process_group("gloo")
rank = dist.get_rank()
W = H = 64; B = 8; wk = 13
plane = fic.synth.structured(W, H, 5)
grey = fic.synth.grey_to_argb(plane)
v = [fic.synth.structured(W, H, s).astype(np.uint32) for s in (5, 6, 7)]
colour = (np.uint32(0xFF000000) | (v[0] << np.uint32(16)) | (v[1] << np.uint32(8)) | v[2]).view(np.int32)
def to_argb(planes):
    p = planes.numpy().astype(np.uint32)
    if p.shape[0] == 1:
        return fic.synth.grey_to_argb(planes[0].numpy())
    return (np.uint32(0xFF000000) | (p[0] << np.uint32(16)) | (p[1] << np.uint32(8)) | p[2]).view(np.int32)
def worker(planes, mode, W, H, B, wk, j0, j1):     # test double: the CPU oracle computes this rank's rows
    iso, rgb = mode == fic.FIC_MODE_GREY_ISO, mode == fic.FIC_MODE_RGB
    info = O.encode(to_argb(planes), B, wk, rgb=rgb, range_begin=j0, range_end=j1, iso=iso)
    q = np.frombuffer(O.write_data(info, W, H, B, wk, rgb=rgb, iso=iso)[20:], ">i4").astype(np.int32).reshape(-1, info.shape[1])
    return torch.from_numpy(info), torch.from_numpy(q.copy())
enc = ShardedEncoder(worker=worker)
# the reference's grey and RGB codes, and the isometry extension
for mode in (fic.FIC_MODE_GREY, fic.FIC_MODE_RGB, fic.FIC_MODE_GREY_ISO):
    iso, rgb = mode == fic.FIC_MODE_GREY_ISO, mode == fic.FIC_MODE_RGB
    argb = colour if rgb else grey
    planes = torch.from_numpy(argb_to_planes(argb, rgb)) if rank == 0 else None
    out = enc.encode(planes, mode, W, H, B, wk, device="cpu")
    if rank == 0:
        info, q = out
        full = O.encode(argb, B, wk, rgb=rgb, iso=iso)
        assert info.numpy().tobytes() == full.tobytes(), "sharded codes differ from the single-process encode"
        want = O.write_data(full, W, H, B, wk, rgb=rgb, iso=iso)
        assert fic.stream_write(q.numpy(), W, H, B, wk, rgb=mode) == want
        print("SHARDED_OK", mode)
    else:
        assert out is None
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert all(f"SHARDED_OK {m}" in r.stdout for m in (0, 1, 2))


def test_rgb_covariance_partial_sums_stay_exact_for_b8():
    """The bound behind the RGB tensor path: for B <= 8 every partial sum of |gR_i * gD_i| is below
    2^24, whatever the pixels.  Checked on adversarial blocks (0/255 patterns, aligned and anti-aligned, all channels
    equal) and random ones, with the integer means the reference uses."""
    rng = np.random.default_rng(5)
    n = 64
    bound = 9 * n * (127.5 ** 2 + 1)
    assert bound < 2 ** 24
    worst = 0

    def centred(block):                      # block: (3, n) ints in [0, 255] -> sum over channels of v - floor(mean)
        return (block - block.sum(1, keepdims=True) // n).sum(0)

    blocks = []
    for k in range(0, n + 1, 4):             # k pixels at 255, the rest 0, all channels equal
        b = np.zeros((3, n), np.int64)
        b[:, :k] = 255
        blocks.append(b)
    for _ in range(200):
        kind = rng.integers(0, 3)
        if kind == 0:
            blocks.append(rng.integers(0, 256, (3, n)))
        elif kind == 1:
            blocks.append(rng.integers(0, 2, (3, n)) * 255)
        else:
            blocks.append(np.repeat(rng.integers(0, 2, (1, n)) * 255, 3, 0))
    g = np.stack([centred(b) for b in blocks])            # (m, n)
    assert np.abs(g).max() <= 765
    l2 = np.sqrt((g.astype(np.float64) ** 2).sum(1))
    assert l2.max() <= 3 * np.sqrt(n * (127.5 ** 2 + 1)) + 1e-9
    total = np.abs(g)[:, None, :] * np.abs(g)[None, :, :]  # |gR_i * gD_i| for every pair of blocks
    worst = int(total.sum(-1).max())
    assert worst <= bound and worst > 9_000_000            # the bound is nearly attained, and it holds
