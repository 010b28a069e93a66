"""GPU parity tests of the round-2 entries: the fused one-launch windowed encode, the 8-bit host entries, the
batched decoder (device-side convergence, images above 2^24 pixels, plane outputs) and the multi-GPU handle.
Everything goes through the C ABI (ctypes) and is compared with the CPU oracle or, for entries that only change the
data format, with the reference-shaped entry bit for bit."""
import os

import numpy as np
import pytest

from conftest import to_argb_grey
from test_gpu_parity import assert_codes_equal, float_bits_equal

pytestmark = pytest.mark.gpu


def _rgb_argb(planes):
    v = [p.astype(np.uint32) for p in planes]
    return (np.uint32(0xFF000000) | (v[0] << np.uint32(16)) | (v[1] << np.uint32(8)) | v[2]).view(np.int32)


# ---------------------------------------------------------------- fused windowed encode (FIC_ENGINE_FUSED)

@pytest.mark.parametrize("name,B,wk", [("lena_grey", 8, 2), ("lena_grey", 8, 16), ("lena_grey", 16, 16), ("lena_grey", 16, 1),
                                       ("lena_grey", 4, 16), ("lena_grey", 4, 3), ("lena64", 8, 2), ("lena64", 8, 13),
                                       ("lena64", 4, 8), ("lena64", 16, 5)])
def test_fused_grey_equals_oracle(fic, handle, oracle, request, name, B, wk):
    """The one-launch encode (decimate + stats + search + solve + quantise per range block) against the oracle,
    forced (FIC_ENGINE_FUSED) and as what AUTO picks for the reference's GUI windows."""
    img = request.getfixturevalue(name)
    H, W = img.shape
    oinfo = oracle.encode(img, B, wk, nthreads=4)
    ostream = oracle.write_data(oinfo, W, H, B, wk)
    handle.set_engine(fic.FIC_ENGINE_FUSED)
    try:
        info, q = handle.encode(img, B, wk, rgb=False)
        assert handle.timings().engine == fic.FIC_ENGINE_FUSED
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
    assert_codes_equal(info, q, oinfo, ostream, 3)
    info, q = handle.encode(img, B, wk, rgb=False)
    assert handle.timings().engine == fic.FIC_ENGINE_FUSED   # Lena64 full pool (wk = 13) is below the tensor-path threshold
    assert_codes_equal(info, q, oinfo, ostream, 3)


@pytest.mark.parametrize("B,wk", [(8, 2), (4, 4), (16, 2), (8, 7), (4, 16)])
def test_fused_rgb_equals_oracle(fic, handle, oracle, lena_colored, B, wk):
    handle.set_engine(fic.FIC_ENGINE_FUSED)
    try:
        info, q = handle.encode(lena_colored, B, wk, rgb=True)
        assert handle.timings().engine == fic.FIC_ENGINE_FUSED
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
    oinfo = oracle.encode(lena_colored, B, wk, rgb=True, nthreads=4)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, 256, 256, B, wk, rgb=True), 5)


def test_fused_stream_equals_the_references_own_file(fic, handle, lena_colored):
    """The reference's own RGB stream (unknown.run, B = 8, wk = 2) from the fused engine, byte for byte."""
    from conftest import GOLD

    with open(os.path.join(GOLD, "unknown_run.bin"), "rb") as f:
        want = f.read()
    _, q = handle.encode(lena_colored, 8, 2, rgb=True)
    assert handle.timings().engine == fic.FIC_ENGINE_FUSED
    assert fic.stream_write(q, 256, 256, 8, 2, True) == want


@pytest.mark.parametrize("W,H,B,wk,rgb", [(96, 64, 8, 4, False), (64, 96, 8, 3, False), (160, 48, 16, 2, False), (96, 64, 4, 9, True),
                                          (48, 128, 8, 2, True)])
def test_fused_non_square(fic, handle, oracle, W, H, B, wk, rgb):
    """Landscape images hit the reference's `x + 1 >= height` decimation slip (FC:993 / FC:940); the fused kernel's
    in-CTA decimation reproduces it like the plane kernels do."""
    if rgb:
        img = _rgb_argb([fic.synth.structured(W, H, s) for s in (1, 2, 3)])
    else:
        img = fic.synth.grey_to_argb(fic.synth.noise(W, H, 4))
    assert img.shape == (H, W)
    info, q = handle.encode(img, B, wk, rgb=rgb)
    assert handle.timings().engine == fic.FIC_ENGINE_FUSED
    oinfo = oracle.encode(img, B, wk, rgb=rgb, nthreads=4)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk, rgb=rgb), 5 if rgb else 3)


def test_fused_flat_and_slices(fic, handle, oracle):
    """Flat content (0/0 = NaN contrast, FC:634) and range-row slices through the fused engine."""
    p = np.full((128, 128), 77, np.uint8)
    p[40:60, 30:90] = 200
    img = fic.synth.grey_to_argb(p)
    oinfo = oracle.encode(img, 8, 4, nthreads=2)
    info = np.zeros_like(oinfo)
    q = np.zeros(oinfo.shape, np.int32)
    for j0, j1 in [(0, 100), (100, 101), (101, 256)]:
        handle.encode(img, 8, 4, rgb=False, range_begin=j0, range_end=j1, info=info, q=q)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, 128, 128, 8, 4), 3)
    assert np.isnan(oinfo[:, 1]).any()


def test_engines_agree_windowed_1024(fic, handle):
    """Fused == multi-kernel direct path on a 1024^2 image at the GUI's largest window (both oracle-exact on the
    small cases; this is the size-independent cross-check)."""
    W = 1024
    img = fic.synth.grey_to_argb(fic.synth.structured(W, W, 9))
    out = {}
    for eng in (fic.FIC_ENGINE_DIRECT, fic.FIC_ENGINE_FUSED):
        handle.set_engine(eng)
        try:
            out[eng] = handle.encode(img, 8, 16, rgb=False)
            assert handle.timings().engine == eng
        finally:
            handle.set_engine(fic.FIC_ENGINE_AUTO)
    (i1, q1), (i2, q2) = out.values()
    assert (q1 == q2).all() and float_bits_equal(i1, i2)


# ---------------------------------------------------------------- 8-bit host entries

@pytest.mark.parametrize("W,B,wk", [(256, 8, 2), (256, 8, 61), (512, 8, 125)])
def test_u8_entries_equal_argb_entries(fic, handle, W, B, wk):
    """fic_encode_grey_u8 / fic_encode_rgb_planes == fic_encode_grey / fic_encode_rgb (fused, direct and tensor-core
    engines), and fic_decode_u8 == the red / R, G, B channels of fic_decode."""
    planes = np.stack([fic.synth.structured(W, W, s) for s in (1, 2, 3)])
    i1, q1 = handle.encode(fic.synth.grey_to_argb(planes[0]), B, wk, rgb=False)
    e1 = handle.timings().engine
    i2, q2 = handle.encode_u8(planes[0], B, wk)
    assert handle.timings().engine == e1
    assert (q1 == q2).all() and float_bits_equal(i1, i2)
    dec, avg, it = handle.decode(q1, W, W, B, wk, False)
    dec8, avg8, it8 = handle.decode_u8(q1, W, W, B, wk, False)
    assert (((dec.view(np.uint32) >> 16) & 0xFF) == dec8).all() and avg == avg8 and it == it8
    if B != 16:
        i1, q1 = handle.encode(_rgb_argb(planes), B, wk, rgb=True)
        i2, q2 = handle.encode_u8(planes, B, wk)
        assert (q1 == q2).all() and float_bits_equal(i1, i2)
        dec, avg, it = handle.decode(q1, W, W, B, wk, True)
        dec8, avg8, it8 = handle.decode_u8(q1, W, W, B, wk, True)
        du = dec.view(np.uint32)
        assert (np.stack([(du >> 16) & 0xFF, (du >> 8) & 0xFF, du & 0xFF]) == dec8).all() and avg == avg8 and it == it8


def test_decode_planes_dev(fic, handle):
    """Device codes in (as fic_encode_planes_dev leaves them), device planes out == the host entries."""
    import torch

    W, B = 512, 8
    wk = 2 * W // B - 3
    p = fic.synth.structured(W, W, 2)
    d_p = torch.from_numpy(p).cuda()
    NR = (W // B) ** 2
    d_q = torch.empty((NR, 3), dtype=torch.int32, device="cuda")
    d_out = torch.empty((W, W), dtype=torch.uint8, device="cuda")
    handle.set_stream(None)
    handle.encode_planes_dev(d_p.data_ptr(), 0, W, W, B, wk, 0, NR, None, d_q.data_ptr())
    handle.sync()
    avg, it = handle.decode_planes_dev(d_q.data_ptr(), W, W, B, wk, False, d_out.data_ptr())
    ref, ravg, rit = handle.decode_u8(d_q.cpu().numpy(), W, W, B, wk, False)
    assert (d_out.cpu().numpy() == ref).all() and avg == ravg and it == rit


# ---------------------------------------------------------------- batched decoder vs oracle

@pytest.mark.parametrize("max_iters", [1, 3, 8, 9, 50])
def test_decode_iteration_limits(fic, handle, oracle, lena_grey, max_iters):
    """The decoder enqueues sweeps in batches of eight and reads the device-side state once per batch: every limit
    around the batch boundary must give the image, avgError and sweep count of a sweep-by-sweep loop.  (The oracle
    has the reference's fixed limit of 50, so shorter limits are compared with a decode that is cut by hand:
    identical prefixes of the same Jacobi iteration.)"""
    info = oracle.encode(lena_grey, 8, 4, nthreads=4)
    stream = oracle.write_data(info, 256, 256, 8, 4)
    q = np.frombuffer(stream[20:], ">i4").astype(np.int32).reshape(-1, 3)
    want_img, want_avg, want_it = oracle.decode(stream)
    img, avg, it = handle.decode(q, 256, 256, 8, 4, False, max_iters=max_iters)
    if max_iters >= want_it:
        assert it == want_it and avg == want_avg and (img == want_img).all()
    else:
        assert it == max_iters and not avg < 1      # not converged: the last sweep's value is kept (FC:416)
        img2, avg2, it2 = handle.decode(q, 256, 256, 8, 4, False, max_iters=max_iters)
        assert (img2 == img).all() and avg2 == avg and it2 == it


@pytest.mark.parametrize("carry", [0.0, 0.7711792, 3.5e7])
def test_decode_carry_in(fic, handle, oracle, lena_grey, carry):
    """FractalCompression.avgError is static and never reset (FC:20): the first sweep starts from whatever the
    previous decode left."""
    info = oracle.encode(lena_grey, 8, 2, nthreads=4)
    stream = oracle.write_data(info, 256, 256, 8, 2)
    q = np.frombuffer(stream[20:], ">i4").astype(np.int32).reshape(-1, 3)
    want_img, want_avg, want_it = oracle.decode(stream, avg_error_in=carry)
    img, avg, it = handle.decode(q, 256, 256, 8, 2, False, avg_error=carry)
    assert it == want_it and avg == want_avg and (img == want_img).all()


def test_decode_above_2_24_pixels(fic, handle, oracle):
    """4096 x 4104 = 16.8 M pixels > 2^24: the float running sum of a converging sweep passes 2^24, where binary32
    addition rounds, so the sweep's value depends on the accumulation order.  The GPU replays it in loop order only
    where it must (k_sweep_finish) and has to reproduce image, avgError and sweep count of the oracle."""
    W, H, B, wk = 4096, 4104, 8, 2
    p = fic.synth.structured(W, H, 3)
    img = fic.synth.grey_to_argb(p)
    info, q = handle.encode(img, B, wk, rgb=False)
    oinfo = oracle.encode(img, B, wk, nthreads=os.cpu_count() or 1)
    stream = oracle.write_data(oinfo, W, H, B, wk)
    assert (q == np.frombuffer(stream[20:], ">i4").astype(np.int32).reshape(-1, 3)).all()
    want_img, want_avg, want_it = oracle.decode(stream)
    out, avg, it = handle.decode_u8(q, W, H, B, wk, False)
    assert it == want_it and avg == want_avg
    assert (out == ((want_img.view(np.uint32) >> 16) & 0xFF)).all()
    assert handle.timings().total_ms < 200.0   # no per-sweep serial cliff: a sweep of this size takes ~30 us


@pytest.mark.parametrize("W,H,B,wk,rgb", [
    (264, 136, 8, 5, False),    # W % 16 == 8: decimated rows start at odd multiples of 4; landscape: the FC:993 tap
    (272, 144, 8, 5, False),    # W % 16 == 0, grey: the decoder loop over the row-pair interleaved plane; landscape
    (136, 264, 8, 4, True),     # portrait RGB, plain plane
    (144, 272, 8, 4, True),     # portrait RGB, W % 16 == 0: interleaved plane
    (256, 256, 16, 4, False),   # blockgroesse 16 (interleaved plane: domain columns are multiples of 4)
    (272, 144, 16, 3, True),
    (64, 64, 8, 13, False),     # whole pool as the window
    (128, 128, 4, 6, False),    # blockgroesse 4: a strip spans two range blocks, explicit start image
    (100, 60, 4, 3, False),     # W % 8 != 0: the quad kernel
])
@pytest.mark.parametrize("max_iters", [50, 2])
def test_decode_random_codes(fic, handle, oracle, W, H, B, wk, rgb, max_iters):
    """Random codes reach every domain position (every alignment of a domain row inside the decimated plane), on
    image shapes that take each of the sweep kernels.  Image, avgError and sweep count must be the oracle's -- the
    first sweep never reads the constant start image (FC:360) it starts from; max_iters = 2 runs the sweeps that also
    write the per-pixel changes (the last allowed sweep and a first sweep with a carried-in avgError are folded by
    the replaying k_sweep_finish)."""
    rng = np.random.default_rng(W * 131 + H * 17 + B + wk)
    NR = (W // B) * (H // B)
    S = 5 if rgb else 3
    q = np.empty((NR, S), np.int32)
    q[:, 0] = rng.integers(0, wk * wk, NR)
    if rgb:
        q[:, 1] = rng.integers(-1200000, 1200000, NR)
        q[:, 2:4] = rng.integers(-100, 300, (NR, 2)) * 100000 + rng.integers(0, 100000, (NR, 2))
        q[:, 4] = rng.integers(-100, 300, NR)   # the blue offset travels as a plain int (FC:254)
    else:
        q[:, 1] = rng.integers(-120, 120, NR)
        q[:, 2] = rng.integers(-100, 300, NR)
    q[: NR // 8, 1] = 0   # flat range blocks
    stream = fic.stream_write(q, W, H, B, wk, rgb)
    img, avg, it = handle.decode(q, W, H, B, wk, rgb, max_iters=max_iters)
    if max_iters == 50:
        want_img, want_avg, want_it = oracle.decode(stream)
        assert it == want_it and np.float32(avg) == np.float32(want_avg) and (img == want_img).all()
        planes, avg8, it8 = handle.decode_u8(q, W, H, B, wk, rgb)
        u = img.view(np.uint32)
        want8 = np.stack([(u >> 16) & 0xFF, (u >> 8) & 0xFF, u & 0xFF]) if rgb else ((u >> 16) & 0xFF)
        assert (planes == want8).all() and avg8 == avg and it8 == it
    else:
        a1 = handle.decode(q, W, H, B, wk, rgb, max_iters=1)
        b2 = handle.decode(q, W, H, B, wk, rgb, max_iters=2, avg_error=0.25)   # carry-in: the first sweep is replayed too
        assert it == 2 and a1[2] == 1 and (b2[0] == img).all()


def _seq_float_sum(terms, carry):
    """Sequential binary32 accumulation (np.add.accumulate is strictly left to right)."""
    x = np.concatenate([[np.float32(carry)], terms.astype(np.float32)])
    return np.add.accumulate(x, dtype=np.float32)[-1]


@pytest.mark.parametrize("kind", ["ones", "small", "uniform", "odd", "sparse", "large", "ramp", "ties4", "mixed"])
@pytest.mark.parametrize("count", [1 << 20, (1 << 22) + 4097, 1 << 24])
def test_avg_error_replay(fic, handle, kind, count):
    """The float running sum behind avgError (FC:407) through both replay forms: one warp below 2^22 terms, per-chunk
    transducers above.  The distributions put the sum into binades 2^24 .. 2^36 with every tie pattern (odd terms on a
    grid of 2, multiples of 2 on a grid of 4, ...), cross binades inside chunks, and carry fractions in."""
    rng = np.random.default_rng(count % 1000 + len(kind))
    hi = 3 * 255 * 255
    if kind == "ones":       # on a grid of 2 an added 1 is a tie that an even mantissa drops: the float sum stalls at 2^24
        t = np.ones(count, np.int64)
    elif kind == "small":
        t = rng.integers(0, 4, count)
    elif kind == "uniform":
        t = rng.integers(0, 256, count)
    elif kind == "odd":
        t = rng.integers(0, 8, count) * 2 + 1
    elif kind == "sparse":
        t = np.where(rng.random(count) < 0.01, rng.integers(0, hi + 1, count), 0)
    elif kind == "large":
        t = rng.integers(hi - 1000, hi + 1, count)
    elif kind == "ramp":
        t = (np.arange(count) * 37 % 1024) * (np.arange(count) // (count // 16) % 2 + 1)
    elif kind == "ties4":
        t = rng.integers(0, 64, count) * 4 + 2
    else:
        t = np.where(np.arange(count) < count // 3, rng.integers(0, 3, count), rng.integers(0, 70000, count))
    t = t.astype(np.int32)
    for carry in (0.0, 0.5863342, 3.5e7):
        got = handle.float_sum(t, carry)
        want = _seq_float_sum(t, carry)
        assert got == want, (kind, count, carry, float(got), float(want))


def test_decode_8192_square(fic, handle, oracle):
    """8192^2 = 2^26 pixels (BASELINE configs[3]'s image size): every sweep's float sum passes 2^24, the converging ones
    end between 2^24 and 2^26 where binary32 addition rounds.  Random codes (no encode needed); image, avgError and sweep
    count equal the oracle's, and no sweep falls back to a 2^26-step serial replay."""
    W = H = 8192
    B, wk = 8, 2
    rng = np.random.default_rng(8192)
    NR = (W // B) * (H // B)
    q = np.empty((NR, 3), np.int32)
    q[:, 0] = rng.integers(0, wk * wk, NR)
    q[:, 1] = rng.integers(-130, 130, NR)   # converges to avgError ~ 0.6: a float sum of ~ 4e7, two binades above 2^24
    q[:, 2] = rng.integers(-60, 280, NR)
    stream = fic.stream_write(q, W, H, B, wk, False)
    want_img, want_avg, want_it = oracle.decode(stream)
    assert 0.3 < want_avg < 1.0
    handle.decode_u8(q, W, H, B, wk, False)
    out, avg, it = handle.decode_u8(q, W, H, B, wk, False)
    assert it == want_it and avg == want_avg
    assert (out == ((want_img.view(np.uint32) >> 16) & 0xFF)).all()
    assert handle.timings().total_ms < 60.0   # 67 MB to the host + a dozen sweeps; the serial replay took 0.2 s per sweep


# ---------------------------------------------------------------- multi-GPU handle behind the C ABI

def _device_count():
    import torch

    return torch.cuda.device_count()


def test_multi_handle_single_device(fic, handle):
    """fic_create_multi over one device (no NCCL involved) == the single-device entries, all pixel formats."""
    m = fic.MultiHandle([0])
    try:
        W, B = 512, 8
        wk = 2 * W // B - 3
        planes = np.stack([fic.synth.structured(W, W, s) for s in (1, 2, 3)])
        argb = fic.synth.grey_to_argb(planes[0])
        i1, q1 = handle.encode(argb, B, wk, rgb=False)
        for px in (argb, planes[0]):
            i2, q2 = m.encode(px, B, wk)
            assert (q1 == q2).all() and float_bits_equal(i1, i2)
        assert m.range_slice(0) == (0, (W // B) ** 2)
        assert m.timings().engine == fic.FIC_ENGINE_UMMA and m.timings().total_ms > 0
        i1, q1 = handle.encode(_rgb_argb(planes), B, 2, rgb=True)
        i2, q2 = m.encode(planes, B, 2)
        assert (q1 == q2).all() and float_bits_equal(i1, i2)
        dec, avg, it = m.handle(0).decode(q1, W, W, B, 2, True)   # the decoder does not shard: per-device handle
        dec1, avg1, it1 = handle.decode(q1, W, W, B, 2, True)
        assert (dec == dec1).all() and avg == avg1 and it == it1
    finally:
        m.close()
    with pytest.raises(fic.FicError):
        fic.MultiHandle([0, 0])


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_handle_equals_single(fic, handle, oracle, n):
    """N devices, one NCCL broadcast of the planes, row slices straight into the caller's arrays == one device,
    byte for byte (grey full pool on the tensor cores, RGB, windowed) and == the oracle on a range sample."""
    if _device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    m = fic.MultiHandle(list(range(n)))
    try:
        W, B = 2048, 8
        wk = 2 * W // B - 3
        planes = np.stack([fic.synth.structured(W, W, s) for s in (1, 2, 3)])
        argb = fic.synth.grey_to_argb(planes[0])
        i1, q1 = handle.encode(argb, B, wk, rgb=False)
        i2, q2 = m.encode(argb, B, wk)
        assert (q1 == q2).all() and float_bits_equal(i1, i2)
        i3, q3 = m.encode(planes[0], B, wk)
        assert (q1 == q3).all() and float_bits_equal(i1, i3)
        slices = [m.range_slice(r) for r in range(n)]
        assert slices[0][0] == 0 and slices[-1][1] == (W // B) ** 2 and all(a[1] == b[0] for a, b in zip(slices, slices[1:]))
        rng = np.random.default_rng(3)
        ranges = np.unique(rng.integers(0, (W // B) ** 2, 48)).astype(np.int64)
        ref = oracle.encode_list(argb, B, wk, ranges, nthreads=os.cpu_count() or 1)
        assert float_bits_equal(i2[ranges], ref)
        i1, q1 = handle.encode(_rgb_argb(planes), B, wk, rgb=True)
        i2, q2 = m.encode(planes, B, wk)
        assert (q1 == q2).all() and float_bits_equal(i1, i2)
        i1, q1 = handle.encode(argb, B, 4, rgb=False)
        i2, q2 = m.encode(argb, B, 4)
        assert (q1 == q2).all() and float_bits_equal(i1, i2)
    finally:
        m.close()
