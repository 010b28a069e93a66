"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs, bit for bit -- codes are integer / index work, and the float a, b are produced by
the same IEEE operations as the reference's."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLD, to_argb_grey, to_argb_rgb

pytestmark = pytest.mark.gpu


def q_from_stream(s, S):
    return np.frombuffer(s[20:], ">i4").astype(np.int32).reshape(-1, S)


def float_bits_equal(a, b):
    """Bit-for-bit equality of float32 arrays, except that any NaN equals any NaN: the 0/0 of a flat
    winner (FC:634) has no observable payload in the reference (it is only ever cast to int -> 0)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and (na == nb).all() and (a.view(np.uint32)[~na] == b.view(np.uint32)[~nb]).all()


def assert_codes_equal(info, q, oinfo, ostream, S):
    assert float_bits_equal(info, oinfo)
    assert (q == q_from_stream(ostream, S)).all()


# ---------------------------------------------------------------- K1 pool builder

@pytest.mark.parametrize("name,B,rgb", [("lena_grey", 8, False), ("lena_grey", 4, False), ("lena_grey", 16, False),
                                        ("lena64", 8, False), ("lena_colored", 8, True), ("lena_colored", 4, True)])
def test_pool_builder(handle, oracle, request, name, B, rgb):
    img = request.getfixturevalue(name)
    dec, s1, s2 = handle.build_pool(img, B, rgb)
    sc = oracle.scale_image(img, rgb=rgb).view(np.uint32)
    planes = [(sc >> 16) & 0xFF, (sc >> 8) & 0xFF, sc & 0xFF] if rgb else [(sc >> 16) & 0xFF]
    for c, p in enumerate(planes):
        assert (dec[c] == p).all()
    pool, mean, var = oracle.create_codebook(img, B, rgb=rgb)
    n = B * B
    if not rgb:
        assert (s1[0] == pool.sum(1)).all() and (s2[0] == (pool.astype(np.int64) ** 2).sum(1)).all()
        m = s1[0] // n
        assert (m == mean).all()
        assert ((s2[0] - 2 * m * s1[0] + n * m * m).astype(np.float32) == var).all()
    else:
        u = pool.view(np.uint32)
        for c, sh in enumerate((16, 8, 0)):
            ch = ((u >> sh) & 0xFF).astype(np.int64)
            assert (s1[c] == ch.sum(1)).all() and (s2[c] == (ch ** 2).sum(1)).all()
            m = s1[c] // n
            assert (m == mean[:, c + 1]).all()
            assert ((s2[c] - 2 * m * s1[c] + n * m * m).astype(np.float32) == var[:, c]).all()


# ---------------------------------------------------------------- encode, reference window sizes

@pytest.mark.parametrize("name,B,wk", [("lena_grey", 8, 2), ("lena_grey", 8, 4), ("lena_grey", 8, 16),
                                       ("lena_grey", 16, 16), ("lena_grey", 4, 16), ("lena_grey", 4, 2),
                                       ("lena64", 8, 2), ("lena64", 4, 8), ("lena64", 16, 2), ("lena64", 16, 5)])
def test_encode_grey_windowed(handle, oracle, request, name, B, wk):
    img = request.getfixturevalue(name)
    H, W = img.shape
    info, q = handle.encode(img, B, wk, rgb=False)
    oinfo = oracle.encode(img, B, wk, nthreads=4)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk), 3)


@pytest.mark.parametrize("B,wk", [(8, 2), (4, 4), (16, 2), (8, 7)])
def test_encode_rgb(handle, oracle, lena_colored, B, wk):
    info, q = handle.encode(lena_colored, B, wk, rgb=True)
    oinfo = oracle.encode(lena_colored, B, wk, rgb=True, nthreads=4)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, 256, 256, B, wk, rgb=True), 5)


def test_rgb_stream_equals_the_references_own_file(fic, handle, lena_colored):
    # unknown.run is the reference's own output for LenaColored.jpg at B=8, wk=2
    ref = open(os.path.join(GOLD, "unknown_run.bin"), "rb").read()
    _, q = handle.encode(lena_colored, 8, 2, rgb=True)
    assert fic.stream_write(q, 256, 256, 8, 2, rgb=True) == ref


def test_golden_streams(fic, handle, golden, lena_grey, lena64, lena_colored):
    imgs = {"lena_grey": lena_grey, "lena64": lena64, "lena_colored": lena_colored}
    for name, g in golden["oracle_streams"].items():
        img = imgs[name.rsplit("_b", 1)[0]]
        _, q = handle.encode(img, g["B"], g["wk"], rgb=g["rgb"])
        s = fic.stream_write(q, g["W"], g["H"], g["B"], g["wk"], rgb=g["rgb"])
        assert hashlib.sha256(s).hexdigest() == g["stream_sha256"], name


# ---------------------------------------------------------------- full pool: direct and tcgen05 engines

FULL = [("lena64", 8, 13), ("lena64", 4, 29), ("lena_grey", 8, 61), ("lena_grey", 4, 125), ("lena_grey", 16, 29)]


@pytest.mark.parametrize("name,B,wk", FULL)
def test_full_pool_direct(fic, handle, oracle, request, name, B, wk):
    img = request.getfixturevalue(name)
    H, W = img.shape
    handle.set_engine(fic.FIC_ENGINE_DIRECT)
    try:
        info, q = handle.encode(img, B, wk, rgb=False)
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
    oinfo = oracle.encode(img, B, wk, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk), 3)


KINDS = ["i8", "f16"]   # tensor-core instruction kind of the fused search; B = 16 always runs kind::i8


def _kind(fic, mma):
    return fic.FIC_UMMA_KIND_I8 if mma == "i8" else fic.FIC_UMMA_KIND_F16


@pytest.mark.parametrize("mma", KINDS)
@pytest.mark.parametrize("name,B,wk", FULL)
def test_full_pool_tcgen05(fic, handle, oracle, request, name, B, wk, mma):
    img = request.getfixturevalue(name)
    H, W = img.shape
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    handle.set_umma_kind(_kind(fic, mma))
    try:
        info, q = handle.encode(img, B, wk, rgb=False)
        assert handle.timings().engine == fic.FIC_ENGINE_UMMA
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
        handle.set_umma_kind(fic.FIC_UMMA_KIND_AUTO)
    oinfo = oracle.encode(img, B, wk, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk), 3)


@pytest.mark.parametrize("mma", KINDS)
@pytest.mark.parametrize("kind,B", [("noise", 8), ("structured", 8), ("structured", 4), ("sparse", 8), ("flat", 8),
                                    ("binary", 8), ("binary", 4), ("noise", 16), ("structured", 16), ("sparse", 16)])
def test_full_pool_tcgen05_synthetic(fic, handle, oracle, kind, B, mma):
    if B == 16 and mma == "f16":
        pytest.skip("B = 16 has no kind::f16 variant")
    W = H = 128
    if kind == "noise":
        p = fic.synth.noise(W, H, 3)
    elif kind == "binary":   # 0 / 255 in 4x4 cells: the largest |kov| the operands can produce
        p = np.kron((fic.synth.noise(W // 4, H // 4, 5) >> 7).astype(np.uint8) * 255, np.ones((4, 4), np.uint8))
    elif kind == "structured":
        p = fic.synth.structured(W, H, 3)
    elif kind == "flat":
        p = np.full((H, W), 77, np.uint8)
    else:  # flat background + sparse dots: ties and vR == 0 rows everywhere
        p = np.full((H, W), 100, np.uint8)
        p[fic.synth.noise(W, H, 9) < 3] = 103
    img = to_argb_grey(p)
    wk = 2 * W // B - 3
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    handle.set_umma_kind(_kind(fic, mma))
    try:
        info, q = handle.encode(img, B, wk, rgb=False)
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
        handle.set_umma_kind(fic.FIC_UMMA_KIND_AUTO)
    oinfo = oracle.encode(img, B, wk, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk), 3)


# ---------------------------------------------------------------- edge cases

def test_flat_image(handle, oracle):
    img = to_argb_grey(np.full((64, 64), 128, np.uint8))
    info, q = handle.encode(img, 8, 13, rgb=False)
    assert (q == 0).all()
    oinfo = oracle.encode(img, 8, 13)
    assert float_bits_equal(info, oinfo)


def test_non_square(fic, handle, oracle):
    # landscape exercises FC:993 (`x + 1 >= image.height`), portrait does not
    for W, H in [(96, 64), (64, 96), (128, 32)]:
        p = fic.synth.structured(W, H, 11)
        img = to_argb_grey(p)
        for B, wk in [(8, 2), (8, 5), (4, 3)]:
            info, q = handle.encode(img, B, wk, rgb=False)
            oinfo = oracle.encode(img, B, wk)
            assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk), 3)
    rgbp = np.stack([fic.synth.noise(96, 64, s) for s in (1, 2, 3)], -1)
    img = to_argb_rgb(rgbp)
    info, q = handle.encode(img, 8, 3, rgb=True)
    oinfo = oracle.encode(img, 8, 3, rgb=True)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, 96, 64, 8, 3, rgb=True), 5)


def test_range_slices_compose(fic, handle, lena_grey):
    full_info, full_q = handle.encode(lena_grey, 8, 61, rgb=False)
    info = np.zeros_like(full_info)
    q = np.zeros_like(full_q)
    for j0, j1 in [(0, 512), (512, 544), (544, 1024)]:   # row slices, as the multi-GPU host shards
        handle.encode(lena_grey, 8, 61, rgb=False, range_begin=j0, range_end=j1, info=info, q=q)
    assert float_bits_equal(info, full_info) and (q == full_q).all()


def test_rejected_arguments(fic, handle):
    for W, H, B, wk in [(64, 64, 2, 1), (60, 64, 8, 2), (64, 64, 8, 14), (64, 64, 8, 0), (8, 8, 8, 1), (64, 64, 32, 1)]:
        with pytest.raises(fic.FicError) as e:
            handle.encode(np.zeros((H, W), np.int32), B, wk, rgb=False)
        assert e.value.code == fic._lib.FIC_E_ARG


# ---------------------------------------------------------------- decoder / collage

@pytest.mark.parametrize("fname", ["lena_grey_b8_wk2.run", "lena64_b8_full.run", "lena64_b4_full.run", "unknown_run.bin"])
def test_decode_golden_streams(fic, handle, oracle, fname):
    s = open(os.path.join(GOLD, fname), "rb").read()
    rgb, W, H, B, wk, q = fic.stream_read(s)
    img, avg, it = handle.decode(q, W, H, B, wk, rgb)
    oimg, oavg, oit = oracle.decode(s)
    assert it == oit and np.float32(avg) == np.float32(oavg)
    assert (img == oimg).all()


def test_decode_avg_error_labels(fic, handle, golden, lena_grey):
    # the reference's own published numbers: Animation.gif "MSE" labels
    for B, wk, label in golden["gif_avg_error"]:
        _, q = handle.encode(lena_grey, B, wk, rgb=False)
        _, avg, _ = handle.decode(q, 256, 256, B, wk, False)
        assert np.float32(label) == np.float32(avg)


def test_decode_carry_and_iteration_cap(fic, handle, oracle):
    s = open(os.path.join(GOLD, "lena_grey_b8_wk2.run"), "rb").read()
    rgb, W, H, B, wk, q = fic.stream_read(s)
    # FractalCompression.avgError is static and never reset (FC:20): carry-in from a previous decode
    img, avg, it = handle.decode(q, W, H, B, wk, rgb, avg_error=0.5863342)
    oimg, oavg, oit = oracle.decode(s, avg_error_in=0.5863342)
    assert it == oit and np.float32(avg) == np.float32(oavg) and (img == oimg).all()
    # iteration cap: stop after 3 sweeps without convergence -> avgError is the unconverged float sum / (W*H)
    img3, avg3, it3 = handle.decode(q, W, H, B, wk, rgb, max_iters=3)
    assert it3 == 3 and avg3 >= 1


def test_collage(fic, handle, oracle, lena_grey, lena_colored):
    for img, B, wk, rgb in [(lena_grey, 8, 2, False), (lena_grey, 8, 61, False), (lena_colored, 8, 2, True)]:
        info, _ = handle.encode(img, B, wk, rgb=rgb)
        oinfo = info.copy()
        got = handle.collage(img, info, B, wk, rgb)
        want = oracle.collage(img, oinfo, B, wk, rgb=rgb)
        assert (got == want).all()


def test_facade_roundtrip(fic, oracle, lena_grey):
    FC = fic.FractalCompression
    FC.blockgroesse, FC.widthKernel, FC.avgError = 8, 4, np.float32(0)
    sink = fic.ByteSink()
    collage = FC.encode(fic.RasterImage.from_argb(lena_grey), sink)
    stream = sink.getvalue()
    oinfo = oracle.encode(lena_grey, 8, 4)
    assert stream == oracle.write_data(oinfo, 256, 256, 8, 4)
    import io

    dec = FC.decode(io.BytesIO(stream))
    oimg, oavg, _ = oracle.decode(stream)
    assert (dec.argb == oimg).all() and FC.getAvgError() == oavg
    assert (collage.argb == oracle.collage(lena_grey, oinfo.copy(), 8, 4)).all()


def test_cpp_host_cli_roundtrip(tmp_path, oracle, lena_grey):
    """The C++ mirror of the reference facade (host/fractal_compression.hpp) through its headless CLI:
    encode + decode of LenaGrey at the reference defaults must print the oracle's avgError."""
    import subprocess

    from conftest import ROOT

    cli = os.path.join(ROOT, "fractal-image-compression_b200", "lib", "fic_cli")
    if not os.path.exists(cli):
        pytest.skip("fic_cli not built")
    pgm = tmp_path / "lena.pgm"
    plane = ((lena_grey.view(np.uint32) >> 16) & 0xFF).astype(np.uint8)
    with open(pgm, "wb") as f:
        f.write(b"P5\n256 256\n255\n" + plane.tobytes())
    run = tmp_path / "lena.run"
    subprocess.run([cli, "encode", str(pgm), str(run), "8", "2"], check=True, capture_output=True, text=True)
    want = oracle.write_data(oracle.encode(lena_grey, 8, 2), 256, 256, 8, 2)
    assert open(run, "rb").read() == want
    out = tmp_path / "dec.pgm"
    r = subprocess.run([cli, "decode", str(run), str(out)], check=True, capture_output=True, text=True)
    _, oavg, oit = oracle.decode(want)
    assert f"({oit} iterations)" in r.stdout
    assert abs(float(r.stdout.split()[1]) - float(oavg)) < 1e-9
    dec = np.frombuffer(open(out, "rb").read().split(b"255\n", 1)[1], np.uint8).reshape(256, 256)
    oimg, _, _ = oracle.decode(want)
    assert (dec == ((oimg.view(np.uint32) >> 16) & 0xFF)).all()


def test_f16_tensor_path_is_exact_on_this_device(handle):
    """The library's own self-test (every binary32 accumulator of a kind::f16 search on worst-case content against
    integer arithmetic) must pass on a B200; a failing device would silently be served by kind::i8."""
    assert handle.f16_exact()


def test_tensor_peak_measurement(fic, handle):
    tops = handle.measure_int8_peak()
    # nominal dense int8 on B200 is 4500 TOP/s; anything far outside means the loop is not measuring the pipe
    assert 1000.0 < tops < 5500.0, tops
    tf = handle.measure_mma_peak(fic.FIC_UMMA_KIND_F16, 128)   # nominal dense f16: 2250
    assert 500.0 < tf < 2750.0, tf


def _random_plane(rng, W, H, kind):
    if kind == 0:
        return rng.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == 1:   # low contrast: many vR == 0 rows, flat domains, float ties
        return (120 + rng.integers(0, 3, (H, W))).astype(np.uint8)
    if kind == 2:   # binary: extreme contrast, |d - dmean| > 127 (the second s8 digit is live)
        return (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
    y, x = np.mgrid[0:H, 0:W]
    return ((x * 5 + y * 3 + rng.integers(0, 8, (H, W))) % 256).astype(np.uint8)


def test_randomised_parity_sweep(fic, handle, oracle):
    """Seeded sweep over image sizes, block sizes, windows, content kinds, grey and RGB: every code and
    quantised int must equal the oracle's; square full-pool cases run on both search engines."""
    rng = np.random.default_rng(20261018)
    cases = 0
    for _ in range(70):
        B = int(rng.choice([4, 8, 16]))
        rw, rh = int(rng.integers(2, 9)), int(rng.integers(2, 9))
        if rng.random() < 0.4:
            rh = rw
        W, H = rw * B, rh * B
        dpw, dph = 2 * rw - 3, 2 * rh - 3
        wk_max = min(dpw, dph)
        wk = wk_max if rng.random() < 0.4 else int(rng.integers(1, wk_max + 1))
        rgb = rng.random() < 0.3
        kind = int(rng.integers(0, 4))
        if rgb:
            img = to_argb_rgb(np.stack([_random_plane(rng, W, H, kind) for _ in range(3)], -1))
        else:
            img = to_argb_grey(_random_plane(rng, W, H, kind))
        oinfo = oracle.encode(img, B, wk, rgb=rgb)
        ostream = oracle.write_data(oinfo, W, H, B, wk, rgb=rgb)
        engines = [(fic.FIC_ENGINE_DIRECT, fic.FIC_UMMA_KIND_AUTO)]
        if not rgb and wk == dpw == dph:
            engines += [(fic.FIC_ENGINE_UMMA, fic.FIC_UMMA_KIND_I8), (fic.FIC_ENGINE_UMMA, fic.FIC_UMMA_KIND_F16)]
        if rgb and wk == dpw == dph and B != 16:   # RGB tensor path: kind::f16 only
            engines += [(fic.FIC_ENGINE_UMMA, fic.FIC_UMMA_KIND_AUTO)]
        for eng, mma in engines:
            handle.set_engine(eng)
            handle.set_umma_kind(mma)
            try:
                info, q = handle.encode(img, B, wk, rgb=rgb)
            finally:
                handle.set_engine(fic.FIC_ENGINE_AUTO)
                handle.set_umma_kind(fic.FIC_UMMA_KIND_AUTO)
            eng = (eng, mma)
            assert float_bits_equal(info, oinfo), (W, H, B, wk, rgb, kind, eng)
            assert (q == q_from_stream(ostream, 5 if rgb else 3)).all(), (W, H, B, wk, rgb, kind, eng)
        # decoder on the same stream
        _, Wd, Hd, Bd, wkd, qd = fic.stream_read(ostream)
        dimg, davg, dit = handle.decode(qd, Wd, Hd, Bd, wkd, rgb)
        oimg, oavg, oit = oracle.decode(ostream)
        assert (dimg == oimg).all() and dit == oit and np.float32(davg) == np.float32(oavg), (W, H, B, wk, rgb, kind)
        cases += 1
    assert cases == 70


def test_tcgen05_b16_extreme_digit(fic, handle, oracle):
    """B = 16: a lone 255 in a domain block of mean 0 gives d - dmean = +255 = 127 + 128, one more than two s8
    digits hold.  The tensor-core path stores such rows negated (-255 = -128 - 127 fits; only |kov| is used) and must
    stay on the tensor cores with the oracle's codes."""
    p = np.zeros((128, 128), np.uint8)
    p[40:42, 40:42] = 255          # one white 2x2 patch -> one decimated pixel of 255 among zeros
    p[100:110, 90:120] = 37        # some structure elsewhere so that ranges are not all flat
    img = to_argb_grey(p)
    wk = 2 * 128 // 16 - 3
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    try:
        info, q = handle.encode(img, 16, wk, rgb=False)
        assert handle.timings().engine == fic.FIC_ENGINE_UMMA
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
    oinfo = oracle.encode(img, 16, wk)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, 128, 128, 16, wk), 3)


def test_pinned_caller_buffers(fic, handle, lena_grey):
    """fic_pin_host_buffer: same results from page-locked caller buffers; double pin / stray unpin are errors."""
    info0, q0 = handle.encode(lena_grey, 8, 2, rgb=False)
    img = np.ascontiguousarray(lena_grey).copy()
    info = np.zeros_like(info0)
    q = np.zeros_like(q0)
    for a in (img, info, q):
        handle.pin(a)
    try:
        handle.encode(img, 8, 2, rgb=False, info=info, q=q)
        assert float_bits_equal(info, info0) and (q == q0).all()
        with pytest.raises(fic.FicError):
            handle.pin(img)              # already registered
    finally:
        for a in (img, info, q):
            handle.unpin(a)
    with pytest.raises(fic.FicError):
        handle.unpin(img)                # not registered any more
    info2, q2 = handle.encode(img, 8, 2, rgb=False)   # the handle is still healthy after the rejected calls
    assert (q2 == q0).all()


def test_planes_dev_rejects_misaligned_input(fic, handle):
    import torch

    W = H = 64
    buf = torch.zeros(W * H + 16, dtype=torch.uint8, device="cuda")
    info = torch.empty((64, 3), dtype=torch.float32, device="cuda")
    q = torch.empty((64, 3), dtype=torch.int32, device="cuda")
    with pytest.raises(fic.FicError) as e:
        handle.encode_planes_dev(buf.data_ptr() + 1, fic.FIC_MODE_GREY, W, H, 8, 2, 0, 64, info.data_ptr(), q.data_ptr())
    assert e.value.code == fic._lib.FIC_E_ARG
    handle.encode_planes_dev(buf.data_ptr(), fic.FIC_MODE_GREY, W, H, 8, 2, 0, 64, info.data_ptr(), q.data_ptr())
    handle.sync()
