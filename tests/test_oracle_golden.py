"""Pins the CPU oracle (oracle/fic_oracle.c) to the only known answers the reference ships:
its own encoded stream unknown.run and the avgError labels of Animation.gif, then freezes
the oracle's output on the parity configs by digest (tests/golden/golden.json)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLD, to_argb_grey


def test_unknown_run_byte_exact(oracle, lena_colored):
    # the reference's own RGB encode of LenaColored.jpg at its defaults (FC:14-15): B=8, wk=2
    ref = open(os.path.join(GOLD, "unknown_run.bin"), "rb").read()
    info = oracle.encode(lena_colored, 8, 2, rgb=True)
    assert oracle.write_data(info, 256, 256, 8, 2, rgb=True) == ref


def test_unknown_run_decodes(oracle):
    ref = open(os.path.join(GOLD, "unknown_run.bin"), "rb").read()
    img, avg, it = oracle.decode(ref)
    assert it == 13 and float(avg) == pytest.approx(0.7711792, abs=1e-7)


@pytest.mark.parametrize("B,wk,label", [(16, 16, "0.3744049"), (8, 16, "0.3647766"), (4, 16, "0.73760986"),
                                        (8, 8, "0.52404785"), (8, 4, "0.36376953")])
def test_gif_avg_error_labels(oracle, lena_grey, B, wk, label):
    # Animation.gif shows "MSE " + Float.toString(avgError) (RLEAppController.java:180)
    info = oracle.encode(lena_grey, B, wk)
    img, avg, it = oracle.decode(oracle.write_data(info, 256, 256, B, wk))
    # Java's Float.toString prints the shortest decimal that round-trips the float
    assert np.float32(label) == avg, (label, repr(avg))


def test_golden_digests(oracle, golden, lena_grey, lena64, lena_colored):
    imgs = {"lena_grey": lena_grey, "lena64": lena64, "lena_colored": lena_colored}
    for name, g in golden["oracle_streams"].items():
        img = imgs[name.rsplit("_b", 1)[0]]
        info = oracle.encode(img, g["B"], g["wk"], rgb=g["rgb"], nthreads=4)
        s = oracle.write_data(info, g["W"], g["H"], g["B"], g["wk"], rgb=g["rgb"])
        assert hashlib.sha256(s).hexdigest() == g["stream_sha256"], name
        dec, avg, it = oracle.decode(s)
        assert hashlib.sha256(dec.tobytes()).hexdigest() == g["decoded_sha256"], name
        assert float(avg).hex() == g["avg_error_hex"] and it == g["iterations"], name


def test_threads_do_not_change_results(oracle, lena_grey):
    a = oracle.encode(lena_grey, 8, 4, nthreads=1)
    b = oracle.encode(lena_grey, 8, 4, nthreads=7)
    assert a.tobytes() == b.tobytes()


def test_range_slices_compose(oracle, lena64):
    full = oracle.encode(lena64, 8, 13)
    part = np.zeros_like(full)
    for j0, j1 in [(0, 24), (24, 25), (25, 64)]:
        p = oracle.encode(lena64, 8, 13, range_begin=j0, range_end=j1)
        part[j0:j1] = p[j0:j1]
    assert part.tobytes() == full.tobytes()


def test_flat_image_codes_are_zero(oracle):
    # 0/0 -> NaN -> (int) -> 0 on flat winners (FC:634, FC:243-244): every code is (0,0,0)
    img = to_argb_grey(np.full((64, 64), 128, np.uint8))
    info = oracle.encode(img, 8, 13)
    s = oracle.write_data(info, 64, 64, 8, 13)
    assert set(np.frombuffer(s[20:], ">i4").tolist()) == {0}
    dec, avg, it = oracle.decode(s)
    assert ((dec.view(np.uint32) >> 16) & 0xFF).max() == 0  # decodes to black


def test_geometry_known_answers(oracle):
    # FC:516-545 / FC:84-100 on a 256x256, B=8 image: rpw=32, dpw=61
    assert oracle.domain_block_index(0, 0, 32, 32, 61, 8) == 1 + 1 * 61
    assert oracle.domain_block_index(16, 8, 32, 32, 61, 8) == 2 + 1 * 61
    assert oracle.domain_block_index(248, 248, 32, 32, 61, 8) == (30 * 2 - 2) + (2 * 30 - 1) * 61
    assert oracle.generate_kernel(61, 61, 62, 2) == (0, 0)
    assert oracle.generate_kernel(61, 61, 60 + 60 * 61, 2) == (59, 59)
    assert oracle.generate_kernel(61, 61, 30 + 30 * 61, 61) == (0, 0)   # full pool: origin forced to 0


@pytest.mark.parametrize("args", [(64, 64, 2, 1), (60, 64, 8, 2), (64, 64, 8, 14), (64, 64, 8, 0), (8, 8, 8, 1)])
def test_arguments_the_reference_throws_on(oracle, args):
    W, H, B, wk = args
    with pytest.raises(ValueError):
        oracle.encode(np.zeros((H, W), np.int32), B, wk)


def test_scale_image_quirks(oracle):
    rng = np.random.default_rng(3)
    p = rng.integers(0, 256, (8, 16, 3), dtype=np.uint8)   # landscape: FC:993 / FC:940 fire for x+1 >= H
    a = p.astype(np.uint32)
    argb = (0xFF000000 | (a[..., 0] << 16) | (a[..., 1] << 8) | a[..., 2]).view(np.int32)
    g = (oracle.scale_image(argb).view(np.uint32) >> 16) & 0xFF
    r = p[..., 0].astype(int)
    for y in range(4):
        for x in range(8):
            fourth = 128 if 2 * x + 1 >= 8 else r[2 * y + 1, 2 * x + 1]
            assert g[y, x] == (r[2 * y, 2 * x] + r[2 * y, 2 * x + 1] + r[2 * y + 1, 2 * x] + fourth) // 4
    c = oracle.scale_image(argb, rgb=True).view(np.uint32)
    for y in range(4):
        for x in range(8):
            for ch, sh in enumerate((16, 8, 0)):
                q = p[..., ch].astype(int)
                fourth = 128 if 2 * x + 1 >= 8 else q[2 * y + 1, 2 * x]   # FC:945 re-reads (x, y+1)
                want = (q[2 * y, 2 * x] + q[2 * y, 2 * x + 1] + q[2 * y + 1, 2 * x] + fourth) // 4
                assert (c[y, x] >> sh) & 0xFF == want
