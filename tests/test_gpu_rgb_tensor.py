"""RGB full-pool search on the tensor cores (tcgen05 kind::f16) against the oracle.

The reference scores RGB candidates with one covariance for the three channels, accumulated sequentially in
binary32 (FC:760-808).  For B <= 8 that sum is provably an exact integer for every pair of blocks ("RGB operands" in
csrc/fic_search_umma.cu), which is what lets the tensor cores compute it; the extreme-contrast cases below sit at
the top of the magnitude range (|kov| up to 9.3e6 < 2^24).  Every case must equal the oracle bit for bit.
"""
import numpy as np
import pytest

from conftest import to_argb_rgb
from test_gpu_parity import assert_codes_equal, float_bits_equal, q_from_stream

pytestmark = pytest.mark.gpu


def _encode_umma(fic, handle, img, B, wk, **kw):
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    try:
        info, q = handle.encode(img, B, wk, rgb=True, **kw)
        assert handle.timings().engine == fic.FIC_ENGINE_UMMA
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
    return info, q


def _planes(fic, kind, W, H):
    if kind == "noise":
        return np.stack([fic.synth.noise(W, H, s) for s in (1, 2, 3)], -1)
    if kind == "structured":
        return np.stack([fic.synth.structured(W, H, s) for s in (4, 5, 6)], -1)
    if kind == "grey":      # r = g = b: every gR, gD is three times the grey value
        p = fic.synth.structured(W, H, 7)
        return np.stack([p, p, p], -1)
    if kind == "binary":    # 0 / 255 in all channels at once: the largest |gR|, |gD| and |kov| (9.3e6) the path can meet
        p = np.kron((fic.synth.noise(W // 4, H // 4, 5) >> 7).astype(np.uint8) * 255, np.ones((4, 4), np.uint8))
        return np.stack([p, p, p], -1)
    if kind == "mixed":     # left half natural, right half extreme contrast
        a = np.stack([fic.synth.structured(W, H, s) for s in (8, 9, 10)], -1)
        b = np.kron((fic.synth.noise(W // 2, H // 2, 11) >> 7).astype(np.uint8) * 255, np.ones((2, 2), np.uint8))
        a[:, W // 2:, :] = b[:, W // 2:, None]
        return a
    if kind == "flat":
        return np.stack([np.full((H, W), v, np.uint8) for v in (10, 200, 77)], -1)
    if kind == "sparse":    # flat background + sparse dots: vR == 0 rows, vD == 0 domains, ties
        p = np.full((H, W, 3), 100, np.uint8)
        for c in range(3):
            p[..., c][fic.synth.noise(W, H, 20 + c) < 3] = 103 + c
        return p
    raise ValueError(kind)


@pytest.mark.parametrize("B", [8, 4])
@pytest.mark.parametrize("kind", ["noise", "structured", "grey", "binary", "mixed", "flat", "sparse"])
def test_rgb_full_pool_tcgen05_synthetic(fic, handle, oracle, kind, B):
    W = H = 128 if B == 8 else 64
    img = to_argb_rgb(_planes(fic, kind, W, H))
    wk = 2 * W // B - 3
    info, q = _encode_umma(fic, handle, img, B, wk)
    oinfo = oracle.encode(img, B, wk, rgb=True, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk, rgb=True), 5)


def test_rgb_full_pool_tcgen05_lena(fic, handle, oracle, lena_colored):
    info, q = _encode_umma(fic, handle, lena_colored, 8, 61)
    oinfo = oracle.encode(lena_colored, 8, 61, rgb=True, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, 256, 256, 8, 61, rgb=True), 5)


def test_rgb_tensor_equals_cuda_core_kernel_at_512(fic, handle):
    """A size the oracle does not finish in seconds: the two engines must agree with each other."""
    W = H = 512
    img = to_argb_rgb(np.stack([fic.synth.structured(W, H, s) for s in (1, 2, 3)], -1))
    wk = 2 * W // 8 - 3
    info, q = _encode_umma(fic, handle, img, 8, wk)
    handle.set_engine(fic.FIC_ENGINE_DIRECT)
    try:
        dinfo, dq = handle.encode(img, 8, wk, rgb=True)
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
    assert float_bits_equal(info, dinfo) and (q == dq).all()


def test_rgb_tensor_range_slices_compose(fic, handle, lena_colored):
    full_info, full_q = _encode_umma(fic, handle, lena_colored, 8, 61)
    info = np.zeros_like(full_info)
    q = np.zeros_like(full_q)
    for j0, j1 in [(0, 100), (100, 101), (101, 700), (700, 1024)]:
        _encode_umma(fic, handle, lena_colored, 8, 61, range_begin=j0, range_end=j1, info=info, q=q)
    assert float_bits_equal(info, full_info) and (q == full_q).all()


def test_rgb_b16_has_no_tensor_path(fic, handle, lena_colored):
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    try:
        with pytest.raises(fic.FicError):
            handle.encode(lena_colored, 16, 29, rgb=True)
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)


def test_rgb_random_cases(fic, handle, oracle):
    rng = np.random.default_rng(77)
    for case in range(24):
        B = int(rng.choice([4, 8]))
        r = int(rng.integers(2, 12))
        W = H = r * B
        wk = 2 * r - 3
        kind = int(rng.integers(0, 4))
        if kind == 0:
            p = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        elif kind == 1:   # low contrast
            p = (120 + rng.integers(0, 3, (H, W, 3))).astype(np.uint8)
        elif kind == 2:   # extreme contrast, channels independent
            p = (rng.integers(0, 2, (H, W, 3)) * 255).astype(np.uint8)
        else:             # extreme contrast, channels equal, in 2x2 cells
            c = (rng.integers(0, 2, (H // 2, W // 2)) * 255).astype(np.uint8)
            p = np.repeat(np.kron(c, np.ones((2, 2), np.uint8))[..., None], 3, -1)
        img = to_argb_rgb(p)
        info, q = _encode_umma(fic, handle, img, B, wk)
        oinfo = oracle.encode(img, B, wk, rgb=True)
        ostream = oracle.write_data(oinfo, W, H, B, wk, rgb=True)
        assert float_bits_equal(info, oinfo), (case, W, B, kind)
        assert (q == q_from_stream(ostream, 5)).all(), (case, W, B, kind)
