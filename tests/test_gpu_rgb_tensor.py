"""RGB full-pool search on the tensor cores (tcgen05 kind::f16) against the oracle.

The reference scores RGB candidates with one covariance for the three channels, accumulated sequentially in
binary32 (FC:760-808).  For B <= 8 that sum is provably an exact integer for every pair of blocks ("RGB operands" in
csrc/fic_search_umma.cu), which is what lets the tensor cores compute it; the extreme-contrast cases below sit at
the top of the magnitude range (|kov| up to 9.3e6 < 2^24).  At B = 16 the sum can round (see the B = 16 section):
the tensor cores filter with widened bounds and the refine step replays the float sum.  Every case must equal the
oracle bit for bit.
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


# ---------------------------------------------------------------- blockgroesse 16 (FC:760-808 at RLEAppView.fxml:55's largest block)
#
# At B = 16 the reference's sequential binary32 covariance is NOT always an exact integer: partial sums reach
# 256 * 382^2 = 3.7e7 > 2^24 on extreme-contrast content, and what the reference ranks by is the ROUNDED sum.  The
# tensor-core pass is then a filter whose chunk bounds are widened by the worst-case rounding of both sums
# (||gR|| ||gD|| * 6e-5 where that product exceeds 2^24), and the refine step walks the reference's float sum
# literally.  "binary" / "binary_px" are the contents where the sums do round; all must equal the oracle bit for bit.

def _planes16(fic, kind, W, H):
    if kind == "binary_px":   # 0 / 255 per pixel and channel, independent
        return np.stack([(fic.synth.noise(W, H, s) >> 7).astype(np.uint8) * 255 for s in (31, 32, 33)], -1)
    if kind == "binary_rows":  # 0 / 255 in 16 x 1 stripes, channels equal: domain blocks reach the largest ||gD||
        p = np.repeat((fic.synth.noise(W // 16, H, 41) >> 7).astype(np.uint8) * 255, 16, axis=1)
        return np.stack([p, p, p], -1)
    return _planes(fic, kind, W, H)


@pytest.mark.parametrize("kind", ["noise", "structured", "grey", "binary", "binary_px", "binary_rows", "mixed", "flat", "sparse"])
def test_rgb_b16_full_pool_tcgen05_synthetic(fic, handle, oracle, kind):
    W = H = 256
    B = 16
    img = to_argb_rgb(_planes16(fic, kind, W, H))
    wk = 2 * W // B - 3
    info, q = _encode_umma(fic, handle, img, B, wk)
    oinfo = oracle.encode(img, B, wk, rgb=True, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, H, B, wk, rgb=True), 5)


def test_rgb_b16_auto_engine_and_lena(fic, handle, oracle, lena_colored):
    """AUTO picks the tensor cores for RGB at B = 16 once the pool is large enough; Lena and a 640^2 natural image."""
    info, q = _encode_umma(fic, handle, lena_colored, 16, 29)
    oinfo = oracle.encode(lena_colored, 16, 29, rgb=True, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, 256, 256, 16, 29, rgb=True), 5)
    W = 640   # 1600 ranges x 5929 domains: above the tensor path's work threshold (2^22 evaluations)
    img = to_argb_rgb(np.stack([fic.synth.structured(W, W, s) for s in (1, 2, 3)], -1))
    wk = 2 * W // 16 - 3
    info, q = handle.encode(img, 16, wk, rgb=True)
    assert handle.timings().engine == fic.FIC_ENGINE_UMMA
    oinfo = oracle.encode(img, 16, wk, rgb=True, nthreads=8)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, W, 16, wk, rgb=True), 5)


@pytest.mark.parametrize("kind", ["structured", "binary_px"])
def test_rgb_b16_tensor_equals_cuda_core_kernel_at_1024(fic, handle, kind):
    """1024^2 (4096 ranges x 15 625 domains x 256 pixels): the tensor-core path against the CUDA-core kernel, which
    walks the reference's float sum literally; row slices compose."""
    W = H = 1024
    img = to_argb_rgb(_planes16(fic, kind, W, H))
    wk = 2 * W // 16 - 3
    info, q = _encode_umma(fic, handle, img, 16, wk)
    handle.set_engine(fic.FIC_ENGINE_DIRECT)
    try:
        dinfo, dq = handle.encode(img, 16, wk, rgb=True)
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
    assert float_bits_equal(info, dinfo) and (q == dq).all()
    info2 = np.zeros_like(info)
    q2 = np.zeros_like(q)
    for j0, j1 in [(0, 1000), (1000, 1001), (1001, 4096)]:
        _encode_umma(fic, handle, img, 16, wk, range_begin=j0, range_end=j1, info=info2, q=q2)
    assert float_bits_equal(info2, info) and (q2 == q).all()


def test_rgb_random_cases(fic, handle, oracle):
    rng = np.random.default_rng(77)
    for case in range(24):
        B = int(rng.choice([4, 8, 16]))
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
