"""Full-size checks through size-independent properties (the oracle cannot finish these
sizes): both search engines must agree on every code, and encode -> decode must
reconstruct the image to the PSNR fractal coding reaches on this content."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


@pytest.mark.parametrize("W,B,kind", [(512, 8, "structured"), (512, 4, "structured"), (1024, 8, "noise"),
                                      (1024, 8, "structured"), (1024, 16, "structured"), (1024, 16, "noise")])
def test_engines_agree_full_pool(fic, handle, W, B, kind):
    p = getattr(fic.synth, kind)(W, W, 7)
    img = fic.synth.grey_to_argb(p)
    wk = 2 * W // B - 3
    handle.set_engine(fic.FIC_ENGINE_DIRECT)
    i1, q1 = handle.encode(img, B, wk, rgb=False)
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    i2, q2 = handle.encode(img, B, wk, rgb=False)
    handle.set_engine(fic.FIC_ENGINE_AUTO)
    from test_gpu_parity import float_bits_equal

    assert (q1 == q2).all() and float_bits_equal(i1, i2)


def test_roundtrip_2048(fic, handle):
    W = 2048
    p = fic.synth.structured(W, W, 1)
    img = fic.synth.grey_to_argb(p)
    wk = 2 * W // 8 - 3
    _, q = handle.encode(img, 8, wk, rgb=False)
    t = handle.timings()
    assert t.engine == fic.FIC_ENGINE_UMMA
    dec, avg, it = handle.decode(q, W, W, 8, wk, False)
    rec = ((dec.view(np.uint32) >> 16) & 0xFF).astype(np.uint8)
    assert it < 50 and avg < 1
    assert psnr(p, rec) > 25.0
    # sharded by range rows == unsharded (what the multi-GPU host relies on)
    info2 = np.zeros((q.shape[0], 3), np.float32)
    q2 = np.zeros_like(q)
    for j0, j1 in [(0, 256 * 100), (256 * 100, 256 * 256)]:
        handle.encode(img, 8, wk, rgb=False, range_begin=j0, range_end=j1, info=info2, q=q2)
    assert (q2 == q).all()


@pytest.mark.parametrize("W,B,period", [(1024, 8, 16), (512, 4, 8)])
def test_periodic_image_ties_and_flag_overflow(fic, handle, W, B, period):
    """A tiled texture makes thousands of domains identical: every copy of the best domain ties, the
    reference keeps the lowest index, and the tcgen05 path's per-row flag lists overflow (the refine
    step then rescans those rows in full).  Both engines must still agree on every code."""
    tile = fic.synth.noise(period, period, 5)
    p = np.tile(tile, (W // period, W // period))
    p = p.copy()
    p[::64, ::64] ^= 1  # a few irregularities so that not every range block is the same
    img = fic.synth.grey_to_argb(p)
    wk = 2 * W // B - 3
    handle.set_engine(fic.FIC_ENGINE_DIRECT)
    i1, q1 = handle.encode(img, B, wk, rgb=False)
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    i2, q2 = handle.encode(img, B, wk, rgb=False)
    handle.set_engine(fic.FIC_ENGINE_AUTO)
    from test_gpu_parity import float_bits_equal

    assert (q1 == q2).all() and float_bits_equal(i1, i2)
