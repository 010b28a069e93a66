"""CTA pairs of the tcgen05 search (tcgen05 cta_group::2, FIC_OPT_UMMA_PAIR): the pair kernel and the single-CTA
kernel must both reproduce the oracle's codes bit for bit (FC:613-644, FC:655-687; RGB FC:697-808), on every
configuration either of them serves, whatever the option says.  AUTO runs pairs at blockgroesse 8 and 16."""
import os

import numpy as np
import pytest

from conftest import to_argb_grey, to_argb_rgb
from test_gpu_parity import assert_codes_equal, float_bits_equal

pytestmark = pytest.mark.gpu


def _content(kind, W, seed):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (W, W), dtype=np.uint8)
    if kind == "binary":  # 0 / 255 cells: the largest covariances the operands can produce
        return (rng.integers(0, 2, (W // 4, W // 4), dtype=np.uint8) * 255).repeat(4, 0).repeat(4, 1)
    if kind == "flat":
        return np.full((W, W), 77, np.uint8)
    if kind == "periodic":  # many exactly equal domains: ties resolved by the lowest index (FC:627)
        y, x = np.mgrid[0:W, 0:W]
        return ((x % 16) * 13 + (y % 16) * 5).astype(np.uint8)
    y, x = np.mgrid[0:W, 0:W]
    return ((x * 3 + y * 2 + rng.integers(0, 8, (W, W))) & 255).astype(np.uint8)


def _encode(fic, handle, img, B, wk, pair, rgb=False, **kw):
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    handle.set_umma_kind(fic.FIC_UMMA_KIND_F16)
    handle.set_umma_pair(pair)
    try:
        out = handle.encode(img, B, wk, rgb=rgb, **kw)
        used = handle.umma_pair_used()
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
        handle.set_umma_kind(fic.FIC_UMMA_KIND_AUTO)
        handle.set_umma_pair(fic.FIC_UMMA_PAIR_AUTO)
    return out, used


@pytest.mark.parametrize("pair", ["on", "off"])
@pytest.mark.parametrize("kind,W,B", [("noise", 256, 8), ("structured", 256, 8), ("binary", 256, 8), ("flat", 128, 8), ("periodic", 256, 8),
                                      ("structured", 384, 8), ("structured", 128, 4), ("binary", 128, 4), ("noise", 192, 4),
                                      ("structured", 512, 16), ("binary", 256, 16), ("noise", 384, 16), ("periodic", 256, 16)])
def test_pair_and_single_equal_oracle_grey(fic, handle, oracle, kind, W, B, pair):
    """Full-pool grey encodes through either kernel against the oracle (384^2: an odd number of 512-row super-blocks,
    the pair's padded partner; 128^2: a single, mostly padded pair; blockgroesse 16: the K-split kind::i8 kernel, whose
    pair form loads half of either part of a tile per CTA -- binary content has low digits, i.e. second parts)."""
    img = to_argb_grey(_content(kind, W, 5))
    wk = 2 * (W // B) - 3
    (info, q), used = _encode(fic, handle, img, B, wk, fic.FIC_UMMA_PAIR_ON if pair == "on" else fic.FIC_UMMA_PAIR_OFF)
    assert used == (pair == "on")
    oinfo = oracle.encode(img, B, wk, nthreads=os.cpu_count() or 1)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, W, B, wk), 3)


@pytest.mark.parametrize("pair", ["on", "off"])
@pytest.mark.parametrize("kind,B", [("structured", 8), ("binary", 8), ("noise", 4), ("structured", 16), ("binary", 16)])
def test_pair_and_single_equal_oracle_rgb(fic, handle, oracle, kind, B, pair):
    W = 128 if B == 4 else (384 if B == 16 else 256)
    rgb = np.stack([_content(kind, W, s) for s in (1, 2, 3)], axis=-1)
    img = to_argb_rgb(rgb)
    wk = 2 * (W // B) - 3
    (info, q), used = _encode(fic, handle, img, B, wk, fic.FIC_UMMA_PAIR_ON if pair == "on" else fic.FIC_UMMA_PAIR_OFF, rgb=True)
    assert used == (pair == "on")
    oinfo = oracle.encode(img, B, wk, rgb=True, nthreads=os.cpu_count() or 1)
    assert_codes_equal(info, q, oinfo, oracle.write_data(oinfo, W, W, B, wk, rgb=True), 5)


def test_pair_auto_policy(fic, handle, lena_grey):
    """AUTO: pairs at blockgroesse 8 and 16 (the K-split kind::i8 kernel), the single-CTA kernel at 4 (epilogue
    bound); an explicit ON is ignored where no pair kernel exists (kind::i8 at blockgroesse 8)."""
    handle.set_engine(fic.FIC_ENGINE_UMMA)
    try:
        handle.encode(lena_grey, 8, 61, rgb=False)
        assert handle.umma_pair_used()
        handle.encode(lena_grey, 4, 125, rgb=False)
        assert not handle.umma_pair_used()
        handle.encode(lena_grey, 16, 29, rgb=False)
        assert handle.umma_pair_used()
        handle.set_umma_pair(fic.FIC_UMMA_PAIR_OFF)
        handle.encode(lena_grey, 16, 29, rgb=False)
        assert not handle.umma_pair_used()
        handle.set_umma_pair(fic.FIC_UMMA_PAIR_ON)
        handle.set_umma_kind(fic.FIC_UMMA_KIND_I8)
        handle.encode(lena_grey, 8, 61, rgb=False)
        assert not handle.umma_pair_used()
    finally:
        handle.set_engine(fic.FIC_ENGINE_AUTO)
        handle.set_umma_kind(fic.FIC_UMMA_KIND_AUTO)
        handle.set_umma_pair(fic.FIC_UMMA_PAIR_AUTO)
    handle.encode(lena_grey, 8, 2, rgb=False)  # a windowed encode is not a tcgen05 search
    assert not handle.umma_pair_used()


@pytest.mark.parametrize("W,B", [(1024, 8), (2048, 16)])
def test_pair_range_slices_compose(fic, handle, W, B):
    """Row shards (what every rank of a multi-GPU encode runs) through the pair kernel: slices whose row counts are
    not multiples of the pair's 1024 rows compose to the whole-image result of the single-CTA kernel (B = 16: the
    K-split pair kernel)."""
    img = to_argb_grey(_content("structured", W, 9))
    rpw = W // B
    wk = 2 * rpw - 3
    NR = rpw * rpw
    (info0, q0), used0 = _encode(fic, handle, img, B, wk, fic.FIC_UMMA_PAIR_OFF)
    assert not used0
    info = np.zeros_like(info0)
    q = np.zeros_like(q0)
    for j0, j1 in [(0, 5 * rpw), (5 * rpw, 5 * rpw + 700), (5 * rpw + 700, NR - 3), (NR - 3, NR)]:
        _, used = _encode(fic, handle, img, B, wk, fic.FIC_UMMA_PAIR_ON, range_begin=j0, range_end=j1, info=info, q=q)
        assert used
    assert (q == q0).all() and float_bits_equal(info, info0)


@pytest.mark.parametrize("B", [8, 16])
def test_pair_equals_single_at_2048(fic, handle, B):
    """The two kernels on a 2048^2 image (65 536 rows, 1.1e9 evaluations per super-block pair at B = 8): identical codes."""
    W = 2048
    img = fic.synth.grey_to_argb(fic.synth.structured(W, W, 4))
    wk = 2 * (W // B) - 3
    (info_p, q_p), used_p = _encode(fic, handle, img, B, wk, fic.FIC_UMMA_PAIR_ON)
    (info_s, q_s), used_s = _encode(fic, handle, img, B, wk, fic.FIC_UMMA_PAIR_OFF)
    assert used_p and not used_s
    assert (q_p == q_s).all() and float_bits_equal(info_p, info_s)
