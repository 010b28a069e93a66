"""The C-ABI library loads and exports every symbol include/fic_b200.h declares; host-only
entries (geometry, .run stream helpers) behave like the reference's; without a GPU every
compute entry fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fic_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fic_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(fic):
    L = fic._lib.load()
    syms = header_symbols()
    assert sorted(fic.ABI_SYMBOLS) == syms
    for s in syms:
        assert getattr(L, s) is not None


def test_no_oracle_in_product():
    # the product must never import, link or call the oracle
    pkg = os.path.join(ROOT, "fractal-image-compression_b200")
    for dp, _, fs in os.walk(pkg):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                assert "oracle" not in open(os.path.join(dp, f)).read().lower().replace("# oracle-free", ""), f


def test_geometry(fic):
    L = fic._lib.load()
    nr, nd = C.c_int64(), C.c_int64()
    assert L.fic_geometry(256, 256, 8, 2, C.byref(nr), C.byref(nd)) == 0
    assert (nr.value, nd.value) == (1024, 3721)
    assert L.fic_geometry(4096, 4096, 8, 1021, C.byref(nr), C.byref(nd)) == 0
    assert (nr.value, nd.value) == (262144, 1042441)
    assert L.fic_geometry(8192, 8192, 8, 2045, C.byref(nr), C.byref(nd)) == 0
    assert (nr.value, nd.value) == (1048576, 4182025)
    # argument sets the reference throws on (FC:1019, FC:124-126, FC:93-96)
    for bad in [(64, 64, 2, 1), (60, 64, 8, 2), (64, 64, 8, 14), (64, 64, 8, 0), (8, 8, 8, 1), (64, 64, 32, 1)]:
        assert L.fic_geometry(*bad, None, None) == fic._lib.FIC_E_ARG
    # pool of 2^24 or more entries does not fit the reference's float index
    assert L.fic_geometry(8192, 8192, 4, 2, None, None) == 0
    assert L.fic_geometry(16384, 8192, 4, 2, None, None) == fic._lib.FIC_E_ARG


def test_stream_roundtrip_matches_reference_format(fic, oracle, lena64):
    info = oracle.encode(lena64, 8, 2)
    want = oracle.write_data(info, 64, 64, 8, 2)
    q = np.frombuffer(want[20:], ">i4").astype(np.int32).reshape(-1, 3)
    got = fic.stream_write(q, 64, 64, 8, 2, rgb=False)
    assert got == want
    rgb, W, H, B, wk, q2 = fic.stream_read(got)
    assert (rgb, W, H, B, wk) == (False, 64, 64, 8, 2) and (q2 == q).all()


def test_stream_reads_the_reference_stream(fic):
    ref = open(os.path.join(ROOT, "tests", "golden", "unknown_run.bin"), "rb").read()
    rgb, W, H, B, wk, q = fic.stream_read(ref)
    assert (rgb, W, H, B, wk) == (True, 256, 256, 8, 2) and q.shape == (1024, 5)
    assert fic.stream_write(q, W, H, B, wk, rgb=True) == ref
    with pytest.raises(fic.FicError):
        fic.stream_read(ref[:100])


def test_no_cpu_fallback(fic):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(fic.FicError) as e:
        fic.Handle(0)
    assert e.value.code == fic._lib.FIC_E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(fic.FicError) as e:   # the multi-GPU handle is built from the same per-device contexts
        fic.MultiHandle([0])
    assert e.value.code == fic._lib.FIC_E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(fic.FicError) as e:
        fic.MultiHandle([0, 0])
    assert e.value.code == fic._lib.FIC_E_ARG
