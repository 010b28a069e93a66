import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def to_argb_grey(plane):
    v = plane.astype(np.uint32)
    return (np.uint32(0xFF000000) | (v << 16) | (v << 8) | v).view(np.int32)


def to_argb_rgb(rgb):
    a = rgb.astype(np.uint32)
    return (np.uint32(0xFF000000) | (a[..., 0] << 16) | (a[..., 1] << 8) | a[..., 2]).view(np.int32)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def lena_grey():
    return to_argb_grey(np.fromfile(os.path.join(GOLD, "lena_grey_256.u8"), np.uint8).reshape(256, 256))


@pytest.fixture(scope="session")
def lena64():
    return to_argb_grey(np.fromfile(os.path.join(GOLD, "lena64.u8"), np.uint8).reshape(64, 64))


@pytest.fixture(scope="session")
def lena_colored():
    return to_argb_rgb(np.fromfile(os.path.join(GOLD, "lena_colored_256.rgb"), np.uint8).reshape(256, 256, 3))


@pytest.fixture(scope="session")
def golden():
    import json

    with open(os.path.join(GOLD, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fic():
    import fractal_image_compression_b200 as f

    return f


@pytest.fixture(scope="session")
def handle(fic):
    h = fic.Handle(0)
    yield h
    h.close()
