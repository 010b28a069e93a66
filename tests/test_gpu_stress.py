"""Randomised cross-check of the CUDA-core and tensor-core search engines (tools/gpu_stress.py) on sizes and
content the oracle is not run on: both instruction kinds, the isometry extension, periodic / binary / mean-0
content.  The direct engine is itself oracle-exact on every small case of test_gpu_parity.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_engines_agree_on_random_cases():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_stress.py"), "120", "3"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "STRESS PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
