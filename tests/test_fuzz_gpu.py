"""Randomised cross-check of the zip kernel's launch shapes and modes against the CPU oracle (tools/fuzz_zip.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_zip_kernel_against_oracle(seed):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_zip.py"), "60", str(seed)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "fuzz ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
