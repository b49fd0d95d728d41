"""Short randomised parity sweep on the GPU (scripts/gpu_fuzz.py): sorts and searches vs the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz.py"), str(seed), "12"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout
