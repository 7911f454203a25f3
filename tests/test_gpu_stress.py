"""A slice of the randomised differential test (tools/stress.py) inside the GPU suite."""
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 2026])
def test_random_differential_vs_oracle(seed):
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "stress.py"), "--cases", "120", "--seed", str(seed),
                        "--max-n", "200000"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout
