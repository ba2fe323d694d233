"""Randomised differential run of the CUDA path against the plain-C oracle (tools/stress_random.py): random 3-CNFs
and multiplier CNFs, random -d / -q, random BFS window lengths (1..16 levels per grid barrier) and random piece-split
targets -- frontier, counters, per-subproblem verdict / node / clause-evaluation counts, winner and model."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [101, 202])
def test_randomised_differential(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_random.py"), "200", str(seed)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout
