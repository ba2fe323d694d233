"""Randomised differential run of the CUDA path against the plain-C oracle (tools/stress_random.py): random 3-CNFs
and multiplier CNFs, random -d / -q, random BFS window lengths (1..16 levels per grid barrier) and random piece-split
targets -- frontier, counters, per-subproblem verdict / node / clause-evaluation counts, winner and model."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed,env", [(101, {}), (202, {}), (303, {"NDP_NO_SUMMARY": "1"}), (404, {"NDP_DFS_SNAPS": "0"})])
def test_randomised_differential(seed, env):
    """env: the fallbacks a big CNF would take -- no summary bitmap (> 32768 clauses), no sibling snapshots (no memory)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_random.py"), "200", str(seed)],
                       env=dict(os.environ, **env),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout
