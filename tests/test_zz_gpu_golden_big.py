"""The CUDA path against the reference-produced fixtures for the LARGE settings (tests/golden/big_*.npz, made by
tools/make_golden_big.py from the unmodified reference: verdict and node count per subproblem from Satisfy_iterative,
NDP-5_6_7.cpp:784-878; clause evaluations from the oracle port after its node counts were checked equal).

These are the settings bench.py quotes -- rsaFACT-32bit / prime-32bit -q 16384 (every subproblem), rsaFACT-48bit -q 4096
(an interleaved sample and the prefix up to the lowest SAT index) and BASELINE configs[2] verbatim, rsaFACT-48bit -q 65536
(the prefix up to its lowest SAT index).  bench.py asserts the same fixtures on every run; this test puts them into the
GPU suite, through the C ABI, with the automatic engine / split policy the bench uses."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(p for p in glob.glob(os.path.join(GOLDEN, "big_*.npz")) if not p.endswith("_sample.npz"))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[4:-4] for p in FIXTURES])
def test_exhaustive_dfs_matches_the_reference_fixture(path, oracle, cnf_text):
    import ndp5_b200 as nd
    assert nd.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    nd.set_engine("auto")
    nd.set_split("auto")
    z = np.load(path)
    assert str(z["mode"]) == "q"
    A = oracle.parse(cnf_text(str(z["cnf"])))
    fr = nd.bfs(A, 0, True, int(z["value"]))
    assert (fr.iterations, fr.task_count, len(fr)) == (int(z["iterations"]), int(z["task_count"]), int(z["size"]))
    r = fr.dfs_shard()                                   # the whole frontier, exhaustively (no early exit)
    idx = z["index"]
    assert (r["verdict"] <= 1).all()
    assert (r["verdict"][idx] == z["verdict"]).all(), "verdicts differ from the reference's"
    assert (r["nodes"][idx].astype(np.int64) == z["nodes"]).all(), "DFS node counts differ from the reference's"
    assert (r["evals"][idx] == z["evals"]).all(), "clause-evaluation counts differ from the oracle's"
    sat = np.nonzero(z["verdict"] == 1)[0]
    if len(sat) and (idx[: int(sat[0]) + 1] == np.arange(int(sat[0]) + 1)).all():
        # the fixture is a prefix that reaches the first model: it pins the lowest SAT index of the frontier
        assert r["winner"] == int(idx[sat[0]])
    fr.free()
