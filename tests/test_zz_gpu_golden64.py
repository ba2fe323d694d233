"""The north-star frontier (rsaFACT-64bit -q 4096) against the UNMODIFIED reference, subproblem by subproblem.

One subproblem of this frontier is ~2.4e8 DFS nodes: hours of one host core for the reference's Satisfy_iterative
(NDP-5_6_7.cpp:784-878), tens of milliseconds for one B200.  tools/make_golden_64bit.py therefore ran the reference on a
handful of hand-picked subproblems (the claimed lowest SAT index, its neighbours, the ends and the middle of the
frontier) and committed verdict / node count / model (and, where the oracle port ran too and agreed on the node count,
clause evaluations) as tests/golden/big_rsaFACT-64bit_q4096_sample.npz.  Here the CUDA path searches exactly those
subproblems -- each as a one-subproblem shard, cut into DFS-order pieces by the automatic policy and searched by the
kernels the 64-bit runs use (lane groups, state slots in global memory) -- and must return the same numbers."""
import os
import re

import numpy as np
import pytest

import util_cnf

pytestmark = pytest.mark.gpu

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big_rsaFACT-64bit_q4096_sample.npz")


def test_64bit_q4096_subproblems_match_the_reference(oracle, cnf_text):
    if not os.path.exists(FIXTURE):
        pytest.skip("no 64-bit reference fixture committed (tools/make_golden_64bit.py takes hours of host time)")
    import ndp5_b200 as nd
    assert nd.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    nd.set_engine("auto")
    nd.set_split("auto")
    z = np.load(FIXTURE)
    text = cnf_text("rsaFACT-64bit")
    A = oracle.parse(text)
    fr = nd.bfs(A, 0, True, 4096)
    assert (fr.iterations, fr.task_count, len(fr)) == (int(z["iterations"]), int(z["task_count"]), int(z["size"]))
    n = len(fr)
    for i, k in enumerate(int(x) for x in z["index"]):
        r = fr.dfs_shard(first=k, stride=n + 1)          # the shard {k}
        assert int(r["verdict"][k]) == int(z["verdict"][i]), ("verdict", k)
        assert int(r["nodes"][k]) == int(z["nodes"][i]), ("nodes", k, int(r["nodes"][k]), int(z["nodes"][i]))
        if int(z["evals"][i]):
            assert int(r["evals"][k]) == int(z["evals"][i]), ("clause evaluations", k)
        assert (np.delete(r["verdict"], k) == 2).all(), ("only the shard's entry is written", k)
        if int(z["verdict"][i]) == 1:
            model = z["model_%d" % k]
            assert r["winner"] == k and r["model"].tolist() == model.tolist(), ("model", k)
            assert util_cnf.satisfies(A, r["model"])
            v1 = [int(x) for x in re.search(r"first input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
            v2 = [int(x) for x in re.search(r"second input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
            d1, d2 = util_cnf.decode_factors(r["model"], v1, v2)
            assert d1 * d2 == 11708736375489890129 and {d1, d2} == {3379623959, 3464508631}
        else:
            assert r["winner"] is None and len(r["model"]) == 0
    fr.free()
