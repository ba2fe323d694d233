"""Pins the plain-C oracle to the UNMODIFIED reference compiled in this container
(oracle/_ref/libndp_ref*.so, built by `make -C oracle ref`).  Skipped where the reference
build is absent (the fixtures in tests/golden/ carry the pin there).  CPU only."""
import numpy as np
import pytest

import cbind
import util_cnf

pytestmark = pytest.mark.skipif(not (cbind.RefLib.available() and cbind.RefLib.available(prof=True)),
                                reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return cbind.RefLib(prof=True)


def _same_frontier(a, b):
    assert len(a) == len(b)
    for (c1, h1), (c2, h2) in zip(a, b):
        assert c1.shape == c2.shape and (c1 == c2).all()
        assert h1.shape == h2.shape and (h1 == h2).all()


def test_parser_matches_reference(oracle, ref, cnf_text):
    nasty = "c comment\np cnf 5 6\n1 -2 3 0\n\n-4 0\n1 2 0\n1 2 3 4 0\n 5 0\nx 1 2 3 0\n+6 -7 8 0\r\n\r\n" \
            "99999999999 1 2 0\n1 2 99999999999 0\n12abc 3 4 0\n- 1 2 3 0\n2 -3 4\n7 8 9 0 10 11 0\n%\n0\n2 -3 4"
    for text in [nasty, "", "\n\n", cnf_text("rsaFACT-8bit"), cnf_text("rsaFACT-16bit")]:
        a, b = ref.parse(text), oracle.parse(text)
        assert a.shape == b.shape and (a == b).all()


def test_choice_and_resolution_step_random(oracle, ref):
    rng = np.random.default_rng(7)
    for trial in range(300):
        nv = int(rng.integers(1, 12))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(0, 30)), p_unit=0.15, p_binary=0.2, p_dup=0.1, p_taut=0.1)
        assert oracle.choice(A) == ref.choice(A)
        for i in {int(rng.integers(1, nv + 1)), max(1, oracle.choice(A))}:
            for wc in (0, 1):
                ra = ref.resolution_step(A, i, wc)
                oa = oracle.resolution_step(A, i, wc)
                assert (ra[0] == oa[0]).all() and (ra[1] == oa[1]).all() and ra[2:] == oa[2:]


def test_bfs_random_settings(oracle, ref):
    rng = np.random.default_rng(11)
    for trial in range(120):
        nv = int(rng.integers(2, 14))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(1, 40)), p_unit=0.08, p_binary=0.15)
        for mode, value in [("d", int(rng.integers(-3, 40))), ("q", int(rng.integers(-1, 12)))]:
            D, ovr, Q = util_cnf.settings_for(mode, value)
            r, o = ref.bfs(A, D, ovr, Q), oracle.bfs(A, D, ovr, Q)
            for key in ("iterations", "task_count", "initial_queue_size", "size"):
                assert r[key] == o[key], (trial, mode, value, key)
            _same_frontier(r["frontier"], o["frontier"])


def test_dfs_random(oracle, ref):
    rng = np.random.default_rng(13)
    sat = 0
    for trial in range(300):
        nv = int(rng.integers(1, 16))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(0, 50)), p_unit=0.05, p_binary=0.15)
        r, o = ref.dfs(A), oracle.dfs(A)
        assert r["models"] == o["models"] and r["nodes"] == o["nodes"], trial
        assert (r["model"] == o["model"]).all(), trial
        sat += r["models"]
        if r["models"]:
            assert util_cnf.satisfies(A, r["model"])
    assert 20 < sat < 290


def test_multiplier_kats(oracle, ref, cnf_text):
    """SURVEY.md §8c known answers, recomputed by the reference and by the oracle."""
    A = ref.parse(cnf_text("rsaFACT-16bit"))
    assert A.shape == (1296, 3)
    r = ref.bfs(A, 969, False, -1)
    assert (r["size"], r["iterations"], r["task_count"]) == (32, 969, 1001)
    tot_nodes = 0
    for cl, ch in r["frontier"]:
        a, b = ref.dfs(cl), oracle.dfs(cl)
        assert a["nodes"] == b["nodes"] and a["models"] == b["models"] and (a["model"] == b["model"]).all()
        tot_nodes += a["nodes"]
    assert tot_nodes == 13914
    r = ref.bfs(A, 0, True, 256)
    assert (r["size"], r["iterations"], r["task_count"]) == (0, 15090, 15090)
    hdr = ref.scrape_header(cnf_text("rsaFACT-16bit"))
    assert (hdr["num_vars"], hdr["num_clauses"], hdr["num_bits"], hdr["product"]) == (336, 1296, 9, "35611")
    m = ref.dfs(ref.bfs(A, 1, False, -1)["frontier"][0][0])
    full = np.concatenate([ref.bfs(A, 1, False, -1)["frontier"][0][1], m["model"]])
    d1, d2, ok = ref.convert(full, hdr["v1"], hdr["v2"], hdr["product"])
    assert ok and {d1, d2} == {149, 239}
    assert util_cnf.decode_factors(full, hdr["v1"], hdr["v2"]) == (d1, d2)
