"""The plain-C oracle (oracle/ndp_oracle.c) against the committed golden vectors, which the
UNMODIFIED reference produced (tools/make_golden.py).  CPU only."""
import numpy as np
import pytest

import util_cnf

# keep the CPU suite to a few minutes: oracle BFS on these is seconds
BFS_BUDGET_ITERATIONS = 80000
DFS_BUDGET_NODES = 120000


def _cases(golden):
    return golden["cases"]


def test_golden_has_the_headline_cases(golden):
    keys = {(c["cnf"], c["mode"], c["value"]) for c in _cases(golden)}
    for k in [("rsaFACT-16bit", "d", 969), ("rsaFACT-32bit", "d", 12), ("rsaFACT-32bit", "q", 1024),
              ("prime-32bit", "q", 1024), ("rsaFACT-64bit", "d", 20)]:
        assert k in keys


def test_oracle_bfs_matches_reference_golden(oracle, golden, cnf_text):
    checked = 0
    for c in _cases(golden):
        if c["iterations"] > BFS_BUDGET_ITERATIONS:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        assert A.shape[0] == c["n_clauses"]
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        o = oracle.bfs(A, D, ovr, Q)
        key = (c["cnf"], c["mode"], c["value"])
        assert o["iterations"] == c["iterations"], key
        assert o["task_count"] == c["task_count"], key
        assert o["initial_queue_size"] == c["initial_queue_size"], key
        assert o["size"] == c["size"], key
        assert [x.shape[0] for x, _ in o["frontier"]] == c["nclauses"], key
        assert [x.shape[0] for _, x in o["frontier"]] == c["nchoices"], key
        assert util_cnf.frontier_digest(o["frontier"]) == c["frontier_sha256"], key
        assert o["clause_evals"] == c["bfs_clause_evals"], key
        checked += 1
    assert checked >= 50


def test_oracle_dfs_matches_reference_golden(oracle, golden, cnf_text):
    checked = 0
    for c in _cases(golden):
        if "verdict" not in c or c["iterations"] > BFS_BUDGET_ITERATIONS or sum(c["nodes"]) > DFS_BUDGET_NODES:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = oracle.bfs(A, D, ovr, Q)["frontier"]
        key = (c["cnf"], c["mode"], c["value"])
        for k, (cl, ch) in enumerate(fr):
            r = oracle.dfs(cl)
            assert (1 if r["models"] else 0) == c["verdict"][k], (key, k)
            assert r["nodes"] == c["nodes"][k], (key, k)
            if "clause_evals" in c:
                assert r["clause_evals"] == c["clause_evals"][k], (key, k)
            if r["models"]:
                full = np.concatenate([ch, r["model"]])
                assert [int(x) for x in full] == c["sat"][str(k)]["model"], (key, k)
                assert util_cnf.satisfies(A, full)
        checked += 1
    assert checked >= 10


def test_golden_factors_are_verified(golden):
    n = 0
    for c in _cases(golden):
        for k, s in c.get("sat", {}).items():
            assert s["verified"] and int(s["fact1"]) * int(s["fact2"]) == int(c["N"])
            n += 1
    assert n >= 10


def test_parser_rules(oracle):
    """parseDimacsString (NDP:619-642): widths 1 and 3 only, c/p/empty lines skipped."""
    text = "c comment\np cnf 5 6\n1 -2 3 0\n\n-4 0\n1 2 0\n1 2 3 4 0\n 5 0\nx 1 2 3 0\n2 -3 4\n7 8 9 0 10 11 0\n"
    A = oracle.parse(text)
    assert A.tolist() == [[1, -2, 3], [0, 0, -4], [0, 0, 5], [2, -3, 4], [7, 8, 9]]
    assert oracle.parse("").shape == (0, 3)
    assert oracle.parse("99999999999 1 2 0\n1 2 99999999999 0\n").tolist() == []
