"""The DFS-order split (ndp-5_b200/csrc/ndp_split.cuh) changes how the work is cut, never the
answers: for every piece target the per-subproblem verdict, node count, clause-evaluation count,
the winner and its model are the reference's (golden vectors) -- including targets so small that
no split happens and so large that the whole search tree is consumed by the split itself."""
import numpy as np
import pytest

import util_cnf

pytestmark = pytest.mark.gpu

TARGETS = [0, 3, 17, 200, 3000, 40000]


@pytest.fixture(scope="module")
def nd():
    import ndp5_b200
    assert ndp5_b200.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    ndp5_b200.set_engine("index")
    yield ndp5_b200
    ndp5_b200.set_engine("auto")
    ndp5_b200.set_split("auto")


def test_split_targets_golden(nd, oracle, golden, cnf_text):
    n = 0
    for c in golden["cases"]:
        if "verdict" not in c or c["bits"] > 24:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        for target in TARGETS:
            nd.set_split(target)
            key = (c["cnf"], c["mode"], c["value"], target)
            r = fr.dfs_shard()
            assert r["verdict"].tolist() == c["verdict"], key
            assert r["nodes"].tolist() == c["nodes"], key
            assert r["evals"].tolist() == c["clause_evals"], key
            if c["winner"] >= 0:
                assert r["winner"] == c["winner"], key
                assert r["model"].tolist() == c["sat"][str(c["winner"])]["model"], key
            else:
                assert r["winner"] is None, key
            for k, s in c["sat"].items():   # every SAT element as a one-subproblem shard
                rk = fr.dfs_shard(first=int(k), stride=len(fr) + 1)
                assert rk["winner"] == int(k) and rk["model"].tolist() == s["model"], (key, k)
                assert int(rk["nodes"][int(k)]) == c["nodes"][int(k)], (key, k)
            if c["winner"] >= 0:
                re_ = fr.dfs_shard(early_exit=True)
                w = c["winner"]
                assert re_["winner"] == w and re_["model"].tolist() == c["sat"][str(w)]["model"], key
                assert re_["verdict"][: w + 1].tolist() == c["verdict"][: w + 1], key
                assert re_["nodes"][: w + 1].tolist() == c["nodes"][: w + 1], key
        fr.free()
        n += 1
    assert n >= 15


def test_split_random_small(nd, oracle):
    """Random 3-CNFs (units, pre-zeroed literals, duplicates, tautologies): split vs oracle."""
    rng = np.random.default_rng(4242)
    sat = 0
    for trial in range(80):
        nv = int(rng.integers(2, 40))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(1, 150)), p_unit=0.03, p_binary=0.12, p_dup=0.04,
                                p_taut=0.04)
        D, ovr, Q = util_cnf.settings_for("q", int(rng.integers(1, 12)))
        o = oracle.bfs(A, D, ovr, Q)
        fr = nd.bfs(A, D, ovr, Q)
        assert len(fr) == o["size"]
        nd.set_split([5, 64, 1000][trial % 3])
        r = fr.dfs_shard()
        first = None
        for k, (cl, ch) in enumerate(o["frontier"]):
            od = oracle.dfs(cl)
            assert int(r["verdict"][k]) == (1 if od["models"] else 0), (trial, k)
            assert int(r["nodes"][k]) == od["nodes"] and int(r["evals"][k]) == od["clause_evals"], (trial, k)
            if od["models"] and first is None:
                first = (k, np.concatenate([ch, od["model"]]).tolist())
            sat += 1 if od["models"] else 0
        if first is None:
            assert r["winner"] is None
        else:
            assert r["winner"] == first[0] and r["model"].tolist() == first[1], trial
            assert util_cnf.satisfies(A, r["model"])
        fr.free()
    assert sat > 0


def test_split_root_edge_cases(nd, oracle):
    """Roots with the reference's quirks ({0,0,0} clauses, unit-only, contradictory) searched as
    -d -1 frontiers (the root itself is the subproblem), split on."""
    cases = {
        "unit_only": np.array([[0, 0, 1], [0, 0, -2], [0, 0, 3]], dtype=np.int32),
        "contradictory": np.array([[0, 0, 1], [0, 0, -1]], dtype=np.int32),
        "one": np.array([[1, 2, 3]], dtype=np.int32),
        "all_zero_first": np.array([[0, 0, 0], [1, 2, 3]], dtype=np.int32),
        "all_zero_mid": np.array([[0, 0, 4], [0, 0, 0], [1, 2, 3]], dtype=np.int32),
        "dups": np.array([[1, 1, 1], [-1, -1, -1]], dtype=np.int32),
        "taut": np.array([[1, -1, 2], [0, -2, 3], [0, 0, -3]], dtype=np.int32),
    }
    for name, A in cases.items():
        for D in (-1, 1, 3):
            o = oracle.bfs(A, D, False, -1)
            fr = nd.bfs(A, D, False, -1)
            assert len(fr) == o["size"], (name, D)
            for target in (0, 8, 500):
                nd.set_split(target)
                r = fr.dfs_shard()
                for k, (cl, ch) in enumerate(o["frontier"]):
                    od = oracle.dfs(cl)
                    assert int(r["verdict"][k]) == (1 if od["models"] else 0), (name, D, target, k)
                    assert int(r["nodes"][k]) == od["nodes"], (name, D, target, k)
                    assert int(r["evals"][k]) == od["clause_evals"], (name, D, target, k)
                    if od["models"]:
                        rk = fr.dfs_shard(first=k, stride=len(fr) + 1)
                        assert rk["model"].tolist() == np.concatenate([ch, od["model"]]).tolist(), (name, D, target, k)
            fr.free()


def test_split_single_subproblem_32bit_d12(nd, oracle, golden, cnf_text):
    """BASELINE configs[1] verbatim (-d 12 = ONE subproblem): with the automatic split the whole
    device searches it and still stops at the reference's first model after exactly the
    reference's 3 713 280 loop iterations; exhaustive-equivalent answers for two more targets."""
    A = oracle.parse(cnf_text("rsaFACT-32bit"))
    fr = nd.bfs(A, 12, False, -1)
    assert len(fr) == 1
    nd.set_split("off")
    whole = fr.dfs_shard(early_exit=True)
    assert int(whole["nodes"][0]) == 3713280
    for target in ("auto", 1000, 30000):
        nd.set_split(target)
        r = fr.dfs_shard(early_exit=True)
        assert r["winner"] == 0 and int(r["nodes"][0]) == 3713280, target
        assert int(r["evals"][0]) == int(whole["evals"][0]), target
        assert r["model"].tolist() == whole["model"].tolist(), target
    assert util_cnf.satisfies(A, whole["model"])
    fr.free()


def test_split_32bit_q1024_counts(nd, oracle, golden, cnf_text):
    """1024 subproblems cut into ~20000 pieces: per-subproblem counters still the reference's."""
    for name in ("rsaFACT-32bit", "prime-32bit"):
        c = next(x for x in golden["cases"] if (x["cnf"], x["mode"], x["value"]) == (name, "q", 1024))
        A = oracle.parse(cnf_text(name))
        fr = nd.bfs(A, 0, True, 1024)
        nd.set_split(20000)
        r = fr.dfs_shard()
        assert r["verdict"].tolist() == c["verdict"]
        assert r["nodes"].tolist() == c["nodes"]
        assert r["evals"].tolist() == c["clause_evals"]
        if c["winner"] >= 0:
            assert r["winner"] == c["winner"] and r["model"].tolist() == c["sat"][str(c["winner"])]["model"]
        fr.free()


def _emulate_ranks(fr, world, early_exit=False):
    """ndp_dfs_pieces_* as `world` ranks would call it, one after the other on this GPU; the collectives
    (MIN of first_model, MAX of verdicts, SUM of counters) are numpy here."""
    runs = [fr.dfs_pieces(rank=r, world=world, early_exit=early_exit) for r in range(world)]
    fm = np.minimum.reduce([run.first_model for run in runs])
    parts = [run.fold(fm) for run in runs]
    verdict = np.maximum.reduce([p["verdict"] for p in parts])
    nodes = np.sum([p["nodes"] for p in parts], axis=0, dtype=np.uint64)
    evals = np.sum([p["evals"] for p in parts], axis=0, dtype=np.uint64)
    winners = {p["winner"] for p in parts}
    assert len(winners) == 1, "every rank must name the same winner"
    models = [p["model"] for p in parts if p["model"].size]
    for run in runs:
        run.free()
    return verdict, nodes, evals, winners.pop(), models


def test_piece_runs_between_ranks_golden(nd, oracle, golden, cnf_text):
    """Piece-level sharing between processes (ndp_dfs_pieces_begin / fold): whatever the cut and the number
    of ranks, the folded per-subproblem verdicts, node / clause-evaluation counts, the winner and its model are
    the reference's; exactly one rank returns the model."""
    n = 0
    for c in golden["cases"]:
        if "verdict" not in c or c["bits"] > 20:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        for target, world in [(0, 1), (0, 2), (0, 3), (50, 2), (700, 3), (5000, 8), ("auto", 4)]:
            nd.set_split(target)
            key = (c["cnf"], c["mode"], c["value"], target, world)
            verdict, nodes, evals, winner, models = _emulate_ranks(fr, world)
            assert verdict.tolist() == c["verdict"], key
            assert nodes.tolist() == c["nodes"], key
            assert evals.tolist() == c["clause_evals"], key
            if c["winner"] >= 0:
                assert winner == c["winner"], key
                assert len(models) == 1 and models[0].tolist() == c["sat"][str(c["winner"])]["model"], key
            else:
                assert winner is None and not models, key
        fr.free()
        n += 1
    assert n >= 10
