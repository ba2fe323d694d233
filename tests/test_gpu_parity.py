"""Parity of the CUDA path (through the C ABI, include/ndp_b200.h) against the oracle and the
golden vectors the unmodified reference produced.  Bar: bit-exact -- same BFS frontier in the
same order (clause lists with their zeros, choices), same iterations/task_count, same
SAT/UNSAT verdict per subproblem, same model, and (stronger than the reference's API needs)
the same DFS node count and clause-evaluation count per subproblem."""
import numpy as np
import pytest

import util_cnf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["index", "index-whole", "index-global", "index-warp", "scan"])
def nd(request):
    """Every parity test runs once per DFS engine / kernel: all must match the oracle bit for bit.
    "index" cuts narrow shards into DFS-order pieces (automatic policy: every test frontier is
    narrower than the device, so the split path is what runs) and, the test CNFs being small, searches
    them with the lane-group kernel (4 pieces per warp, state slots in shared memory); "index-whole"
    searches each subproblem whole; "index-global" forces the lane-group kernel with its state slots in
    global memory (what large CNFs get); "index-warp" forces the one-piece-per-warp kernel; "scan" is the
    clause-streaming engine."""
    import os
    import ndp5_b200
    assert ndp5_b200.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    ndp5_b200.set_engine(request.param.split("-")[0])
    ndp5_b200.set_split("off" if request.param == "index-whole" else "auto")
    env = {"index-global": {"NDP_DFS_GLOBAL_STATE": "1"}, "index-warp": {"NDP_DFS_GROUPS": "1", "NDP_DFS_GLOBAL_STATE": "0"}}
    for k, v in env.get(request.param, {}).items():
        os.environ[k] = v
    yield ndp5_b200
    for k in env.get(request.param, {}):
        os.environ.pop(k, None)
    ndp5_b200.set_engine("auto")
    ndp5_b200.set_split("auto")


def gpu_frontier_list(fr):
    return [fr.get(k) for k in range(len(fr))]


def check_bfs_against_oracle(nd, oracle, A, D, ovr, Q, tag):
    o = oracle.bfs(A, D, ovr, Q)
    fr = nd.bfs(A, D, ovr, Q)
    assert fr.iterations == o["iterations"], tag
    assert fr.task_count == o["task_count"], tag
    assert fr.initial_queue_size == o["initial_queue_size"], tag
    assert len(fr) == o["size"], tag
    got = gpu_frontier_list(fr)
    for k, ((c1, h1), (c2, h2)) in enumerate(zip(got, o["frontier"])):
        assert c1.shape == c2.shape and (c1 == c2).all(), (tag, k, "clauses")
        assert h1.shape == h2.shape and (h1 == h2).all(), (tag, k, "choices")
        assert fr.dims(k) == (c2.shape[0], h2.shape[0])
    assert fr.bfs_stats()["clause_evals"] == o["clause_evals"], tag
    return fr, o


# ------------------------------------------------------------------------------------------ BFS
def test_bfs_golden_all_cases(nd, oracle, golden, cnf_text):
    """Every golden BFS case: counters, per-element sizes and the frontier digest."""
    for c in golden["cases"]:
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        key = (c["cnf"], c["mode"], c["value"])
        assert (fr.iterations, fr.task_count, fr.initial_queue_size, len(fr)) == \
            (c["iterations"], c["task_count"], c["initial_queue_size"], c["size"]), key
        assert [fr.dims(k) for k in range(len(fr))] == list(zip(c["nclauses"], c["nchoices"])), key
        assert util_cnf.frontier_digest(gpu_frontier_list(fr)) == c["frontier_sha256"], key
        assert fr.bfs_stats()["clause_evals"] == c["bfs_clause_evals"], key
        fr.free()


def test_bfs_random_small(nd, oracle):
    """Random 3-CNFs with units, pre-zeroed literals, duplicates, tautologies; -d and -q sweeps
    including the degenerate settings (-d -1, negative -d, -q 0/1)."""
    rng = np.random.default_rng(2024)
    for trial in range(60):
        nv = int(rng.integers(2, 40))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(1, 120)), p_unit=0.05, p_binary=0.15, p_dup=0.05,
                                p_taut=0.05)
        for mode, value in [("d", int(rng.integers(-3, 60))), ("q", int(rng.integers(-1, 20)))]:
            D, ovr, Q = util_cnf.settings_for(mode, value)
            fr, _ = check_bfs_against_oracle(nd, oracle, A, D, ovr, Q, (trial, mode, value))
            fr.free()


def test_bfs_edge_cases(nd, oracle):
    unit_only = np.array([[0, 0, 1], [0, 0, -2], [0, 0, 3]], dtype=np.int32)
    contradictory = np.array([[0, 0, 1], [0, 0, -1]], dtype=np.int32)
    one = np.array([[1, 2, 3]], dtype=np.int32)
    all_zero_first = np.array([[0, 0, 0], [1, 2, 3]], dtype=np.int32)
    all_zero_mid = np.array([[0, 0, 4], [0, 0, 0], [1, 2, 3]], dtype=np.int32)
    sparse_ids = np.array([[30000, -32766, 5], [0, 0, -30000], [0, 32766, 7]], dtype=np.int32)
    for name, A in [("unit_only", unit_only), ("contradictory", contradictory), ("one", one),
                    ("all_zero_first", all_zero_first), ("all_zero_mid", all_zero_mid), ("sparse_ids", sparse_ids)]:
        for D, ovr, Q in [(-1, False, -1), (1, False, -1), (5, False, -1), (100, False, -1), (0, True, 1),
                          (0, True, 2), (0, True, 100), (0, False, -1), (-7, False, -1), (0, True, 0)]:
            fr, _ = check_bfs_against_oracle(nd, oracle, A, D, ovr, Q, (name, D, ovr, Q))
            fr.free()


def test_sparse_variable_ids_up_to_2_pow_30(nd, oracle, cnf_text):
    """The reference's Clause3 is `int l[3]` (ClauseSetPool.hpp:37-39): any variable id is legal.  The kernels pack
    literals in 16-bit codes, so the library renumbers at the ABI when the ids do not fit (only variables that occur
    need a code) and maps every literal back on the way out: frontier clauses, choices and models carry the
    caller's ids.  Parity against the oracle on the caller's CNF, BFS + DFS + resume with root."""
    rng = np.random.default_rng(77)
    cases = []
    for trial in range(12):
        nv = int(rng.integers(3, 40))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(5, 150)), p_unit=0.03, p_binary=0.12, p_dup=0.04, p_taut=0.04)
        cases.append(A)
    cases.append(oracle.parse(cnf_text("rsaFACT-12bit")))
    cases.append(oracle.parse(cnf_text("prime-12bit")))
    for trial, A in enumerate(cases):
        ids = np.unique(np.abs(A[A != 0]))
        big = rng.choice(np.arange(40000, 1 << 30, 1 << 12), size=len(ids), replace=False) + rng.integers(0, 1 << 12, len(ids))
        if trial % 3 == 0:
            big[0] = (1 << 30)                                  # the largest id the test names
        if trial % 3 == 1:
            big[: len(big) // 2] = ids[: len(big) // 2]         # a mix of small and huge ids
        lut = dict(zip(ids.tolist(), big.tolist()))
        B = np.vectorize(lambda l: 0 if l == 0 else (lut[abs(l)] if l > 0 else -lut[abs(l)]))(A).astype(np.int32)
        for D, ovr, Q in [(0, True, int(rng.integers(2, 12))), (int(rng.integers(1, 60)), False, -1)]:
            fr, o = check_bfs_against_oracle(nd, oracle, B, D, ovr, Q, ("sparse", trial, D, Q))
            r = check_dfs_against_oracle(nd, oracle, fr, o["frontier"], ("sparse", trial, D, Q))
            if r["winner"] is not None:
                assert util_cnf.satisfies(B, r["model"])
            if len(fr):
                fr2 = nd.frontier_from_host(o["frontier"])       # resume path with the caller's ids, then its root
                assert fr2.attach_root(B) is True
                r2 = fr2.dfs_shard()
                assert r2["verdict"].tolist() == r["verdict"].tolist() and r2["nodes"].tolist() == r["nodes"].tolist()
                assert r2["winner"] == r["winner"] and r2["model"].tolist() == r["model"].tolist()
                c2, h2 = fr2.get(len(fr) - 1)
                assert (c2 == o["frontier"][-1][0]).all() and (h2 == o["frontier"][-1][1]).all()
                fr2.free()
            fr.free()


def test_too_many_distinct_variables_fails_loudly(nd):
    """What remains unsupported is a resource limit, not an id range: more than 32766 DISTINCT variables."""
    v = np.arange(1, 3 * 11000 + 1, dtype=np.int32).reshape(-1, 3) + 100000
    with pytest.raises(nd.NdpError) as ei:
        nd.bfs(v, 1, False, -1)
    assert ei.value.code == nd.NDP_E_UNSUPPORTED
    with pytest.raises(nd.NdpError) as ei:
        nd.frontier_from_host([(v, np.zeros(0, np.int32))])
    assert ei.value.code == nd.NDP_E_UNSUPPORTED
    with pytest.raises(nd.NdpError) as ei:
        nd.bfs(np.array([[1, 2, -2147483648]], dtype=np.int32), 1, False, -1)
    assert ei.value.code == nd.NDP_E_UNSUPPORTED


# ------------------------------------------------------------------------------------------ DFS
def check_dfs_against_oracle(nd, oracle, fr, ofr, tag):
    r = fr.dfs_shard()
    first_sat = None
    for k, (cl, ch) in enumerate(ofr):
        od = oracle.dfs(cl)
        assert int(r["verdict"][k]) == (1 if od["models"] else 0), (tag, k, "verdict")
        assert int(r["nodes"][k]) == od["nodes"], (tag, k, "nodes")
        assert int(r["evals"][k]) == od["clause_evals"], (tag, k, "clause_evals")
        if od["models"] and first_sat is None:
            first_sat = (k, np.concatenate([ch, od["model"]]))
    if first_sat is None:
        assert r["winner"] is None and r["model"].size == 0, tag
    else:
        assert r["winner"] == first_sat[0], tag
        assert r["model"].tolist() == first_sat[1].tolist(), tag
    return r


def test_dfs_golden_small_and_medium(nd, oracle, golden, cnf_text):
    """Golden DFS cases up to 24-bit: verdict vector, node counts, clause evals, every SAT
    subproblem's model (by making it the winner through a one-element shard)."""
    n = 0
    for c in golden["cases"]:
        if "verdict" not in c or c["bits"] > 24:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        r = fr.dfs_shard()
        key = (c["cnf"], c["mode"], c["value"])
        assert r["verdict"].tolist() == c["verdict"], key
        assert r["nodes"].tolist() == c["nodes"], key
        assert r["evals"].tolist() == c["clause_evals"], key
        assert r["stats"]["nodes"] == sum(c["nodes"]) and r["stats"]["clause_evals"] == sum(c["clause_evals"]), key
        if c["winner"] >= 0:
            assert r["winner"] == c["winner"], key
            assert r["model"].tolist() == c["sat"][str(c["winner"])]["model"], key
        else:
            assert r["winner"] is None, key
        for k, s in c["sat"].items():   # every SAT element, not just the winner
            rk = fr.dfs_shard(first=int(k), stride=len(fr) + 1)
            assert rk["winner"] == int(k) and rk["model"].tolist() == s["model"], (key, k)
            assert util_cnf.satisfies(A, rk["model"]), (key, k)
        fr.free()
        n += 1
    assert n >= 15


def test_dfs_generic_kernel_on_generated_cnfs(nd, golden, oracle, cnf_text, monkeypatch):
    """Generated factoring CNFs take the kernels' kPlain instantiation (B[2] summary, no {0,0,0} clause, no
    variable twice in a clause).  NDP_NO_PLAIN forces the generic instantiation on the same inputs: the
    golden verdicts, node and clause-evaluation counts must not move."""
    monkeypatch.setenv("NDP_NO_PLAIN", "1")
    n = 0
    for c in golden["cases"]:
        if "verdict" not in c or c["bits"] not in (12, 16, 20):
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        r = fr.dfs_shard()
        key = (c["cnf"], c["mode"], c["value"])
        assert r["verdict"].tolist() == c["verdict"], key
        assert r["nodes"].tolist() == c["nodes"] and r["evals"].tolist() == c["clause_evals"], key
        fr.free()
        n += 1
    assert n >= 8


def test_bfs_level_wider_than_the_device_loop_can_hold(nd, golden, oracle, cnf_text, monkeypatch):
    """A level with more nodes per CTA than the persistent BFS kernel's shared scan arrays hold (2048 x SMs:
    -q above ~300k) must go through the per-level kernels instead of relaunching the device loop for ever
    (round-1 advisor finding).  NDP_PERSIST_BLOCK_NODES lowers the limit so small inputs take that path."""
    monkeypatch.setenv("NDP_PERSIST_BLOCK_NODES", "1")
    n = 0
    for c in golden["cases"]:
        if c["cnf"] not in ("rsaFACT-16bit", "prime-16bit", "rsaFACT-24bit") or c["size"] < 8:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        key = (c["cnf"], c["mode"], c["value"])
        assert (fr.iterations, fr.task_count, len(fr)) == (c["iterations"], c["task_count"], c["size"]), key
        assert util_cnf.frontier_digest(gpu_frontier_list(fr)) == c["frontier_sha256"], key
        assert fr.bfs_stats()["clause_evals"] == c["bfs_clause_evals"], key
        fr.free()
        n += 1
    assert n >= 6


def test_dfs_random_small(nd, oracle):
    rng = np.random.default_rng(99)
    sat = 0
    for trial in range(60):
        nv = int(rng.integers(2, 40))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(1, 150)), p_unit=0.03, p_binary=0.12, p_dup=0.04,
                                p_taut=0.04)
        D, ovr, Q = util_cnf.settings_for("q", int(rng.integers(1, 12)))
        fr, o = check_bfs_against_oracle(nd, oracle, A, D, ovr, Q, ("dfs-random", trial))
        r = check_dfs_against_oracle(nd, oracle, fr, o["frontier"], ("dfs-random", trial))
        sat += int((r["verdict"] == 1).sum())
        if r["winner"] is not None:
            assert util_cnf.satisfies(A, r["model"])
        fr.free()
    assert sat > 0


def test_dfs_edge_cases_through_resume_path(nd, oracle):
    """ndp_frontier_from_host (the -r path) with hand-made subproblems: empty set (model = its
    choices), a set whose first clause is {0,0,0} (the reference's choice()==0 quirk), unit-only,
    contradictory units, duplicates, a ragged mix of sizes."""
    items = [
        (np.zeros((0, 3), np.int32), np.array([5, -6], np.int32)),
        (np.array([[0, 0, 0], [1, 2, 3]], np.int32), np.array([9], np.int32)),
        (np.array([[0, 0, 1], [0, 0, -2], [0, 0, 3]], np.int32), np.zeros(0, np.int32)),
        (np.array([[0, 0, 1], [0, 0, -1]], np.int32), np.array([4], np.int32)),
        (np.array([[1, 1, 1], [-1, -1, -1]], np.int32), np.zeros(0, np.int32)),
        (np.array([[1, 2, 3]], np.int32), np.array([-7, 8, 9], np.int32)),
        (np.array([[0, 0, 4], [0, 0, 0], [1, 2, 3]], np.int32), np.zeros(0, np.int32)),
        (np.array([[1, -2, 3], [-1, 2, 0], [0, -3, 2], [1, 2, 3], [-1, -2, -3]], np.int32), np.array([11], np.int32)),
    ]
    fr = nd.frontier_from_host(items)
    assert len(fr) == len(items)
    for k, (cl, ch) in enumerate(items):
        c2, h2 = fr.get(k)
        assert (c2 == cl).all() and (h2 == ch).all()
    r = fr.dfs_shard()
    for k, (cl, ch) in enumerate(items):
        od = oracle.dfs(cl)
        assert int(r["verdict"][k]) == (1 if od["models"] else 0), k
        assert int(r["nodes"][k]) == od["nodes"] and int(r["evals"][k]) == od["clause_evals"], k
        rk = fr.dfs_shard(first=k, stride=len(items) + 1)
        if od["models"]:
            assert rk["winner"] == k and rk["model"].tolist() == np.concatenate([ch, od["model"]]).tolist(), k
        else:
            assert rk["winner"] is None
    fr.free()
    empty = nd.frontier_from_host([])
    assert len(empty) == 0
    r = empty.dfs_shard()
    assert r["winner"] is None and r["verdict"].size == 0


def test_resume_equals_straight_run(nd, oracle, cnf_text):
    """-s / -r round trip at the ABI: a frontier rebuilt from its own exported elements gives the
    same verdicts, counters and model (BASELINE config 5's save/resume requirement)."""
    for name, q in [("prime-16bit", 64), ("rsaFACT-16bit", 32)]:
        A = oracle.parse(cnf_text(name))
        fr = nd.bfs(A, 0, True, q)
        r1 = fr.dfs_shard()
        fr2 = nd.frontier_from_host(gpu_frontier_list(fr))
        r2 = fr2.dfs_shard()
        assert r1["verdict"].tolist() == r2["verdict"].tolist()
        assert r1["nodes"].tolist() == r2["nodes"].tolist() and r1["evals"].tolist() == r2["evals"].tolist()
        assert r1["winner"] == r2["winner"] and r1["model"].tolist() == r2["model"].tolist()
        fr.free()
        fr2.free()


def test_resume_with_root_uses_the_index_engine(nd, oracle, cnf_text):
    """-r with the DIMACS root attached (ndp_frontier_attach_root): the library proves on the device
    that the loaded frontier is the root filtered by its choices, then searches it like a straight
    run -- same verdicts, counters and model.  A foreign root, a tampered clause or a tampered choice
    is refused and the frontier stays searchable as loaded."""
    for name, q in [("prime-16bit", 64), ("rsaFACT-16bit", 32), ("rsaFACT-24bit", 100)]:
        A = oracle.parse(cnf_text(name))
        fr = nd.bfs(A, 0, True, q)
        r1 = fr.dfs_shard()
        items = gpu_frontier_list(fr)
        fr2 = nd.frontier_from_host(items)
        assert fr2.attach_root(A) is True
        assert fr2.attach_root(A) is True          # idempotent
        for k in (0, len(items) - 1):
            c2, h2 = fr2.get(k)
            assert (c2 == items[k][0]).all() and (h2 == items[k][1]).all()
        r2 = fr2.dfs_shard()
        assert r1["verdict"].tolist() == r2["verdict"].tolist()
        assert r1["nodes"].tolist() == r2["nodes"].tolist() and r1["evals"].tolist() == r2["evals"].tolist()
        assert r1["winner"] == r2["winner"] and r1["model"].tolist() == r2["model"].tolist()
        # refused: another CNF of the same shape, a flipped literal in one clause list, a flipped choice
        other = oracle.parse(cnf_text("prime-16bit" if name != "prime-16bit" else "rsaFACT-16bit"))
        bad_clause = [(c.copy(), h.copy()) for c, h in items]
        bad_clause[len(items) // 2][0][3, 1] *= -1
        bad_choice = [(c.copy(), h.copy()) for c, h in items]
        bad_choice[1][1][-1] *= -1
        for label, its, root in [("other root", items, other), ("clause", bad_clause, A), ("choice", bad_choice, A)]:
            fr3 = nd.frontier_from_host(its)
            assert fr3.attach_root(root) is False, (name, label)
            r3 = fr3.dfs_shard()                       # still searchable, on exactly the loaded sets
            for k in (0, 1, len(its) // 2):
                od = oracle.dfs(its[k][0])
                assert int(r3["verdict"][k]) == (1 if od["models"] else 0), (name, label, k)
                assert int(r3["nodes"][k]) == od["nodes"], (name, label, k)
            fr3.free()
        fr.free()
        fr2.free()


def test_sharding_does_not_change_answers(nd, oracle, cnf_text):
    """Static interleaved shards k -> k mod G (G = 1,2,4,8) reproduce the single-shard verdict
    vector; the global winner is the minimum over shards."""
    A = oracle.parse(cnf_text("rsaFACT-16bit"))
    fr = nd.bfs(A, 2240, False, -1)
    full = fr.dfs_shard()
    for G in (2, 4, 8):
        verdict = np.full(len(fr), 2, np.uint8)
        winners = []
        for rank in range(G):
            r = fr.dfs_shard(first=rank, stride=G)
            mine = np.arange(rank, len(fr), G)
            verdict[mine] = r["verdict"][mine]
            assert (r["verdict"][np.setdiff1d(np.arange(len(fr)), mine)] == 2).all()
            if r["winner"] is not None:
                winners.append((r["winner"], r["model"].tolist()))
        assert verdict.tolist() == full["verdict"].tolist()
        assert min(winners)[0] == full["winner"] and min(winners)[1] == full["model"].tolist()
    fr.free()


def test_early_exit_is_deterministic(nd, oracle, golden, cnf_text):
    """early_exit=1: the winner is the lowest-index SAT subproblem and its model is the
    exhaustive run's; everything at or below the winner is resolved, nothing above it is SAT
    with a lower index."""
    for c in golden["cases"]:
        if "verdict" not in c or c["winner"] < 0 or c["bits"] > 24:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        for _ in range(3):
            r = fr.dfs_shard(early_exit=True)
            w = c["winner"]
            assert r["winner"] == w and r["model"].tolist() == c["sat"][str(w)]["model"]
            assert r["verdict"][: w + 1].tolist() == c["verdict"][: w + 1]
            assert r["exit_flag"] == w
        fr.free()


# ------------------------------------------------------------------------------------ full size
def test_full_size_32bit_q1024(nd, oracle, golden, cnf_text):
    """BASELINE configs[1]'s CNF at the bench's frontier width: 1024 subproblems, exhaustive.
    Golden verdicts / node counts are the reference's; clause evals the oracle's.  Also the
    size-independent properties: the model satisfies the ORIGINAL CNF and decodes to the factors."""
    for name in ("rsaFACT-32bit", "prime-32bit"):
        c = next(x for x in golden["cases"] if (x["cnf"], x["mode"], x["value"]) == (name, "q", 1024))
        A = oracle.parse(cnf_text(name))
        fr = nd.bfs(A, 0, True, 1024)
        assert (fr.iterations, fr.task_count, len(fr)) == (c["iterations"], c["task_count"], 1024)
        assert util_cnf.frontier_digest(gpu_frontier_list(fr)) == c["frontier_sha256"]
        r = fr.dfs_shard()
        assert r["verdict"].tolist() == c["verdict"]
        assert r["nodes"].tolist() == c["nodes"]
        assert r["evals"].tolist() == c["clause_evals"]
        if c["winner"] >= 0:
            assert r["winner"] == c["winner"]
            s = c["sat"][str(c["winner"])]
            assert r["model"].tolist() == s["model"]
            assert util_cnf.satisfies(A, r["model"])
            import re
            text = cnf_text(name)
            v1 = [int(x) for x in re.search(r"first input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
            v2 = [int(x) for x in re.search(r"second input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
            d1, d2 = util_cnf.decode_factors(r["model"], v1, v2)
            assert d1 * d2 == int(c["N"]) and (str(d1), str(d2)) == (s["fact1"], s["fact2"])
        else:
            assert r["winner"] is None and not r["verdict"].any()
        fr.free()


def test_single_subproblem_32bit_d12(nd, oracle, golden, cnf_text):
    """BASELINE configs[1] verbatim: -d 12 gives ONE subproblem (SURVEY.md finding 1); the DFS to
    the first model is 3 713 280 reference loop iterations (SURVEY.md §8c KAT)."""
    A = oracle.parse(cnf_text("rsaFACT-32bit"))
    fr = nd.bfs(A, 12, False, -1)
    assert len(fr) == 1 and fr.dims(0) == (5629, 12)
    r = fr.dfs_shard(early_exit=True)
    assert r["winner"] == 0 and int(r["nodes"][0]) == 3713280
    assert util_cnf.satisfies(A, r["model"])
    import re
    text = cnf_text("rsaFACT-32bit")
    v1 = [int(x) for x in re.search(r"first input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
    v2 = [int(x) for x in re.search(r"second input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
    d1, d2 = util_cnf.decode_factors(r["model"], v1, v2)
    assert {d1, d2} == {45887, 54493}
    fr.free()


def test_full_size_48bit_and_64bit(nd, oracle, cnf_text, request):
    """BASELINE configs[2] and [3] at full size.  The reference cannot finish these searches on host
    cores, so parity rests on (a) the BFS counters it did produce (48-bit -q 4096 ->
    395 680 expansions / 399 776 tasks, 4096 x 9934 clauses; 64-bit -q 4096 -> 395 696 / 399 792,
    4096 x 20 014 clauses, 643 choices) and (b) size-independent properties: the model satisfies the
    ORIGINAL CNF and its decoded factors multiply to N; both engines agree on the winner."""
    import re
    A = oracle.parse(cnf_text("rsaFACT-64bit"))
    fr = nd.bfs(A, 0, True, 4096)
    assert (fr.iterations, fr.task_count, len(fr)) == (395696, 399792, 4096)
    assert fr.dims(0) == (20014, 643) and fr.dims(4095)[0] == 20014
    fr.free()
    text = cnf_text("rsaFACT-48bit")
    A = oracle.parse(text)
    fr = nd.bfs(A, 0, True, 4096)
    assert (fr.iterations, fr.task_count, len(fr)) == (395680, 399776, 4096)
    assert fr.dims(0) == (9934, 627)
    cl, ch = fr.get(17)
    o = oracle.bfs(A, 0, True, 64)          # a prefix the oracle finishes quickly, for the clause-list format
    assert len(o["frontier"]) == 64 and cl.shape == (9934, 3) and ch.shape == (627,)
    if request.node.callspec.params["nd"] != "scan":   # 4.1e8 nodes: the streaming engine holds 2 warps/SM at this size
        r = fr.dfs_shard(early_exit=True)
        assert r["winner"] == 301
        assert util_cnf.satisfies(A, r["model"])
        v1 = [int(x) for x in re.search(r"first input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
        v2 = [int(x) for x in re.search(r"second input \[msb,...,lsb\]: \[(.*?)\]", text).group(1).split(",")]
        d1, d2 = util_cnf.decode_factors(r["model"], v1, v2)
        assert d1 * d2 == 209727620799133 and {d1, d2} == {15908531, 13183343}
        assert (r["verdict"][:301] == 0).all() and r["verdict"][301] == 1
    fr.free()


def test_bfs_96bit_beyond_the_summary_bitmap(nd, oracle):
    """A 96-bit multiplier CNF (53 856 clauses, 13 680 variables): more than 32 768 clauses, so the one-word-per-32-words
    summary of the unit-clause plane no longer fits a warp and the engines scan the plane instead; granulation and a
    short DFS prefix against the oracle."""
    import gen_cnf
    text = gen_cnf.dimacs_text((1 << 95) + 12345678901234567, 96)   # any 96-bit product encodes; it need not be a semiprime
    A = oracle.parse(text)
    assert A.shape[0] > 32768
    for mode, value in (("d", 300), ("q", 24)):
        D, ovr, Q = util_cnf.settings_for(mode, value)
        fr, o = check_bfs_against_oracle(nd, oracle, A, D, ovr, Q, ("96bit", mode, value))
        fr.free()


def test_exit_word_single_process(nd, oracle, golden, cnf_text):
    """ndp_dfs_shard_shared with this process's own early-exit word and no peers equals ndp_dfs_shard(early_exit):
    same winner, model and verdicts up to the winner; the word ends up holding the winner and can be reset.
    (The cross-process exchange needs two GPUs: tests/test_gpu_multi.py / tools/check_shared_exit.py.)"""
    word = nd.ExitWord()
    assert word.read() is None and len(word.export()) == nd.EXIT_HANDLE_BYTES
    n = 0
    for c in golden["cases"]:
        if "verdict" not in c or c["bits"] > 20:
            continue
        A = oracle.parse(cnf_text(c["cnf"]))
        D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
        fr = nd.bfs(A, D, ovr, Q)
        word.reset()
        r = fr.dfs_shard_shared(word, [], early_exit=True)
        w = c["winner"]
        if w >= 0:
            assert r["winner"] == w and r["model"].tolist() == c["sat"][str(w)]["model"]
            assert r["verdict"][: w + 1].tolist() == c["verdict"][: w + 1]
            assert word.read() == w
        else:
            assert r["winner"] is None and word.read() is None and r["verdict"].tolist() == c["verdict"]
        # a word that already knows a lower index makes the whole shard stand down
        if w > 0:
            r2 = fr.dfs_shard_shared(word, [], first=w, stride=len(fr) + 1, early_exit=True)   # word still holds w: not below it
            assert r2["winner"] == w
        fr.free()
        n += 1
    assert n >= 8
    word.free()
