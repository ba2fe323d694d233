#!/usr/bin/env python3
"""Randomised differential stress (GPU box): BFS + DFS of the CUDA path vs the plain-C oracle on random 3-CNFs and on
multiplier CNFs with random -d / -q, random BFS window lengths and random split targets.
    python tools/stress_random.py [trials] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ndp-5_b200/py", "tests", "tools", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import cbind  # noqa: E402
import gen_cnf  # noqa: E402
import util_cnf  # noqa: E402
import ndp5_b200 as nd  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
orc = cbind.OracleLib()
nd.set_engine("index")
mult = {name: orc.parse(gen_cnf.dimacs_text(*gen_cnf.NAMED[name])) for name in
        ("rsaFACT-8bit", "rsaFACT-12bit", "rsaFACT-16bit", "prime-12bit", "prime-16bit", "rsaFACT-20bit")}
t0 = time.time()
bad = 0
for trial in range(trials):
    if trial % 3 == 0:
        name = list(mult)[int(rng.integers(0, len(mult)))]
        A = mult[name]
        mode, value = ("d", int(rng.integers(1, 3000))) if rng.random() < 0.5 else ("q", int(rng.integers(1, 200)))
    else:
        name = "random"
        nv = int(rng.integers(2, 60))
        A = util_cnf.random_cnf(rng, nv, int(rng.integers(1, 220)), p_unit=0.03, p_binary=0.12, p_dup=0.04, p_taut=0.04)
        mode, value = ("d", int(rng.integers(-3, 80))) if rng.random() < 0.5 else ("q", int(rng.integers(-1, 40)))
    D, ovr, Q = util_cnf.settings_for(mode, value)
    os.environ["NDP_BFS_WINDOW"] = str(int(rng.choice([1, 2, 3, 5, 8, 16, 24, 32])))
    nd.set_split(int(rng.choice([-1, 0, 2, 7, 50, 400, 5000])))
    o = orc.bfs(A, D, ovr, Q)
    fr = nd.bfs(A, D, ovr, Q)
    tag = (trial, name, mode, value, os.environ["NDP_BFS_WINDOW"])
    ok = (fr.iterations, fr.task_count, len(fr)) == (o["iterations"], o["task_count"], o["size"])
    ok = ok and fr.bfs_stats()["clause_evals"] == o["clause_evals"]
    if ok:
        ks = range(len(fr)) if len(fr) <= 24 else sorted(set(int(x) for x in rng.integers(0, len(fr), 12)))
        for k in ks:
            c1, h1 = fr.get(k)
            c2, h2 = o["frontier"][k]
            ok = ok and c1.shape == c2.shape and (c1 == c2).all() and h1.shape == h2.shape and (h1 == h2).all()
    if ok and len(fr):
        r = fr.dfs_shard()
        first = None
        for k, (cl, ch) in enumerate(o["frontier"]):
            od = orc.dfs(cl)
            ok = ok and int(r["verdict"][k]) == (1 if od["models"] else 0) and int(r["nodes"][k]) == od["nodes"] \
                and int(r["evals"][k]) == od["clause_evals"]
            if od["models"] and first is None:
                first = (k, np.concatenate([ch, od["model"]]).tolist())
        if first is None:
            ok = ok and r["winner"] is None
        else:
            ok = ok and r["winner"] == first[0] and r["model"].tolist() == first[1]
    if not ok:
        bad += 1
        print("MISMATCH", tag, flush=True)
    fr.free()
print("stress: %d trials, %d mismatches, %.1f s (seed %d)" % (trials, bad, time.time() - t0, seed))
sys.exit(1 if bad else 0)
