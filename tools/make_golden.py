#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libndp_ref*.so).

Run in the build container, where /root/reference exists and `make -C oracle ref` has been
run.  The fixtures are what pins parity on machines without the reference (the GPU box):
every number in them was produced by the reference's own Satisfy_iterative_BFS /
Satisfy_iterative / convert (via oracle/ref_harness.cpp), except the `clause_evals` fields,
which the reference cannot report and which come from the plain-C oracle AFTER its node counts
were checked equal to the reference's PROFILE_SCOPE counts.

    python tools/make_golden.py            # writes tests/golden/golden.json
"""
import hashlib
import json
import multiprocessing
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cbind  # noqa: E402
import gen_cnf  # noqa: E402


def frontier_digest(frontier):
    """sha256 over (n, m, clauses int32 LE, choices int32 LE) of every element in queue order."""
    h = hashlib.sha256()
    for cl, ch in frontier:
        h.update(np.asarray([cl.shape[0], ch.shape[0]], dtype="<i4").tobytes())
        h.update(np.ascontiguousarray(cl, dtype="<i4").tobytes())
        h.update(np.ascontiguousarray(ch, dtype="<i4").tobytes())
    return h.hexdigest()


def element_digest(cl, ch):
    return frontier_digest([(cl, ch)])[:16]


# (cnf, mode, value, dfs?)  mode 'd': -d value ; mode 'q': -q value
BFS_CASES = []
for d in [-5, -1, 1, 2, 12, 60, 99, 100, 101, 333, 969, 1200, 2239, 2240, 2241, 5000, 20000]:
    BFS_CASES.append(("rsaFACT-16bit", "d", d))
for q in [1, 2, 3, 5, 8, 9, 31, 32, 33, 64, 100, 127, 256]:
    BFS_CASES.append(("rsaFACT-16bit", "q", q))
for d in [1, 7, 40, 150, 500, 4000]:
    BFS_CASES.append(("rsaFACT-8bit", "d", d))
    BFS_CASES.append(("rsaFACT-12bit", "d", d))
for q in [2, 4, 7, 16, 50]:
    BFS_CASES.append(("rsaFACT-8bit", "q", q))
    BFS_CASES.append(("rsaFACT-12bit", "q", q))
BFS_CASES += [("prime-16bit", "d", 2240), ("prime-16bit", "q", 64), ("prime-8bit", "q", 8), ("prime-12bit", "q", 16),
              ("rsaFACT-20bit", "q", 128), ("rsaFACT-24bit", "q", 256), ("prime-24bit", "q", 256),
              ("rsaFACT-32bit", "d", 12), ("rsaFACT-32bit", "q", 64), ("rsaFACT-32bit", "q", 1024),
              ("prime-32bit", "q", 1024), ("rsaFACT-48bit", "q", 512), ("rsaFACT-64bit", "d", 20),
              ("rsaFACT-64bit", "q", 256)]

# cases whose whole frontier is also searched by the reference DFS (exhaustively)
DFS_CASES = {("rsaFACT-16bit", "d", 1), ("rsaFACT-16bit", "d", 12), ("rsaFACT-16bit", "d", 60),
             ("rsaFACT-16bit", "d", 969), ("rsaFACT-16bit", "d", 2240), ("rsaFACT-16bit", "q", 32),
             ("rsaFACT-16bit", "q", 100), ("rsaFACT-16bit", "d", -1),
             ("rsaFACT-8bit", "d", 1), ("rsaFACT-8bit", "q", 7), ("rsaFACT-12bit", "d", 40), ("rsaFACT-12bit", "q", 16),
             ("prime-8bit", "q", 8), ("prime-12bit", "q", 16), ("prime-16bit", "d", 2240), ("prime-16bit", "q", 64),
             ("rsaFACT-20bit", "q", 128), ("rsaFACT-24bit", "q", 256), ("prime-24bit", "q", 256),
             ("rsaFACT-32bit", "q", 1024), ("prime-32bit", "q", 1024)}


_FR = None
_ORC = None


def _oracle_dfs(k):
    global _ORC
    if _ORC is None:
        _ORC = cbind.OracleLib()
    od = _ORC.dfs(_FR[k][0])
    return od["nodes"], od["models"], od["clause_evals"]


def main():
    ref = cbind.RefLib()
    refp = cbind.RefLib(prof=True)
    orc = cbind.OracleLib()
    workers = max(1, (os.cpu_count() or 2) - 1)
    out = {"generator": "tools/make_golden.py", "source": "reference NDP-5_6_7.cpp via oracle/ref_harness.cpp",
           "cases": []}
    texts = {}
    for cnf, mode, value in BFS_CASES:
        if cnf not in texts:
            texts[cnf] = gen_cnf.dimacs_text(*gen_cnf.NAMED[cnf])
        text = texts[cnf]
        A = ref.parse(text)
        hdr = ref.scrape_header(text)
        D, ovr, Q = (value, False, -1) if mode == "d" else (0, True, value)
        t0 = time.time()
        h, info = refp.bfs_handle(A, D, ovr, Q)
        fr = refp.frontier_list(h)
        case = {"cnf": cnf, "N": str(gen_cnf.NAMED[cnf][0]), "bits": gen_cnf.NAMED[cnf][1], "mode": mode,
                "value": value, "n_clauses": int(A.shape[0]), "num_vars": hdr["num_vars"],
                "iterations": info["iterations"], "task_count": info["task_count"],
                "initial_queue_size": info["initial_queue_size"], "size": len(fr),
                "frontier_sha256": frontier_digest(fr),
                "nclauses": [int(c.shape[0]) for c, _ in fr], "nchoices": [int(c.shape[0]) for _, c in fr],
                "element_sha": [element_digest(c, ch) for c, ch in fr] if len(fr) <= 256 else None}
        o = orc.bfs(A, D, ovr, Q, materialise=False)
        assert (o["iterations"], o["task_count"], o["size"]) == (info["iterations"], info["task_count"], len(fr))
        case["bfs_clause_evals"] = o["clause_evals"]
        if (cnf, mode, value) in DFS_CASES and len(fr) > 0:
            r = refp.dfs_frontier(h, 0, len(fr), workers, early_exit=False)
            case["verdict"] = [int(v) for v in r["verdict"]]
            case["nodes"] = [int(v) for v in r["nodes"]]
            case["winner"] = int(r["winner"])
            # every SAT subproblem's full model + decoded factors, from the reference itself
            sat = {}
            for k, v in enumerate(r["verdict"]):
                if v == 1:
                    m = refp.dfs(fr[k][0])
                    full = np.concatenate([fr[k][1], m["model"]])
                    d1, d2, ok = ref.convert(full, hdr["v1"], hdr["v2"], hdr["product"])
                    sat[str(k)] = {"model": [int(x) for x in full], "fact1": str(d1), "fact2": str(d2),
                                   "verified": bool(ok)}
            case["sat"] = sat
            # clause_evals from the oracle, only after its node counts match the reference's
            global _FR
            _FR = fr
            with multiprocessing.get_context("fork").Pool(workers) as pool:
                res = pool.map(_oracle_dfs, range(len(fr)), chunksize=4)
            for k, (nodes_k, models_k, ev_k) in enumerate(res):
                assert nodes_k == case["nodes"][k], (cnf, mode, value, k)
                assert (models_k > 0) == (case["verdict"][k] == 1)
            case["clause_evals"] = [int(r[2]) for r in res]
        refp.frontier_free(h)
        out["cases"].append(case)
        print("%-14s -%s %-6d size %-6d it %-8d tasks %-8d %s  %.1fs" % (
            cnf, mode, value, len(fr), info["iterations"], info["task_count"],
            "dfs nodes=%d" % sum(case["nodes"]) if "nodes" in case else "", time.time() - t0), flush=True)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
