#!/usr/bin/env python3
"""Ad-hoc timing probe (GPU box): BFS + exhaustive DFS on a few named workloads.
    python tools/perf_probe.py rsaFACT-32bit:q:1024 prime-32bit:q:1024 rsaFACT-32bit:d:12:e
A trailing :e runs DFS with early exit."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ndp-5_b200/py", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import cbind  # noqa: E402
import gen_cnf  # noqa: E402
import ndp5_b200 as nd  # noqa: E402

orc = cbind.OracleLib()
for spec in sys.argv[1:]:
    parts = spec.split(":")
    name, mode, value = parts[0], parts[1], int(parts[2])
    early = len(parts) > 3 and parts[3] == "e"
    A = orc.parse(gen_cnf.dimacs_text(*gen_cnf.NAMED[name]))
    D, ovr, Q = (value, False, -1) if mode == "d" else (0, True, value)
    for rep in range(2):
        t0 = time.time()
        fr = nd.bfs(A, D, ovr, Q)
        t1 = time.time()
        r = fr.dfs_shard(early_exit=early, counters=True)
        t2 = time.time()
        bs = fr.bfs_stats()
        st = r["stats"]
        n = len(fr)
        print("%s rep%d: frontier %d (n0=%s) | BFS wall %.3fs kernel %.1fms launches %d evals %.3g | DFS wall %.3fs "
              "kernel %.2fms nodes %d evals %.4g rebuilds %d -> %.1f subproblems/s, %.1f Gce/s (%.2f of HBM roofline) "
              "winner %s" % (spec, rep, n, fr.dims(0)[0] if n else None, t1 - t0, bs["kernel_ms"], bs["launches"],
                             bs["clause_evals"], t2 - t1, st["kernel_ms"], st["nodes"], st["clause_evals"],
                             st["rebuilds"], n / (st["kernel_ms"] / 1e3 + 1e-12),
                             st["clause_evals"] / (st["kernel_ms"] / 1e3 + 1e-12) / 1e9,
                             st["clause_evals"] * 12 / (st["kernel_ms"] / 1e3 + 1e-12) / 6544.3e9, r["winner"]),
              flush=True)
        fr.free()
