#!/usr/bin/env python3
"""Repeat the DFS kernel on a resident frontier and print kernel times (GPU box).
    python tools/dfs_reps.py rsaFACT-32bit:q:16384 [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ndp-5_b200/py", "tests", "tools", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
import bench  # noqa: E402
import ndp5_b200 as nd  # noqa: E402

spec = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # emulate rank 0 of `stride` ranks
name, mode, value, (D, ovr, Q) = bench.parse_workload(spec)
A = bench.clause_array(name)[0]
fr = nd.bfs(A, D, ovr, Q)
ts = []
for _ in range(reps):
    r = fr.dfs_shard(first=0, stride=stride, early_exit=False, counters=False)
    ts.append(r["stats"]["kernel_ms"])
print(spec, "stride=%d" % stride, "n=%d" % len(fr), "kernel_ms", " ".join("%.3f" % t for t in ts), "nodes", r["stats"]["nodes"],
      "launches", r["stats"]["launches"])
