#!/usr/bin/env python3
"""Wall-clock breakdown of one e2e step (GPU box):  python tools/e2e_trace.py rsaFACT-32bit:q:16384"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ndp-5_b200/py", "tests", "tools", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
import bench  # noqa: E402
import ndp5_b200 as nd  # noqa: E402
spec = sys.argv[1]
name, mode, value, (D, ovr, Q) = bench.parse_workload(spec)
A = bench.clause_array(name)[0]
for rep in range(4):
    t0 = time.time(); f = nd.bfs(A, D, ovr, Q)
    t1 = time.time(); r = f.dfs_shard(early_exit=False, counters=False)
    t2 = time.time(); f.free()
    t3 = time.time()
    print("%s rep%d: bfs %.2f ms | dfs %.2f ms | free %.2f ms | total %.2f ms (kernels %.2f + %.2f)" % (
        spec, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0), f.h is None or 0,
        r["stats"]["kernel_ms"]), flush=True)
