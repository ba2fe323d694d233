#!/usr/bin/env python3
"""One short DFS launch on a large CNF for `ncu --set full` (a full 64-bit launch runs for minutes; ncu replays a
kernel ~40 times).  NDP_PROFILE_MAX_UNITS caps the launch at the first K pieces; everything else is the product path.

    NDP_PROFILE_MAX_UNITS=20000 python tools/ncu_probe.py rsaFACT-64bit 4096
Prints nodes, clause evals and kernel ms of the capped launch (divide the ncu instruction count by `nodes`)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ndp-5_b200", "py"))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import ndp5_b200 as nd  # noqa: E402


def main(name, q):
    A = bench.clause_array(name)[0]
    fr = nd.bfs(A, 0, True, int(q))
    out = None
    for _ in range(2):      # the first launch warms the pools and the code; profile the second (ncu -s 1 -c 1)
        r = fr.dfs_shard(early_exit=False, counters=True)
        out = {"cnf": name, "q": int(q), "frontier": len(fr), "units_cap": os.environ.get("NDP_PROFILE_MAX_UNITS"),
               "nodes": int(r["stats"]["nodes"]), "clause_evals": int(r["stats"]["clause_evals"]),
               "kernel_ms": r["stats"]["kernel_ms"], "launches": int(r["stats"]["launches"])}
    print(json.dumps(out))
    fr.free()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
