#!/usr/bin/env python3
"""One short DFS launch on a large CNF for `ncu --set full` (a full 64-bit launch runs for minutes; ncu replays a
kernel ~40 times).  NDP_PROFILE_MAX_UNITS caps the launch at the first K pieces; everything else is the product path.

    NDP_PROFILE_MAX_STEPS=20000 python tools/ncu_probe.py rsaFACT-64bit 4096     (every warp stops after 20000 unit steps)
    NDP_PROFILE_MAX_UNITS=20000 python tools/ncu_probe.py rsaFACT-64bit 4096     (only the first 20000 pieces are searched)
Prints the nodes the DFS kernel visited in the capped launch (divide the ncu instruction count by `kernel_nodes`)."""
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
    dump = "/tmp/ncu_probe_pieces.csv"
    os.environ["NDP_DUMP_PIECES"] = dump
    for _ in range(2):      # the first launch warms the pools and the code; profile the second (ncu -s 1 -c 1)
        r = fr.dfs_shard(early_exit=False, counters=True)
        kernel_nodes = 0
        with open(dump) as f:            # nodes the DFS kernel itself visited (the cut's interior nodes are not its work)
            for line in f:
                kernel_nodes += int(line.rsplit(",", 1)[1])
        out = {"cnf": name, "q": int(q), "frontier": len(fr), "units_cap": os.environ.get("NDP_PROFILE_MAX_UNITS"),
               "steps_cap": os.environ.get("NDP_PROFILE_MAX_STEPS"), "kernel_nodes": kernel_nodes,
               "nodes_incl_cut": int(r["stats"]["nodes"]), "kernel_ms_incl_cut": r["stats"]["kernel_ms"],
               "launches": int(r["stats"]["launches"])}
    print(json.dumps(out))
    fr.free()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
