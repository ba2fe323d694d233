#!/usr/bin/env python3
"""Multi-GPU check (GPU box, >= 2 devices): ndp_dfs_batch over G devices must return the single-GPU
verdict vector, winner and model -- with the piece split active on every shard.
    python tools/check_batch.py G"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ndp-5_b200/py", "tests", "tools", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import ndp5_b200 as nd  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
assert nd.device_count() >= G, "needs %d devices" % G
ok = True
for spec, early in [("rsaFACT-32bit:q:1024", False), ("rsaFACT-32bit:d:12", True), ("prime-32bit:q:1024", False),
                    ("rsaFACT-32bit:q:16384", False), ("rsaFACT-48bit:q:4096", True)]:
    name, mode, value, (D, ovr, Q) = bench.parse_workload(spec)
    A = bench.clause_array(name)[0]
    fr = nd.bfs(A, D, ovr, Q)
    one = fr.dfs_shard(early_exit=early)
    for rep in range(2):
        t0 = time.time()
        many = fr.dfs_batch(n_gpus=G, early_exit=early)
        dt = time.time() - t0
    w = one["winner"]
    same_v = (many["verdict"] == one["verdict"]).all() if not early else \
        (many["verdict"][: (w + 1 if w is not None else None)] == one["verdict"][: (w + 1 if w is not None else None)]).all()
    good = bool(same_v) and many["winner"] == w and many["model"].tolist() == one["model"].tolist()
    ok &= good
    print("%s early=%d G=%d: %s  winner %s  1-GPU kernel %.2f ms | %d-GPU kernel %.2f ms, wall %.1f ms, nodes %d vs %d" % (
        spec, early, G, "OK" if good else "MISMATCH", many["winner"], one["stats"]["kernel_ms"], G,
        many["stats"]["kernel_ms"], dt * 1e3, many["stats"]["nodes"], one["stats"]["nodes"]), flush=True)
    fr.free()
sys.exit(0 if ok else 1)
