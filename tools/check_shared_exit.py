#!/usr/bin/env python3
"""Cross-process early exit (GPU box, torchrun with >= 2 ranks): every rank granulates, searches its interleaved shard
with device-visible exit words (ndp_dfs_shard_shared) and the job's winner / model must be the single-GPU ones; ranks
whose shard lies above the winner abandon their work.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_shared_exit.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ndp-5_b200/py", "tests", "tools", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
import ndp5_b200 as nd  # noqa: E402
from ndp5_b200 import shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
nd.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
own = nd.ExitWord()
peers = [nd.ExitWord.from_handle(h) for h in shard.exchange_exit_words(dist, own.export(), rank, world)]
ok = True
for spec, expect in [("rsaFACT-32bit:q:16384", 409), ("rsaFACT-48bit:q:4096", 301), ("rsaFACT-32bit:q:1024", 25),
                     ("prime-32bit:q:1024", None)]:
    name, mode, value, (D, ovr, Q) = bench.parse_workload(spec)
    A = bench.clause_array(name)[0]
    fr = nd.bfs(A, D, ovr, Q)
    for rep in range(2):
        own.reset()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.time()
        r = fr.dfs_shard_shared(own, peers, first=rank, stride=world, early_exit=True)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.time() - t0
    gw = shard.lowest_winner(dist, r["winner"], world, device="cuda")
    gm = shard.broadcast_model(dist, r["model"] if r["winner"] == gw else None, gw, world, device="cuda")
    seen = own.read()
    good = gw == expect and seen == expect
    if expect is not None:
        one = fr.dfs_shard(early_exit=True) if rank == 0 else None
        if rank == 0:
            good = good and gm.tolist() == one["model"].tolist()
    notrun = int((r["verdict"][rank::world] == 2).sum())
    ok = ok and good
    print("rank %d %s: job winner %s (expected %s), word %s, my winner %s, abandoned %d of %d, nodes %d, %.1f ms %s" % (
        rank, spec, gw, expect, seen, r["winner"], notrun, len(r["verdict"][rank::world]), r["stats"]["nodes"], dt * 1e3,
        "OK" if good else "MISMATCH"), flush=True)
    fr.free()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
