"""The N>1 host logic (static interleaved shards, verdict all-gather, lowest-winner reduction,
model broadcast) over gloo with world_size 2 and 3 on the CPU.  The per-shard verdicts come from
the oracle here (no GPU in this suite); on the GPU box the same functions run over NCCL in
bench.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "ndp-5_b200", "py")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ndp_shard", os.path.join(ROOT, "ndp-5_b200", "py", "ndp5_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    import cbind
    import gen_cnf
    orc = cbind.OracleLib()
    name, D, ovr, Q = case
    A = orc.parse(gen_cnf.dimacs_text(*gen_cnf.NAMED[name]))
    fr = orc.bfs(A, D, ovr, Q)["frontier"]          # every rank derives the same frontier (NDP:1651)
    n = len(fr)
    verdict = np.full(n, shard.NOT_RUN, dtype=np.uint8)
    winner, model = None, None
    for k in shard.shard_indices(n, rank, world):   # this rank's shard only
        r = orc.dfs(fr[k][0])
        verdict[k] = 1 if r["models"] else 0
        if r["models"] and winner is None:
            winner, model = int(k), np.concatenate([fr[k][1], r["model"]])
    merged = shard.gather_verdicts(dist, verdict, n, rank, world)
    gw = shard.lowest_winner(dist, winner, world)
    gm = shard.broadcast_model(dist, model if winner == gw else None, gw, world)
    total = shard.all_sum(dist, [int((verdict == 1).sum())], world)
    # the 64-byte early-exit word handles travel the same way (here: stand-in bytes; real handles need GPUs)
    mine = bytes([rank + 1]) * 64
    others = shard.exchange_exit_words(dist, mine, rank, world, device="cpu")
    assert others == [bytes([r + 1]) * 64 for r in range(world) if r != rank]
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), merged=merged, winner=-1 if gw is None else gw,
             model=np.zeros(0, np.int32) if gm is None else gm, n_sat=total[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", [("rsaFACT-16bit", 2240, False, -1), ("prime-16bit", 0, True, 64),
                                  ("rsaFACT-12bit", 0, True, 16)])
def test_sharded_run_equals_single_rank(world, case, tmp_path, oracle, cnf_text):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    name, D, ovr, Q = case
    fr = oracle.bfs(oracle.parse(cnf_text(name)), D, ovr, Q)["frontier"]
    want, first = [], None
    for k, (cl, ch) in enumerate(fr):
        r = oracle.dfs(cl)
        want.append(1 if r["models"] else 0)
        if r["models"] and first is None:
            first = (k, np.concatenate([ch, r["model"]]).tolist())
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert z["merged"].tolist() == want
        assert int(z["n_sat"]) == sum(want)
        if first is None:
            assert int(z["winner"]) == -1 and z["model"].size == 0
        else:
            assert int(z["winner"]) == first[0] and z["model"].tolist() == first[1]


class _OracleRun:
    """What ndp5_b200.PieceRun is on the GPU box, restated with the oracle for a frontier that is not cut (the
    pieces are the subproblems): this rank searched pieces rank, rank + world, ..."""

    def __init__(self, orc, fr, rank, world):
        self.n, self.rank, self.world = len(fr), rank, world
        self.first_model = np.full(self.n, 0xffffffff, dtype=np.uint32)
        self.res = {}
        for k in range(rank, self.n, world):
            r = orc.dfs(fr[k][0])
            self.res[k] = (r, np.concatenate([fr[k][1], r["model"]]) if r["models"] else None)
            if r["models"]:
                self.first_model[k] = k

    def fold(self, fm):
        verdict = np.zeros(self.n, np.uint8)
        nodes = np.zeros(self.n, np.uint64)
        evals = np.zeros(self.n, np.uint64)
        sat = np.nonzero(np.asarray(fm) != 0xffffffff)[0]
        verdict[sat] = 1
        for k, (r, _) in self.res.items():
            nodes[k], evals[k] = r["nodes"], r["clause_evals"]
        winner = int(sat[0]) if len(sat) else None
        model = np.zeros(0, np.int32)
        if winner is not None and int(fm[winner]) % self.world == self.rank:
            model = self.res[winner][1]
        return dict(verdict=verdict, nodes=nodes, evals=evals, winner=winner, model=model,
                    stats=dict(nodes=int(nodes.sum()), clause_evals=int(evals.sum())))


def _piece_worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ndp_shard", os.path.join(ROOT, "ndp-5_b200", "py", "ndp5_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    import cbind
    import gen_cnf
    orc = cbind.OracleLib()
    name, D, ovr, Q = case
    fr = orc.bfs(orc.parse(gen_cnf.dimacs_text(*gen_cnf.NAMED[name])), D, ovr, Q)["frontier"]
    out = shard.fold_piece_runs(dist, _OracleRun(orc, fr, rank, world), world)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), verdict=out["verdict"], nodes=out["nodes"], evals=out["evals"],
             winner=-1 if out["winner"] is None else out["winner"],
             model=np.zeros(0, np.int32) if out["winner"] is None else out["model"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", [("rsaFACT-16bit", 969, False, -1), ("prime-12bit", 0, True, 16)])
def test_piece_fold_collectives(world, case, tmp_path, oracle, cnf_text):
    """shard.fold_piece_runs (MIN of first_model, MAX of verdicts, SUM of counters, model from the rank that searched
    the winning piece) over gloo: every rank ends with the single-rank verdicts, counters, winner and model."""
    port = _free_port()
    mp.spawn(_piece_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    name, D, ovr, Q = case
    fr = oracle.bfs(oracle.parse(cnf_text(name)), D, ovr, Q)["frontier"]
    want_v, want_n, want_e, first = [], [], [], None
    for k, (cl, ch) in enumerate(fr):
        r = oracle.dfs(cl)
        want_v.append(1 if r["models"] else 0)
        want_n.append(r["nodes"])
        want_e.append(r["clause_evals"])
        if r["models"] and first is None:
            first = (k, np.concatenate([ch, r["model"]]).tolist())
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert z["verdict"].tolist() == want_v and z["nodes"].tolist() == want_n and z["evals"].tolist() == want_e
        if first is None:
            assert int(z["winner"]) == -1
        else:
            assert int(z["winner"]) == first[0] and z["model"].tolist() == first[1]


def test_shard_indices_partition():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ndp_shard", os.path.join(ROOT, "ndp-5_b200", "py", "ndp5_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            allidx = np.concatenate([shard.shard_indices(n, r, world) for r in range(world)])
            assert sorted(allidx.tolist()) == list(range(n))
    assert shard.gather_verdicts(None, [0, 1, 2], 3, 0, 1).tolist() == [0, 1, 2]
    assert shard.lowest_winner(None, 5, 1) == 5 and shard.lowest_winner(None, None, 1) is None
    assert shard.exchange_exit_words(None, b"x" * 64, 0, 1) == []
