"""Host-side logic of the multi-rank run (one process per GPU): the static interleaved split of
the frontier and the two tiny exchanges the north star allows -- per-shard verdict bytes and the
SAT-found flag / winner.  Works over any torch.distributed backend (NCCL on the GPU box, gloo in
the CPU tests); with dist=None it is the single-rank identity."""
import numpy as np

NOT_RUN = 2
_ABSENT = 255


def shard_indices(n, rank, world):
    """Global subproblem indices of one rank: k -> k mod world (ndp_dfs_shard(first=rank, stride=world))."""
    return np.arange(rank, n, world)


def gather_verdicts(dist, verdict, n, rank, world, device="cpu"):
    """All-gather the verdict bytes: every rank contributes the entries of its own shard."""
    verdict = np.asarray(verdict, dtype=np.uint8)
    if dist is None or world == 1:
        return verdict.copy()
    import torch
    mine = np.full(n, _ABSENT, dtype=np.uint8)
    idx = shard_indices(n, rank, world)
    mine[idx] = verdict[idx]
    t = torch.from_numpy(mine).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    merged = torch.stack(outs).min(dim=0).values.cpu().numpy()
    assert (merged != _ABSENT).all(), "a shard did not report"
    return merged


def lowest_winner(dist, winner, world, device="cpu"):
    """The deterministic winner: the lowest SAT index over all shards (None if none).  This is also
    the early-exit word exchanged between kernel waves (a MIN all-reduce of one int64)."""
    if dist is None or world == 1:
        return winner
    import torch
    big = np.iinfo(np.int64).max
    t = torch.tensor([big if winner is None else int(winner)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    v = int(t.item())
    return None if v == big else v


def broadcast_model(dist, model, winner, world, device="cpu", cap=1 << 16):
    """The winner's model travels from the rank that owns it (winner mod world) to everyone."""
    if dist is None or world == 1 or winner is None:
        return None if winner is None else np.asarray(model, dtype=np.int32)
    import torch
    src = int(winner) % world
    buf = torch.zeros(cap + 1, dtype=torch.int32, device=device)
    if dist.get_rank() == src:
        m = np.asarray(model, dtype=np.int32)
        buf[0] = m.size
        buf[1:1 + m.size] = torch.from_numpy(m).to(device)
    dist.broadcast(buf, src=src)
    k = int(buf[0].item())
    return buf[1:1 + k].cpu().numpy()


def all_sum(dist, values, world, device="cpu"):
    if dist is None or world == 1:
        return [int(v) for v in values]
    import torch
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return [int(v) for v in t.cpu().tolist()]


def all_max(dist, value, world, device="cpu"):
    if dist is None or world == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def exchange_exit_words(dist, own_handle, rank, world, device="cuda"):
    """All-gather the ranks' 64-byte early-exit word handles (CUDA IPC); returns the OTHER ranks' handles in rank order.
    One tiny collective at set-up: the words themselves are then lowered by the kernels over NVLink, no further traffic."""
    if dist is None or world == 1:
        return []
    import torch
    t = torch.from_numpy(np.frombuffer(own_handle, dtype=np.uint8).copy()).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [o.cpu().numpy().tobytes() for r, o in enumerate(outs) if r != rank]


# ---- piece-level sharing (ndp_dfs_pieces_*): the two collectives between begin and the job-wide result ----
def min_first_model(dist, first_model, world, device="cpu"):
    """MIN all-reduce of the ranks' first_model vectors (lowest piece with a model per subproblem; 0xffffffff none)."""
    fm = np.asarray(first_model, dtype=np.uint32)
    if dist is None or world == 1:
        return fm.copy()
    import torch
    t = torch.from_numpy(fm.astype(np.int64)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t.cpu().numpy().astype(np.uint32)


def combine_partials(dist, part, world, device="cpu"):
    """Job-wide per-subproblem results from the ranks' ndp_dfs_pieces_fold partials: verdict MAX (UNSAT 0 < SAT 1 <
    NOT_RUN 2, SAT reported by every rank), nodes / evals SUM.  One all-reduce each (three vectors of frontier size)."""
    verdict = np.asarray(part["verdict"], dtype=np.uint8)
    nodes = np.asarray(part["nodes"], dtype=np.uint64)
    evals = np.asarray(part["evals"], dtype=np.uint64)
    if dist is None or world == 1:
        return verdict.copy(), nodes.copy(), evals.copy()
    import torch
    v = torch.from_numpy(verdict.astype(np.int64)).to(device)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    c = torch.from_numpy(np.stack([nodes.astype(np.int64), evals.astype(np.int64)])).to(device)
    dist.all_reduce(c)
    c = c.cpu().numpy()
    return v.cpu().numpy().astype(np.uint8), c[0].astype(np.uint64), c[1].astype(np.uint64)


def broadcast_model_from(dist, model, src, world, device="cpu", cap=1 << 16):
    """The model travels from rank `src` (the one that searched the winning piece) to everyone."""
    if dist is None or world == 1:
        return np.asarray(model, dtype=np.int32)
    import torch
    buf = torch.zeros(cap + 1, dtype=torch.int32, device=device)
    if dist.get_rank() == src:
        m = np.asarray(model, dtype=np.int32)
        buf[0] = m.size
        buf[1:1 + m.size] = torch.from_numpy(m).to(device)
    dist.broadcast(buf, src=src)
    k = int(buf[0].item())
    return buf[1:1 + k].cpu().numpy()


def fold_piece_runs(dist, run, world, device="cpu", want_model=True):
    """begin -> [MIN first_model] -> fold -> [MAX verdict, SUM nodes/evals] -> winner + model on every rank.
    `run` is this rank's ndp5_b200.PieceRun.  Returns dict(verdict, nodes, evals, winner, model, stats)."""
    fm = min_first_model(dist, run.first_model, world, device)
    part = run.fold(fm)
    verdict, nodes, evals = combine_partials(dist, part, world, device)
    winner, model = part["winner"], part["model"]
    if winner is not None and want_model:
        src = int(fm[winner]) % world          # pieces are dealt out piece mod world
        model = broadcast_model_from(dist, model, src, world, device)
    return dict(verdict=verdict, nodes=nodes, evals=evals, winner=winner, model=model, stats=part["stats"])
