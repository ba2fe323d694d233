#!/usr/bin/env python3
"""bench.py -- the NDP hot path (BFS granulation -> DFS over the frontier) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME:q|d:VALUE]
    torchrun ... bench.py --gpus N ...            (one rank per GPU; RANK/LOCAL_RANK/WORLD_SIZE from env)

Workload (config.workload, the same string in both arms): the rsaFACT-48bit factoring CNF of BASELINE.json
configs[2] (13104 clauses / 3312 variables, tools/gen_cnf.py) granulated with -q 4096 and searched
EXHAUSTIVELY (no early exit): 4096 subproblems, 3.0e9 DFS nodes, ~0.6 s of GPU work -- enough for an 8-GPU
step to be many times the longest single piece, while the reference arm can still MEASURE its BFS (56 s; the
verbatim -q 65536 of configs[2] needs a 1000 s CPU BFS).  The round-1 headline (32-bit -q 16384, 2 ms of work)
is kept as the extra key "legacy_32bit_q16384"; configs[1] verbatim (-d 12: ONE subproblem) as "tts_d12".

A step = one pass of the hot path: ndp_bfs(host clause array) -> piece-level DFS (ndp_dfs_pieces_begin/fold)
-> per-subproblem verdicts on the host.  Every step is run once and read two ways:
  value : subproblems resolved per second of DEVICE time of that step (CUDA events inside the library on the
          launching stream: BFS level loop + DFS cut + DFS kernel), max over ranks.  Both arms define value over
          BFS + DFS.  The frontier is shared out piece by piece (piece p -> rank p mod N); total work is fixed =>
          "scaling": "strong".
  e2e   : the same step by the host clock: host clause array in, H2D/D2H copies, launches, the two small
          collectives (MIN of first-model pieces, MAX/SUM of verdicts/counters over NCCL at N > 1), verdict
          vector on the host; barrier + synchronize on both sides, max over ranks.
  --impl reference : the reference's own CPU code (oracle/_ref = the unmodified NDP-5_6_7.cpp behind
          oracle/ref_harness.cpp; the plain-C oracle port if that is absent): BFS once on one core, as the
          reference does on every rank (NDP:1651); each step = DFS over a DIFFERENT interleaved subset of the
          frontier (subproblem sizes are heavy-tailed: never a prefix) with nproc-1 worker processes pulling
          indices FIFO (NDP:1322-1346), scaled to the whole frontier.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "ndp-5_b200", "py"), os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

DEFAULT_WORKLOAD = "rsaFACT-48bit:q:4096"
LEGACY_WORKLOAD = "rsaFACT-32bit:q:16384"
ALGO_BYTES_PER_EVAL = 12  # sizeof(Clause3), ClauseSetPool.hpp:37-39
SMSP_PER_GPU = 148 * 4    # warp schedulers of a B200: one warp instruction per clock each
# Known answers checked on every run.  BFS counters: the reference's (tests/golden).  n_sat /
# first_sat: reference-pinned where tests/golden/big_*.npz covers them (tools/make_golden_big.py) -- both
# 32-bit -q 16384 frontiers exhaustively, 48-bit -q 65536 and -q 4096 by the prefix 0..first_sat; the 64-bit
# entry is this implementation's own (its model is GMP-verified, the UNSAT verdicts below it are not pinned).
KAT = {
    "rsaFACT-32bit:q:16384": dict(size=16384, iterations=1761168, task_count=1777552, n_sat=2, first_sat=409),
    "rsaFACT-32bit:q:1024": dict(size=1024, iterations=69520, task_count=70544, n_sat=2, first_sat=25),
    "prime-32bit:q:1024": dict(size=1024, iterations=69520, task_count=70544, n_sat=0, first_sat=None),
    "prime-32bit:q:16384": dict(size=16384, iterations=1761168, task_count=1777552, n_sat=0, first_sat=None),
    "rsaFACT-32bit:d:12": dict(size=1, iterations=12, task_count=13, n_sat=1, first_sat=0),
    "rsaFACT-48bit:q:65536": dict(size=65536, iterations=7862688, task_count=7928224, n_sat=None, first_sat=4822),
    "rsaFACT-48bit:q:4096": dict(size=4096, iterations=395680, task_count=399776, n_sat=None, first_sat=301),
    "rsaFACT-64bit:q:4096": dict(size=4096, iterations=395696, task_count=399792, n_sat=None, first_sat=666),
}


def workload_label(name, mode, value):
    """config.workload: identical in both arms so the driver can match them."""
    return "%s -%s %d, exhaustive DFS (no early exit)" % (name, mode, value)


def clause_array(name):
    """The reference's in-memory clause array for a named CNF (unit x -> {0,0,x}, NDP:632-637),
    built straight from the generator -- no oracle code involved."""
    import gen_cnf
    N, bits = gen_cnf.NAMED[name]
    nvars, clauses, A, B = gen_cnf.multiplier_cnf(N, bits)
    out = np.zeros((len(clauses), 3), dtype=np.int32)
    for k, c in enumerate(clauses):
        if len(c) == 1:
            out[k, 2] = c[0]
        else:
            out[k] = c
    return out, nvars, N, A[::-1], B[::-1]


def parse_workload(w):
    name, mode, value = w.split(":")
    value = int(value)
    return name, mode, value, ((value, False, -1) if mode == "d" else (0, True, value))


def golden_big(workload):
    """Reference-produced per-subproblem verdicts / node counts / clause evals for (a subset of) this frontier,
    if tools/make_golden_big.py committed them."""
    name, mode, value = workload.split(":")
    out = []
    for suffix in ("", "_prefix", "_sample"):   # _sample (tools/make_golden_64bit.py): clause evals only where the port ran too (else 0)
        path = os.path.join(ROOT, "tests", "golden", "big_%s_%s%s%s.npz" % (name, mode, value, suffix))
        if os.path.exists(path):
            out.append(np.load(path))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def recorded_profile(workload):
    """What the committed ncu --set full capture of the DFS kernel says for this workload (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload, {})
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------ reference (CPU)
def interleaved(n, k, offset=0):
    """k indices spread over [0, n): offset, offset + s, ... with s = ceil(n / k) -- never a prefix."""
    s = max(1, -(-n // max(k, 1)))
    return list(range(offset % s, n, s))


def cpu_dfs_sample(get_item, n, budget_s, want_cores=None):
    """Time the reference CPU DFS (oracle/_ref; else the oracle port) over an interleaved sample of the
    frontier (get_item(k) -> (clauses, choices)) with nproc-1 worker processes.  Returns (cpu_baseline dict,
    sample indices, reference verdicts of the sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cbind
    cores = want_cores or max(1, (os.cpu_count() or 2) - 1)
    if cbind.RefLib.available():
        ref = cbind.RefLib()
        # one subproblem tells the rate; then one subproblem per worker (or more if they are small)
        probe = [n // 2]
        h = ref.frontier_from_arrays([get_item(k) for k in probe])
        r = ref.dfs_frontier(h, 0, 1, 1, early_exit=False)
        ref.frontier_free(h)
        per = max(r["seconds"], 1e-6)
        k = int(max(4, min(n, cores * budget_s / per)))
        idx = interleaved(n, min(k, n))
        h = ref.frontier_from_arrays([get_item(i) for i in idx])
        r = ref.dfs_frontier(h, 0, len(idx), cores, early_exit=False)
        ref.frontier_free(h)
        busy = float(ref.last_busy_seconds(len(idx)).sum())
        base = {"value": len(idx) * cores / max(busy, 1e-9), "unit": "subproblems/s", "cores": cores, "kind": "reference",
                "sample": "DFS phase only (BFS excluded): %d of %d subproblems of the same frontier, every %d-th "
                          "(interleaved, not a prefix), unmodified Satisfy_iterative via oracle/_ref, %d forked workers, "
                          "%.2f s of wall clock, %.1f worker seconds; value = subproblems per worker second x %d cores"
                          % (len(idx), n, max(1, -(-n // len(idx))), cores, r["seconds"], busy, cores),
                "n_sat_in_sample": int((r["verdict"] == 1).sum())}
        return base, idx, r["verdict"]
    orc = cbind.OracleLib()
    idx = interleaved(n, 8)
    t0 = time.time()
    verdict = []
    for i in idx:
        verdict.append(1 if orc.dfs(get_item(i)[0])["models"] else 0)
        if time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    idx = idx[:len(verdict)]
    base = {"value": len(idx) / dt, "unit": "subproblems/s", "cores": 1, "kind": "port",
            "sample": "DFS phase only: %d of %d subproblems (interleaved), plain-C oracle port, 1 thread, %.2f s"
                      % (len(idx), n, dt), "n_sat_in_sample": int(sum(verdict))}
    return base, idx, np.asarray(verdict, dtype=np.uint8)


def run_reference(args, workload):
    """--impl reference: the reference's own CPU code on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cbind
    name, mode, value, (D, ovr, Q) = parse_workload(workload)
    A, nvars, N, v1, v2 = clause_array(name)
    cores = max(1, (os.cpu_count() or 2) - 1)
    line = {"impl": "reference", "metric": "subproblems resolved/s", "unit": "subproblems/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_label(name, mode, value),
                       "timed_region": "BFS (1 core, measured once) + DFS (host worker processes), CPU"}}
    if not cbind.RefLib.available():
        # the oracle port stands in when the compiled reference did not travel
        orc = cbind.OracleLib()
        t0 = time.time()
        h, info = orc.bfs_handle(A, D, ovr, Q)
        bfs_s = time.time() - t0
        n = orc._fsize(h)

        def get_item(k):
            ncl, nch = orc._fncl(h, k), orc._fnch(h, k)
            cl = np.zeros((ncl, 3), dtype=np.int32)
            ch = np.zeros(nch, dtype=np.int32)
            orc._fcopy(h, k, cl.ctypes.data_as(cbind._i32p), ch.ctypes.data_as(cbind._i32p))
            return cl, ch
        base, idx, _ = cpu_dfs_sample(get_item, n, 20.0)
        orc._ffree(h)
        v = n / (bfs_s + n / base["value"])
        line.update(value=v, ms_per_step=1e3 * n / v, cpu_baseline=dict(base, value=v),
                    e2e={"value": v, "unit": "subproblems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return 0
    ref = cbind.RefLib()
    h, info = ref.bfs_handle(A, D, ovr, Q)          # single-threaded, as on every reference rank (NDP:1651)
    bfs_s = info["seconds"]
    n = ref._fsize(h)

    def element(k):
        ncl, nch = ref._fncl(h, k), ref._fnch(h, k)
        cl = np.zeros((ncl, 3), dtype=np.int32)
        ch = np.zeros(nch, dtype=np.int32)
        ref._fcopy(h, k, cl.ctypes.data_as(cbind._i32p), ch.ctypes.data_as(cbind._i32p))
        return cl, ch

    # Bound the run: ~3 minutes of DFS for warmup + steps together.  Step i owns the interleaved subset
    # {i, i + s, i + 2s, ...}[:m] of the frontier (subproblem sizes are heavy-tailed: never a prefix; no two steps
    # share a subproblem).  All subsets go through ONE FIFO queue served by nproc-1 forked workers -- the
    # reference's one-task-per-request dispatch (NDP:1322-1346) with every core busy -- and each subproblem is
    # timed by the worker that searches it.  A step's DFS time is its subset's worker seconds scaled to the whole
    # frontier on `cores` workers: busy(subset) x n / (m x cores).
    total_steps = args.warmup + args.steps
    hp = ref.frontier_from_arrays([element(n // 2)])
    per = max(ref.dfs_frontier(hp, 0, 1, 1, early_exit=False)["seconds"], 1e-6)   # one subproblem on one core
    ref.frontier_free(hp)
    budget = 180.0
    m = int(max(1, min(n // max(total_steps, 1), cores * budget / total_steps / per)))
    s = max(total_steps, n // m)
    order = []                                  # warmup subsets first, then the timed ones, in FIFO order
    for i in range(total_steps):
        order.extend(list(range(i, n, s))[:m])
    hs = ref.frontier_from_arrays([element(j) for j in order])
    r = ref.dfs_frontier(hs, 0, len(order), cores, early_exit=False)
    busy = ref.last_busy_seconds(len(order))
    ref.frontier_free(hs)
    ref.frontier_free(h)
    dts, sat, sampled = [], 0, 0
    for i in range(args.warmup, total_steps):
        sl = slice(i * m, (i + 1) * m)
        dts.append(float(busy[sl].sum()) * n / (m * cores))
        sat += int((r["verdict"][sl] == 1).sum())
        sampled += m
    dfs_s_full = sum(dts) / len(dts)
    v = n / (bfs_s + dfs_s_full)
    sample = ("BFS %.2f s on 1 core (measured once; the reference runs it single-threaded on every rank) + DFS: step i "
              "searches the interleaved subset {i, i + %d, ...} of %d subproblems (%d of %d sampled over the %d timed steps, "
              "no prefix, no repeats); all subsets served FIFO by %d forked workers in %.1f s of wall clock, each subproblem "
              "timed by its worker; step DFS time = its worker seconds x %d / (%d x %d cores) = mean %.1f s for the whole "
              "frontier; unmodified NDP-5_6_7.cpp via oracle/_ref"
              % (bfs_s, s, m, sampled, n, len(dts), cores, r["seconds"], n, m, cores, dfs_s_full))
    line.update(value=v, ms_per_step=1e3 * (bfs_s + dfs_s_full),
                cpu_baseline={"value": v, "unit": "subproblems/s", "cores": cores, "kind": "reference",
                              "sample": sample, "bfs_s": bfs_s, "dfs_s": dfs_s_full, "n_sat_in_sample": sat,
                              "dfs_s_per_step": dts},
                e2e={"value": v, "unit": "subproblems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    # BASELINE configs[1] verbatim (-d 12 = one subproblem = one worker) and the whole unmodified program, once
    if not args.no_tts and args.gpus == 1:
        A32 = clause_array("rsaFACT-32bit")[0]
        h12, info12 = ref.bfs_handle(A32, 12, False, -1)
        r12 = ref.dfs_frontier(h12, 0, 1, 1, early_exit=True)
        ref.frontier_free(h12)
        line["tts_d12"] = {"workload": "rsaFACT-32bit -d 12: 1 subproblem, unmodified Satisfy_iterative to its first model, "
                                       "1 worker", "seconds": info12["seconds"] + r12["seconds"],
                           "bfs_seconds": info12["seconds"], "dfs_seconds": r12["seconds"]}
        whole = whole_program_reference("rsaFACT-32bit", "q", 16384, cores + 1)
        if whole:
            line["whole_program"] = whole
    print(json.dumps(line))
    return 0


def whole_program_reference(name, mode, value, np_ranks):
    """The WHOLE unmodified reference program (main + process_queue dispatcher + workers) as np forked ranks over
    oracle/shim_fork/mpi.h -- `mpirun -np N ./NDP-5_6_7 file -q Q` (README.md:188) without OpenMPI.  It stops at the
    first model (NDP:1436-1447), so this is a time-to-solution, not the exhaustive metric above."""
    import re
    import tempfile
    import gen_cnf
    exe = os.path.join(ROOT, "oracle", "_ref", "ndp_ref_mpirun")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, name + ".dimacs")
        gen_cnf.write_named(name, path)
        try:
            out = subprocess.run([exe, "-np", str(np_ranks), path, "-" + mode, str(value)], cwd=td, stdout=subprocess.PIPE,
                                 stderr=subprocess.DEVNULL, text=True, errors="replace", timeout=900).stdout
        except subprocess.TimeoutExpired:
            return {"error": "timeout after 900 s"}
    m = re.search(r"\[ref_mpirun\] np=(\d+) wall=([0-9.]+) s", out)
    if not m:
        return {"error": "no launcher summary", "tail": out[-300:]}
    bfs = re.search(r"BFS time: ([0-9.]+) seconds \(", out)
    dfs = re.search(r"DFS time: ([0-9.]+) seconds \(", out)
    facts = re.findall(r"FACT \d: (\d+)", out)
    return {"what": "unmodified NDP-5_6_7 program, %s ranks (1 dispatcher + %d workers) forked over oracle/shim_fork/mpi.h, "
                    "%s -%s %d, to its first model" % (m.group(1), int(m.group(1)) - 1, name, mode, value),
            "wall_seconds": float(m.group(2)), "bfs_seconds": float(bfs.group(1)) if bfs else None,
            "dfs_seconds": float(dfs.group(1)) if dfs else None, "factors": facts, "verified": "verified." in out,
            "prime": "Prime!" in out}


# ------------------------------------------------------------------------------------------ B200
def run_b200(args, workload):
    import torch
    import ndp5_b200 as nd
    from ndp5_b200 import shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NDP hot path has no CPU fallback")
    torch.cuda.set_device(local)
    nd.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    name, mode, wl_value, (D, ovr, Q) = parse_workload(workload)
    A, nvars, N, v1, v2 = clause_array(name)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    kat = KAT.get(workload)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        return shard.all_max(dist, x, world, device="cuda")

    def step(clauses, Dq, early_exit=False, own=None, peers=(), want_model=False):
        """One pass of the hot path through the C ABI from host buffers; returns the job-wide result on every rank."""
        f = nd.bfs(clauses, *Dq)
        run = f.dfs_pieces(rank=rank, world=world, early_exit=early_exit, own=own, peers=peers)
        out = shard.fold_piece_runs(dist, run, world, device="cuda", want_model=want_model)
        out["bfs"] = f.bfs_stats()
        out["n"] = len(f)
        out["iterations"], out["task_count"] = f.iterations, f.task_count
        run.free()
        f.free()
        return out

    def decode(model):
        m = set(int(x) for x in model)
        d1 = int("".join("1" if v in m else "0" for v in v1), 2)
        d2 = int("".join("1" if v in m else "0" for v in v2), 2)
        return d1, d2

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    warm_note = "%d untimed steps of the same workload" % args.warmup
    if args.warmup_legacy:
        # a step of minutes (64-bit exhaustive): warm the device, the pools and the code paths on the small legacy
        # workload instead of repeating the long step three times
        Aw = clause_array(parse_workload(LEGACY_WORKLOAD)[0])[0]
        for _ in range(args.warmup):
            flush.zero_()
            step(Aw, parse_workload(LEGACY_WORKLOAD)[3])
        warm_note = "%d untimed steps of %s (a step of this workload takes minutes)" % (args.warmup, LEGACY_WORKLOAD)
    else:
        for _ in range(args.warmup):
            flush.zero_()
            out = step(A, (D, ovr, Q))
    barrier()
    t0 = time.time()
    step_s, dev_ms, launches = [], [], 0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        ta = time.time()
        out = step(A, (D, ovr, Q), want_model=True)
        torch.cuda.synchronize()
        step_s.append(time.time() - ta)
        dev_ms.append(out["bfs"]["kernel_ms"] + out["stats"]["kernel_ms"])
        launches += int(out["bfs"]["launches"] + out["stats"]["launches"])
    barrier()
    t1 = time.time()
    n = out["n"]
    verdict, nodes_k, evals_k = out["verdict"], out["nodes"], out["evals"]
    nodes, evals = int(nodes_k.sum()), int(evals_k.sum())
    sat_idx = np.nonzero(verdict == 1)[0]
    if kat:
        assert (n, out["iterations"], out["task_count"]) == (kat["size"], kat["iterations"], kat["task_count"]), \
            "BFS disagrees with the reference's known answer"
        assert (verdict <= 1).all() and (kat["n_sat"] is None or len(sat_idx) == kat["n_sat"]), \
            "verdict vector disagrees with the reference"
        assert (int(sat_idx[0]) if len(sat_idx) else None) == kat["first_sat"]
    model_ok = None
    if out["winner"] is not None:
        assert out["winner"] == int(sat_idx[0])
        d1, d2 = decode(out["model"])
        model_ok = d1 * d2 == N
        assert model_ok, "recovered factors do not multiply to N"
    # reference-produced fixtures for this frontier (verdict, node count, clause evals per subproblem)
    golden_checked = 0
    for z in golden_big(workload):
        idx = z["index"]
        assert int(z["size"]) == n and int(z["iterations"]) == out["iterations"]
        assert (verdict[idx] == z["verdict"]).all(), "verdicts differ from the reference's (tests/golden/big_*.npz)"
        assert (nodes_k[idx].astype(np.int64) == z["nodes"]).all(), "DFS node counts differ from the reference's"
        have = z["evals"] != 0 if "ref_seconds" in z.files else np.ones(len(idx), dtype=bool)
        assert (evals_k[idx][have] == z["evals"][have]).all(), "clause-evaluation counts differ from the oracle's"
        golden_checked += len(idx)
    # every step is timed on its own; value = device time (CUDA events in the library), e2e = host clock.  The
    # headline e2e is the MEDIAN step (a wall-clock mean over a few steps is at the mercy of one scheduler
    # hiccup on a shared host), the mean over the whole bracketed region is reported beside it.
    ms_per_step = allmax(float(np.mean(dev_ms)))
    value = n / (ms_per_step / 1e3)
    e2e_s = allmax(float(np.median(step_s)))
    e2e_mean_s = allmax((t1 - t0) / args.steps)
    h2d = out["bfs"]["h2d_bytes"] + out["stats"]["h2d_bytes"]   # counted by the library from the copies it issues
    d2h = out["bfs"]["d2h_bytes"] + out["stats"]["d2h_bytes"]
    bfs_ms, dfs_ms = allmax(out["bfs"]["kernel_ms"]), allmax(out["stats"]["kernel_ms"])
    e2e = {"value": n / e2e_s, "unit": "subproblems/s", "ms_per_step": 1e3 * e2e_s, "stat": "median of %d steps" % args.steps,
           "ms_per_step_mean": 1e3 * e2e_mean_s, "ms_per_step_min": 1e3 * float(min(step_s)),
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "bfs_kernel_ms": bfs_ms, "bfs_launches": int(out["bfs"]["launches"]), "dfs_kernel_ms": dfs_ms}
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # ---- time to the first model with the reference's "first solution ends the job" (NDP:1297-1303, 1436-1447):
    # device-visible exit words, one per rank, exchanged once as CUDA IPC handles; host buffers -> verified winner
    tts_exit = None
    if not args.no_tts:
        own = nd.ExitWord()
        peers = [nd.ExitWord.from_handle(hh) for hh in shard.exchange_exit_words(dist, own.export(), rank, world)]
        best = None
        for it_ in range(3):
            own.reset()
            barrier()
            ta = time.time()
            oe = step(A, (D, ovr, Q), early_exit=True, own=own, peers=peers, want_model=True)
            torch.cuda.synchronize()
            dt_ = allmax(time.time() - ta)
            if kat and kat.get("first_sat") is not None:
                assert oe["winner"] == kat["first_sat"], "early-exit winner disagrees with the exhaustive run"
                assert decode(oe["model"])[0] * decode(oe["model"])[1] == N
            if it_ > 0 and (best is None or dt_ < best):
                best = dt_
        tts_exit = {"what": "BFS + DFS with early exit from host buffers to the job's winner and model on every rank "
                            "(device-visible exit words over NVLink at N > 1)", "seconds": best, "winner": oe["winner"]}
        for w_ in peers + [own]:
            w_.free()

    # ---- round-1 headline workload (2 ms of work in total), kept so round-over-round comparison survives
    legacy = None
    if not args.no_tts and workload != LEGACY_WORKLOAD:
        ln, lm, lv, lD = parse_workload(LEGACY_WORKLOAD)
        A32 = clause_array(ln)[0]
        ls, lb, ld = [], [], []
        for it_ in range(5):
            barrier()
            ta = time.time()
            ol = step(A32, lD)
            torch.cuda.synchronize()
            if it_ >= 2:
                ls.append(time.time() - ta)
                lb.append(ol["bfs"]["kernel_ms"])
                ld.append(ol["stats"]["kernel_ms"])
        lk = KAT[LEGACY_WORKLOAD]
        lsat = np.nonzero(ol["verdict"] == 1)[0]
        assert (ol["n"], ol["iterations"], ol["task_count"], len(lsat), int(lsat[0])) == \
            (lk["size"], lk["iterations"], lk["task_count"], lk["n_sat"], lk["first_sat"])
        for z in golden_big(LEGACY_WORKLOAD):
            assert (ol["verdict"][z["index"]] == z["verdict"]).all()
            assert (ol["nodes"][z["index"]].astype(np.int64) == z["nodes"]).all() and (ol["evals"][z["index"]] == z["evals"]).all()
        legacy = {"workload": workload_label(ln, lm, lv), "e2e_ms_per_step": 1e3 * allmax(float(np.median(ls))),
                  "bfs_kernel_ms": allmax(float(np.median(lb))), "dfs_kernel_ms": allmax(float(np.median(ld))),
                  "nodes": int(ol["nodes"].sum()), "clause_evals": int(ol["evals"].sum()),
                  "reference_checked": "every verdict, node and clause-eval count (tests/golden/big_rsaFACT-32bit_q16384.npz)"
                  if golden_big(LEGACY_WORKLOAD) else "BFS counters, n_sat, first_sat"}

    # ---- single-subproblem time-to-solution: BASELINE configs[1] verbatim (-d 12), rank 0 only
    tts = None
    if rank == 0 and world == 1 and not args.no_tts:
        A32 = clause_array("rsaFACT-32bit")[0]
        best = None
        for _ in range(4):   # first pass warms the pools; report the best of the next three
            torch.cuda.synchronize()
            ta = time.time()
            f = nd.bfs(A32, 12, False, -1)
            tm = time.time()
            rr = f.dfs_shard(early_exit=True, counters=True)
            tb = time.time()
            cur = {"workload": "rsaFACT-32bit -d 12 (BASELINE configs[1] verbatim): 1 subproblem, cut into DFS-order "
                               "pieces on the device, searched to the reference's first model",
                   "seconds": tb - ta, "bfs_seconds": tm - ta, "dfs_seconds": tb - tm,
                   "dfs_nodes": int(rr["nodes"][0]), "dfs_kernel_ms": rr["stats"]["kernel_ms"],
                   "launches": int(rr["stats"]["launches"]),
                   "reference_cpu": "see tts_d12 of `bench.py --impl reference` on the same box"}
            assert cur["dfs_nodes"] == 3713280, "node count differs from the reference's known answer"
            f.free()
            if _ > 0 and (best is None or cur["seconds"] < best["seconds"]):
                best = cur
        tts = best

    # ---- the single-process multi-GPU path (ndp_dfs_batch: what the C++ program's -g N uses) on the same
    # frontier, rank 0 only (every device is visible under torchrun); the other ranks wait at the barrier
    batch_check = None
    if world > 1 and not args.no_tts:
        barrier()
        if rank == 0:
            f = nd.bfs(A, D, ovr, Q)
            tb0 = time.time()
            rb = f.dfs_batch(n_gpus=world, early_exit=False)
            tb1 = time.time()
            equal = bool((rb["verdict"] == verdict).all() and rb["winner"] == out["winner"] and
                         rb["model"].tolist() == np.asarray(out["model"]).tolist() and
                         int(rb["stats"]["nodes"]) == nodes and int(rb["stats"]["clause_evals"]) == evals)
            batch_check = {"n_gpus": world, "ms": 1e3 * (tb1 - tb0), "kernel_ms": rb["stats"]["kernel_ms"], "equal": equal,
                           "what": "ndp_dfs_batch(n_gpus=N) from one process vs the one-process-per-GPU run: verdict "
                                   "vector, winner, model, node and clause-eval totals"}
            f.free()
            assert equal, "ndp_dfs_batch disagrees with the sharded run"
        barrier()

    if rank == 0:
        peak, peak_src = measured_peak()
        dfs_s = dfs_ms / 1e3
        achieved = evals * ALGO_BYTES_PER_EVAL / world / dfs_s / 1e9   # per GPU, DFS kernel (the dominant one)
        prof = recorded_profile(workload)
        prof_note = ""
        if not prof:   # no capture of this very workload: the one of the same kernel build (large CNFs: state slots in global memory)
            other = DEFAULT_WORKLOAD if int(A.shape[0]) > 8192 else LEGACY_WORKLOAD
            prof = recorded_profile(other)
            prof_note = " [capture of %s, same kernel build; no ncu capture of this workload]" % other
        ipn = prof.get("warp_instructions_per_node")
        # real DRAM bytes of one DFS launch on one GPU (ncu): the capture's own figure for this workload, else bytes per node x nodes
        traffic = None
        if not prof_note and prof.get("dram_bytes_per_dfs_launch") and world == 1:
            traffic = prof["dram_bytes_per_dfs_launch"]
        elif not prof_note and prof.get("dram_bytes_per_node"):
            traffic = prof["dram_bytes_per_node"] * nodes / world
        elif not prof_note and prof.get("dram_bytes_per_dfs_launch"):
            traffic = prof["dram_bytes_per_dfs_launch"] / world
        sm_hz = (clocks["sm_mhz"] or 1965.0) * 1e6 if clocks else 1965.0e6
        issue = None
        if ipn:
            ach = nodes * ipn / world / dfs_s                      # warp instructions per second per GPU
            issue = {"bound": "issue", "achieved": ach / 1e9, "peak": SMSP_PER_GPU * sm_hz / 1e9, "unit": "G warp-instr/s",
                     "frac": ach / (SMSP_PER_GPU * sm_hz), "warp_instr_per_node": ipn,
                     "source": "nodes/s measured live x warp instructions per node from the committed ncu capture (%s)%s; "
                               "peak = 592 warp schedulers x SM clock" % (prof.get("source", "profiles/traffic.json"), prof_note)}
        line = {
            "metric": "subproblems resolved/s", "value": value, "unit": "subproblems/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_label(name, mode, wl_value),
                       "frontier": "%d subproblems, BFS %d expansions; DFS shared out piece by piece, piece p -> rank p mod %d"
                                   % (n, out["iterations"], world),
                       "cnf": {"clauses": int(A.shape[0]), "vars": int(nvars), "N": str(N)},
                       "timed_region": "BFS level loop + DFS cut + DFS kernel, device time (CUDA events on the launching "
                                       "stream, inside the library), mean over steps, max over ranks",
                       "l2": "256 MB flush write before every step; the DFS reads > 126 MB of piece state slots per step",
                       "warmup": warm_note,
                       "parity": "BFS counters, n_sat, first_sat and (golden_checked) per-subproblem verdict / node / "
                                 "clause-eval counts asserted against reference-produced fixtures on every run"},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "traffic_frac": traffic / dfs_s / 1e9 / peak if traffic else None,
                         "kernel": "the DFS kernel (dfs_index_group_kernel: 4 searches per warp), %.3f ms per launch incl. the cut "
                                   "(CUDA events on the launching stream)" % dfs_ms,
                         "algorithmic": "%d clause evaluations per step x 12 B (SURVEY.md §8d), counted by the device and "
                                        "equal to the oracle's on every reference-checked subproblem; the occurrence index "
                                        "does not perform the reference's per-node rescans, so frac exceeds 1 by "
                                        "construction -- the bound the kernel is subject to is instruction issue, see "
                                        "`issue`" % evals,
                         "issue": issue},
            "dfs": {"nodes": nodes, "clause_evals": evals, "gnodes_per_s": nodes / dfs_s / 1e9,
                    "gce_per_s": evals / dfs_s / 1e9, "n_sat": int(len(sat_idx)),
                    "first_sat": int(sat_idx[0]) if len(sat_idx) else None, "factors_verified": model_ok,
                    "golden_checked": int(golden_checked)},
            "clocks": clocks,
        }
        if tts:
            line["tts_d12"] = tts
        if tts_exit:
            line["tts_early_exit"] = tts_exit
        if legacy:
            line["legacy_32bit_q16384"] = legacy
        if batch_check:
            line["batch_check"] = batch_check
        if world == 1 and not args.no_cpu_baseline:
            f = nd.bfs(A, D, ovr, Q)
            base, idx, ref_verdict = cpu_dfs_sample(lambda k: f.get(k), len(f), 20.0)
            f.free()
            # the reference's verdicts of the sample against the GPU's (free parity check on the box itself)
            assert (np.asarray(ref_verdict) == verdict[np.asarray(idx)]).all(), \
                "reference CPU verdicts of the sample differ from the GPU's"
            base["parity_checked"] = int(len(idx))
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tts", action="store_true")
    ap.add_argument("--warmup-legacy", action="store_true",
                    help="run the warm-up steps on the small legacy workload (for workloads whose step takes minutes)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.steps = max(args.steps, 1)
    if args.impl == "reference":
        return run_reference(args, args.workload)
    return run_b200(args, args.workload)


if __name__ == "__main__":
    sys.exit(main())
