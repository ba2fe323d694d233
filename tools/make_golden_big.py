#!/usr/bin/env python3
"""Reference-produced fixtures for the LARGE settings the bench and the GPU parity tests quote
(tests/golden/big_*.npz).  Same rules as tools/make_golden.py: every verdict and node count comes
from the UNMODIFIED reference (oracle/_ref/libndp_ref_prof.so: Satisfy_iterative_BFS for the
frontier, Satisfy_iterative per subproblem, PROFILE_SCOPE call counts = DFS nodes); clause_evals
come from the plain-C oracle port AFTER its node count was checked equal to the reference's.

    python tools/make_golden_big.py [case ...]        # default: all cases, in order

Cases (subproblem subsets are index lists into the reference's own frontier):
  rsaFACT-32bit_q16384   every subproblem (exhaustive): the old bench headline
  prime-32bit_q16384     every subproblem (all UNSAT)
  rsaFACT-48bit_q4096    every 8th subproblem (an interleaved sample: 512 of 4096, ~70 min on 6 cores incl. the port; the
                         committed file also holds 4, 20, 36, ... from a second run merged in: 768), then
                         (second file, *_prefix) indices 0..first_sat so that "lowest SAT index" is pinned
  rsaFACT-48bit_q65536   indices 0..first_sat+1 (BASELINE configs[2]; the reference BFS alone takes ~17 min); the committed
                         big_rsaFACT-48bit_q65536.npz adds every 128th subproblem from 4879 on (same functions, run by hand)
Each .npz holds: index (int64), verdict (uint8), nodes (int64), evals (uint64), and the BFS
counters iterations / task_count / size / bfs_clause_evals, plus cnf / mode / value as strings.
"""
import os
import sys
import time
import multiprocessing

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cbind  # noqa: E402
import gen_cnf  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WORKERS = int(os.environ.get("GOLDEN_WORKERS", max(1, (os.cpu_count() or 2) - 2)))

_ITEMS = None
_ORC = None


def _oracle_dfs(i):
    global _ORC
    if _ORC is None:
        _ORC = cbind.OracleLib()
    od = _ORC.dfs(_ITEMS[i][0])
    return od["nodes"], od["models"], od["clause_evals"]


def run_subset(refp, items, tag):
    """reference DFS (verdict, nodes) + oracle clause evals over `items`."""
    global _ITEMS
    t0 = time.time()
    h = refp.frontier_from_arrays(items)
    r = refp.dfs_frontier(h, 0, len(items), WORKERS, early_exit=False)
    refp.frontier_free(h)
    print("  [%s] reference DFS over %d subproblems: %.1f s wall, %d nodes" % (tag, len(items), time.time() - t0,
                                                                             int(r["nodes"].sum())), flush=True)
    _ITEMS = items
    t0 = time.time()
    with multiprocessing.get_context("fork").Pool(WORKERS) as pool:
        res = pool.map(_oracle_dfs, range(len(items)), chunksize=1)
    evals = np.zeros(len(items), dtype=np.uint64)
    for i, (nodes_i, models_i, ev_i) in enumerate(res):
        assert nodes_i == int(r["nodes"][i]), (tag, i, nodes_i, int(r["nodes"][i]))
        assert (models_i > 0) == (int(r["verdict"][i]) == 1), (tag, i)
        evals[i] = ev_i
    print("  [%s] oracle port agrees on every node count; clause evals %d (%.1f s)" % (tag, int(evals.sum()),
                                                                                    time.time() - t0), flush=True)
    return r["verdict"].astype(np.uint8), r["nodes"].astype(np.int64), evals


def frontier_of(ref, cnf, mode, value):
    text = gen_cnf.dimacs_text(*gen_cnf.NAMED[cnf])
    A = ref.parse(text)
    D, ovr, Q = (value, False, -1) if mode == "d" else (0, True, value)
    t0 = time.time()
    h, info = ref.bfs_handle(A, D, ovr, Q)
    n = ref._fsize(h)
    print("%s -%s %d: reference BFS %.1f s, frontier %d, iterations %d, task_count %d" % (
        cnf, mode, value, time.time() - t0, n, info["iterations"], info["task_count"]), flush=True)
    return h, info, n, A


def element(ref, h, k):
    ncl, nch = ref._fncl(h, k), ref._fnch(h, k)
    cl = np.zeros((ncl, 3), dtype=np.int32)
    ch = np.zeros(nch, dtype=np.int32)
    ref._fcopy(h, k, cl.ctypes.data_as(cbind._i32p), ch.ctypes.data_as(cbind._i32p))
    return cl, ch


def save(name, cnf, mode, value, info, n, index, verdict, nodes, evals, note):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "big_%s.npz" % name)
    np.savez_compressed(path, index=np.asarray(index, dtype=np.int64), verdict=verdict, nodes=nodes, evals=evals,
                        iterations=np.int64(info["iterations"]), task_count=np.int64(info["task_count"]),
                        size=np.int64(n), cnf=np.str_(cnf), mode=np.str_(mode), value=np.int64(value),
                        note=np.str_(note))
    print("wrote %s (%d bytes): %d subproblems, %d SAT, %d nodes, %d clause evals" % (
        path, os.path.getsize(path), len(index), int((verdict == 1).sum()), int(nodes.sum()), int(evals.sum())), flush=True)


def case_exhaustive(ref, refp, cnf, value):
    h, info, n, _ = frontier_of(ref, cnf, "q", value)
    items = [element(ref, h, k) for k in range(n)]
    ref.frontier_free(h)
    v, nd, ev = run_subset(refp, items, "%s q%d" % (cnf, value))
    save("%s_q%d" % (cnf, value), cnf, "q", value, info, n, np.arange(n), v, nd, ev,
         "every subproblem, exhaustive, reference verdict + nodes, oracle-port clause evals")


def case_sample_then_prefix(ref, refp, cnf, value, every, first_sat_claim, prefix_extra=1, sample=True):
    h, info, n, _ = frontier_of(ref, cnf, "q", value)
    if sample:
        idx = list(range(0, n, every))
        items = [element(ref, h, k) for k in idx]
        v, nd, ev = run_subset(refp, items, "%s q%d every %d" % (cnf, value, every))
        save("%s_q%d" % (cnf, value), cnf, "q", value, info, n, idx, v, nd, ev,
             "every %d-th subproblem (interleaved sample), reference verdict + nodes, oracle-port clause evals" % every)
    # prefix up to and including the claimed lowest SAT index (+ extra), in chunks so progress survives
    hi = min(n, first_sat_claim + 1 + prefix_extra)
    vs, nds, evs = [], [], []
    chunk = 16 * WORKERS
    for k0 in range(0, hi, chunk):
        idx = list(range(k0, min(hi, k0 + chunk)))
        items = [element(ref, h, k) for k in idx]
        v, nd, ev = run_subset(refp, items, "%s q%d prefix %d..%d" % (cnf, value, idx[0], idx[-1]))
        vs.append(v), nds.append(nd), evs.append(ev)
        save("%s_q%d_prefix" % (cnf, value), cnf, "q", value, info, n, np.arange(0, idx[-1] + 1),
             np.concatenate(vs), np.concatenate(nds), np.concatenate(evs),
             "subproblems 0..%d (prefix pinning the lowest SAT index), reference verdict + nodes, oracle-port "
             "clause evals" % idx[-1])
    ref.frontier_free(h)
    allv = np.concatenate(vs)
    sat = np.nonzero(allv == 1)[0]
    print("%s -q %d: lowest SAT index in the prefix: %s (claimed %d)" % (cnf, value, sat[0] if len(sat) else None,
                                                                       first_sat_claim), flush=True)


CASES = {
    "rsaFACT-32bit_q16384": lambda ref, refp: case_exhaustive(ref, refp, "rsaFACT-32bit", 16384),
    "prime-32bit_q16384": lambda ref, refp: case_exhaustive(ref, refp, "prime-32bit", 16384),
    "rsaFACT-48bit_q4096": lambda ref, refp: case_sample_then_prefix(ref, refp, "rsaFACT-48bit", 4096, 8, 301),
    "rsaFACT-48bit_q65536": lambda ref, refp: case_sample_then_prefix(ref, refp, "rsaFACT-48bit", 65536, 0, 4822,
                                                                     sample=False),
}


def main(argv):
    ref = cbind.RefLib()
    refp = cbind.RefLib(prof=True)
    for name in (argv[1:] or list(CASES)):
        t0 = time.time()
        CASES[name](ref, refp)
        print("== %s done in %.0f s" % (name, time.time() - t0), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
