#!/usr/bin/env python3
"""Reference-produced fixture for the NORTH-STAR frontier: rsaFACT-64bit -q 4096 (tests/golden/big_rsaFACT-64bit_q4096_sample.npz).

One subproblem of this frontier is ~2.4e8 DFS nodes (~1.4e12 clause evaluations): hours of one host core for the
reference, ~40 ms for one B200.  So the fixture holds a handful of subproblems, each searched by the UNMODIFIED
reference (oracle/_ref/libndp_ref_prof.so: Satisfy_iterative_BFS for the frontier, Satisfy_iterative(A, true) for the
subproblem, PROFILE_SCOPE call count = DFS nodes), one process per subproblem, results written as they finish:

    python tools/make_golden_64bit.py start [--ref 666,0,665,...] [--port 666,0]   # BFS once, then one nice'd process per index
    python tools/make_golden_64bit.py collect                                       # whatever has finished -> the .npz
    python tools/make_golden_64bit.py worker ref|port <index>                       # (internal)

`--ref` indices get verdict / node count / model from the reference; `--port` indices get clause evaluations (and a second
node count) from the plain-C restatement oracle/ndp_oracle.c.  collect stores evals only where the port's node count
equals the reference's (same rule as tools/make_golden_big.py), 0 = not available.
The .npz holds: index, verdict, nodes, evals (0 = no port run), ref_seconds, model_<k> for SAT subproblems
(choices_k ++ DFS choices = final_choices of NDP:1430-1431), the BFS counters, cnf / mode / value.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cbind  # noqa: E402
import gen_cnf  # noqa: E402

WORK = os.environ.get("GOLDEN64_WORK", "/tmp/golden64")
OUT = os.path.join(ROOT, "tests", "golden", "big_rsaFACT-64bit_q4096_sample.npz")
CNF, Q = "rsaFACT-64bit", 4096


def start(ref_idx, port_idx):
    os.makedirs(WORK, exist_ok=True)
    ref = cbind.RefLib()
    A = ref.parse(gen_cnf.dimacs_text(*gen_cnf.NAMED[CNF]))
    t0 = time.time()
    h, info = ref.bfs_handle(A, 0, True, Q)
    n = ref._fsize(h)
    print("%s -q %d: reference BFS %.1f s, frontier %d, iterations %d, task_count %d" % (
        CNF, Q, time.time() - t0, n, info["iterations"], info["task_count"]), flush=True)
    json.dump(dict(iterations=info["iterations"], task_count=info["task_count"], size=n), open(os.path.join(WORK, "bfs.json"), "w"))
    for k in sorted(set(ref_idx) | set(port_idx)):
        ncl, nch = ref._fncl(h, k), ref._fnch(h, k)
        cl = np.zeros((ncl, 3), dtype=np.int32)
        ch = np.zeros(nch, dtype=np.int32)
        ref._fcopy(h, k, cl.ctypes.data_as(cbind._i32p), ch.ctypes.data_as(cbind._i32p))
        np.savez(os.path.join(WORK, "item_%d.npz" % k), clauses=cl, choices=ch)
    ref.frontier_free(h)
    for kind, idx in (("ref", ref_idx), ("port", port_idx)):
        for k in idx:
            log = open(os.path.join(WORK, "%s_%d.log" % (kind, k)), "w")
            subprocess.Popen(["nice", "-n", "19", sys.executable, os.path.abspath(__file__), "worker", kind, str(k)],
                             stdout=log, stderr=subprocess.STDOUT, start_new_session=True)
            print("started %s worker for subproblem %d" % (kind, k), flush=True)


def worker(kind, k):
    item = np.load(os.path.join(WORK, "item_%d.npz" % k))
    cl, ch = item["clauses"], item["choices"]
    t0 = time.time()
    if kind == "ref":
        r = cbind.RefLib(prof=True).dfs(cl)
        res = dict(index=k, kind=kind, verdict=1 if r["models"] else 0, nodes=int(r["nodes"]), seconds=time.time() - t0,
                   model=[int(x) for x in ch] + [int(x) for x in r["model"]] if r["models"] else [])
    else:
        r = cbind.OracleLib().dfs(cl)
        res = dict(index=k, kind=kind, verdict=1 if r["models"] else 0, nodes=int(r["nodes"]), evals=int(r["clause_evals"]),
                   seconds=time.time() - t0,
                   model=[int(x) for x in ch] + [int(x) for x in r["model"]] if r["models"] else [])
    tmp = os.path.join(WORK, "%s_%d.json.tmp" % (kind, k))
    json.dump(res, open(tmp, "w"))
    os.replace(tmp, os.path.join(WORK, "%s_%d.json" % (kind, k)))
    print("done", {x: res[x] for x in res if x != "model"}, flush=True)


def collect():
    bfs = json.load(open(os.path.join(WORK, "bfs.json")))
    refs, ports = {}, {}
    for name in os.listdir(WORK):
        if name.endswith(".json") and name != "bfs.json":
            r = json.load(open(os.path.join(WORK, name)))
            (refs if r["kind"] == "ref" else ports)[r["index"]] = r
    idx = sorted(refs)
    if not idx:
        print("no reference result has finished yet")
        return 1
    evals = []
    for k in idx:
        p = ports.get(k)
        if p is not None:
            assert p["nodes"] == refs[k]["nodes"] and p["verdict"] == refs[k]["verdict"] and p["model"] == refs[k]["model"], \
                ("port and reference disagree on subproblem", k, p["nodes"], refs[k]["nodes"])
        evals.append(p["evals"] if p is not None else 0)
    extra = {"model_%d" % k: np.asarray(refs[k]["model"], dtype=np.int32) for k in idx if refs[k]["verdict"] == 1}
    np.savez_compressed(OUT, index=np.asarray(idx, dtype=np.int64),
                        verdict=np.asarray([refs[k]["verdict"] for k in idx], dtype=np.uint8),
                        nodes=np.asarray([refs[k]["nodes"] for k in idx], dtype=np.int64),
                        evals=np.asarray(evals, dtype=np.uint64),
                        ref_seconds=np.asarray([refs[k]["seconds"] for k in idx], dtype=np.float64),
                        iterations=np.int64(bfs["iterations"]), task_count=np.int64(bfs["task_count"]),
                        size=np.int64(bfs["size"]), cnf=np.str_(CNF), mode=np.str_("q"), value=np.int64(Q),
                        note=np.str_("hand-picked subproblems of the north-star frontier; verdict / nodes / model from the "
                                     "unmodified reference, clause evals (0 = none) from the oracle port where its node "
                                     "count equals the reference's"), **extra)
    for k in idx:
        print("subproblem %4d: verdict %d, %d nodes, evals %d, reference %.0f s" % (
            k, refs[k]["verdict"], refs[k]["nodes"], evals[idx.index(k)], refs[k]["seconds"]))
    print("wrote %s (%d bytes)" % (OUT, os.path.getsize(OUT)))
    return 0


def _ints(s):
    return [int(x) for x in s.split(",") if x]


def main(argv):
    if len(argv) >= 2 and argv[1] == "worker":
        worker(argv[2], int(argv[3]))
        return 0
    if len(argv) >= 2 and argv[1] == "collect":
        return collect()
    if len(argv) >= 2 and argv[1] == "start":
        ref_idx, port_idx = [666, 0, 665, 667, 2048, 4095], [666, 0]
        for i, a in enumerate(argv):
            if a == "--ref":
                ref_idx = _ints(argv[i + 1])
            if a == "--port":
                port_idx = _ints(argv[i + 1])
        start(ref_idx, port_idx)
        return 0
    print(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv))
