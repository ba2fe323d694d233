"""End-to-end runs of the drop-in program (ndp-5_b200/bin/NDP-5_6_7): the reference's command
line over the B200 hot path -- report text, result file, -s JSON (checked against the oracle's
frontier), -r resume, Prime!, the empty-frontier case, usage errors."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import gen_cnf
import util_cnf

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ndp-5_b200", "bin", "NDP-5_6_7")


def run(args, cwd, stdin=None):
    return subprocess.run([BIN] + args, cwd=cwd, input=stdin, capture_output=True, text=True, timeout=300)


@pytest.fixture()
def work(tmp_path):
    for name in ("rsaFACT-16bit", "prime-16bit", "rsaFACT-24bit"):
        gen_cnf.write_named(name, str(tmp_path / (name + ".dimacs")))
    return tmp_path


def test_solve_save_resume(work, oracle):
    r = run(["rsaFACT-16bit.dimacs", "-d", "60", "-s"], str(work))
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "saving BFS after completion!" in out and "found a solution!" in out
    assert "Input Number: 35611" in out and "verified." in out
    assert {int(re.search(r"FACT 1: (\d+)", out).group(1)), int(re.search(r"FACT 2: (\d+)", out).group(1))} == {149, 239}
    assert re.search(r"  Queue Size: 4\n", out) and re.search(r"       Depth: 60\n", out)
    assert re.search(r" Total Cores: 2\n", out) and " NDP-version: 5.6.7" in out
    # result file: same block + the Assignments line; name pattern of formatFilename
    saved = re.search(r"Result saved: (.*)\n", out).group(1)
    assert re.fullmatch(r"NDP-5_6_7_rsaFACT-16bit_[0-9a-f]{5}-d60\.txt", os.path.basename(saved))
    text = open(saved).read()
    model = [int(x) for x in re.search(r" Assignments: \[(.*)\]", text).group(1).split(", ")]
    A = oracle.parse(open(work / "rsaFACT-16bit.dimacs").read())
    assert util_cnf.satisfies(A, model)
    # reference order of the model: choices of the winning subproblem, then its DFS choices
    o = oracle.bfs(A, 60, False, -1)
    first = next((np.concatenate([ch, oracle.dfs(cl)["model"]]).tolist() for cl, ch in o["frontier"]
                  if oracle.dfs(cl)["models"]))
    assert model == first
    # -s JSON == the oracle's frontier
    jname = re.search(r"BFS results saved to (.*)\n", out).group(1)
    assert re.fullmatch(r"bfs_NDP-5_6_7_rsaFACT-16bit_[0-9a-f]{5}-d60\.json", jname)
    j = json.load(open(work / jname))
    assert [e["choices"] for e in j["bfs_results"]] == [ch.tolist() for _, ch in o["frontier"]]
    assert [e["clauses"] for e in j["bfs_results"]] == [cl.tolist() for cl, _ in o["frontier"]]
    # -r <file>: same answer from the saved frontier
    r2 = run(["rsaFACT-16bit.dimacs", "-r", jname], str(work))
    assert r2.returncode == 0, r2.stdout + r2.stderr
    assert "Loading: " + jname in r2.stdout and "Queue size: 4" in r2.stdout and "verified." in r2.stdout
    saved2 = re.search(r"Result saved: (.*)\n", r2.stdout).group(1)
    model2 = [int(x) for x in re.search(r" Assignments: \[(.*)\]", open(saved2).read()).group(1).split(", ")]
    assert model2 == model
    # -r without a name: interactive prompt over ./*bfs_* (NDP:381-421)
    r3 = run(["rsaFACT-16bit.dimacs", "-r"], str(work), stdin="1\n")
    assert r3.returncode == 0 and "Available BFS results files:" in r3.stdout and "verified." in r3.stdout


def test_prime_and_default_depth(work):
    r = run(["prime-16bit.dimacs"], str(work))          # default depth formula, world_size 2 => 969
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Input Number: 35051\n              Prime!" in r.stdout
    assert re.search(r"       Depth: 969\n", r.stdout) and re.search(r"  Queue Size: 32\n", r.stdout)
    assert re.search(r"NDP-5_6_7_prime-16bit_[0-9a-f]{5}-auto\.txt", r.stdout)
    r = run(["prime-16bit.dimacs", "-q", "64", "-x", "-s"], str(work))
    assert r.returncode == 0 and "Prime!" in r.stdout and "Proc. Queues: 64" in r.stdout


def test_24bit_q256(work):
    r = run(["rsaFACT-24bit.dimacs", "-q", "256"], str(work))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "verified." in r.stdout
    assert {int(re.search(r"FACT 1: (\d+)", r.stdout).group(1)), int(re.search(r"FACT 2: (\d+)", r.stdout).group(1))} == {3323, 3889}


def test_empty_frontier_and_errors(work):
    r = run(["rsaFACT-16bit.dimacs", "-q", "256"], str(work))     # BFS exhausts the tree (SURVEY.md finding 3)
    assert r.returncode == 2 and "frontier is empty" in r.stdout
    assert run(["rsaFACT-16bit.dimacs", "-d", "5", "-q", "7"], str(work)).returncode == 1
    assert "Unknown or invalid argument" in run(["rsaFACT-16bit.dimacs", "-z"], str(work)).stdout
    assert "Could not open file" in run(["nope.dimacs"], str(work)).stdout
    (work / "bad.dimacs").write_text("c no header\n1 2 3 0\n")
    assert "Could not extract number of variables" in run(["bad.dimacs"], str(work)).stdout


# ---- the reference itself, with its two hot-path call sites bound to the library ----
REF_BOUND = os.path.join(ROOT, "oracle", "_ref", "NDP-5_6_7_b200")


def _report_block(out):
    """The report lines that do not depend on time, host, process id or program name."""
    keep = ("Input Number:", "FACT 1:", "FACT 2:", "verified.", "Prime!", "Bits:", "VARs:", "Clauses:", "Queue Size:",
            "Depth:", "NDP-version:", "DIMACS:")
    return [l.strip() for l in out.splitlines() if any(l.strip().startswith(k) for k in keep)]


@pytest.mark.skipif(not os.path.exists(REF_BOUND), reason="the patched reference was not built (needs /root/reference at build time)")
def test_reference_program_bound_to_the_library(work, oracle):
    """oracle/_ref/NDP-5_6_7_b200 = the reference TU with Satisfy_iterative_BFS (NDP:1651) -> ndp_bfs and process_queue
    (NDP:1673) -> ndp_dfs_batch, everything else the reference's own code: same factors / Prime!, same report block as
    the drop-in program, the reference's own -s JSON equal to the oracle's frontier, and -r through its own reader."""
    def ref(args, stdin=None):
        return subprocess.run([REF_BOUND] + args, cwd=str(work), input=stdin, capture_output=True, text=True, timeout=300)
    r = ref(["rsaFACT-16bit.dimacs", "-d", "60", "-s"])
    assert r.returncode == 0, r.stdout[-2000:]
    out = r.stdout
    assert "found a solution!" in out and "verified." in out
    assert {int(re.search(r"FACT 1: (\d+)", out).group(1)), int(re.search(r"FACT 2: (\d+)", out).group(1))} == {149, 239}
    mine = run(["rsaFACT-16bit.dimacs", "-d", "60"], str(work)).stdout
    assert _report_block(out) == _report_block(mine)
    saved = re.search(r"Result saved: (.*)\n", out).group(1)
    model = [int(x) for x in re.search(r" Assignments: \[(.*)\]", open(saved).read()).group(1).split(", ")]
    A = oracle.parse(open(work / "rsaFACT-16bit.dimacs").read())
    o = oracle.bfs(A, 60, False, -1)
    first = next((np.concatenate([ch, oracle.dfs(cl)["model"]]).tolist() for cl, ch in o["frontier"]
                  if oracle.dfs(cl)["models"]))
    assert model == first
    # the reference's own JSON writer over the library's frontier == the oracle's frontier
    jname = [p for p in os.listdir(work) if p.startswith("bfs_NDP-5_6_7_b200_rsaFACT-16bit_")][0]
    j = json.load(open(work / jname))
    assert [e["choices"] for e in j["bfs_results"]] == [ch.tolist() for _, ch in o["frontier"]]
    assert [e["clauses"] for e in j["bfs_results"]] == [cl.tolist() for cl, _ in o["frontier"]]
    # -r: the reference's reader -> ndp_frontier_from_host + ndp_frontier_attach_root -> same model
    r2 = ref(["rsaFACT-16bit.dimacs", "-r", jname])
    assert r2.returncode == 0 and "verified." in r2.stdout, r2.stdout[-2000:]
    saved2 = re.search(r"Result saved: (.*)\n", r2.stdout).group(1)
    model2 = [int(x) for x in re.search(r" Assignments: \[(.*)\]", open(saved2).read()).group(1).split(", ")]
    assert model2 == model
    # UNSAT input: Prime!, same block as the drop-in program
    rp = ref(["prime-16bit.dimacs", "-q", "64"])
    assert rp.returncode == 0 and "Input Number: 35051\n              Prime!" in rp.stdout, rp.stdout[-2000:]
    assert _report_block(rp.stdout) == _report_block(run(["prime-16bit.dimacs", "-q", "64"], str(work)).stdout)
