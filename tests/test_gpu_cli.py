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
