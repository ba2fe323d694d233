"""oracle/_ref/ndp_ref_mpirun runs the WHOLE unmodified reference program (main, dispatcher, workers) as
forked ranks over oracle/shim_fork/mpi.h.  These tests pin that harness: it must reproduce what the
program prints under mpirun -- the factors with `verified.` on a SAT input, `Prime!` on an UNSAT one --
and agree with the direct-call oracle on the BFS counters it reports."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ndp_ref_mpirun")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="the compiled reference did not travel here")


def run(tmp_path, cnf_text, name, np_ranks, *flags):
    path = tmp_path / (name + ".dimacs")
    path.write_text(cnf_text(name))
    out = subprocess.run([EXE, "-np", str(np_ranks), str(path), *flags], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True, errors="replace", timeout=120).stdout
    assert re.search(r"\[ref_mpirun\] np=%d wall=[0-9.]+ s" % np_ranks, out), out[-400:]
    return out


def test_whole_program_sat(tmp_path, cnf_text, oracle):
    out = run(tmp_path, cnf_text, "rsaFACT-16bit", 4, "-d", "60")
    assert "found a solution!" in out and "verified." in out
    assert sorted(re.findall(r"FACT \d: (\d+)", out)) == ["149", "239"]
    o = oracle.bfs(oracle.parse(cnf_text("rsaFACT-16bit")), 60, False, -1, materialise=False)
    assert "Queue size: %d - Depth: %d - Tasks: %d" % (o["size"], o["iterations"], o["task_count"]) in out
    assert list(tmp_path.glob("NDP-5_6_7_rsaFACT-16bit_*-d60.txt"))


def test_whole_program_unsat_and_save(tmp_path, cnf_text):
    out = run(tmp_path, cnf_text, "prime-16bit", 4, "-q", "64", "-s")
    assert "Prime!" in out and "found a solution" not in out
    assert "Queue Size: 64" in out
    assert list(tmp_path.glob("bfs_NDP-5_6_7_prime-16bit_*-q64.json"))
