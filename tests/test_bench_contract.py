"""bench.py's reference arm on a workload small enough for the CPU suite: the JSON line carries the keys the
driver's contract names (impl, metric, unit, value, higher_is_better, cpu_baseline{kind, cores, sample}, e2e{...})
and the numbers are consistent with each other."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libndp_ref.so")) and \
            not os.path.exists(os.path.join(ROOT, "oracle", "libndp_oracle.so")):
        pytest.skip("no oracle library built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "rsaFACT-16bit:q:32", "--no-tts"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "subproblems resolved/s" and line["unit"] == "subproblems/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "subproblems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "16bit" in line["config"]["workload"]


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
