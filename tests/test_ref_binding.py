"""The reference-side binding of INTEGRATION.md §1-§2, compiled: the reference translation
unit with its two hot-path call sites swapped for ndp_bfs / ndp_dfs_batch by the committed patch script
(oracle/ref_binding/patch_reference.py), built against oracle/shim/mpi.h and linked against libndp_b200.so.
Here (no GPU): the anchors are still where the patch expects them, the patched program builds, links the product
library and -- without a device -- fails loudly instead of computing anything on the CPU.  On the GPU box
tests/test_gpu_cli.py runs it end to end."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/NDP-5_6_7.cpp"
EXE = os.path.join(ROOT, "oracle", "_ref", "NDP-5_6_7_b200")
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_binding"))

pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the reference sources are not on this machine")


def test_patch_applies_to_a_copy_and_touches_only_the_two_call_sites(tmp_path):
    import patch_reference
    src = tmp_path / "NDP-5_6_7.cpp"
    shutil.copy(REF, src)
    out = tmp_path / "NDP-5_6_7_b200.cpp"
    assert patch_reference.main(["patch_reference.py", str(src), str(out)]) == 0
    a, b = open(src, errors="replace").read().splitlines(), open(out, errors="replace").read().splitlines()
    import difflib
    removed = [l[2:] for l in difflib.ndiff(a, b) if l.startswith("- ")]
    # exactly: the BFS call line and the five lines of the process_queue call
    assert len(removed) == 6, removed
    assert "Satisfy_iterative_BFS(clauses, depth" in removed[0] and "process_queue(bfs_queue" in removed[1]
    text = "\n".join(b)
    for name in ("ndp_bfs(", "ndp_dfs_batch(", "ndp_frontier_from_host(", "ndp_frontier_attach_root(", "generate_output(found"):
        assert name in text
    assert "Satisfy_iterative(" in text      # the reference's own DFS is still in the TU, just no longer called from main


def test_patched_reference_builds_and_links_the_product_library():
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/NDP-5_6_7_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(EXE)
    ldd = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libndp_b200.so" in ldd and "not found" not in ldd.split("libndp_b200.so")[1].splitlines()[0]


def test_no_cpu_fallback_behind_the_binding(tmp_path):
    import gen_cnf
    try:
        import ndp5_b200
        if ndp5_b200.device_count() > 0:
            pytest.skip("a GPU is present: the end-to-end run is tests/test_gpu_cli.py")
    except ImportError:
        pytest.skip("product library not built")
    if not os.path.exists(EXE):
        pytest.skip("binding not built")
    gen_cnf.write_named("rsaFACT-16bit", str(tmp_path / "r16.dimacs"))
    r = subprocess.run([EXE, "r16.dimacs", "-d", "60"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0                       # ndp_bfs -> NDP_E_CUDA -> the binding throws (NDP has no CPU path here)
    assert "Input Number: 35611" in r.stdout and "FACT" not in r.stdout
