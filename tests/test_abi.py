"""The C-ABI library loads and exports every symbol include/ndp_b200.h declares (no compute
without a GPU), and fails loudly -- not silently on a CPU path -- when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ndp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ndp_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ["ndp_bfs", "ndp_frontier_size", "ndp_frontier_get", "ndp_frontier_from_host", "ndp_frontier_free",
                 "ndp_dfs_shard", "ndp_dfs_batch", "ndp_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    import ndp5_b200
    for s in declared_symbols():
        assert hasattr(ndp5_b200.lib, s), "libndp_b200.so does not export %s" % s
    assert b"sm_100a" in ndp5_b200.lib.ndp_version()


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "ndp_b200.h")).read()
    for cite in ["NDP:1651", "NDP:881", "NDP:1424", "NDP:784", "NDP:512", "NDP:1207"]:
        assert cite in text


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C11 (no C++ types in the signatures) and a C
    program must link against the library."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include "ndp_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { int n = -1; ndp_device_count(&n); printf("%s %d\\n", ndp_version(), n >= 0); return 0; }\n')
    exe = tmp_path / "use"
    lib = os.path.join(ROOT, "ndp-5_b200", "lib")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", str(exe), "-L", lib, "-lndp_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ndp-b200")


def test_invalid_arguments_are_rejected_without_a_device():
    import ndp5_b200 as nd
    assert nd.lib.ndp_bfs(None, 0, 1, 0, -1, None, None, None, None) == nd.NDP_E_INVALID
    assert b"null" in nd.lib.ndp_last_error()
    assert nd.lib.ndp_frontier_size(None) == 0
    assert nd.lib.ndp_dfs_shard(None, 0, 1, 0, None, None, None, None, 0, None, None, None, None) == nd.NDP_E_INVALID


def test_progress_and_piece_calls_reject_bad_arguments_without_a_device():
    import ctypes as C
    import ndp5_b200 as nd
    assert nd.dfs_progress() == (0, 0, 0)                      # nothing is searching
    h = C.c_void_p(None)
    assert nd.lib.ndp_dfs_pieces_begin(None, 0, 1, 0, None, None, 0, C.byref(h), None, None) == nd.NDP_E_INVALID
    assert nd.lib.ndp_dfs_pieces_fold(None, None, None, None, None, None, None, 0, None, None) == nd.NDP_E_INVALID
    nd.lib.ndp_dfs_pieces_end(None)                            # a no-op, like free(NULL)


def test_no_cpu_fallback_when_no_gpu():
    """Where there is no CUDA device the compute entry points must fail with NDP_E_CUDA."""
    import ndp5_b200 as nd
    if nd.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(nd.NdpError) as ei:
        nd.bfs(np.array([[0, 0, 1]], dtype=np.int32), 1, False, -1)
    assert ei.value.code == nd.NDP_E_CUDA
    with pytest.raises(nd.NdpError) as ei:
        nd.frontier_from_host([(np.array([[0, 0, 1]], dtype=np.int32), np.zeros(0, np.int32))])
    assert ei.value.code == nd.NDP_E_CUDA


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under ndp-5_b200/ or include/ may name it."""
    bad = []
    for base in ("ndp-5_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".py", "Makefile")):
                    t = open(os.path.join(dp, fn), errors="replace").read()
                    if re.search(r"ndp_oracle|orc_[a-z]+\(|libndp_ref|oracle/", t):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
