"""ndp_dfs_batch over several devices of one process (static split k -> k mod G, peer-visible
early-exit words): same verdicts, winner and model as one device.  Needs >= 2 GPUs (skipped on
a single-GPU box; tools/check_batch.py is the same check as a script for `gpurun --gpus N`)."""
import numpy as np
import pytest

import util_cnf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine", ["index", "scan"])
def test_batch_equals_single_device(engine, oracle, golden, cnf_text):
    import ndp5_b200 as nd
    G = min(nd.device_count(), 4)
    if G < 2:
        pytest.skip("needs at least 2 GPUs")
    nd.set_engine(engine)
    try:
        for c in golden["cases"]:
            if "verdict" not in c or c["bits"] > 24:
                continue
            A = oracle.parse(cnf_text(c["cnf"]))
            D, ovr, Q = util_cnf.settings_for(c["mode"], c["value"])
            fr = nd.bfs(A, D, ovr, Q)
            r = fr.dfs_batch(n_gpus=G, early_exit=False)
            assert r["verdict"].tolist() == c["verdict"]
            assert r["stats"]["nodes"] == sum(c["nodes"])
            if c["winner"] >= 0:
                assert r["winner"] == c["winner"] and r["model"].tolist() == c["sat"][str(c["winner"])]["model"]
                e = fr.dfs_batch(n_gpus=G, early_exit=True)
                assert e["winner"] == c["winner"] and e["model"].tolist() == r["model"].tolist()
            else:
                assert r["winner"] is None
            fr.free()
    finally:
        nd.set_engine("auto")


def test_cross_process_early_exit_words():
    """One process per GPU (torchrun, 2 ranks): exit words exchanged as CUDA IPC handles over NCCL, lowered by the kernels
    over NVLink; the job's winner and model are the single-GPU ones and the other rank stands down."""
    import os
    import subprocess
    import sys
    import ndp5_b200 as nd
    if nd.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29591", os.path.join(root, "tools", "check_shared_exit.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "MISMATCH" not in r.stdout and r.stdout.count(" OK") >= 8
