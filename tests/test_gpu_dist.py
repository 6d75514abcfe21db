"""Multi-GPU equivalence on real hardware: launches tests/dist_check.py under torchrun on 2 GPUs (V, W and F
cycles; iterate bit-identical to one GPU, norms to 1e-12).  Skipped on a single-GPU box -- the schedule itself is
proven on CPU over gloo (test_dist_gloo.py, test_dist_gloo_fcycle.py)."""
import os
import subprocess
import sys

import pytest

import pmg_b200 as pmg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_gpus_match_one_gpu(p2p):
    if pmg.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PMG_P2P=p2p)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541" if p2p == "1" else "29542",
                        os.path.join(ROOT, "tests", "dist_check.py"), "1025"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "FAIL" not in p.stdout and p.stdout.count(" OK") >= 5
