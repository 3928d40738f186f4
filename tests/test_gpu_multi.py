"""Real multi-GPU run (NCCL between processes) of the sharded pipelines and the CLI; skipped on a single-GPU box.
The same properties are covered on one GPU by tests/test_gpu_sharded.py (threads + a fake collective)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs >= 2 GPUs (tests/test_gpu_persistent.py runs the peer exchange with "
                                             "several row blocks on ONE GPU)")
@pytest.mark.timeout(600)
def test_sharded_pipelines_and_cli_under_torchrun():
    world = min(_gpu_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "multigpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=580)
    assert p.returncode == 0 and "ALL PASS" in p.stdout and "FAIL" not in p.stdout, p.stdout[-3000:]
