"""Sharded crowd over >= 2 GPUs (NCCL all-gather of the payload every step, or the peer-memory exchange
folded into the step's kernels) equals the single-GPU crowd.  Skipped on a one-GPU box; tools/check_sharded.py is the rank program."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_sharded_equals_single_gpu(exchange):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533" if exchange == "nccl" else "29534",
           os.path.join(ROOT, "tools", "check_sharded.py")] + (["--peer"] if exchange == "peer" else [])
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('"test": "sharded_vs_single_gpu"') >= 3 and '"ok": false' not in r.stdout
