"""Multi-GPU parity through real NCCL ranks: runs tests/dist_check.py under torchrun when the box has
at least two GPUs (the single-GPU box skips it; the sharded logic is covered on CPU by test_sharding)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_engine_matches_oracle_on_all_gpus():
    import torch
    assert torch.cuda.is_available()
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (n.bit_length() - 1)
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "dist_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "dist_check: all passed" in r.stdout, r.stdout[-4000:]
