"""Multi-GPU parity on real devices: clips sharded over 2 ranks (one process per GPU,
NCCL), the all-gathered per-clip MSE vector and SR frames must equal the single-GPU result bit for bit and arrive
in clip order (model/pfnl.py:90,139-141).  Skipped on a box with fewer than 2 GPUs (the world-size-2 host logic is
covered on CPU with gloo in tests/test_dist_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_single_gpu_bit_for_bit(built_lib):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert tail.count("sharded==single-GPU: True") == 2, tail
