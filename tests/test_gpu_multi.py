"""GPU (-m gpu, needs >= 2 visible GPUs, skips otherwise): the template-sharded matcher under torchrun, one process per GPU.
Every rank must return exactly the oracle's match list, frame after frame, with both exchange modes (peer-memory push fused into
the sort kernel, and the NCCL all-gather fallback).  The worker is tools/check_multi_gpu.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_match_parity_under_torchrun(world):
    n = _n_gpus()
    if n < world:
        pytest.skip("%d GPU(s) visible, the test needs %d" % (n, world))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail
    assert "MULTI-GPU PARITY PASS" in r.stdout, tail
