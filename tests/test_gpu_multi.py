"""Two-rank checks on real GPUs (skipped on boxes with one GPU; the gloo tests cover the host logic
there): the peer-memory transport of the sharded dedup and of the sharded anti-join equals the NCCL transport,
the exact-size paths and the url-id ground truth, for keep = first / last / False, with and without NaN cells,
on repeated steps (the buffers reset by the previous step)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_exchange_equals_nccl_exchange():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", str(ROOT / "tools" / "xchg_check.py"), "400000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "p2p == nccl: True  p2p == exact-size path: True  dedup == url-id ground truth: True" in out.stdout, out.stdout[-2000:]
    assert "antijoin p2p == nccl: True  antijoin p2p == exact-size path: True  antijoin == url-id ground truth: True" in out.stdout, \
        out.stdout[-2000:]
    assert "joint exchange == separate exchanges (both transports): True" in out.stdout, out.stdout[-2000:]
    assert "requested p2p: using p2p" in out.stdout, out.stdout[-2000:]
