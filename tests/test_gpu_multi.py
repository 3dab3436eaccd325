"""Multi-GPU test of the one exchange on the path (config 5: exact sums -> NCCL all-reduce -> identical statistics).
Needs >= 2 GPUs in the box; skipped otherwise (the host logic is covered on CPU by test_distributed_cpu.py over gloo)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_config5_statistics_exchange_over_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_c5_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and f"NCCL_C5_OK {world}" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
