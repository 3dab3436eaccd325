"""Config 5 behind the C-ABI: vacv_cuda_normalize_batch_global (NCCL transport, libvacv_dist.so) and
vacv_cuda_normalize_batch_global_p2p (peer-memory transport, libvacv_cuda.so).  Reference semantics: the auto-statistics branch of
Normalize::normalize_naive (src/cv/normalize.cpp:84-121) extended to the batch; exact u64 sums.

The single-rank and the two-ranks-on-one-GPU cases run on a 1-GPU box; the multi-GPU case needs >= 2 GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _torchrun(world, *args):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_c5_worker.py"), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
@pytest.mark.parametrize("layout", ["NHWC", "NCHW"])
def test_batch_global_single_rank_equals_unfused_operators(oracle, transport, layout):
    """world = 1 (a real NCCL communicator of one rank / an unconnected exchange): the C entry must equal
    sums_u8 -> finalize -> normalize on the same frames, and the oracle."""
    import vacv_b200 as vacv
    from arm_neon_opencv_b200 import distributed as vd
    from oracle_lib import NCHW, NHWC
    lay, lay_o = (vacv.NHWC, NHWC) if layout == "NHWC" else (vacv.NCHW, NCHW)
    rng = np.random.default_rng(5)
    b, h, w = 5, 72, 128
    frames = rng.integers(0, 256, (b, h, w, 3) if layout == "NHWC" else (b, 3, h, w), dtype=np.uint8)
    src = torch.from_numpy(frames).cuda()
    t = vd.NcclComm() if transport == "nccl" else vd.P2PExchange()
    try:
        out, ms = vd.normalize_batch_global(t, src, lay)
        sums = vacv.sums_u8(src, lay)
        mean, std = vacv.finalize_mean_stddev(sums, b * w * h)
        assert torch.equal(ms[0], mean[0]) and torch.equal(ms[1], std[0])
        assert torch.equal(out, vacv.normalize(src, lay, mean[0], std[0]))
        m, s = oracle.finalize_mean_stddev(sums.cpu().numpy()[0].astype(np.uint64).ravel(), 3, b * w * h)
        want = oracle.normalize(frames[1], w * h, 3, lay_o, m, s)
        assert np.array_equal(out[1].cpu().numpy().view(np.uint32).ravel(), want.view(np.uint32).ravel())
    finally:
        t.close()


def test_batch_global_p2p_is_cuda_graph_capturable():
    """The peer exchange keeps its epoch on the device, so the whole config-5 step replays from a CUDA graph."""
    import vacv_b200 as vacv
    from arm_neon_opencv_b200 import distributed as vd
    src = torch.randint(0, 256, (3, 64, 96, 3), dtype=torch.uint8, device="cuda")
    t = vd.P2PExchange()
    try:
        want, ms_want = vd.normalize_batch_global(t, src, vacv.NHWC)
        out = torch.empty_like(want)
        work = torch.empty(16, dtype=torch.int64, device="cuda")
        ms = torch.empty((2, 3), dtype=torch.float32, device="cuda")
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            vd.normalize_batch_global(t, src, vacv.NHWC, out=out, work=work, mean_std=ms)   # warm-up outside capture
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                vd.normalize_batch_global(t, src, vacv.NHWC, out=out, work=work, mean_std=ms)
        for _ in range(3):
            out.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, want) and torch.equal(ms, ms_want)
        assert t.timed_out() == 0
    finally:
        t.close()


def test_batch_global_p2p_two_ranks_sharing_one_gpu():
    """Two processes on cuda:0 exchange their sums through IPC-mapped peer memory: the multi-rank path on a 1-GPU box."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    r = _torchrun(2, "p2p", "1")
    assert r.returncode == 0 and "C5_DIST_OK 2 p2p" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_batch_global_over_nvlink_multi_gpu():
    """One process per GPU: the NCCL all-reduce and the peer-memory exchange over NVLink / NVSwitch."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    r = _torchrun(world, "nccl,p2p", "0")
    assert r.returncode == 0 and f"C5_DIST_OK {world} nccl,p2p" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
