"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame sharding and the batch-global statistics
exchange (config 5).  The per-rank sums come from the oracle here (no GPU); on the GPU the same exchange runs over
NCCL in bench_ops.py --workload c5."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, w, h, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("vacv_distributed", os.path.join(ROOT, "arm-neon-opencv_b200", "distributed.py"))
    vd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vd)
    from oracle_lib import NHWC, Oracle
    oracle = Oracle()
    frames = np.random.default_rng(42).integers(0, 256, (n_frames, h, w, 3), dtype=np.uint8)   # same on every rank
    b, e = vd.shard_range(n_frames, rank, world)
    sums = np.zeros(6, np.uint64)
    for i in range(b, e):
        oracle.sums_u8(frames[i], w * h, 3, NHWC, sums)
    t = torch.from_numpy(sums.astype(np.int64).reshape(1, 3, 2))
    vd.allreduce_sums(t)
    mean, std = vd.finalize_mean_stddev_host(t, n_frames * w * h)
    # every rank normalises only its shard with the GLOBAL statistics
    out = np.stack([oracle.normalize(frames[i], w * h, 3, NHWC, mean[0].numpy(), std[0].numpy()) for i in range(b, e)]) if e > b else np.zeros((0, h, w, 3), np.float32)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), out)
    np.save(os.path.join(out_dir, f"stats{rank}.npy"), np.stack([mean[0].numpy(), std[0].numpy()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_global_stats_exchange_world2(tmp_path, n_frames):
    w, h, world = 64, 48, 2
    mp.spawn(_worker, args=(world, free_port(), n_frames, w, h, str(tmp_path)), nprocs=world, join=True)
    from oracle_lib import NHWC, Oracle
    oracle = Oracle()
    frames = np.random.default_rng(42).integers(0, 256, (n_frames, h, w, 3), dtype=np.uint8)
    sums = np.zeros(6, np.uint64)
    for f in frames:
        oracle.sums_u8(f, w * h, 3, NHWC, sums)
    mean, std = oracle.finalize_mean_stddev(sums, 3, n_frames * w * h)
    want = np.stack([oracle.normalize(f, w * h, 3, NHWC, mean, std) for f in frames])
    s0, s1 = np.load(tmp_path / "stats0.npy"), np.load(tmp_path / "stats1.npy")
    assert np.array_equal(s0, s1), "ranks disagree on the global statistics"
    assert np.array_equal(s0.view(np.uint32), np.stack([mean, std]).view(np.uint32)), "sharded statistics != single-process statistics"
    got = np.concatenate([np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_shard_range_covers_everything():
    import importlib.util
    spec = importlib.util.spec_from_file_location("vacv_distributed", os.path.join(ROOT, "arm-neon-opencv_b200", "distributed.py"))
    vd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vd)
    for n in (0, 1, 7, 256, 1024):
        for world in (1, 2, 3, 8):
            rs = [vd.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1
