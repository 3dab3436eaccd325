"""Worker of tests/test_gpu_dist.py (one process per rank under torchrun): config 5 through the C entries
vacv_cuda_normalize_batch_global (NCCL, libvacv_dist.so) and vacv_cuda_normalize_batch_global_p2p (peer memory).

    _c5_worker.py <transports: nccl,p2p> <same_gpu: 0|1>

Every rank owns a ragged contiguous shard of the same seeded frames.  Checks per transport: mean / stddev have identical bits
on every rank and equal the oracle's finalize of the whole batch's exact integer sums; this rank's normalised shard equals the
oracle bit for bit; repeated calls (epoch / slot-parity handling of the peer exchange) keep giving the same result; the bare
all-reduce sums correctly.  same_gpu = 1 puts every rank on cuda:0 (gloo rendezvous, CUDA IPC between processes on one device),
which lets a 1-GPU box exercise the multi-rank exchange.  Exit code != 0 on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    transports, same_gpu = sys.argv[1].split(","), sys.argv[2] == "1"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = 0 if same_gpu else local
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")   # set-up plumbing only (ids / IPC handles); the data path never touches it
    import vacv_b200 as vacv
    from arm_neon_opencv_b200 import distributed as vd
    from oracle_lib import NCHW, NHWC, Oracle
    o = Oracle()
    n, h, w = 7, 90, 160
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)   # identical on every rank
    frames[:, :, :, 1] //= 2                                       # distinct channel statistics
    b, e = vd.shard_range(n, rank, world)
    whole_sums = np.stack([[int(frames[..., k].astype(np.uint64).sum()), int((frames[..., k].astype(np.uint64) ** 2).sum())]
                           for k in range(3)]).astype(np.uint64)
    m_want, s_want = o.finalize_mean_stddev(whole_sums.ravel(), 3, n * w * h)
    for name in transports:
        t = vd.NcclComm() if name == "nccl" else vd.P2PExchange()
        assert t.world == world
        for layout, lay_o in ((vacv.NHWC, NHWC), (vacv.NCHW, NCHW)):
            mine_np = frames[b:e] if layout == vacv.NHWC else np.ascontiguousarray(frames[b:e].transpose(0, 3, 1, 2))
            mine = torch.from_numpy(mine_np).cuda()
            for rep in range(3):   # odd and even epochs
                out, ms = vd.normalize_batch_global(t, mine, layout)
                torch.cuda.synchronize()
                ms = ms.cpu().numpy()
                assert np.array_equal(ms[0].view(np.uint32), m_want.view(np.uint32)), (name, rank, rep, ms, m_want)
                assert np.array_equal(ms[1].view(np.uint32), s_want.view(np.uint32)), (name, rank, rep, ms, s_want)
                got = out.cpu().numpy()
                for i in range(e - b):
                    want = o.normalize(mine_np[i], w * h, 3, lay_o, m_want, s_want)
                    assert np.array_equal(got[i].view(np.uint32).ravel(), want.view(np.uint32).ravel()), (name, rank, rep, i)
        # the bare exchange: values that differ per rank, including > 2^53 (exact integer arithmetic, not fp64)
        buf = torch.tensor([rank + 1, (1 << 60) + rank, 7], dtype=torch.int64, device="cuda")
        for rep in range(4):
            t.allreduce_u64(buf)
        torch.cuda.synchronize()
        want = [rank + 1, (1 << 60) + rank, 7]
        for rep in range(4):   # what four successive in-place sum all-reduces give
            tot = [sum(r + 1 for r in range(world)), sum((1 << 60) + r for r in range(world)), 7 * world]
            want = tot if rep == 0 else [(v * world) & 0xFFFFFFFFFFFFFFFF for v in want]
        got = [int(v) & 0xFFFFFFFFFFFFFFFF for v in buf.cpu().tolist()]
        assert got == want, (name, rank, got, want)
        if name == "p2p":
            assert t.timed_out() == 0, "peer exchange timed out"
        gathered = [None] * world
        dist.all_gather_object(gathered, ms.tobytes())
        assert all(g == gathered[0] for g in gathered), "statistics differ between ranks"
        dist.barrier()
        t.close()
    dist.barrier()
    if rank == 0:
        print("C5_DIST_OK", world, ",".join(transports))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
