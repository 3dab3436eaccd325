"""Worker of tests/test_gpu_multi.py (launched under torchrun, one process per GPU, NCCL): config 5 on a sharded batch.
Every rank owns a contiguous shard of the frames, reduces it to exact integer sums on its GPU, the sums are all-reduced
(48 bytes), every rank finalises mean/stddev and normalises its shard.  Checks: the all-reduced sums equal the sums of the
whole batch computed on one GPU; mean/stddev are bit-identical on every rank; a frame of the shard equals the oracle's
normalise with the global statistics.  Exit code != 0 on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import vacv_b200 as vacv
    from arm_neon_opencv_b200 import distributed as vd
    from oracle_lib import NHWC, Oracle
    n, h, w = 13, 360, 640                                     # ragged shards on purpose
    frames = np.random.default_rng(7).integers(0, 256, (n, h, w, 3), dtype=np.uint8)   # same on every rank
    b, e = vd.shard_range(n, rank, world)
    mine = torch.from_numpy(frames[b:e]).cuda()
    sums = vacv.sums_u8(mine, vacv.NHWC, per_frame=False)
    vd.allreduce_sums(sums)
    whole = vacv.sums_u8(torch.from_numpy(frames).cuda(), vacv.NHWC, per_frame=False)
    assert torch.equal(sums, whole), "all-reduced shard sums != sums of the whole batch"
    mean, std = vacv.finalize_mean_stddev(sums, n * w * h)
    gathered = [torch.empty_like(mean) for _ in range(world)]
    dist.all_gather(gathered, mean)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "mean differs between ranks"
    out = vacv.normalize(mine, vacv.NHWC, mean[0], std[0])
    o = Oracle()
    m, s = o.finalize_mean_stddev(whole.cpu().numpy()[0].astype(np.uint64).ravel(), 3, n * w * h)
    want = o.normalize(frames[b], w * h, 3, NHWC, m, s)
    assert np.array_equal(out[0].cpu().numpy().view(np.uint32), want.view(np.uint32)), "normalised frame != oracle"
    dist.barrier()
    if rank == 0:
        print("NCCL_C5_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
