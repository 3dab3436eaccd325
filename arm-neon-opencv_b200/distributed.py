"""Multi-GPU plumbing of the vacv path (one process per GPU, torch.distributed).

Frames are independent, so every operator shards the batch with NO data-path collective.  The single exchange on the
path is the batch-global statistic of config 5: per-rank exact integer sums -> one all-reduce of 2*c int64 values
(48 bytes for BGR) -> identical mean/stddev on every rank, independent of the number of ranks.
Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [begin, end) of n_items for `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_sums(sums):
    """In-place SUM all-reduce of the exact (sum x, sum x^2) counters (int64 tensor [sets, c, 2])."""
    if sums.dtype != torch.int64:
        raise TypeError("sums must be int64 (exact integer statistics)")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def finalize_mean_stddev_host(sums, n_per_channel):
    """Host mirror of vacv_cuda_finalize_mean_stddev (fp64 -> fp32); used where no GPU is present (CPU tests)."""
    s = sums.to(torch.float64)
    mean = s[..., 0] / n_per_channel
    var = (s[..., 1] / n_per_channel - mean * mean).clamp_min(0)
    return mean.to(torch.float32), var.sqrt().to(torch.float32)
