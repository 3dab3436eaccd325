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


# ------------------------------------------------------------------------------------------------------------------------
# Config 5 through the C-ABI (include/vacv_cuda.h, include/vacv_dist.h).  torch.distributed is used ONLY to ship the 128-byte
# NCCL id / the 64-byte IPC handles between the ranks at set-up time; the data path is one C call per step.
import ctypes as _C


def _pkg():
    import sys
    return sys.modules[__name__.rsplit(".", 1)[0]]


def _bytes_all_gather(payload, group=None):
    """Every rank contributes len(payload) bytes; returns the list of all ranks' payloads (works on gloo and nccl)."""
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, bytes(payload), group=group)
    return out


class NcclComm:
    """An ncclComm_t of this library's own (vacv_dist_nccl_comm_create), one rank per process on the current CUDA device."""

    def __init__(self, group=None):
        v = _pkg()
        self._lib = v.dist_lib()
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
        ident = _C.create_string_buffer(128)
        if rank == 0:
            v._check(self._lib.vacv_dist_nccl_unique_id(_C.cast(ident, _C.c_void_p)))
        if world > 1:
            ident = _C.create_string_buffer(_bytes_all_gather(ident.raw, group)[0], 128)
        self.handle = _C.c_void_p()
        v._check(self._lib.vacv_dist_nccl_comm_create(_C.byref(self.handle), world, rank, _C.cast(ident, _C.c_void_p)))
        self.rank, self.world = rank, world

    def allreduce_u64(self, buf):
        v = _pkg()
        v._check(self._lib.vacv_dist_allreduce_u64(self.handle, buf.data_ptr(), buf.numel(), v._stream()))
        return buf

    def close(self):
        if self.handle:
            _pkg()._check(self._lib.vacv_dist_nccl_comm_destroy(self.handle))
            self.handle = _C.c_void_p()


class P2PExchange:
    """The peer-memory exchange (vacv_cuda_p2p_*): slot buffers mapped into every rank of the node through CUDA IPC."""

    def __init__(self, group=None):
        v = _pkg()
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
        mine = _C.create_string_buffer(64)
        self.handle = _C.c_void_p()
        v._check(v.lib.vacv_cuda_p2p_create(_C.byref(self.handle), world, rank, _C.cast(mine, _C.c_void_p)))
        if world > 1:
            everyone = b"".join(_bytes_all_gather(mine.raw, group))
            v._check(v.lib.vacv_cuda_p2p_connect(self.handle, _C.cast(_C.create_string_buffer(everyone, 64 * world), _C.c_void_p)))
            dist.barrier(group)   # nobody publishes before every rank has mapped every buffer
        self.rank, self.world = rank, world

    def allreduce_u64(self, buf):
        v = _pkg()
        v._check(v.lib.vacv_cuda_p2p_allreduce_u64(self.handle, buf.data_ptr(), buf.numel(), v._stream()))
        return buf

    def timed_out(self):
        n = _C.c_int(0)
        v = _pkg()
        v._check(v.lib.vacv_cuda_p2p_status(self.handle, _C.byref(n)))
        return n.value

    def close(self):
        if self.handle:
            _pkg()._check(_pkg().lib.vacv_cuda_p2p_destroy(self.handle))
            self.handle = _C.c_void_p()


def normalize_batch_global(transport, src, layout, out=None, work=None, mean_std=None, ev_sums_done=None, ev_stats_ready=None):
    """Config 5 in one C call: src uint8 [B,h,w,c] (NHWC) / [B,c,h,w] (NCHW) = this rank's shard -> fp32 normalised with the
    statistics of ALL ranks' frames.  transport: NcclComm (vacv_cuda_normalize_batch_global) or P2PExchange (..._p2p).
    Returns (dst, mean_std) with mean_std = float32 [2, c] (mean row, stddev row)."""
    v = _pkg()
    src = v._dev(src, torch.uint8)
    if layout == v.NHWC:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    dst = out if out is not None else torch.empty(src.shape, dtype=torch.float32, device=src.device)
    work = work if work is not None else torch.empty(16, dtype=torch.int64, device=src.device)
    mean_std = mean_std if mean_std is not None else torch.empty((2, c), dtype=torch.float32, device=src.device)
    e0 = ev_sums_done.cuda_event if ev_sums_done is not None else None
    e1 = ev_stats_ready.cuda_event if ev_stats_ready is not None else None
    if isinstance(transport, NcclComm):
        fn = v.dist_lib().vacv_cuda_normalize_batch_global
    elif isinstance(transport, P2PExchange):
        fn = v.lib.vacv_cuda_normalize_batch_global_p2p
    else:
        raise TypeError("transport must be NcclComm or P2PExchange")
    v._check(fn(transport.handle, src.data_ptr(), dst.data_ptr(), b, w, h, c, layout, work.data_ptr(), mean_std.data_ptr(), e0, e1,
                v._stream()))
    return dst, mean_std
