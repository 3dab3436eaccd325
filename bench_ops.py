#!/usr/bin/env python
"""bench_ops.py -- per-operator / per-config measurements next to the headline bench.py (not the driver contract).

    python bench_ops.py [--workload all|cpu|e2e|c1|c2|c2g|c2s|c3|c4|c5|ops] [--iters N] [--json out.jsonl]
    torchrun --nproc-per-node N bench_ops.py --workload c5        # batch-global statistics with the all-reduce

For every kernel: device-resident time per launch (CUDA events, median of N after warm-up, batch >> L2), output
Mpix/s, achieved GB/s = algorithmic bytes / time, fraction of the measured HBM copy peak.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import vacv_b200 as vacv  # noqa: E402
from bench import MEAN, STD, measured_peak  # noqa: E402

PEAK, _ = measured_peak()
RESULTS = []


GATE_CYCLES = 400_000   # ~0.2 ms of torch.cuda._sleep ahead of every sample
INNER = 4               # back-to-back launches per sample


def timeit(fn, iters, warmup=5):
    """Median / min DEVICE time of one call of fn, in ms.  A sample = INNER back-to-back calls between two CUDA events on the launching
    stream, queued while the GPU still spins in a gate kernel -- so the events bracket the launches' execution only.  (Until round 2
    a sample was one call issued to an idle GPU: the host's 10-25 us between recording the first event and the launch were inside
    every figure, which matters for the operators that take 0.05-0.15 ms.)"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(GATE_CYCLES)
        e0.record()
        for _ in range(INNER):
            fn()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1) / INNER)
    return statistics.median(times), min(times)


def report(name, ms, out_pix, algo_bytes, note=""):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    r = {"name": name, "ms": round(ms, 4), "out_Mpix_s": round(out_pix / (ms * 1e-3) / 1e6, 1), "GB_s": round(gbs, 1),
         "frac_of_measured_peak": round(gbs / PEAK, 3), "algo_bytes": algo_bytes, "note": note}
    RESULTS.append(r)
    print(f"{name:58s} {ms:9.4f} ms {r['out_Mpix_s']:12.1f} Mpix/s {gbs:8.1f} GB/s  frac {r['frac_of_measured_peak']:.3f}  {note}", flush=True)


def rand_u8(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)


def stats():
    return torch.tensor(MEAN, device="cuda"), torch.tensor(STD, device="cuda")


def c1(iters):   # resize INTER_LINEAR u8 BGR 1920x1080 -> 640x360, batch 256
    b = 256
    src = rand_u8(b, 1080, 1920, 3)
    ms, _ = timeit(lambda: vacv.resize(src, vacv.NHWC, 640, 360), iters)
    report("c1 resize linear u8 hwc 1920x1080->640x360 x256", ms, b * 640 * 360, b * (1920 * 1080 * 3 + 640 * 360 * 3),
           "SURVEY 8d bytes (whole source frame); the kernel must touch one source row in three -> next line")
    # bytes the operator MUST touch: ratio 3.0 samples source row 3*dy+1 only (fractional weight 0), and of that row the pixel
    # 3*dx+1 of every 3 -- whole 32-byte sectors of the sampled rows are fetched, i.e. the full row: 360 rows x 5760 B + output
    report("   same launch, touched bytes (360 sampled rows, whole sectors)", ms, b * 640 * 360, b * (360 * 1920 * 3 + 640 * 360 * 3),
           "the roofline that applies: frac must be <= 1")
    ms, _ = timeit(lambda: vacv.resize(src, vacv.NHWC, 500, 300), iters)
    report("   resize linear u8 hwc 1920x1080->500x300 x256", ms, b * 500 * 300, b * (1920 * 1080 * 3 + 500 * 300 * 3))


def c2(iters, wo=640, ho=640, w=1920, h=1080, tag="c2"):
    b = 256
    mean, std = stats()
    src = rand_u8(b, w * h * 3 // 2)
    out = torch.empty((b, 3, ho, wo), dtype=torch.float32, device="cuda")
    ms, _ = timeit(lambda: vacv.nv_resize_normalize_chw(src, w, h, wo, ho, mean, std, True, out=out), iters)
    report(f"{tag} fused nv12->chw f32 {w}x{h}->{wo}x{ho} x{b}", ms, b * wo * ho, b * (w * h * 3 // 2 + wo * ho * 12))


def c2_surfaces(iters):
    """Next rows 8f-1 / 8f-3: pitched and planar surfaces in, fp32 / fp16 planes out (same shape as config 2)."""
    b, w, h, wo, ho = 256, 1920, 1080, 640, 640
    mean, std = stats()
    for name, fmt, yp, half in [("nv12 dense  -> f32", vacv.YUV_NV12, 1920, False), ("nv12 p2048  -> f32", vacv.YUV_NV12, 2048, False),
                                ("i420 dense  -> f32", vacv.YUV_I420, 1920, False), ("nv12 dense  -> f16", vacv.YUV_NV12, 1920, True),
                                ("i420 p2048  -> f16", vacv.YUV_I420, 2048, True)]:
        cp = yp // 2 if fmt >= vacv.YUV_I420 else yp
        per = yp * h * 3 // 2
        src = rand_u8(b * per)
        out = torch.empty((b, 3, ho, wo), dtype=torch.float16 if half else torch.float32, device="cuda")
        ms, _ = timeit(lambda: vacv.yuv_resize_normalize_chw(src, fmt, w, h, wo, ho, mean, std, y_pitch=yp, c_pitch=cp, batch=b,
                                                             half=half, out=out), iters)
        report(f"c2s fused {name} 1080p->640x640 x{b}", ms, b * wo * ho, b * (w * h * 3 // 2 + wo * ho * (6 if half else 12)))
        if yp != w:   # A/B of the three staging forms on padded surfaces: tensor-map boxes (default), one bulk copy per row, whole bands
            for mode, what in ((1, "one bulk copy per row, padding not read (PIPE_ROWS=1)"), (2, "whole bands, padding read (PIPE_ROWS=2, the round-1 path)")):
                vacv.lib.vacv_cuda_set_tuning(b"PIPE_ROWS", mode)
                ms, _ = timeit(lambda: vacv.yuv_resize_normalize_chw(src, fmt, w, h, wo, ho, mean, std, y_pitch=yp, c_pitch=cp, batch=b,
                                                                     half=half, out=out), iters)
                vacv.lib.vacv_cuda_set_tuning(b"PIPE_ROWS", 0)
                report(f"   same, {what}", ms, b * wo * ho, b * (w * h * 3 // 2 + wo * ho * (6 if half else 12)))
    src = rand_u8(b * w * h * 3 // 2)
    for name, dt, tdt, eb in [("f32", vacv.FP32, torch.float32, 4), ("bf16", vacv.BF16, torch.bfloat16, 2)]:
        out = torch.empty((b, 3, ho, wo), dtype=tdt, device="cuda")
        ms, _ = timeit(lambda: vacv.yuv_letterbox_normalize_chw(src, vacv.YUV_NV12, w, h, wo, ho, [104., 117., 123.], [58., 57., 57.], batch=b,
                                                                out_dtype=dt, out=out), iters)
        report(f"c2s letterbox nv12 -> {name} 1080p->640x360 in 640x640 x{b}", ms, b * wo * ho, b * (w * h * 3 // 2 + wo * ho * 3 * eb))


def c2_unfused(iters):
    b = 64
    mean, std = stats()
    src = rand_u8(b, 1920 * 1080 * 3 // 2)

    def chain():
        bgr = vacv.cvt_nv2bgr(src, 1920, 1080)
        small = vacv.resize(bgr, vacv.NHWC, 640, 640)
        norm = vacv.normalize(small, vacv.NHWC, mean, std)
        return vacv.layout_change(norm, vacv.NHWC, vacv.NCHW)
    ms, _ = timeit(chain, iters)
    report("   unfused CUDA chain (4 kernels) 1080p->640x640 x64", ms, b * 640 * 640, b * (1920 * 1080 * 3 // 2 + 640 * 640 * 12),
           "algo bytes of the fused op")


def face_matrices(n, w, h, wo, seed=7):
    r = np.random.default_rng(seed)
    ms = []
    for _ in range(n):
        s = r.uniform(0.3, 0.6)
        a = np.deg2rad(r.uniform(-15, 15))
        cx, cy = r.uniform(0.25 * w, 0.75 * w), r.uniform(0.3 * h, 0.7 * h)
        al, be = s * np.cos(a), s * np.sin(a)
        ms.append(vacv.invert_affine([al, be, wo / 2 - al * cx - be * cy, -be, al, wo / 2 + be * cx - al * cy]))
    return torch.tensor(np.array(ms, np.float32), device="cuda"), r


def c3(iters):   # warp_affine face crops: 4096 x (1280x720 -> 112x112 fp32 normalised), frame pool of 512
    n, nf, w, h, wo = 4096, 512, 1280, 720, 112
    mean, std = stats()
    frames = rand_u8(nf, h, w, 3)
    minv, _ = face_matrices(n, w, h, wo)
    idx = (torch.arange(n, device="cuda") % nf).to(torch.int32)
    roi = int(np.mean([(wo / s) ** 2 * 3 for s in np.random.default_rng(7).uniform(0.3, 0.6, 1000)]))
    ms, _ = timeit(lambda: vacv.warp_affine_normalize(frames, minv, wo, wo, mean, std, idx), iters)
    report("c3 warp_affine_normalize 1280x720->112x112 f32 x4096", ms, n * wo * wo, n * (roi + wo * wo * 12), "bytes = mean source ROI + out")
    vacv.lib.vacv_cuda_set_tuning(b"WARP_GATHER", 1)    # A/B: the direct gather kernel instead of the TMA-staged default
    ms, _ = timeit(lambda: vacv.warp_affine_normalize(frames, minv, wo, wo, mean, std, idx), iters)
    vacv.lib.vacv_cuda_set_tuning(b"WARP_GATHER", 0)
    report("   same, direct gather kernel (WARP_GATHER=1)", ms, n * wo * wo, n * (roi + wo * wo * 12))
    ms, _ = timeit(lambda: vacv.warp_affine(frames, vacv.NHWC, minv, wo, wo, idx), iters)
    report("   warp_affine u8 1280x720->112x112 x4096", ms, n * wo * wo, n * (roi + wo * wo * 3))
    for v, what in ((2, "pack kernel with two pixels per thread in flight (WARP_V=2)"), (1, "first-generation gather kernel (WARP_V=1)")):   # A/B
        vacv.lib.vacv_cuda_set_tuning(b"WARP_V", v)
        ms, _ = timeit(lambda: vacv.warp_affine(frames, vacv.NHWC, minv, wo, wo, idx), iters)
        vacv.lib.vacv_cuda_set_tuning(b"WARP_V", 0)
        report(f"   warp_affine u8, {what}", ms, n * wo * wo, n * (roi + wo * wo * 3))
    ms, _ = timeit(lambda: vacv.warp_affine(frames, vacv.NHWC, minv, wo, wo, idx, vacv.FLAG_TILED), iters)
    report("   warp_affine u8 (TMA-staged kernel, opt-in) x4096", ms, n * wo * wo, n * (roi + wo * wo * 3))
    ms, _ = timeit(lambda: vacv.warp_affine_normalize(frames, minv, wo, wo, mean, std, idx, out_layout=vacv.NCHW), iters)
    report("   warp_affine_normalize -> CHW planes x4096", ms, n * wo * wo, n * (roi + wo * wo * 12))


def c4(iters):   # resize INTER_CUBIC u8 2560x1440 -> 1920x1080, batch 128
    b = 128
    src = rand_u8(b, 1440, 2560, 3)
    ms, _ = timeit(lambda: vacv.resize(src, vacv.NHWC, 1920, 1080, vacv.INTER_CUBIC), iters)
    report("c4 resize cubic u8 hwc 2560x1440->1920x1080 x128", ms, b * 1920 * 1080, b * (2560 * 1440 * 3 + 1920 * 1080 * 3))
    for v, name in ((1, "first-generation walker, 2 columns / thread (CUBIC_V=1)"), (2, "second generation, 2 columns / thread (CUBIC_V=2)")):
        vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", v)
        ms, _ = timeit(lambda: vacv.resize(src, vacv.NHWC, 1920, 1080, vacv.INTER_CUBIC), iters)
        vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 0)
        report(f"   same, {name}", ms, b * 1920 * 1080, b * (2560 * 1440 * 3 + 1920 * 1080 * 3))
    big = rand_u8(64, 1080, 1920, 3)
    ms, _ = timeit(lambda: vacv.resize(big, vacv.NHWC, 1280, 720, vacv.INTER_CUBIC), iters)
    report("   resize cubic u8 hwc 1920x1080->1280x720 x64", ms, 64 * 1280 * 720, 64 * (1920 * 1080 * 3 + 1280 * 720 * 3), "3 : 2, periodic walker with realignment")
    uhd = rand_u8(32, 2160, 3840, 3)
    ms, _ = timeit(lambda: vacv.resize(uhd, vacv.NHWC, 1920, 1080, vacv.INTER_CUBIC), iters)
    report("   resize cubic u8 hwc 3840x2160->1920x1080 x32", ms, 32 * 1920 * 1080, 32 * (3840 * 2160 * 3 + 1920 * 1080 * 3), "2 : 1, periodic walker")
    del uhd
    srcf = src[:32].to(torch.float32)
    ms, _ = timeit(lambda: vacv.resize(srcf, vacv.NHWC, 1920, 1080, vacv.INTER_CUBIC), iters)
    report("   resize cubic f32 hwc 2560x1440->1920x1080 x32", ms, 32 * 1920 * 1080, 32 * 4 * (2560 * 1440 * 3 + 1920 * 1080 * 3))


def c5(iters):   # batch-global mean/stddev + normalize, 4K frames, 128 per GPU: ONE C call per step, NCCL and peer-memory transports
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    dist = None
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from bench import C5_B, C5_H, C5_W, c5_measure

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    res = c5_measure(torch, vacv, dev, rank, world, iters, max_over_ranks, barrier)
    b, w, h = C5_B, C5_W, C5_H
    if rank == 0:
        for name, t in res["transports"].items():
            report(f"c5 normalize_batch_global[{name}] 4K x{b}/GPU, {world} GPU(s)", t["ms_per_step"], world * b * w * h, world * b * w * h * 18,
                   f"per-GPU frac {t['per_gpu_frac']:.3f}; exchange gap median {t['gap_us_median']} us (min {t['gap_us_min']} us)")
        RESULTS.append({"name": "c5 object", **res})
    if world == 1:
        frames = rand_u8(b, h, w, 3)
        out = torch.empty((b, h, w, 3), dtype=torch.float32, device="cuda")
        sums = torch.zeros((1, 3, 2), dtype=torch.int64, device="cuda")
        ms, _ = timeit(lambda: vacv.sums_u8(frames, vacv.NHWC, False, sums), iters)
        report("   sums_u8 (stats pass) 4K x128", ms, b * w * h, b * w * h * 3)
        mean, std = stats()
        ms, _ = timeit(lambda: vacv.normalize(frames, vacv.NHWC, mean, std, out=out), iters)
        report("   normalize u8->f32 (apply pass) 4K x128", ms, b * w * h, b * w * h * 3 * 5)
    if dist:
        dist.destroy_process_group()


def pcie(iters):
    """The platform ceiling of the end-to-end path: the chunked cudaMemcpyAsync pattern of run_host_pipeline
    (csrc/vacv_host.cu) -- H2D of NV12 chunks and D2H of fp32 plane chunks on two streams -- with NO kernels, on every rank at once;
    H2D alone and D2H alone; pinned (default) vs write-combined host source.  Run under torchrun for N GPUs."""
    import ctypes as C
    import time
    from bench import BATCH, HO, IN_FRAME, WO, bind_to_gpu_numa, pcie_ceiling
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    numa = bind_to_gpu_numa(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    h_in = torch.randint(0, 256, (BATCH, IN_FRAME), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((BATCH, 3, HO, WO), dtype=torch.float32).pin_memory()
    out = {"name": "pcie", "n_gpus": world, "host_binding": numa, "chunks": {}}
    for chunk in (8, 32):
        out["chunks"][chunk] = pcie_ceiling(torch, dev, h_in, h_out, chunk, max(3, iters // 4), barrier, max_over_ranks)
    # write-combined source for the H2D direction (cudaHostAllocWriteCombined through the C-ABI)
    p = C.c_void_p()
    nbytes = BATCH * IN_FRAME
    if vacv.lib.vacv_cuda_host_alloc_flags(C.byref(p), nbytes, 1) == 0:
        d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        C.memset(p, 7, nbytes)
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            vacv.lib.vacv_cuda_memcpy_h2d(d.data_ptr(), p, nbytes, s)
        barrier()
        t0 = time.perf_counter()
        n = max(3, iters // 4)
        for _ in range(n):
            vacv.lib.vacv_cuda_memcpy_h2d(d.data_ptr(), p, nbytes, s)
        torch.cuda.synchronize()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n
        out["h2d_write_combined_gbs_per_gpu"] = round(nbytes / (ms * 1e-3) / 1e9, 2)
        vacv.lib.vacv_cuda_host_free(p)
    # the real e2e call beside it
    run = lambda: vacv.nv_resize_normalize_chw_host(h_in, h_out, 1920, 1080, WO, HO, MEAN, STD, True, 8)
    run()
    barrier()
    t0 = time.perf_counter()
    n = max(3, iters // 4)
    for _ in range(n):
        run()
    ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n
    out["e2e_ms"] = round(ms, 3)
    out["e2e_gbs_per_gpu"] = round(BATCH * (IN_FRAME + 3 * HO * WO * 4) / (ms * 1e-3) / 1e9, 2)
    out["e2e_frac_of_ceiling"] = round(out["e2e_gbs_per_gpu"] / out["chunks"][8]["h2d+d2h"]["gbs_per_gpu"], 3)
    if rank == 0:
        RESULTS.append(out)
        print(json.dumps(out), flush=True)
    if dist:
        dist.destroy_process_group()


def ops(iters):
    b = 128
    nv = rand_u8(b, 1920 * 1080 * 3 // 2)
    ms, _ = timeit(lambda: vacv.cvt_nv2bgr(nv, 1920, 1080), iters)
    report("op cvt_nv2bgr 1080p x128", ms, b * 1920 * 1080, b * 1920 * 1080 * 4.5)
    bgr = rand_u8(b, 1080, 1920, 3)
    ms, _ = timeit(lambda: vacv.layout_change(bgr, vacv.NHWC, vacv.NCHW), iters)
    report("op layout hwc->chw u8 1080p x128", ms, b * 1920 * 1080, b * 1920 * 1080 * 6)
    chw = vacv.layout_change(bgr, vacv.NHWC, vacv.NCHW)
    ms, _ = timeit(lambda: vacv.layout_change(chw, vacv.NCHW, vacv.NHWC), iters)
    report("op layout chw->hwc u8 1080p x128", ms, b * 1920 * 1080, b * 1920 * 1080 * 6)
    f = torch.rand((32, 1080, 1920, 3), device="cuda") * 255
    ms, _ = timeit(lambda: vacv.layout_change(f, vacv.NHWC, vacv.NCHW), iters)
    report("op layout hwc->chw f32 1080p x32", ms, 32 * 1920 * 1080, 32 * 1920 * 1080 * 24)
    ms, _ = timeit(lambda: vacv.dtype_change(bgr, vacv.FP32), iters)
    report("op dtype u8->f32 1080p x128", ms, b * 1920 * 1080, b * 1920 * 1080 * 15)
    ms, _ = timeit(lambda: vacv.dtype_change(f, vacv.INT8), iters)
    report("op dtype f32->u8 1080p x32", ms, 32 * 1920 * 1080, 32 * 1920 * 1080 * 15)
    mean, std = stats()
    ms, _ = timeit(lambda: vacv.normalize(f, vacv.NHWC, mean, std), iters)
    report("op normalize f32 hwc 1080p x32", ms, 32 * 1920 * 1080, 32 * 1920 * 1080 * 24)
    ms, _ = timeit(lambda: vacv.crop(bgr, vacv.NHWC, 321, 181, 1280, 720), iters)
    report("op crop u8 hwc 1080p->1280x720 @(321,181) x128", ms, b * 1280 * 720, b * 1280 * 720 * 6)
    big = rand_u8(64, 1440, 2560, 3)
    ms, _ = timeit(lambda: vacv.resize(big, vacv.NHWC, 320, 180), iters)
    report("op resize linear u8 2560x1440->320x180 x64", ms, 64 * 320 * 180, 64 * (2560 * 1440 * 3 // 4 + 320 * 180 * 3), "bytes: 2 of 8 rows touched")
    ms, _ = timeit(lambda: vacv.resize_normalize(bgr, 640, 640, mean, std, vacv.NCHW), iters)
    report("op resize_normalize u8 hwc 1080p->640x640 chw x128", ms, b * 640 * 640, b * (1920 * 1080 * 3 + 640 * 640 * 12))
    ms, _ = timeit(lambda: vacv.resize_normalize(bgr, 640, 640, mean, std, vacv.NHWC), iters)
    report("op resize_normalize u8 hwc 1080p->640x640 hwc x128", ms, b * 640 * 640, b * (1920 * 1080 * 3 + 640 * 640 * 12))


def ops2(iters):   # shapes off the headline path: planar layouts, fp32 gathers, generic channel counts
    b = 64
    bgr = rand_u8(b, 1080, 1920, 3)
    chw = vacv.layout_change(bgr, vacv.NHWC, vacv.NCHW)
    f = bgr[:16].to(torch.float32)
    fchw = chw[:16].to(torch.float32)
    mean, std = stats()
    ms, _ = timeit(lambda: vacv.resize(chw, vacv.NCHW, 640, 360), iters)
    report("op2 resize linear u8 chw 1080p->640x360 x64", ms, b * 640 * 360, b * (1920 * 1080 * 3 * 2 // 3 + 640 * 360 * 3), "2 of 3 rows touched")
    for (w, h, wo, ho, n, what) in ((1920, 1080, 1280, 720, 64, "3 : 2"), (2560, 1440, 1920, 1080, 32, "4 : 3"), (3840, 2160, 1920, 1080, 16, "2 : 1")):
        planes = chw if w == 1920 else rand_u8(n, 3, h, w)
        for v in (0, 1):
            vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", v)
            ms, _ = timeit(lambda: vacv.resize(planes, vacv.NCHW, wo, ho), iters)
            vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 0)
            report(f"op2 resize linear u8 chw {w}x{h}->{wo}x{ho} x{n}" if v == 0 else "   same, planes pipeline / gather kernel (LINEAR_V=1)", ms, n * wo * ho,
                   n * 3 * (w * h + wo * ho), f"{what}, periodic walker on planes" if v == 0 else "")
        del planes
    ms, _ = timeit(lambda: vacv.resize(f, vacv.NHWC, 640, 360), iters)
    report("op2 resize linear f32 hwc 1080p->640x360 x16", ms, 16 * 640 * 360, 16 * 4 * (1920 * 1080 * 3 * 2 // 3 + 640 * 360 * 3), "2 of 3 rows touched")
    ms, _ = timeit(lambda: vacv.resize(f, vacv.NHWC, 1280, 720), iters)
    report("op2 resize linear f32 hwc 1080p->1280x720 x16", ms, 16 * 1280 * 720, 16 * 4 * (1920 * 1080 * 3 + 1280 * 720 * 3))
    ms, _ = timeit(lambda: vacv.resize(bgr, vacv.NHWC, 1280, 720), iters)
    report("op2 resize linear u8 hwc 1080p->1280x720 x64", ms, b * 1280 * 720, b * (1920 * 1080 * 3 + 1280 * 720 * 3), "3 : 2, periodic walker")
    vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 1)   # A/B: the persistent pipeline
    ms, _ = timeit(lambda: vacv.resize(bgr, vacv.NHWC, 1280, 720), iters)
    vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 0)
    report("   same, persistent pipeline (LINEAR_V=1)", ms, b * 1280 * 720, b * (1920 * 1080 * 3 + 1280 * 720 * 3))
    for (w, h, wo, ho, n, what) in ((2560, 1440, 1920, 1080, 32, "4 : 3"), (3840, 2160, 1920, 1080, 16, "2 : 1"), (1920, 1080, 768, 432, 64, "5 : 2"),
                                    (1920, 1080, 1152, 648, 64, "5 : 3")):
        big = bgr if w == 1920 else rand_u8(n, h, w, 3)
        for v in (0, 1):
            vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", v)
            ms, _ = timeit(lambda: vacv.resize(big, vacv.NHWC, wo, ho), iters)
            vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 0)
            report(f"op2 resize linear u8 hwc {w}x{h}->{wo}x{ho} x{n}" if v == 0 else "   same, persistent pipeline (LINEAR_V=1)", ms, n * wo * ho,
                   n * (w * h * 3 + wo * ho * 3), f"{what}, periodic walker" if v == 0 else "")
        if big is not bgr:
            del big
    ms, _ = timeit(lambda: vacv.resize(chw, vacv.NCHW, 1280, 720, vacv.INTER_CUBIC) if False else vacv.resize(fchw, vacv.NCHW, 1280, 720, vacv.INTER_CUBIC), iters)
    report("op2 resize cubic f32 chw 1080p->1280x720 x16", ms, 16 * 1280 * 720, 16 * 4 * (1920 * 1080 * 3 + 1280 * 720 * 3), "3 : 2, periodic walker")
    vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 1)   # A/B: the one-column walker
    ms, _ = timeit(lambda: vacv.resize(fchw, vacv.NCHW, 1280, 720, vacv.INTER_CUBIC), iters)
    vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 0)
    report("   same, one-column walker (CUBIC_V=1)", ms, 16 * 1280 * 720, 16 * 4 * (1920 * 1080 * 3 + 1280 * 720 * 3))
    qhd = rand_u8(8, 3, 1440, 2560).to(torch.float32)
    ms, _ = timeit(lambda: vacv.resize(qhd, vacv.NCHW, 1920, 1080, vacv.INTER_CUBIC), iters)
    report("op2 resize cubic f32 chw 2560x1440->1920x1080 x8", ms, 8 * 1920 * 1080, 8 * 4 * 3 * (2560 * 1440 + 1920 * 1080), "4 : 3, periodic walker")
    del qhd
    ms, _ = timeit(lambda: vacv.normalize(chw, vacv.NCHW, mean, std), iters)
    report("op2 normalize u8 chw 1080p x64", ms, b * 1920 * 1080, b * 1920 * 1080 * 15)
    ms, _ = timeit(lambda: vacv.crop(chw, vacv.NCHW, 321, 181, 1280, 720), iters)
    report("op2 crop u8 chw 1080p->1280x720 x64", ms, b * 1280 * 720, b * 1280 * 720 * 6)
    ms, _ = timeit(lambda: vacv.crop(f, vacv.NHWC, 321, 181, 1280, 720), iters)
    report("op2 crop f32 hwc 1080p->1280x720 x16", ms, 16 * 1280 * 720, 16 * 1280 * 720 * 24)
    n, nf = 1024, 64
    frames = bgr[:nf].to(torch.float32)
    minv, _ = face_matrices(n, 1920, 1080, 112)
    idx = (torch.arange(n, device="cuda") % nf).to(torch.int32)
    ms, _ = timeit(lambda: vacv.warp_affine(frames, vacv.NHWC, minv, 112, 112, idx), iters)
    report("op2 warp_affine f32 hwc 1080p->112x112 x1024", ms, n * 112 * 112, n * (250000 * 4 + 112 * 112 * 12))
    grey = rand_u8(nf, 1080, 1920, 1)
    ms, _ = timeit(lambda: vacv.warp_affine(grey, vacv.NHWC, minv, 112, 112, idx), iters)
    report("op2 warp_affine u8 grey 1080p->112x112 x1024", ms, n * 112 * 112, n * (83000 + 112 * 112))
    rgba = rand_u8(b, 1080, 1920, 4)
    ms, _ = timeit(lambda: vacv.layout_change(rgba, vacv.NHWC, vacv.NCHW), iters)
    report("op2 layout hwc->chw u8 c=4 1080p x64", ms, b * 1920 * 1080, b * 1920 * 1080 * 8)


def e2e(iters):
    """Host buffers in, host buffers out (pinned), through the chunked three-stream pipeline: the config-2 shape with fp32 and
    fp16 planes.  Wall clock; PCIe-bound (the fp16 planes halve the D2H bytes)."""
    import time
    b, w, h, wo, ho = 256, 1920, 1080, 640, 640
    h_in = torch.randint(0, 256, (b, w * h * 3 // 2), dtype=torch.uint8).pin_memory()
    for name, dt, tdt, eb in [("fp32", vacv.FP32, torch.float32, 4), ("fp16", vacv.FP16, torch.float16, 2)]:
        h_out = torch.empty((b, 3, ho, wo), dtype=tdt).pin_memory()
        run = lambda: vacv.yuv_normalize_chw_host(h_in, h_out, vacv.YUV_NV12, w, h, wo, ho, MEAN, STD, batch=b, out_dtype=dt, chunk_frames=8)
        for _ in range(2):
            run()
        n = max(3, iters // 4)
        t0 = time.perf_counter()
        for _ in range(n):
            run()
        ms = (time.perf_counter() - t0) / n * 1e3
        report(f"e2e host nv12 -> host {name} planes 1080p->640x640 x{b}", ms, b * wo * ho, b * (w * h * 3 // 2 + wo * ho * 3 * eb), "PCIe bytes (H2D + D2H)")


def cpu_reference():
    """SURVEY 8(d) 'CPU reference alongside': the compiled reference (oracle/_ref, test infrastructure) timed on the box's host
    cores for every config -- (i) one thread, as the reference runs, (ii) frames spread over all host cores.  Bounded samples."""
    import ctypes as C
    import time
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    from oracle_lib import COLOR_YUV2BGR_NV21, Ref, ref_available
    if not ref_available():
        print("cpu: oracle/_ref not staged (make -C oracle ref)")
        return
    r = Ref()
    cores = len(os.sched_getaffinity(0))
    rng = np.random.default_rng(0)
    mean, std = np.array(MEAN, np.float32), np.array(STD, np.float32)

    r.lib.ref_time_config.restype = C.c_double
    r.lib.ref_time_config.argtypes = [C.c_int, C.c_int, C.c_int]

    def rate(name, cfg, out_pix, reps):
        one = reps * out_pix / r.lib.ref_time_config(cfg, 1, reps) / 1e6
        allc = cores * reps * out_pix / r.lib.ref_time_config(cfg, cores, reps) / 1e6
        rec = {"name": f"cpu {name}", "Mpix_s_1_thread": round(one, 1), f"Mpix_s_{cores}_threads": round(allc, 1), "cores": cores,
               "kind": "reference (oracle/_ref): C++ threads in the harness, one frame stream per thread", "frames_per_thread": reps}
        RESULTS.append(rec)
        print(f"cpu {name:58s} 1 thread {one:8.1f} Mpix/s   {cores} threads {allc:9.1f} Mpix/s", flush=True)

    rate("c1 resize linear u8 1080p->640x360", 1, 640 * 360, 100)
    rate("c3 warp_affine 720p->112x112 + f32 + normalize", 3, 112 * 112, 1500)
    rate("c4 cv::resize u8 cubic 1440p->1080p (bundled OpenCV 2.4)", 4, 1920 * 1080, 20)
    rate("c5 normalize u8 4K, own mean/stddev", 5, 3840 * 2160, 8)
    nv = rng.integers(0, 256, (cores, 1920 * 1080 * 3 // 2), dtype=np.uint8)
    t0 = time.perf_counter(); r.pipeline(nv[:1], 1920, 1080, COLOR_YUV2BGR_NV21, 640, 640, mean, std, batch=1, threads=1); t1 = time.perf_counter()
    for _ in range(3):
        r.pipeline(nv, 1920, 1080, COLOR_YUV2BGR_NV21, 640, 640, mean, std, batch=cores, threads=cores)
    t2 = time.perf_counter()
    one, allc = 640 * 640 / (t1 - t0) / 1e6, 3 * cores * 640 * 640 / (t2 - t1) / 1e6
    RESULTS.append({"name": "cpu c2 chain", "Mpix_s_1_thread": round(one, 1), f"Mpix_s_{cores}_threads": round(allc, 1), "cores": cores})
    print(f"cpu {'c2 chain nv21->bgr->640x640->f32->normalize->chw':58s} 1 thread {one:8.1f} Mpix/s   {cores} threads {allc:9.1f} Mpix/s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="all")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    wl = args.workload
    if wl == "cpu":   # host cores only
        cpu_reference()
        if args.json:
            with open(args.json, "w") as f:
                for r in RESULTS:
                    f.write(json.dumps(r) + "\n")
        return
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if wl in ("all", "e2e"):
        e2e(args.iters)
    if wl in ("all", "c2"):
        c2(args.iters)
    if wl in ("all", "c2g"):   # general (non-integer-ratio) shapes of the fused kernel
        c2(args.iters, 608, 608, tag="c2g")
        c2(args.iters, 640, 640, 1280, 720, tag="c2g")
        c2(args.iters, 416, 416, tag="c2g")
        c2_unfused(args.iters)
    if wl in ("all", "c2s"):
        c2_surfaces(args.iters)
    if wl in ("all", "c1"):
        c1(args.iters)
    if wl in ("all", "c3"):
        c3(args.iters)
    if wl in ("all", "c4"):
        c4(args.iters)
    if wl in ("all", "ops"):
        ops(args.iters)
    if wl in ("ops2",):
        ops2(args.iters)
    if wl in ("all", "c5"):
        c5(max(5, args.iters // 2))
    if wl in ("pcie",):
        pcie(args.iters)
    if args.json and int(os.environ.get("RANK", 0)) == 0:
        with open(args.json, "a") as f:
            for r in RESULTS:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
