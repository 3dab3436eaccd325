"""Periodic bilinear walker: vertical segments per column strip (VACV_WALK_SEGS) on the shapes of bench_ops.   python profiles/_lin_segs.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vacv_b200 as vacv
from bench_ops import rand_u8, timeit
cases = [("chw 4K->1080p x16", (16, 3, 2160, 3840), vacv.NCHW, 1920, 1080), ("chw 1440p->1080p x32", (32, 3, 1440, 2560), vacv.NCHW, 1920, 1080),
         ("chw 1080p->720p x64", (64, 3, 1080, 1920), vacv.NCHW, 1280, 720), ("hwc 4K->1080p x16", (16, 2160, 3840, 3), vacv.NHWC, 1920, 1080),
         ("hwc 1080p->720p x64", (64, 1080, 1920, 3), vacv.NHWC, 1280, 720), ("hwc 1080p->720p x1", (1, 1080, 1920, 3), vacv.NHWC, 1280, 720)]
for name, shape, layout, wo, ho in cases:
    src = rand_u8(*shape)
    line = []
    for segs in (0, 4, 8, 16, 34, 68, 135):
        vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", segs)
        ms, mn = timeit(lambda: vacv.resize(src, layout, wo, ho), 20)
        line.append(f"{segs}: {ms:.4f}")
    vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", 0)
    print(name, "  ".join(line), flush=True)
    del src
