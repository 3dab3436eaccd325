"""Periodic / generic bicubic walkers on small batches: vertical segments per column strip (VACV_WALK_SEGS).   python profiles/_cubic_segs.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vacv_b200 as vacv
from bench_ops import rand_u8, timeit
cases = [("u8 hwc 1440p->1080p", (2560, 1440), (1920, 1080), torch.uint8), ("u8 hwc 1080p->720p", (1920, 1080), (1280, 720), torch.uint8),
         ("u8 hwc 1080p->1000x700", (1920, 1080), (1000, 700), torch.uint8), ("f32 hwc 1440p->1080p", (2560, 1440), (1920, 1080), torch.float32)]
for name, (w, h), (wo, ho), dt in cases:
    for b in (1, 4, 16):
        src = rand_u8(b, h, w, 3).to(dt)
        line = []
        for segs in (0, 8, 17, 34, 68, 135):
            vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", segs)
            ms, mn = timeit(lambda: vacv.resize(src, vacv.NHWC, wo, ho, vacv.INTER_CUBIC), 15)
            line.append(f"{segs}: {ms:.4f}")
        vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", 0)
        print(f"{name} x{b}", "  ".join(line), flush=True)
        del src
