"""Sweep of the columns per thread of the u8 BGR bilinear pipeline (tuning knob RPIPE_NCOL) on a few shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vacv_b200 as vacv
from bench_ops import rand_u8, timeit, stats
mean, std = stats()
for (w, h, wo, ho, b) in ((1920, 1080, 1280, 720, 64), (1920, 1080, 640, 360, 256), (1920, 1080, 960, 540, 64), (1280, 720, 640, 360, 128)):
    src = rand_u8(b, h, w, 3)
    for v in (0, 2, 3, 4, 0):
        vacv.lib.vacv_cuda_set_tuning(b"RPIPE_NCOL", v)
        ms, mn = timeit(lambda: vacv.resize(src, vacv.NHWC, wo, ho), 30)
        print(f"resize {w}x{h}->{wo}x{ho} x{b} RPIPE_NCOL={v}: {ms:.4f} ms (min {mn:.4f})", flush=True)
        if (wo, ho) == (640, 360) or (wo, ho) == (1280, 720):
            ms, mn = timeit(lambda: vacv.resize_normalize(src, wo, ho, mean, std, vacv.NCHW), 30) if hasattr(vacv, "resize_normalize") else (0, 0)
            print(f"   resize_normalize chw RPIPE_NCOL={v}: {ms:.4f} ms", flush=True)
    del src
vacv.lib.vacv_cuda_set_tuning(b"RPIPE_NCOL", 0)
