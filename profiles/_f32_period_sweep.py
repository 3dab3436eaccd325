"""fp32 planes bicubic at 1080p -> 1280x720: periodic walker vs the one-column walker (CUBIC_V=1), sweep of the vertical segments."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vacv_b200 as vacv
from bench_ops import rand_u8, timeit
for (w, h, wo, ho, b) in ((1920, 1080, 1280, 720, 16), (1920, 1080, 1280, 720, 64), (3840, 2160, 1920, 1080, 4), (3840, 2160, 1920, 1080, 16)):
    src = rand_u8(b, 3, h, w).to(torch.float32)
    nbytes = b * 3 * 4 * (w * h + wo * ho)
    for v, segs in [(1, 0), (0, 0)] + [(0, int(a)) for a in sys.argv[1:]]:
        vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", v); vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", segs)
        ms, mn = timeit(lambda: vacv.resize(src, vacv.NCHW, wo, ho, vacv.INTER_CUBIC), 30)
        print(f"{w}x{h}->{wo}x{ho} x{b} CUBIC_V={v} WALK_SEGS={segs}: {ms:.4f} ms (min {mn:.4f}) {nbytes / ms / 1e6:.0f} GB/s", flush=True)
vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 0); vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", 0)
