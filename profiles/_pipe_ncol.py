import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vacv_b200 as vacv
from bench_ops import rand_u8, timeit, stats
mean, std = stats()
for (w, h, wo, ho) in ((1920, 1080, 608, 608), (1280, 720, 640, 640), (1920, 1080, 416, 416), (1920, 1080, 1280, 720), (1920, 1080, 224, 224), (3840, 2160, 640, 384)):
    src = rand_u8(256 if w < 3000 else 64, w * h * 3 // 2)
    out = torch.empty((src.shape[0], 3, ho, wo), dtype=torch.float32, device="cuda")
    for v in (0, 2, 3, 4, 0):
        vacv.lib.vacv_cuda_set_tuning(b"PIPE_NCOL", v)
        try:
            ms, mn = timeit(lambda: vacv.nv_resize_normalize_chw(src, w, h, wo, ho, mean, std, True, out=out), 30)
            print(f"{w}x{h}->{wo}x{ho} PIPE_NCOL={v}: {ms:.4f} ms (min {mn:.4f})", flush=True)
        except Exception as e:
            print(f"{w}x{h}->{wo}x{ho} PIPE_NCOL={v}: {e}", flush=True)
    del src, out
vacv.lib.vacv_cuda_set_tuning(b"PIPE_NCOL", 0)
