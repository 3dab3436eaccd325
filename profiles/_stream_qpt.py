"""Sweep of the streaming kernels' grid rule (tuning knob STREAM_QPT = 16-byte groups per thread):  python profiles/_stream_qpt.py 4 8 16 32 64 2368"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vacv_b200 as vacv
from bench_ops import rand_u8, timeit, stats
mean, std = stats()
u8_4k = rand_u8(128, 2160, 3840, 3)
f32_4k = torch.empty((128, 2160, 3840, 3), dtype=torch.float32, device="cuda")
u8_hd = rand_u8(128, 1080, 1920, 3)
f32_hd = u8_hd[:32].to(torch.float32)
chw = rand_u8(64, 3, 1080, 1920)
cases = [("normalize u8 hwc 4K x128", lambda: vacv.normalize(u8_4k, vacv.NHWC, mean, std, out=f32_4k), u8_4k.numel() * 5),
         ("dtype u8->f32 1080p x128", lambda: vacv.dtype_change(u8_hd, vacv.FP32), u8_hd.numel() * 5),
         ("dtype f32->u8 1080p x32", lambda: vacv.dtype_change(f32_hd, vacv.INT8), f32_hd.numel() * 5),
         ("normalize f32 hwc 1080p x32", lambda: vacv.normalize(f32_hd, vacv.NHWC, mean, std), f32_hd.numel() * 8),
         ("normalize u8 chw 1080p x64", lambda: vacv.normalize(chw, vacv.NCHW, mean, std), chw.numel() * 5)]
for name, fn, nbytes in cases:
    for q in [int(a) for a in sys.argv[1:]]:
        vacv.lib.vacv_cuda_set_tuning(b"STREAM_QPT", q)
        ms, mn = timeit(fn, 15)
        print(f"{name:30s} STREAM_QPT={q:5d}: {ms:.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)
vacv.lib.vacv_cuda_set_tuning(b"STREAM_QPT", 0)
