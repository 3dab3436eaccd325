import sys; sys.path.insert(0,'.')
import torch, vacv_b200 as vacv
from bench_ops import rand_u8, timeit
chw = rand_u8(64, 3, 1080, 1920)
for (wo,ho) in ((1280,720),(960,540)):
    ms, mn = timeit(lambda: vacv.resize(chw, vacv.NCHW, wo, ho), 20)
    gb = 64*3*(1920*1080 + wo*ho)/ (ms*1e-3)/1e9
    print(f"chw 1080p->{wo}x{ho} x64: {ms:.4f} ms {gb:.0f} GB/s frac {gb/6565.5:.3f}")
q = rand_u8(32, 3, 1440, 2560)
ms, mn = timeit(lambda: vacv.resize(q, vacv.NCHW, 1920, 1080), 20)
gb = 32*3*(2560*1440 + 1920*1080)/(ms*1e-3)/1e9
print(f"chw 1440p->1080p x32: {ms:.4f} ms {gb:.0f} GB/s frac {gb/6565.5:.3f}")
f = rand_u8(16, 1080, 1920, 3).to(torch.float32)
for (wo,ho) in ((1280,720),(960,540)):
    ms, mn = timeit(lambda: vacv.resize(f, vacv.NHWC, wo, ho), 20)
    gb = 16*12*(1920*1080 + wo*ho)/ (ms*1e-3)/1e9
    print(f"f32 hwc 1080p->{wo}x{ho} x16: {ms:.4f} ms {gb:.0f} GB/s frac {gb/6565.5:.3f}")
src = rand_u8(128, 1080, 1920, 3)
mean = torch.tensor([104.,117.,123.], device="cuda"); std = torch.tensor([58.,57.,57.], device="cuda")
for (wo,ho) in ((1280,720),(960,540),(608,608)):
    ms, mn = timeit(lambda: vacv.resize_normalize(src, wo, ho, mean, std, vacv.NCHW), 20)
    gb = 128*(1920*1080*3 + wo*ho*12)/(ms*1e-3)/1e9
    print(f"resize_normalize u8 hwc 1080p->{wo}x{ho} chw x128: {ms:.4f} ms {gb:.0f} GB/s frac {gb/6565.5:.3f}")
