import sys; sys.path.insert(0,'.')
import torch, vacv_b200 as vacv
from bench_ops import rand_u8, timeit
big = rand_u8(16, 2160, 3840, 3)
for v in (0, 2, 1):
    vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", v)
    out = vacv.resize(big, vacv.NHWC, 1920, 1080)
    if v == 0: ref = out.clone()
    ms, mn = timeit(lambda: vacv.resize(big, vacv.NHWC, 1920, 1080), 30)
    print("2:1 LINEAR_V", v, ms, mn, bool(torch.equal(out, ref)))
