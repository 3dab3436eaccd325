"""u8 bicubic walkers on large batches: vertical segments per column strip (VACV_WALK_SEGS); the launcher defaults are within 5 % of the best.   python profiles/_cubic_segs_big.py"""
import sys; sys.path.insert(0,'.')
import torch, vacv_b200 as vacv
from bench_ops import rand_u8, timeit
for name,(w,h),(wo,ho),b in (("c4 1440p->1080p x128",(2560,1440),(1920,1080),128),("1080p->720p x64",(1920,1080),(1280,720),64),("4K->1080p x32",(3840,2160),(1920,1080),32)):
    src = rand_u8(b,h,w,3); line=[]
    for segs in (0,4,7,12,17,24,34,68):
        vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", segs)
        ms,mn = timeit(lambda: vacv.resize(src, vacv.NHWC, wo, ho, vacv.INTER_CUBIC), 15)
        line.append(f"{segs}: {ms:.4f}")
    vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", 0)
    print(name, "  ".join(line), flush=True); del src
