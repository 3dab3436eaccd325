"""u8 BGR warp_affine on config 3's shape: the kernel variants behind VACV_WARP_V (0 pack kernel, 2 two pixels per thread in flight, 3 32-bit
tap loads, 1 first-generation gather kernel), each checked against variant 1 byte for byte.   python profiles/_warp_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vacv_b200 as vacv
from bench_ops import rand_u8, timeit, face_matrices
n, nf, w, h, wo = 4096, 512, 1280, 720, 112
frames = rand_u8(nf, h, w, 3)
minv, _ = face_matrices(n, w, h, wo)
idx = (torch.arange(n, device="cuda") % nf).to(torch.int32)
vacv.lib.vacv_cuda_set_tuning(b"WARP_V", 1)
want = vacv.warp_affine(frames, vacv.NHWC, minv, wo, wo, idx).clone()
for v in (0, 2, 3, 1):
    vacv.lib.vacv_cuda_set_tuning(b"WARP_V", v)
    same = bool(torch.equal(vacv.warp_affine(frames, vacv.NHWC, minv, wo, wo, idx), want))
    ms, mn = timeit(lambda: vacv.warp_affine(frames, vacv.NHWC, minv, wo, wo, idx), 30)
    print(f"WARP_V={v}: {ms:.4f} ms (min {mn:.4f})  bit-exact vs V=1: {same}", flush=True)
vacv.lib.vacv_cuda_set_tuning(b"WARP_V", 0)
