"""One launch of each C3 warp kernel (for ncu): staged f32 HWC, staged u8, direct-gather u8."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_ops as B
import vacv_b200 as vacv
n, nf, w, h, wo = 4096, 512, 1280, 720, 112
mean, std = B.stats()
frames = B.rand_u8(nf, h, w, 3)
minv, _ = B.face_matrices(n, w, h, wo)
idx = (torch.arange(n, device="cuda") % nf).to(torch.int32)
for _ in range(2):
    vacv.warp_affine_normalize(frames, minv, wo, wo, mean, std, idx)
    vacv.warp_affine(frames, vacv.NHWC, minv, wo, wo, idx)
torch.cuda.synchronize()
