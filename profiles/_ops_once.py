"""One launch of the crop and the plane-normalize kernels at their bench_ops shapes (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_ops as B
import vacv_b200 as vacv
mean, std = B.stats()
bgr = B.rand_u8(128, 1080, 1920, 3)
chw = B.rand_u8(64, 3, 1080, 1920)
for _ in range(2):
    vacv.crop(bgr, vacv.NHWC, 321, 181, 1280, 720)
    vacv.normalize(chw, vacv.NCHW, mean, std)
torch.cuda.synchronize()
