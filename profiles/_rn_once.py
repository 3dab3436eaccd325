"""One launch of resize_normalize (BGR 1080p -> 640x640 fp32 CHW, x128) for ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_ops as B
import vacv_b200 as vacv
mean, std = B.stats()
bgr = B.rand_u8(128, 1080, 1920, 3)
for _ in range(3):
    vacv.resize_normalize(bgr, 640, 640, mean, std, vacv.NCHW)
torch.cuda.synchronize()
