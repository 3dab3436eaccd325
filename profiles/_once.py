"""A couple of launches of ONE operator at its bench_ops shape, for ncu captures:  python profiles/_once.py <case>"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_ops as B
import vacv_b200 as vacv

case = sys.argv[1]
mean, std = B.stats()
if case == "c1":
    src = B.rand_u8(256, 1080, 1920, 3)
    fn = lambda: vacv.resize(src, vacv.NHWC, 640, 360)
elif case == "c2":
    src = B.rand_u8(256, 1920 * 1080 * 3 // 2)
    out = torch.empty((256, 3, 640, 640), dtype=torch.float32, device="cuda")
    fn = lambda: vacv.nv_resize_normalize_chw(src, 1920, 1080, 640, 640, mean, std, True, out=out)
elif case == "c2p":   # pitched decoder surfaces (pitch 2048)
    src = B.rand_u8(256 * 2048 * 1080 * 3 // 2)
    out = torch.empty((256, 3, 640, 640), dtype=torch.float32, device="cuda")
    fn = lambda: vacv.yuv_resize_normalize_chw(src, vacv.YUV_NV12, 1920, 1080, 640, 640, mean, std, y_pitch=2048, c_pitch=2048, batch=256, out=out)
elif case == "cvt":
    src = B.rand_u8(128, 1920 * 1080 * 3 // 2)
    fn = lambda: vacv.cvt_nv2bgr(src, 1920, 1080)
elif case == "layout":
    src = B.rand_u8(128, 1080, 1920, 3)
    fn = lambda: vacv.layout_change(src, vacv.NHWC, vacv.NCHW)
elif case == "sums":
    src = B.rand_u8(128, 2160, 3840, 3)
    sums = torch.zeros((1, 3, 2), dtype=torch.int64, device="cuda")
    fn = lambda: vacv.sums_u8(src, vacv.NHWC, False, sums)
elif case == "normu8":
    src = B.rand_u8(128, 2160, 3840, 3)
    out = torch.empty((128, 2160, 3840, 3), dtype=torch.float32, device="cuda")
    fn = lambda: vacv.normalize(src, vacv.NHWC, mean, std, out=out)
elif case == "dtype":
    src = B.rand_u8(128, 1080, 1920, 3)
    fn = lambda: vacv.dtype_change(src, vacv.FP32)
elif case == "c4":
    src = B.rand_u8(128, 1440, 2560, 3)
    fn = lambda: vacv.resize(src, vacv.NHWC, 1920, 1080, vacv.INTER_CUBIC)
elif case == "c3u8":
    frames = B.rand_u8(512, 720, 1280, 3)
    minv, _ = B.face_matrices(4096, 1280, 720, 112)
    idx = (torch.arange(4096, device="cuda") % 512).to(torch.int32)
    fn = lambda: vacv.warp_affine(frames, vacv.NHWC, minv, 112, 112, idx)
elif case == "lin720":
    src = B.rand_u8(64, 1080, 1920, 3)
    fn = lambda: vacv.resize(src, vacv.NHWC, 1280, 720)
elif case == "linchw720":   # planes walker, 3 : 2
    src = B.rand_u8(64, 3, 1080, 1920)
    fn = lambda: vacv.resize(src, vacv.NCHW, 1280, 720)
elif case == "linchw4k":    # planes walker, 2 : 1
    src = B.rand_u8(16, 3, 2160, 3840)
    fn = lambda: vacv.resize(src, vacv.NCHW, 1920, 1080)
elif case == "linchw":
    src = B.rand_u8(64, 3, 1080, 1920)
    fn = lambda: vacv.resize(src, vacv.NCHW, 640, 360)
elif case == "cropchw":
    src = B.rand_u8(64, 3, 1080, 1920)
    fn = lambda: vacv.crop(src, vacv.NCHW, 321, 181, 1280, 720)
elif case == "cropf32":
    src = B.rand_u8(16, 1080, 1920, 3).to(torch.float32)
    fn = lambda: vacv.crop(src, vacv.NHWC, 321, 181, 1280, 720)
elif case == "cubf32chw":
    src = B.rand_u8(16, 3, 1080, 1920).to(torch.float32)
    fn = lambda: vacv.resize(src, vacv.NCHW, 1280, 720, vacv.INTER_CUBIC)
else:
    raise SystemExit(f"unknown case {case}")
for _ in range(3):
    fn()
torch.cuda.synchronize()
