"""Device time of every operator on ONE frame (the drop-in API's shape: one image per call), to find launchers that under-fill the GPU
on small batches.   python profiles/_batch1.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vacv_b200 as vacv
from bench_ops import rand_u8, timeit, stats, face_matrices
mean, std = stats()
def t(name, fn):
    ms, mn = timeit(fn, 20)
    print(f"{name:60s} {ms*1000:8.1f} us", flush=True)
bgr = rand_u8(1, 1080, 1920, 3); chw = rand_u8(1, 3, 1080, 1920); nv = rand_u8(1, 1920 * 1080 * 3 // 2); f = bgr.to(torch.float32)
uhd = rand_u8(1, 2160, 3840, 3)
t("cvt_nv2bgr 1080p", lambda: vacv.cvt_nv2bgr(nv, 1920, 1080))
t("fused nv12 -> 640x640 chw f32", lambda: vacv.nv_resize_normalize_chw(nv, 1920, 1080, 640, 640, mean, std))
t("fused nv12 -> 608x608 chw f32", lambda: vacv.nv_resize_normalize_chw(nv, 1920, 1080, 608, 608, mean, std))
t("resize linear u8 hwc 1080p->640x360", lambda: vacv.resize(bgr, vacv.NHWC, 640, 360))
t("resize linear u8 hwc 1080p->640x640", lambda: vacv.resize(bgr, vacv.NHWC, 640, 640))
t("resize linear u8 hwc 1080p->1280x720", lambda: vacv.resize(bgr, vacv.NHWC, 1280, 720))
t("resize linear u8 hwc 1080p->1000x500", lambda: vacv.resize(bgr, vacv.NHWC, 1000, 500))
t("resize linear u8 chw 1080p->640x360", lambda: vacv.resize(chw, vacv.NCHW, 640, 360))
t("resize linear f32 hwc 1080p->640x360", lambda: vacv.resize(f, vacv.NHWC, 640, 360))
t("resize cubic u8 hwc 1080p->1280x720", lambda: vacv.resize(bgr, vacv.NHWC, 1280, 720, vacv.INTER_CUBIC))
t("resize cubic u8 hwc 4K->1080p", lambda: vacv.resize(uhd, vacv.NHWC, 1920, 1080, vacv.INTER_CUBIC))
t("resize cubic f32 hwc 1080p->1280x720", lambda: vacv.resize(f, vacv.NHWC, 1280, 720, vacv.INTER_CUBIC))
t("resize_normalize u8 hwc 1080p->640x640 chw", lambda: vacv.resize_normalize(bgr, 640, 640, mean, std, vacv.NCHW))
t("normalize u8 hwc 1080p", lambda: vacv.normalize(bgr, vacv.NHWC, mean, std))
t("normalize f32 hwc 1080p", lambda: vacv.normalize(f, vacv.NHWC, mean, std))
t("layout hwc->chw u8 1080p", lambda: vacv.layout_change(bgr, vacv.NHWC, vacv.NCHW))
t("dtype u8->f32 1080p", lambda: vacv.dtype_change(bgr, vacv.FP32))
t("crop u8 hwc 1080p->1280x720", lambda: vacv.crop(bgr, vacv.NHWC, 321, 181, 1280, 720))
t("sums_u8 4K", lambda: vacv.sums_u8(uhd, vacv.NHWC, False))
minv, _ = face_matrices(8, 1920, 1080, 112)
idx = torch.zeros(8, dtype=torch.int32, device="cuda")
t("warp_affine u8 8 crops 1080p->112x112", lambda: vacv.warp_affine(bgr, vacv.NHWC, minv, 112, 112, idx))
t("warp_affine_normalize 8 crops 1080p->112x112", lambda: vacv.warp_affine_normalize(bgr, minv, 112, 112, mean, std, idx))
