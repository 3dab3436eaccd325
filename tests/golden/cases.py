"""Shared case list for the golden digests: each case maps a name to a callable(api) -> ndarray, where `api` is either
the compiled reference (make_golden.py) or the C oracle (test_golden.py).  Inputs are seeded, so the digests are
reproducible anywhere; they are sha256 over the raw output bytes."""
import hashlib

import numpy as np

FP32, INT8 = 0, 2
NCHW, NHWC = 0, 1
MEAN = np.array([103.53, 116.28, 123.675], np.float32)
STD = np.array([57.375, 57.12, 58.395], np.float32)
M_TEST = [0.849158, 0.012257, -474.827, -0.01225, 0.849158, -379.18]   # reference test_warp_affine.cpp:31-32
ROT = dict(scale=1.073914, rot=-3.314525, aux=[738.518372, 537.672852, 204.766998, 73.329681])   # :198-205


def u8(seed, *shape):
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def f32(seed, *shape):
    return (np.random.default_rng(seed).random(shape, dtype=np.float32) * 255).astype(np.float32)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# (name, reference-side callable, oracle-side callable)
def build_cases():
    c = []
    nv = u8(1, 640 * 360 * 3 // 2)
    c.append(("nv21_to_bgr_640x360", lambda r: r.cvt_color(nv, 640, 360, 93), lambda o: o.nv_to_bgr(nv, 640, 360, 1)))
    c.append(("nv12code_to_bgr_640x360", lambda r: r.cvt_color(nv, 640, 360, 91), lambda o: o.nv_to_bgr(nv, 640, 360, 1)))
    img = u8(2, 360, 640, 3)
    imgf = f32(3, 360, 640, 3)
    chw = np.ascontiguousarray(img.transpose(2, 0, 1))
    c.append(("crop_hwc_u8", lambda r: r.crop(img, 640, 360, 3, NHWC, 7.9, 3.2, 200.5, 99.7), lambda o: o.crop(img, 640, 360, 3, NHWC, 7, 3, 192, 96)))
    c.append(("crop_chw_f32", lambda r: r.crop(np.ascontiguousarray(imgf.transpose(2, 0, 1)), 640, 360, 3, NCHW, 0, 0, 320, 180),
              lambda o: o.crop(np.ascontiguousarray(imgf.transpose(2, 0, 1)), 640, 360, 3, NCHW, 0, 0, 320, 180)))
    c.append(("hwc_to_chw_u8", lambda r: r.change_layout(img, 640, 360, 3, NHWC, NCHW), lambda o: o.hwc_to_chw(img, 640, 360, 3)))
    c.append(("u8_to_f32", lambda r: r.change_dtype(img, 640, 360, 3, NHWC, FP32), lambda o: o.u8_to_f32(img)))
    c.append(("f32_to_u8", lambda r: r.change_dtype(imgf, 640, 360, 3, NHWC, INT8), lambda o: o.f32_to_u8(imgf)))
    c.append(("resize_linear_u8_hwc_320x180", lambda r: r.resize(img, 640, 360, 3, NHWC, 320, 180, 1), lambda o: o.resize_linear(img, 640, 360, 3, NHWC, 320, 180)))
    c.append(("resize_linear_u8_chw_213x97", lambda r: r.resize(chw, 640, 360, 3, NCHW, 213, 97, 1), lambda o: o.resize_linear(chw, 640, 360, 3, NCHW, 213, 97)))
    c.append(("resize_linear_u8_neon_source_chw_213x97", lambda r: r.resize_neon(chw, 640, 360, NCHW, 213, 97),   # resize_neon.cpp over neon_emul
              lambda o: o.resize_linear_neon_rule(chw, 640, 360, 3, NCHW, 213, 97)))
    c.append(("resize_linear_u8_neon_source_hwc_320x180", lambda r: r.resize_neon(img, 640, 360, NHWC, 320, 180),
              lambda o: o.resize_linear_neon_rule(img, 640, 360, 3, NHWC, 320, 180)))
    c.append(("resize_linear_f32_hwc_500x300", lambda r: r.resize(imgf, 640, 360, 3, NHWC, 500, 300, 1), lambda o: o.resize_linear(imgf, 640, 360, 3, NHWC, 500, 300)))
    c.append(("resize_cubic_f32_hwc_300x300", lambda r: r.resize(imgf, 640, 360, 3, NHWC, 300, 300, 2), lambda o: o.resize_cubic_f32(imgf, 640, 360, 3, NHWC, 300, 300)))
    c.append(("resize_cubic_f32_hwc_480x270_fixed", lambda r: r.resize_cubic_f32_fixed(imgf, 640, 360, 3, NHWC, 480, 270), lambda o: o.resize_cubic_f32(imgf, 640, 360, 3, NHWC, 480, 270)))
    c.append(("resize_cubic_u8_cv24_480x270", lambda r: r.cv_resize(img, 640, 360, 3, 480, 270, 2), lambda o: o.resize_cubic_u8(img, 640, 360, 3, 480, 270)))
    c.append(("resize_cubic_u8_cv24_up_803x451", lambda r: r.cv_resize(img, 640, 360, 3, 803, 451, 2), lambda o: o.resize_cubic_u8(img, 640, 360, 3, 803, 451)))
    big = u8(4, 720, 1280, 3)
    c.append(("warp_affine_u8_240x240", lambda r: r.warp_affine(big, 1280, 720, 3, NHWC, 240, 240, M_TEST)[0],
              lambda o: o.warp_affine(big, 1280, 720, 3, NHWC, 240, 240, o.invert_affine(M_TEST))))
    bigf = big.astype(np.float32)
    c.append(("warp_affine_f32_240x240", lambda r: r.warp_affine(bigf, 1280, 720, 3, NHWC, 240, 240, M_TEST)[0],
              lambda o: o.warp_affine(bigf, 1280, 720, 3, NHWC, 240, 240, o.invert_affine(M_TEST))))
    grey = u8(5, 720, 1280, 1)
    c.append(("warp_affine_rot_grey_140x210", lambda r: r.warp_affine_rot(grey, 1280, 720, 1, NHWC, 140, 210, ROT["scale"], ROT["rot"], ROT["aux"]),
              lambda o: o.warp_affine(grey, 1280, 720, 1, NHWC, 140, 210, o.invert_affine(o.rotation_matrix(ROT["scale"], ROT["rot"], ROT["aux"])))))
    c.append(("normalize_u8_hwc", lambda r: r.normalize(img, 640, 360, 3, NHWC, MEAN, STD), lambda o: o.normalize(img, 640 * 360, 3, NHWC, MEAN, STD)))
    c.append(("normalize_f32_chw", lambda r: r.normalize(np.ascontiguousarray(imgf.transpose(2, 0, 1)), 640, 360, 3, NCHW, MEAN, STD),
              lambda o: o.normalize(np.ascontiguousarray(imgf.transpose(2, 0, 1)), 640 * 360, 3, NCHW, MEAN, STD)))
    c.append(("mean_stddev_f32_sequential_hwc", lambda r: np.concatenate(r.mean_stddev_f32(imgf, 640, 360, 3, NHWC)),
              lambda o: np.concatenate(o.mean_stddev_f32_sequential(imgf, 640 * 360, 3, NHWC))))
    nv2 = u8(6, 1920 * 1080 * 3 // 2)
    c.append(("pipeline_c2_1080p_to_640x640", lambda r: r.pipeline(nv2, 1920, 1080, 93, 640, 640, MEAN, STD),
              lambda o: o.nv_resize_normalize_chw(nv2, 1920, 1080, 1, 640, 640, MEAN, STD)))
    return c
