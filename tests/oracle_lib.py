"""TEST INFRASTRUCTURE: numpy-facing ctypes bindings for the two CPU checkers.

* ``Oracle``  -> oracle/libvacv_oracle.so   (plain-C restatement, oracle/vacv_oracle.c)
* ``Ref``     -> oracle/_ref/liboracle_ref.so (the unmodified reference compiled from its own sources,
                 oracle/ref_shim.cpp; present only where `make -C oracle ref` has been run)

Nothing in the product imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
FIXTURE_DIR = os.path.join(REF_DIR, "fixtures")

FP32, FP16, INT8 = 0, 1, 2          # vision::DType (tensor.h:12-18)
NCHW, NHWC = 0, 1                   # vision::DLayout (tensor.h:21-24)
INTER_LINEAR, INTER_CUBIC = 1, 2    # va_cv::VInterMode (cv.h:28-36)
COLOR_YUV2BGR_NV12, COLOR_YUV2BGR_NV21 = 91, 93   # cv.h:62-72

_p = C.c_void_p
_i = C.c_int
_f = C.c_float


def _ptr(a):
    return a.ctypes.data_as(_p)


def _c(a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def build_oracle():
    """(Re)build oracle/libvacv_oracle.so if missing or stale."""
    so = os.path.join(ORACLE_DIR, "libvacv_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("vacv_oracle.c", "vacv_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libvacv_oracle.so"], stdout=subprocess.DEVNULL)
    return so


class Oracle:
    """The C restatement."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        self.lib.orc_rotation_matrix.argtypes = [_f, _f, _p, _p]
        for name in dir(self.lib):
            pass

    # colour -----------------------------------------------------------------------------------
    def nv_to_bgr(self, src, w, h, v_first=1):
        src = _c(src, np.uint8)
        dst = np.empty((h, w, 3), np.uint8)
        self.lib.orc_nv_to_bgr(_ptr(src), _i(w), _i(h), _i(v_first), _ptr(dst))
        return dst

    def yuv_to_bgr(self, src, fmt, w, h, y_pitch, c_pitch):
        src = _c(src, np.uint8)
        dst = np.empty((h, w, 3), np.uint8)
        self.lib.orc_yuv_to_bgr(_ptr(src), _i(fmt), _i(w), _i(h), _i(y_pitch), _i(c_pitch), _ptr(dst))
        return dst

    def bgr_to_nv21(self, bgr):
        bgr = _c(bgr, np.uint8)
        h, w = bgr.shape[:2]
        dst = np.empty(w * h * 3 // 2, np.uint8)
        self.lib.orc_bgr_to_nv21(_ptr(bgr), _i(w), _i(h), _ptr(dst))
        return dst

    # copies -----------------------------------------------------------------------------------
    def crop(self, src, w, h, c, layout, left, top, cw, ch):
        src = _c(src)
        dst = np.empty(cw * ch * c, src.dtype)
        self.lib.orc_crop(_ptr(src), _i(w), _i(h), _i(c), _i(src.itemsize), _i(layout),
                          _i(left), _i(top), _i(cw), _i(ch), _ptr(dst))
        return dst.reshape((ch, cw, c) if layout == NHWC else (c, ch, cw))

    def hwc_to_chw(self, src, w, h, c):
        src = _c(src)
        dst = np.empty((c, h, w), src.dtype)
        self.lib.orc_hwc_to_chw(_ptr(src), _i(w), _i(h), _i(c), _i(src.itemsize), _ptr(dst))
        return dst

    def chw_to_hwc(self, src, w, h, c):
        src = _c(src)
        dst = np.empty((h, w, c), src.dtype)
        self.lib.orc_chw_to_hwc(_ptr(src), _i(w), _i(h), _i(c), _i(src.itemsize), _ptr(dst))
        return dst

    def u8_to_f32(self, src):
        src = _c(src, np.uint8)
        dst = np.empty(src.shape, np.float32)
        self.lib.orc_u8_to_f32(_ptr(src), C.c_size_t(src.size), _ptr(dst))
        return dst

    def f32_to_u8(self, src):
        src = _c(src, np.float32)
        dst = np.empty(src.shape, np.uint8)
        self.lib.orc_f32_to_u8(_ptr(src), C.c_size_t(src.size), _ptr(dst))
        return dst

    # resize -----------------------------------------------------------------------------------
    @staticmethod
    def _shape(layout, w, h, c):
        return (h, w, c) if layout == NHWC else (c, h, w)

    def resize_linear(self, src, w, h, c, layout, wo, ho, signed_char=0):
        src = _c(src)
        dst = np.empty(self._shape(layout, wo, ho, c), src.dtype)
        if src.dtype == np.uint8:
            self.lib.orc_resize_linear_u8(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(dst), _i(wo), _i(ho),
                                          _i(signed_char))
        else:
            self.lib.orc_resize_linear_f32(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(dst), _i(wo), _i(ho))
        return dst

    def resize_linear_neon_rule(self, src, w, h, c, layout, wo, ho):
        src = _c(src, np.uint8)
        dst = np.empty(self._shape(layout, wo, ho, c), np.uint8)
        self.lib.orc_resize_linear_u8_neon_rule(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(dst), _i(wo), _i(ho))
        return dst

    def resize_cubic_f32(self, src, w, h, c, layout, wo, ho):
        src = _c(src, np.float32)
        dst = np.empty(self._shape(layout, wo, ho, c), np.float32)
        self.lib.orc_resize_cubic_f32(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(dst), _i(wo), _i(ho))
        return dst

    def resize_cubic_u8(self, src, w, h, c, wo, ho):
        src = _c(src, np.uint8)
        dst = np.empty((ho, wo, c), np.uint8)
        self.lib.orc_resize_cubic_u8_cv24(_ptr(src), _i(w), _i(h), _i(c), _ptr(dst), _i(wo), _i(ho))
        return dst

    # warp -------------------------------------------------------------------------------------
    def invert_affine(self, m):
        m = np.array(m, np.float32).reshape(6).copy()
        self.lib.orc_invert_affine(_ptr(m))
        return m

    def rotation_matrix(self, scale, rot, aux):
        aux = np.array(aux, np.float64)
        m = np.empty(6, np.float32)
        self.lib.orc_rotation_matrix(_f(scale), _f(rot), _ptr(aux), _ptr(m))
        return m

    def warp_affine(self, src, w, h, c, layout, wo, ho, m_inv, signed_char=0, fill=0):
        src = _c(src)
        m_inv = _c(m_inv, np.float32)
        dst = np.full(self._shape(layout, wo, ho, c), fill, src.dtype)
        if src.dtype == np.uint8:
            self.lib.orc_warp_affine_u8(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(dst), _i(wo), _i(ho),
                                        _ptr(m_inv), _i(signed_char))
        else:
            self.lib.orc_warp_affine_f32(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(dst), _i(wo), _i(ho),
                                         _ptr(m_inv))
        return dst

    # statistics / normalize -------------------------------------------------------------------
    def sums_u8(self, src, pixels, c, layout, sums=None):
        src = _c(src, np.uint8)
        if sums is None:
            sums = np.zeros(2 * c, np.uint64)
        self.lib.orc_sums_u8(_ptr(src), C.c_size_t(pixels), _i(c), _i(layout), _ptr(sums))
        return sums

    def sums_f32(self, src, pixels, c, layout):
        src = _c(src, np.float32)
        sums = np.zeros(2 * c, np.float64)
        self.lib.orc_sums_f32(_ptr(src), C.c_size_t(pixels), _i(c), _i(layout), _ptr(sums))
        return sums

    def finalize_mean_stddev_f64(self, sums, c, n):
        sums = _c(sums, np.float64)
        mean = np.empty(c, np.float32)
        std = np.empty(c, np.float32)
        self.lib.orc_finalize_mean_stddev_f64(_ptr(sums), _i(c), C.c_uint64(n), _ptr(mean), _ptr(std))
        return mean, std

    def finalize_mean_stddev(self, sums, c, n):
        sums = _c(sums, np.uint64)
        mean = np.empty(c, np.float32)
        std = np.empty(c, np.float32)
        self.lib.orc_finalize_mean_stddev(_ptr(sums), _i(c), C.c_uint64(n), _ptr(mean), _ptr(std))
        return mean, std

    def mean_stddev_f32_sequential(self, src, pixels, c, layout):
        src = _c(src, np.float32)
        mean = np.empty(c, np.float32)
        std = np.empty(c, np.float32)
        self.lib.orc_mean_stddev_f32_sequential(_ptr(src), C.c_size_t(pixels), _i(c), _i(layout), _ptr(mean), _ptr(std))
        return mean, std

    def normalize(self, src, pixels, c, layout, mean, std):
        src = _c(src)
        mean = _c(mean, np.float32)
        std = _c(std, np.float32)
        dst = np.empty(src.shape, np.float32)
        fn = self.lib.orc_normalize_u8 if src.dtype == np.uint8 else self.lib.orc_normalize_f32
        fn(_ptr(src), C.c_size_t(pixels), _i(c), _i(layout), _ptr(mean), _ptr(std), _ptr(dst))
        return dst

    # compositions -----------------------------------------------------------------------------
    def nv_resize_normalize_chw(self, src, w, h, v_first, wo, ho, mean, std, batch=None, threads=1):
        src = _c(src, np.uint8)
        mean = _c(mean, np.float32)
        std = _c(std, np.float32)
        if batch is None:
            dst = np.empty((3, ho, wo), np.float32)
            self.lib.orc_nv_resize_normalize_chw(_ptr(src), _i(w), _i(h), _i(v_first), _i(wo), _i(ho),
                                                 _ptr(mean), _ptr(std), _ptr(dst))
        else:
            dst = np.empty((batch, 3, ho, wo), np.float32)
            self.lib.orc_nv_resize_normalize_chw_batch(_ptr(src), _i(batch), _i(w), _i(h), _i(v_first), _i(wo), _i(ho),
                                                       _ptr(mean), _ptr(std), _ptr(dst), _i(threads))
        return dst

    def warp_affine_normalize(self, src, w, h, c, m_inv, wo, ho, mean, std):
        src = _c(src, np.uint8)
        m_inv = _c(m_inv, np.float32)
        mean = _c(mean, np.float32)
        std = _c(std, np.float32)
        dst = np.empty((ho, wo, c), np.float32)
        self.lib.orc_warp_affine_normalize(_ptr(src), _i(w), _i(h), _i(c), _ptr(m_inv), _i(wo), _i(ho),
                                           _ptr(mean), _ptr(std), _ptr(dst))
        return dst


def ref_available(schar=False):
    return os.path.exists(os.path.join(REF_DIR, "liboracle_ref_schar.so" if schar else "liboracle_ref.so"))


class Ref:
    """The unmodified reference (naive CPU path + bundled OpenCV 2.4.13), via oracle/ref_shim.cpp."""

    def __init__(self, schar=False):
        self.lib = C.CDLL(os.path.join(REF_DIR, "liboracle_ref_schar.so" if schar else "liboracle_ref.so"))
        self.lib.ref_warp_affine_rot.argtypes = [_p, _i, _i, _i, _i, _i, _f, _f, _p, _p, _i, _i]
        self.lib.ref_crop.argtypes = [_p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _p]

    @staticmethod
    def _dt(a):
        return INT8 if a.dtype == np.uint8 else FP32

    @staticmethod
    def _shape(layout, w, h, c):
        return (h, w, c) if layout == NHWC else (c, h, w)

    def cvt_color(self, src, w, h, code=COLOR_YUV2BGR_NV21):
        src = _c(src, np.uint8)
        dst = np.empty((h, w, 3), np.uint8)
        self.lib.ref_cvt_color(_ptr(src), _i(w), _i(h), _i(code), _ptr(dst))
        return dst

    def bgr2nv21(self, bgr):
        bgr = _c(bgr, np.uint8)
        h, w = bgr.shape[:2]
        dst = np.empty(w * h * 3 // 2, np.uint8)
        self.lib.ref_bgr2nv21(_ptr(bgr), _ptr(dst), _i(w), _i(h))
        return dst

    def crop(self, src, w, h, c, layout, left, top, right, bottom):
        src = _c(src)
        cw, ch = int(np.float32(right) - np.float32(left)), int(np.float32(bottom) - np.float32(top))
        dst = np.empty(self._shape(layout, cw, ch, c), src.dtype)
        self.lib.ref_crop(_ptr(src), w, h, c, self._dt(src), layout, left, top, right, bottom, _ptr(dst))
        return dst

    def change_layout(self, src, w, h, c, layout, new_layout):
        src = _c(src)
        dst = np.empty(self._shape(new_layout, w, h, c), src.dtype)
        self.lib.ref_change_layout(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _i(layout), _i(new_layout), _ptr(dst))
        return dst

    def change_dtype(self, src, w, h, c, layout, new_dtype):
        src = _c(src)
        dst = np.empty(src.shape, np.uint8 if new_dtype == INT8 else np.float32)
        self.lib.ref_change_dtype(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _i(layout), _i(new_dtype), _ptr(dst))
        return dst

    def resize(self, src, w, h, c, layout, wo, ho, interpolation=INTER_LINEAR):
        src = _c(src)
        dst = np.empty(self._shape(layout, wo, ho, c), src.dtype)
        self.lib.ref_resize(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _i(layout), _ptr(dst), _i(wo), _i(ho),
                            _i(interpolation))
        return dst

    def resize_neon(self, src, w, h, layout, wo, ho):
        """The reference's NEON bilinear source (resize_neon.cpp) executed through oracle/neon_emul; 3 channels."""
        src = _c(src, np.uint8)
        pad = 8   # NHWC upscales make the reference read past the row / buffer end (its right-edge clamp uses the tripled width)
        buf = np.zeros(src.size + pad, np.uint8)
        buf[:src.size] = src.ravel()
        dst = np.empty(self._shape(layout, wo, ho, 3), np.uint8)
        self.lib.ref_resize_neon(_ptr(buf), _i(w), _i(h), _i(layout), _ptr(dst), _i(wo), _i(ho))
        return dst

    def resize_cubic_f32_fixed(self, src, w, h, c, layout, wo, ho):
        src = _c(src, np.float32)
        dst = np.empty(self._shape(layout, wo, ho, c), np.float32)
        self.lib.ref_resize_cubic_f32_fixed(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(dst), _i(wo), _i(ho))
        return dst

    def cv_resize(self, src, w, h, c, wo, ho, interpolation, threads=1):
        src = _c(src)
        dst = np.empty((ho, wo, c), src.dtype)
        self.lib.ref_cv_resize(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _ptr(dst), _i(wo), _i(ho),
                               _i(interpolation), _i(threads))
        return dst

    def warp_affine(self, src, w, h, c, layout, wo, ho, m_forward, fill=0):
        """Returns (dst, inverted m) -- the reference overwrites the caller's matrix."""
        src = _c(src)
        m = np.array(m_forward, np.float32).reshape(6).copy()
        dst = np.full(self._shape(layout, wo, ho, c), fill, src.dtype)
        self.lib.ref_warp_affine(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _i(layout), _ptr(m), _ptr(dst),
                                 _i(wo), _i(ho))
        return dst, m

    def warp_affine_rot(self, src, w, h, c, layout, wo, ho, scale, rot, aux, fill=0):
        src = _c(src)
        aux = np.array(aux, np.float64)
        dst = np.full(self._shape(layout, wo, ho, c), fill, src.dtype)
        self.lib.ref_warp_affine_rot(_ptr(src), w, h, c, self._dt(src), layout, scale, rot, _ptr(aux), _ptr(dst), wo, ho)
        return dst

    def cv_warp_affine(self, src, w, h, c, wo, ho, m_forward):
        src = _c(src)
        m = _c(m_forward, np.float32)
        dst = np.zeros((ho, wo, c), src.dtype)
        self.lib.ref_cv_warp_affine(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _ptr(m), _ptr(dst), _i(wo), _i(ho))
        return dst

    def normalize(self, src, w, h, c, layout, mean=None, std=None):
        src = _c(src)
        dst = np.empty(self._shape(layout, w, h, c), np.float32)
        if mean is None:
            self.lib.ref_normalize(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _i(layout), None, None, _ptr(dst))
        else:
            mean = _c(mean, np.float32)
            std = _c(std, np.float32)
            self.lib.ref_normalize(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _i(layout), _ptr(mean), _ptr(std),
                                   _ptr(dst))
        return dst

    def mean_stddev_f32(self, src, w, h, c, layout):
        src = _c(src, np.float32)
        mean = np.empty(c, np.float32)
        std = np.empty(c, np.float32)
        self.lib.ref_mean_stddev_f32(_ptr(src), _i(w), _i(h), _i(c), _i(layout), _ptr(mean), _ptr(std))
        return mean, std

    def cv_mean_stddev(self, src, w, h, c):
        src = _c(src)
        mean = np.zeros(4, np.float64)
        std = np.zeros(4, np.float64)
        self.lib.ref_cv_mean_stddev(_ptr(src), _i(w), _i(h), _i(c), _i(self._dt(src)), _ptr(mean), _ptr(std))
        return mean[:c], std[:c]

    def imread(self, path, color=1):
        w, h, c = _i(), _i(), _i()
        if self.lib.ref_imread(path.encode(), _i(color), None, C.byref(w), C.byref(h), C.byref(c)) != 0:
            raise FileNotFoundError(path)
        dst = np.empty((h.value, w.value, c.value), np.uint8)
        self.lib.ref_imread(path.encode(), _i(color), _ptr(dst), C.byref(w), C.byref(h), C.byref(c))
        return dst

    def pipeline(self, src, w, h, code, wo, ho, mean, std, batch=None, threads=1, out=None):
        src = _c(src, np.uint8)
        mean = _c(mean, np.float32)
        std = _c(std, np.float32)
        if batch is None:
            dst = np.empty((3, ho, wo), np.float32)
            self.lib.ref_pipeline_nv_resize_norm_chw(_ptr(src), _i(w), _i(h), _i(code), _i(wo), _i(ho), _ptr(mean),
                                                     _ptr(std), _ptr(dst))
        else:
            dst = out if out is not None else np.empty((batch, 3, ho, wo), np.float32)
            self.lib.ref_pipeline_nv_resize_norm_chw_batch(_ptr(src), _i(batch), _i(w), _i(h), _i(code), _i(wo), _i(ho),
                                                           _ptr(mean), _ptr(std), _ptr(dst), _i(threads))
        return dst


def load_fixture(name):
    """Decoded reference JPEG (by the bundled OpenCV) from oracle/_ref/fixtures, or None if not staged."""
    if not os.path.isdir(FIXTURE_DIR):
        return None
    for f in os.listdir(FIXTURE_DIR):
        if f.startswith(name + "_") and f.endswith(".bin"):
            w, h, c = (int(v) for v in f[len(name) + 1:-4].split("x"))
            return np.fromfile(os.path.join(FIXTURE_DIR, f), np.uint8).reshape(h, w, c)
    return None
