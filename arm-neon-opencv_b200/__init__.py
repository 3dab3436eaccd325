"""vacv-b200: Python plumbing over the C-ABI of libvacv_cuda.so (include/vacv_cuda.h).

The product is the CUDA library and its C/C++ host layer; this module only lends torch's device memory,
streams and torch.distributed to tests/ and bench.py.  Every function takes/returns CUDA torch tensors and
calls straight through ctypes into the extension.  There is NO fallback: if the shared library is missing the
import fails, and every call checks the status code and raises with the library's message.

(The directory name has hyphens, so import it through the repo-root shim:  ``import vacv_b200 as vacv``.)
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VACV_B200_LIB") or os.path.join(_HERE, "libvacv_cuda.so")   # override: A/B runs of two builds on one box

FP32, FP16, INT8, FP64 = 0, 1, 2, 3
BF16 = 16   # extension: output type of the fused pipeline only
NCHW, NHWC = 0, 1
INTER_LINEAR, INTER_CUBIC = 1, 2
FLAG_NONE, FLAG_NEON_RULE, FLAG_SIGNED_CHAR, FLAG_DIRECT_GATHER, FLAG_TILED = 0, 1, 2, 0x100, 0x200

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
lib = C.CDLL(LIB_PATH)
lib.vacv_cuda_last_error.restype = C.c_char_p

_vp, _i, _sz, _u64, _f = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_float
_SIGS = {
    "vacv_cuda_cvt_nv2bgr": [_vp, _vp, _i, _i, _i, _i, _vp],
    "vacv_cuda_crop": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vacv_cuda_layout_change": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vacv_cuda_dtype_change": [_vp, _vp, _sz, _i, _i, _vp],
    "vacv_cuda_resize": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vacv_cuda_warp_affine": [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp],
    "vacv_cuda_sums_u8": [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "vacv_cuda_finalize_mean_stddev": [_vp, _i, _i, _u64, _vp, _vp, _vp],
    "vacv_cuda_sums_f32": [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "vacv_cuda_finalize_mean_stddev_f64": [_vp, _i, _i, _u64, _vp, _vp, _vp],
    "vacv_cuda_normalize": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "vacv_cuda_nv_resize_normalize_chw": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "vacv_cuda_yuv_resize_normalize_chw": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "vacv_cuda_cvt_yuv2bgr": [_vp, _vp, _vp, _i, _vp],
    "vacv_cuda_yuv_letterbox_normalize_chw": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "vacv_letterbox_rect": [_i, _i, _i, _i, _vp],
    "vacv_cuda_yuv_normalize_chw_host": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i],
    "vacv_cuda_nv_resize_normalize_chw_host": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i],
    "vacv_cuda_resize_normalize": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "vacv_cuda_warp_affine_normalize": [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp],
    "vacv_cuda_normalize_batch_global_cb": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "vacv_cuda_normalize_batch_global_p2p": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "vacv_cuda_p2p_create": [_vp, _i, _i, _vp],
    "vacv_cuda_p2p_connect": [_vp, _vp],
    "vacv_cuda_p2p_destroy": [_vp],
    "vacv_cuda_p2p_allreduce_u64": [_vp, _vp, _i, _vp],
    "vacv_cuda_p2p_status": [_vp, _vp],
    "vacv_cuda_set_last_error": [_i, C.c_char_p],
    "vacv_cuda_device_count": [_vp],
    "vacv_cuda_set_device": [_i],
    "vacv_cuda_get_device": [_vp],
    "vacv_cuda_set_tuning": [C.c_char_p, _i],
    "vacv_cuda_malloc": [_vp, _sz],
    "vacv_cuda_free": [_vp],
    "vacv_cuda_host_alloc": [_vp, _sz],
    "vacv_cuda_host_alloc_flags": [_vp, _sz, _i],
    "vacv_cuda_host_free": [_vp],
    "vacv_cuda_memcpy_h2d": [_vp, _vp, _sz, _vp],
    "vacv_cuda_memcpy_d2h": [_vp, _vp, _sz, _vp],
    "vacv_cuda_memcpy2d_h2d": [_vp, _sz, _vp, _sz, _sz, _sz, _vp],
    "vacv_cuda_memset": [_vp, _i, _sz, _vp],
    "vacv_cuda_stream_create": [_vp],
    "vacv_cuda_stream_destroy": [_vp],
    "vacv_cuda_stream_sync": [_vp],
    "vacv_invert_affine": [_vp],
    "vacv_rotation_matrix": [_f, _f, _vp, _vp],
}
for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = None if _name.startswith(("vacv_invert", "vacv_rotation", "vacv_letterbox")) else _i
EXPORTS = ["vacv_cuda_abi_version", "vacv_cuda_last_error"] + list(_SIGS)


# libvacv_dist.so = the NCCL transport of config 5 (include/vacv_dist.h); loaded on first use so that single-GPU users of this
# module never touch NCCL.  Import torch.distributed's NCCL first if you want both to share one libnccl.so.2.
DIST_LIB_PATH = os.path.join(_HERE, "libvacv_dist.so")
_DIST_SIGS = {
    "vacv_cuda_normalize_batch_global": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "vacv_dist_allreduce_u64": [_vp, _vp, _i, _vp],
    "vacv_dist_nccl_version": [_vp],
    "vacv_dist_nccl_unique_id": [_vp],
    "vacv_dist_nccl_comm_create": [_vp, _i, _i, _vp],
    "vacv_dist_nccl_comm_destroy": [_vp],
}
_dist_lib = None


def dist_lib():
    global _dist_lib
    if _dist_lib is None:
        if not os.path.exists(DIST_LIB_PATH):
            raise ImportError(f"{DIST_LIB_PATH} not built -- run __graft_entry__.build()")
        d = C.CDLL(DIST_LIB_PATH)
        for name, args in _DIST_SIGS.items():
            fn = getattr(d, name)
            fn.argtypes = args
            fn.restype = _i
        _dist_lib = d
    return _dist_lib


class VacvError(RuntimeError):
    pass


def _check(status):
    if status != 0:
        raise VacvError(f"vacv status {status}: {lib.vacv_cuda_last_error().decode()}")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev(t, dtype=None):
    if not t.is_cuda:
        raise VacvError("expected a CUDA tensor (the C-ABI takes device pointers)")
    if dtype is not None and t.dtype != dtype:
        raise VacvError(f"expected dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def _dt(t):
    if t.dtype == torch.uint8:
        return INT8
    if t.dtype == torch.float32:
        return FP32
    if t.dtype in (torch.float16, torch.int16):
        return FP16
    raise VacvError(f"unsupported dtype {t.dtype}")


def _shape(layout, n, w, h, c):
    return (n, h, w, c) if layout == NHWC else (n, c, h, w)


# ------------------------------------------------------------------------------------------------ operators
def cvt_nv2bgr(src, w, h, v_first=True):
    """src: uint8 [B, w*h*3/2] -> uint8 [B, h, w, 3] (vacv_cuda_cvt_nv2bgr)."""
    src = _dev(src, torch.uint8)
    b = src.numel() // (w * h * 3 // 2)
    dst = torch.empty((b, h, w, 3), dtype=torch.uint8, device=src.device)
    _check(lib.vacv_cuda_cvt_nv2bgr(src.data_ptr(), dst.data_ptr(), b, w, h, int(bool(v_first)), _stream()))
    return dst


def crop(src, layout, left, top, cw, ch):
    """src: [B,h,w,c] (NHWC) or [B,c,h,w] (NCHW), uint8/float32."""
    src = _dev(src)
    if layout == NHWC:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    dst = torch.empty(_shape(layout, b, cw, ch, c), dtype=src.dtype, device=src.device)
    _check(lib.vacv_cuda_crop(src.data_ptr(), dst.data_ptr(), b, w, h, c, _dt(src), layout, left, top, cw, ch, _stream()))
    return dst


def layout_change(src, from_layout, to_layout):
    src = _dev(src)
    if from_layout == NHWC:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    dst = torch.empty(_shape(to_layout, b, w, h, c), dtype=src.dtype, device=src.device)
    _check(lib.vacv_cuda_layout_change(src.data_ptr(), dst.data_ptr(), b, w, h, c, _dt(src), from_layout, to_layout, _stream()))
    return dst


def dtype_change(src, to_dtype):
    src = _dev(src)
    dst = torch.empty(src.shape, dtype=torch.float32 if to_dtype == FP32 else torch.uint8, device=src.device)
    _check(lib.vacv_cuda_dtype_change(src.data_ptr(), dst.data_ptr(), src.numel(), _dt(src), to_dtype, _stream()))
    return dst


def resize(src, layout, w_out, h_out, interpolation=INTER_LINEAR, flags=FLAG_NONE):
    src = _dev(src)
    if layout == NHWC:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    dst = torch.empty(_shape(layout, b, w_out, h_out, c), dtype=src.dtype, device=src.device)
    _check(lib.vacv_cuda_resize(src.data_ptr(), dst.data_ptr(), b, w, h, c, _dt(src), layout, w_out, h_out,
                                interpolation, flags, _stream()))
    return dst


def invert_affine(m):
    """Host helper: forward 2x3 (6 floats) -> inverse, with the reference's arithmetic."""
    arr = (C.c_float * 6)(*[float(v) for v in m])
    lib.vacv_invert_affine(C.cast(arr, _vp))
    return list(arr)


def rotation_matrix(scale, rot_deg, aux):
    a = (C.c_double * 4)(*[float(v) for v in aux])
    m = (C.c_float * 6)()
    lib.vacv_rotation_matrix(scale, rot_deg, C.cast(a, _vp), C.cast(m, _vp))
    return list(m)


def warp_affine(frames, layout, minv, w_out, h_out, frame_idx=None, flags=FLAG_NONE):
    """frames: [F,h,w,c]/[F,c,h,w]; minv: float32 [N,6] INVERTED matrices (device); frame_idx: int32 [N] or None."""
    frames = _dev(frames)
    minv = _dev(minv, torch.float32)
    if layout == NHWC:
        f, h, w, c = frames.shape
    else:
        f, c, h, w = frames.shape
    n = minv.numel() // 6
    idx = _dev(frame_idx, torch.int32) if frame_idx is not None else None
    dst = torch.empty(_shape(layout, n, w_out, h_out, c), dtype=frames.dtype, device=frames.device)
    _check(lib.vacv_cuda_warp_affine(frames.data_ptr(), f, w, h, c, _dt(frames), layout,
                                     idx.data_ptr() if idx is not None else None, minv.data_ptr(), n,
                                     dst.data_ptr(), w_out, h_out, flags, _stream()))
    return dst


def sums_u8(src, layout, per_frame=False, sums=None):
    """Exact per-channel (sum x, sum x^2) as int64 [sets, c, 2]; accumulates into `sums` if given."""
    src = _dev(src, torch.uint8)
    if layout == NHWC:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    if sums is None:
        sums = torch.zeros((b if per_frame else 1, c, 2), dtype=torch.int64, device=src.device)
    _check(lib.vacv_cuda_sums_u8(src.data_ptr(), b, w, h, c, layout, sums.data_ptr(), int(per_frame), _stream()))
    return sums


def finalize_mean_stddev(sums, n_per_channel):
    sums = _dev(sums, torch.int64)
    sets, c = sums.shape[0], sums.shape[1]
    mean = torch.empty((sets, c), dtype=torch.float32, device=sums.device)
    std = torch.empty((sets, c), dtype=torch.float32, device=sums.device)
    _check(lib.vacv_cuda_finalize_mean_stddev(sums.data_ptr(), sets, c, n_per_channel, mean.data_ptr(), std.data_ptr(), _stream()))
    return mean, std


def normalize(src, layout, mean, std, stats_per_frame=False, out=None):
    src = _dev(src)
    mean, std = _dev(mean, torch.float32), _dev(std, torch.float32)
    if layout == NHWC:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    dst = out if out is not None else torch.empty(src.shape, dtype=torch.float32, device=src.device)
    _check(lib.vacv_cuda_normalize(src.data_ptr(), dst.data_ptr(), b, w, h, c, _dt(src), layout, mean.data_ptr(),
                                   std.data_ptr(), int(stats_per_frame), _stream()))
    return dst


def nv_resize_normalize_chw(src, w, h, w_out, h_out, mean, std, v_first=True, out=None):
    """The fused config-2 pipeline.  src: uint8 [B, w*h*3/2] -> float32 [B, 3, h_out, w_out]."""
    src = _dev(src, torch.uint8)
    mean, std = _dev(mean, torch.float32), _dev(std, torch.float32)
    b = src.numel() // (w * h * 3 // 2)
    dst = out if out is not None else torch.empty((b, 3, h_out, w_out), dtype=torch.float32, device=src.device)
    _check(lib.vacv_cuda_nv_resize_normalize_chw(src.data_ptr(), dst.data_ptr(), b, w, h, int(bool(v_first)), w_out, h_out,
                                                 mean.data_ptr(), std.data_ptr(), _stream()))
    return dst


YUV_NV21, YUV_NV12, YUV_I420, YUV_YV12 = 0, 1, 2, 3


class YuvLayout(C.Structure):
    """Mirror of vacv_yuv_layout (include/vacv_cuda.h)."""
    _fields_ = [("format", C.c_int), ("w", C.c_int), ("h", C.c_int), ("y_pitch", C.c_int), ("c_pitch", C.c_int),
                ("frame_stride", C.c_size_t)]


def _torch_out_dtype(out_dtype):
    return {FP32: torch.float32, FP16: torch.float16, BF16: torch.bfloat16}[out_dtype]


def _yuv_batch(src, fmt, w, h, y_pitch, c_pitch, frame_stride):
    yp = y_pitch or w
    cp = c_pitch or (w // 2 if fmt >= YUV_I420 else w)
    per = frame_stride or (yp * h + cp * (h // 2) * (2 if fmt >= YUV_I420 else 1))
    return src.numel() // per


def yuv_resize_normalize_chw(src, fmt, w, h, w_out, h_out, mean, std, y_pitch=0, c_pitch=0, frame_stride=0, batch=None,
                             half=False, out=None, out_dtype=None):
    """Fused pipeline on pitched / planar decoder surfaces; fp32, fp16 or bf16 CHW planes out.  src: uint8 device tensor
    holding `batch` frames `frame_stride` bytes apart (dense when 0)."""
    src = _dev(src, torch.uint8)
    mean, std = _dev(mean, torch.float32), _dev(std, torch.float32)
    lay = YuvLayout(fmt, w, h, y_pitch, c_pitch, frame_stride)
    if batch is None:
        batch = _yuv_batch(src, fmt, w, h, y_pitch, c_pitch, frame_stride)
    if out_dtype is None:
        out_dtype = FP16 if half else FP32
    dst = out if out is not None else torch.empty((batch, 3, h_out, w_out), dtype=_torch_out_dtype(out_dtype), device=src.device)
    _check(lib.vacv_cuda_yuv_resize_normalize_chw(src.data_ptr(), C.addressof(lay), dst.data_ptr(), out_dtype, batch,
                                                  w_out, h_out, mean.data_ptr(), std.data_ptr(), _stream()))
    return dst


def cvt_yuv2bgr(src, fmt, w, h, y_pitch=0, c_pitch=0, frame_stride=0, batch=None):
    """Colour conversion of pitched / planar surfaces -> dense HWC BGR uint8 [batch, h, w, 3]."""
    src = _dev(src, torch.uint8)
    lay = YuvLayout(fmt, w, h, y_pitch, c_pitch, frame_stride)
    if batch is None:
        batch = _yuv_batch(src, fmt, w, h, y_pitch, c_pitch, frame_stride)
    dst = torch.empty((batch, h, w, 3), dtype=torch.uint8, device=src.device)
    _check(lib.vacv_cuda_cvt_yuv2bgr(src.data_ptr(), C.addressof(lay), dst.data_ptr(), batch, _stream()))
    return dst


class Rect(C.Structure):
    """Mirror of vacv_rect."""
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int)]


def letterbox_rect(w, h, canvas_w, canvas_h):
    r = Rect()
    lib.vacv_letterbox_rect(w, h, canvas_w, canvas_h, C.addressof(r))
    return r.x, r.y, r.w, r.h


def yuv_letterbox_normalize_chw(src, fmt, w, h, canvas_w, canvas_h, mean, std, pad_bgr=(114, 114, 114), content=None, y_pitch=0,
                                c_pitch=0, frame_stride=0, batch=None, out_dtype=FP32, out=None):
    """Aspect-preserving variant: the frame is resized into `content` (x, y, w, h; default vacv_letterbox_rect) of a
    canvas_w x canvas_h canvas filled with pad_bgr, then normalised -> CHW planes.  mean/std: 3-element host sequences."""
    src = _dev(src, torch.uint8)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    dm = torch.tensor(list(m), dtype=torch.float32, device=src.device)
    ds = torch.tensor(list(s), dtype=torch.float32, device=src.device)
    lay = YuvLayout(fmt, w, h, y_pitch, c_pitch, frame_stride)
    if batch is None:
        batch = _yuv_batch(src, fmt, w, h, y_pitch, c_pitch, frame_stride)
    r = Rect(*(content if content is not None else letterbox_rect(w, h, canvas_w, canvas_h)))
    pad = (C.c_uint8 * 3)(*[int(v) for v in pad_bgr])
    dst = out if out is not None else torch.empty((batch, 3, canvas_h, canvas_w), dtype=_torch_out_dtype(out_dtype), device=src.device)
    _check(lib.vacv_cuda_yuv_letterbox_normalize_chw(src.data_ptr(), C.addressof(lay), dst.data_ptr(), out_dtype, batch, canvas_w, canvas_h,
                                                     C.addressof(r), C.cast(pad, _vp), dm.data_ptr(), ds.data_ptr(), C.cast(m, _vp),
                                                     C.cast(s, _vp), _stream()))
    return dst


def resize_normalize(src, w_out, h_out, mean, std, out_layout=NHWC):
    src = _dev(src, torch.uint8)
    mean, std = _dev(mean, torch.float32), _dev(std, torch.float32)
    b, h, w, c = src.shape
    dst = torch.empty(_shape(out_layout, b, w_out, h_out, c), dtype=torch.float32, device=src.device)
    _check(lib.vacv_cuda_resize_normalize(src.data_ptr(), dst.data_ptr(), b, w, h, c, w_out, h_out, mean.data_ptr(),
                                          std.data_ptr(), out_layout, _stream()))
    return dst


def warp_affine_normalize(frames, minv, w_out, h_out, mean, std, frame_idx=None, out_layout=NHWC):
    frames = _dev(frames, torch.uint8)
    minv = _dev(minv, torch.float32)
    mean, std = _dev(mean, torch.float32), _dev(std, torch.float32)
    f, h, w, c = frames.shape
    n = minv.numel() // 6
    idx = _dev(frame_idx, torch.int32) if frame_idx is not None else None
    dst = torch.empty(_shape(out_layout, n, w_out, h_out, c), dtype=torch.float32, device=frames.device)
    _check(lib.vacv_cuda_warp_affine_normalize(frames.data_ptr(), f, w, h, c, idx.data_ptr() if idx is not None else None,
                                               minv.data_ptr(), n, dst.data_ptr(), w_out, h_out, mean.data_ptr(),
                                               std.data_ptr(), out_layout, _stream()))
    return dst


def nv_resize_normalize_chw_host(h_src, h_out, w, h, w_out, h_out_, mean, std, v_first=True, chunk_frames=32):
    """End-to-end host-buffer call: h_src uint8 [B, w*h*3/2] and h_out float32 [B,3,h_out,w_out] are HOST tensors
    (pinned for full PCIe speed); mean/std are 3-element host sequences.  Synchronous."""
    if h_src.is_cuda or h_out.is_cuda:
        raise VacvError("nv_resize_normalize_chw_host takes host tensors")
    b = h_src.numel() // (w * h * 3 // 2)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    _check(lib.vacv_cuda_nv_resize_normalize_chw_host(h_src.data_ptr(), h_out.data_ptr(), b, w, h, int(bool(v_first)), w_out, h_out_,
                                                      C.cast(m, _vp), C.cast(s, _vp), chunk_frames))
    return h_out


def yuv_normalize_chw_host(h_src, h_out, fmt, w, h, canvas_w, canvas_h, mean, std, content=None, pad_bgr=(114, 114, 114), y_pitch=0,
                           c_pitch=0, frame_stride=0, batch=None, out_dtype=FP32, chunk_frames=8):
    """Host-buffer form of yuv_resize_normalize_chw (content None) / yuv_letterbox_normalize_chw: host surfaces in, host planes
    (fp32 / fp16 / bf16, dtype of h_out must match out_dtype) out.  Synchronous."""
    if h_src.is_cuda or h_out.is_cuda:
        raise VacvError("yuv_normalize_chw_host takes host tensors")
    if h_out.dtype != _torch_out_dtype(out_dtype):
        raise VacvError("h_out dtype does not match out_dtype")
    lay = YuvLayout(fmt, w, h, y_pitch, c_pitch, frame_stride)
    if batch is None:
        batch = _yuv_batch(h_src, fmt, w, h, y_pitch, c_pitch, frame_stride)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    pad = (C.c_uint8 * 3)(*[int(v) for v in pad_bgr])
    r = Rect(*content) if content is not None else None
    _check(lib.vacv_cuda_yuv_normalize_chw_host(h_src.data_ptr(), C.addressof(lay), h_out.data_ptr(), out_dtype, batch, canvas_w, canvas_h,
                                                C.addressof(r) if r is not None else None, C.cast(pad, _vp), C.cast(m, _vp), C.cast(s, _vp),
                                                chunk_frames))
    return h_out


def sums_f32(src, layout, per_frame=False):
    """fp64 per-channel (sum x, sum x^2) of fp32 pixels: float64 [sets, c, 2]."""
    src = _dev(src, torch.float32)
    if layout == NHWC:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    sums = torch.zeros((b if per_frame else 1, c, 2), dtype=torch.float64, device=src.device)
    _check(lib.vacv_cuda_sums_f32(src.data_ptr(), b, w, h, c, layout, sums.data_ptr(), int(per_frame), _stream()))
    return sums


def finalize_mean_stddev_f64(sums, n_per_channel):
    sums = _dev(sums, torch.float64)
    sets, c = sums.shape[0], sums.shape[1]
    mean = torch.empty((sets, c), dtype=torch.float32, device=sums.device)
    std = torch.empty((sets, c), dtype=torch.float32, device=sums.device)
    _check(lib.vacv_cuda_finalize_mean_stddev_f64(sums.data_ptr(), sets, c, n_per_channel, mean.data_ptr(), std.data_ptr(), _stream()))
    return mean, std
