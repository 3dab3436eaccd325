"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the symbols
include/vacv_cuda.h declares; argument validation and the host-side matrix helpers work without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vacv_cuda.h")
LIB = os.path.join(ROOT, "arm-neon-opencv_b200", "libvacv_cuda.so")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"VACV_API\s+[\w\s\*]+?\b(vacv_\w+)\s*\(", text)))


def test_header_declares_the_operator_set():
    syms = declared_symbols()
    for op in ["cvt_nv2bgr", "crop", "layout_change", "dtype_change", "resize", "warp_affine", "sums_u8",
               "finalize_mean_stddev", "normalize", "nv_resize_normalize_chw", "resize_normalize",
               "warp_affine_normalize"]:
        assert f"vacv_cuda_{op}" in syms
    assert "vacv_invert_affine" in syms and "vacv_rotation_matrix" in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    stray = [s for s in exported if not s.startswith("vacv_")]
    assert not stray, f"non-API symbols leak from the library: {stray}"


def test_dist_library_exports_what_vacv_dist_h_declares_and_validates_arguments():
    """libvacv_dist.so (NCCL transport of config 5): symbols of include/vacv_dist.h, NCCL reachable, argument checks before any
    CUDA / NCCL call; libvacv_cuda.so itself must stay free of an NCCL dependency."""
    dist_lib = os.path.join(ROOT, "arm-neon-opencv_b200", "libvacv_dist.so")
    assert os.path.exists(dist_lib), "run __graft_entry__.build() first"
    text = open(os.path.join(ROOT, "include", "vacv_dist.h")).read()
    declared = sorted(set(re.findall(r"VACV_API\s+[\w\s\*]+?\b(vacv_\w+)\s*\(", text)))
    assert "vacv_cuda_normalize_batch_global" in declared and "vacv_dist_nccl_comm_create" in declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", dist_lib], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert sorted(exported) == declared
    needed = subprocess.check_output(["readelf", "-d", LIB], text=True)
    assert "nccl" not in needed.lower()
    import vacv_b200 as vacv
    d = vacv.dist_lib()
    ver = C.c_int(0)
    assert d.vacv_dist_nccl_version(C.byref(ver)) == 0 and ver.value >= 22700
    buf = C.create_string_buffer(256)
    p = C.cast(buf, C.c_void_p)
    assert d.vacv_cuda_normalize_batch_global(None, p, p, 1, 8, 8, 3, vacv.NHWC, p, p, None, None, None) == -1
    assert b"communicator" in vacv.lib.vacv_cuda_last_error()
    assert d.vacv_dist_allreduce_u64(None, p, 7, None) == -1
    assert d.vacv_dist_nccl_comm_create(p, 2, 5, p) == -1
    # the transport-agnostic entries of libvacv_cuda.so
    assert vacv.lib.vacv_cuda_normalize_batch_global_p2p(None, p, p, 1, 8, 8, 3, vacv.NHWC, p, p, None, None, None) == -1
    assert vacv.lib.vacv_cuda_normalize_batch_global_cb(None, None, p, p, 1, 8, 8, 3, vacv.NHWC, p, p, None, None, None) == -1
    assert vacv.lib.vacv_cuda_p2p_create(p, 99, 0, p) == -1 and b"ranks" in vacv.lib.vacv_cuda_last_error()


def test_library_is_built_for_sm_100a_only():
    out = subprocess.check_output(["cuobjdump", "--list-elf", LIB], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_python_binding_covers_the_header():
    import vacv_b200 as vacv
    assert sorted(vacv.EXPORTS) == declared_symbols()
    assert vacv.lib.vacv_cuda_abi_version() == 1


def test_argument_validation_without_gpu():
    import vacv_b200 as vacv
    # validation happens before any CUDA call, so these return error codes even with no device
    assert vacv.lib.vacv_cuda_cvt_nv2bgr(None, None, 1, 16, 16, 1, None) == -1
    assert b"null" in vacv.lib.vacv_cuda_last_error()
    buf = C.create_string_buffer(64)
    p = C.cast(buf, C.c_void_p)
    assert vacv.lib.vacv_cuda_cvt_nv2bgr(p, p, 1, 15, 16, 1, None) == -1          # odd width
    assert vacv.lib.vacv_cuda_crop(p, p, 1, 10, 10, 3, vacv.INT8, vacv.NHWC, 5, 5, 10, 10, None) == -1   # rect outside
    assert vacv.lib.vacv_cuda_dtype_change(p, p, 8, vacv.FP16, vacv.FP32, None) == -2   # unsupported pair
    assert vacv.lib.vacv_cuda_resize(p, p, 1, 8, 8, 3, vacv.INT8, vacv.NHWC, 4, 4, 0, 0, None) == -2   # INTER_NEAREST


def test_tuning_switches_without_gpu():
    """vacv_cuda_set_tuning is host-only: every switch DESIGN.md lists is known by name, unknown names are rejected, and a switch can be
    set back (the kernels never read the environment on the call path)."""
    import vacv_b200 as vacv
    for name in (b"PIPE_NCOL", b"RPIPE_NCOL", b"WARP_GATHER", b"NO_RPIPE", b"RESIZE_NORMALIZE_GATHER", b"WALK_SEGS", b"WALK_SYNC", b"WALK2_SYNC",
                 b"CUBIC3", b"CUBIC_V", b"PIPE_ROWS", b"STREAM_QPT", b"WARP_V", b"LINEAR_V"):
        assert vacv.lib.vacv_cuda_set_tuning(name, 1) == 0, name
        assert vacv.lib.vacv_cuda_set_tuning(name, 0) == 0, name
    assert vacv.lib.vacv_cuda_set_tuning(b"NO_SUCH_SWITCH", 1) == -1
    assert b"unknown" in vacv.lib.vacv_cuda_last_error()


def test_host_matrix_helpers_match_oracle(oracle):
    import vacv_b200 as vacv
    m = [0.849158, 0.012257, -474.827, -0.01225, 0.849158, -379.18]   # test_warp_affine.cpp:31-32
    assert np.array_equal(np.array(vacv.invert_affine(m), np.float32).view(np.uint32), oracle.invert_affine(m).view(np.uint32))
    aux = [738.518372, 537.672852, 204.766998, 73.329681]
    got = np.array(vacv.rotation_matrix(1.073914, -3.314525, aux), np.float32)
    assert np.array_equal(got.view(np.uint32), oracle.rotation_matrix(1.073914, -3.314525, aux).view(np.uint32))
    assert np.array_equal(np.array(vacv.invert_affine([0, 0, 1, 0, 0, 1]), np.float32)[[0, 4]], [0, 0])   # singular: D = 0


def test_argument_validation_of_the_surface_entries_without_gpu():
    """The decoder-surface / letterbox entry points validate layout, output type and rectangle before any CUDA call."""
    import vacv_b200 as vacv
    buf = C.create_string_buffer(256)
    p = C.cast(buf, C.c_void_p)
    ok = vacv.YuvLayout(vacv.YUV_NV12, 64, 32, 0, 0, 0)
    odd = vacv.YuvLayout(vacv.YUV_NV12, 63, 32, 0, 0, 0)
    badfmt = vacv.YuvLayout(7, 64, 32, 0, 0, 0)
    small_pitch = vacv.YuvLayout(vacv.YUV_I420, 64, 32, 48, 0, 0)
    f = vacv.lib.vacv_cuda_yuv_resize_normalize_chw
    assert f(None, C.addressof(ok), p, vacv.FP32, 1, 32, 16, p, p, None) == -1
    assert f(p, C.addressof(odd), p, vacv.FP32, 1, 32, 16, p, p, None) == -1
    assert b"even" in vacv.lib.vacv_cuda_last_error()
    assert f(p, C.addressof(badfmt), p, vacv.FP32, 1, 32, 16, p, p, None) == -2
    assert f(p, C.addressof(small_pitch), p, vacv.FP32, 1, 32, 16, p, p, None) == -1
    assert f(p, C.addressof(ok), p, vacv.INT8, 1, 32, 16, p, p, None) == -2            # output type
    assert f(p, C.addressof(ok), p, vacv.FP32, 1, 64, 32, p, p, None) == -2            # same-size resize
    g = vacv.lib.vacv_cuda_yuv_letterbox_normalize_chw
    outside = vacv.Rect(10, 0, 60, 30)
    assert g(p, C.addressof(ok), p, vacv.FP32, 1, 64, 64, C.addressof(outside), p, p, p, p, p, None) == -1
    assert b"outside" in vacv.lib.vacv_cuda_last_error()
    assert vacv.lib.vacv_cuda_cvt_yuv2bgr(p, C.addressof(odd), p, 1, None) == -1
    assert vacv.lib.vacv_cuda_yuv_normalize_chw_host(p, C.addressof(ok), p, vacv.FP64, 1, 32, 16, None, None, p, p, 4) == -2
    assert vacv.letterbox_rect(1920, 1080, 640, 640) == (0, 140, 640, 360)           # pure host helper
