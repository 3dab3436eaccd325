"""CPU: libvacv.so (the C++ drop-in) is link- and layout-compatible with the reference.

* every `va_cv::*` / `vision::Tensor::*` / `vision::VRect::*` symbol the compiled reference exports is exported by
  libvacv.so under the same mangled name (needs oracle/_ref);
* `vision::Tensor` has the same size and member offsets as the reference's class (needs /root/reference headers);
* the drop-in fails loudly without a GPU (no CPU fallback)."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBVACV = os.path.join(ROOT, "arm-neon-opencv_b200", "libvacv.so")
CVMAT = os.path.join(ROOT, "arm-neon-opencv_b200", "libvacv_cvmat.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liboracle_ref.so")
REF_SRC = "/root/reference/src"


def exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return {l.split()[-1] for l in out.splitlines() if " T " in l or " W " in l}


def test_dropin_exports_reference_symbols():
    assert os.path.exists(LIBVACV), "run __graft_entry__.build() first"
    ours = exported(LIBVACV)
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    want = {s for s in exported(REF_SO)
            if s.startswith("_ZN5va_cv") and not any(cls in s for cls in
               ("6Resize", "8CvtColor", "9Normalize", "10WarpAffine", "4Crop", "11ResizeNaive", "14NormalizeNaive",
                "15WarpAffineNaive", "15ResizeNormalize", "19WarpAffineNormalize", "13MatchTemplate", "8ImEncode",
                "10CudaDevice", "8CropCuda", "10ResizeNeon", "13NormalizeNeon"))}      # public free functions only
    want |= {s for s in exported(REF_SO) if "N6vision6Tensor" in s and s.startswith(("_ZN6vision6Tensor", "_ZNK6vision6Tensor"))}
    want |= {s for s in exported(REF_SO) if s.startswith(("_ZN6vision5VRect", "_ZNK6vision5VRect"))}
    # TensorConverter<cv::Mat> (tensor_converter.cpp:15-83): the reference's test target links only `vacv` + OpenCV, so these two
    # must be linkable too; they live in libvacv_cvmat.so (built against the caller's OpenCV), libvacv.so stays OpenCV-free
    conv = {s for s in exported(REF_SO) if "15TensorConverter" in s}
    assert len(conv) == 2
    assert os.path.exists(CVMAT), "libvacv_cvmat.so not built (make -C oracle ut)"
    want |= conv
    ours |= exported(CVMAT)
    assert not any("opencv" in l for l in subprocess.check_output(["readelf", "-d", LIBVACV], text=True).splitlines() if "NEEDED" in l)
    assert len(want) > 40
    missing = sorted(want - ours)
    assert not missing, f"reference symbols missing from libvacv.so: {missing}"


LAYOUT_PROBE = r"""
#include <cstddef>
#include <cstdio>
#include "common/tensor.h"
#include "cv/cv.h"
struct Peek : vision::Tensor {};
int main() {
    using T = vision::Tensor;
    std::printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu | %zu %zu %zu\n", sizeof(T), offsetof(T, w), offsetof(T, h), offsetof(T, c),
                offsetof(T, stride), offsetof(T, dims), offsetof(T, data), offsetof(T, dtype), offsetof(T, layout),
                sizeof(va_cv::VSize), sizeof(va_cv::VScalar), sizeof(vision::VRect));
    std::printf("%d %d %d %d %d %d %d\n", (int)vision::FP32, (int)vision::INT8, (int)vision::NHWC, (int)va_cv::INTER_CUBIC,
                (int)va_cv::BORDER_CONSTANT, (int)va_cv::COLOR_YUV2BGR_NV12, (int)va_cv::COLOR_YUV2BGR_NV21);
}
"""


def probe(include_dir, workdir, tag):
    src = os.path.join(workdir, f"probe_{tag}.cpp")
    exe = os.path.join(workdir, f"probe_{tag}")
    open(src, "w").write(LAYOUT_PROBE)
    subprocess.check_call(["g++", "-std=c++14", "-Wno-invalid-offsetof", "-I", include_dir, src, "-o", exe])
    return subprocess.check_output([exe], text=True)


def test_tensor_layout_matches_reference():
    if not os.path.isdir(REF_SRC):
        pytest.skip("/root/reference not mounted")
    with tempfile.TemporaryDirectory() as d:
        assert probe(os.path.join(ROOT, "include", "vacv"), d, "ours") == probe(REF_SRC, d, "ref")


def test_dropin_fails_loudly_without_gpu():
    exe = os.path.join(ROOT, "tests", "cpp", "test_dropin")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/test_dropin not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stdout + r.stderr


HEADER_ONLY_PROBE = r"""
#include <opencv2/core/core.hpp>
#include "common/tensor_converter.h"
#include <cstdio>
int main() {
    unsigned char px[2 * 3 * 3] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18};
    cv::Mat m(2, 3, CV_8UC3, px);
    vision::Tensor t = vision::TensorConverter::convert_from<cv::Mat>(m);            // borrows
    vision::Tensor tc = vision::TensorConverter::convert_from<cv::Mat>(m, true);     // copies
    cv::Mat back = vision::TensorConverter::convert_to<cv::Mat>(tc, true);
    const bool ok = t.w == 3 && t.h == 2 && t.c == 3 && t.dtype == vision::INT8 && t.layout == vision::NHWC && t.data == px &&
                    tc.data != px && back.rows == 2 && back.cols == 3 && back.type() == CV_8UC3 && back.data[17] == 18 &&
                    vision::TensorConverter::convert_to<cv::Mat>(vision::Tensor()).empty();
    cv::Mat f(4, 5, CV_32FC1), d(1, 1, CV_64FC2), s16(1, 2, CV_16SC1);
    const bool types = vision::TensorConverter::convert_from<cv::Mat>(f).dtype == vision::FP32 &&
                       vision::TensorConverter::convert_from<cv::Mat>(d).dtype == vision::FP64 &&
                       vision::TensorConverter::convert_from<cv::Mat>(s16).dtype == vision::FP16;
    std::printf("%d %d\n", (int)ok, (int)types);
    return ok && types ? 0 : 1;
}
"""


def test_tensor_converter_header_instantiates_in_the_callers_tu():
    """include/vacv/common/tensor_converter.h is header-only: a caller TU built against the reference's bundled OpenCV 2.4 headers
    gets both specialisations (type map of tensor_converter.cpp:26-36,56-74) without any OpenCV dependency in libvacv.so."""
    ocv = "/root/reference/thirdparty/opencv_2.4.13.4/linux-x86_64"
    libdir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(ocv) or not os.path.exists(os.path.join(libdir, "libopencv_core.so.2.4")):
        pytest.skip("bundled OpenCV headers / binaries not available")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "conv.cpp"), os.path.join(d, "conv")
        open(src, "w").write(HEADER_ONLY_PROBE)
        pkg = os.path.join(ROOT, "arm-neon-opencv_b200")
        subprocess.check_call(["g++", "-std=c++14", "-I", os.path.join(ocv, "include"), "-I", os.path.join(ROOT, "include", "vacv"), src, "-o", exe,
                               "-L", pkg, "-l:libvacv.so", "-l:libvacv_cuda.so", "-L", libdir, "-l:libopencv_core.so.2.4",
                               f"-Wl,-rpath,{pkg}", f"-Wl,-rpath,{libdir}"])
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip() == "1 1", r.stdout + r.stderr
