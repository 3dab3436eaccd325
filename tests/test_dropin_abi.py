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
