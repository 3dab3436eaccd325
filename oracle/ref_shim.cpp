// TEST INFRASTRUCTURE ONLY -- flat C entry points over the *unmodified reference* (oracle/_ref).
//
// Everything here calls code compiled from /root/reference (its naive CPU path) or the
// OpenCV 2.4.13 binaries bundled with it.  It exists so that tests/ and bench.py's CPU arm can
// (a) pin the C restatement in oracle/vacv_oracle.c against the real thing and (b) time the
// reference's own CPU path.  The product (libvacv_cuda.so / libvacv.so) never links or loads it.
//
// dtype codes = vision::DType (tensor.h:12-18): FP32=0 FP16=1 INT8=2 ; layout = vision::DLayout
// (tensor.h:21-24): NCHW=0 NHWC=1.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#include "opencv2/core/core.hpp"
#include "opencv2/imgproc/imgproc.hpp"
#include "opencv2/highgui/highgui.hpp"

#include "common/tensor.h"
#include "common/vision_structs.h"
#include "cv/cv.h"
#include "cv/crop_cuda.h"
#include "cv/cuda_device.h"
#include "cv/normalize_naive.h"
#include "cv/resize_naive.h"
#include "cv/resize_neon.h"
#include "util/image_util.h"

using vision::DLayout;
using vision::DType;
using vision::Tensor;

// App. B shim 3: Crop::crop_cuda (crop.cpp:146-168) is compiled unconditionally and references these;
// they are never reached on the non-USE_CUDA path.
namespace va_cv {
int CudaDevice::get_device_count() { return 0; }
int CudaDevice::set_device(int) { return -1; }
void CropCuda::crop_cuda_chw_int8(const unsigned char*, int, int, int, unsigned char*, int, int, int, int) {}
void CropCuda::crop_cuda_rgb_hwc_int8(const unsigned char*, int, int, unsigned char*, int, int, int, int) {}
}  // namespace va_cv

namespace {
inline Tensor wrap(const void* p, int w, int h, int c, int dtype, int layout) {
    return Tensor(w, h, c, const_cast<void*>(p), static_cast<DType>(dtype), static_cast<DLayout>(layout));
}
inline int cv_depth(int dtype) { return dtype == vision::INT8 ? CV_8U : CV_32F; }
}  // namespace

extern "C" {

// va_cv::cvt_color (cv.cpp:22-24).  src = w x (h*3/2) single channel, dst = w x h x 3 HWC u8.
void ref_cvt_color(const uint8_t* src, int w, int h, int code, uint8_t* dst) {
    Tensor s = wrap(src, w, h / 2 * 3, 1, vision::INT8, vision::NCHW);
    Tensor d = wrap(dst, w, h, 3, vision::INT8, vision::NHWC);
    va_cv::cvt_color(s, d, code);
}

// ImageUtil::bgr2nv21 (image_util.cpp:9-40) -- the reference's NV21 fixture generator.
void ref_bgr2nv21(const uint8_t* bgr, uint8_t* nv21, int w, int h) {
    ImageUtil::bgr2nv21(const_cast<uint8_t*>(bgr), nv21, w, h);
}

// va_cv::crop (cv.cpp:75-77).
void ref_crop(const void* src, int w, int h, int c, int dtype, int layout,
              float left, float top, float right, float bottom, void* dst) {
    vision::VRect r(left, top, right, bottom);
    Tensor s = wrap(src, w, h, c, dtype, layout);
    Tensor d = wrap(dst, (int)r.width(), (int)r.height(), c, dtype, layout);
    va_cv::crop(s, d, r);
}

// Tensor::change_layout (tensor.cpp:393-457).
void ref_change_layout(const void* src, int w, int h, int c, int dtype, int layout, int new_layout, void* dst) {
    Tensor s = wrap(src, w, h, c, dtype, layout);
    Tensor t = s.change_layout(static_cast<DLayout>(new_layout));
    memcpy(dst, t.data, t.len());
}

// Tensor::change_dtype (tensor.cpp:459-502).
void ref_change_dtype(const void* src, int w, int h, int c, int dtype, int layout, int new_dtype, void* dst) {
    Tensor s = wrap(src, w, h, c, dtype, layout);
    Tensor t = s.change_dtype(static_cast<DType>(new_dtype));
    memcpy(dst, t.data, t.len());
}

// va_cv::resize (cv.cpp:16-20).  NB: INTER_CUBIC fp32 with w_out != h_out hits the scratch-buffer
// aliasing defect (resize_naive.cpp:537-538) -- use ref_resize_cubic_f32_fixed for that.
void ref_resize(const void* src, int w, int h, int c, int dtype, int layout,
                void* dst, int w_out, int h_out, int interpolation) {
    Tensor s = wrap(src, w, h, c, dtype, layout);
    Tensor d = wrap(dst, w_out, h_out, c, dtype, layout);
    va_cv::resize(s, d, va_cv::VSize(w_out, h_out), 0, 0, interpolation);
}

// The reference's NEON bilinear path, glue of Resize::resize_neon (resize.cpp:132-146) restated here because that
// function only exists in aarch64 builds; the kernels themselves are the reference's resize_neon.cpp compiled against
// oracle/neon_emul/arm_neon.h.  NHWC: 3 channels, called with tripled widths exactly like the glue; NCHW: 3 planes.
void ref_resize_neon(const uint8_t* src, int w, int h, int layout, uint8_t* dst, int w_out, int h_out) {
    if (layout == vision::NHWC) {
        va_cv::ResizeNeon::resize_neon_inter_linear_three_channel(src, w * 3, h, dst, w_out * 3, h_out);
    } else {
        for (int i = 0; i < 3; i++)
            va_cv::ResizeNeon::resize_neon_inter_linear_one_channel(src + (size_t)w * h * i, w, h, dst + (size_t)w_out * h_out * i, w_out, h_out);
    }
}

// The reference's own cubic building blocks (resize_naive.cpp:143-529) driven with *non-aliased*
// coefficient buffers: the intended behaviour of resize_naive_inter_cubic_fp32_{hwc,chw} (App. C-2).
void ref_resize_cubic_f32_fixed(const float* src, int w, int h, int c, int layout,
                                float* dst, int w_out, int h_out) {
    std::vector<int> xofs(w_out), yofs(h_out);
    std::vector<float> alpha(w_out * 4), beta(h_out * 4);
    va_cv::ResizeNaive::cubic_coeffs_naive(w, w_out, xofs.data(), alpha.data());
    va_cv::ResizeNaive::cubic_coeffs_naive(h, h_out, yofs.data(), beta.data());
    if (layout == vision::NHWC) {
        va_cv::ResizeNaive::resize_naive_inter_cubic_fp32_three_channel(
            const_cast<float*>(src), w, h, dst, w_out, h_out, alpha.data(), xofs.data(), beta.data(), yofs.data());
    } else {
        for (int k = 0; k < c; ++k) {
            va_cv::ResizeNaive::resize_naive_inter_cubic_fp32_one_channel(
                const_cast<float*>(src) + (size_t)w * h * k, w, h, dst + (size_t)w_out * h_out * k, w_out, h_out,
                alpha.data(), xofs.data(), beta.data(), yofs.data());
        }
    }
}

// Bundled OpenCV 2.4.13 cv::resize -- the only executable reference for u8 INTER_CUBIC
// (resize.cpp:33-36 is dead without USE_OPENCV).  HWC only.
void ref_cv_resize(const void* src, int w, int h, int c, int dtype, void* dst, int w_out, int h_out,
                   int interpolation, int threads) {
    cv::setNumThreads(threads);
    cv::Mat s(h, w, CV_MAKETYPE(cv_depth(dtype), c), const_cast<void*>(src));
    cv::Mat d(h_out, w_out, CV_MAKETYPE(cv_depth(dtype), c), dst);
    cv::resize(s, d, cv::Size(w_out, h_out), 0, 0, interpolation);
}

// va_cv::warp_affine, matrix overload (cv.cpp:31-35).  m[6] is inverted IN PLACE like the reference
// does to the caller's tensor (warp_affine.cpp:121-133).  dst must be pre-zeroed by the caller:
// out-of-bounds destination pixels are left untouched (warp_affine_naive.cpp:28-30,36-38).
void ref_warp_affine(const void* src, int w, int h, int c, int dtype, int layout,
                     float* m, void* dst, int w_out, int h_out) {
    Tensor s = wrap(src, w, h, c, dtype, layout);
    Tensor d = wrap(dst, w_out, h_out, c, dtype, layout);
    Tensor M = wrap(m, 3, 2, 1, vision::FP32, vision::NCHW);
    va_cv::warp_affine(s, d, M, va_cv::VSize(w_out, h_out));
}

// va_cv::warp_affine, scale/rot overload (cv.cpp:37-42).  m_out receives nothing (the reference
// builds and inverts a private matrix).
void ref_warp_affine_rot(const void* src, int w, int h, int c, int dtype, int layout,
                         float scale, float rot, const double* aux4, void* dst, int w_out, int h_out) {
    Tensor s = wrap(src, w, h, c, dtype, layout);
    Tensor d = wrap(dst, w_out, h_out, c, dtype, layout);
    va_cv::VScalar aux;
    aux.v0 = aux4[0]; aux.v1 = aux4[1]; aux.v2 = aux4[2]; aux.v3 = aux4[3];
    va_cv::warp_affine(s, d, scale, rot, va_cv::VSize(w_out, h_out), aux);
}

// va_cv::normalize (cv.cpp:26-29).  mean/stddev == nullptr -> auto statistics (normalize.cpp:98-108).
void ref_normalize(const void* src, int w, int h, int c, int dtype, int layout,
                   const float* mean, const float* stddev, float* dst) {
    Tensor s = wrap(src, w, h, c, dtype, layout);
    Tensor d = wrap(dst, w, h, c, vision::FP32, layout);
    if (mean && stddev) {
        Tensor m = wrap(mean, c, 1, 1, vision::FP32, vision::NCHW);
        Tensor sd = wrap(stddev, c, 1, 1, vision::FP32, vision::NCHW);
        va_cv::normalize(s, d, m, sd);
    } else {
        va_cv::normalize(s, d);
    }
}

// NormalizeNaive::mean_stddev_naive_{hwc_bgr,chw} (normalize_naive.cpp:7-72): sequential fp32.
void ref_mean_stddev_f32(const float* src, int w, int h, int c, int layout, float* mean, float* stddev) {
    if (layout == vision::NHWC) {
        va_cv::NormalizeNaive::mean_stddev_naive_hwc_bgr(const_cast<float*>(src), w * h, mean, stddev);
    } else {
        va_cv::NormalizeNaive::mean_stddev_naive_chw(const_cast<float*>(src), w * h, c, mean, stddev);
    }
}

// Bundled cv::meanStdDev -- the truth the reference's own test uses (test_normalize.cpp:31).  HWC.
void ref_cv_mean_stddev(const void* src, int w, int h, int c, int dtype, double* mean, double* stddev) {
    cv::Mat s(h, w, CV_MAKETYPE(cv_depth(dtype), c), const_cast<void*>(src));
    cv::Scalar m, sd;
    cv::meanStdDev(s, m, sd);
    for (int k = 0; k < c && k < 4; ++k) { mean[k] = m[k]; stddev[k] = sd[k]; }
}

// Bundled cv::warpAffine (cross-check only; the reference's tests compare against it).
void ref_cv_warp_affine(const void* src, int w, int h, int c, int dtype, const float* m,
                        void* dst, int w_out, int h_out) {
    cv::Mat s(h, w, CV_MAKETYPE(cv_depth(dtype), c), const_cast<void*>(src));
    cv::Mat d(h_out, w_out, CV_MAKETYPE(cv_depth(dtype), c), dst);
    cv::Mat M(2, 3, CV_32FC1, const_cast<float*>(m));
    cv::warpAffine(s, d, M, cv::Size(w_out, h_out), cv::INTER_LINEAR, cv::BORDER_CONSTANT, cv::Scalar(0, 0, 0, 0));
}

// JPEG decode through the *bundled* highgui (other decoders differ by a few LSBs).  Returns 0 on
// success; *w,*h,*c receive the shape; if dst != nullptr it receives w*h*c bytes (BGR HWC).
int ref_imread(const char* path, int color, uint8_t* dst, int* w, int* h, int* c) {
    cv::Mat m = cv::imread(path, color ? 1 : 0);
    if (m.empty()) return 1;
    *w = m.cols; *h = m.rows; *c = m.channels();
    if (dst) memcpy(dst, m.data, (size_t)m.cols * m.rows * m.channels());
    return 0;
}

// The unfused config-2 chain exactly as a reference user would write it:
// cvt_color -> resize(INTER_LINEAR) -> normalize(mean,std) [does change_dtype] -> change_layout(NCHW).
void ref_pipeline_nv_resize_norm_chw(const uint8_t* src, int w, int h, int code, int w_out, int h_out,
                                     const float* mean, const float* stddev, float* dst) {
    Tensor s = wrap(src, w, h / 2 * 3, 1, vision::INT8, vision::NCHW);
    Tensor bgr, small, norm;
    va_cv::cvt_color(s, bgr, code);
    va_cv::resize(bgr, small, va_cv::VSize(w_out, h_out));
    Tensor m = wrap(mean, 3, 1, 1, vision::FP32, vision::NCHW);
    Tensor sd = wrap(stddev, 3, 1, 1, vision::FP32, vision::NCHW);
    va_cv::normalize(small, norm, m, sd);
    Tensor chw = norm.change_layout(vision::NCHW);
    memcpy(dst, chw.data, chw.len());
}

// Batch driver for CPU-baseline timing: frames are independent, so the harness (not the reference)
// may spread them over host threads (SURVEY 8d "CPU reference alongside" (ii)).
void ref_pipeline_nv_resize_norm_chw_batch(const uint8_t* src, int n, int w, int h, int code, int w_out,
                                           int h_out, const float* mean, const float* stddev, float* dst,
                                           int threads) {
    const size_t in_stride = (size_t)w * h * 3 / 2, out_stride = (size_t)w_out * h_out * 3;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1))
            ref_pipeline_nv_resize_norm_chw(src + in_stride * i, w, h, code, w_out, h_out, mean, stddev,
                                            dst + out_stride * i);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}


// SURVEY 8(d) "CPU reference alongside": wall-clock seconds for `threads` host threads to run `reps` frames each of one
// config through the reference's public API (each thread owns its tensors; frames are independent).  The harness, not the
// reference, provides the threads.  cfg: 1 = resize linear u8 1080p -> 640x360, 3 = warp_affine 720p -> 112x112 + fp32 +
// normalize, 4 = cv::resize u8 cubic 1440p -> 1080p (bundled OpenCV 2.4, the reference's only u8 cubic), 5 = normalize u8 4K
// with its own mean/stddev.
double ref_time_config(int cfg, int threads, int reps) {
    auto fill = [](std::vector<uint8_t>& v, uint32_t seed) { for (auto& x : v) { seed = seed * 1664525u + 1013904223u; x = (uint8_t)(seed >> 24); } };
    const float mean_v[3] = {103.53f, 116.28f, 123.675f}, std_v[3] = {57.375f, 57.12f, 58.395f};
    cv::setNumThreads(1);
    auto work = [&](int tid) {
        if (cfg == 1) {
            std::vector<uint8_t> img((size_t)1920 * 1080 * 3); fill(img, 1 + tid);
            Tensor s = wrap(img.data(), 1920, 1080, 3, vision::INT8, vision::NHWC), d;
            for (int i = 0; i < reps; ++i) va_cv::resize(s, d, va_cv::VSize(640, 360));
        } else if (cfg == 3) {
            std::vector<uint8_t> img((size_t)1280 * 720 * 3); fill(img, 3 + tid);
            Tensor s = wrap(img.data(), 1280, 720, 3, vision::INT8, vision::NHWC), d, n;
            Tensor m = wrap(mean_v, 3, 1, 1, vision::FP32, vision::NCHW), sd = wrap(std_v, 3, 1, 1, vision::FP32, vision::NCHW);
            for (int i = 0; i < reps; ++i) {
                float mat[6] = {0.4f, 0.05f, -100.f, -0.05f, 0.4f, -20.f};   // warp_affine inverts it in place
                Tensor M = wrap(mat, 3, 2, 1, vision::FP32, vision::NCHW);
                va_cv::warp_affine(s, d, M, va_cv::VSize(112, 112));
                va_cv::normalize(d, n, m, sd);
            }
        } else if (cfg == 4) {
            std::vector<uint8_t> img((size_t)2560 * 1440 * 3), out((size_t)1920 * 1080 * 3); fill(img, 4 + tid);
            cv::Mat sm(1440, 2560, CV_8UC3, img.data()), dm(1080, 1920, CV_8UC3, out.data());
            for (int i = 0; i < reps; ++i) cv::resize(sm, dm, cv::Size(1920, 1080), 0, 0, cv::INTER_CUBIC);
        } else if (cfg == 5) {
            std::vector<uint8_t> img((size_t)3840 * 2160 * 3); fill(img, 5 + tid);
            Tensor s = wrap(img.data(), 3840, 2160, 3, vision::INT8, vision::NHWC), n;
            for (int i = 0; i < reps; ++i) va_cv::normalize(s, n);
        }
    };
    std::vector<std::thread> pool;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
