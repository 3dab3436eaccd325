// Per-call latency of the reference's public API (va_cv::*, vision::Tensor; host images in, host images out) on single
// images, the way the reference's own tests and its author's camera pipeline call it.  The SAME source is linked twice:
//   bench_dropin_b200  against arm-neon-opencv_b200/libvacv.so   (H2D copy + CUDA kernel + D2H copy inside every call)
//   bench_dropin_ref   against oracle/_ref/liboracle_ref.so      (the unmodified reference CPU implementation)
// so the two columns are what a caller sees before / after re-linking.  Prints one JSON object per operator.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <vector>

#include "common/tensor.h"
#include "cv/cv.h"

using namespace vision;

static double median_ms(const std::function<void()>& fn, int warm, int iters) {
    for (int i = 0; i < warm; ++i) fn();
    std::vector<double> t(iters);
    for (int i = 0; i < iters; ++i) {
        auto a = std::chrono::steady_clock::now();
        fn();
        auto b = std::chrono::steady_clock::now();
        t[i] = std::chrono::duration<double, std::milli>(b - a).count();
    }
    std::sort(t.begin(), t.end());
    return t[iters / 2];
}

int main(int argc, char** argv) {
    const char* impl = argc > 1 ? argv[1] : "?";
    const int iters = argc > 2 ? std::atoi(argv[2]) : 30;
    uint32_t seed = 1;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (uint8_t)(seed >> 24); };
    const int w = 1920, h = 1080;
    std::vector<uint8_t> bgr((size_t)w * h * 3), nv((size_t)w * h * 3 / 2), big((size_t)2560 * 1440 * 3);
    for (auto& x : bgr) x = rnd();
    for (auto& x : nv) x = rnd();
    for (auto& x : big) x = rnd();
    Tensor t_bgr(w, h, 3, bgr.data(), INT8, NHWC), t_nv(w, h * 3 / 2, 1, nv.data(), INT8, NCHW), t_big(2560, 1440, 3, big.data(), INT8, NHWC);
    const float mean_v[3] = {103.53f, 116.28f, 123.675f}, std_v[3] = {57.375f, 57.12f, 58.395f};
    Tensor mean(3, 1, 1, (void*)mean_v, FP32, NCHW), stddev(3, 1, 1, (void*)std_v, FP32, NCHW);
    Tensor small;
    va_cv::resize(t_bgr, small, va_cv::VSize(640, 640));
    Tensor small_f = small.change_dtype(FP32);

    struct Case { const char* name; std::function<void()> fn; };
    Tensor d0, d1, d2, d3, d4, d5, d6, d7;
    std::vector<Case> cases = {
        {"cvt_color NV21 1080p -> BGR", [&] { va_cv::cvt_color(t_nv, d0, va_cv::COLOR_YUV2BGR_NV21); }},
        {"resize linear u8 1080p -> 640x360 (config 1)", [&] { va_cv::resize(t_bgr, d1, va_cv::VSize(640, 360)); }},
        {"resize linear u8 1080p -> 640x640", [&] { va_cv::resize(t_bgr, d2, va_cv::VSize(640, 640)); }},
        {"change_dtype u8 -> fp32 640x640x3", [&] { d3 = small.change_dtype(FP32); }},
        {"normalize fp32 640x640x3 (given mean/std)", [&] { va_cv::normalize(small_f, d4, mean, stddev); }},
        {"change_layout HWC -> CHW fp32 640x640x3", [&] { d5 = small_f.change_layout(NCHW); }},
        {"chain: cvt_color -> resize -> dtype -> normalize -> layout (config 2, 1 frame)", [&] {
             Tensor a, b, c, e;
             va_cv::cvt_color(t_nv, a, va_cv::COLOR_YUV2BGR_NV21);
             va_cv::resize(a, b, va_cv::VSize(640, 640));
             c = b.change_dtype(FP32);
             va_cv::normalize(c, e, mean, stddev);
             d6 = e.change_layout(NCHW);
         }},
        {"crop u8 1080p -> 1280x720", [&] { vision::VRect r(321, 181, 321 + 1280, 181 + 720); va_cv::crop(t_bgr, d7, r); }},
    };
    for (auto& c : cases) {
        const double ms = median_ms(c.fn, 3, iters);
        std::printf("{\"impl\": \"%s\", \"op\": \"%s\", \"ms_per_call\": %.3f}\n", impl, c.name, ms);
        std::fflush(stdout);
    }
    return 0;
}
