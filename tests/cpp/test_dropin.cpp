// Drop-in test: a caller written against the reference's headers (vision::Tensor, va_cv::*) linked with
// libvacv.so, checked element-wise against the oracle (oracle/vacv_oracle.c, linked into THIS test binary only).
// Mirrors the call patterns of the reference's own tests (src/test/src/impl/test_*.cpp) with their known-answer
// parameters; pass criterion is bit equality (the reference's cosine>=1-5e-4 is far looser).
// Needs a GPU.  Exit code = number of failed cases.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <algorithm>
#include <vector>

#include "common/tensor.h"
#include "cv/cv.h"
#include "vacv_cuda.h"
#include "vacv_oracle.h"

using namespace vision;

static int g_failed = 0, g_run = 0;

static void report(const char* name, bool ok) {
    ++g_run;
    if (!ok) ++g_failed;
    std::printf("%-52s %s\n", name, ok ? "PASS" : "FAIL");
}

static uint32_t g_seed = 12345;
static uint8_t rnd8() { g_seed = g_seed * 1664525u + 1013904223u; return (uint8_t)(g_seed >> 24); }
static std::vector<uint8_t> random_u8(size_t n) { std::vector<uint8_t> v(n); for (auto& x : v) x = rnd8(); return v; }
static std::vector<float> to_f32(const std::vector<uint8_t>& v) { return std::vector<float>(v.begin(), v.end()); }

template <typename T>
static bool same(const Tensor& t, const std::vector<T>& want) {
    return t.len() == want.size() * sizeof(T) && std::memcmp(t.data, want.data(), t.len()) == 0;
}

int main() {
    const int w = 640, h = 360, c = 3;
    std::vector<uint8_t> img = random_u8((size_t)w * h * c);
    std::vector<float> imgf = to_f32(img);
    Tensor src_u8(w, h, c, img.data(), INT8, NHWC);      // borrowing ctor, like TensorConverter::convert_from
    Tensor src_f32(w, h, c, imgf.data(), FP32, NHWC);
    const float mean_v[3] = {103.53f, 116.28f, 123.675f}, std_v[3] = {57.375f, 57.12f, 58.395f};
    Tensor mean(3, 1, 1, (void*)mean_v, FP32, NCHW), stddev(3, 1, 1, (void*)std_v, FP32, NCHW);

    try {
        {   // test_resize.cpp: bilinear HWC u8 / fp32, CHW u8
            Tensor dst;
            va_cv::resize(src_u8, dst, va_cv::VSize(320, 180));
            std::vector<uint8_t> want(320 * 180 * 3);
            orc_resize_linear_u8(img.data(), w, h, c, 1, want.data(), 320, 180, 0);
            report("resize INTER_LINEAR hwc u8 -> 320x180", dst.w == 320 && dst.h == 180 && dst.c == 3 && dst.dtype == INT8 &&
                                                              dst.layout == NHWC && same(dst, want));
            Tensor dstf;
            va_cv::resize(src_f32, dstf, va_cv::VSize(200, 111), 0, 0, va_cv::INTER_LINEAR);
            std::vector<float> wantf(200 * 111 * 3);
            orc_resize_linear_f32(imgf.data(), w, h, c, 1, wantf.data(), 200, 111);
            report("resize INTER_LINEAR hwc fp32 -> 200x111", same(dstf, wantf));
            Tensor chw = src_u8.change_layout(NCHW), dchw;
            va_cv::resize(chw, dchw, va_cv::VSize(320, 180));
            std::vector<uint8_t> in_chw(img.size()), want_chw(320 * 180 * 3);
            orc_hwc_to_chw(img.data(), w, h, c, 1, in_chw.data());
            orc_resize_linear_u8(in_chw.data(), w, h, c, 0, want_chw.data(), 320, 180, 0);
            report("change_layout + resize INTER_LINEAR chw u8", same(chw, in_chw) && dchw.layout == NCHW && same(dchw, want_chw));
            Tensor dcub;
            va_cv::resize(src_f32, dcub, va_cv::VSize(300, 170), 0, 0, va_cv::INTER_CUBIC);
            std::vector<float> wantc(300 * 170 * 3);
            orc_resize_cubic_f32(imgf.data(), w, h, c, 1, wantc.data(), 300, 170);
            report("resize INTER_CUBIC hwc fp32 -> 300x170", same(dcub, wantc));
            Tensor dcub8;
            va_cv::resize(src_u8, dcub8, va_cv::VSize(480, 270), 0, 0, va_cv::INTER_CUBIC);
            std::vector<uint8_t> wantc8(480 * 270 * 3);
            orc_resize_cubic_u8_cv24(img.data(), w, h, c, wantc8.data(), 480, 270);
            report("resize INTER_CUBIC hwc u8 (OpenCV-2.4 rule)", same(dcub8, wantc8));
            Tensor reuse(320, 180, 3, INT8, NHWC);   // pre-created dst of the right shape is reused, not reallocated
            void* before = reuse.data;
            va_cv::resize(src_u8, reuse, va_cv::VSize(320, 180));
            report("resize reuses a matching dst buffer", reuse.data == before && same(reuse, want));
        }
        {   // test_cvt_color.cpp: BGR -> bgr2nv21 -> cvt_color
            std::vector<uint8_t> nv((size_t)w * h * 3 / 2);
            orc_bgr_to_nv21(img.data(), w, h, nv.data());
            Tensor nvt(w, h * 3 / 2, 1, nv.data(), INT8, NCHW), bgr21, bgr12;
            va_cv::cvt_color(nvt, bgr21, va_cv::COLOR_YUV2BGR_NV21);
            va_cv::cvt_color(nvt, bgr12, va_cv::COLOR_YUV2BGR_NV12);
            std::vector<uint8_t> want((size_t)w * h * 3);
            orc_nv_to_bgr(nv.data(), w, h, 1, want.data());
            report("cvt_color NV21 -> BGR", bgr21.w == w && bgr21.h == h && bgr21.c == 3 && bgr21.layout == NHWC && same(bgr21, want));
            report("cvt_color NV12 code decodes V-first like the reference", same(bgr12, want));
            // COLOR_YUV2BGR_YV12 (declared at cv.h:73, no reference implementation): the same samples as planar Y, V, U
            std::vector<uint8_t> yv((size_t)w * h * 3 / 2);
            std::copy(nv.begin(), nv.begin() + (size_t)w * h, yv.begin());
            for (size_t i = 0; i < (size_t)w * h / 4; ++i) {
                yv[(size_t)w * h + i] = nv[(size_t)w * h + 2 * i];                          // V plane
                yv[(size_t)w * h + (size_t)w * h / 4 + i] = nv[(size_t)w * h + 2 * i + 1];  // U plane
            }
            Tensor yvt(w, h * 3 / 2, 1, yv.data(), INT8, NCHW), bgryv;
            va_cv::cvt_color(yvt, bgryv, va_cv::COLOR_YUV2BGR_YV12);
            report("cvt_color YV12 (planar) -> BGR", same(bgryv, want));
        }
        {   // test_change_dtype.cpp
            Tensor f = src_u8.change_dtype(FP32);
            report("change_dtype u8 -> fp32", f.dtype == FP32 && same(f, imgf));
            Tensor b = f.change_dtype(INT8);
            report("change_dtype fp32 -> u8", b.dtype == INT8 && same(b, img));
        }
        {   // test_normalize.cpp: given statistics, and statistics of the image (exact sums)
            Tensor dst;
            va_cv::normalize(src_u8, dst, mean, stddev);
            std::vector<float> want(img.size());
            orc_normalize_u8(img.data(), (size_t)w * h, c, 1, mean_v, std_v, want.data());
            report("normalize hwc u8, given mean/stddev", dst.dtype == FP32 && same(dst, want));
            Tensor dstf;
            va_cv::normalize(src_f32, dstf, mean, stddev);
            report("normalize hwc fp32, given mean/stddev", same(dstf, want));
            Tensor dauto;
            va_cv::normalize(src_u8, dauto);
            uint64_t sums[6] = {0};
            float m[3], s[3];
            orc_sums_u8(img.data(), (size_t)w * h, c, 1, sums);
            orc_finalize_mean_stddev(sums, c, (uint64_t)w * h, m, s);
            orc_normalize_u8(img.data(), (size_t)w * h, c, 1, m, s, want.data());
            report("normalize hwc u8, automatic statistics", same(dauto, want));
        }
        {   // test_warp_affine.cpp: matrix variant (:31-32) and scale/rotation variant (:198-205)
            const int W = 1280, H = 720;
            std::vector<uint8_t> big = random_u8((size_t)W * H * 3);
            Tensor src(W, H, 3, big.data(), INT8, NHWC), dst;
            float mv[6] = {0.849158f, 0.012257f, -474.827f, -0.01225f, 0.849158f, -379.18f};
            Tensor M(3, 2, 1, NCHW, FP32);
            std::memcpy(M.data, mv, sizeof(mv));
            va_cv::warp_affine(src, dst, M, va_cv::VSize(240, 240));
            float inv[6];
            std::memcpy(inv, mv, sizeof(mv));
            orc_invert_affine(inv);
            std::vector<uint8_t> want(240 * 240 * 3, 0);
            orc_warp_affine_u8(big.data(), W, H, 3, 1, want.data(), 240, 240, inv, 0);
            report("warp_affine(M) hwc u8 -> 240x240", same(dst, want));
            report("warp_affine(M) leaves the inverse in M", std::memcmp(M.data, inv, sizeof(inv)) == 0);

            std::vector<uint8_t> grey = random_u8((size_t)W * H);
            Tensor gsrc(W, H, 1, grey.data(), INT8, NHWC), gdst;
            va_cv::VScalar aux;
            aux.v0 = 738.518372f; aux.v1 = 537.672852f; aux.v2 = 204.766998f; aux.v3 = 73.329681f;
            va_cv::warp_affine(gsrc, gdst, 1.073914f, -3.314525f, va_cv::VSize(140, 210), aux);
            const double a4[4] = {aux.v0, aux.v1, aux.v2, aux.v3};
            float rm[6];
            orc_rotation_matrix(1.073914f, -3.314525f, a4, rm);
            orc_invert_affine(rm);
            std::vector<uint8_t> wantg(140 * 210, 0);
            orc_warp_affine_u8(grey.data(), W, H, 1, 1, wantg.data(), 140, 210, rm, 0);
            report("warp_affine(scale, rot, aux) grey -> 140x210", same(gdst, wantg));

            Tensor M2(3, 2, 1, NCHW, FP32), dn;
            std::memcpy(M2.data, mv, sizeof(mv));
            va_cv::warp_affine_normalize(src, dn, M2, va_cv::VSize(112, 112), va_cv::INTER_LINEAR, va_cv::BORDER_CONSTANT,
                                         va_cv::VScalar(), mean, stddev);
            std::vector<float> wantn(112 * 112 * 3);
            orc_warp_affine_normalize(big.data(), W, H, 3, inv, 112, 112, mean_v, std_v, wantn.data());
            report("warp_affine_normalize hwc u8 -> 112x112 fp32", dn.dtype == FP32 && same(dn, wantn));
        }
        {   // resize_normalize (fused)
            Tensor dn;
            va_cv::resize_normalize(src_u8, dn, va_cv::VSize(224, 224), 0, 0, va_cv::INTER_LINEAR, mean, stddev);
            std::vector<uint8_t> sm(224 * 224 * 3);
            std::vector<float> want(sm.size());
            orc_resize_linear_u8(img.data(), w, h, c, 1, sm.data(), 224, 224, 0);
            orc_normalize_u8(sm.data(), 224 * 224, 3, 1, mean_v, std_v, want.data());
            report("resize_normalize hwc u8 -> 224x224 fp32", same(dn, want));
        }
        {   // test_crop.cpp rects (:16-20) + an off-origin, fractional one
            const float rects[][4] = {{0, 0, 5, 5}, {0, 0, 320, 180}, {7.9f, 3.2f, 200.5f, 99.7f}};
            bool ok = true;
            for (auto& r : rects) {
                vision::VRect rect(r[0], r[1], r[2], r[3]);
                Tensor d8, df;
                va_cv::crop(src_u8, d8, rect);
                va_cv::crop(src_f32, df, rect);
                const int l = (int)r[0], t = (int)r[1], cw = (int)(r[2] - r[0]), ch = (int)(r[3] - r[1]);
                std::vector<uint8_t> w8((size_t)cw * ch * 3);
                std::vector<float> wf((size_t)cw * ch * 3);
                orc_crop(img.data(), w, h, c, 1, 1, l, t, cw, ch, w8.data());
                orc_crop(imgf.data(), w, h, c, 4, 1, l, t, cw, ch, wf.data());
                ok = ok && d8.w == cw && d8.h == ch && same(d8, w8) && same(df, wf);
            }
            report("crop hwc u8 / fp32", ok);
        }
        {   // the fused entry points with the header's DEFAULT arguments (cv.h:154-201: mean / stddev empty): statistics of the
            // resized / warped image (resize_normalize.cpp:58), and the layouts / dtypes / modes the composition supports
            Tensor dn;
            va_cv::resize_normalize(src_u8, dn, va_cv::VSize(224, 200));
            std::vector<uint8_t> sm(224 * 200 * 3);
            std::vector<float> want(sm.size());
            orc_resize_linear_u8(img.data(), w, h, c, 1, sm.data(), 224, 200, 0);
            uint64_t sums[6] = {0};
            float m[3], sd[3];
            orc_sums_u8(sm.data(), 224 * 200, 3, 1, sums);
            orc_finalize_mean_stddev(sums, 3, 224 * 200, m, sd);
            orc_normalize_u8(sm.data(), 224 * 200, 3, 1, m, sd, want.data());
            report("resize_normalize, default arguments (own statistics)", dn.dtype == FP32 && dn.layout == NHWC && same(dn, want));
            Tensor dcub;
            va_cv::resize_normalize(src_f32, dcub, va_cv::VSize(300, 170), 0, 0, va_cv::INTER_CUBIC, mean, stddev);
            std::vector<float> cub(300 * 170 * 3), wantc(cub.size());
            orc_resize_cubic_f32(imgf.data(), w, h, c, 1, cub.data(), 300, 170);
            orc_normalize_f32(cub.data(), 300 * 170, 3, 1, mean_v, std_v, wantc.data());
            report("resize_normalize fp32 INTER_CUBIC (composition)", same(dcub, wantc));
            Tensor chw = src_u8.change_layout(NCHW), dchw;
            va_cv::resize_normalize(chw, dchw, va_cv::VSize(320, 180), 0, 0, va_cv::INTER_LINEAR, mean, stddev);
            std::vector<uint8_t> in_chw(img.size()), r_chw(320 * 180 * 3);
            std::vector<float> want_chw(r_chw.size());
            orc_hwc_to_chw(img.data(), w, h, c, 1, in_chw.data());
            orc_resize_linear_u8(in_chw.data(), w, h, c, 0, r_chw.data(), 320, 180, 0);
            orc_normalize_u8(r_chw.data(), 320 * 180, 3, 0, mean_v, std_v, want_chw.data());
            report("resize_normalize chw u8 (composition)", dchw.layout == NCHW && same(dchw, want_chw));

            float mv[6] = {0.5f, 0.05f, 10.f, -0.05f, 0.5f, 20.f}, inv[6];
            std::memcpy(inv, mv, sizeof(mv));
            orc_invert_affine(inv);
            Tensor M(3, 2, 1, NCHW, FP32), dw;
            std::memcpy(M.data, mv, sizeof(mv));
            va_cv::warp_affine_normalize(src_u8, dw, M, va_cv::VSize(96, 80));
            std::vector<uint8_t> wu(96 * 80 * 3, 0);
            std::vector<float> wantw(wu.size());
            orc_warp_affine_u8(img.data(), w, h, 3, 1, wu.data(), 96, 80, inv, 0);
            uint64_t s2[6] = {0};
            orc_sums_u8(wu.data(), 96 * 80, 3, 1, s2);
            orc_finalize_mean_stddev(s2, 3, 96 * 80, m, sd);
            orc_normalize_u8(wu.data(), 96 * 80, 3, 1, m, sd, wantw.data());
            report("warp_affine_normalize, default arguments", same(dw, wantw) && std::memcmp(M.data, inv, sizeof(inv)) == 0);
        }
        {   // crop uploads only the ROI: off-origin rectangles at every column phase, both layouts, both dtypes
            bool ok = true;
            Tensor chw8 = src_u8.change_layout(NCHW), chwf = src_f32.change_layout(NCHW);
            std::vector<uint8_t> in8(img.size());
            std::vector<float> inf(img.size());
            orc_hwc_to_chw(img.data(), w, h, c, 1, in8.data());
            orc_hwc_to_chw(imgf.data(), w, h, c, 4, inf.data());
            const int rects[][4] = {{17, 5, 101, 77}, {16, 0, 624, 360}, {1, 359, 3, 1}, {333, 100, 307, 259}, {0, 0, 640, 360}, {48, 31, 64, 64}};
            for (auto& r : rects) {
                vision::VRect rect((float)r[0], (float)r[1], (float)(r[0] + r[2]), (float)(r[1] + r[3]));
                Tensor a, b, cc, d;
                va_cv::crop(src_u8, a, rect); va_cv::crop(src_f32, b, rect); va_cv::crop(chw8, cc, rect); va_cv::crop(chwf, d, rect);
                std::vector<uint8_t> w8((size_t)r[2] * r[3] * 3), w8c(w8.size());
                std::vector<float> wf(w8.size()), wfc(w8.size());
                orc_crop(img.data(), w, h, c, 1, 1, r[0], r[1], r[2], r[3], w8.data());
                orc_crop(imgf.data(), w, h, c, 4, 1, r[0], r[1], r[2], r[3], wf.data());
                orc_crop(in8.data(), w, h, c, 1, 0, r[0], r[1], r[2], r[3], w8c.data());
                orc_crop(inf.data(), w, h, c, 4, 0, r[0], r[1], r[2], r[3], wfc.data());
                ok = ok && same(a, w8) && same(b, wf) && cc.layout == NCHW && same(cc, w8c) && same(d, wfc);
            }
            report("crop off-origin, hwc / chw, u8 / fp32 (ROI-only upload)", ok);
        }
        {   // one thread, alternating shapes (and GPUs when the box has two): per-device contexts and LRU plan caches
            int n_dev = 0;
            vacv_cuda_device_count(&n_dev);
            const int sizes[][2] = {{320, 180}, {200, 120}, {416, 234}};
            std::vector<std::vector<uint8_t>> wants;
            for (auto& sz : sizes) {
                std::vector<uint8_t> wv((size_t)sz[0] * sz[1] * 3);
                orc_resize_linear_u8(img.data(), w, h, c, 1, wv.data(), sz[0], sz[1], 0);
                wants.push_back(wv);
            }
            std::vector<uint8_t> nv((size_t)w * h * 3 / 2);
            orc_bgr_to_nv21(img.data(), w, h, nv.data());
            bool ok = true;
            for (int it = 0; it < 12; ++it) {
                if (n_dev >= 2) ok = ok && vacv_cuda_set_device(it & 1) == 0;
                const int k = it % 3;
                Tensor d;
                va_cv::resize(src_u8, d, va_cv::VSize(sizes[k][0], sizes[k][1]));
                ok = ok && same(d, wants[k]);
                // the host-buffer form of the fused pipeline on the same thread / device
                const int wo = 160 + 32 * k, ho = 96 + 16 * k;
                std::vector<float> got((size_t)3 * wo * ho), wantp(got.size());
                ok = ok && vacv_cuda_nv_resize_normalize_chw_host(nv.data(), got.data(), 1, w, h, 1, wo, ho, mean_v, std_v, 1) == 0;
                orc_nv_resize_normalize_chw(nv.data(), w, h, 1, wo, ho, mean_v, std_v, wantp.data());
                ok = ok && std::memcmp(got.data(), wantp.data(), got.size() * sizeof(float)) == 0;
            }
            if (n_dev >= 2) vacv_cuda_set_device(0);
            report(n_dev >= 2 ? "one thread alternating 2 GPUs x 3 shapes" : "one thread alternating 3 shapes (1 GPU)", ok);
        }
        {   // ownership: ref counting and the error path
            Tensor a(16, 16, 3, INT8, NHWC);
            bool ok = a.get_ref_count() == 1;
            { Tensor b = a; ok = ok && a.get_ref_count() == 2 && b.data == a.data; }
            ok = ok && a.get_ref_count() == 1;
            Tensor cl = a.clone();
            ok = ok && cl.data != a.data && cl.get_ref_count() == 1;
            bool threw = false;
            try { Tensor d; va_cv::resize(src_u8, d, va_cv::VSize(10, 10), 0, 0, va_cv::INTER_AREA); } catch (const std::runtime_error&) { threw = true; }
            report("Tensor ref counting; unsupported mode raises", ok && threw);
        }
    } catch (const std::exception& e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 100;
    }
    std::printf("%d / %d cases passed\n", g_run - g_failed, g_run);
    return g_failed;
}
