// va_cv::* for libvacv.so: the reference's public operator API (src/cv/cv.cpp:16-89) bound to the CUDA C-ABI.
//
// The reference routes each call through a class-static dispatcher that picks a backend at compile time
// (e.g. Resize::resize -> resize_naive, src/cv/resize.cpp:19-27) and then does shape bookkeeping +
// dst.create(...) before calling a raw-pointer kernel.  This file is that dispatcher layer with the CUDA backend:
// same bookkeeping, then  host tensor -> device scratch -> vacv_cuda_* -> host tensor, synchronously.
#include "cv/cv.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include "device_context.h"
#include "vacv_cuda.h"

namespace va_cv {

using vision::Tensor;
using vacv_host::DeviceContext;

namespace {

[[noreturn]] void unsupported(const std::string& what) {
    throw std::runtime_error("vacv: " + what + " has no native implementation (neither in the reference nor here)");
}

// mean / stddev tensors as the reference expects them: c fp32 values (normalize.cpp:109-112)
void require_stats(const Tensor& mean, const Tensor& stddev, int c) {
    if (mean.empty() || stddev.empty() || static_cast<int>(mean.size()) != c || static_cast<int>(stddev.size()) != c ||
        mean.dtype != vision::FP32 || stddev.dtype != vision::FP32)
        throw std::runtime_error("vacv: mean / stddev must be FP32 tensors with one value per channel");
}

// uploads mean|stddev into one scratch slot; returns device pointers
void upload_stats(DeviceContext& ctx, int slot, const Tensor& mean, const Tensor& stddev, int c, float*& d_mean, float*& d_std) {
    float tmp[2 * 16];
    if (c > 16) throw std::runtime_error("vacv: more than 16 channels");
    std::memcpy(tmp, mean.data, sizeof(float) * c);
    std::memcpy(tmp + c, stddev.data, sizeof(float) * c);
    d_mean = static_cast<float*>(ctx.upload(slot, tmp, sizeof(float) * 2 * c));
    ctx.sync();   // tmp is a stack buffer
    d_std = d_mean + c;
}

void build_matrix(float scale, float rot, const VScalar& aux, float m[6]) {
    const double a[4] = {aux.v0, aux.v1, aux.v2, aux.v3};
    vacv_rotation_matrix(scale, rot, a, m);   // warp_affine.cpp:76-109
}

void check_warp_mode(const Tensor& src, int flags, int borderMode) {
    if ((src.dtype != vision::INT8 && src.dtype != vision::FP32) || borderMode != BORDER_CONSTANT || flags != INTER_LINEAR)
        unsupported("warp_affine with this dtype / border mode / interpolation (warp_affine.cpp:114-119)");
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- resize
void resize(const Tensor& src, Tensor& dst, VSize dsize, double /*fx*/, double /*fy*/, int interpolation) {
    if (src.empty()) throw std::runtime_error("vacv: resize of an empty tensor");
    if (interpolation != INTER_LINEAR && interpolation != INTER_CUBIC) unsupported("resize with this interpolation");
    if (src.dtype != vision::INT8 && src.dtype != vision::FP32) unsupported("resize of this dtype");
    const Tensor in = src;   // keep the source alive if dst aliases it
    dst.create(dsize.w, dsize.h, in.c, in.dtype, in.layout);   // resize.cpp:56
    DeviceContext& ctx = DeviceContext::current();
    void* d_in = ctx.upload(0, in.data, in.len());
    void* d_out = ctx.scratch(1, dst.len());
    ctx.check(vacv_cuda_resize(d_in, d_out, 1, in.w, in.h, in.c, in.dtype, in.layout, dsize.w, dsize.h, interpolation,
                               VACV_FLAG_NONE, ctx.stream()));
    ctx.download(dst.data, d_out, dst.len());
}

// ---------------------------------------------------------------------------------------------------- cvt_color
void cvt_color(const Tensor& src, Tensor& dst, int code) {
    if (code != COLOR_YUV2BGR_NV21 && code != COLOR_YUV2BGR_NV12 && code != COLOR_YUV2BGR_YV12)
        unsupported("cvt_color with this code (cvt_color.cpp:139-141)");
    if (src.empty() || src.dtype != vision::INT8) throw std::runtime_error("vacv: cvt_color needs an INT8 NV21/NV12/YV12 tensor");
    const Tensor in = src;
    const int w = in.w, h = in.h / 3 * 2;   // cvt_color.cpp:151-152
    dst.create(w, h, 3, vision::NHWC, vision::INT8);
    DeviceContext& ctx = DeviceContext::current();
    void* d_in = ctx.upload(0, in.data, (size_t)w * h * 3 / 2);
    void* d_out = ctx.scratch(1, dst.len());
    if (code == COLOR_YUV2BGR_YV12) {   // declared by the reference (cv.h:73) but never implemented there: planar Y, V, U
        vacv_yuv_layout lay = {VACV_YUV_YV12, w, h, 0, 0, 0};
        ctx.check(vacv_cuda_cvt_yuv2bgr(static_cast<const uint8_t*>(d_in), &lay, static_cast<uint8_t*>(d_out), 1, ctx.stream()));
        ctx.download(dst.data, d_out, dst.len());
        return;
    }
    // the reference decodes both codes with V-first chroma (the swap at :146 tests the wrong enum), kept for parity
    ctx.check(vacv_cuda_cvt_nv2bgr(static_cast<const uint8_t*>(d_in), static_cast<uint8_t*>(d_out), 1, w, h, 1, ctx.stream()));
    ctx.download(dst.data, d_out, dst.len());
}

// ---------------------------------------------------------------------------------------------------- normalize
void normalize(const Tensor& src, Tensor& dst, const Tensor& mean, const Tensor& stddev) {
    if (src.empty()) throw std::runtime_error("vacv: normalize of an empty tensor");
    if (src.dtype != vision::INT8 && src.dtype != vision::FP32) unsupported("normalize of this dtype");
    const Tensor in = src;
    dst.create(in.w, in.h, in.c, vision::FP32, in.layout);   // normalize.cpp:90
    DeviceContext& ctx = DeviceContext::current();
    void* d_in = ctx.upload(0, in.data, in.len());
    float* d_out = static_cast<float*>(ctx.scratch(1, dst.len()));
    float *d_mean, *d_std;
    if (mean.empty() && stddev.empty()) {   // normalize.cpp:98: statistics of the image itself
        void* d_sums = ctx.scratch(2, 8 * 2 * in.c);   // u64 or fp64 counters
        d_mean = static_cast<float*>(ctx.scratch(3, sizeof(float) * 2 * in.c));
        d_std = d_mean + in.c;
        ctx.check(vacv_cuda_memset(d_sums, 0, 8 * 2 * in.c, ctx.stream()));
        const unsigned long long n = (unsigned long long)in.w * in.h;
        if (in.dtype == vision::INT8) {   // exact integer sums
            auto* su = static_cast<unsigned long long*>(d_sums);
            ctx.check(vacv_cuda_sums_u8(static_cast<const uint8_t*>(d_in), 1, in.w, in.h, in.c, in.layout, su, 0, ctx.stream()));
            ctx.check(vacv_cuda_finalize_mean_stddev(su, 1, in.c, n, d_mean, d_std, ctx.stream()));
        } else {                          // fp64 sums (exact for integer-valued pixels)
            auto* sd = static_cast<double*>(d_sums);
            ctx.check(vacv_cuda_sums_f32(static_cast<const float*>(d_in), 1, in.w, in.h, in.c, in.layout, sd, 0, ctx.stream()));
            ctx.check(vacv_cuda_finalize_mean_stddev_f64(sd, 1, in.c, n, d_mean, d_std, ctx.stream()));
        }
    } else {
        require_stats(mean, stddev, in.c);
        upload_stats(ctx, 3, mean, stddev, in.c, d_mean, d_std);
    }
    ctx.check(vacv_cuda_normalize(d_in, d_out, 1, in.w, in.h, in.c, in.dtype, in.layout, d_mean, d_std, 0, ctx.stream()));
    ctx.download(dst.data, d_out, dst.len());
}

// ---------------------------------------------------------------------------------------------------- warp_affine
void warp_affine(const Tensor& src, Tensor& dst, const Tensor& M, VSize dsize, int flags, int borderMode,
                 const VScalar& /*borderValue*/) {
    if (src.empty() || M.empty() || M.size() < 6 || M.dtype != vision::FP32)
        throw std::runtime_error("vacv: warp_affine needs a source and a 2x3 FP32 matrix");
    check_warp_mode(src, flags, borderMode);
    float* m = static_cast<float*>(M.data);
    vacv_invert_affine(m);   // in place in the caller's tensor, like warp_affine.cpp:121-133
    const Tensor in = src;
    dst.create(dsize.w, dsize.h, in.c, in.dtype, in.layout);
    DeviceContext& ctx = DeviceContext::current();
    void* d_in = ctx.upload(0, in.data, in.len());
    void* d_out = ctx.scratch(1, dst.len());
    float* d_m = static_cast<float*>(ctx.upload(2, m, sizeof(float) * 6));
    ctx.check(vacv_cuda_warp_affine(d_in, 1, in.w, in.h, in.c, in.dtype, in.layout, nullptr, d_m, 1, d_out, dsize.w, dsize.h,
                                    VACV_FLAG_NONE, ctx.stream()));
    ctx.download(dst.data, d_out, dst.len());
}

void warp_affine(const Tensor& src, Tensor& dst, float scale, float rot, VSize dsize, const VScalar& aux_param, int flags,
                 int borderMode, const VScalar& borderValue) {
    Tensor M(3, 2, 1, vision::FP32, vision::NCHW);
    build_matrix(scale, rot, aux_param, static_cast<float*>(M.data));
    warp_affine(src, dst, M, dsize, flags, borderMode, borderValue);
}

// ---------------------------------------------------------------------------------------------------- fused ops
// Semantics (the reference's only implemented body, resize_normalize.cpp:39-103): resize -> fp32 -> per-channel
// (x - m) / (s + 1e-6), with m / s taken from the RESIZED image when `mean` or `stddev` is empty (:58) -- the defaults of
// cv.h:154-159.  u8 HWC + INTER_LINEAR + explicit statistics is one fused kernel; everything else the composition supports
// (fp32 input, CHW, INTER_CUBIC, automatic statistics) runs as resize() + normalize(), which defines the fused result anyway.
void resize_normalize(const Tensor& src, Tensor& dst, VSize dsize, double fx, double fy, int interpolation,
                      const Tensor& mean, const Tensor& stddev) {
    if (src.empty()) throw std::runtime_error("vacv: resize_normalize of an empty tensor");
    const bool auto_stats = mean.empty() || stddev.empty();
    const bool fused = !auto_stats && src.dtype == vision::INT8 && src.layout == vision::NHWC && interpolation == INTER_LINEAR;
    if (!fused) {
        Tensor resized;
        resize(src, resized, dsize, fx, fy, interpolation);
        if (auto_stats) normalize(resized, dst, Tensor(), Tensor());
        else normalize(resized, dst, mean, stddev);
        return;
    }
    require_stats(mean, stddev, src.c);
    const Tensor in = src;
    dst.create(dsize.w, dsize.h, in.c, vision::FP32, vision::NHWC);
    DeviceContext& ctx = DeviceContext::current();
    void* d_in = ctx.upload(0, in.data, in.len());
    float* d_out = static_cast<float*>(ctx.scratch(1, dst.len()));
    float *d_mean, *d_std;
    upload_stats(ctx, 3, mean, stddev, in.c, d_mean, d_std);
    if (dsize.w == in.w && dsize.h == in.h)   // resize.cpp:58-61: plain copy, so only the normalisation remains
        ctx.check(vacv_cuda_normalize(d_in, d_out, 1, in.w, in.h, in.c, vision::INT8, vision::NHWC, d_mean, d_std, 0, ctx.stream()));
    else
        ctx.check(vacv_cuda_resize_normalize(static_cast<const uint8_t*>(d_in), d_out, 1, in.w, in.h, in.c, dsize.w, dsize.h,
                                             d_mean, d_std, VACV_NHWC, ctx.stream()));
    ctx.download(dst.data, d_out, dst.len());
}

// Same contract for the warp (warp_affine_normalize.cpp:47-120): statistics of the WARPED image when mean / stddev are empty;
// fp32 input or CHW layout run as warp_affine() + normalize().
void warp_affine_normalize(const Tensor& src, Tensor& dst, const Tensor& M, VSize dsize, int flags, int borderMode,
                           const VScalar& borderValue, const Tensor& mean, const Tensor& stddev) {
    if (src.empty() || M.empty() || M.size() < 6 || M.dtype != vision::FP32)
        throw std::runtime_error("vacv: warp_affine_normalize needs a source and a 2x3 FP32 matrix");
    check_warp_mode(src, flags, borderMode);
    const bool auto_stats = mean.empty() || stddev.empty();
    if (auto_stats || src.dtype != vision::INT8 || src.layout != vision::NHWC) {
        Tensor warped;
        warp_affine(src, warped, M, dsize, flags, borderMode, borderValue);   // inverts M in place, like the fused path below
        if (auto_stats) normalize(warped, dst, Tensor(), Tensor());
        else normalize(warped, dst, mean, stddev);
        return;
    }
    require_stats(mean, stddev, src.c);
    float* m = static_cast<float*>(M.data);
    vacv_invert_affine(m);
    const Tensor in = src;
    dst.create(dsize.w, dsize.h, in.c, vision::FP32, vision::NHWC);
    DeviceContext& ctx = DeviceContext::current();
    void* d_in = ctx.upload(0, in.data, in.len());
    float* d_out = static_cast<float*>(ctx.scratch(1, dst.len()));
    float* d_m = static_cast<float*>(ctx.upload(2, m, sizeof(float) * 6));
    float *d_mean, *d_std;
    upload_stats(ctx, 3, mean, stddev, in.c, d_mean, d_std);
    ctx.check(vacv_cuda_warp_affine_normalize(static_cast<const uint8_t*>(d_in), 1, in.w, in.h, in.c, nullptr, d_m, 1, d_out,
                                              dsize.w, dsize.h, d_mean, d_std, VACV_NHWC, ctx.stream()));
    ctx.download(dst.data, d_out, dst.len());
}

void warp_affine_normalize(const Tensor& src, Tensor& dst, float scale, float rot, VSize dsize, const VScalar& aux_param,
                           int flags, int borderMode, const VScalar& borderValue, const Tensor& mean, const Tensor& stddev) {
    Tensor M(3, 2, 1, vision::FP32, vision::NCHW);
    build_matrix(scale, rot, aux_param, static_cast<float*>(M.data));
    warp_affine_normalize(src, dst, M, dsize, flags, borderMode, borderValue, mean, stddev);
}

// ---------------------------------------------------------------------------------------------------- crop
void crop(const Tensor& src, Tensor& dst, const vision::VRect& rect) {
    if (src.empty()) throw std::runtime_error("vacv: crop of an empty tensor");
    if (src.dtype != vision::INT8 && src.dtype != vision::FP32) unsupported("crop of this dtype (crop.cpp:133-141)");
    const int left = static_cast<int>(rect.left), top = static_cast<int>(rect.top);           // crop.cpp:128-131
    const int cw = static_cast<int>(rect.width()), ch = static_cast<int>(rect.height());
    const Tensor in = src;
    if (cw <= 0 || ch <= 0 || left < 0 || top < 0 || left + cw > in.w || top + ch > in.h)
        throw std::runtime_error("vacv: crop rectangle outside the source (the reference would read out of bounds)");
    dst.create(cw, ch, in.c, in.dtype, in.layout);
    DeviceContext& ctx = DeviceContext::current();
    // Only the ROI crosses PCIe: the rows top .. top+ch, and of each row the columns around the rectangle (widened to whole
    // 16-pixel groups so the DMA engine moves aligned runs); the kernel then crops the few residual columns on the device.
    const size_t es = in.dtype == vision::INT8 ? 1 : 4;
    const int x0 = left & ~15, x1 = std::min(in.w, (left + cw + 15) & ~15), sub_w = x1 - x0;
    const bool hwc = in.layout == vision::NHWC;
    const size_t px = hwc ? es * in.c : es;                   // bytes per pixel step along a row
    const size_t row_bytes = (size_t)sub_w * px, src_pitch = (size_t)in.w * px;
    const int planes = hwc ? 1 : in.c;
    uint8_t* d_in = static_cast<uint8_t*>(ctx.scratch(0, row_bytes * ch * planes));
    for (int k = 0; k < planes; ++k) {
        const uint8_t* h0 = static_cast<const uint8_t*>(in.data) + (size_t)k * in.w * in.h * es + (size_t)top * src_pitch + (size_t)x0 * px;
        ctx.check(vacv_cuda_memcpy2d_h2d(d_in + (size_t)k * row_bytes * ch, row_bytes, h0, src_pitch, row_bytes, ch, ctx.stream()));
    }
    void* d_out = ctx.scratch(1, dst.len());
    ctx.check(vacv_cuda_crop(d_in, d_out, 1, sub_w, ch, in.c, in.dtype, in.layout, left - x0, 0, cw, ch, ctx.stream()));
    ctx.download(dst.data, d_out, dst.len());
}

// ---------------------------------------------------------------------------------------------------- not on the path
// Without USE_OPENCV the reference's bodies are empty (match_template.cpp:48-60, imencode.cpp:11-15).
void match_template(const Tensor&, const Tensor&, Tensor&, int) {}
void minMaxIdx(const Tensor&, double*, double*, int*, int*, const Tensor&) {}
void imencode(const Tensor&, std::vector<unsigned char>&, const char*) {}

}  // namespace va_cv
