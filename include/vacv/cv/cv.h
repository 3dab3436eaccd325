// va_cv -- the public operator API of the vacv drop-in library (libvacv.so).
//
// Same namespace, enums, value types, function names, parameter lists and defaults as the reference's
// src/cv/cv.h:11-239, so existing callers switch by re-linking.  Every operator here copies its host tensors to
// the GPU, runs the sm_100a kernels behind include/vacv_cuda.h on the calling thread's stream, and copies the
// result back before returning (the reference API is synchronous).  Batched, device-resident use goes through
// the C-ABI directly.
//
// Behaviour notes (see DESIGN.md "compat decisions"):
//   * fx / fy of resize and borderValue of warp_affine are accepted and ignored, as in the reference's native
//     paths (resize.cpp:42-100, warp_affine.cpp:111-169).
//   * cvt_color decodes COLOR_YUV2BGR_NV12 with V-first chroma exactly like the reference (cvt_color.cpp:139-149).
//   * warp_affine(M) overwrites M with its inverse like the reference (warp_affine.cpp:121-133); destination pixels
//     that map outside the source are 0.
//   * normalize with empty mean/stddev computes exact statistics (the reference accumulates sequentially in fp32).
//   * combinations the reference has no native code for throw std::runtime_error instead of recursing forever.
#ifndef VISION_CV_H
#define VISION_CV_H

#include <vector>

#include "../common/tensor.h"
#include "../common/vision_structs.h"

namespace va_cv {

struct VSize {
    int w;
    int h;
    VSize() : w(0), h(0) {}
    VSize(int _w, int _h) : w(_w), h(_h) {}
};

struct VScalar {
    double v0, v1, v2, v3;
    VScalar() : v0(0), v1(0), v2(0), v3(0) {}
};

enum VInterMode { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4,
                  INTER_MAX = 7, WARP_INVERSE_MAP = 16 };

enum VBorderMode { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
                   BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_TRANSPARENT = 5,
                   BORDER_ISOLATED = 16 };

enum VMatchMode { TM_SQDIFF = 0, TM_SQDIFF_NORMED = 1, TM_CCORR = 2, TM_CCORR_NORMED = 3, TM_CCOEFF = 4,
                  TM_CCOEFF_NORMED = 5 };

enum InputImageFormat { COLOR_GRAY2RGB = 8, COLOR_GRAY2BGR = COLOR_GRAY2RGB, COLOR_YUV2RGB_NV12 = 90,
                        COLOR_YUV2BGR_NV12 = 91, COLOR_YUV2RGB_NV21 = 92, COLOR_YUV2BGR_NV21 = 93,
                        COLOR_YUV2RGBA_NV12 = 94, COLOR_YUV2BGRA_NV12 = 95, COLOR_YUV2RGBA_NV21 = 96,
                        COLOR_YUV2BGRA_NV21 = 97, COLOR_YUV2BGR_YV12 = 99 };

// INTER_LINEAR (u8 / fp32) and INTER_CUBIC (fp32; u8 = OpenCV-2.4 rule, HWC); dst takes src's dtype and layout.
void resize(const vision::Tensor& src, vision::Tensor& dst, VSize dsize, double fx = 0, double fy = 0,
            int interpolation = INTER_LINEAR);

// NV21 / NV12 (w x h*3/2, one channel) -> BGR HWC u8.
void cvt_color(const vision::Tensor& src, vision::Tensor& dst, int code);

// (x - mean[k]) / (stddev[k] + 1e-6) -> fp32, layout kept; both statistics empty => computed per channel.
void normalize(const vision::Tensor& src, vision::Tensor& dst, const vision::Tensor& mean = vision::Tensor(),
               const vision::Tensor& stddev = vision::Tensor());

// Bilinear warp with the forward 2x3 matrix M (fp32 tensor, 6 values).
void warp_affine(const vision::Tensor& src, vision::Tensor& dst, const vision::Tensor& M, VSize dsize,
                 int flags = INTER_LINEAR, int borderMode = BORDER_CONSTANT, const VScalar& borderValue = VScalar());

// Same, M built from scale / rotation (degrees) about the origin plus the aux translation (v0..v3).
void warp_affine(const vision::Tensor& src, vision::Tensor& dst, float scale, float rot, VSize dsize,
                 const VScalar& aux_param = VScalar(), int flags = INTER_LINEAR, int borderMode = BORDER_CONSTANT,
                 const VScalar& borderValue = VScalar());

// Fused resize -> fp32 -> normalize (u8 HWC input), one kernel.
void resize_normalize(const vision::Tensor& src, vision::Tensor& dst, VSize dsize, double fx = 0, double fy = 0,
                      int interpolation = INTER_LINEAR, const vision::Tensor& mean = vision::Tensor(),
                      const vision::Tensor& stddev = vision::Tensor());

// Fused warp_affine -> fp32 -> normalize (u8 HWC input), one kernel.
void warp_affine_normalize(const vision::Tensor& src, vision::Tensor& dst, const vision::Tensor& M, VSize dsize,
                           int flags = INTER_LINEAR, int borderMode = BORDER_CONSTANT,
                           const VScalar& borderValue = VScalar(), const vision::Tensor& mean = vision::Tensor(),
                           const vision::Tensor& stddev = vision::Tensor());

void warp_affine_normalize(const vision::Tensor& src, vision::Tensor& dst, float scale, float rot, VSize dsize,
                           const VScalar& aux_param = VScalar(), int flags = INTER_LINEAR,
                           int borderMode = BORDER_CONSTANT, const VScalar& borderValue = VScalar(),
                           const vision::Tensor& mean = vision::Tensor(), const vision::Tensor& stddev = vision::Tensor());

// ROI copy; rect edges are truncated to int like the reference (crop.cpp:128-131).
void crop(const vision::Tensor& src, vision::Tensor& dst, const vision::VRect& rect);

// Declared for link compatibility; like the reference without USE_OPENCV these have no native implementation
// (match_template.cpp:48-60, imencode.cpp:11-15) and do nothing.
void match_template(const vision::Tensor& src, const vision::Tensor& target, vision::Tensor& result, int method);
void minMaxIdx(const vision::Tensor& src, double* minVal, double* maxVal, int* minIdx = nullptr, int* maxIdx = nullptr,
               const vision::Tensor& mask = vision::Tensor());
void imencode(const vision::Tensor& src, std::vector<unsigned char>& buf, const char* format);

}  // namespace va_cv

#endif  // VISION_CV_H
