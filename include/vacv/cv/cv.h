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

typedef vision::Tensor VTensor;   // shorthand for the declarations below (same type, same mangled names)

// ---- colour, geometry -------------------------------------------------------------------------------------------
// NV21 / NV12 / YV12 frame (w x h*3/2, one channel, u8) -> BGR HWC u8.
void cvt_color(const VTensor& yuv, VTensor& bgr, int code);

// ROI copy; rect edges are truncated to int like the reference (crop.cpp:128-131).
void crop(const VTensor& image, VTensor& roi, const vision::VRect& rect);

// INTER_LINEAR (u8 / fp32) and INTER_CUBIC (fp32; u8 = OpenCV-2.4 rule, HWC).  The result takes the input's dtype and layout;
// scale_x / scale_y are accepted and ignored like in the reference's native paths.
void resize(const VTensor& image, VTensor& resized, VSize out_size, double scale_x = 0, double scale_y = 0,
            int interpolation = INTER_LINEAR);

// Bilinear warp with the forward 2x3 matrix (fp32 tensor, 6 values); the matrix is overwritten with its inverse.
void warp_affine(const VTensor& image, VTensor& warped, const VTensor& matrix, VSize out_size, int flags = INTER_LINEAR,
                 int border_mode = BORDER_CONSTANT, const VScalar& border_value = VScalar());
// Same, the matrix built from scale / rotation (degrees) about the origin plus the aux translation (v0..v3).
void warp_affine(const VTensor& image, VTensor& warped, float scale, float rotation_deg, VSize out_size,
                 const VScalar& aux = VScalar(), int flags = INTER_LINEAR, int border_mode = BORDER_CONSTANT,
                 const VScalar& border_value = VScalar());

// ---- statistics, normalisation -------------------------------------------------------------------------------------
// (x - mean[k]) / (stddev[k] + 1e-6) -> fp32, layout kept; both statistics empty => computed per channel from the image.
void normalize(const VTensor& image, VTensor& normalized, const VTensor& mean = VTensor(), const VTensor& stddev = VTensor());

// ---- fused variants (u8 HWC input, one kernel each) -------------------------------------------------------------
void resize_normalize(const VTensor& image, VTensor& out, VSize out_size, double scale_x = 0, double scale_y = 0,
                      int interpolation = INTER_LINEAR, const VTensor& mean = VTensor(), const VTensor& stddev = VTensor());
void warp_affine_normalize(const VTensor& image, VTensor& out, const VTensor& matrix, VSize out_size, int flags = INTER_LINEAR,
                           int border_mode = BORDER_CONSTANT, const VScalar& border_value = VScalar(),
                           const VTensor& mean = VTensor(), const VTensor& stddev = VTensor());
void warp_affine_normalize(const VTensor& image, VTensor& out, float scale, float rotation_deg, VSize out_size,
                           const VScalar& aux = VScalar(), int flags = INTER_LINEAR, int border_mode = BORDER_CONSTANT,
                           const VScalar& border_value = VScalar(), const VTensor& mean = VTensor(), const VTensor& stddev = VTensor());

// ---- declared for link compatibility ---------------------------------------------------------------------------
// Like the reference without USE_OPENCV these have no native implementation (match_template.cpp:48-60, imencode.cpp:11-15)
// and do nothing.
void imencode(const VTensor& image, std::vector<unsigned char>& bytes, const char* format);
void match_template(const VTensor& image, const VTensor& pattern, VTensor& scores, int method);
void minMaxIdx(const VTensor& values, double* min_value, double* max_value, int* min_index = nullptr, int* max_index = nullptr,
               const VTensor& mask = VTensor());

}  // namespace va_cv

#endif  // VISION_CV_H
