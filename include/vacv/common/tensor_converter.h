// vision::TensorConverter -- cv::Mat <-> vision::Tensor, the helper the reference's callers and its own test-suite use
// (reference: src/common/tensor_converter.h:8-17, src/common/tensor_converter.cpp:15-83).
//
// Header-only on purpose: libvacv.so must not depend on any particular OpenCV (cv::Mat's layout differs between the
// reference's bundled 2.4.13 and current releases), so the two specialisations instantiate in the CALLER's translation
// unit against whichever <opencv2/core/core.hpp> that caller builds with.  Include an OpenCV core header first.
// Code that was compiled against the reference's declaration-only header (and therefore needs the two symbols at link
// time) links libvacv_cvmat.so instead: csrc/host/tensor_converter.cpp emits exactly these definitions, non-inline,
// for the OpenCV whose headers it is built with (make -C arm-neon-opencv_b200/csrc cvmat OPENCV_INC=...).
//
// Type map, as in the reference: Tensor -> Mat: FP32 -> CV_32F, FP16 -> CV_16U (16-bit payload, moved as integers),
// INT8 -> CV_8U, FP64 -> CV_64F; Mat -> Tensor: 8U/8S -> INT8, 16U/16S -> FP16, 32S/32F -> FP32, 64F -> FP64, layout NHWC.
// copy = false shares the pixels (Mat header over tensor.data / borrowing Tensor over mat.data), copy = true duplicates.
#ifndef VISION_TENSOR_CONVERTER_H
#define VISION_TENSOR_CONVERTER_H

#include <cstring>
#include <stdexcept>

#include "tensor.h"

#ifndef VACV_TENSOR_CONVERTER_INLINE
#define VACV_TENSOR_CONVERTER_INLINE inline
#endif

namespace vision {

class TensorConverter {
public:
    template <typename T>
    static T convert_to(const Tensor& tensor, bool copy = false);

    template <typename T>
    static Tensor convert_from(const T& mat, bool copy = false);
};

#if defined(CV_VERSION) || defined(CV_MAJOR_VERSION) || defined(__OPENCV_CORE_HPP__) || defined(OPENCV_CORE_HPP)

template <>
VACV_TENSOR_CONVERTER_INLINE cv::Mat TensorConverter::convert_to<cv::Mat>(const Tensor& tensor, bool copy) {
    if (tensor.empty()) return cv::Mat();
    int depth;
    switch (tensor.dtype) {
        case FP32: depth = CV_32F; break;
        case FP16: depth = CV_16U; break;
        case INT8: depth = CV_8U; break;
        case FP64: depth = CV_64F; break;
        default: throw std::runtime_error("TensorConverter: tensor dtype has no cv::Mat equivalent");
    }
    const int type = CV_MAKETYPE(depth, tensor.c);
    if (!copy) return cv::Mat(tensor.h, tensor.w, type, tensor.data);   // header over the tensor's pixels
    cv::Mat owned(tensor.h, tensor.w, type);
    std::memcpy(owned.data, tensor.data, tensor.len());
    return owned;
}

template <>
VACV_TENSOR_CONVERTER_INLINE Tensor TensorConverter::convert_from<cv::Mat>(const cv::Mat& mat, bool copy) {
    if (mat.empty()) return Tensor();
    DType dtype;
    switch (mat.depth()) {
        case CV_8U: case CV_8S: dtype = INT8; break;
        case CV_16U: case CV_16S: dtype = FP16; break;
        case CV_32S: case CV_32F: dtype = FP32; break;
        case CV_64F: dtype = FP64; break;
        default: throw std::runtime_error("TensorConverter: cv::Mat depth has no tensor dtype");
    }
    if (!copy) return Tensor(mat.cols, mat.rows, mat.channels(), mat.data, dtype, NHWC);   // borrows mat's pixels
    Tensor owned(mat.cols, mat.rows, mat.channels(), dtype, NHWC);
    std::memcpy(owned.data, mat.data, owned.len());
    return owned;
}

#endif  // an OpenCV core header was included first

}  // namespace vision

#endif  // VISION_TENSOR_CONVERTER_H
