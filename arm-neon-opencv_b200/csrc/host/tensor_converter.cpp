// libvacv_cvmat.so: out-of-line definitions of vision::TensorConverter::convert_to / convert_from <cv::Mat> for callers
// that were compiled against the reference's declaration-only header (src/common/tensor_converter.h:8-17) and need the
// symbols at link time -- e.g. the reference's own test-suite (src/test/CMakeLists.txt:19-23 links only `vacv` + OpenCV).
// The definitions are the ones of include/vacv/common/tensor_converter.h, emitted non-inline for the OpenCV whose
// headers this file is compiled against (OPENCV_INC); libvacv.so itself stays OpenCV-free.
#include <opencv2/core/core.hpp>

#define VACV_TENSOR_CONVERTER_INLINE __attribute__((visibility("default")))
#include "common/tensor_converter.h"
