// TEST INFRASTRUCTURE ONLY.  main() for the reference's own test-suite (src/test/src/impl/*.cpp + profile/cv_profile.cpp,
// compiled unmodified where they lie) with ALL 36 cases registered -- the reference's test_main.cpp:19-63 lists the same
// cases but has 32 of them commented out.  The binary is linked twice (oracle/Makefile):
//   oracle/_ref/va_cv_ut_b200 : against arm-neon-opencv_b200/libvacv.so  (the drop-in under test)
//   oracle/_ref/va_cv_ut_ref  : against oracle/_ref/liboracle_ref.so     (the reference itself, for comparison)
// Run from a directory containing res/ (the reference's JPEG fixtures).  After the reference's own report it prints one
// machine-readable line per case:  "CASE <name> <mean cosine>".
#include <cstdio>
#include <vector>

#include "impl/test_change_dtype.h"
#include "impl/test_change_layout.h"
#include "impl/test_crop.h"
#include "impl/test_cvt_color.h"
#include "impl/test_normalize.h"
#include "impl/test_resize.h"
#include "impl/test_warp_affine.h"
#include "profile/cv_profile.h"

using namespace vacv;

#define CASE(cls, fn) {cls::fn, #fn}

int main() {
    std::vector<CvProfile::SpeedResult> speed;
    std::vector<CvProfile::OutputResult> output;
    CvProfile::TestFuncList cases{
        CASE(TestCrop, test_crop_hwc_5x5), CASE(TestCrop, test_crop_hwc_5x5_FP32), CASE(TestCrop, test_crop_hwc_320x180),
        CASE(TestCrop, test_crop_hwc_640x360), CASE(TestCrop, test_crop_hwc_1280x720), CASE(TestCrop, test_crop_hwc_1920x1080),
        CASE(TestCrop, test_crop_chw_320x180), CASE(TestCrop, test_crop_chw_320x180_FP32), CASE(TestCrop, test_crop_chw_640x360),
        CASE(TestCrop, test_crop_chw_5x5), CASE(TestCrop, test_crop_chw_5x5_FP32),
        CASE(TestResize, test_resize_bilinear_hwc_u8_320x180), CASE(TestResize, test_resize_bilinear_chw_u8_320x180),
        CASE(TestResize, test_resize_bilinear_hwc_fp32_320x180), CASE(TestResize, test_resize_bilinear_chw_fp32_320x180),
        CASE(TestResize, test_resize_cubic_hwc_fp32_320x180), CASE(TestResize, test_resize_cubic_chw_fp32_320x180),
        CASE(TestChangeDtype, test_change_dtype_u8_to_fp32_176x144), CASE(TestChangeDtype, test_change_dtype_fp32_to_u8_176x144),
        CASE(TestChangeLayout, test_change_layout_hwc_to_chw_u8_176x144), CASE(TestChangeLayout, test_change_layout_hwc_to_chw_fp32_176x144),
        CASE(TestNormalize, test_normalize_hwc_176x144), CASE(TestNormalize, test_normalize_chw_176x144), CASE(TestNormalize, test_normalize_hwc_284x214),
        CASE(TestWarpAffine, test_warp_affine_hwc_u8), CASE(TestWarpAffine, test_warp_affine_hwc_fp32),
        CASE(TestWarpAffine, test_get_rotation_matrix_hwc_u8), CASE(TestWarpAffine, test_get_rotation_matrix_hwc_fp32),
        CASE(TestWarpAffine, test_warp_affine_chw_u8), CASE(TestWarpAffine, test_warp_affine_chw_fp32), CASE(TestWarpAffine, test_get_rotation_matrix_chw_u8),
        CASE(TestCvtColor, test_nv21_to_bgr_176x144), CASE(TestCvtColor, test_nv21_to_bgr_640x360), CASE(TestCvtColor, test_nv21_to_bgr_1280x720),
        CASE(TestCvtColor, test_nv21_to_bgr_1920x1080), CASE(TestCvtColor, test_nv21_to_bgr_2560x1440),
    };
    CvProfile::profile(cases, nullptr, nullptr, speed, output);
    for (const auto& o : output) std::printf("CASE %s %.9f\n", o.second.c_str(), o.first[0]);
    return 0;
}
