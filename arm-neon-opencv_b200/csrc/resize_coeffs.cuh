// Coefficient generators of the resize family, evaluated on the device in the reference's exact operation order
// (built --fmad=false).  Shared by the direct-gather kernels (resize.cu) and the tiled kernels (resize_tiled.cu).
#pragma once
#include "vacv_common.cuh"

namespace vacv {

// ----------------------------------------------------------------------------------------------------
// a8 bicubic fp32 (resize_naive.cpp:130-185 coefficients with border folding; :230,:345 accumulation order)
// scale = (double)n_in / n_out (resize_naive.cpp:144)
__device__ __forceinline__ void cubic_naive_scaled(int d, int n_in, double scale, int& ofs, float (&a)[4]) {
    float fx = (float)(((double)d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    const float A = -0.75f;
    const float fx0 = fx + 1, fx1 = fx, fx2 = 1 - fx;
    a[0] = A * fx0 * fx0 * fx0 - 5 * A * fx0 * fx0 + 8 * A * fx0 - 4 * A;
    a[1] = (A + 2) * fx1 * fx1 * fx1 - (A + 3) * fx1 * fx1 + 1;
    a[2] = (A + 2) * fx2 * fx2 * fx2 - (A + 3) * fx2 * fx2 + 1;
    a[3] = 1.f - a[0] - a[1] - a[2];
    if (sx <= -1) { sx = 1; a[0] = 1.f - a[3]; a[1] = a[3]; a[2] = 0.f; a[3] = 0.f; }
    if (sx == 0) { sx = 1; a[0] = a[0] + a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = 0.f; }
    if (sx == n_in - 2) { sx = n_in - 3; a[3] = a[2] + a[3]; a[2] = a[1]; a[1] = a[0]; a[0] = 0.f; }
    if (sx >= n_in - 1) { sx = n_in - 3; a[3] = 1.f - a[0]; a[2] = a[0]; a[1] = 0.f; a[0] = 0.f; }
    ofs = sx;
}
__device__ __forceinline__ void cubic_naive(int d, int n_in, int n_out, int& ofs, float (&a)[4]) {
    cubic_naive_scaled(d, n_in, (double)n_in / (double)n_out, ofs, a);
}


// ----------------------------------------------------------------------------------------------------
// a9 bicubic u8 = OpenCV 2.4.13 cv::resize(CV_8UCn, INTER_CUBIC): int32 horizontal pass with
// 11-bit coefficients (round-half-even), fp32 vertical pass with round-half-even for the first (W*cn & ~7)
// elements of a row (the SSE2 body) and an integer (+2^21)>>22 vertical pass for the <=7 tail elements.
__device__ __forceinline__ void cubic_cv(float x, float (&k)[4]) {
    const float A = -0.75f;
    k[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    k[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    k[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    k[3] = 1.f - k[0] - k[1] - k[2];
}
__device__ __forceinline__ int sat_short_rhe(float v) { return max(min(__float2int_rn(v), 32767), -32768); }


// OpenCV 2.4.13 cubic source index / fixed-point taps for one output coordinate: scale = 1/(n_out/n_in)
// in double; the [0, n_in-1] clamp of (s, f) is applied along x only.
__device__ __forceinline__ void cubic_cv_coord_scaled(int d, int n_in, double scale, bool is_x, int& s, int (&q)[4]) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    s = (int)floorf(f);
    f -= (float)s;
    if (is_x) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= n_in - 1) { f = 0.f; s = n_in - 1; }
    }
    float k[4];
    cubic_cv(f, k);
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = sat_short_rhe(k[j] * 2048.f);
}
__device__ __forceinline__ void cubic_cv_coord(int d, int n_in, int n_out, bool is_x, int& s, int (&q)[4]) {
    cubic_cv_coord_scaled(d, n_in, 1. / ((double)n_out / (double)n_in), is_x, s, q);
}

}  // namespace vacv
