/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the vacv operator arithmetic (the "oracle").
 *
 * Plain-C re-derivation of what the reference (b1xian/arm-neon-opencv, naive x86 path + the bundled
 * OpenCV 2.4.13 for u8 INTER_CUBIC) computes for the hot path named by BASELINE.json.  Each function
 * cites the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * arm may use it -- as the checker, never as the product.
 *
 * PARITY PINNING: every function here is checked bit-for-bit (integer ops) / to the stated tolerance
 * (fp32 ops) against the reference itself compiled from its own sources (oracle/_ref, see Makefile)
 * by tests/test_oracle_vs_ref.py, and against committed golden digests (tests/golden/) that were
 * produced by that compiled reference (tests/golden/make_golden.py).
 *
 * Conventions: dense tensors, no row pitch (tensor.cpp:524).  layout: 0 = CHW planes, 1 = HWC.
 * "u8" pixels have unsigned-char semantics unless signed_char != 0.
 */
#ifndef VACV_ORACLE_H
#define VACV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cvt_color.cpp:39-135.  src = Y plane (w*h) + interleaved chroma (w*h/2); dst = HWC BGR.
 * v_first=1: chroma byte 0 is V (NV21; also what the reference does for code NV12, App. C-3);
 * v_first=0: chroma byte 0 is U (true NV12).  w,h even. */
void orc_nv_to_bgr(const uint8_t* src, int w, int h, int v_first, uint8_t* dst);

/* Next-row extension (no reference implementation: cv.h:73 enum only): pitched NV21/NV12 (format 0/1) and
 * planar I420/YV12 (2/3) surfaces through the same matrix, by re-packing into the dense VU form of orc_nv_to_bgr. */
void orc_yuv_to_bgr(const uint8_t* src, int format, int w, int h, int y_pitch, int c_pitch, uint8_t* dst);

/* image_util.cpp:9-40 (fixture generator). */
void orc_bgr_to_nv21(const uint8_t* bgr, int w, int h, uint8_t* dst);

/* crop.cpp:44-142.  elem = bytes per element (1 or 4). */
void orc_crop(const void* src, int w, int h, int c, int elem, int layout,
              int left, int top, int cw, int ch, void* dst);

/* tensor.cpp:160-182. */
void orc_hwc_to_chw(const void* src, int w, int h, int c, int elem, void* dst);
void orc_chw_to_hwc(const void* src, int w, int h, int c, int elem, void* dst);

/* tensor.cpp:459-502. */
void orc_u8_to_f32(const uint8_t* src, size_t n, float* dst);
void orc_f32_to_u8(const float* src, size_t n, uint8_t* dst);     /* truncation toward zero, domain [0,256) */

/* resize_naive.cpp:10-68 / 70-128; layout CHW = per-plane calls with c=1 (resize.cpp:73-87). */
void orc_resize_linear_u8(const uint8_t* src, int w, int h, int c, int layout,
                          uint8_t* dst, int w_out, int h_out, int signed_char);
void orc_resize_linear_f32(const float* src, int w, int h, int c, int layout,
                           float* dst, int w_out, int h_out);

/* resize_neon.cpp:12-347 (scalar restatement of the NEON rule; aarch64 hosts only in the reference -- pinned against that
 * source compiled over oracle/neon_emul/arm_neon.h, tests/test_oracle_vs_ref.py). */
void orc_resize_linear_u8_neon_rule(const uint8_t* src, int w, int h, int c, int layout,
                                    uint8_t* dst, int w_out, int h_out);

/* resize_naive.cpp:130-569 with the intended (non-aliased) coefficient buffers (App. C-2).
 * HWC requires c == 3 in the reference; any c is accepted here. */
void orc_resize_cubic_f32(const float* src, int w, int h, int c, int layout,
                          float* dst, int w_out, int h_out);

/* OpenCV 2.4.13 cv::resize(CV_8UCn, INTER_CUBIC) (source not in the reference tree; algorithm restated
 * from the published 2.4 imgwarp.cpp and pinned against the bundled binary).  HWC. */
void orc_resize_cubic_u8_cv24(const uint8_t* src, int w, int h, int c,
                              uint8_t* dst, int w_out, int h_out);

/* warp_affine.cpp:121-133: forward 2x3 -> inverse, in place, mixed float/double exactly as written. */
void orc_invert_affine(float m[6]);
/* warp_affine.cpp:76-109: get_rotation_matrix_2D((0,0),rot,scale) + aux translation (forward matrix). */
void orc_rotation_matrix(float scale, float rot_deg, const double aux[4], float m[6]);
/* warp_affine_naive.cpp:9-106.  m = INVERTED matrix.  Out-of-range destination pixels are left
 * untouched (caller pre-fills dst). */
void orc_warp_affine_u8(const uint8_t* src, int w, int h, int c, int layout,
                        uint8_t* dst, int w_out, int h_out, const float m[6], int signed_char);
void orc_warp_affine_f32(const float* src, int w, int h, int c, int layout,
                         float* dst, int w_out, int h_out, const float m[6]);

/* Exact statistics (decision App. C-4): per-channel sum and sum of squares in u64. sums[2*k]=Sx, [2*k+1]=Sxx.
 * They are ACCUMULATED into sums (caller zeroes). */
void orc_sums_u8(const uint8_t* src, size_t pixels, int c, int layout, uint64_t* sums);
/* fp32 pixels: the same two sums accumulated in fp64 (exact for integer-valued data). ACCUMULATED into sums. */
void orc_sums_f32(const float* src, size_t pixels, int c, int layout, double* sums);
void orc_finalize_mean_stddev_f64(const double* sums, int c, uint64_t n_per_channel, float* mean, float* stddev);
/* mean = Sx/N, std = sqrt(max(Sxx/N - mean^2, 0)) in double, rounded to fp32. */
void orc_finalize_mean_stddev(const uint64_t* sums, int c, uint64_t n_per_channel, float* mean, float* stddev);
/* normalize_naive.cpp:7-72 verbatim semantics (sequential fp32) -- for reporting the deviation only. */
void orc_mean_stddev_f32_sequential(const float* src, size_t pixels, int c, int layout, float* mean, float* stddev);

/* normalize_naive.cpp:74-90: (float)((double)(x - mean) / ((double)std + 1e-6)). */
void orc_normalize_f32(const float* src, size_t pixels, int c, int layout,
                       const float* mean, const float* stddev, float* dst);
void orc_normalize_u8(const uint8_t* src, size_t pixels, int c, int layout,
                      const float* mean, const float* stddev, float* dst);

/* SURVEY A.9 composition for config 2: nv->bgr, resize linear u8, u8->f32, normalize, HWC->CHW. */
void orc_nv_resize_normalize_chw(const uint8_t* src, int w, int h, int v_first, int w_out, int h_out,
                                 const float mean[3], const float stddev[3], float* dst);
/* batch of frames over `threads` host threads (frames are independent). */
void orc_nv_resize_normalize_chw_batch(const uint8_t* src, int n, int w, int h, int v_first, int w_out, int h_out,
                                       const float mean[3], const float stddev[3], float* dst, int threads);
/* config 3 composition: warp_affine u8 HWC (zero-filled OOB) -> f32 -> normalize.  m = inverted matrix. */
void orc_warp_affine_normalize(const uint8_t* src, int w, int h, int c, const float m[6], int w_out, int h_out,
                               const float* mean, const float* stddev, float* dst);

#ifdef __cplusplus
}
#endif
#endif
