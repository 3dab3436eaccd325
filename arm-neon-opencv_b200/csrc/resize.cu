// a5-a9: resize INTER_LINEAR / INTER_CUBIC (reference: src/cv/resize.cpp:42-100, src/cv/resize_naive.cpp,
// src/cv/resize_neon.cpp; u8 cubic = OpenCV 2.4.13 cv::resize, the reference's only path for it).
//
// Coefficients are rebuilt on the device with the reference's exact expression order (fp32 / fp64 mix,
// --fmad=false), so the entry points are stateless like the reference's.  A CTA owns a 32x8 tile of output
// pixels of one image; x-coefficients are computed once per column and y-coefficients once per row of the tile
// into shared memory, then each thread gathers its taps through the read-only path.
#include <vector>
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include "host_util.cuh"
#include "vacv_common.cuh"
#include "resize_coeffs.cuh"
#include "gather_u8c3.cuh"
#include "resize_pipe_u8c3.cuh"
#include "resize_pipe_u8c1.cuh"

namespace vacv {

constexpr int kTileX = 32, kTileY = 8;

struct ResizeGeom {
    int w, h, c, wo, ho;
    size_t src_image, dst_image;   // elements between consecutive images (frames for HWC, planes for CHW)
};

// ----------------------------------------------------------------------------------------------------
// a5 bilinear u8, naive rule (resize_naive.cpp:10-68) / a7 NEON rule (resize_neon.cpp:12-347)
template <bool kSigned, bool kNeonRule>
__global__ void __launch_bounds__(kTileX * kTileY) resize_linear_u8_kernel(const uint8_t* __restrict__ src,
                                                                            uint8_t* __restrict__ dst, ResizeGeom g) {
    __shared__ int s_sx[kTileX], s_cx[kTileX], s_sy[kTileY], s_cy[kTileY];   // c = c0 | c1 << 16
    const int dx0 = blockIdx.x * kTileX, dy0 = blockIdx.y * kTileY;
    const int t = threadIdx.y * kTileX + threadIdx.x;
    if (t < kTileX + kTileY) {
        const bool isx = t < kTileX;
        const int d = isx ? dx0 + t : dy0 + (t - kTileX);
        const int n_in = isx ? g.w : g.h, n_out = isx ? g.wo : g.ho;
        // naive: fp32 scale (resize_naive.cpp:17-18); NEON: fp64 scale (resize_neon.cpp:17-18)
        const double scale = kNeonRule ? (double)n_in / (double)n_out : (double)((float)n_in / (float)n_out);
        int s; float f;
        linear_coord(min(d, n_out - 1), scale, n_in, s, f);
        const int c0 = sat_short((1.f - f) * 2048.f), c1 = sat_short(f * 2048.f);
        if (isx) { s_sx[t] = s; s_cx[t] = (c0 & 0xffff) | (c1 << 16); }
        else { s_sy[t - kTileX] = s; s_cy[t - kTileX] = (c0 & 0xffff) | (c1 << 16); }
    }
    __syncthreads();
    const int dx = dx0 + threadIdx.x, dy = dy0 + threadIdx.y;
    if (dx >= g.wo || dy >= g.ho) return;
    const int sx = s_sx[threadIdx.x], sy = s_sy[threadIdx.y];
    const int cx0 = (short)(s_cx[threadIdx.x] & 0xffff), cx1 = s_cx[threadIdx.x] >> 16;
    const int cy0 = (short)(s_cy[threadIdx.y] & 0xffff), cy1 = s_cy[threadIdx.y] >> 16;
    const uint8_t* img = src + blockIdx.z * g.src_image;
    const uint8_t* lt = img + ((size_t)sy * g.w + sx) * g.c;
    const uint8_t* lb = lt + (size_t)g.w * g.c;
    uint8_t* o = dst + blockIdx.z * g.dst_image + ((size_t)dy * g.wo + dx) * g.c;
    for (int k = 0; k < g.c; ++k) {
        const int p00 = pix<kSigned>(__ldg(lt + k)), p01 = pix<kSigned>(__ldg(lt + g.c + k));
        const int p10 = pix<kSigned>(__ldg(lb + k)), p11 = pix<kSigned>(__ldg(lb + g.c + k));
        int v;
        if (kNeonRule) {
            const int r0 = (short)((p00 * cx0 + p01 * cx1) >> 4), r1 = (short)((p10 * cx0 + p11 * cx1) >> 4);
            v = clamp255(((short)((cy0 * r0) >> 16) + (short)((cy1 * r1) >> 16) + 2) >> 2);
        } else {
            v = (p00 * cx0 * cy0 + p10 * cx0 * cy1 + p01 * cx1 * cy0 + p11 * cx1 * cy1) >> 22;   // :60-65
        }
        o[k] = (uint8_t)v;
    }
}

// ----------------------------------------------------------------------------------------------------
// a5 / a7 for interleaved 3-channel u8 (the C1 shape).  The generic kernel above issues 12 byte loads and 3 strided byte
// stores per pixel and is LSU-bound; here the 6 contiguous bytes of a tap row (2 pixels x BGR) come from 2-3 aligned
// 32-bit words + funnel shifts, and each warp (32 consecutive pixels of one output row) re-chunks its 96 output bytes
// through shared memory into lane-contiguous 32-bit stores.
constexpr int kC3Rows = 32;   // output rows per CTA of the u8c3 kernel (4 passes of 8 rows): amortises the coefficient set-up

template <bool kSigned, bool kNeonRule, bool kLowRow>
__device__ __forceinline__ void resize_linear_u8c3_rows(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const ResizeGeom& g,
                                                        const int* s_sx, const int* s_cx, const int* s_sy, const int* s_cy,
                                                        uint32_t (*stage)[24]) {
    const int dx0 = blockIdx.x * kTileX, dy00 = blockIdx.y * kC3Rows;
    const int lane = threadIdx.x;
    const int n = min(32, g.wo - dx0);
    const unsigned sx3 = (unsigned)s_sx[lane] * 3u;
    const uint32_t cx = (uint32_t)s_cx[lane];   // cx0 | cx1 << 16
    const uint8_t* img = src + blockIdx.z * g.src_image;   // every byte offset inside one image fits 32 bits (host check)
    const unsigned row3 = (unsigned)g.w * 3u;
    uint8_t* sb = reinterpret_cast<uint8_t*>(stage[threadIdx.y]);
    uint8_t* o = dst + blockIdx.z * g.dst_image + ((size_t)(dy00 + threadIdx.y) * g.wo + dx0) * 3;
    const size_t o_step = (size_t)kTileY * g.wo * 3;
#pragma unroll 2
    for (int pass = 0; pass < kC3Rows / kTileY; ++pass, o += o_step) {
        const int ry = pass * kTileY + threadIdx.y;
        if (dy00 + ry >= g.ho) break;   // whole warp
        const int cy0 = (short)(s_cy[ry] & 0xffff), cy1 = s_cy[ry] >> 16;
        int v[3] = {0, 0, 0};
        if (lane < n) {
            const unsigned a = (unsigned)s_sy[ry] * row3 + sx3;
            uint32_t t0, t1, u0, u1;
            linear_taps_u8c3(img, a, t0, t1);
            if (kLowRow) linear_taps_u8c3(img, a + row3, u0, u1);
            int Ht[3], Hb[3] = {0, 0, 0};
            hsum_u8c3<kSigned>(t0, t1, cx, Ht);                 // p00*cx0 + p01*cx1
            if (kLowRow) hsum_u8c3<kSigned>(u0, u1, cx, Hb);    // p10*cx0 + p11*cx1
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (kNeonRule) {   // resize_neon.cpp:145-181
                    const int r0 = (short)(Ht[k] >> 4), r1 = (short)(Hb[k] >> 4);
                    v[k] = clamp255(((short)((cy0 * r0) >> 16) + (short)((cy1 * r1) >> 16) + 2) >> 2);
                } else {           // resize_naive.cpp:60-65, the same integer regrouped row-wise (no overflow: < 2^31)
                    v[k] = (Ht[k] * cy0 + Hb[k] * cy1) >> 22;
                }
            }
        }
        sb[3 * lane] = (uint8_t)v[0]; sb[3 * lane + 1] = (uint8_t)v[1]; sb[3 * lane + 2] = (uint8_t)v[2];
        __syncwarp();
        if ((reinterpret_cast<uintptr_t>(o) & 3) == 0 && n == 32) { if (lane < 24) st_stream4(o + 4 * lane, stage[threadIdx.y][lane]); }
        else for (int bb = lane; bb < 3 * n; bb += 32) o[bb] = sb[bb];
        __syncwarp();
    }
}


template <bool kSigned, bool kNeonRule>
__global__ void __launch_bounds__(kTileX * kTileY) resize_linear_u8c3_kernel(const uint8_t* __restrict__ src,
                                                                              uint8_t* __restrict__ dst, ResizeGeom g) {
    __shared__ int s_sx[kTileX], s_cx[kTileX], s_sy[kC3Rows], s_cy[kC3Rows];
    __shared__ __align__(16) uint32_t stage[kTileY][24];
    const int dx0 = blockIdx.x * kTileX, dy00 = blockIdx.y * kC3Rows;
    const int t = threadIdx.y * kTileX + threadIdx.x;
    if (t < kTileX + kC3Rows) {
        const bool isx = t < kTileX;
        const int d = isx ? dx0 + t : dy00 + (t - kTileX);
        const int n_in = isx ? g.w : g.h, n_out = isx ? g.wo : g.ho;
        const double scale = kNeonRule ? (double)n_in / (double)n_out : (double)((float)n_in / (float)n_out);
        int s; float f;
        linear_coord(min(d, n_out - 1), scale, n_in, s, f);
        const int c0 = sat_short((1.f - f) * 2048.f), c1 = sat_short(f * 2048.f);
        if (isx) { s_sx[t] = s; s_cx[t] = (c0 & 0xffff) | (c1 << 16); }
        else { s_sy[t - kTileX] = s; s_cy[t - kTileX] = (c0 & 0xffff) | (c1 << 16); }
    }
    __syncthreads();
    // Integer ratios (1080 -> 360) put every sample on a source row: the lower tap row has weight 0 everywhere.  When that holds
    // for all rows of this CTA the lower row is not read at all (a third of the source traffic instead of two thirds); the
    // test is CTA-uniform so the general case keeps its straight-line loop.
    static_assert(kC3Rows == 32 && kTileX == 32, "one lane per row of the CTA");
    const int any_low = __any_sync(0xffffffffu, dy00 + (int)threadIdx.x < g.ho && (s_cy[threadIdx.x] >> 16) != 0);
    if (any_low) resize_linear_u8c3_rows<kSigned, kNeonRule, true>(src, dst, g, s_sx, s_cx, s_sy, s_cy, stage);
    else resize_linear_u8c3_rows<kSigned, kNeonRule, false>(src, dst, g, s_sx, s_cx, s_sy, s_cy, stage);
}

// ----------------------------------------------------------------------------------------------------
// a5 / a7 for single-channel u8 planes (CHW tensors are resized plane by plane, resize.cpp:73-87).  Same structure as the
// 3-channel kernel: a warp owns 128 consecutive output pixels of a row (lane: pixels lane, lane+32, +64, +96 so that every
// gather instruction reads neighbouring addresses), the two tap bytes of a row come from one aligned 32-bit word (two when
// they straddle it), one IDP.2A forms L*cx0 + R*cx1, and the 128 output bytes leave as one lane-contiguous 32-bit store.
constexpr int kC1Cols = 128;

template <bool kSigned, bool kNeonRule, bool kLowRow>
__device__ __forceinline__ void resize_linear_u8c1_rows(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const ResizeGeom& g,
                                                        const int* s_sx, const int* s_cx, const int* s_sy, const int* s_cy,
                                                        uint32_t (*stage)[kC1Cols / 4]) {
    const int dx0 = blockIdx.x * kC1Cols, dy00 = blockIdx.y * kC3Rows;
    const int lane = threadIdx.x;
    const int n = min(kC1Cols, g.wo - dx0);
    unsigned sx[4]; uint32_t cx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sx[j] = (unsigned)s_sx[lane + 32 * j]; cx[j] = (uint32_t)s_cx[lane + 32 * j]; }
    const uint8_t* img = src + blockIdx.z * g.src_image;   // every byte offset inside one plane fits 32 bits (host check)
    uint8_t* sb = reinterpret_cast<uint8_t*>(stage[threadIdx.y]);
    uint8_t* o = dst + blockIdx.z * g.dst_image + (size_t)(dy00 + threadIdx.y) * g.wo + dx0;
    const size_t o_step = (size_t)kTileY * g.wo;
    auto taps = [&](unsigned a) -> uint32_t {   // bytes a, a+1 of the plane in the low half-word
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (a & ~3u));
        const int r = (int)(a & 3u);
        const uint32_t w0 = __ldg(wp), w1 = r == 3 ? __ldg(wp + 1) : 0u;
        return __funnelshift_r(w0, w1, r * 8);
    };
    auto hsum = [&](uint32_t pair, uint32_t c) -> int {
        return kSigned ? __dp2a_lo((int)c, (int)pair, 0) : (int)__dp2a_lo(c, pair, 0u);
    };
    for (int pass = 0; pass < kC3Rows / kTileY; ++pass, o += o_step) {
        const int ry = pass * kTileY + threadIdx.y;
        if (dy00 + ry >= g.ho) break;   // whole warp
        const int cy0 = (short)(s_cy[ry] & 0xffff), cy1 = s_cy[ry] >> 16;
        const unsigned row = (unsigned)s_sy[ry] * (unsigned)g.w;
        uint32_t t[4], u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // all loads first
            t[j] = lane + 32 * j < n ? taps(row + sx[j]) : 0u;
            u[j] = kLowRow && lane + 32 * j < n ? taps(row + (unsigned)g.w + sx[j]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int Ht = hsum(t[j], cx[j]), Hb = kLowRow ? hsum(u[j], cx[j]) : 0;
            int v;
            if (kNeonRule) {   // resize_neon.cpp:145-181
                const int r0 = (short)(Ht >> 4), r1 = (short)(Hb >> 4);
                v = clamp255(((short)((cy0 * r0) >> 16) + (short)((cy1 * r1) >> 16) + 2) >> 2);
            } else {           // resize_naive.cpp:60-65, the same integer regrouped row-wise (no overflow: < 2^31)
                v = (Ht * cy0 + Hb * cy1) >> 22;
            }
            sb[lane + 32 * j] = (uint8_t)v;
        }
        __syncwarp();
        if ((reinterpret_cast<uintptr_t>(o) & 3) == 0 && n == kC1Cols) st_stream4(o + 4 * lane, stage[threadIdx.y][lane]);
        else for (int bb = lane; bb < n; bb += 32) o[bb] = sb[bb];
        __syncwarp();
    }
}

template <bool kSigned, bool kNeonRule>
__global__ void __launch_bounds__(kTileX * kTileY) resize_linear_u8c1_kernel(const uint8_t* __restrict__ src,
                                                                              uint8_t* __restrict__ dst, ResizeGeom g) {
    __shared__ int s_sx[kC1Cols], s_cx[kC1Cols], s_sy[kC3Rows], s_cy[kC3Rows];
    __shared__ __align__(16) uint32_t stage[kTileY][kC1Cols / 4];
    const int dx0 = blockIdx.x * kC1Cols, dy00 = blockIdx.y * kC3Rows;
    const int t = threadIdx.y * kTileX + threadIdx.x;
    if (t < kC1Cols + kC3Rows) {
        const bool isx = t < kC1Cols;
        const int d = isx ? dx0 + t : dy00 + (t - kC1Cols);
        const int n_in = isx ? g.w : g.h, n_out = isx ? g.wo : g.ho;
        const double scale = kNeonRule ? (double)n_in / (double)n_out : (double)((float)n_in / (float)n_out);
        int s; float f;
        linear_coord(min(d, n_out - 1), scale, n_in, s, f);
        const int c0 = sat_short((1.f - f) * 2048.f), c1 = sat_short(f * 2048.f);
        if (isx) { s_sx[t] = s; s_cx[t] = (c0 & 0xffff) | (c1 << 16); }
        else { s_sy[t - kC1Cols] = s; s_cy[t - kC1Cols] = (c0 & 0xffff) | (c1 << 16); }
    }
    __syncthreads();
    const int any_low = __any_sync(0xffffffffu, dy00 + (int)threadIdx.x < g.ho && (s_cy[threadIdx.x] >> 16) != 0);   // see the 3-channel kernel
    if (any_low) resize_linear_u8c1_rows<kSigned, kNeonRule, true>(src, dst, g, s_sx, s_cx, s_sy, s_cy, stage);
    else resize_linear_u8c1_rows<kSigned, kNeonRule, false>(src, dst, g, s_sx, s_cx, s_sy, s_cy, stage);
}

// ----------------------------------------------------------------------------------------------------
// a6 bilinear fp32 (resize_naive.cpp:70-128)
__global__ void __launch_bounds__(kTileX * kTileY) resize_linear_f32_kernel(const float* __restrict__ src,
                                                                             float* __restrict__ dst, ResizeGeom g) {
    __shared__ int s_sx[kTileX], s_sy[kTileY];
    __shared__ float s_fx[kTileX], s_fy[kTileY];
    const int dx0 = blockIdx.x * kTileX, dy0 = blockIdx.y * kTileY;
    const int t = threadIdx.y * kTileX + threadIdx.x;
    if (t < kTileX + kTileY) {
        const bool isx = t < kTileX;
        const int d = isx ? dx0 + t : dy0 + (t - kTileX);
        const int n_in = isx ? g.w : g.h, n_out = isx ? g.wo : g.ho;
        int s; float f;
        linear_coord(min(d, n_out - 1), (double)((float)n_in / (float)n_out), n_in, s, f);
        if (isx) { s_sx[t] = s; s_fx[t] = f; } else { s_sy[t - kTileX] = s; s_fy[t - kTileX] = f; }
    }
    __syncthreads();
    const int dx = dx0 + threadIdx.x, dy = dy0 + threadIdx.y;
    if (dx >= g.wo || dy >= g.ho) return;
    const int sx = s_sx[threadIdx.x], sy = s_sy[threadIdx.y];
    const float fx = s_fx[threadIdx.x], fy = s_fy[threadIdx.y];
    const float cx0 = 1.f - fx, cx1 = fx, cy0 = 1.f - fy, cy1 = fy;
    const float* img = src + blockIdx.z * g.src_image;
    const float* lt = img + ((size_t)sy * g.w + sx) * g.c;
    const float* lb = lt + (size_t)g.w * g.c;
    float* o = dst + blockIdx.z * g.dst_image + ((size_t)dy * g.wo + dx) * g.c;
    for (int k = 0; k < g.c; ++k)   // :121-124 evaluation order
        o[k] = __ldg(lt + k) * cx0 * cy0 + __ldg(lb + k) * cx0 * cy1 + __ldg(lt + g.c + k) * cx1 * cy0 +
               __ldg(lb + g.c + k) * cx1 * cy1;
}

// a6 for interleaved 3-channel fp32 (HWC): same arithmetic as resize_linear_f32_kernel, but a warp owns 32 consecutive output
// pixels of a row for 4 passes (coefficients in registers), issues the 12 tap loads of a pixel before any use, and re-chunks its
// 96 output floats through shared memory into three lane-contiguous 128-byte stores (the generic kernel's three 12-byte-strided
// stores per lane cost 3-4 LSU wavefronts each).
__global__ void __launch_bounds__(kTileX * kTileY) resize_linear_f32c3_kernel(const float* __restrict__ src, float* __restrict__ dst, ResizeGeom g) {
    __shared__ int s_sx[kTileX], s_sy[kC3Rows];
    __shared__ float s_fx[kTileX], s_fy[kC3Rows];
    __shared__ __align__(16) float stage[kTileY][96];
    const int dx0 = blockIdx.x * kTileX, dy00 = blockIdx.y * kC3Rows;
    const int t = threadIdx.y * kTileX + threadIdx.x;
    if (t < kTileX + kC3Rows) {
        const bool isx = t < kTileX;
        const int d = isx ? dx0 + t : dy00 + (t - kTileX);
        const int n_in = isx ? g.w : g.h, n_out = isx ? g.wo : g.ho;
        int s; float f;
        linear_coord(min(d, n_out - 1), (double)((float)n_in / (float)n_out), n_in, s, f);
        if (isx) { s_sx[t] = s; s_fx[t] = f; } else { s_sy[t - kTileX] = s; s_fy[t - kTileX] = f; }
    }
    __syncthreads();
    const int lane = threadIdx.x;
    const int n = min(32, g.wo - dx0);
    const int sx = s_sx[lane];
    const float fx = s_fx[lane], cx0 = 1.f - fx, cx1 = fx;
    const float* img = src + blockIdx.z * g.src_image;
    float* sb = stage[threadIdx.y];
    for (int pass = 0; pass < kC3Rows / kTileY; ++pass) {
        const int ry = pass * kTileY + threadIdx.y, dy = dy00 + ry;
        if (dy >= g.ho) break;   // whole warp
        const float fy = s_fy[ry], cy0 = 1.f - fy, cy1 = fy;
        float v[3] = {0.f, 0.f, 0.f};
        if (lane < n) {
            const float* lt = img + ((size_t)s_sy[ry] * g.w + sx) * 3;
            const float* lb = lt + (size_t)g.w * 3;
            float a[6], b[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { a[i] = __ldg(lt + i); b[i] = __ldg(lb + i); }
#pragma unroll
            for (int k = 0; k < 3; ++k)   // resize_naive.cpp:121-124 evaluation order
                v[k] = a[k] * cx0 * cy0 + b[k] * cx0 * cy1 + a[3 + k] * cx1 * cy0 + b[3 + k] * cx1 * cy1;
        }
        sb[3 * lane] = v[0]; sb[3 * lane + 1] = v[1]; sb[3 * lane + 2] = v[2];
        __syncwarp();
        float* o = dst + blockIdx.z * g.dst_image + ((size_t)dy * g.wo + dx0) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (lane + 32 * k < 3 * n) st_stream4f(o + lane + 32 * k, sb[lane + 32 * k]);
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileX * kTileY) resize_cubic_f32_kernel(const float* __restrict__ src,
                                                                            float* __restrict__ dst, ResizeGeom g) {
    __shared__ int s_sx[kTileX], s_sy[kTileY];
    __shared__ float s_a[kTileX][4], s_b[kTileY][4];
    const int dx0 = blockIdx.x * kTileX, dy0 = blockIdx.y * kTileY;
    const int t = threadIdx.y * kTileX + threadIdx.x;
    if (t < kTileX + kTileY) {
        const bool isx = t < kTileX;
        const int d = isx ? dx0 + t : dy0 + (t - kTileX);
        int ofs; float a[4];
        cubic_naive(min(d, (isx ? g.wo : g.ho) - 1), isx ? g.w : g.h, isx ? g.wo : g.ho, ofs, a);
        if (isx) { s_sx[t] = ofs; for (int j = 0; j < 4; ++j) s_a[t][j] = a[j]; }
        else { s_sy[t - kTileX] = ofs; for (int j = 0; j < 4; ++j) s_b[t - kTileX][j] = a[j]; }
    }
    __syncthreads();
    const int dx = dx0 + threadIdx.x, dy = dy0 + threadIdx.y;
    if (dx >= g.wo || dy >= g.ho) return;
    const int sx = s_sx[threadIdx.x], sy = s_sy[threadIdx.y];
    const float a0 = s_a[threadIdx.x][0], a1 = s_a[threadIdx.x][1], a2 = s_a[threadIdx.x][2], a3 = s_a[threadIdx.x][3];
    const float b0 = s_b[threadIdx.y][0], b1 = s_b[threadIdx.y][1], b2 = s_b[threadIdx.y][2], b3 = s_b[threadIdx.y][3];
    const float* img = src + blockIdx.z * g.src_image;
    float* o = dst + blockIdx.z * g.dst_image + ((size_t)dy * g.wo + dx) * g.c;
    const int c = g.c;
    for (int k = 0; k < c; ++k) {
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* S = img + ((size_t)(sy - 1 + j) * g.w + sx) * c + k;
            r[j] = __ldg(S - c) * a0 + __ldg(S) * a1 + __ldg(S + c) * a2 + __ldg(S + 2 * c) * a3;
        }
        o[k] = r[0] * b0 + r[1] * b1 + r[2] * b2 + r[3] * b3;
    }
}

// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileX * kTileY) resize_cubic_u8_cv24_kernel(const uint8_t* __restrict__ src,
                                                                                uint8_t* __restrict__ dst, ResizeGeom g) {
    __shared__ int s_sx[kTileX], s_sy[kTileY];
    __shared__ short s_a[kTileX][4], s_b[kTileY][4];
    const int dx0 = blockIdx.x * kTileX, dy0 = blockIdx.y * kTileY;
    const int t = threadIdx.y * kTileX + threadIdx.x;
    if (t < kTileX + kTileY) {
        const bool isx = t < kTileX;
        const int n_in = isx ? g.w : g.h, n_out = isx ? g.wo : g.ho;
        const int d = min(isx ? dx0 + t : dy0 + (t - kTileX), n_out - 1);
        const double scale = 1. / ((double)n_out / (double)n_in);
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (isx) {   // 2.4: the clamp is applied along x only
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= n_in - 1) { f = 0.f; s = n_in - 1; }
        }
        float k[4];
        cubic_cv(f, k);
        for (int j = 0; j < 4; ++j) {
            const short q = (short)sat_short_rhe(k[j] * 2048.f);
            if (isx) s_a[t][j] = q; else s_b[t - kTileX][j] = q;
        }
        if (isx) s_sx[t] = s; else s_sy[t - kTileX] = s;
    }
    __syncthreads();
    const int dx = dx0 + threadIdx.x, dy = dy0 + threadIdx.y;
    if (dx >= g.wo || dy >= g.ho) return;
    const int sx = s_sx[threadIdx.x], sy = s_sy[threadIdx.y];
    int xi[4], a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        xi[j] = min(max(sx - 1 + j, 0), g.w - 1) * g.c;
        a[j] = s_a[threadIdx.x][j];
        b[j] = s_b[threadIdx.y][j];
    }
    const uint8_t* img = src + blockIdx.z * g.src_image;
    const uint8_t* rows[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rows[j] = img + (size_t)min(max(sy - 1 + j, 0), g.h - 1) * g.w * g.c;
    uint8_t* o = dst + blockIdx.z * g.dst_image + ((size_t)dy * g.wo + dx) * g.c;
    const int vec_end = (g.wo * g.c) & ~7;
    const float sc = 1.f / (2048 * 2048);
    const float fb0 = (float)b[0] * sc, fb1 = (float)b[1] * sc, fb2 = (float)b[2] * sc, fb3 = (float)b[3] * sc;
    for (int k = 0; k < g.c; ++k) {
        int H[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            H[j] = __ldg(rows[j] + xi[0] + k) * a[0] + __ldg(rows[j] + xi[1] + k) * a[1] +
                   __ldg(rows[j] + xi[2] + k) * a[2] + __ldg(rows[j] + xi[3] + k) * a[3];
        int v;
        if (dx * g.c + k < vec_end) {
            float f = (float)H[0] * fb0;
            f = f + (float)H[1] * fb1;
            f = f + (float)H[2] * fb2;
            f = f + (float)H[3] * fb3;
            v = max(min(__float2int_rn(f), 32767), -32768);
        } else {
            v = (H[0] * b[0] + H[1] * b[1] + H[2] * b[2] + H[3] * b[3] + (1 << 21)) >> 22;
        }
        o[k] = (uint8_t)clamp255(v);
    }
}

}  // namespace vacv

namespace vacv {
int try_launch_resize_tiled(int kind, const void* src, void* dst, int images, int w, int h, int c, int wo, int ho, cudaStream_t s);
}
using namespace vacv;

// Host copy of linear_coord's source index (identical IEEE arithmetic on x86-64: no FMA, same rounding).
static int host_linear_index_r(int d, double scale, int n_in) {
    float fx = (float)(((double)d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    if (sx < 0) sx = 0;
    if (sx >= n_in - 1) sx = n_in - 2;
    return sx;
}

template <bool kSigned, bool kBand, int OUT = kRpOutU8>
static const void* resize_pipe_kernel_for(int ncol) {
    switch (ncol) {
        case 1: return (const void*)resize_linear_u8c3_pipe_kernel<kSigned, 1, kBand, false, OUT>;
        case 2: return (const void*)resize_linear_u8c3_pipe_kernel<kSigned, 2, kBand, false, OUT>;
        case 3: return (const void*)resize_linear_u8c3_pipe_kernel<kSigned, 3, kBand, false, OUT>;
        default: return (const void*)resize_linear_u8c3_pipe_kernel<kSigned, 4, kBand, false, OUT>;
    }
}

template <bool kSigned, bool kBand>
static const void* resize_pipe_c1_kernel_for(int ncol) {
    switch (ncol) {
        case 1: return (const void*)resize_linear_u8c1_pipe_kernel<kSigned, 1, kBand>;
        case 2: return (const void*)resize_linear_u8c1_pipe_kernel<kSigned, 2, kBand>;
        case 3: return (const void*)resize_linear_u8c1_pipe_kernel<kSigned, 3, kBand>;
        default: return (const void*)resize_linear_u8c1_pipe_kernel<kSigned, 4, kBand>;
    }
}

// Persistent TMA kernel for u8 bilinear, BGR (resize_pipe_u8c3.cuh) or single planes (resize_pipe_u8c1.cuh).
// 1 = launched, 0 = shape not eligible, < 0 = error.
// Shape-dependent part of a launch, cached per host thread (a stream of equally shaped calls pays it once).
struct ResizePipePlan {
    int w, h, wo, ho, device, out_mode, c, knob_gen; bool signed_char, dst4;   // key (out_mode: kRpOut*, c: 3 = BGR, 1 = planes; dst4: 4-byte aligned dst)
    bool eligible; ResizePipeGeom g; const void* kern; int threads, per_sm, sms; size_t smem;
    int quad_tx;   // > 0: the quad kernel for planes (resize_linear_u8c1_quad_kernel), threads per row group
};

static int build_resize_pipe_plan(ResizePipePlan& plan) {
    const int w = plan.w, h = plan.h, wo = plan.wo, ho = plan.ho;
    const bool signed_char = plan.signed_char;
    plan.eligible = false;
    const int c = plan.c;
    if (((size_t)w * c) % 16 != 0) return 0;                                         // bulk copies: 16-byte granularity
    if (wo > kRpThreads * kRpMaxCols || ho > 8192) return 0;
    const double scale_y = (double)((float)h / (float)ho);
    std::vector<int> sy(ho), cy(ho);   // cy: only "lower tap has weight" (bit 16), which is all tile_rows() looks at
    for (int d = 0; d < ho; ++d) {
        sy[d] = host_linear_index_r(d, scale_y, h);
        float fy = (float)(((double)d + 0.5) * scale_y - 0.5);
        int s0 = (int)floorf(fy);
        fy -= (float)s0;
        if (s0 < 0) fy = 0.f;
        if (s0 >= h - 1) fy = 1.f;
        const float x = 2048.f * fy;
        cy[d] = (int)(x + (x >= 0.f ? 0.5f : -0.5f)) != 0 ? 1 << 16 : 0;
    }
    // measured on B200 (bench_ops.py c1 / ops / ops2): the contiguous-band variant wins for small ratios (1080p -> 720p: gather
    // 0.254 ms, band 0.162 ms), the row-list variant for integer ratios where no lower tap has weight (1080p -> 640x360: gather
    // 0.233 ms, row list 0.206 ms); in between (ratio > 2 with live lower taps) the gather kernel is the fastest.
    bool any_low = false;
    for (int d = 0; d < ho; ++d) any_low = any_low || cy[d] != 0;
    const bool band = h <= 2 * ho && any_low;
    if (!band && any_low) return 0;
    const size_t row_bytes = (size_t)w * c;
    ResizePipeGeom& g = plan.g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = row_bytes * h; g.dst_image = (size_t)wo * ho * c;
    plan.quad_tx = 0;
    if (c == 1 && (wo % 4) == 0 && plan.dst4 && knob(kKnobRpipeNcol) == 0) {   // planes: quads of four adjacent columns per thread
        const int quads = wo / 4;
        const int nq = (quads + kRpThreads - 1) / kRpThreads;                  // quads per thread: 1 (w_out <= 1536) or 2
        if (nq <= 2) {
            const int tx = (((quads + nq - 1) / nq) + 31) & ~31;
            int best_TH = 0; size_t best_smem = 0;
            for (int TH = kQuadMaxTH; TH >= 1; --TH) {
                int rows = 0, tmp[2 * kQuadMaxTH];
                for (int d0 = 0; d0 < ho; d0 += TH)
                    rows = std::max(rows, band ? sy[std::min(d0 + TH, ho) - 1] + 1 - sy[d0] + 1 : tile_rows(sy.data(), cy.data(), d0, std::min(TH, ho - d0), tmp, nullptr));
                const size_t stage = ((size_t)rows * row_bytes + 16 + 127) & ~(size_t)127;
                const int table_bytes = ((3 * ho + (band ? 0 : ((ho + TH - 1) / TH) * (1 + 2 * kQuadMaxTH))) * (int)sizeof(int) + 127) & ~127;
                const size_t smem = table_bytes + kQuadStages * stage;
                if (smem + 64 <= 113 * 1024 || (TH == 1 && smem + 64 <= 226 * 1024)) { best_TH = TH; best_smem = smem; g.stage_bytes = (int)stage; g.table_bytes = table_bytes; break; }
            }
            if (best_TH) {
                g.TH = best_TH;
                g.tiles_per_frame = (ho + best_TH - 1) / best_TH;
                g.total_tiles = 0;
                // row groups: as many as fit the CTA, each with at least two consecutive output rows of the tile
                const int ry = std::max(1, std::min(kRpThreads / tx, best_TH / 2));
                const void* kern = nq == 1 ? (band ? (signed_char ? (const void*)resize_linear_u8c1_quad_kernel<true, true, 1> : (const void*)resize_linear_u8c1_quad_kernel<false, true, 1>)
                                                   : (signed_char ? (const void*)resize_linear_u8c1_quad_kernel<true, false, 1> : (const void*)resize_linear_u8c1_quad_kernel<false, false, 1>))
                                           : (band ? (signed_char ? (const void*)resize_linear_u8c1_quad_kernel<true, true, 2> : (const void*)resize_linear_u8c1_quad_kernel<false, true, 2>)
                                                   : (signed_char ? (const void*)resize_linear_u8c1_quad_kernel<true, false, 2> : (const void*)resize_linear_u8c1_quad_kernel<false, false, 2>));
                int optin = 0, per_sm = 0;
                cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, plan.device);
                cudaFuncAttributes fa;
                cudaError_t e = cudaFuncGetAttributes(&fa, kern);
                if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
                if ((size_t)(optin - (int)fa.sharedSizeBytes) >= best_smem) {
                    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
                    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
                    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tx * ry, best_smem);
                    if (e == cudaSuccess && per_sm >= 1) {
                        plan.kern = kern; plan.threads = tx * ry; plan.per_sm = per_sm; plan.sms = sm_count(plan.device); plan.smem = best_smem;
                        plan.quad_tx = tx;
                        plan.eligible = true;
                        return 0;
                    }
                }
            }
        }
    }
    int ncol = (wo + kRpThreads - 1) / kRpThreads;
    if (ncol < 2) ncol = 2;
    if (const int v = knob(kKnobRpipeNcol)) { if (v >= 1 && v <= kRpMaxCols && (wo + v - 1) / v <= kRpThreads) ncol = v; }   // tuning knob
    const int threads = std::min(kRpThreads, ((wo + ncol - 1) / ncol + 31) & ~31);
    const size_t lines = (size_t)(threads / 32) * ncol * (c == 1 ? 32 : plan.out_mode == kRpOutF32HWC ? 384 : 96);
    int best_TH = 0; size_t best_smem = 0;
    for (int TH = kRpMaxTH; TH >= 1; --TH) {
        int rows = 0, tmp[2 * kRpMaxTH];
        for (int d0 = 0; d0 < ho; d0 += TH)
            rows = std::max(rows, band ? sy[std::min(d0 + TH, ho) - 1] + 1 - sy[d0] + 1 : tile_rows(sy.data(), cy.data(), d0, std::min(TH, ho - d0), tmp, nullptr));
        const size_t stage = ((size_t)rows * row_bytes + 16 + 127) & ~(size_t)127;
        const int table_bytes = ((3 * ho + (band ? 0 : ((ho + TH - 1) / TH) * (1 + 2 * kRpMaxTH))) * (int)sizeof(int) + 127) & ~127;
        const size_t smem = table_bytes + 2 * stage + lines;
        if (smem + 64 <= 113 * 1024 || (TH == 1 && smem + 64 <= 226 * 1024)) { best_TH = TH; best_smem = smem; g.stage_bytes = (int)stage; g.table_bytes = table_bytes; break; }
    }
    if (!best_TH) return 0;
    g.TH = best_TH;
    g.tiles_per_frame = (ho + best_TH - 1) / best_TH;
    g.total_tiles = 0;   // per call
    bool any_right = false;   // does any right tap have weight? (same evaluation as linear_coord + sat_short on the device)
    {
        const double scale_x = (double)((float)w / (float)wo);
        for (int d = 0; d < wo && !any_right; ++d) {
            float fx = (float)(((double)d + 0.5) * scale_x - 0.5);
            int sx = (int)floorf(fx);
            fx -= (float)sx;
            if (sx < 0) fx = 0.f;
            if (sx >= w - 1) fx = 1.f;
            const float x = 2048.f * fx;
            any_right = (int)(x + (x >= 0.f ? 0.5f : -0.5f)) != 0;
        }
    }
    const void* kern = c == 1 ? (band ? (signed_char ? resize_pipe_c1_kernel_for<true, true>(ncol) : resize_pipe_c1_kernel_for<false, true>(ncol))
                                      : (signed_char ? resize_pipe_c1_kernel_for<true, false>(ncol) : resize_pipe_c1_kernel_for<false, false>(ncol)))
                     : plan.out_mode == kRpOutF32CHW ? (band ? resize_pipe_kernel_for<false, true, kRpOutF32CHW>(ncol) : resize_pipe_kernel_for<false, false, kRpOutF32CHW>(ncol))
                     : plan.out_mode == kRpOutF32HWC ? (band ? resize_pipe_kernel_for<false, true, kRpOutF32HWC>(ncol) : resize_pipe_kernel_for<false, false, kRpOutF32HWC>(ncol))
                     : band ? (signed_char ? resize_pipe_kernel_for<true, true>(ncol) : resize_pipe_kernel_for<false, true>(ncol))
                     : !any_right && ncol == 2 ? (const void*)resize_linear_u8c3_pipe_kernel<false, 2, false, true>   // pure byte moves: signedness irrelevant
                            : (signed_char ? resize_pipe_kernel_for<true, false>(ncol) : resize_pipe_kernel_for<false, false>(ncol));
    const int dev = plan.device;
    int optin = 0, per_sm = 0;
    const int sms = sm_count(dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    if ((size_t)(optin - (int)fa.sharedSizeBytes) < best_smem) return 0;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, best_smem);
    if (e != cudaSuccess || per_sm < 1) return 0;
    plan.kern = kern; plan.threads = threads; plan.per_sm = per_sm; plan.sms = sms; plan.smem = best_smem;
    plan.eligible = true;
    return 0;
}

// Persistent TMA kernel for u8 BGR bilinear (resize_pipe_u8c3.cuh).  1 = launched, 0 = shape not eligible, < 0 = error.
// out_mode kRpOutU8: dst = u8 BGR; kRpOutF32CHW / kRpOutF32HWC (resize_normalize): dst = fp32, mean / stddev = 3 floats each (device).
int vacv::try_launch_resize_pipe_u8c3(const uint8_t* src, void* dst, int images, int w, int h, int wo, int ho, bool signed_char, int out_mode,
                                      const float* mean, const float* stddev, cudaStream_t s, int c) {
    if (((uintptr_t)src & 15) != 0) return 0;
    static thread_local PlanCache<ResizePipePlan, 8> cache;
    const int dev = current_device(), gen = knob_generation();
    const bool dst4 = ((uintptr_t)dst & 3) == 0;
    ResizePipePlan* pp = cache.find([&](const ResizePipePlan& p) {
        return p.w == w && p.h == h && p.wo == wo && p.ho == ho && p.signed_char == signed_char && p.device == dev && p.out_mode == out_mode &&
               p.c == c && p.knob_gen == gen && p.dst4 == dst4;
    });
    if (!pp) {
        pp = cache.claim();
        pp->w = w; pp->h = h; pp->wo = wo; pp->ho = ho; pp->signed_char = signed_char; pp->device = dev; pp->out_mode = out_mode; pp->c = c; pp->knob_gen = gen; pp->dst4 = dst4;
        const int rc = build_resize_pipe_plan(*pp);
        if (rc < 0) return rc;
        cache.commit();
        ++cache.builds;
    }
    const ResizePipePlan& plan = *pp;
    if (!plan.eligible) return 0;
    ResizePipeGeom g = plan.g;
    const long long total = (long long)g.tiles_per_frame * images;
    if (total > 0x7fffffffLL - 4096) return 0;
    g.total_tiles = (int)total;
    const int grid = (int)std::min<long long>(total, (long long)plan.sms * plan.per_sm);
    int tx = plan.quad_tx;
    void* args[] = {(void*)&src, (void*)&dst, (void*)&g, (void*)&mean, (void*)&stddev};
    void* quad_args[] = {(void*)&src, (void*)&dst, (void*)&g, (void*)&tx};
    const cudaError_t e = cudaLaunchKernel(plan.kern, dim3(grid), dim3(plan.threads), tx > 0 ? quad_args : args, plan.smem, s);
    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    return 1;
}


extern "C" int vacv_cuda_resize(const void* src, void* dst, int batch, int w, int h, int c, int dtype, int layout,
                                int w_out, int h_out, int interpolation, int flags, void* stream) {
    VACV_REQUIRE(src && dst, "resize: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0 && w_out > 0 && h_out > 0, "resize: non-positive size");
    if (interpolation != VACV_INTER_LINEAR && interpolation != VACV_INTER_CUBIC)
        return set_error(VACV_ERR_UNSUPPORTED, "resize: interpolation %d (reference native paths: LINEAR, CUBIC)", interpolation);
    if (dtype != VACV_INT8 && dtype != VACV_FP32) return set_error(VACV_ERR_UNSUPPORTED, "resize: dtype %d", dtype);
    cudaStream_t s = as_stream(stream);
    const size_t es = elem_size(dtype);
    if (w_out == w && h_out == h) {   // resize.cpp:58-61 (intended: the whole buffer, App. C-7)
        cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)batch * w * h * c * es, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
        return VACV_OK;
    }
    const bool cubic = interpolation == VACV_INTER_CUBIC;
    if (cubic) VACV_REQUIRE(dtype == VACV_INT8 || (w >= 4 && h >= 4), "resize: fp32 cubic needs w,h >= 4");
    else VACV_REQUIRE(w >= 2 && h >= 2, "resize: bilinear needs w,h >= 2");
    if (cubic && dtype == VACV_INT8 && layout != VACV_NHWC && c != 1)
        return set_error(VACV_ERR_UNSUPPORTED, "resize: u8 INTER_CUBIC is HWC only (cv::resize on a cv::Mat)");
    ResizeGeom g;
    g.w = w; g.h = h; g.wo = w_out; g.ho = h_out;
    int images;
    if (layout == VACV_NHWC) { g.c = c; images = batch; }
    else { g.c = 1; images = batch * c; }   // resize.cpp:73-87: per-plane calls with c = 1
    g.src_image = (size_t)w * h * g.c;
    g.dst_image = (size_t)w_out * h_out * g.c;
    {   // shared-memory tiled kernels (resize_tiled.cu); kinds: 0 u8 linear, 1 signed-char, 2 NEON rule, 3 f32 linear, 4 f32 cubic, 5 u8 cubic
        const int kind = cubic ? (dtype == VACV_FP32 ? 4 : 5)
                               : (dtype == VACV_FP32 ? 3 : (flags & VACV_FLAG_NEON_RULE) ? 2 : (flags & VACV_FLAG_SIGNED_CHAR) ? 1 : 0);
        // measured on B200: staging pays for the 16-tap bicubic kernels, not (yet) for the 4-tap bilinear ones
        const bool tiled = (flags & VACV_FLAG_TILED) || (cubic && !(flags & VACV_FLAG_DIRECT_GATHER));
        if (tiled) {
            const int rc = try_launch_resize_tiled(kind, src, dst, images, w, h, g.c, w_out, h_out, s);
            if (rc < 0) return rc;
            if (rc > 0) return check_launch("resize (tiled)");
        }
    }
    dim3 block(kTileX, kTileY);
    for (int i0 = 0; i0 < images; i0 += 65535) {
        const int ni = min(images - i0, 65535);
        dim3 grid(ceil_div(w_out, kTileX), ceil_div(h_out, kTileY), ni);
        const uint8_t* sp = (const uint8_t*)src + (size_t)i0 * g.src_image * es;
        uint8_t* dp = (uint8_t*)dst + (size_t)i0 * g.dst_image * es;
        const bool c3_words = g.c == 3 && (((size_t)w * h * 3) % 4) == 0 && ((uintptr_t)src % 4) == 0 && (size_t)w * h * 3 < 0xfffffff0ull;
        if (!cubic && dtype == VACV_INT8 && c3_words && !(flags & (VACV_FLAG_NEON_RULE | VACV_FLAG_DIRECT_GATHER)) && !knob(kKnobNoRpipe)) {
            const int rcp = try_launch_resize_linear3_period(sp, dp, ni, w, h, w_out, h_out, (flags & VACV_FLAG_SIGNED_CHAR) != 0, s);   // rational scales
            if (rcp < 0) return rcp;
            if (rcp > 0) continue;
            const int rc = try_launch_resize_pipe_u8c3(sp, dp, ni, w, h, w_out, h_out, (flags & VACV_FLAG_SIGNED_CHAR) != 0, kRpOutU8, nullptr, nullptr, s);
            if (rc < 0) return rc;
            if (rc > 0) continue;
        }
        if (!cubic && dtype == VACV_INT8 && c3_words) {
            const bool sc = flags & VACV_FLAG_SIGNED_CHAR;
            grid.y = ceil_div(h_out, kC3Rows);
            if (flags & VACV_FLAG_NEON_RULE) resize_linear_u8c3_kernel<false, true><<<grid, block, 0, s>>>(sp, dp, g);
            else if (sc) resize_linear_u8c3_kernel<true, false><<<grid, block, 0, s>>>(sp, dp, g);
            else resize_linear_u8c3_kernel<false, false><<<grid, block, 0, s>>>(sp, dp, g);
        } else if (!cubic && dtype == VACV_INT8 && g.c == 1 && (((size_t)w * h) % 4) == 0 && ((uintptr_t)src % 4) == 0 && (size_t)w * h < 0xfffffff0ull) {
            if (!(flags & (VACV_FLAG_NEON_RULE | VACV_FLAG_DIRECT_GATHER)) && !knob(kKnobNoRpipe)) {   // persistent TMA pipeline for planes
                const int rcp = try_launch_resize_linear1_period(sp, dp, ni, w, h, w_out, h_out, (flags & VACV_FLAG_SIGNED_CHAR) != 0, s);   // rational scales
                if (rcp < 0) return rcp;
                if (rcp > 0) continue;
                const int rc = try_launch_resize_pipe_u8c3(sp, dp, ni, w, h, w_out, h_out, (flags & VACV_FLAG_SIGNED_CHAR) != 0, kRpOutU8, nullptr, nullptr, s, 1);
                if (rc < 0) return rc;
                if (rc > 0) continue;
            }
            const bool sc = flags & VACV_FLAG_SIGNED_CHAR;
            grid.x = ceil_div(w_out, kC1Cols);
            grid.y = ceil_div(h_out, kC3Rows);
            if (flags & VACV_FLAG_NEON_RULE) resize_linear_u8c1_kernel<false, true><<<grid, block, 0, s>>>(sp, dp, g);
            else if (sc) resize_linear_u8c1_kernel<true, false><<<grid, block, 0, s>>>(sp, dp, g);
            else resize_linear_u8c1_kernel<false, false><<<grid, block, 0, s>>>(sp, dp, g);
        } else if (!cubic && dtype == VACV_INT8) {
            const bool sc = flags & VACV_FLAG_SIGNED_CHAR;
            if (flags & VACV_FLAG_NEON_RULE) resize_linear_u8_kernel<false, true><<<grid, block, 0, s>>>(sp, dp, g);
            else if (sc) resize_linear_u8_kernel<true, false><<<grid, block, 0, s>>>(sp, dp, g);
            else resize_linear_u8_kernel<false, false><<<grid, block, 0, s>>>(sp, dp, g);
        } else if (!cubic && g.c == 3) {
            grid.y = ceil_div(h_out, kC3Rows);
            resize_linear_f32c3_kernel<<<grid, block, 0, s>>>((const float*)sp, (float*)dp, g);
        } else if (!cubic) {
            resize_linear_f32_kernel<<<grid, block, 0, s>>>((const float*)sp, (float*)dp, g);
        } else if (dtype == VACV_FP32) {
            resize_cubic_f32_kernel<<<grid, block, 0, s>>>((const float*)sp, (float*)dp, g);
        } else {
            resize_cubic_u8_cv24_kernel<<<grid, block, 0, s>>>(sp, dp, g);
        }
    }
    return check_launch("resize");
}
