// a13 / config 2: the fused single-pass pipeline
//     NV12/NV21 -> BGR (u8) -> resize INTER_LINEAR (u8) -> u8->fp32 -> normalize -> HWC->CHW
// reference chain: cvt_color.cpp:39-135 -> resize_naive.cpp:10-68 -> tensor.cpp:459-502 ->
// normalize_naive.cpp:74-90 -> tensor.cpp:160-170 (five passes over HBM); here: one read of the NV12 frame, one
// write of the fp32 planes, every intermediate stays in registers / shared memory.
//
// Bit-exactness: each of the four taps is converted to BGR and clamped to u8 exactly as the unfused chain would
// have stored it; the bilinear MAC is the reference's integer expression regrouped row-wise
//     (p00*cx0 + p01*cx1)*cy0 + (p10*cx0 + p11*cx1)*cy1          (same integer, no overflow: < 2^31)
// then >>22 gives the u8 the chain would have stored, and the final value is table[c][u8], the table holding the
// reference's exact (float)((double)(x-mean)/((double)std+1e-6)) for the 256 possible inputs.
//
// Work decomposition: a CTA owns TH output rows x TW output columns of one frame.  It stages the source rows those
// outputs touch (a contiguous band of the Y plane and of the interleaved chroma plane) in shared memory with
// 128-bit loads, so every HBM sector is fetched once and the 2x2 gathers, the shared chroma and the row reuse
// between consecutive output rows are served on chip.  Each thread owns output columns and walks down the rows,
// carrying the horizontally interpolated row (3 ints) over when the next output row starts on it.
// Stores: lanes = consecutive columns -> 128 B per warp per plane, streaming.
#include "host_util.cuh"
#include "vacv_common.cuh"
#include "fused_pipeline.cuh"
#include "tma_host.cuh"
#include "gather_u8c3.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>
#include <cstdlib>
#include <vector>

namespace vacv {

struct FusedGeom {
    int w, h, wo, ho;
    int TW, TH;            // output tile
    int ypitch, cpitch;    // shared-memory row pitch (bytes) of the Y / chroma bands
    int yrows, crows;      // band capacity in rows
    int vec;               // 1: 16-byte staging legal (pitches, offsets and base 16-byte aligned)
    // source surface and destination canvas, as in PipeGeom
    int y_pitch, c_pitch;
    size_t frame_stride, c_off, c2_off;
    int canvas_w, canvas_h, x0, y0;
    int bf16;
};

template <int FMT>
__device__ __forceinline__ void hrow(const uint8_t* __restrict__ yrow, const uint8_t* __restrict__ crow, int vplane_off,
                                     int yo, int ca, int cb, int cx0, int cx1, int (&H)[3]) {
    const int Y0 = yrow[yo], Y1 = yrow[yo + 1];
    ChromaTerms ta, tb;
    if (FMT == kFmtPlanar) {   // crow = U row, crow + vplane_off = V row
        ta = chroma_terms(crow[vplane_off + ca], crow[ca]);
        tb = chroma_terms(crow[vplane_off + cb], crow[cb]);
    } else {
        const unsigned pa = *reinterpret_cast<const uint16_t*>(crow + ca);
        const unsigned pb = *reinterpret_cast<const uint16_t*>(crow + cb);
        ta = FMT == kFmtVU ? chroma_terms(pa & 0xff, pa >> 8) : chroma_terms(pa >> 8, pa & 0xff);
        tb = FMT == kFmtVU ? chroma_terms(pb & 0xff, pb >> 8) : chroma_terms(pb >> 8, pb & 0xff);
    }
    H[0] = add_clamp255(Y0, ta.ba) * cx0 + add_clamp255(Y1, tb.ba) * cx1;
    H[1] = add_clamp255(Y0, -ta.ga) * cx0 + add_clamp255(Y1, -tb.ga) * cx1;
    H[2] = add_clamp255(Y0, ta.ra) * cx0 + add_clamp255(Y1, tb.ra) * cx1;
}

__device__ __forceinline__ void store_out(float* p, float v) { st_stream4f(p, v); }
__device__ __forceinline__ void store_out(unsigned short* p, unsigned short v) { asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory"); }

// Tiled kernel for every shape the persistent TMA pipeline does not take (pitch / width not a multiple of 16, bands too large
// for its stages, w_out > 1536).  OutT = float or unsigned short (fp16 / bf16 bits).
template <int FMT, typename OutT>
__global__ void __launch_bounds__(320) yuv_resize_normalize_chw_tiled_kernel(const uint8_t* __restrict__ src, OutT* __restrict__ dst,
                                                                             FusedGeom g, const float* __restrict__ mean,
                                                                             const float* __restrict__ stddev) {
    extern __shared__ __align__(16) uint8_t smem[];
    OutT* lut = reinterpret_cast<OutT*>(smem);                         // [3][256]
    int* s_sy = reinterpret_cast<int*>(smem + 3072);                   // [TH]
    int* s_cy = s_sy + g.TH;                                           // [TH]  cy0 | cy1 << 16
    int* s_misc = s_cy + g.TH;                                         // [4]   sx_first, sx_last
    uint8_t* ybuf = smem + 3072 + ((8 * g.TH + 16 + 15) & ~15);
    uint8_t* cbuf = ybuf + (size_t)g.yrows * g.ypitch;                 // semi-planar: VU / UV rows; planar: U rows then V rows
    const int vplane_off = g.crows * g.cpitch;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int tiles_x = (g.wo + g.TW - 1) / g.TW;
    const int dx0 = (blockIdx.x % tiles_x) * g.TW, dy0 = (blockIdx.x / tiles_x) * g.TH;
    const int tw = min(g.TW, g.wo - dx0), th = min(g.TH, g.ho - dy0);
    const size_t frame = blockIdx.y;
    const uint8_t* yplane = src + frame * g.frame_stride;
    const uint8_t* cplane = yplane + g.c_off;
    const uint8_t* c2plane = yplane + g.c2_off;

    // ---- coefficients (resize_naive.cpp:17-53) + normalisation table
    const double scale_x = (double)((float)g.w / (float)g.wo), scale_y = (double)((float)g.h / (float)g.ho);
    if (tid < th) {
        int s; float f;
        linear_coord(dy0 + tid, scale_y, g.h, s, f);
        s_sy[tid] = s;
        s_cy[tid] = (sat_short((1.f - f) * 2048.f) & 0xffff) | (sat_short(2048.f * f) << 16);
    } else if (tid == nthr - 1) {
        int s; float f;
        linear_coord(dx0, scale_x, g.w, s, f); s_misc[0] = s;
        linear_coord(dx0 + tw - 1, scale_x, g.w, s, f); s_misc[1] = s;
    }
    for (int t = tid; t < 768; t += nthr)
        lut[t] = OutOps<typename std::conditional<sizeof(OutT) == 4, float, __half>::type>::make(
            normalize_one((float)(t & 255), __ldg(mean + (t >> 8)), (double)__ldg(stddev + (t >> 8)) + 1e-6), g.bf16);
    __syncthreads();

    // ---- stage the source band
    const int y_first = s_sy[0], y_last = s_sy[th - 1] + 1;            // Y rows [y_first, y_last]
    const int c_first = y_first >> 1, c_last = y_last >> 1;            // chroma rows
    const int xb0 = s_misc[0] & ~15;                                   // Y byte columns [xb0, xb1)
    const int xb1 = min((s_misc[1] + 2 + 15) & ~15, (g.w + 15) & ~15);
    const int cb0 = FMT == kFmtPlanar ? xb0 >> 1 : xb0;                // first chroma byte column (semi-planar: 2 bytes per 2 px)
    const int ywidth = xb1 - xb0;
    if (g.vec && FMT != kFmtPlanar) {
        const int cpr = ywidth >> 4;                                   // 16-byte chunks per row
        const int ny = (y_last - y_first + 1) * cpr, nc = (c_last - c_first + 1) * cpr;
        for (int i = tid; i < ny + nc; i += nthr) {
            const bool isy = i < ny;
            const int j = isy ? i : i - ny;
            const int r = j / cpr, q = j - r * cpr;
            const uint8_t* gp = (isy ? yplane + (size_t)(y_first + r) * g.y_pitch : cplane + (size_t)(c_first + r) * g.c_pitch) + xb0 + 16 * q;
            uint8_t* sp = (isy ? ybuf + r * g.ypitch : cbuf + r * g.cpitch) + 16 * q;
            *reinterpret_cast<uint4*>(sp) = ld_stream16(gp);
        }
    } else {
        const int wy = min(xb1, g.w) - xb0;
        const int wc = FMT == kFmtPlanar ? (wy + 1) >> 1 : wy;          // chroma bytes per row of one plane
        const int ny = (y_last - y_first + 1) * wy, nc = (c_last - c_first + 1) * wc;
        const int total = ny + (FMT == kFmtPlanar ? 2 * nc : nc);
        for (int i = tid; i < total; i += nthr) {
            if (i < ny) {
                const int r = i / wy, q = i - r * wy;
                ybuf[r * g.ypitch + q] = __ldg(yplane + (size_t)(y_first + r) * g.y_pitch + xb0 + q);
            } else {
                int j = i - ny;
                const bool second = j >= nc;                            // planar: the V plane
                if (second) j -= nc;
                const int r = j / wc, q = j - r * wc;
                cbuf[(second ? vplane_off : 0) + r * g.cpitch + q] = __ldg((second ? c2plane : cplane) + (size_t)(c_first + r) * g.c_pitch + cb0 + q);
            }
        }
    }
    __syncthreads();

    // ---- compute: thread-owned columns, walk down the tile rows
    const size_t plane = (size_t)g.canvas_w * g.canvas_h;
    OutT* out = dst + frame * 3 * plane + (size_t)g.y0 * g.canvas_w + g.x0;
    for (int col = tid; col < tw; col += nthr) {
        const int dx = dx0 + col;
        int sx; float fx;
        linear_coord(dx, scale_x, g.w, sx, fx);
        const int cx0 = sat_short((1.f - fx) * 2048.f), cx1 = sat_short(2048.f * fx);
        const int yo = sx - xb0;
        const int ca = (FMT == kFmtPlanar ? sx >> 1 : sx & ~1) - cb0, cb = (FMT == kFmtPlanar ? (sx + 1) >> 1 : (sx + 1) & ~1) - cb0;
        int H0[3], H1[3];
        int have = -2;   // source row held in H1
        for (int ty = 0; ty < th; ++ty) {
            const int sy = s_sy[ty];
            const int cy = s_cy[ty];
            const int cy0 = (short)(cy & 0xffff), cy1 = cy >> 16;
            if (sy == have) {   // the row below the previous output's pair is this output's top row
                H0[0] = H1[0]; H0[1] = H1[1]; H0[2] = H1[2];
            } else if (sy + 1 != have) {
                hrow<FMT>(ybuf + (sy - y_first) * g.ypitch, cbuf + ((sy >> 1) - c_first) * g.cpitch, vplane_off, yo, ca, cb, cx0, cx1, H0);
            }
            if (sy + 1 != have) {
                hrow<FMT>(ybuf + (sy + 1 - y_first) * g.ypitch, cbuf + (((sy + 1) >> 1) - c_first) * g.cpitch, vplane_off, yo, ca, cb, cx0, cx1, H1);
                have = sy + 1;
            }
            const size_t o = (size_t)(dy0 + ty) * g.canvas_w + dx;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int v = (H0[k] * cy0 + H1[k] * cy1) >> 22;
                store_out(out + k * plane + o, lut[k * 256 + (v & 0xff)]);
            }
        }
    }
}

// ResizeNormalize::resize_normalize (resize_normalize.cpp:15-31): resize INTER_LINEAR u8 HWC -> fp32 normalised.
// One CTA per (frame, band of output rows); exact table per CTA; taps through the read-only path.
template <int C>
__global__ void __launch_bounds__(256) resize_normalize_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                                int w, int h, int wo, int ho, int rows_per_cta,
                                                                const float* __restrict__ mean, const float* __restrict__ stddev,
                                                                int out_layout) {
    __shared__ float lut[C][256];
    for (int t = threadIdx.x; t < 256 * C; t += blockDim.x)
        lut[t >> 8][t & 255] = normalize_one((float)(t & 255), __ldg(mean + (t >> 8)), (double)__ldg(stddev + (t >> 8)) + 1e-6);
    __syncthreads();
    const double scale_x = (double)((float)w / (float)wo), scale_y = (double)((float)h / (float)ho);
    const uint8_t* img = src + (size_t)blockIdx.y * w * h * C;
    const size_t plane = (size_t)wo * ho;
    float* out = dst + (size_t)blockIdx.y * plane * C;
    const int y0 = blockIdx.x * rows_per_cta, y1 = min(y0 + rows_per_cta, ho);
    for (int i = y0 * wo + threadIdx.x; i < y1 * wo; i += blockDim.x) {
        const int dy = i / wo, dx = i - dy * wo;
        int sx, sy; float fx, fy;
        linear_coord(dx, scale_x, w, sx, fx);
        linear_coord(dy, scale_y, h, sy, fy);
        const int cx0 = sat_short((1.f - fx) * 2048.f), cx1 = sat_short(2048.f * fx);
        const int cy0 = sat_short((1.f - fy) * 2048.f), cy1 = sat_short(2048.f * fy);
        const uint8_t* lt = img + ((size_t)sy * w + sx) * C;
        const uint8_t* lb = lt + (size_t)w * C;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int v = (__ldg(lt + k) * cx0 * cy0 + __ldg(lb + k) * cx0 * cy1 + __ldg(lt + C + k) * cx1 * cy0 +
                           __ldg(lb + C + k) * cx1 * cy1) >> 22;
            const float r = lut[k][v & 0xff];
            if (out_layout == VACV_NHWC) out[(size_t)i * C + k] = r;
            else st_stream4f(out + k * plane + i, r);
        }
    }
}

// 3-channel fast path of resize_normalize: the structure of resize_linear_u8c3_kernel (resize.cu) -- a warp owns 32
// consecutive output pixels of a row, tap bytes come as aligned 32-bit words + funnel shifts, PRMT + IDP.2A form the
// horizontal sums -- followed by the exact normalisation table.  CHW: one lane-contiguous 128-byte store per plane;
// HWC: the warp's 96 floats are re-chunked through shared memory into three lane-contiguous 128-byte stores.
constexpr int kRnRows = 32;   // output rows per CTA (4 passes of 8 rows)

template <bool kCHW, bool kLowRow>
__device__ __forceinline__ void resize_normalize_u8c3_rows(const uint8_t* __restrict__ src, float* __restrict__ dst, int w, int wo, int ho,
                                                           const int* s_sx, const int* s_cx, const int* s_sy, const int* s_cy,
                                                           const float* lut, float (*stage)[96]) {
    const int dx0 = blockIdx.x * 32, dy00 = blockIdx.y * kRnRows;
    const int lane = threadIdx.x;
    const int n = min(32, wo - dx0);
    const unsigned sx3 = (unsigned)s_sx[lane] * 3u;
    const uint32_t cx = (uint32_t)s_cx[lane];   // cx0 | cx1 << 16
    const unsigned row3 = (unsigned)w * 3u;
    const size_t plane = (size_t)wo * ho;
    float* out = dst + (size_t)blockIdx.z * plane * 3;
    for (int pass = 0; pass < kRnRows / 8; ++pass) {
        const int ry = pass * 8 + threadIdx.y, dy = dy00 + ry;
        if (dy >= ho) break;   // whole warp
        const int cy0 = (short)(s_cy[ry] & 0xffff), cy1 = s_cy[ry] >> 16;
        float r[3] = {0.f, 0.f, 0.f};
        if (lane < n) {
            const unsigned a = (unsigned)s_sy[ry] * row3 + sx3;
            uint32_t t0, t1, u0, u1;
            linear_taps_u8c3(src, a, t0, t1);
            if (kLowRow) linear_taps_u8c3(src, a + row3, u0, u1);
            int Ht[3], Hb[3] = {0, 0, 0};
            hsum_u8c3<false>(t0, t1, cx, Ht);                 // p00*cx0 + p01*cx1
            if (kLowRow) hsum_u8c3<false>(u0, u1, cx, Hb);    // p10*cx0 + p11*cx1
#pragma unroll
            for (int k = 0; k < 3; ++k) r[k] = lut[k * 256 + (((Ht[k] * cy0 + Hb[k] * cy1) >> 22) & 0xff)];   // resize_naive.cpp:60-65
        }
        if (kCHW) {
            if (lane < n) {
                const size_t o = (size_t)dy * wo + dx0 + lane;
#pragma unroll
                for (int k = 0; k < 3; ++k) st_stream4f(out + k * plane + o, r[k]);
            }
        } else {
            float* sb = stage[threadIdx.y];
            sb[3 * lane] = r[0]; sb[3 * lane + 1] = r[1]; sb[3 * lane + 2] = r[2];
            __syncwarp();
            float* o = out + ((size_t)dy * wo + dx0) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (lane + 32 * k < 3 * n) st_stream4f(o + lane + 32 * k, sb[lane + 32 * k]);
            __syncwarp();
        }
    }
}

template <bool kCHW>
__global__ void __launch_bounds__(256) resize_normalize_u8c3_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int w, int h,
                                                                     int wo, int ho, size_t src_image, const float* __restrict__ mean,
                                                                     const float* __restrict__ stddev) {
    __shared__ int s_sx[32], s_cx[32], s_sy[kRnRows], s_cy[kRnRows];
    __shared__ float lut[768];
    __shared__ __align__(16) float stage[8][96];
    const int dx0 = blockIdx.x * 32, dy00 = blockIdx.y * kRnRows;
    const int t = threadIdx.y * 32 + threadIdx.x;
    if (t < 32 + kRnRows) {
        const bool isx = t < 32;
        const int d = isx ? dx0 + t : dy00 + (t - 32);
        const int n_in = isx ? w : h, n_out = isx ? wo : ho;
        int s; float f;
        linear_coord(min(d, n_out - 1), (double)((float)n_in / (float)n_out), n_in, s, f);
        const int c0 = sat_short((1.f - f) * 2048.f), c1 = sat_short(f * 2048.f);
        if (isx) { s_sx[t] = s; s_cx[t] = (c0 & 0xffff) | (c1 << 16); }
        else { s_sy[t - 32] = s; s_cy[t - 32] = (c0 & 0xffff) | (c1 << 16); }
    }
    for (int i = t; i < 768; i += 256)
        lut[i] = normalize_one((float)(i & 255), __ldg(mean + (i >> 8)), (double)__ldg(stddev + (i >> 8)) + 1e-6);
    __syncthreads();
    const uint8_t* img = src + blockIdx.z * src_image;
    const int any_low = __any_sync(0xffffffffu, dy00 + (int)threadIdx.x < ho && (s_cy[threadIdx.x] >> 16) != 0);   // see resize_linear_u8c3_kernel
    if (any_low) resize_normalize_u8c3_rows<kCHW, true>(img, dst, w, wo, ho, s_sx, s_cx, s_sy, s_cy, lut, stage);
    else resize_normalize_u8c3_rows<kCHW, false>(img, dst, w, wo, ho, s_sx, s_cx, s_sy, s_cy, lut, stage);
}

}  // namespace vacv

using namespace vacv;

// Host copy of linear_coord's source index (identical IEEE arithmetic on x86-64: no FMA, same rounding).
static int host_linear_index(int d, double scale, int n_in) {
    float fx = (float)(((double)d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    if (sx < 0) sx = 0;
    if (sx >= n_in - 1) sx = n_in - 2;
    return sx;
}

template <int FMT, typename OutT, bool kDense = false, int RIGHT = -1, bool kMaps = false>
static const void* pipe_kernel_ncol(int ncol) {
    switch (ncol) {
        case 1: return (const void*)nv_resize_normalize_chw_pipe_kernel<FMT, OutT, 1, kDense, RIGHT, kMaps>;
        case 2: return (const void*)nv_resize_normalize_chw_pipe_kernel<FMT, OutT, 2, kDense, RIGHT, kMaps>;
        case 3: return (const void*)nv_resize_normalize_chw_pipe_kernel<FMT, OutT, 3, kDense, RIGHT, kMaps>;
        default: return (const void*)nv_resize_normalize_chw_pipe_kernel<FMT, OutT, 4, kDense, RIGHT, kMaps>;
    }
}
template <int FMT, int RIGHT>
static const void* pipe_kernel_pairs(int ncol) {   // fp16 column pairs: even column counts only
    return ncol <= 2 ? (const void*)nv_resize_normalize_chw_pipe_kernel<FMT, __half2, 2, false, RIGHT>
                     : (const void*)nv_resize_normalize_chw_pipe_kernel<FMT, __half2, 4, false, RIGHT>;
}
static const void* pipe_kernel_for(int fmt, bool half_out, bool pairs, bool dense, int ncol, bool any_right, bool dense_maps) {
    if (dense_maps) {   // padded NV12 / NV21 surfaces through tensor maps, fp32 planes without letterbox: the dense tile loops
        if (any_right) return fmt == kFmtVU ? pipe_kernel_ncol<kFmtVU, float, true, 1, true>(ncol) : pipe_kernel_ncol<kFmtUV, float, true, 1, true>(ncol);
        return fmt == kFmtVU ? pipe_kernel_ncol<kFmtVU, float, true, 0, true>(ncol) : pipe_kernel_ncol<kFmtUV, float, true, 0, true>(ncol);
    }
    if (half_out && pairs) {
        if (any_right) return fmt == kFmtVU ? pipe_kernel_pairs<kFmtVU, 1>(ncol) : fmt == kFmtUV ? pipe_kernel_pairs<kFmtUV, 1>(ncol) : pipe_kernel_pairs<kFmtPlanar, 1>(ncol);
        return fmt == kFmtVU ? pipe_kernel_pairs<kFmtVU, 0>(ncol) : fmt == kFmtUV ? pipe_kernel_pairs<kFmtUV, 0>(ncol) : pipe_kernel_pairs<kFmtPlanar, 0>(ncol);
    }
    if (dense && !half_out && fmt != kFmtPlanar) {   // the reference's own case keeps its dedicated instantiations: one per tap rule, so
        // that the path a launch never takes does not shape the register allocation and schedule of the one it does
        if (any_right) return fmt == kFmtVU ? pipe_kernel_ncol<kFmtVU, float, true, 1>(ncol) : pipe_kernel_ncol<kFmtUV, float, true, 1>(ncol);
        return fmt == kFmtVU ? pipe_kernel_ncol<kFmtVU, float, true, 0>(ncol) : pipe_kernel_ncol<kFmtUV, float, true, 0>(ncol);
    }
    if (half_out) return fmt == kFmtVU ? pipe_kernel_ncol<kFmtVU, __half>(ncol) : fmt == kFmtUV ? pipe_kernel_ncol<kFmtUV, __half>(ncol) : pipe_kernel_ncol<kFmtPlanar, __half>(ncol);
    if (any_right) return fmt == kFmtVU ? pipe_kernel_ncol<kFmtVU, float, false, 1>(ncol) : fmt == kFmtUV ? pipe_kernel_ncol<kFmtUV, float, false, 1>(ncol) : pipe_kernel_ncol<kFmtPlanar, float, false, 1>(ncol);
    return fmt == kFmtVU ? pipe_kernel_ncol<kFmtVU, float, false, 0>(ncol) : fmt == kFmtUV ? pipe_kernel_ncol<kFmtUV, float, false, 0>(ncol) : pipe_kernel_ncol<kFmtPlanar, float, false, 0>(ncol);
}

// Source description resolved from a vacv_yuv_layout (or the dense NV12/NV21 default).
struct YuvSource {
    int fmt;                 // kFmtVU / kFmtUV / kFmtPlanar
    int w, h, y_pitch, c_pitch;
    size_t frame_stride, c_off, c2_off;
};

// Returns 1 if the persistent TMA pipeline was launched, 0 if the shape does not qualify (caller falls back to the
// tiled kernel where one exists), < 0 on error.
struct Canvas { int w, h, x0, y0; };   // destination plane size and where the w_out x h_out result goes inside it

// Everything about a launch that depends on the shapes only: computed once per (shape, layout, output) and cached per host
// thread, so that a steady stream of equally shaped calls (one frame or one small batch per call) pays ~nothing on the host.
struct PipePlan {
    // key
    YuvSource y; int w_out, h_out, out_dtype; Canvas cv; bool pairs_ok; int device, knob_gen;
    // plan
    bool eligible; PipeGeom g; const void* kern; int threads, per_sm, sms; size_t smem;
    // padded surfaces through tensor maps (g.tile_maps): the maps bind the surface pool's address and are re-encoded when it changes
    // (a decoder pool rotates through a few surface buffers: kMapSlots encoded sets are kept, replaced round-robin)
    static constexpr int kMapSlots = 4;
    PipeMaps maps;                                   // the band heights; tensor maps of slot 0 when encoded
    PipeMaps map_slots[kMapSlots]; const void* maps_src[kMapSlots]; int maps_batch[kMapSlots]; int maps_next;
};

static bool same_key(const PipePlan& p, const YuvSource& y, int w_out, int h_out, int out_dtype, const Canvas& cv, bool pairs_ok, int device) {
    return p.y.fmt == y.fmt && p.y.w == y.w && p.y.h == y.h && p.y.y_pitch == y.y_pitch && p.y.c_pitch == y.c_pitch && p.y.frame_stride == y.frame_stride &&
           p.y.c_off == y.c_off && p.y.c2_off == y.c2_off && p.w_out == w_out && p.h_out == h_out && p.out_dtype == out_dtype && p.cv.w == cv.w &&
           p.cv.h == cv.h && p.cv.x0 == cv.x0 && p.cv.y0 == cv.y0 && p.pairs_ok == pairs_ok && p.device == device;
}

// Fills plan.eligible / g / kern / threads / per_sm / smem.  Returns < 0 on a CUDA error.
static int build_pipe_plan(PipePlan& plan) {
    const YuvSource& y = plan.y;
    const int w_out = plan.w_out, h_out = plan.h_out, out_dtype = plan.out_dtype;
    const Canvas& cv = plan.cv;
    plan.eligible = false;
    const bool half_out = out_dtype != VACV_FP32;   // fp16 / bf16: 16-bit table entries
    const int w = y.w, h = y.h;
    // bulk copies need 16-byte granularity of every band start and length
    if ((y.y_pitch % 16) != 0 || (y.c_pitch % 16) != 0 || (y.frame_stride % 16) != 0 || (y.c_off % 16) != 0 || (y.c2_off % 16) != 0) return 0;
    if (w_out > kPipeThreads * kPipeMaxCols || h_out > 8192) return 0;
    // row source indices, exactly as the device computes them -> exact band sizes per tile
    const double scale_y = (double)((float)h / (float)h_out);
    std::vector<int> sy(h_out);
    for (int d = 0; d < h_out; ++d) sy[d] = host_linear_index(d, scale_y, h);
    const int table_bytes = (2 * h_out * (int)sizeof(int) + 127) & ~127;
    const int static_bytes = 768 * 4 + 64;
    PipeGeom& g = plan.g;
    g.w = w; g.h = h; g.wo = w_out; g.ho = h_out; g.table_bytes = table_bytes;
    g.y_pitch = y.y_pitch; g.c_pitch = y.c_pitch; g.frame_stride = y.frame_stride; g.c_off = y.c_off; g.c2_off = y.c2_off;
    g.canvas_w = cv.w; g.canvas_h = cv.h; g.x0 = cv.x0; g.y0 = cv.y0; g.bf16 = out_dtype == VACV_BF16 ? 1 : 0;
    const bool planar = y.fmt == kFmtPlanar;
    // padded surfaces (decoder pools: pitch 2048 for 1920-wide frames).  Default: a band is ONE copy, padding included (7 % more DRAM
    // reads at pitch 2048).  VACV_PIPE_ROWS=1 stages only the rows' own bytes with one bulk copy per row -- measured slower on B200
    // (1080p -> 640x640 x256, same box: nv12 p2048 fp32 0.447 vs 0.407 ms, i420 p2048 fp16 0.389 vs 0.309 ms): 25-40 copies of
    // <= 1920 bytes per tile cost more in the copy engine than the padding costs in DRAM.
    // Round 2, default: one tensor-map box per band and plane (rows' own bytes only, ONE copy per plane): needs rows of whole 16-byte
    // chunks, <= 256 8-byte elements per box row and at most kPipeMapHeights distinct band heights; otherwise whole bands.
    // VACV_PIPE_ROWS=2 forces whole bands (A/B).
    const int y_row16 = (w + 15) & ~15, c_row16 = ((planar ? w / 2 : w) + 15) & ~15;
    const bool padded = y.y_pitch > y_row16 || y.c_pitch > c_row16;
    const bool by_row = padded && knob(kKnobPipeRows) == 1;
    bool by_map = padded && knob(kKnobPipeRows) == 0 && (w % 16) == 0 && (!planar || (w % 32) == 0) && w <= 2048 && encode_tiled_fn() != nullptr;
    int best_TH = 0;
    size_t best_smem = 0;
    for (int attempt = 0; attempt < 2 && !best_TH; ++attempt) {
        g.sy_pitch = by_row || by_map ? y_row16 : y.y_pitch;
        g.sc_pitch = by_row || by_map ? c_row16 : y.c_pitch;
        for (int TH = 8; TH >= 1; --TH) {
            int yrows = 0, crows = 0;
            for (int d0 = 0; d0 < h_out; d0 += TH) {
                const int d1 = std::min(d0 + TH, h_out) - 1;
                const int y0 = sy[d0], y1 = sy[d1] + 1;
                yrows = std::max(yrows, y1 - y0 + 1);
                crows = std::max(crows, (y1 >> 1) - (y0 >> 1) + 1);
            }
            const size_t ystage = ((size_t)yrows * g.sy_pitch + 127) & ~(size_t)127;
            const size_t cband = ((size_t)crows * g.sc_pitch + 127) & ~(size_t)127;
            const size_t cstage = planar ? 2 * cband : cband;
            const size_t smem = table_bytes + 2 * (ystage + cstage);
            if (smem + static_bytes <= 113 * 1024 || (TH == 1 && smem + static_bytes <= 226 * 1024)) {   // 2 CTAs / SM
                best_TH = TH; best_smem = smem; g.ystage = (int)ystage; g.cstage = (int)cstage; g.vstage_off = (int)cband;
                break;
            }
        }
        if (!by_map) break;
        if (best_TH) {   // the band heights that occur with this tile height: one tensor map each
            std::memset(&plan.maps, 0, sizeof(plan.maps));
            int ny = 0, nc = 0;
            bool fits = true;
            for (int d0 = 0; d0 < h_out && fits; d0 += best_TH) {
                const int d1 = std::min(d0 + best_TH, h_out) - 1;
                const int y0 = sy[d0], y1 = sy[d1] + 1;
                const int yr = y1 - y0 + 1, cr = (y1 >> 1) - (y0 >> 1) + 1;
                if (std::find(plan.maps.yh, plan.maps.yh + ny, yr) == plan.maps.yh + ny) { if (ny < kPipeMapHeights) plan.maps.yh[ny++] = yr; else fits = false; }
                if (std::find(plan.maps.ch, plan.maps.ch + nc, cr) == plan.maps.ch + nc) { if (nc < kPipeMapHeights) plan.maps.ch[nc++] = cr; else fits = false; }
            }
            if (fits) break;
        }
        by_map = false;   // second attempt: whole bands, padding included
        best_TH = 0;
    }
    g.tile_maps = by_map ? 1 : 0;
    for (int i = 0; i < PipePlan::kMapSlots; ++i) { plan.maps_src[i] = nullptr; plan.maps_batch[i] = 0; }
    plan.maps_next = 0;
    if (!best_TH) return 0;
    g.TH = best_TH;
    g.tiles_per_frame = (h_out + best_TH - 1) / best_TH;
    g.total_tiles = 0;   // per call
    // Columns per thread.  Measured on B200 (bench_ops.py c2/c2g): when some right tap has weight (general ratios) the
    // loop is issue-bound and 4 columns per thread (more ILP, less per-row overhead) win; when every cx1 is 0 (odd
    // integer x ratio) the kernel is HBM-bound and 2 columns per thread (twice the warps) win.
    bool any_right = false;
    {
        const double scale_x = (double)((float)w / (float)w_out);
        for (int d = 0; d < w_out && !any_right; ++d) {
            float fx = (float)(((double)d + 0.5) * scale_x - 0.5);
            int sx = (int)floorf(fx);
            fx -= (float)sx;
            if (sx < 0) fx = 0.f;
            if (sx >= w - 1) fx = 1.f;
            const float x = 2048.f * fx;
            any_right = (int)(x + (x >= 0.f ? 0.5f : -0.5f)) != 0;
        }
    }
    // Columns per thread: as few as cover the row.  (Round 1 gave launches with right taps 4 columns per thread up to 768 columns; with one
    // tap rule per kernel and the parity register sets 2 columns measure faster again -- 1080p -> 608x608 0.371 -> 0.358 ms, 720p ->
    // 640x640 0.317 -> 0.313, 1080p -> 416x416 0.234 -> 0.230, same box.)
    int ncol = (w_out + kPipeThreads - 1) / kPipeThreads;
    if (const int v = knob(kKnobPipeNcol)) { if (v >= 1 && v <= kPipeMaxCols && (w_out + v - 1) / v <= (v == 1 ? 640 : kPipeThreads)) ncol = v; }   // tuning knob
    const bool pairs = half_out && (w_out % 2) == 0 && (cv.w % 2) == 0 && (cv.x0 % 2) == 0 && plan.pairs_ok;   // 16-bit outputs: 32-bit stores of column pairs
    if (pairs) ncol = ncol <= 2 ? 2 : 4;
    const int threads = std::min(ncol == 1 ? 640 : kPipeThreads, ((w_out + ncol - 1) / ncol + 31) & ~31);
    const bool dense = y.y_pitch == w && y.c_pitch == w && y.c_off == (size_t)w * h && y.frame_stride == (size_t)w * h * 3 / 2 &&
                       cv.w == w_out && cv.h == h_out;
    const bool dense_maps = by_map && !half_out && !planar && cv.w == w_out && cv.h == h_out && g.sy_pitch == w && g.sc_pitch == w;
    const void* kern = pipe_kernel_for(y.fmt, half_out, pairs, dense, ncol, any_right, dense_maps);
    // the opt-in limit, not this shape's size: plans of other host threads for the same kernel must stay launchable
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, plan.device);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "nv_resize_normalize_chw: %s", cudaGetErrorString(e));
    const int max_dyn = optin - (int)fa.sharedSizeBytes;   // static + dynamic <= opt-in limit
    if ((size_t)max_dyn < best_smem) return 0;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "nv_resize_normalize_chw: %s", cudaGetErrorString(e));
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, best_smem);
    if (e != cudaSuccess || per_sm < 1) return 0;
    const int sms = sm_count(plan.device);
    plan.kern = kern; plan.threads = threads; plan.per_sm = per_sm; plan.sms = sms; plan.smem = best_smem;
    plan.eligible = true;
    return 0;
}

// Returns 1 if the persistent TMA pipeline was launched, 0 if the shape does not qualify (caller falls back to the
// tiled kernel), < 0 on error.
static int try_launch_pipe(const uint8_t* src, void* dst, int out_dtype, int batch, const YuvSource& y, int w_out, int h_out,
                           const Canvas& cv, const float* mean, const float* stddev, cudaStream_t s) {
    if ((((uintptr_t)src) & 15) != 0) return 0;   // bulk copies need 16-byte aligned band starts
    static thread_local PlanCache<PipePlan, 8> cache;   // keyed by device + shape: alternating shapes / GPUs on one thread do not rebuild
    const int dev = current_device(), gen = knob_generation();
    const bool pairs_ok = (((uintptr_t)dst) & 3) == 0;
    PipePlan* pp = cache.find([&](const PipePlan& p) { return p.knob_gen == gen && same_key(p, y, w_out, h_out, out_dtype, cv, pairs_ok, dev); });
    if (!pp) {
        pp = cache.claim();
        pp->y = y; pp->w_out = w_out; pp->h_out = h_out; pp->out_dtype = out_dtype; pp->cv = cv; pp->pairs_ok = pairs_ok; pp->device = dev; pp->knob_gen = gen;
        const int rc = build_pipe_plan(*pp);
        if (rc < 0) return rc;
        cache.commit();
        ++cache.builds;
    }
    const PipePlan& plan = *pp;
    if (!plan.eligible) return 0;
    PipeGeom g = plan.g;
    const long long total = (long long)g.tiles_per_frame * batch;
    if (total > 0x7fffffffLL - 4096) return 0;
    g.total_tiles = (int)total;
    const int grid = (int)std::min<long long>(total, (long long)plan.sms * plan.per_sm);
    const PipeMaps* maps = &pp->maps;   // plans without tensor maps pass the (unused) height table
    if (g.tile_maps) {   // the encoded maps of this surface pool (frames are the third dimension), or encode them into the next slot
        int slot = -1;
        for (int i = 0; i < PipePlan::kMapSlots; ++i)
            if (pp->maps_src[i] == src && pp->maps_batch[i] == batch) slot = i;
        if (slot < 0) {
            slot = pp->maps_next;
            pp->maps_next = (pp->maps_next + 1) % PipePlan::kMapSlots;
            PipeMaps& m = pp->map_slots[slot];
            m = pp->maps;
            pp->maps_src[slot] = nullptr;
            const YuvSource& ys = plan.y;
            const bool planar = ys.fmt == kFmtPlanar;
            const cuuint64_t frames = (cuuint64_t)batch;
            const int cw8 = (planar ? ys.w / 2 : ys.w) / 8, chh = (ys.h + 1) / 2;
            bool ok = true;
            for (int k = 0; k < kPipeMapHeights && ok; ++k) {
                if (m.yh[k])
                    ok = encode_map_3d(&m.y[k], CU_TENSOR_MAP_DATA_TYPE_UINT64, src, ys.w / 8, ys.h, frames, ys.y_pitch, ys.frame_stride, ys.w / 8, m.yh[k], 1);
                if (ok && m.ch[k]) {
                    ok = encode_map_3d(&m.c[k], CU_TENSOR_MAP_DATA_TYPE_UINT64, src + ys.c_off, cw8, chh, frames, ys.c_pitch, ys.frame_stride, cw8, m.ch[k], 1);
                    if (ok && planar)
                        ok = encode_map_3d(&m.c2[k], CU_TENSOR_MAP_DATA_TYPE_UINT64, src + ys.c2_off, cw8, chh, frames, ys.c_pitch, ys.frame_stride, cw8, m.ch[k], 1);
                }
            }
            if (!ok) return set_error(VACV_ERR_CUDA, "nv_resize_normalize_chw: cuTensorMapEncodeTiled failed");
            pp->maps_src[slot] = src; pp->maps_batch[slot] = batch;
        }
        maps = &pp->map_slots[slot];
    }
    void* args[] = {(void*)&src, (void*)&dst, (void*)&g, (void*)&mean, (void*)&stddev, (void*)maps};
    const cudaError_t e = cudaLaunchKernel(plan.kern, dim3(grid), dim3(plan.threads), args, plan.smem, s);
    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "nv_resize_normalize_chw: %s", cudaGetErrorString(e));
    return 1;
}


template <int FMT>
static const void* tiled_kernel_for(bool half_out) {
    return half_out ? (const void*)yuv_resize_normalize_chw_tiled_kernel<FMT, unsigned short> : (const void*)yuv_resize_normalize_chw_tiled_kernel<FMT, float>;
}

// The tiled kernel: any pitch, any width.  Returns a status code.
static int launch_fused_tiled(const char* who, const uint8_t* src, void* dst, int out_dtype, int batch, const YuvSource& y, int w_out, int h_out,
                              const Canvas& cv, const float* mean, const float* stddev, cudaStream_t s) {
    const int w = y.w, h = y.h;
    FusedGeom g;
    g.w = w; g.h = h; g.wo = w_out; g.ho = h_out;
    g.y_pitch = y.y_pitch; g.c_pitch = y.c_pitch; g.frame_stride = y.frame_stride; g.c_off = y.c_off; g.c2_off = y.c2_off;
    g.canvas_w = cv.w; g.canvas_h = cv.h; g.x0 = cv.x0; g.y0 = cv.y0; g.bf16 = out_dtype == VACV_BF16 ? 1 : 0;
    g.vec = (y.y_pitch % 16) == 0 && (y.c_pitch % 16) == 0 && (y.frame_stride % 16) == 0 && (y.c_off % 16) == 0 && (((uintptr_t)src) & 15) == 0;
    const double sx = (double)w / w_out, sy = (double)h / h_out;
    // tile: full output rows when the source span fits comfortably, else split in x
    const int budget = 44 * 1024;   // ~4 CTAs / SM
    const int cplanes = y.fmt == kFmtPlanar ? 2 : 1;
    int TW = w_out, TH = 8;
    auto span = [](double scale, int n) { return (int)(scale * (n - 1)) + 4; };
    auto bytes = [&](int tw, int th) {
        const int yp = (span(sx, tw) + 32 + 15) & ~15;
        const int yr = span(sy, th) + 1, cr = yr / 2 + 2;
        return (size_t)yp * (yr + cplanes * cr);
    };
    while (TH > 1 && bytes(TW, TH) > (size_t)budget) --TH;
    while (TW > 32 && bytes(TW, TH) > (size_t)budget) TW = (TW + 1) / 2;
    TW = (TW + 31) & ~31;
    if (bytes(TW, TH) > 200 * 1024) return set_error(VACV_ERR_UNSUPPORTED, "%s: scale too large for the staged tile", who);
    g.TW = TW; g.TH = TH;
    g.ypitch = g.cpitch = (span(sx, TW) + 32 + 15) & ~15;
    g.yrows = span(sy, TH) + 1; g.crows = g.yrows / 2 + 2;
    const size_t smem = 3072 + ((8 * TH + 16 + 15) & ~15) + (size_t)g.ypitch * (g.yrows + cplanes * g.crows);
    const int threads = 32 * std::max(1, std::min(10, (std::min(TW, w_out) + 63) / 64));   // ~2 columns per thread, <= 320
    const int tiles = ceil_div(w_out, TW) * ceil_div(h_out, TH);
    const bool half_out = out_dtype != VACV_FP32;
    const void* kern = y.fmt == kFmtVU ? tiled_kernel_for<kFmtVU>(half_out) : y.fmt == kFmtUV ? tiled_kernel_for<kFmtUV>(half_out) : tiled_kernel_for<kFmtPlanar>(half_out);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
    }
    const size_t dst_frame = (size_t)3 * cv.w * cv.h * (half_out ? 2 : 4);
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const uint8_t* sp = src + (size_t)b0 * y.frame_stride;
        void* dp = (uint8_t*)dst + (size_t)b0 * dst_frame;
        void* args[] = {(void*)&sp, (void*)&dp, (void*)&g, (void*)&mean, (void*)&stddev};
        cudaError_t e = cudaLaunchKernel(kern, dim3(tiles, std::min(batch - b0, 65535)), dim3(threads), args, smem, s);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
    }
    return check_launch(who);
}

extern "C" int vacv_cuda_nv_resize_normalize_chw(const uint8_t* src, float* dst, int batch, int w, int h, int v_first,
                                                 int w_out, int h_out, const float* mean, const float* stddev, void* stream) {
    VACV_REQUIRE(src && dst && mean && stddev, "nv_resize_normalize_chw: null pointer");
    VACV_REQUIRE(batch > 0 && w >= 2 && h >= 2 && w_out > 0 && h_out > 0, "nv_resize_normalize_chw: bad size");
    VACV_REQUIRE((w % 2) == 0 && (h % 2) == 0, "nv_resize_normalize_chw: w and h must be even (got %dx%d)", w, h);
    if (w_out == w && h_out == h)   // resize.cpp:58-61 memcpy shortcut == identity taps; not on the fused fast path
        return set_error(VACV_ERR_UNSUPPORTED, "nv_resize_normalize_chw: same-size resize (compose cvt_nv2bgr + normalize + layout_change)");
    cudaStream_t s = as_stream(stream);
    YuvSource ys;
    ys.fmt = v_first ? kFmtVU : kFmtUV; ys.w = w; ys.h = h; ys.y_pitch = w; ys.c_pitch = w;
    ys.frame_stride = (size_t)w * h * 3 / 2; ys.c_off = (size_t)w * h; ys.c2_off = 0;
    const Canvas whole = {w_out, h_out, 0, 0};
    if (int rc = try_launch_pipe(src, dst, VACV_FP32, batch, ys, w_out, h_out, whole, mean, stddev, s)) {
        if (rc < 0) return rc;
        return check_launch("nv_resize_normalize_chw (persistent)");
    }
    return launch_fused_tiled("nv_resize_normalize_chw", src, dst, VACV_FP32, batch, ys, w_out, h_out, whole, mean, stddev, s);
}

extern "C" int vacv_cuda_resize_normalize(const uint8_t* src, float* dst, int batch, int w, int h, int c,
                                          int w_out, int h_out, const float* mean, const float* stddev,
                                          int out_layout, void* stream) {
    VACV_REQUIRE(src && dst && mean && stddev, "resize_normalize: null pointer");
    VACV_REQUIRE(batch > 0 && batch <= 65535 && w >= 2 && h >= 2 && w_out > 0 && h_out > 0, "resize_normalize: bad size");
    if (c != 1 && c != 3) return set_error(VACV_ERR_UNSUPPORTED, "resize_normalize: c must be 1 or 3 (got %d)", c);
    if (w_out == w && h_out == h)
        return set_error(VACV_ERR_UNSUPPORTED, "resize_normalize: same-size resize (use vacv_cuda_normalize)");
    const int rows_per_cta = max(1, min(h_out, (8192 + w_out - 1) / w_out));
    dim3 grid(ceil_div(h_out, rows_per_cta), batch);
    cudaStream_t s = as_stream(stream);
    if (c == 3 && !knob(kKnobRnGather)) {   // persistent TMA pipeline where the shape qualifies (small or integer ratios, 16-byte rows)
        const int rc = try_launch_resize_pipe_u8c3(src, dst, batch, w, h, w_out, h_out, false, out_layout == VACV_NCHW ? 1 : 2, mean, stddev, s);
        if (rc < 0) return rc;
        if (rc > 0) return check_launch("resize_normalize (persistent)");
    }
    const bool words = c == 3 && (((size_t)w * h * 3) % 4) == 0 && ((uintptr_t)src % 4) == 0 && (size_t)w * h * 3 < 0xfffffff0ull;
    if (words) {   // word-granular taps
        dim3 g3(ceil_div(w_out, 32), ceil_div(h_out, kRnRows), batch), b3(32, 8);
        if (out_layout == VACV_NCHW) resize_normalize_u8c3_kernel<true><<<g3, b3, 0, s>>>(src, dst, w, h, w_out, h_out, (size_t)w * h * 3, mean, stddev);
        else resize_normalize_u8c3_kernel<false><<<g3, b3, 0, s>>>(src, dst, w, h, w_out, h_out, (size_t)w * h * 3, mean, stddev);
    } else if (c == 3) resize_normalize_kernel<3><<<grid, 256, 0, s>>>(src, dst, w, h, w_out, h_out, rows_per_cta, mean, stddev, out_layout);
    else resize_normalize_kernel<1><<<grid, 256, 0, s>>>(src, dst, w, h, w_out, h_out, rows_per_cta, mean, stddev, out_layout);
    return check_launch("resize_normalize");
}

// Next-row extensions of the fused pipeline: decoder-style surfaces (row pitch, planar I420 / YV12
// chroma) in; fp32, fp16 or bf16 CHW planes out; optional letterbox placement.  Same arithmetic as
// vacv_cuda_nv_resize_normalize_chw; 16-bit outputs = the exact fp32 result rounded to nearest-even.  These run on the
// persistent TMA pipeline only (pitches / strides must be multiples of 16).
static int parse_yuv_layout(const char* who, const vacv_yuv_layout* layout, YuvSource& ys) {
    const int w = layout->w, h = layout->h;
    VACV_REQUIRE(w >= 2 && h >= 2 && (w % 2) == 0 && (h % 2) == 0, "%s: w and h must be even and >= 2 (got %dx%d)", who, w, h);
    const bool planar = layout->format == VACV_YUV_I420 || layout->format == VACV_YUV_YV12;
    if (!planar && layout->format != VACV_YUV_NV12 && layout->format != VACV_YUV_NV21)
        return set_error(VACV_ERR_UNSUPPORTED, "%s: format %d", who, layout->format);
    ys.w = w; ys.h = h;
    ys.y_pitch = layout->y_pitch ? layout->y_pitch : w;
    ys.c_pitch = layout->c_pitch ? layout->c_pitch : (planar ? w / 2 : w);
    VACV_REQUIRE(ys.y_pitch >= w && ys.c_pitch >= (planar ? w / 2 : w), "%s: pitch smaller than the row", who);
    const size_t y_bytes = (size_t)ys.y_pitch * h, c_bytes = (size_t)ys.c_pitch * (h / 2);
    ys.frame_stride = layout->frame_stride ? layout->frame_stride : y_bytes + (planar ? 2 * c_bytes : c_bytes);
    VACV_REQUIRE(ys.frame_stride >= y_bytes + (planar ? 2 * c_bytes : c_bytes), "%s: frame_stride too small", who);
    if (planar) {
        ys.fmt = kFmtPlanar;
        const size_t first = y_bytes, second = y_bytes + c_bytes;
        ys.c_off = layout->format == VACV_YUV_I420 ? first : second;    // U plane
        ys.c2_off = layout->format == VACV_YUV_I420 ? second : first;   // V plane
    } else {
        ys.fmt = layout->format == VACV_YUV_NV21 ? kFmtVU : kFmtUV;
        ys.c_off = y_bytes; ys.c2_off = 0;
    }
    return VACV_OK;
}

extern "C" int vacv_cuda_yuv_resize_normalize_chw(const uint8_t* src, const vacv_yuv_layout* layout, void* dst, int out_dtype,
                                                  int batch, int w_out, int h_out, const float* mean, const float* stddev, void* stream) {
    VACV_REQUIRE(src && layout && dst && mean && stddev, "yuv_resize_normalize_chw: null pointer");
    VACV_REQUIRE(batch > 0 && w_out > 0 && h_out > 0, "yuv_resize_normalize_chw: bad size");
    if (out_dtype != VACV_FP32 && out_dtype != VACV_FP16 && out_dtype != VACV_BF16)
        return set_error(VACV_ERR_UNSUPPORTED, "yuv_resize_normalize_chw: out dtype %d (FP32, FP16 or BF16)", out_dtype);
    YuvSource ys;
    if (int rc = parse_yuv_layout("yuv_resize_normalize_chw", layout, ys)) return rc;
    if (w_out == ys.w && h_out == ys.h) return set_error(VACV_ERR_UNSUPPORTED, "yuv_resize_normalize_chw: same-size resize");
    const Canvas whole = {w_out, h_out, 0, 0};
    const int rc = try_launch_pipe(src, dst, out_dtype, batch, ys, w_out, h_out, whole, mean, stddev, as_stream(stream));
    if (rc < 0) return rc;
    if (rc == 0) return launch_fused_tiled("yuv_resize_normalize_chw", src, dst, out_dtype, batch, ys, w_out, h_out, whole, mean, stddev, as_stream(stream));
    return check_launch("yuv_resize_normalize_chw");
}

// Aspect-preserving placement of a w x h frame on a canvas (the usual detector "letterbox"): scale = min(cw/w, ch/h),
// content size rounded to nearest, centred.
extern "C" void vacv_letterbox_rect(int w, int h, int canvas_w, int canvas_h, vacv_rect* content) {
    const double sc = std::min((double)canvas_w / w, (double)canvas_h / h);
    int cw = (int)std::floor(w * sc + 0.5), ch = (int)std::floor(h * sc + 0.5);
    cw = std::max(1, std::min(cw, canvas_w)); ch = std::max(1, std::min(ch, canvas_h));
    content->x = (canvas_w - cw) / 2; content->y = (canvas_h - ch) / 2; content->w = cw; content->h = ch;
}

template <typename T>
static void launch_pad(void* dst, int batch, int cw_, int ch_, const vacv_rect& r, const float pad[3], cudaStream_t s, T (*conv)(float)) {
    const long long total = (long long)(ch_ - r.h) * cw_ + (long long)r.h * (cw_ - r.w);   // border elements per plane
    const int chunks = (int)std::max<long long>(1, std::min<long long>((total + 256 * 8 - 1) / (256 * 8), 64));
    for (int b0 = 0; b0 < batch; b0 += 20000) {
        dim3 grid(chunks, 3 * std::min(batch - b0, 20000));
        letterbox_pad_kernel<T><<<grid, 256, 0, s>>>((T*)dst + (size_t)b0 * 3 * cw_ * ch_, cw_, ch_, r.x, r.y, r.w, r.h, conv(pad[0]), conv(pad[1]), conv(pad[2]));
    }
}
static float conv_f32(float v) { return v; }
static unsigned short conv_f16(float v) { return __half_as_ushort(__float2half_rn(v)); }
static unsigned short conv_bf16(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

// Letterbox: frame -> bilinear resize (reference rule) to content->w x content->h -> placed at (content->x, content->y) of a
// canvas_w x canvas_h canvas filled with pad_bgr -> normalise -> CHW planes.  Equals the unfused chain on the padded u8 canvas.
extern "C" int vacv_cuda_yuv_letterbox_normalize_chw(const uint8_t* src, const vacv_yuv_layout* layout, void* dst, int out_dtype, int batch,
                                                     int canvas_w, int canvas_h, const vacv_rect* content, const uint8_t* pad_bgr,
                                                     const float* mean, const float* stddev, const float* mean_host, const float* stddev_host,
                                                     void* stream) {
    VACV_REQUIRE(src && layout && dst && content && pad_bgr && mean && stddev && mean_host && stddev_host, "yuv_letterbox_normalize_chw: null pointer");
    VACV_REQUIRE(batch > 0 && canvas_w > 0 && canvas_h > 0, "yuv_letterbox_normalize_chw: bad size");
    VACV_REQUIRE(content->w > 0 && content->h > 0 && content->x >= 0 && content->y >= 0 && content->x + content->w <= canvas_w &&
                 content->y + content->h <= canvas_h, "yuv_letterbox_normalize_chw: content rectangle outside the canvas");
    if (out_dtype != VACV_FP32 && out_dtype != VACV_FP16 && out_dtype != VACV_BF16)
        return set_error(VACV_ERR_UNSUPPORTED, "yuv_letterbox_normalize_chw: out dtype %d (FP32, FP16 or BF16)", out_dtype);
    YuvSource ys;
    if (int rc = parse_yuv_layout("yuv_letterbox_normalize_chw", layout, ys)) return rc;
    if (content->w == ys.w && content->h == ys.h) return set_error(VACV_ERR_UNSUPPORTED, "yuv_letterbox_normalize_chw: same-size resize");
    cudaStream_t s = as_stream(stream);
    // border first (it never overlaps the content rectangle); the pad colour goes through the same exact expression
    float pad[3];
    for (int k = 0; k < 3; ++k) pad[k] = (float)((double)((float)pad_bgr[k] - mean_host[k]) / ((double)stddev_host[k] + 1e-6));   // normalize_naive.cpp:74-90
    if (content->w != canvas_w || content->h != canvas_h) {
        if (out_dtype == VACV_FP32) launch_pad<float>(dst, batch, canvas_w, canvas_h, *content, pad, s, conv_f32);
        else launch_pad<unsigned short>(dst, batch, canvas_w, canvas_h, *content, pad, s, out_dtype == VACV_FP16 ? conv_f16 : conv_bf16);
    }
    const Canvas cv = {canvas_w, canvas_h, content->x, content->y};
    const int rc = try_launch_pipe(src, dst, out_dtype, batch, ys, content->w, content->h, cv, mean, stddev, s);
    if (rc < 0) return rc;
    if (rc == 0) return launch_fused_tiled("yuv_letterbox_normalize_chw", src, dst, out_dtype, batch, ys, content->w, content->h, cv, mean, stddev, s);
    return check_launch("yuv_letterbox_normalize_chw");
}
