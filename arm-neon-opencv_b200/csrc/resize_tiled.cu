// Tiled resize kernels (a5-a9): the source window of an output tile is staged ONCE in shared memory with coalesced
// 128-bit loads, every tap is then a shared-memory read, and the output tile leaves through shared memory as
// lane-contiguous 128-bit stores.  Replaces the byte-granular global gathers of the direct kernels in resize.cu
// (kept as fallback), which were LSU-bound (ncu: c4 u8 cubic at 7 % of the HBM roofline).
//
//   tile          TH output rows x up to 256 output ELEMENTS (pixels x channels) of one image; thread t owns element
//                 column t of the tile and walks down the rows, so its x taps/coefficients live in registers.
//   source rows   only the rows some tap of the tile touches are staged (row -> slot table), so large vertical
//                 down-scales do not fetch unused rows.
//   bilinear      (u8 naive rule / NEON rule / signed-char compat, fp32): the reference's 4-tap expression evaluated
//                 directly on the staged taps (bit-exact operation order).
//   bicubic       separable like the reference (resize_naive.cpp:187-366) and like OpenCV 2.4: pass 1 forms the
//                 horizontal sums H[slot][element] in shared memory (each source row once per tile instead of once
//                 per output row: 2.5x fewer MACs at 4:3), pass 2 combines 4 slots vertically.
//                 u8: H is an exact integer carried as fp32 (|H| < 2^22), vertical pass = OpenCV's fp32 SSE2 body
//                 with round-half-even, or its integer tail for the last <= 7 elements of an output row.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include "resize_coeffs.cuh"
#include "resize_cubic3.cuh"
#include "resize_cubic3_walk.cuh"
#include "resize_cubic3_walkn.cuh"
#include "resize_cubic3_period.cuh"
#include "resize_cubic_f32_period.cuh"
#include "host_util.cuh"
#include "vacv_common.cuh"

namespace vacv {

enum { kLinU8 = 0, kLinU8Signed = 1, kLinU8Neon = 2, kLinF32 = 3, kCubF32 = 4, kCubU8 = 5 };

constexpr int kRtThreads = 256;
constexpr int kRtMaxTH = 16;
constexpr int kRtMaxBand = 96;   // source rows spanned by one tile (host picks TH accordingly)

struct TiledGeom {
    int w, h, c, wo, ho;
    int TWE, TH;              // tile: elements (multiple of c, <= 256) x rows
    int hpitch;               // floats per row of the horizontal-sum buffer (TWE rounded up to 4)
    int tiles_x, tiles_y;
    int src_pitch;            // shared-memory bytes per staged source row (multiple of 16)
    int max_slots;            // staged row capacity
    int opitch;               // shared-memory bytes per output row (multiple of 16, >= TWE*es + 16)
    int align;                // staging granularity in bytes: 16, 4 or 1
    int tiles_per_cta;        // consecutive y tiles walked by one CTA (x coefficients are computed once)
    int dense;                // 1: vertical scale <= taps, every row of a tile's band is touched -> slot = row - first row
    double scale_x, scale_y;  // coordinate scales, formed on the host with the reference's expression for the kind
    size_t src_image, dst_image;   // elements between images
};

template <int KIND> struct Kind {
    using S = uint8_t;
    static constexpr int TAPS = 2;
};
template <> struct Kind<kLinF32> { using S = float; static constexpr int TAPS = 2; };
template <> struct Kind<kCubF32> { using S = float; static constexpr int TAPS = 4; };
template <> struct Kind<kCubU8> { using S = uint8_t; static constexpr int TAPS = 4; };

// Tap indices (absolute, along one axis) and coefficients (int, or float bits) of output coordinate d.
template <int KIND>
__device__ __forceinline__ void axis_coefs(int d, int n_in, double scale, bool is_x, int (&idx)[Kind<KIND>::TAPS],
                                           int (&coef)[Kind<KIND>::TAPS]) {
    if constexpr (KIND == kLinU8 || KIND == kLinU8Signed || KIND == kLinU8Neon || KIND == kLinF32) {
        int s; float f;
        linear_coord(d, scale, n_in, s, f);
        idx[0] = s; idx[1] = s + 1;
        if constexpr (KIND == kLinF32) { coef[0] = __float_as_int(1.f - f); coef[1] = __float_as_int(f); }
        else { coef[0] = sat_short((1.f - f) * 2048.f); coef[1] = sat_short(f * 2048.f); }
    } else if constexpr (KIND == kCubF32) {
        int ofs; float a[4];
        cubic_naive_scaled(d, n_in, scale, ofs, a);
#pragma unroll
        for (int j = 0; j < 4; ++j) { idx[j] = ofs - 1 + j; coef[j] = __float_as_int(a[j]); }
    } else {
        int s, q[4];
        cubic_cv_coord_scaled(d, n_in, scale, is_x, s, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) { idx[j] = min(max(s - 1 + j, 0), n_in - 1); coef[j] = q[j]; }
    }
}

__device__ __forceinline__ float int_to_float_exact(int v) {   // |v| < 2^22: two ALU ops instead of an XU conversion
    return __int_as_float(0x4B400000 + v) - 12582912.0f;
}
__device__ __forceinline__ int float_to_int_rhe(float f) {     // |f| < 2^22, round half to even (cvtps2dq)
    return __float_as_int(f + 12582912.0f) - 0x4B400000;
}

template <int KIND>
__global__ void __launch_bounds__(kRtThreads) resize_tiled_kernel(const void* __restrict__ src_, void* __restrict__ dst_, TiledGeom g) {
    using S = typename Kind<KIND>::S;
    constexpr int K = Kind<KIND>::TAPS;
    constexpr int ES = sizeof(S);
    constexpr bool kTwoPass = K == 4;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_yidx[kRtMaxTH][K];
    __shared__ __align__(16) int s_ycoef[kRtMaxTH][4];    // int, or float bits (fp32 kinds; u8 bicubic: (float)ibeta * 2^-22)
    __shared__ __align__(16) int s_yint[kRtMaxTH][4];     // u8 bicubic: integer ibeta for OpenCV's scalar tail
    __shared__ __align__(16) int s_yoff[kRtMaxTH][4];     // byte offset of each tap's row inside tile (1-pass) / hbuf (2-pass)
    __shared__ int s_ooff[kRtMaxTH];                      // byte offset of the row inside obuf (includes its 16-byte phase)
    __shared__ short s_slot[kRtMaxBand];
    __shared__ int s_rows[kRtMaxBand];
    __shared__ int s_x[2], s_nslots;

    uint8_t* tile = smem;                                                                 // [max_slots][src_pitch]
    uint8_t* hbuf = smem + (size_t)g.max_slots * g.src_pitch;                             // [max_slots][hpitch] floats (two-pass)
    uint8_t* obuf = hbuf + (kTwoPass ? (size_t)g.max_slots * g.hpitch * 4 : 0);          // [TH][opitch]

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % g.tiles_x, ychunk = blockIdx.x / g.tiles_x;
    const int e0 = tile_x * g.TWE;
    const int twe = min(g.TWE, g.wo * g.c - e0);
    const uint8_t* img = reinterpret_cast<const uint8_t*>(src_) + blockIdx.y * g.src_image * ES;
    uint8_t* out_img = reinterpret_cast<uint8_t*>(dst_) + blockIdx.y * g.dst_image * ES;
    const int row_bytes = g.w * g.c * ES;
    const size_t out_row_bytes = (size_t)g.wo * g.c * ES;
    const int seg = twe * ES;

    // ---- once per CTA: this thread's element column -> x taps / coefficients in registers, staged column range
    const bool active = tid < twe;
    const int xe = active ? tid : twe - 1;
    const int dxl = xe / g.c, ch = xe - dxl * g.c;
    int xidx[K], xcoef[K];
    axis_coefs<KIND>(e0 / g.c + dxl, g.w, g.scale_x, true, xidx, xcoef);
    if (tid == 0) s_x[0] = xidx[0];
    if (tid == twe - 1) s_x[1] = xidx[K - 1];
    __syncthreads();
    const int xb0 = s_x[0] * g.c * ES, xb1 = (s_x[1] + 1) * g.c * ES;
    const int xb0a = xb0 & ~(g.align - 1);
    const int width = ((xb1 + g.align - 1) & ~(g.align - 1)) - xb0a;
    int toff[K];   // byte offset of each x tap of this thread's element inside a staged row
#pragma unroll
    for (int j = 0; j < K; ++j) toff[j] = (xidx[j] * g.c + ch) * ES - xb0a;
    const int vec_end = (g.wo * g.c) & ~7;   // OpenCV's SSE2 body covers x < (width & ~7)
    const bool in_vec_body = e0 + tid < vec_end;

    for (int t = 0; t < g.tiles_per_cta; ++t) {
        const int tile_y = ychunk * g.tiles_per_cta + t;
        if (tile_y >= g.tiles_y) break;
        const int dy0 = tile_y * g.TH, th = min(g.TH, g.ho - dy0);

        // ---- row taps / coefficients of the tile
        if (tid < th) {
            int yi[K], yc[K];
            axis_coefs<KIND>(dy0 + tid, g.h, g.scale_y, false, yi, yc);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                s_yidx[tid][j] = yi[j];
                if constexpr (KIND == kCubU8) { s_ycoef[tid][j] = __float_as_int((float)yc[j] * (1.f / (2048 * 2048))); s_yint[tid][j] = yc[j]; }
                else s_ycoef[tid][j] = yc[j];
            }
            const uintptr_t ga = reinterpret_cast<uintptr_t>(out_img) + (size_t)(dy0 + tid) * out_row_bytes + (size_t)e0 * ES;
            s_ooff[tid] = tid * g.opitch + (int)(ga & 15);
        }
        __syncthreads();

        // ---- which source rows does the tile touch?  row -> slot
        const int y_lo = s_yidx[0][0], band = s_yidx[th - 1][K - 1] - y_lo + 1;
        int nslots;
        if (g.dense) {   // consecutive tap windows leave no gaps: the band IS the slot list
            nslots = band;
            for (int i = tid; i < th * K; i += kRtThreads)
                s_yoff[i / K][i % K] = (s_yidx[i / K][i % K] - y_lo) * (kTwoPass ? g.hpitch * 4 : g.src_pitch);
        } else {         // large vertical down-scale: stage only the rows some tap touches
            for (int r = tid; r < band; r += kRtThreads) s_slot[r] = -1;
            __syncthreads();
            for (int i = tid; i < th * K; i += kRtThreads) s_slot[s_yidx[i / K][i % K] - y_lo] = 0;
            __syncthreads();
            if (tid < 32) {   // warp 0: slot = number of touched rows below (ballot + popc prefix)
                int base = 0;
                for (int r0 = 0; r0 < band; r0 += 32) {
                    const int r = r0 + tid;
                    const bool used = r < band && s_slot[r] == 0;
                    const unsigned m = __ballot_sync(0xffffffffu, used);
                    if (used) {
                        const int slot = base + __popc(m & ((1u << tid) - 1));
                        s_slot[r] = (short)slot;
                        s_rows[slot] = y_lo + r;
                    }
                    base += __popc(m);
                }
                if (tid == 0) s_nslots = base;
            }
            __syncthreads();
            nslots = s_nslots;
            for (int i = tid; i < th * K; i += kRtThreads)
                s_yoff[i / K][i % K] = s_slot[s_yidx[i / K][i % K] - y_lo] * (kTwoPass ? g.hpitch * 4 : g.src_pitch);
        }
        const int row0 = g.dense ? y_lo : 0;   // dense: staged row r is source row y_lo + r

        // ---- stage the touched rows, columns [x_lo, x_hi], at the widest legal granularity
        if (g.align == 16) {
            const int units = width >> 4;
            for (int i = tid; i < nslots * units; i += kRtThreads) {
                const int r = i / units, u = i - r * units;
                *reinterpret_cast<uint4*>(tile + r * g.src_pitch + 16 * u) = ld_stream16(img + (size_t)(g.dense ? row0 + r : s_rows[r]) * row_bytes + xb0a + 16 * u);
            }
        } else if (g.align == 4) {
            const int units = width >> 2;
            for (int i = tid; i < nslots * units; i += kRtThreads) {
                const int r = i / units, u = i - r * units;
                *reinterpret_cast<uint32_t*>(tile + r * g.src_pitch + 4 * u) = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)(g.dense ? row0 + r : s_rows[r]) * row_bytes + xb0a) + u);
            }
        } else {
            for (int i = tid; i < nslots * width; i += kRtThreads) {
                const int r = i / width, u = i - r * width;
                tile[r * g.src_pitch + u] = __ldg(img + (size_t)(g.dense ? row0 + r : s_rows[r]) * row_bytes + xb0a + u);
            }
        }
        __syncthreads();

        // ---- pass 1 (bicubic): horizontal sums per staged row
        if constexpr (kTwoPass) {
            if (active) {
                const uint8_t* r0 = tile + toff[0];
                const uint8_t* r1 = tile + toff[1];
                const uint8_t* r2 = tile + toff[2];
                const uint8_t* r3 = tile + toff[3];
                float* hp = reinterpret_cast<float*>(hbuf) + tid;
#pragma unroll 4
                for (int s = 0; s < nslots; ++s) {
                    float hval;
                    if constexpr (KIND == kCubF32) {   // resize_naive.cpp:230: S[-1]*a0 + S[0]*a1 + S[1]*a2 + S[2]*a3, left to right
                        hval = *reinterpret_cast<const float*>(r0) * __int_as_float(xcoef[0]) + *reinterpret_cast<const float*>(r1) * __int_as_float(xcoef[1]) +
                               *reinterpret_cast<const float*>(r2) * __int_as_float(xcoef[2]) + *reinterpret_cast<const float*>(r3) * __int_as_float(xcoef[3]);
                    } else {
                        hval = int_to_float_exact(*r0 * xcoef[0] + *r1 * xcoef[1] + *r2 * xcoef[2] + *r3 * xcoef[3]);
                    }
                    *hp = hval;
                    r0 += g.src_pitch; r1 += g.src_pitch; r2 += g.src_pitch; r3 += g.src_pitch;
                    hp += g.hpitch;
                }
            }
            __syncthreads();
        }

        // ---- pass 2: one output row at a time into the output tile (same 16-byte phase as its global destination)
        if (active) {
            for (int ty = 0; ty < th; ++ty) {
                const int4 yo = *reinterpret_cast<const int4*>(s_yoff[ty]);
                const int4 yc = *reinterpret_cast<const int4*>(s_ycoef[ty]);
                S* o = reinterpret_cast<S*>(obuf + s_ooff[ty]) + tid;
                if constexpr (KIND == kLinU8 || KIND == kLinU8Signed || KIND == kLinU8Neon) {
                    constexpr bool kS = KIND == kLinU8Signed;
                    const uint8_t* r0 = tile + yo.x;
                    const uint8_t* r1 = tile + yo.y;
                    const int p00 = pix<kS>(r0[toff[0]]), p01 = pix<kS>(r0[toff[1]]), p10 = pix<kS>(r1[toff[0]]), p11 = pix<kS>(r1[toff[1]]);
                    const int cx0 = xcoef[0], cx1 = xcoef[1], cy0 = yc.x, cy1 = yc.y;
                    int v;
                    if constexpr (KIND == kLinU8Neon) {   // resize_neon.cpp:145-181
                        const int h0 = (short)((p00 * cx0 + p01 * cx1) >> 4), h1 = (short)((p10 * cx0 + p11 * cx1) >> 4);
                        v = clamp255(((short)((cy0 * h0) >> 16) + (short)((cy1 * h1) >> 16) + 2) >> 2);
                    } else {                              // resize_naive.cpp:60-65
                        v = (p00 * cx0 * cy0 + p10 * cx0 * cy1 + p01 * cx1 * cy0 + p11 * cx1 * cy1) >> 22;
                    }
                    *o = (uint8_t)v;
                } else if constexpr (KIND == kLinF32) {   // resize_naive.cpp:121-124 evaluation order
                    const uint8_t* r0 = tile + yo.x;
                    const uint8_t* r1 = tile + yo.y;
                    const float lt = *reinterpret_cast<const float*>(r0 + toff[0]), rt = *reinterpret_cast<const float*>(r0 + toff[1]);
                    const float lb = *reinterpret_cast<const float*>(r1 + toff[0]), rb = *reinterpret_cast<const float*>(r1 + toff[1]);
                    const float cx0 = __int_as_float(xcoef[0]), cx1 = __int_as_float(xcoef[1]);
                    const float cy0 = __int_as_float(yc.x), cy1 = __int_as_float(yc.y);
                    *o = lt * cx0 * cy0 + lb * cx0 * cy1 + rt * cx1 * cy0 + rb * cx1 * cy1;
                } else {
                    const uint8_t* hb = hbuf + 4 * tid;
                    const float h0 = *reinterpret_cast<const float*>(hb + yo.x), h1 = *reinterpret_cast<const float*>(hb + yo.y);
                    const float h2 = *reinterpret_cast<const float*>(hb + yo.z), h3 = *reinterpret_cast<const float*>(hb + yo.w);
                    const float b0 = __int_as_float(yc.x), b1 = __int_as_float(yc.y), b2 = __int_as_float(yc.z), b3 = __int_as_float(yc.w);
                    if constexpr (KIND == kCubF32) {      // resize_naive.cpp:345
                        *o = h0 * b0 + h1 * b1 + h2 * b2 + h3 * b3;
                    } else {
                        int v;
                        if (in_vec_body) {                // fp32 body: mul, then add, one rounding each; cvtps2dq; packs; packus
                            float f = h0 * b0;
                            f = f + h1 * b1;
                            f = f + h2 * b2;
                            f = f + h3 * b3;
                            v = max(min(float_to_int_rhe(f), 32767), -32768);
                        } else {                          // scalar tail: FixedPtCast<int, uchar, 22>
                            const int4 ib = *reinterpret_cast<const int4*>(s_yint[ty]);
                            v = (__float2int_rn(h0) * ib.x + __float2int_rn(h1) * ib.y + __float2int_rn(h2) * ib.z + __float2int_rn(h3) * ib.w + (1 << 21)) >> 22;
                        }
                        *o = (uint8_t)clamp255(v);
                    }
                }
            }
        }
        __syncthreads();

        // ---- copy the output tile out: 16-byte aligned chunks, partial chunks at the row ends byte by byte
        const int chunks_per_row = (seg + 15 + 15) >> 4;   // upper bound incl. phase
        for (int i = tid; i < th * chunks_per_row; i += kRtThreads) {
            const int ty = i / chunks_per_row, q = i - ty * chunks_per_row;
            const int mis = s_ooff[ty] - ty * g.opitch;
            const int lo = max(mis, 16 * q), hi = min(mis + seg, 16 * q + 16);
            if (lo >= hi) continue;
            const uint8_t* sp = obuf + ty * g.opitch + 16 * q;
            uint8_t* gp = out_img + (size_t)(dy0 + ty) * out_row_bytes + (size_t)e0 * ES - mis + 16 * q;
            if (hi - lo == 16) st_stream16(gp, *reinterpret_cast<const uint4*>(sp));
            else for (int b = lo - 16 * q; b < hi - 16 * q; ++b) gp[b] = sp[b];
        }
        __syncthreads();   // tile / hbuf / obuf / row tables are reused by the next y tile
    }
}

// =====================================================================================================
// Bicubic, 3-channel interleaved (the C4 shape): PIXEL-per-thread variant.  The element-per-thread kernel above is
// bound by shared-memory instructions (ncu: LSU wavefronts 84 % of peak, profiles/r1_cubic_u8_tiled_v2_ncu_raw.txt);
// here a thread owns all 3 channels of one output column, so
//   pass 1 reads the 12 contiguous tap bytes of a pixel as four aligned words + funnel shifts (4 loads instead of 12)
//          and writes one float4 (H_b, H_g, H_r, -) per staged row;
//   pass 2 reads four float4 per output pixel (4 loads instead of 12) and the per-row tables once per pixel.
constexpr int kC3Px = 128;        // tile width in pixels
constexpr int kC3Threads = 256;   // two threads per pixel column: they split the staged rows (pass 1) and the output rows (pass 2)

struct Cubic3Geom {
    int w, h, wo, ho;
    int TH, tiles_x, tiles_y, tiles_per_cta;
    int src_pitch, max_slots, opitch, align;
    double scale_x, scale_y;
    size_t src_image, dst_image;   // elements between images
};

template <int KIND>
__global__ void __launch_bounds__(kC3Threads) resize_cubic3_kernel(const void* __restrict__ src_, void* __restrict__ dst_, Cubic3Geom g) {
    using S = typename Kind<KIND>::S;
    constexpr int ES = sizeof(S), C = 3, PX = C * ES;   // bytes per pixel
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_yidx[kRtMaxTH][4];
    __shared__ __align__(16) int s_ycoef[kRtMaxTH][4];   // float bits: naive beta, or (float)ibeta * 2^-22
    __shared__ __align__(16) int s_yint[kRtMaxTH][4];    // u8: integer ibeta for OpenCV's scalar tail
    __shared__ __align__(16) int s_yoff[kRtMaxTH][4];    // byte offset of each tap's row inside hbuf
    __shared__ int s_ooff[kRtMaxTH];
    __shared__ int s_x[2];

    uint8_t* tile = smem;                                             // [max_slots][src_pitch]
    uint8_t* hbuf = smem + (size_t)g.max_slots * g.src_pitch;         // [max_slots][128] float4
    uint8_t* obuf = hbuf + (size_t)g.max_slots * kC3Px * 16;     // [TH][opitch]
    constexpr int kHRow = kC3Px * 16;

    const int tid = threadIdx.x, px = tid & (kC3Px - 1), half = tid >> 7;
    const int tile_x = blockIdx.x % g.tiles_x, ychunk = blockIdx.x / g.tiles_x;
    const int dx0 = tile_x * kC3Px;
    const int tw = min(kC3Px, g.wo - dx0);
    const uint8_t* img = reinterpret_cast<const uint8_t*>(src_) + blockIdx.y * g.src_image * ES;
    uint8_t* out_img = reinterpret_cast<uint8_t*>(dst_) + blockIdx.y * g.dst_image * ES;
    const int row_bytes = g.w * PX;
    const size_t out_row_bytes = (size_t)g.wo * PX;
    const int seg = tw * PX;

    // ---- once per CTA: x taps of this thread's column
    const bool active = px < tw;
    int xidx[4], xcoef[4];
    axis_coefs<KIND>(dx0 + (active ? px : tw - 1), g.w, g.scale_x, true, xidx, xcoef);
    if (tid == 0) s_x[0] = xidx[0];
    if (tid == tw - 1) s_x[1] = xidx[3];
    __syncthreads();
    const int xb0 = s_x[0] * PX, xb1 = (s_x[1] + 1) * PX;
    const int xb0a = xb0 & ~(g.align - 1);
    const int width = ((xb1 + g.align - 1) & ~(g.align - 1)) - xb0a;
    int toff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) toff[j] = xidx[j] * PX - xb0a;
    const bool consecutive = toff[1] == toff[0] + PX && toff[2] == toff[1] + PX && toff[3] == toff[2] + PX;   // false only where taps are clamped
    const int vec_end = (g.wo * C) & ~7;
    const int e_first = (dx0 + px) * C;   // element index of this pixel's first channel in the output row

    for (int t = 0; t < g.tiles_per_cta; ++t) {
        const int tile_y = ychunk * g.tiles_per_cta + t;
        if (tile_y >= g.tiles_y) break;
        const int dy0 = tile_y * g.TH, th = min(g.TH, g.ho - dy0);
        if (tid < th) {
            int yi[4], yc[4];
            axis_coefs<KIND>(dy0 + tid, g.h, g.scale_y, false, yi, yc);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_yidx[tid][j] = yi[j];
                if constexpr (KIND == kCubU8) { s_ycoef[tid][j] = __float_as_int((float)yc[j] * (1.f / (2048 * 2048))); s_yint[tid][j] = yc[j]; }
                else s_ycoef[tid][j] = yc[j];
            }
            const uintptr_t ga = reinterpret_cast<uintptr_t>(out_img) + (size_t)(dy0 + tid) * out_row_bytes + (size_t)dx0 * PX;
            s_ooff[tid] = tid * g.opitch + (int)(ga & 15);
        }
        __syncthreads();
        // vertical scale <= 4 on this path (host check): every row of the band is touched -> slot = row - first row
        const int y_lo = s_yidx[0][0], nslots = s_yidx[th - 1][3] - y_lo + 1;
        for (int i = tid; i < th * 4; i += kC3Threads) s_yoff[i >> 2][i & 3] = (s_yidx[i >> 2][i & 3] - y_lo) * kHRow;

        // ---- stage rows [y_lo, y_lo + nslots), byte columns [xb0a, xb0a + width)
        if (g.align == 16) {
            const int units = width >> 4;
            for (int i = tid; i < nslots * units; i += kC3Threads) {
                const int r = i / units, u = i - r * units;
                *reinterpret_cast<uint4*>(tile + r * g.src_pitch + 16 * u) = ld_stream16(img + (size_t)(y_lo + r) * row_bytes + xb0a + 16 * u);
            }
        } else if (g.align == 4) {
            const int units = width >> 2;
            for (int i = tid; i < nslots * units; i += kC3Threads) {
                const int r = i / units, u = i - r * units;
                *reinterpret_cast<uint32_t*>(tile + r * g.src_pitch + 4 * u) = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)(y_lo + r) * row_bytes + xb0a) + u);
            }
        } else {
            for (int i = tid; i < nslots * width; i += kC3Threads) {
                const int r = i / width, u = i - r * width;
                tile[r * g.src_pitch + u] = __ldg(img + (size_t)(y_lo + r) * row_bytes + xb0a + u);
            }
        }
        __syncthreads();

        // ---- pass 1: horizontal sums of the 3 channels per staged row -> float4
        if (active) {
            float4* hp = reinterpret_cast<float4*>(hbuf) + half * kC3Px + px;
            if constexpr (KIND == kCubU8) {
                if (consecutive) {   // 12 contiguous bytes: four aligned words, funnel-shifted
                    const uint8_t* base = tile + half * g.src_pitch + (toff[0] & ~3);
                    const int sh = (toff[0] & 3) * 8;
#pragma unroll 2
                    for (int s = half; s < nslots; s += 2, base += 2 * g.src_pitch, hp += 2 * kC3Px) {
                        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(base), w1 = *reinterpret_cast<const uint32_t*>(base + 4);
                        const uint32_t w2 = *reinterpret_cast<const uint32_t*>(base + 8), w3 = *reinterpret_cast<const uint32_t*>(base + 12);
                        const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
                        // bytes: b0 = [t0.b t0.g t0.r t1.b]  b1 = [t1.g t1.r t2.b t2.g]  b2 = [t2.r t3.b t3.g t3.r]
                        const int hb = (int)(b0 & 0xff) * xcoef[0] + (int)(b0 >> 24) * xcoef[1] + (int)((b1 >> 16) & 0xff) * xcoef[2] + (int)((b2 >> 8) & 0xff) * xcoef[3];
                        const int hg = (int)((b0 >> 8) & 0xff) * xcoef[0] + (int)(b1 & 0xff) * xcoef[1] + (int)(b1 >> 24) * xcoef[2] + (int)((b2 >> 16) & 0xff) * xcoef[3];
                        const int hr = (int)((b0 >> 16) & 0xff) * xcoef[0] + (int)((b1 >> 8) & 0xff) * xcoef[1] + (int)(b2 & 0xff) * xcoef[2] + (int)(b2 >> 24) * xcoef[3];
                        *hp = make_float4(int_to_float_exact(hb), int_to_float_exact(hg), int_to_float_exact(hr), 0.f);
                    }
                } else {
                    const uint8_t* row = tile + half * g.src_pitch;
                    for (int s = half; s < nslots; s += 2, row += 2 * g.src_pitch, hp += 2 * kC3Px) {
                        int h[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            h[k] = row[toff[0] + k] * xcoef[0] + row[toff[1] + k] * xcoef[1] + row[toff[2] + k] * xcoef[2] + row[toff[3] + k] * xcoef[3];
                        *hp = make_float4(int_to_float_exact(h[0]), int_to_float_exact(h[1]), int_to_float_exact(h[2]), 0.f);
                    }
                }
            } else {
                const uint8_t* row = tile + half * g.src_pitch;
                const float a0 = __int_as_float(xcoef[0]), a1 = __int_as_float(xcoef[1]), a2 = __int_as_float(xcoef[2]), a3 = __int_as_float(xcoef[3]);
                for (int s = half; s < nslots; s += 2, row += 2 * g.src_pitch, hp += 2 * kC3Px) {
                    float h[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {   // resize_naive.cpp:230 order
                        const float t0 = *reinterpret_cast<const float*>(row + toff[0] + 4 * k), t1 = *reinterpret_cast<const float*>(row + toff[1] + 4 * k);
                        const float t2 = *reinterpret_cast<const float*>(row + toff[2] + 4 * k), t3 = *reinterpret_cast<const float*>(row + toff[3] + 4 * k);
                        h[k] = t0 * a0 + t1 * a1 + t2 * a2 + t3 * a3;
                    }
                    *hp = make_float4(h[0], h[1], h[2], 0.f);
                }
            }
        }
        __syncthreads();

        // ---- pass 2: vertical combination, 3 channels per thread
        if (active) {
            const uint8_t* hb = hbuf + 16 * px;
            for (int ty = half; ty < th; ty += 2) {
                const int4 yo = *reinterpret_cast<const int4*>(s_yoff[ty]);
                const int4 yc = *reinterpret_cast<const int4*>(s_ycoef[ty]);
                const float4 h0 = *reinterpret_cast<const float4*>(hb + yo.x), h1 = *reinterpret_cast<const float4*>(hb + yo.y);
                const float4 h2 = *reinterpret_cast<const float4*>(hb + yo.z), h3 = *reinterpret_cast<const float4*>(hb + yo.w);
                const float b0 = __int_as_float(yc.x), b1 = __int_as_float(yc.y), b2 = __int_as_float(yc.z), b3 = __int_as_float(yc.w);
                const float c0[3] = {h0.x, h0.y, h0.z}, c1[3] = {h1.x, h1.y, h1.z}, c2[3] = {h2.x, h2.y, h2.z}, c3[3] = {h3.x, h3.y, h3.z};
                S* o = reinterpret_cast<S*>(obuf + s_ooff[ty]) + C * px;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if constexpr (KIND == kCubF32) {   // resize_naive.cpp:345
                        o[k] = c0[k] * b0 + c1[k] * b1 + c2[k] * b2 + c3[k] * b3;
                    } else {
                        int v;
                        if (e_first + k < vec_end) {   // OpenCV SSE2 body
                            float f = c0[k] * b0;
                            f = f + c1[k] * b1;
                            f = f + c2[k] * b2;
                            f = f + c3[k] * b3;
                            v = max(min(float_to_int_rhe(f), 32767), -32768);
                        } else {                       // scalar tail
                            const int4 ib = *reinterpret_cast<const int4*>(s_yint[ty]);
                            v = (__float2int_rn(c0[k]) * ib.x + __float2int_rn(c1[k]) * ib.y + __float2int_rn(c2[k]) * ib.z + __float2int_rn(c3[k]) * ib.w + (1 << 21)) >> 22;
                        }
                        o[k] = (uint8_t)clamp255(v);
                    }
                }
            }
        }
        __syncthreads();

        // ---- copy out
        const int chunks_per_row = (seg + 15 + 15) >> 4;
        for (int i = tid; i < th * chunks_per_row; i += kC3Threads) {
            const int ty = i / chunks_per_row, q = i - ty * chunks_per_row;
            const int mis = s_ooff[ty] - ty * g.opitch;
            const int lo = max(mis, 16 * q), hi = min(mis + seg, 16 * q + 16);
            if (lo >= hi) continue;
            const uint8_t* sp = obuf + ty * g.opitch + 16 * q;
            uint8_t* gp = out_img + (size_t)(dy0 + ty) * out_row_bytes + (size_t)dx0 * PX - mis + 16 * q;
            if (hi - lo == 16) st_stream16(gp, *reinterpret_cast<const uint4*>(sp));
            else for (int b = lo - 16 * q; b < hi - 16 * q; ++b) gp[b] = sp[b];
        }
        __syncthreads();
    }
}

template <int KIND>
static int launch_cubic3(const void* src, void* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    constexpr int ES = sizeof(typename Kind<KIND>::S), PX = 3 * ES;
    Cubic3Geom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * 3; g.dst_image = (size_t)wo * ho * 3;
    const double sx = (double)w / wo, sy = (double)h / ho;
    if (sy > 3.75 || sx > 6.0) return 0;   // dense band assumption / staged span; larger down-scales use the generic kernel
    if (KIND == kCubU8) { g.scale_x = 1. / ((double)wo / (double)w); g.scale_y = 1. / ((double)ho / (double)h); }
    else { g.scale_x = (double)w / (double)wo; g.scale_y = (double)h / (double)ho; }
    const size_t row_bytes = (size_t)w * PX;
    g.align = ((row_bytes % 16) == 0 && ((uintptr_t)src % 16) == 0) ? 16 : ((row_bytes % 4) == 0 && ((uintptr_t)src % 4) == 0) ? 4 : 1;
    const int span_px = std::min(w, (int)(sx * (kC3Px - 1)) + 4 + 3);
    g.src_pitch = (int)(((size_t)span_px * PX + 2 * g.align + 15 + 16) & ~(size_t)15);   // +16: the funnel-shift loads read one word past the last tap
    g.opitch = (kC3Px * PX + 16 + 15) & ~15;
    const size_t budget = 44 * 1024;   // 5 CTAs x 8 warps per SM
    for (int TH = kRtMaxTH; TH >= 1; --TH) {
        const int slots = (int)(sy * (TH - 1)) + 4 + 3;
        const size_t smem = (size_t)slots * g.src_pitch + (size_t)slots * kC3Px * 16 + (size_t)TH * g.opitch;
        if (smem > budget) continue;
        g.TH = TH; g.max_slots = slots;
        g.tiles_x = (wo + kC3Px - 1) / kC3Px; g.tiles_y = (ho + TH - 1) / TH;
        const long long ctas1 = (long long)g.tiles_x * g.tiles_y * images;
        g.tiles_per_cta = (int)std::max<long long>(1, std::min<long long>(8, ctas1 / (current_sm_count() * 16LL)));
        const int ychunks = (g.tiles_y + g.tiles_per_cta - 1) / g.tiles_per_cta;
        auto kern = resize_cubic3_kernel<KIND>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
        }
        for (int i0 = 0; i0 < images; i0 += 65535) {
            dim3 grid(g.tiles_x * ychunks, std::min(images - i0, 65535));
            kern<<<grid, kC3Threads, smem, s>>>((const uint8_t*)src + (size_t)i0 * g.src_image * ES, (uint8_t*)dst + (size_t)i0 * g.dst_image * ES, g);
        }
        return 1;
    }
    return 0;
}


// Rolling bicubic for interleaved 3-channel images (resize_cubic3.cuh).  1 = launched, 0 = shape not eligible.
template <bool kU8>
static int launch_cubic3_rolling(const void* src, void* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    constexpr int ES = kU8 ? 1 : 4, PX = 3 * ES;
    const double sy = (double)h / ho;
    if (((size_t)w * PX) % 4 != 0 || ((uintptr_t)src % 4) != 0) return 0;       // pass 1 reads aligned 32-bit words
    int G = kRollMaxG;
    while (G > 1 && (int)(sy * (G - 1)) + 6 > kRollRing) --G;                   // a group's source rows must fit the ring
    if ((int)(sy * (G - 1)) + 6 > kRollRing) return 0;
    RollGeom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho; g.G = G;
    g.src_image = (size_t)w * h * 3; g.dst_image = (size_t)wo * ho * 3;
    if (kU8) { g.scale_x = 1. / ((double)wo / (double)w); g.scale_y = 1. / ((double)ho / (double)h); }   // OpenCV 2.4
    else { g.scale_x = (double)w / (double)wo; g.scale_y = (double)h / (double)ho; }                     // resize_naive.cpp:144
    g.strips = (wo + kRollPx - 1) / kRollPx;
    g.opitch = (kRollPx * PX + 16 + 15) & ~15;
    // vertical segments: enough CTAs for >= ~6 waves of 5 CTAs/SM, each segment a multiple of G rows
    const long long want = 6LL * 5 * current_sm_count();
    long long segs = std::max<long long>(1, std::min<long long>((want + (long long)g.strips * images - 1) / ((long long)g.strips * images), (ho + G - 1) / G));
    int rps = (int)((ho + segs - 1) / segs);
    rps = std::min(kRollMaxRows / G * G, (rps + G - 1) / G * G);
    g.rows_per_seg = rps;
    g.segs = (ho + rps - 1) / rps;
    const size_t smem = (size_t)kRollRing * kRollPx * 16 + (size_t)rps * sizeof(RowEntry) + (((size_t)rps * 4 + 15) & ~(size_t)15) + (size_t)G * g.opitch;
    auto kern = resize_cubic3_rolling_kernel<kU8>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    }
    for (int i0 = 0; i0 < images; i0 += 65535) {
        dim3 grid(g.strips * g.segs, std::min(images - i0, 65535));
        kern<<<grid, kRollThreads, smem, s>>>((const uint8_t*)src + (size_t)i0 * g.src_image * ES, (uint8_t*)dst + (size_t)i0 * g.dst_image * ES, g);
    }
    return 1;
}

// Column-walking bicubic for fp32 images, C = 3 interleaved or C = 1 planes (resize_cubic3_walk.cuh).  1 = launched, 0 = shape
// not eligible.
// Most vertical segments a column strip is cut into: segments of >= 32 output rows when the batch already supplies CTAs (a segment
// re-reads its first tap rows and pays the table prologue), >= 8 rows for small batches, which need the CTAs more -- one 1440p frame per
// call: u8 bicubic 0.0315 -> 0.0130 ms, four frames 0.043 -> 0.034, 16 frames best at 32 rows (profiles/_cubic_segs.py).
static inline long long seg_cap(int ho, long long per_seg) {
    const int min_rows = per_seg >= 32 ? 32 : 8;
    return (ho + min_rows - 1) / min_rows;
}

template <int C>
static int launch_cubic_walk_f32(const float* src, float* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    constexpr int PX = 4 * C;
    if (w < 4 || h < 4 || (size_t)w * h * PX >= 0xffffffffull || (size_t)wo * ho * PX >= 0xffffffffull) return 0;
    if ((double)h / ho > 4.0) return 0;                                           // the walk filters every source row in a segment
    WalkGeom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * C; g.dst_image = (size_t)wo * ho * C;
    g.scale_x = (double)w / (double)wo; g.scale_y = (double)h / (double)ho;      // resize_naive.cpp:144
    g.strips = (wo + kWalkThreads - 1) / kWalkThreads;
    g.store16 = (((size_t)wo * PX) % 16 == 0 && ((uintptr_t)dst % 16) == 0) ? 1 : 0;
    const long long want = 8LL * 8 * current_sm_count();
    const long long per_seg = (long long)g.strips * images;
    long long segs = std::max<long long>(1, std::min<long long>((want + per_seg - 1) / per_seg, seg_cap(ho, per_seg)));
    if (const int v = knob(kKnobWalkSegs)) segs = std::max(1, v);   // tuning knob
    int rps = (int)((ho + segs - 1) / segs);
    rps = std::min(kWalkMaxRows, std::max(rps, 1));
    g.rows_per_seg = rps;
    g.segs = (ho + rps - 1) / rps;
    // widest byte span of a warp's 32 columns in a source row: x_first differences <= ceil(31 * scale) + 2 pixels, + 4 tap
    // pixels, + 15 (alignment of the span start), rounded up
    const int span_px = (int)std::ceil(31.0 * g.scale_x) + 2 + 4;
    g.ring_pitch = (span_px * PX + 15 + 15) & ~15;
    const bool sync_only = knob(kKnobWalkSync) != 0;   // tuning knob: register prefetch instead of the cp.async ring
    const bool async = !sync_only && ((size_t)w * PX) % 16 == 0 && ((size_t)w * h * PX) % 16 == 0 && ((uintptr_t)src % 16) == 0 && g.ring_pitch <= 1024;
    const size_t smem = (size_t)(rps + 1) * sizeof(WalkRow) + (size_t)(kWalkThreads / 32) * kWalkStageRows * 32 * PX +
                        (async ? (size_t)(kWalkThreads / 32) * kWalkRing * g.ring_pitch : 0);
    auto kern = async ? resize_cubic3_walk_f32_kernel<C, true> : resize_cubic3_walk_f32_kernel<C, false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    }
    for (int i0 = 0; i0 < images; i0 += 65535) {
        dim3 grid(g.strips * g.segs, std::min(images - i0, 65535));
        kern<<<grid, kWalkThreads, smem, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, g);
    }
    return 1;
}

// u8: two columns per thread, packed fp32 vertical pass, tail pixels rewritten by a second small kernel.
static int launch_cubic3_walk2(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    if (((size_t)w * 3) % 4 != 0 || ((uintptr_t)src % 4) != 0) return 0;        // rows are read as aligned 32-bit words
    if (w < 4 || h < 4 || (size_t)w * h * 3 >= 0xffffffffull || (size_t)wo * ho * 3 >= 0xffffffffull) return 0;
    if ((double)h / ho > 4.0) return 0;                                           // the walk filters every source row in a segment
    Walk2Geom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * 3; g.dst_image = (size_t)wo * ho * 3;
    g.scale_x = 1. / ((double)wo / (double)w); g.scale_y = 1. / ((double)ho / (double)h);   // OpenCV 2.4
    g.strips = (wo + kWalk2Cols - 1) / kWalk2Cols;
    g.store16 = (((size_t)wo * 3) % 16 == 0 && ((uintptr_t)dst % 16) == 0) ? 1 : 0;
    g.one2 = 0x3F8000003F800000ull; g.negzero2 = 0x8000000080000000ull; g.magic2 = 0x4B4000004B400000ull; g.negmagic2 = 0xCB400000CB400000ull;
    const long long want = 8LL * 6 * current_sm_count();
    const long long per_seg = (long long)g.strips * images;
    long long segs = std::max<long long>(1, std::min<long long>((want + per_seg - 1) / per_seg, seg_cap(ho, per_seg)));
    if (const int v = knob(kKnobWalkSegs)) segs = std::max(1, v);   // tuning knob
    int rps = (int)((ho + segs - 1) / segs);
    rps = std::min(kWalkMaxRows, std::max(rps, 1));
    g.rows_per_seg = rps;
    g.segs = (ho + rps - 1) / rps;
    // widest byte span of a warp's 64 columns in a source row: x_first differences <= ceil(63 * scale) + 2 pixels, + 12 tap
    // bytes + 4 (word 3) + 15 (alignment of the span start), rounded up, + 16 slack
    const int span_px = (int)std::ceil(63.0 * g.scale_x) + 2;
    g.ring_pitch = ((span_px * 3 + 12 + 4 + 15 + 15) & ~15) + 16;
    const bool sync_only = knob(kKnobWalk2Sync) != 0;   // tuning knob: register prefetch instead of the cp.async ring
    const bool async = !sync_only && ((size_t)w * 3) % 16 == 0 && ((uintptr_t)src % 16) == 0 && g.ring_pitch <= 1024;
    const size_t smem = (size_t)(rps + 1) * sizeof(Walk2Row) + (size_t)(kWalkThreads / 32) * kWalkStageRows * 64 * 3 +
                        (async ? (size_t)(kWalkThreads / 32) * kWalk2Ring * g.ring_pitch : 0);
    auto kern = async ? resize_cubic3_walk2_kernel<true> : resize_cubic3_walk2_kernel<false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    }
    const int first_px = ((wo * 3) & ~7) / 3;
    for (int i0 = 0; i0 < images; i0 += 65535) {
        const int n = std::min(images - i0, 65535);
        dim3 grid(g.strips * g.segs, n);
        kern<<<grid, kWalkThreads, smem, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, g);
        if (first_px < wo) {
            const long long items = (long long)n * ho * (wo - first_px);
            resize_cubic3_tail_kernel<<<(unsigned)((items + 127) / 128), 128, 0, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, w, h, wo, ho,
                                                                                  g.scale_x, g.scale_y, g.src_image, g.dst_image, first_px, n);
        }
    }
    return 1;
}

// u8, second generation (resize_cubic3_walkn.cuh): NC = 4 (or 2) columns per thread, 2..4 warps per CTA chosen for the least padding.
// Needs 16-byte aligned source rows (cp.async ring); everything else stays on launch_cubic3_walk2.
struct WalkNPlan { int h, ho; bool down, valid; };
static bool walkn_rows_strictly_increase(int h, int ho, double scale_y) {
    // kDown's precondition, checked with the device's own arithmetic (cubic_cv_coord_scaled): every output row ends on a later source row
    static thread_local PlanCache<WalkNPlan, 8> cache;
    if (WalkNPlan* p = cache.find([&](const WalkNPlan& q) { return q.h == h && q.ho == ho; })) return p->down;
    bool down = true;
    int prev = INT_MIN;
    for (int d = 0; d < ho && down; ++d) {
        const float f = (float)(((double)d + 0.5) * scale_y - 0.5);
        const int sy = (int)floorf(f);
        down = sy > prev;
        prev = sy;
    }
    WalkNPlan* p = cache.claim();
    p->h = h; p->ho = ho; p->down = down;
    cache.commit();
    return down;
}

template <int NC, int MAXREG = (NC == 4 ? 128 : 80), bool kPre = true>
static int launch_cubic3_walkn_nc(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    WalkNGeom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * 3; g.dst_image = (size_t)wo * ho * 3;
    g.scale_x = 1. / ((double)wo / (double)w); g.scale_y = 1. / ((double)ho / (double)h);   // OpenCV 2.4
    const int span_px = (int)std::ceil((32.0 * NC - 1) * g.scale_x) + 2;
    g.ring_pitch = ((span_px * 3 + 12 + 4 + 15 + 15) & ~15) + 16;
    if (g.ring_pitch > 1024) return 0;                 // two 16-byte chunks per lane and row at most
    g.warp_strips = (wo + 32 * NC - 1) / (32 * NC);
    int warps = 4, best_pad = INT_MAX;
    for (int wv = 4; wv >= 2; --wv) {
        const int pad = (g.warp_strips + wv - 1) / wv * wv - g.warp_strips;
        if (pad < best_pad) { best_pad = pad; warps = wv; }
    }
    g.cta_strips = (g.warp_strips + warps - 1) / warps;
    g.store16 = (((size_t)wo * 3) % 16 == 0 && ((uintptr_t)dst % 16) == 0) ? 1 : 0;
    g.one2 = 0x3F8000003F800000ull; g.negzero2 = 0x8000000080000000ull; g.magic2 = 0x4B4000004B400000ull; g.negmagic2 = 0xCB400000CB400000ull;
    const int per_sm = 65536 / (MAXREG * 32 * warps);   // resident CTAs per SM at the kernel's register budget
    const long long want = 8LL * per_sm * current_sm_count();
    const long long per_seg = (long long)g.cta_strips * images;
    long long segs = std::max<long long>(1, std::min<long long>((want + per_seg - 1) / per_seg, seg_cap(ho, per_seg)));
    if (const int v = knob(kKnobWalkSegs)) segs = std::max(1, v);   // tuning knob
    int rps = (int)((ho + segs - 1) / segs);
    rps = std::min(kWalkMaxRows, std::max(rps, 1));
    g.rows_per_seg = rps;
    g.segs = (ho + rps - 1) / rps;
    const size_t smem = (size_t)(rps + 1) * sizeof(Walk2Row) + (size_t)warps * kWnStageRows * 32 * NC * 3 + (size_t)warps * kWnRing * g.ring_pitch;
    const bool down = walkn_rows_strictly_increase(h, ho, g.scale_y);
    auto kern = down ? resize_cubic3_walkn_kernel<NC, true, MAXREG, kPre> : resize_cubic3_walkn_kernel<NC, false, MAXREG, kPre>;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return 0;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    }
    const int first_px = ((wo * 3) & ~7) / 3;
    for (int i0 = 0; i0 < images; i0 += 65535) {
        const int n = std::min(images - i0, 65535);
        dim3 grid(g.cta_strips * g.segs, n);
        kern<<<grid, 32 * warps, smem, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, g);
        if (first_px < wo) {
            const long long items = (long long)n * ho * (wo - first_px);
            resize_cubic3_tail_kernel<<<(unsigned)((items + 127) / 128), 128, 0, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, w, h, wo, ho,
                                                                                  g.scale_x, g.scale_y, g.src_image, g.dst_image, first_px, n);
        }
    }
    return 1;
}

// u8, third generation for rational horizontal scales (resize_cubic3_period.cuh): w : wo = P : Q, a thread owns KP periods of adjacent
// columns.  Returns 0 (not eligible) unless every column's taps sit where the kernel's compile-time pattern expects them.
struct PeriodPlan { int w, wo, P, Q, KP; bool ok; };
template <int P, int Q, int KP>
static bool period_taps_match(int w, int wo, double scale_x) {
    using S = PeriodShape<P, Q, KP>;
    static thread_local PlanCache<PeriodPlan, 8> cache;
    if (PeriodPlan* p = cache.find([&](const PeriodPlan& q) { return q.w == w && q.wo == wo && q.P == P && q.Q == Q && q.KP == KP; })) return p->ok;
    bool ok = true;
    for (int dx = 0; dx < wo && ok; ++dx) {            // the device's own coordinate arithmetic (cubic_cv_coord_scaled, is_x)
        float f = (float)(((double)dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(f);
        if (sx < 0) sx = 0;
        if (sx >= w - 1) sx = w - 1;
        const int pt = dx / S::NCOL, c = dx - pt * S::NCOL;
        const int base = P * KP * pt - 1 + pd::tap0(P, Q, c);
        for (int j = 0; j < 4; ++j) {
            const int pos = std::min(std::max(sx - 1 + j, 0), w - 1) - base;
            ok = ok && pos >= 0 && pos <= 3;
        }
    }
    PeriodPlan* p = cache.claim();
    p->w = w; p->wo = wo; p->P = P; p->Q = Q; p->KP = KP; p->ok = ok;
    cache.commit();
    return ok;
}

// kVStat's precondition, checked with the device's own arithmetic: in every segment of rps output rows, row r completes at walk step
// 3 + r + r / 3 (steps counted from the segment's first tap row).
struct PeriodRowsPlan { int h, ho, rps; bool ok; };
static bool period_rows_match_4to3(int h, int ho, int rps, double scale_y) {
    static thread_local PlanCache<PeriodRowsPlan, 8> cache;
    if (PeriodRowsPlan* p = cache.find([&](const PeriodRowsPlan& q) { return q.h == h && q.ho == ho && q.rps == rps; })) return p->ok;
    auto row_index = [&](int d) { return (int)floorf((float)(((double)d + 0.5) * scale_y - 0.5)); };
    bool ok = true;
    for (int d0 = 0; d0 < ho && ok; d0 += rps) {
        const int t_first = row_index(d0) - 1, nrows = std::min(rps, ho - d0);
        for (int r = 0; r < nrows && ok; ++r) ok = row_index(d0 + r) + 2 - t_first == 3 + r + r / 3;
    }
    PeriodRowsPlan* p = cache.claim();
    p->h = h; p->ho = ho; p->rps = rps; p->ok = ok;
    cache.commit();
    return ok;
}

template <int P, int Q, int KP, int MAXREG>
static int launch_cubic3_period(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    using S = PeriodShape<P, Q, KP>;
    if ((long long)w * Q != (long long)wo * P || wo % S::NCOL != 0) return 0;
    if (((size_t)wo * 3) % 16 != 0 || ((uintptr_t)dst % 16) != 0) return 0;      // staged rows leave as aligned 16-byte chunks
    PeriodGeom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * 3; g.dst_image = (size_t)wo * ho * 3;
    g.scale_x = 1. / ((double)wo / (double)w); g.scale_y = 1. / ((double)ho / (double)h);   // OpenCV 2.4
    if (!period_taps_match<P, Q, KP>(w, wo, g.scale_x)) return 0;
    g.ring_pitch = S::kNeed;
    g.warp_strips = (wo + 32 * S::NCOL - 1) / (32 * S::NCOL);
    int warps = 4, best_pad = INT_MAX;
    for (int wv = 4; wv >= 2; --wv) {
        const int pad = (g.warp_strips + wv - 1) / wv * wv - g.warp_strips;
        if (pad < best_pad) { best_pad = pad; warps = wv; }
    }
    g.cta_strips = (g.warp_strips + warps - 1) / warps;
    g.one2 = 0x3F8000003F800000ull; g.negzero2 = 0x8000000080000000ull; g.magic2 = 0x4B4000004B400000ull; g.negmagic2 = 0xCB400000CB400000ull;
    const int per_sm = std::max(1, 65536 / (MAXREG * 32 * warps));   // resident CTAs per SM at the kernel's register budget
    const long long want = 8LL * per_sm * current_sm_count();
    const long long per_seg = (long long)g.cta_strips * images;
    long long segs = std::max<long long>(1, std::min<long long>((want + per_seg - 1) / per_seg, seg_cap(ho, per_seg)));
    if (const int v = knob(kKnobWalkSegs)) segs = std::max(1, v);   // tuning knob
    int rps = (int)((ho + segs - 1) / segs);
    rps = std::min(kWalkMaxRows, std::max(rps, 1));
    // vertical 4 : 3 as well (config 4): segments of whole periods (3 output rows), so that every segment sees the same emit pattern
    // -- output row r of a segment completes at walk step 3 + r + r / 3 -- which the kernel's kVStat variant has compiled in
    bool vstat = (long long)h * 3 == (long long)ho * 4 && knob(kKnobCubicV) != 22;
    if (vstat) rps = std::min(kWalkMaxRows / 3 * 3, (rps + 2) / 3 * 3);
    g.rows_per_seg = rps;
    g.segs = (ho + rps - 1) / rps;
    if (vstat) vstat = period_rows_match_4to3(h, ho, rps, g.scale_y);
    const size_t smem = (size_t)(rps + 1) * sizeof(Walk2Row) + (size_t)warps * (kPdStageRows * S::kWarpRow + kPdRing * g.ring_pitch + kPdRing * 8);
    const bool down = walkn_rows_strictly_increase(h, ho, g.scale_y);
    auto kern = down ? (vstat ? resize_cubic3_period_kernel<P, Q, KP, true, MAXREG, true> : resize_cubic3_period_kernel<P, Q, KP, true, MAXREG>)
                     : resize_cubic3_period_kernel<P, Q, KP, false, MAXREG>;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return 0;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    }
    const int first_px = ((wo * 3) & ~7) / 3;
    for (int i0 = 0; i0 < images; i0 += 65535) {
        const int n = std::min(images - i0, 65535);
        dim3 grid(g.cta_strips * g.segs, n);
        kern<<<grid, 32 * warps, smem, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, g);
        if (first_px < wo) {
            const long long items = (long long)n * ho * (wo - first_px);
            resize_cubic3_tail_kernel<<<(unsigned)((items + 127) / 128), 128, 0, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, w, h, wo, ho,
                                                                                  g.scale_x, g.scale_y, g.src_image, g.dst_image, first_px, n);
        }
    }
    return 1;
}

// fp32 planes at rational horizontal scales (resize_cubic_f32_period.cuh).  Host twin of cubic_naive_scaled (resize_coeffs.cuh; same
// IEEE operations: this file is built with contraction off on both sides) for the launcher's checks.
static void host_cubic_naive(int d, int n_in, double scale, int& ofs, float (&a)[4]) {
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
struct PeriodF32Plan { int w, wo, h, ho, P, Q, KP; bool ok, down; };
template <int P, int Q, int KP>
static const PeriodF32Plan* period_f32_plan(int w, int h, int wo, int ho, double scale_x, double scale_y) {
    using S = PeriodF32Shape<1, P, Q, KP>;             // the tap pattern does not depend on the channel count
    static thread_local PlanCache<PeriodF32Plan, 8> cache;
    if (PeriodF32Plan* p = cache.find([&](const PeriodF32Plan& q) { return q.w == w && q.wo == wo && q.h == h && q.ho == ho && q.P == P && q.Q == Q && q.KP == KP; }))
        return p;
    bool ok = true;
    for (int dx = 0; dx < wo && ok; ++dx) {            // every tap with a non-zero coefficient inside the column's compile-time window
        int ofs; float a[4];
        host_cubic_naive(dx, w, scale_x, ofs, a);
        const int pt = dx / S::NCOL, c = dx - pt * S::NCOL;
        const int base = P * KP * pt - 1 + pd::tap0(P, Q, c);
        for (int j = 0; j < 4; ++j) {
            const int pos = ofs - 1 + j - base;
            if (a[j] != 0.f) ok = ok && pos >= 0 && pos <= 3;
        }
    }
    bool down = true;                                  // kDown's precondition: every output row ends on a later source row
    int prev = INT_MIN;
    for (int d = 0; d < ho && down; ++d) {
        int ofs; float a[4];
        host_cubic_naive(d, h, scale_y, ofs, a);
        down = ofs > prev;
        prev = ofs;
    }
    PeriodF32Plan* p = cache.claim();
    p->w = w; p->wo = wo; p->h = h; p->ho = ho; p->P = P; p->Q = Q; p->KP = KP; p->ok = ok; p->down = down;
    cache.commit();
    return p;
}
template <int C, int P, int Q, int KP>
static int launch_cubic_f32_period(const float* src, float* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    using S = PeriodF32Shape<C, P, Q, KP>;
    if ((long long)w * Q != (long long)wo * P || wo % S::NCOL != 0 || ((w * C) % 4) != 0 || ((wo * C) % 4) != 0) return 0;   // rows are whole 16-byte chunks
    if (((uintptr_t)src % 16) != 0 || ((uintptr_t)dst % 16) != 0 || ((size_t)w * h * C) % 4 != 0 || ((size_t)wo * ho * C) % 4 != 0) return 0;
    if (w < 4 || h < 4 || (size_t)w * h * 4 * C >= 0xffffffffull || (size_t)wo * ho * 4 * C >= 0xffffffffull || (double)h / ho > 4.0) return 0;
    PeriodF32Geom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * C; g.dst_image = (size_t)wo * ho * C;
    g.scale_x = (double)w / (double)wo; g.scale_y = (double)h / (double)ho;      // resize_naive.cpp:144
    const PeriodF32Plan* plan = period_f32_plan<P, Q, KP>(w, h, wo, ho, g.scale_x, g.scale_y);
    if (!plan->ok) return 0;
    g.warp_strips = (wo + 32 * S::NCOL - 1) / (32 * S::NCOL);
    int warps = 4, best_pad = INT_MAX;
    for (int wv = 4; wv >= 1; --wv) {
        const int pad = (g.warp_strips + wv - 1) / wv * wv - g.warp_strips;
        if (pad < best_pad) { best_pad = pad; warps = wv; }
    }
    g.cta_strips = (g.warp_strips + warps - 1) / warps;
    const long long want = 8LL * 16 * current_sm_count();
    const long long per_seg = (long long)g.cta_strips * images;
    long long segs = std::max<long long>(1, std::min<long long>((want + per_seg - 1) / per_seg, seg_cap(ho, per_seg)));
    if (const int v = knob(kKnobWalkSegs)) segs = std::max(1, v);   // tuning knob
    int rps = (int)((ho + segs - 1) / segs);
    rps = std::min(kWalkMaxRows, std::max(rps, 1));
    g.rows_per_seg = rps;
    g.segs = (ho + rps - 1) / rps;
    const size_t smem = (size_t)(rps + 1) * sizeof(WalkRow) + (size_t)warps * (kPdRing * S::kNeed + (S::kDirect ? 0 : 2 * S::kWarpRow) + kPdRing * 8);
    auto kern = plan->down ? resize_cubic_f32_period_kernel<C, P, Q, KP, true> : resize_cubic_f32_period_kernel<C, P, Q, KP, false>;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return 0;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    }
    for (int i0 = 0; i0 < images; i0 += 65535) {
        dim3 grid(g.cta_strips * g.segs, std::min(images - i0, 65535));
        kern<<<grid, 32 * warps, smem, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, g);
    }
    return 1;
}

static int launch_cubic3_walkn(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    if (((size_t)w * 3) % 16 != 0 || ((uintptr_t)src % 16) != 0) return 0;     // the cp.async ring copies aligned 16-byte chunks
    if (w < 4 || h < 4 || (size_t)w * h * 3 >= 0xffffffffull || (size_t)wo * ho * 3 >= 0xffffffffull) return 0;
    if ((double)h / ho > 4.0) return 0;                                           // the walk filters every source row in a segment
    const int v = knob(kKnobCubicV);                                              // tuning knob: 0 = automatic, 2 / 4 = columns per thread
    int rc = 0;
    if (v == 0 || v == 20 || v == 22) {                                           // rational horizontal scales: periodic walker (22: without the compiled-in vertical pattern)
        rc = launch_cubic3_period<4, 3, 2, 168>(src, dst, images, w, h, wo, ho, s);   // 4 : 3 (config 4: 2560 -> 1920)
        if (rc == 0) rc = launch_cubic3_period<2, 1, 4, 128>(src, dst, images, w, h, wo, ho, s);   // 2 : 1 (3840 -> 1920)
        if (rc == 0) rc = launch_cubic3_period<3, 2, 2, 128>(src, dst, images, w, h, wo, ho, s);   // 3 : 2 (1920 -> 1280)
    }
    if (v == 21) rc = launch_cubic3_period<4, 3, 2, 128>(src, dst, images, w, h, wo, ho, s);   // experiment: 128 registers, 16 warps per SM
    if (rc != 0) return rc;
    if (v == 0 || v == 4) rc = launch_cubic3_walkn_nc<4>(src, dst, images, w, h, wo, ho, s);
    if (rc == 0 && v != 1) rc = launch_cubic3_walkn_nc<2>(src, dst, images, w, h, wo, ho, s);
    return rc;
}

template <int KIND>
static int launch_tiled_kind(const void* src, void* dst, int images, TiledGeom g, size_t smem, cudaStream_t s) {
    auto kern = resize_tiled_kernel<KIND>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
    }
    constexpr int ES = sizeof(typename Kind<KIND>::S);
    const int ychunks = (g.tiles_y + g.tiles_per_cta - 1) / g.tiles_per_cta;
    for (int i0 = 0; i0 < images; i0 += 65535) {
        dim3 grid(g.tiles_x * ychunks, std::min(images - i0, 65535));
        kern<<<grid, kRtThreads, smem, s>>>((const uint8_t*)src + (size_t)i0 * g.src_image * ES,
                                            (uint8_t*)dst + (size_t)i0 * g.dst_image * ES, g);
    }
    return 1;
}

// Returns 1 if launched, 0 if the shape does not fit the tiled kernel (caller uses the direct kernels), < 0 on error.
int try_launch_resize_tiled(int kind, const void* src, void* dst, int images, int w, int h, int c, int wo, int ho, cudaStream_t s) {
    if (c == 3 && (kind == kCubU8 || kind == kCubF32)) {   // interleaved BGR: rolling separable kernel, else the tiled pixel-per-thread one
        const bool roll = knob(kKnobCubic3Roll) != 0;   // tuning knob: shared-memory ring kernel instead of the column walker
        int rc = 0;
        if (kind == kCubU8 && !roll && knob(kKnobCubicV) != 1) rc = launch_cubic3_walkn((const uint8_t*)src, (uint8_t*)dst, images, w, h, wo, ho, s);
        if (rc != 0) return rc;
        if (kind == kCubU8 && !roll) rc = launch_cubic3_walk2((const uint8_t*)src, (uint8_t*)dst, images, w, h, wo, ho, s);   // CUBIC_V=1: first generation
        if (rc != 0) return rc;
        if (kind == kCubF32 && !roll && knob(kKnobCubicV) != 1) {   // rational horizontal scales: periodic walker for interleaved fp32
            rc = launch_cubic_f32_period<3, 4, 3, 1>((const float*)src, (float*)dst, images, w, h, wo, ho, s);          // 4 : 3 (2560 -> 1920)
            if (rc == 0) rc = launch_cubic_f32_period<3, 3, 2, 2>((const float*)src, (float*)dst, images, w, h, wo, ho, s);   // 3 : 2 (1920 -> 1280)
            if (rc != 0) return rc;
        }
        if (kind == kCubF32 && !roll) rc = launch_cubic_walk_f32<3>((const float*)src, (float*)dst, images, w, h, wo, ho, s);
        if (rc != 0) return rc;
        rc = kind == kCubU8 ? launch_cubic3_rolling<true>(src, dst, images, w, h, wo, ho, s) : launch_cubic3_rolling<false>(src, dst, images, w, h, wo, ho, s);
        if (rc != 0) return rc;
        rc = kind == kCubU8 ? launch_cubic3<kCubU8>(src, dst, images, w, h, wo, ho, s) : launch_cubic3<kCubF32>(src, dst, images, w, h, wo, ho, s);
        if (rc != 0) return rc;
    }
    if (c == 1 && kind == kCubF32) {   // planes of a CHW tensor
        int rc = 0;
        if (knob(kKnobCubicV) != 1) {   // rational horizontal scales: periodic walker (CUBIC_V=1: the one-column walker everywhere)
            rc = launch_cubic_f32_period<1, 3, 2, 4>((const float*)src, (float*)dst, images, w, h, wo, ho, s);          // 3 : 2 (1920 -> 1280)
            if (rc == 0) rc = launch_cubic_f32_period<1, 4, 3, 4>((const float*)src, (float*)dst, images, w, h, wo, ho, s);   // 4 : 3 (2560 -> 1920)
            // (2 : 1 stays on the one-column walker: its lanes are 8 bytes apart there already, 4.0 - 4.7 TB/s; the periodic form measured the same)
            if (rc != 0) return rc;
        }
        rc = launch_cubic_walk_f32<1>((const float*)src, (float*)dst, images, w, h, wo, ho, s);
        if (rc != 0) return rc;
    }
    const int es = (kind == kLinF32 || kind == kCubF32) ? 4 : 1;
    const int K = (kind == kCubF32 || kind == kCubU8) ? 4 : 2;
    if (c > 64 || (size_t)wo * c > 0x3fffffff) return 0;
    TiledGeom g;
    g.w = w; g.h = h; g.c = c; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * c; g.dst_image = (size_t)wo * ho * c;
    g.TWE = std::min(256, wo * c) / c * c;
    if (g.TWE <= 0) return 0;
    const size_t row_bytes = (size_t)w * c * es;
    g.align = ((row_bytes % 16) == 0 && ((uintptr_t)src % 16) == 0) ? 16 : ((row_bytes % 4) == 0 && ((uintptr_t)src % 4) == 0) ? 4 : 1;
    const double sx = (double)w / wo, sy = (double)h / ho;
    // coordinate scales with the reference's expression per kind (host IEEE arithmetic == device, no FMA involved)
    if (kind == kCubU8) { g.scale_x = 1. / ((double)wo / (double)w); g.scale_y = 1. / ((double)ho / (double)h); }            // OpenCV 2.4
    else if (kind == kCubF32 || kind == kLinU8Neon) { g.scale_x = (double)w / (double)wo; g.scale_y = (double)h / (double)ho; }  // resize_naive.cpp:144, resize_neon.cpp:17-18
    else { g.scale_x = (double)((float)w / (float)wo); g.scale_y = (double)((float)h / (float)ho); }                          // resize_naive.cpp:17-18
    const int tw = g.TWE / c;
    const int span_px = std::min(w, (int)(sx * (tw - 1)) + K + 3);
    g.src_pitch = (int)((((size_t)span_px * c * es + 2 * g.align + 15)) & ~(size_t)15);
    g.opitch = (g.TWE * es + 16 + 15) & ~15;
    g.hpitch = (g.TWE + 3) & ~3;
    g.dense = sy <= (double)K ? 1 : 0;
    const size_t budget = 56 * 1024;
    int TH = kRtMaxTH;
    for (; TH >= 1; --TH) {
        const int band = (int)(sy * (TH - 1)) + K + 3;
        const int slots = g.dense ? band : std::min(band, TH * K);
        const size_t smem = (size_t)slots * g.src_pitch + (K == 4 ? (size_t)slots * g.hpitch * 4 : 0) + (size_t)TH * g.opitch;
        if (band <= kRtMaxBand && smem <= budget) {
            g.TH = TH; g.max_slots = slots;
            g.tiles_x = (wo * c + g.TWE - 1) / g.TWE; g.tiles_y = (ho + TH - 1) / TH;
            if ((long long)g.tiles_x * g.tiles_y > 0x7fffffffLL) return 0;
            // several y tiles per CTA amortise the per-CTA column set-up, but keep >= ~8 CTAs per SM in flight
            const long long ctas1 = (long long)g.tiles_x * g.tiles_y * images;
            g.tiles_per_cta = (int)std::max<long long>(1, std::min<long long>(8, ctas1 / (current_sm_count() * 8LL)));
            switch (kind) {
                case kLinU8: return launch_tiled_kind<kLinU8>(src, dst, images, g, smem, s);
                case kLinU8Signed: return launch_tiled_kind<kLinU8Signed>(src, dst, images, g, smem, s);
                case kLinU8Neon: return launch_tiled_kind<kLinU8Neon>(src, dst, images, g, smem, s);
                case kLinF32: return launch_tiled_kind<kLinF32>(src, dst, images, g, smem, s);
                case kCubF32: return launch_tiled_kind<kCubF32>(src, dst, images, g, smem, s);
                default: return launch_tiled_kind<kCubU8>(src, dst, images, g, smem, s);
            }
        }
    }
    return 0;
}

}  // namespace vacv
