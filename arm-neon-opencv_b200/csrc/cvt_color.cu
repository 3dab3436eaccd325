// a1: semi-planar YUV 4:2:0 -> interleaved BGR (reference: src/cv/cvt_color.cpp:39-135).
//
// HBM-bound streaming op: 1.5 B read + 3 B written per pixel.  A thread owns a 2-row x 16-pixel strip, so the
// 8 chroma pairs it needs are loaded once (the reference does the same per 2x2 quad, cvt_color.cpp:68-131):
// three 128-bit loads in, 2 x 48 B out.  The 48-byte-per-lane output is re-chunked through a per-warp shared
// memory buffer so that every global store is a full-sector, lane-contiguous 128-bit access.
#include <algorithm>

#include "vacv_common.cuh"

namespace vacv {

// Packed chroma terms of one 2x1 pixel pair: each 32-bit word holds the same 16-bit term twice, so that one DPX
// VIADDMNMX.S16x2.RELU adds it to two luma samples and clamps both to [0, 255] (cvt_color.cpp:76-100).
struct ChromaTerms2 { uint32_t ra2, nga2, ba2; };
__device__ __forceinline__ ChromaTerms2 chroma_terms2(int v, int u) {
    const ChromaTerms t = chroma_terms(v, u);
    ChromaTerms2 p;
    p.ra2 = __byte_perm((uint32_t)t.ra, 0u, 0x1010);
    p.nga2 = __byte_perm((uint32_t)(-t.ga), 0u, 0x1010);
    p.ba2 = __byte_perm((uint32_t)t.ba, 0u, 0x1010);
    return p;
}

// 16 pixels of one row: y = 16 luma bytes, terms for the 8 chroma pairs -> 12 output words (48 bytes of BGR).
// Per pixel pair: 1 PRMT (two luma bytes -> two 16-bit lanes), 3 packed add-clamps, then PRMT byte shuffles into BGR order.
__device__ __forceinline__ void convert16(const uint4& y, const ChromaTerms2 (&t)[8], uint32_t (&out)[12]) {
    const uint32_t yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // 4 pixels = pairs 2q, 2q+1 = luma word q -> 3 output words
        const uint32_t ya = __byte_perm(yw[q], 0u, 0x4140), yb = __byte_perm(yw[q], 0u, 0x4342);   // [y0 0 y1 0], [y2 0 y3 0]
        const ChromaTerms2 &ta = t[2 * q], &tb = t[2 * q + 1];
        const uint32_t Ba = __viaddmin_s16x2_relu(ya, ta.ba2, 0x00ff00ffu), Ga = __viaddmin_s16x2_relu(ya, ta.nga2, 0x00ff00ffu),
                       Ra = __viaddmin_s16x2_relu(ya, ta.ra2, 0x00ff00ffu);
        const uint32_t Bb = __viaddmin_s16x2_relu(yb, tb.ba2, 0x00ff00ffu), Gb = __viaddmin_s16x2_relu(yb, tb.nga2, 0x00ff00ffu),
                       Rb = __viaddmin_s16x2_relu(yb, tb.ra2, 0x00ff00ffu);
        const uint32_t xa = __byte_perm(Ba, Ga, 0x6240), xb = __byte_perm(Bb, Gb, 0x6240);   // [b0 g0 b1 g1], [b2 g2 b3 g3]
        out[3 * q + 0] = __byte_perm(xa, Ra, 0x2410);                                          // b0 g0 r0 b1
        out[3 * q + 1] = __byte_perm(__byte_perm(xa, Ra, 0x0063), xb, 0x5410);                 // g1 r1 b2 g2
        out[3 * q + 2] = __byte_perm(xb, Rb, 0x6324);                                          // r2 b3 g3 r3
    }
}

constexpr int kCvtThreads = 128;

enum { kCvtVU = 0, kCvtUV = 1, kCvtPlanar = 2 };   // interleaved chroma V-first (NV21) / U-first (NV12), separate U and V planes (I420 / YV12)

struct CvtGeom {
    int w, h;
    int y_pitch, c_pitch;          // bytes per luma / chroma row (planar: per U or V row)
    size_t frame_stride;           // bytes between frames
    size_t c_off, c2_off;          // chroma plane (semi-planar) or U and V planes (planar) inside a frame
};

// grid.x = batch * h/2 (row pairs), grid.y = ceil(w/16 / 128).   Requires w % 16 == 0 and 16-byte aligned rows (planar chroma: 8).
template <int FMT>
__global__ void __launch_bounds__(kCvtThreads) yuv2bgr_strip16_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, CvtGeom g) {
    __shared__ __align__(16) uint4 stage[kCvtThreads / 32][2][96];   // per warp: 2 rows x 1536 B
    const int w = g.w, h = g.h;
    const int strips = w >> 4;
    const int strip = blockIdx.y * kCvtThreads + threadIdx.x;
    const int pair = blockIdx.x % (h >> 1);
    const int frame = blockIdx.x / (h >> 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t out_frame = (size_t)w * h * 3;
    const uint8_t* f = src + frame * g.frame_stride;
    const uint8_t* y0p = f + (size_t)(2 * pair) * g.y_pitch;
    const bool active = strip < strips;

    if (active) {
        const uint4 y0 = ld_stream16(y0p + 16 * strip);
        const uint4 y1 = ld_stream16(y0p + g.y_pitch + 16 * strip);
        ChromaTerms2 t[8];
        if (FMT == kCvtPlanar) {
            const uint2 u8 = __ldg(reinterpret_cast<const uint2*>(f + g.c_off + (size_t)pair * g.c_pitch + 8 * strip));
            const uint2 v8 = __ldg(reinterpret_cast<const uint2*>(f + g.c2_off + (size_t)pair * g.c_pitch + 8 * strip));
            const uint32_t uw[2] = {u8.x, u8.y}, vw[2] = {v8.x, v8.y};
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = chroma_terms2((vw[i >> 2] >> (8 * (i & 3))) & 0xff, (uw[i >> 2] >> (8 * (i & 3))) & 0xff);
        } else {
            const uint4 vu = ld_stream16(f + g.c_off + (size_t)pair * g.c_pitch + 16 * strip);
            const uint32_t cw[4] = {vu.x, vu.y, vu.z, vu.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t pr = (cw[i >> 1] >> (16 * (i & 1))) & 0xffff;
                int c0 = pr & 0xff, c1 = pr >> 8;
                t[i] = FMT == kCvtVU ? chroma_terms2(c0, c1) : chroma_terms2(c1, c0);
            }
        }
        uint32_t o[12];
        convert16(y0, t, o);
#pragma unroll
        for (int k = 0; k < 3; ++k) stage[warp][0][3 * lane + k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        convert16(y1, t, o);
#pragma unroll
        for (int k = 0; k < 3; ++k) stage[warp][1][3 * lane + k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
    __syncwarp();
    // the warp's 32 strips are contiguous in the row: chunk q of the warp buffer -> byte 16*q of the warp's span
    const int warp_strip0 = blockIdx.y * kCvtThreads + warp * 32;
    const int valid_chunks = 3 * max(0, min(32, strips - warp_strip0));
    uint8_t* d0 = dst + frame * out_frame + (size_t)(2 * pair) * w * 3 + (size_t)warp_strip0 * 48;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int q = lane + 32 * k;
            if (q < valid_chunks) st_stream16(d0 + (size_t)r * w * 3 + 16 * q, stage[warp][r][q]);
        }
}

// any even w, h, any pitch: one thread per 2x2 quad (the reference's own unit of work)
template <int FMT>
__global__ void yuv2bgr_quad_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, CvtGeom g, size_t quads) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    const int w = g.w, h = g.h;
    const int qw = w >> 1, qh = h >> 1;
    const int qx = (int)(q % qw);
    const int qy = (int)((q / qw) % qh);
    const size_t frame = q / ((size_t)qw * qh);
    const uint8_t* f = src + frame * g.frame_stride;
    ChromaTerms t;
    if (FMT == kCvtPlanar) {
        t = chroma_terms(f[g.c2_off + (size_t)qy * g.c_pitch + qx], f[g.c_off + (size_t)qy * g.c_pitch + qx]);
    } else {
        const uint8_t* cp = f + g.c_off + (size_t)qy * g.c_pitch + 2 * qx;
        const int c0 = cp[0], c1 = cp[1];
        t = FMT == kCvtVU ? chroma_terms(c0, c1) : chroma_terms(c1, c0);
    }
    uint8_t* o = dst + frame * ((size_t)w * h * 3);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const int row = 2 * qy + r, col = 2 * qx + x;
            const size_t p = (size_t)row * w + col;
            int Y = f[(size_t)row * g.y_pitch + col];
            o[3 * p + 0] = (uint8_t)add_clamp255(Y, t.ba);
            o[3 * p + 1] = (uint8_t)add_clamp255(Y, -t.ga);
            o[3 * p + 2] = (uint8_t)add_clamp255(Y, t.ra);
        }
}

}  // namespace vacv

using namespace vacv;

template <int FMT>
static void launch_cvt(const uint8_t* src, uint8_t* dst, int batch, const CvtGeom& g, bool aligned, cudaStream_t s) {
    if (aligned) {
        const long long pairs = (long long)batch * (g.h / 2);
        for (long long p0 = 0; p0 < pairs; p0 += 0x7fffff00LL / (g.h / 2) * (g.h / 2)) {   // whole frames per launch, grid.x < 2^31
            const long long np = std::min<long long>(pairs - p0, 0x7fffff00LL / (g.h / 2) * (g.h / 2));
            const size_t f0 = (size_t)(p0 / (g.h / 2));
            dim3 grid((unsigned)np, ceil_div(g.w / 16, kCvtThreads));
            yuv2bgr_strip16_kernel<FMT><<<grid, kCvtThreads, 0, s>>>(src + f0 * g.frame_stride, dst + f0 * (size_t)g.w * g.h * 3, g);
        }
    } else {
        const size_t quads = (size_t)batch * (g.w / 2) * (g.h / 2);
        yuv2bgr_quad_kernel<FMT><<<ceil_div(quads, 256), 256, 0, s>>>(src, dst, g, quads);
    }
}

static int cvt_dispatch(const char* who, const uint8_t* src, uint8_t* dst, int batch, int fmt, const CvtGeom& g, cudaStream_t s) {
    const int cpix = fmt == kCvtPlanar ? 8 : 16;   // chroma bytes per 16-pixel strip
    const bool aligned = (g.w % 16) == 0 && (g.y_pitch % 16) == 0 && (g.c_pitch % cpix) == 0 && (g.frame_stride % 16) == 0 &&
                         (g.c_off % cpix) == 0 && (g.c2_off % cpix) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    if (fmt == kCvtVU) launch_cvt<kCvtVU>(src, dst, batch, g, aligned, s);
    else if (fmt == kCvtUV) launch_cvt<kCvtUV>(src, dst, batch, g, aligned, s);
    else launch_cvt<kCvtPlanar>(src, dst, batch, g, aligned, s);
    return check_launch(who);
}

extern "C" int vacv_cuda_cvt_nv2bgr(const uint8_t* src, uint8_t* dst, int batch, int w, int h, int v_first, void* stream) {
    VACV_REQUIRE(src && dst, "cvt_nv2bgr: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0, "cvt_nv2bgr: non-positive size");
    VACV_REQUIRE((w % 2) == 0 && (h % 2) == 0, "cvt_nv2bgr: w and h must be even (got %dx%d)", w, h);
    CvtGeom g;
    g.w = w; g.h = h; g.y_pitch = w; g.c_pitch = w; g.frame_stride = (size_t)w * h * 3 / 2; g.c_off = (size_t)w * h; g.c2_off = 0;
    return cvt_dispatch("cvt_nv2bgr", src, dst, batch, v_first ? kCvtVU : kCvtUV, g, as_stream(stream));
}

// Next row 8f-1: the same colour matrix on decoder surfaces -- row pitch, NV12 / NV21 or planar I420 / YV12 chroma
// (the reference declares COLOR_YUV2BGR_YV12, cv.h:73, without implementing it).  dst: dense HWC BGR.
extern "C" int vacv_cuda_cvt_yuv2bgr(const uint8_t* src, const vacv_yuv_layout* layout, uint8_t* dst, int batch, void* stream) {
    VACV_REQUIRE(src && layout && dst, "cvt_yuv2bgr: null pointer");
    const int w = layout->w, h = layout->h;
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0, "cvt_yuv2bgr: non-positive size");
    VACV_REQUIRE((w % 2) == 0 && (h % 2) == 0, "cvt_yuv2bgr: w and h must be even (got %dx%d)", w, h);
    const bool planar = layout->format == VACV_YUV_I420 || layout->format == VACV_YUV_YV12;
    if (!planar && layout->format != VACV_YUV_NV12 && layout->format != VACV_YUV_NV21)
        return set_error(VACV_ERR_UNSUPPORTED, "cvt_yuv2bgr: format %d", layout->format);
    CvtGeom g;
    g.w = w; g.h = h;
    g.y_pitch = layout->y_pitch ? layout->y_pitch : w;
    g.c_pitch = layout->c_pitch ? layout->c_pitch : (planar ? w / 2 : w);
    VACV_REQUIRE(g.y_pitch >= w && g.c_pitch >= (planar ? w / 2 : w), "cvt_yuv2bgr: pitch smaller than the row");
    const size_t y_bytes = (size_t)g.y_pitch * h, c_bytes = (size_t)g.c_pitch * (h / 2);
    g.frame_stride = layout->frame_stride ? layout->frame_stride : y_bytes + (planar ? 2 * c_bytes : c_bytes);
    VACV_REQUIRE(g.frame_stride >= y_bytes + (planar ? 2 * c_bytes : c_bytes), "cvt_yuv2bgr: frame_stride too small");
    int fmt;
    if (planar) {
        fmt = kCvtPlanar;
        g.c_off = layout->format == VACV_YUV_I420 ? y_bytes : y_bytes + c_bytes;    // U plane
        g.c2_off = layout->format == VACV_YUV_I420 ? y_bytes + c_bytes : y_bytes;   // V plane
    } else {
        fmt = layout->format == VACV_YUV_NV21 ? kCvtVU : kCvtUV;
        g.c_off = y_bytes; g.c2_off = 0;
    }
    return cvt_dispatch("cvt_yuv2bgr", src, dst, batch, fmt, g, as_stream(stream));
}
