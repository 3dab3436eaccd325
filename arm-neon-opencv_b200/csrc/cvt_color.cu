// a1: semi-planar YUV 4:2:0 -> interleaved BGR (reference: src/cv/cvt_color.cpp:39-135).
//
// HBM-bound streaming op: 1.5 B read + 3 B written per pixel.  A thread owns a 2-row x 16-pixel strip, so the
// 8 chroma pairs it needs are loaded once (the reference does the same per 2x2 quad, cvt_color.cpp:68-131):
// three 128-bit loads in, 2 x 48 B out.  The 48-byte-per-lane output is re-chunked through a per-warp shared
// memory buffer so that every global store is a full-sector, lane-contiguous 128-bit access.
#include "vacv_common.cuh"

namespace vacv {

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}

// 16 pixels of one row: y = 16 luma bytes, terms for the 8 chroma pairs -> 12 output words (48 bytes of BGR)
__device__ __forceinline__ void convert16(const uint4& y, const ChromaTerms (&t)[8], uint32_t (&out)[12]) {
    const uint32_t yw[4] = {y.x, y.y, y.z, y.w};
    uint8_t px[48];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int Y = (yw[i >> 2] >> (8 * (i & 3))) & 0xff;
        const ChromaTerms& c = t[i >> 1];
        px[3 * i + 0] = (uint8_t)add_clamp255(Y, c.ba);
        px[3 * i + 1] = (uint8_t)add_clamp255(Y, -c.ga);
        px[3 * i + 2] = (uint8_t)add_clamp255(Y, c.ra);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) out[j] = pack4(px[4 * j], px[4 * j + 1], px[4 * j + 2], px[4 * j + 3]);
}

constexpr int kCvtThreads = 128;

// grid.x = batch * h/2 (row pairs), grid.y = ceil(w/16 / 128).   Requires w % 16 == 0 and 16-byte aligned frames.
template <bool kVFirst>
__global__ void __launch_bounds__(kCvtThreads) nv2bgr_strip16_kernel(const uint8_t* __restrict__ src,
                                                                       uint8_t* __restrict__ dst, int w, int h) {
    __shared__ __align__(16) uint4 stage[kCvtThreads / 32][2][96];   // per warp: 2 rows x 1536 B
    const int strips = w >> 4;
    const int strip = blockIdx.y * kCvtThreads + threadIdx.x;
    const int pair = blockIdx.x % (h >> 1);
    const int frame = blockIdx.x / (h >> 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t in_frame = (size_t)w * h * 3 / 2, out_frame = (size_t)w * h * 3;
    const uint8_t* y0p = src + frame * in_frame + (size_t)(2 * pair) * w;
    const uint8_t* cp = src + frame * in_frame + (size_t)w * h + (size_t)pair * w;
    const bool active = strip < strips;

    if (active) {
        const uint4 y0 = ld_stream16(y0p + 16 * strip);
        const uint4 y1 = ld_stream16(y0p + w + 16 * strip);
        const uint4 vu = ld_stream16(cp + 16 * strip);
        const uint32_t cw[4] = {vu.x, vu.y, vu.z, vu.w};
        ChromaTerms t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t pr = (cw[i >> 1] >> (16 * (i & 1))) & 0xffff;
            int c0 = pr & 0xff, c1 = pr >> 8;
            t[i] = kVFirst ? chroma_terms(c0, c1) : chroma_terms(c1, c0);
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

// any even w, h: one thread per 2x2 quad (the reference's own unit of work)
template <bool kVFirst>
__global__ void nv2bgr_quad_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w, int h, size_t quads) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    const int qw = w >> 1, qh = h >> 1;
    const int qx = (int)(q % qw);
    const int qy = (int)((q / qw) % qh);
    const size_t frame = q / ((size_t)qw * qh);
    const uint8_t* f = src + frame * ((size_t)w * h * 3 / 2);
    const uint8_t* cp = f + (size_t)w * h + (size_t)qy * w + 2 * qx;
    const int c0 = cp[0], c1 = cp[1];
    const ChromaTerms t = kVFirst ? chroma_terms(c0, c1) : chroma_terms(c1, c0);
    uint8_t* o = dst + frame * ((size_t)w * h * 3);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            size_t p = (size_t)(2 * qy + r) * w + 2 * qx + x;
            int Y = f[p];
            o[3 * p + 0] = (uint8_t)add_clamp255(Y, t.ba);
            o[3 * p + 1] = (uint8_t)add_clamp255(Y, -t.ga);
            o[3 * p + 2] = (uint8_t)add_clamp255(Y, t.ra);
        }
}

}  // namespace vacv

using namespace vacv;

extern "C" int vacv_cuda_cvt_nv2bgr(const uint8_t* src, uint8_t* dst, int batch, int w, int h, int v_first, void* stream) {
    VACV_REQUIRE(src && dst, "cvt_nv2bgr: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0, "cvt_nv2bgr: non-positive size");
    VACV_REQUIRE((w % 2) == 0 && (h % 2) == 0, "cvt_nv2bgr: w and h must be even (got %dx%d)", w, h);
    cudaStream_t s = as_stream(stream);
    const bool aligned = (w % 16) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    if (aligned) {
        dim3 grid((unsigned)(batch * (h / 2)), ceil_div(w / 16, kCvtThreads));
        if (v_first) nv2bgr_strip16_kernel<true><<<grid, kCvtThreads, 0, s>>>(src, dst, w, h);
        else nv2bgr_strip16_kernel<false><<<grid, kCvtThreads, 0, s>>>(src, dst, w, h);
    } else {
        size_t quads = (size_t)batch * (w / 2) * (h / 2);
        if (v_first) nv2bgr_quad_kernel<true><<<ceil_div(quads, 256), 256, 0, s>>>(src, dst, w, h, quads);
        else nv2bgr_quad_kernel<false><<<ceil_div(quads, 256), 256, 0, s>>>(src, dst, w, h, quads);
    }
    return check_launch("cvt_nv2bgr");
}
