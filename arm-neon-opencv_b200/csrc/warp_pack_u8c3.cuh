// a10, u8 BGR output: second form of the direct gather kernel (reference loop: src/cv/warp_affine_naive.cpp:9-58).
//
// ncu on warp_affine_u8c3_kernel<u8> (profiles/r2_warp_u8_ncu_raw.txt): 154 instructions per pixel at 86 % issue utilisation and
// 74 % of the L1 wavefront peak -- issue slots and L1 wavefronts together, DRAM at 0.59.  This form cuts both:
//   * a thread owns ONE output column: the CTA is `rows` whole output rows wide (blockDim = rows * wo), so m0*dx and m3*dx are
//     per-thread constants and m1*dy, m4*dy change by row -- the same fp32 products and sums in the same order, so bit-equal;
//   * the 6 tap bytes of a source row come from one aligned 64-bit load plus a second one only when they cross it (5 lanes in
//     8): 1.6 requests per tap row instead of 2.25, and with rows that are multiples of 8 bytes both tap rows share the
//     alignment arithmetic;
//   * the vertical weights carry a factor 4, so the blended value is byte 3 of the 32-bit sum (no shift);
//   * the three bytes of a pixel leave through ONE shuffle: lane j of a quad builds word j of the quad's 12 bytes from its own
//     pixel and its right neighbour's with a per-lane constant PRMT selector; 24 lanes of a warp store 96 contiguous bytes.
//     No staging line, no warp sync.
// Requirements (launcher): frames 8-byte aligned, w % 8 == 0 and (w*h*3) % 8 == 0, wo % 4 == 0, dst 4-byte aligned, wo <= 256.
#pragma once
#include "gather_u8c3.cuh"
#include "vacv_common.cuh"

namespace vacv {

__device__ __forceinline__ uint2 ldg_u64(const uint8_t* p) {
    return __ldg(reinterpret_cast<const uint2*>(p));
}

// grid = (crops, bands); blockDim.x = rows * wo rounded up to whole warps (rows = output rows per pass).  Branch-free: a pixel
// outside the frame or the band loads from offset 0 and its result is discarded.  U = passes per loop iteration, the tap loads of U
// pixels of a thread issued before the first blend: U = 2 measured no faster than U = 1 (0.223 vs 0.220 ms on config 3's shape) --
// long-scoreboard is the top stall, but what bounds the kernel is the L1 data pipe (75 % of its wavefront peak), not load latency.
template <bool kSigned, int U, bool kWide = true>
__global__ void __launch_bounds__(256) warp_affine_u8c3_pack_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ frame_idx,
                                                                     const float* __restrict__ minv, uint8_t* __restrict__ dst,
                                                                     int w, int h, int wo, int ho, size_t frame_bytes,
                                                                     int rows, int rows_per_cta, int crop0) {
    const int crop = crop0 + blockIdx.x;
    const float* mp = minv + 6 * (size_t)crop;
    const float m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3), m4 = __ldg(mp + 4), m5 = __ldg(mp + 5);
    const size_t f = frame_idx ? (size_t)__ldg(frame_idx + crop) : (size_t)crop;
    const uint8_t* img = frames + f * frame_bytes;
    const int tid = threadIdx.x, lane = tid & 31, j = lane & 3;
    const int ty = tid / wo, dx = tid - ty * wo;
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(y0 + rows_per_cta, ho);
    const float fdx = (float)dx;
    const float ax = m0 * fdx, ay = m3 * fdx;   // warp_affine_naive.cpp:23-24, first products
    const unsigned row = (unsigned)w * 3u;
    const int wm1 = w - 1, hm1 = h - 1;
    // word j of a quad's 12 bytes from lo = [. v0 v1 v2] of this lane and the same of the next lane
    const uint32_t sel = j == 0 ? 0x5321u : j == 1 ? 0x6532u : 0x7653u;
    uint8_t* o = dst + ((size_t)crop * wo * ho + (size_t)y0 * wo + tid) * 3 + j;   // byte 3 i + j: lane j's word of the quad
    const size_t o_step = (size_t)rows * wo * 3;   // blockDim.x may hold padding lanes
    float fdy = (float)(y0 + ty);
    const float frows = (float)rows;
    const bool live = ty < rows;                   // false: padding lanes of the last warp
    for (int yb = y0; yb < y1; yb += U * rows) {   // CTA-uniform loop: the shuffle below needs whole warps
        bool valid[U], in[U];
        uint32_t cx[U], q0[U], q1[U];
        unsigned r[U];
        uint2 t0[U], u0[U], t1[U], u1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            valid[u] = live && yb + u * rows + ty < y1;
            float fx = (ax + m1 * fdy) + m2;
            float fy = (ay + m4 * fdy) + m5;
            fdy += frows;
            const float flx = floorf(fx), fly = floorf(fy);
            const int sx = (int)flx, sy = (int)fly;
            in[u] = valid[u] && !(sy < 0 || sy >= hm1 || sx < 0 || sx >= wm1);
            fx -= flx;
            fy -= fly;
            // SATURATE_CAST_SHORT of values in [0, 2048] (macro.h:25-30) == trunc(x + 0.5f); see warp_taps_fast
            const int cx0 = (int)((1.f - fx) * 2048.f + 0.5f), cy0 = (int)((1.f - fy) * 2048.f + 0.5f);
            cx[u] = (uint32_t)cx0 | ((uint32_t)(2048 - cx0) << 16);
            q0[u] = (uint32_t)(cy0 << 2);          // 4 * cy: 255 * 2048 * 8192 < 2^32, the result is byte 3 of the sum
            q1[u] = 8192u - q0[u];
            const unsigned a = in[u] ? ((unsigned)sy * (unsigned)w + (unsigned)sx) * 3u : 0u;
            if (kWide) {
                r[u] = a & 7u;
                const uint8_t* p = img + (a & ~7u);
                t0[u] = ldg_u64(p); u0[u] = ldg_u64(p + row);
                t1[u] = make_uint2(0u, 0u); u1[u] = make_uint2(0u, 0u);
                if (r[u] > 2u) { t1[u] = ldg_u64(p + 8); u1[u] = ldg_u64(p + row + 8); }
            } else {   // 32-bit tap loads (linear_taps_u8c3): t0 / u0 hold the funnel-shifted words directly
                r[u] = 0u;
                linear_taps_u8c3(img, a, t0[u].x, t0[u].y);
                linear_taps_u8c3(img, a + row, u0[u].x, u0[u].y);
                t1[u] = make_uint2(0u, 0u); u1[u] = make_uint2(0u, 0u);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool hiw = r[u] >= 4u;
            const unsigned sh = r[u] * 8u;   // funnel shifts use the amount mod 32
            const uint32_t x0 = hiw ? t0[u].y : t0[u].x, x1 = hiw ? t1[u].x : t0[u].y, x2 = hiw ? t1[u].y : t1[u].x;
            const uint32_t b0 = kWide ? __funnelshift_r(x0, x1, sh) : t0[u].x, b1 = kWide ? __funnelshift_r(x1, x2, sh) : t0[u].y;
            const uint32_t z0 = hiw ? u0[u].y : u0[u].x, z1 = hiw ? u1[u].x : u0[u].y, z2 = hiw ? u1[u].y : u1[u].x;
            const uint32_t c0 = kWide ? __funnelshift_r(z0, z1, sh) : u0[u].x, c1 = kWide ? __funnelshift_r(z1, z2, sh) : u0[u].y;
            int Ht[3], Hb[3];
            hsum_u8c3<kSigned>(b0, b1, cx[u], Ht);   // p00*cx0 + p01*cx1
            hsum_u8c3<kSigned>(c0, c1, cx[u], Hb);   // p10*cx0 + p11*cx1
            // warp_affine_naive.cpp:50-54 regrouped row-wise, times 4: bits 22..29 of the reference's sum are byte 3 here
            // (unsigned arithmetic: the signed-char sums wrap mod 2^32 by design)
            const uint32_t s0 = (uint32_t)Ht[0] * q0[u] + (uint32_t)Hb[0] * q1[u], s1 = (uint32_t)Ht[1] * q0[u] + (uint32_t)Hb[1] * q1[u],
                           s2 = (uint32_t)Ht[2] * q0[u] + (uint32_t)Hb[2] * q1[u];
            const uint32_t lo = in[u] ? __byte_perm(__byte_perm(s0, s1, 0x0730), s2, 0x7210) : 0u;   // [. v0 v1 v2]; outside the frame: 0 (App. C-5)
            const uint32_t nxt = __shfl_down_sync(0xffffffffu, lo, 1);
            if (valid[u] && j != 3) st_stream4(o, __byte_perm(lo, nxt, sel));
            o += o_step;
        }
    }
}

}  // namespace vacv
