// Bicubic resize of 3-channel interleaved images (the C4 shape): rolling, separable, one pass over the source.
//
// Reference arithmetic: u8 = OpenCV 2.4.13 cv::resize (the reference's only u8 cubic path,
// src/cv/resize.cpp:33-36); fp32 = ResizeNaive::resize_naive_inter_cubic_fp32_three_channel
// (src/cv/resize_naive.cpp:187-366).  Both are separable with a rolling window of horizontally filtered rows; this
// kernel keeps that structure on the GPU:
//
//   CTA      owns a strip of 128 output columns x a vertical segment of one image and walks down it in groups of G
//            output rows.  256 threads = 2 per column (they split the rows of each pass).
//   pass 1   every source row the segment touches is filtered horizontally exactly ONCE: a thread reads the 12
//            contiguous tap bytes of its column as aligned 32-bit words straight from global memory (lanes read
//            adjacent words -> 128-byte coalesced, overlap served by L1), regroups them per channel with PRMT and
//            forms sum(tap * coef) with two dp4a (11-bit coefficients split into a signed high and an unsigned low
//            byte).  The three sums go into a 16-row ring in shared memory as one float4 (exact: |H| < 2^22).
//   pass 2   an output pixel = four float4 ring reads, OpenCV's fp32 mul/add chain per channel, round-half-even and
//            clamp in one FADD + one DPX instruction; its last <= 7 elements per output row use OpenCV's integer tail.
//   output   rows are assembled in shared memory with the 16-byte phase of their destination and stored as
//            lane-contiguous 128-bit chunks.
#pragma once
#include <climits>
#include <type_traits>

#include "resize_coeffs.cuh"
#include "vacv_common.cuh"

namespace vacv {

constexpr int kRollPx = 128;        // output columns per CTA
constexpr int kRollThreads = 256;
constexpr int kRollRing = 16;       // ring capacity in source rows (power of two)
constexpr int kRollMaxG = 6;        // output rows per group (<= 8 new source rows per group at 4:3 -> 4 per thread)
constexpr int kRollMaxRows = 512;   // output rows per segment (row tables live in shared memory)

struct RollGeom {
    int w, h, wo, ho;
    int strips, segs, rows_per_seg, G;
    int opitch;
    double scale_x, scale_y;
    size_t src_image, dst_image;   // elements between images
};

struct __align__(16) RowEntry {   // per output row of the segment
    int off[4];     // byte offset (inside the ring) of the row of each vertical tap
    float b[4];     // the four vertical weights
};

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {   // unsigned bytes of a x signed bytes of b
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <bool kU8>
__global__ void __launch_bounds__(kRollThreads) resize_cubic3_rolling_kernel(const void* __restrict__ src_, void* __restrict__ dst_, RollGeom g) {
    using S = typename std::conditional<kU8, uint8_t, float>::type;
    constexpr int ES = sizeof(S), PX = 3 * ES;
    extern __shared__ __align__(16) uint8_t smem[];
    float4* ring = reinterpret_cast<float4*>(smem);                                  // [kRollRing][kRollPx]
    RowEntry* rows = reinterpret_cast<RowEntry*>(smem + kRollRing * kRollPx * 16);   // [rows_per_seg]
    int* row_y0 = reinterpret_cast<int*>(rows + g.rows_per_seg);                      // [rows_per_seg] first tap row (unclamped)
    uint8_t* obuf = smem + kRollRing * kRollPx * 16 + g.rows_per_seg * (int)sizeof(RowEntry) + ((g.rows_per_seg * 4 + 15) & ~15);   // [G][opitch]

    const int tid = threadIdx.x, px = tid & (kRollPx - 1), half = tid >> 7;
    const int strip = blockIdx.x % g.strips, seg = blockIdx.x / g.strips;
    const int dx0 = strip * kRollPx, tw = min(kRollPx, g.wo - dx0);
    const int dy_begin = seg * g.rows_per_seg, dy_end = min(g.ho, dy_begin + g.rows_per_seg);
    const uint8_t* img = reinterpret_cast<const uint8_t*>(src_) + blockIdx.y * g.src_image * ES;
    uint8_t* out_img = reinterpret_cast<uint8_t*>(dst_) + blockIdx.y * g.dst_image * ES;
    const size_t row_bytes = (size_t)g.w * PX, out_row_bytes = (size_t)g.wo * PX;
    const int seg_bytes = tw * PX;
    const bool active = px < tw;

    // ---- once per CTA: x taps of this column, vertical tables of the segment
    int xidx[4], xcoef[4];
    {
        const int dx = dx0 + (active ? px : tw - 1);
        if (kU8) {
            int s;
            cubic_cv_coord_scaled(dx, g.w, g.scale_x, true, s, xcoef);
#pragma unroll
            for (int j = 0; j < 4; ++j) xidx[j] = min(max(s - 1 + j, 0), g.w - 1);
        } else {
            int ofs; float a[4];
            cubic_naive_scaled(dx, g.w, g.scale_x, ofs, a);
#pragma unroll
            for (int j = 0; j < 4; ++j) { xidx[j] = ofs - 1 + j; xcoef[j] = __float_as_int(a[j]); }
        }
    }
    for (int r = tid; r < dy_end - dy_begin; r += kRollThreads) {
        RowEntry e;
        int y0;
        if (kU8) {
            int s, q[4];
            cubic_cv_coord_scaled(dy_begin + r, g.h, g.scale_y, false, s, q);
            y0 = s - 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) e.b[j] = (float)q[j] * (1.f / (2048 * 2048));
        } else {
            int ofs;
            cubic_naive_scaled(dy_begin + r, g.h, g.scale_y, ofs, e.b);
            y0 = ofs - 1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rr = kU8 ? min(max(y0 + j, 0), g.h - 1) : y0 + j;   // OpenCV clamps the row index; the naive rule is in range by construction
            e.off[j] = (rr & (kRollRing - 1)) * (kRollPx * 16);
        }
        rows[r] = e;
        row_y0[r] = y0;
    }
    const bool consecutive = xidx[1] == xidx[0] + 1 && xidx[2] == xidx[1] + 1 && xidx[3] == xidx[2] + 1;   // false only where taps are clamped
    const int a0 = xidx[0] * PX;               // byte offset of the first tap in a source row
    const int sh = (a0 & 3) * 8;
    // 11-bit coefficients as (signed high byte, unsigned low byte): a = 256*ah + al
    unsigned al = 0; int ah = 0;
    if (kU8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { al |= (unsigned)(xcoef[j] & 0xff) << (8 * j); ah |= ((xcoef[j] >> 8) & 0xff) << (8 * j); }
    }
    const int vec_end = (g.wo * 3) & ~7;       // OpenCV's SSE2 body covers x < (width & ~7)
    const int e_first = (dx0 + px) * 3;
    const bool has_tail = kU8 && e_first + 2 >= vec_end;
    // 16-byte phase of output row r of the segment = (mis0 + r * row_step) & 15
    const unsigned mis0 = (unsigned)((reinterpret_cast<uintptr_t>(out_img) + (size_t)dy_begin * out_row_bytes + (size_t)dx0 * PX) & 15);
    const unsigned row_step = (unsigned)(out_row_bytes & 15);
    __syncthreads();

    auto clamp_row = [&](int r) { return kU8 ? min(max(r, 0), g.h - 1) : r; };

    // Pass 1, u8 fast path (unclamped columns), software-pipelined: the 32-bit words of up to NB rows per thread are
    // loaded into registers one group AHEAD (prefetch), i.e. their DRAM/L2 latency overlaps the previous group's
    // pass 2 (ncu showed pass 1 latency-bound on long_scoreboard otherwise); consume() turns them into ring rows.
    constexpr int NB = 4;
    uint32_t pf[NB][4];
    int pf_first = 0, pf_hi = -1;   // prefetched rows: pf_first + half + 2*i <= pf_hi
    auto prefetch = [&](int first_new, int hi) {
        pf_first = first_new; pf_hi = hi;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int rr = min(first_new + half + 2 * i, hi);   // past-the-end slots reload row hi and are never consumed
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (size_t)max(rr, 0) * row_bytes + (a0 & ~3));
            pf[i][0] = __ldg(wp); pf[i][1] = __ldg(wp + 1); pf[i][2] = __ldg(wp + 2); pf[i][3] = sh ? __ldg(wp + 3) : 0u;
        }
    };
    auto hsum_store = [&](int r, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
        const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
        // 12 bytes  b0 = [t0.b t0.g t0.r t1.b]  b1 = [t1.g t1.r t2.b t2.g]  b2 = [t2.r t3.b t3.g t3.r]  -> one word per channel
        const unsigned cb = __byte_perm(__byte_perm(b0, b1, 0x0630), b2, 0x5210);   // t0.b t1.b t2.b t3.b
        const unsigned cg = __byte_perm(__byte_perm(b0, b1, 0x0741), b2, 0x6210);   // t0.g t1.g t2.g t3.g
        const unsigned cr = __byte_perm(__byte_perm(b0, b1, 0x0052), b2, 0x7410);   // t0.r t1.r t2.r t3.r
        // sum(tap * (256*ah + al)) + bit pattern of 1.5*2^23; then exact int -> float for |H| < 2^22
        const int hb = dp4a_us(cb, ah, 0) * 256 + (int)__dp4a(cb, al, 0x4B400000u);
        const int hg = dp4a_us(cg, ah, 0) * 256 + (int)__dp4a(cg, al, 0x4B400000u);
        const int hr = dp4a_us(cr, ah, 0) * 256 + (int)__dp4a(cr, al, 0x4B400000u);
        ring[(r & (kRollRing - 1)) * kRollPx + px] =
            make_float4(__int_as_float(hb) - 12582912.0f, __int_as_float(hg) - 12582912.0f, __int_as_float(hr) - 12582912.0f, 0.f);
    };
    auto consume = [&]() {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int r = pf_first + half + 2 * i;
            if (r <= pf_hi) hsum_store(r, pf[i][0], pf[i][1], pf[i][2], pf[i][3]);
        }
        for (int r = pf_first + half + 2 * NB; r <= pf_hi; r += 2) {   // more new rows than the prefetch holds (first group of a segment)
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (size_t)r * row_bytes + (a0 & ~3));
            hsum_store(r, __ldg(wp), __ldg(wp + 1), __ldg(wp + 2), sh ? __ldg(wp + 3) : 0u);
        }
    };
    auto hfilter = [&](int r) {   // one row: clamped (edge) columns of the u8 kind, and the fp32 kind
        const uint8_t* rowp = img + (size_t)r * row_bytes;
        float4 hv;
        if (kU8) {
            int h[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                h[k] = __ldg(rowp + xidx[0] * 3 + k) * xcoef[0] + __ldg(rowp + xidx[1] * 3 + k) * xcoef[1] +
                       __ldg(rowp + xidx[2] * 3 + k) * xcoef[2] + __ldg(rowp + xidx[3] * 3 + k) * xcoef[3];
            hv = make_float4(__int_as_float(h[0] + 0x4B400000) - 12582912.0f, __int_as_float(h[1] + 0x4B400000) - 12582912.0f,
                             __int_as_float(h[2] + 0x4B400000) - 12582912.0f, 0.f);
        } else {
            const float* fp = reinterpret_cast<const float*>(rowp) + xidx[0] * 3;   // taps are consecutive on this path
            const float a_0 = __int_as_float(xcoef[0]), a_1 = __int_as_float(xcoef[1]), a_2 = __int_as_float(xcoef[2]), a_3 = __int_as_float(xcoef[3]);
            float t[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) t[i] = __ldg(fp + i);
            float h[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)   // resize_naive.cpp:230 order
                h[k] = t[k] * a_0 + t[3 + k] * a_1 + t[6 + k] * a_2 + t[9 + k] * a_3;
            hv = make_float4(h[0], h[1], h[2], 0.f);
        }
        ring[(r & (kRollRing - 1)) * kRollPx + px] = hv;
    };

    const bool fast = kU8 && consecutive && active;
    auto group_rows = [&](int dy0, int& lo, int& hi) {
        const int gth = min(g.G, dy_end - dy0);
        lo = clamp_row(row_y0[dy0 - dy_begin]);
        hi = clamp_row(row_y0[dy0 + gth - 1 - dy_begin] + 3);
    };
    int y_done = INT_MIN;   // highest source row already in the ring (or prefetched)
    int lo, hi;
    group_rows(dy_begin, lo, hi);
    if (fast) prefetch(lo, hi);
    for (int dy0 = dy_begin; dy0 < dy_end; dy0 += g.G) {
        const int gth = min(g.G, dy_end - dy0);
        // ---- pass 1: source rows [lo, hi] of this group that are not in the ring yet
        const int first_new = max(lo, y_done + 1);
        if (fast) consume();
        else if (active) for (int r = first_new + half; r <= hi; r += 2) hfilter(r);
        y_done = hi;
        __syncthreads();
        if (dy0 + g.G < dy_end) {   // next group's rows: start their loads now, they land while pass 2 runs
            group_rows(dy0 + g.G, lo, hi);
            if (fast) prefetch(max(lo, y_done + 1), hi);
        }
        // ---- pass 2: the group's output rows into obuf
        if (active) {
            const uint8_t* rp = reinterpret_cast<const uint8_t*>(ring) + 16 * px;
            for (int ty = half; ty < gth; ty += 2) {
                const RowEntry* e = rows + (dy0 + ty - dy_begin);
                const int4 off = *reinterpret_cast<const int4*>(e->off);
                const float4 bw = *reinterpret_cast<const float4*>(e->b);
                const float4 h0 = *reinterpret_cast<const float4*>(rp + off.x), h1 = *reinterpret_cast<const float4*>(rp + off.y);
                const float4 h2 = *reinterpret_cast<const float4*>(rp + off.z), h3 = *reinterpret_cast<const float4*>(rp + off.w);
                const float c0[3] = {h0.x, h0.y, h0.z}, c1[3] = {h1.x, h1.y, h1.z}, c2[3] = {h2.x, h2.y, h2.z}, c3[3] = {h3.x, h3.y, h3.z};
                const unsigned mis = (mis0 + (unsigned)(dy0 + ty - dy_begin) * row_step) & 15u;
                S* o = reinterpret_cast<S*>(obuf + ty * g.opitch + mis) + 3 * px;
                if (!kU8) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)   // resize_naive.cpp:345
                        o[k] = (S)(c0[k] * bw.x + c1[k] * bw.y + c2[k] * bw.z + c3[k] * bw.w);
                } else if (!has_tail) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {   // fp32 body: mul, add, one rounding each; cvtps2dq (half-even); packs; packus
                        float f = c0[k] * bw.x;
                        f = f + c1[k] * bw.y;
                        f = f + c2[k] * bw.z;
                        f = f + c3[k] * bw.w;
                        // |f| < 2^22: adding 1.5*2^23 rounds half-to-even at integer granularity; subtracting its bit pattern and
                        // clamping to [0,255] is one DPX op (the intermediate s16 saturation of packs cannot change the result)
                        o[k] = (S)__viaddmin_s32_relu(__float_as_int(f + 12582912.0f), -0x4B400000, 255);
                    }
                } else {                            // this pixel reaches into the last (width & 7) elements of the row
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        int v;
                        if (e_first + k < vec_end) {
                            float f = c0[k] * bw.x;
                            f = f + c1[k] * bw.y;
                            f = f + c2[k] * bw.z;
                            f = f + c3[k] * bw.w;
                            v = __float_as_int(f + 12582912.0f) - 0x4B400000;
                        } else {                    // scalar tail: FixedPtCast<int, uchar, 22>; ibeta = b * 2^22 exactly
                            v = (__float2int_rn(c0[k]) * __float2int_rn(bw.x * 4194304.f) + __float2int_rn(c1[k]) * __float2int_rn(bw.y * 4194304.f) +
                                 __float2int_rn(c2[k]) * __float2int_rn(bw.z * 4194304.f) + __float2int_rn(c3[k]) * __float2int_rn(bw.w * 4194304.f) + (1 << 21)) >> 22;
                        }
                        o[k] = (S)clamp255(v);
                    }
                }
            }
        }
        __syncthreads();
        // ---- store the group: 16-byte aligned chunks (the next group's pass 1 overlaps with this)
        const int chunks_per_row = (seg_bytes + 15 + 15) >> 4;
        for (int i = tid; i < gth * chunks_per_row; i += kRollThreads) {
            const int ty = i / chunks_per_row, q = i - ty * chunks_per_row;
            uint8_t* grow = out_img + (size_t)(dy0 + ty) * out_row_bytes + (size_t)dx0 * PX;
            const int mis = (int)(reinterpret_cast<uintptr_t>(grow) & 15);
            const int lo_b = max(mis, 16 * q), hi_b = min(mis + seg_bytes, 16 * q + 16);
            if (lo_b >= hi_b) continue;
            const uint8_t* sp = obuf + ty * g.opitch + 16 * q;
            uint8_t* gp = grow - mis + 16 * q;
            if (hi_b - lo_b == 16) st_stream16(gp, *reinterpret_cast<const uint4*>(sp));
            else for (int b = lo_b - 16 * q; b < hi_b - 16 * q; ++b) gp[b] = sp[b];
        }
        // no barrier here: the next iteration's pass 1 touches only the ring, and its barrier orders these obuf reads
        // before the next pass 2 writes obuf
    }
}

}  // namespace vacv
