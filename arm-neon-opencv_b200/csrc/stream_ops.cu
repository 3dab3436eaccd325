// a2 crop, a3 layout_change, a4 dtype_change, a12 normalize -- the pure streaming operators.
// All are HBM-bound byte movers: the design rule is "every global access a warp issues is lane-contiguous and
// as wide as the alignment allows (128-bit in the common shapes)".
#include <algorithm>
#include <cstdlib>

#include "host_util.cuh"
#include "vacv_common.cuh"

namespace vacv {

// =====================================================================================================
// a2 crop (src/cv/crop.cpp:44-142).  Both layouts reduce to "copy R byte-rows of RB bytes": HWC rows are
// cw*c*elem bytes, CHW rows are cw*elem bytes for each of c planes.  One warp per destination row.  The
// destination is written in 16-byte aligned chunks; the (generally misaligned) source bytes for a chunk come
// from the two aligned 16-byte chunks around them (128-bit loads; L1 serves the overlap between neighbouring lanes),
// merged with funnel shifts.  Aligned chunks that hold valid bytes never leave the allocation (allocations are 256-byte
// granular).  (A pure copy-engine variant -- cp.async.bulk.tensor load + store through a 4-stage shared-memory ring, one
// thread per CTA -- was measured on 16-byte aligned crops: 0.150 ms against 0.134 ms for this kernel on 128 x 1080p ->
// 1280x720, and TMA cannot shift bytes, so misaligned left edges would need this kernel anyway; it was dropped.)
// Division by a launch constant without the 20-instruction runtime-divisor sequence: q = (n * mul) >> sh, exact for n < 2^31.
struct FastDiv {
    unsigned mul, sh, d;
    void init(unsigned div) {
        d = div;
        unsigned l = 0;
        while ((1ull << l) < div) ++l;
        sh = 31 + l;
        mul = (unsigned)(((1ull << sh) + div - 1) / div);
    }
    __device__ __forceinline__ unsigned div(unsigned n) const { return (unsigned)(((unsigned long long)n * mul) >> sh); }
};

struct CropGeom {
    int rows_per_frame;   // ch (HWC) or c*ch (CHW)
    int ch;               // rows per plane
    int RB;               // destination row bytes
    size_t src_frame, src_plane, src_pitch, src_ofs;   // bytes
    FastDiv by_rows_per_frame, by_ch, by_segs;
    int segs;             // warps per destination row (long rows are split so that no warp loops)
    int cpr;              // kFlat: 16-byte chunks per destination row
    FastDiv by_cpr;
};

constexpr int kCropU = 4;                      // 16-byte chunks per lane: all loads of a warp are issued before its first store
constexpr int kCropSeg = 32 * kCropU;          // chunks per warp

// One warp per (destination row, segment of 128 chunks).  Round 1's kernel gave a whole row to one warp, looping chunk by chunk
// with one load pair in flight per lane and three runtime divisions per thread: fp32 rows (15 KB) were latency-bound and the
// short CHW plane rows (1.3 KB) instruction-bound (bench_ops: 0.61 / 0.65 of the copy peak against 0.81 for u8 HWC rows).
__global__ void __launch_bounds__(256) crop_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                         CropGeom g, unsigned total_warps) {
    const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= total_warps) return;
    const int lane = threadIdx.x & 31;
    const unsigned row = g.segs == 1 ? wid : g.by_segs.div(wid);
    const int seg = (int)(wid - row * g.segs);
    const unsigned frame = g.by_rows_per_frame.div(row);
    const unsigned rr = row - frame * g.rows_per_frame;
    const unsigned plane = g.by_ch.div(rr), y = rr - plane * g.ch;
    const uint8_t* s = src + frame * g.src_frame + plane * g.src_plane + (size_t)y * g.src_pitch + g.src_ofs;
    uint8_t* d = dst + (size_t)row * (size_t)g.RB;

    const int head = min((int)((16 - ((uintptr_t)d & 15)) & 15), g.RB);
    if (seg == 0 && lane < head) d[lane] = __ldg(s + lane);
    const int nchunks = (g.RB - head) >> 4;
    const uint8_t* sb = s + head;
    uint8_t* db = d + head;
    // 16-byte granular: output chunk q = source bytes [sb + 16 q, +16) = words k .. k+4 (shifted by 8 (m & 3) bits) of the two
    // aligned 16-byte chunks around it; both come in as 128-bit loads (the second one is the next lane's first: an L1 hit).
    // m is the same for every chunk of the row.
    const int m = (int)((uintptr_t)sb & 15), k = m >> 2, sh = 8 * (m & 3);
    const uint4* sa = reinterpret_cast<const uint4*>(sb - m);
    // never read past the 16-byte chunk that holds the last source byte of the row (it may be the last of the allocation)
    const uint4* last = reinterpret_cast<const uint4*>((uintptr_t)(s + g.RB - 1) & ~(uintptr_t)15);
    const int q0 = seg * kCropSeg + lane;
    uint4 a[kCropU], b4[kCropU];
#pragma unroll
    for (int j = 0; j < kCropU; ++j) {
        const int q = q0 + 32 * j;
        a[j] = q < nchunks ? ld_stream16(sa + q) : make_uint4(0u, 0u, 0u, 0u);
        b4[j] = (q < nchunks && m != 0 && sa + q + 1 <= last) ? __ldg(sa + q + 1) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < kCropU; ++j) {
        const int q = q0 + 32 * j;
        if (q >= nchunks) break;
        uint32_t w0, w1, w2, w3, w4;
        switch (k) {
            case 0: w0 = a[j].x; w1 = a[j].y; w2 = a[j].z; w3 = a[j].w; w4 = b4[j].x; break;
            case 1: w0 = a[j].y; w1 = a[j].z; w2 = a[j].w; w3 = b4[j].x; w4 = b4[j].y; break;
            case 2: w0 = a[j].z; w1 = a[j].w; w2 = b4[j].x; w3 = b4[j].y; w4 = b4[j].z; break;
            default: w0 = a[j].w; w1 = b4[j].x; w2 = b4[j].y; w3 = b4[j].z; w4 = b4[j].w; break;
        }
        st_stream16(db + 16 * q, make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                                            __funnelshift_r(w3, w4, sh)));
    }
    const int done = head + 16 * nchunks;
    if (seg == g.segs - 1 && lane < g.RB - done) d[done + lane] = __ldg(s + done + lane);
}

// Rows that are whole 16-byte chunks in an aligned destination (the usual case) need no head / tail handling, so the chunks of ALL
// rows form one flat sequence and a warp simply takes the next 128 of them -- short rows (CHW plane rows: 80 chunks) no longer
// leave lanes idle.  A lane's chunks may belong to different rows, so row, source pointer and byte phase are per lane.
__global__ void __launch_bounds__(256) crop_flat_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, CropGeom g,
                                                         unsigned total_chunks) {
    const unsigned base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kCropSeg + (threadIdx.x & 31);
    uint4 a[kCropU], b4[kCropU];
    int m[kCropU];
#pragma unroll
    for (int j = 0; j < kCropU; ++j) {
        const unsigned item = base + 32 * j;
        a[j] = b4[j] = make_uint4(0u, 0u, 0u, 0u);
        m[j] = 0;
        if (item >= total_chunks) continue;
        const unsigned row = g.by_cpr.div(item), q = item - row * g.cpr;
        const unsigned frame = g.by_rows_per_frame.div(row);
        const unsigned rr = row - frame * g.rows_per_frame;
        const unsigned plane = g.by_ch.div(rr), y = rr - plane * g.ch;
        const uint8_t* s = src + frame * g.src_frame + plane * g.src_plane + (size_t)y * g.src_pitch + g.src_ofs;
        m[j] = (int)((uintptr_t)s & 15);
        const uint4* sa = reinterpret_cast<const uint4*>(s - m[j]) + q;
        a[j] = ld_stream16(sa);
        // the chunk behind holds bytes of this row unless this is the row's last chunk and the row ends inside `a`
        if (m[j] != 0) b4[j] = __ldg(sa + 1);   // m != 0: the row's last byte lies in the chunk after `a`, never past it
    }
#pragma unroll
    for (int j = 0; j < kCropU; ++j) {
        const unsigned item = base + 32 * j;
        if (item >= total_chunks) break;
        const int k = m[j] >> 2, sh = 8 * (m[j] & 3);
        const uint32_t w[8] = {a[j].x, a[j].y, a[j].z, a[j].w, b4[j].x, b4[j].y, b4[j].z, b4[j].w};
        uint32_t v[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = k == 0 ? w[i] : k == 1 ? w[i + 1] : k == 2 ? w[i + 2] : w[i + 3];
        st_stream16(dst + (size_t)item * 16, make_uint4(__funnelshift_r(v[0], v[1], sh), __funnelshift_r(v[1], v[2], sh),
                                                        __funnelshift_r(v[2], v[3], sh), __funnelshift_r(v[3], v[4], sh)));
    }
}

// =====================================================================================================
// a3 layout_change (src/common/tensor.cpp:160-182): chw[k*wh + j] = hwc[j*c + k].
// c in {2, 3, 4} fast path: a thread moves V = 16/sizeof(T) pixels: 16*c contiguous bytes on the interleaved side
// (c 128-bit accesses; neighbouring lanes complete each other's sectors) and one 128-bit access per plane.
template <typename T, int C>
__device__ __forceinline__ void deinterleave(const uint4 (&in)[C], uint4 (&out)[C]) {
    constexpr int V = 16 / sizeof(T);
    const T* a = reinterpret_cast<const T*>(in);
    T* o = reinterpret_cast<T*>(out);
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
        for (int k = 0; k < C; ++k) o[k * V + i] = a[C * i + k];
}
template <typename T, int C>
__device__ __forceinline__ void interleave(const uint4 (&in)[C], uint4 (&out)[C]) {
    constexpr int V = 16 / sizeof(T);
    const T* a = reinterpret_cast<const T*>(in);
    T* o = reinterpret_cast<T*>(out);
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
        for (int k = 0; k < C; ++k) o[C * i + k] = a[k * V + i];
}

// groups = batch * wh / V ; wh % V == 0
template <typename T, int C, bool kToCHW>
__global__ void __launch_bounds__(256) layout_cn_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                         size_t groups, size_t groups_per_frame) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const size_t frame = g / groups_per_frame, j = g % groups_per_frame;
    const size_t inter = frame * groups_per_frame * C + j * C;      // uint4 index on the HWC side
    const size_t planar = frame * groups_per_frame * C + j;         // uint4 index of plane 0 on the CHW side
    uint4 a[C], b[C];
    if (kToCHW) {
#pragma unroll
        for (int k = 0; k < C; ++k) a[k] = __ldg(src + inter + k);
        deinterleave<T, C>(a, b);
#pragma unroll
        for (int k = 0; k < C; ++k) st_stream16(dst + planar + k * groups_per_frame, b[k]);
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) a[k] = ld_stream16(src + planar + k * groups_per_frame);
        interleave<T, C>(a, b);
#pragma unroll
        for (int k = 0; k < C; ++k) st_stream16(dst + inter + k, b[k]);
    }
}

template <typename T, bool kToCHW>
__global__ void layout_generic_kernel(const T* __restrict__ src, T* __restrict__ dst, int wh, int c, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // index on the destination side
    if (i >= total) return;
    const size_t per = (size_t)wh * c;
    const size_t frame = i / per, r = i % per;
    size_t s;
    if (kToCHW) { size_t k = r / wh, j = r % wh; s = j * c + k; }
    else        { size_t j = r / c, k = r % c;   s = k * wh + j; }
    dst[i] = src[frame * per + s];
}

// =====================================================================================================
// a4 dtype_change (src/common/tensor.cpp:459-502)
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n4) {
    // 4 elements per thread and step: 32-bit load (128 B / warp), 128-bit store (512 B / warp); kU independent loads in flight
    // before the first store (with one load per iteration the loop waits out a DRAM latency per 16 output bytes)
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int kU = 4;
    for (; i < n4; i += kU * stride) {
        uint32_t v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) v[u] = i + u * stride < n4 ? ld_stream4(src + 4 * (i + u * stride)) : 0u;
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (i + u * stride < n4)
                st_stream16f(dst + 4 * (i + u * stride), make_float4((float)(v[u] & 0xff), (float)((v[u] >> 8) & 0xff), (float)((v[u] >> 16) & 0xff),
                                                                     (float)(v[u] >> 24)));
    }
}
__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int kU = 4;
    for (; i < n4; i += kU * stride) {
        uint4 r[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) r[u] = i + u * stride < n4 ? ld_stream16(src + 4 * (i + u * stride)) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (i + u * stride >= n4) continue;
            // static_cast<char>(float) on x86 = cvttss2si then low byte (tensor.cpp:488-492)
            uint32_t b0 = (uint32_t)(int)__uint_as_float(r[u].x) & 0xff, b1 = (uint32_t)(int)__uint_as_float(r[u].y) & 0xff;
            uint32_t b2 = (uint32_t)(int)__uint_as_float(r[u].z) & 0xff, b3 = (uint32_t)(int)__uint_as_float(r[u].w) & 0xff;
            st_stream4(dst + 4 * (i + u * stride), b0 | (b1 << 8) | (b2 << 16) | (b3 << 24));
        }
    }
}
__global__ void dtype_tail_kernel(const void* src, void* dst, size_t begin, size_t n, int to_f32) {
    size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (to_f32) ((float*)dst)[i] = (float)((const uint8_t*)src)[i];
    else ((uint8_t*)dst)[i] = (uint8_t)(int)((const float*)src)[i];
}

// =====================================================================================================
// a12 normalize (src/cv/normalize_naive.cpp:74-90): (float)((double)(x - mean) / ((double)std + 1e-6)).
//
// u8 input: only 256 inputs per channel exist, so each CTA builds the 256 x c table once with that exact expression
//   and the stream becomes  4-byte load -> 4 shared-memory lookups -> 16-byte store.
// fp32 input: q = (double)(x - mean) * (1/den) is within 2 ulp(double) of the correctly rounded quotient, so
//   (float)q equals the reference's (float)(d/den) unless q sits within a few double-ulps of a float rounding
//   boundary; exactly those (about 1 in 2^25) are redone with the IEEE double division.  Bit-exact, 1 DMUL per element.
//
// Channel of an element: HWC -> (index mod c) tracked incrementally (no per-element division);
//                        CHW -> one plane per blockIdx.y, the channel is a CTA constant.
constexpr int kNormMaxC = 4;

__device__ __forceinline__ float normalize_fast_exact(float x, float mean, double den, double rden) {
    const double d = (double)(x - mean);
    const double q = d * rden;
    // low 29 bits of the fp64 mantissa are what the fp32 rounding discards; 0x10000000 is the midpoint
    const unsigned lo = (unsigned)__double2loint(q) & 0x1fffffffu;
    const float f = (float)q;
    // redo exactly when q is within 8 double-ulps of an fp32 rounding boundary (|q - exact| <= 2.5 ulp), or when
    // the fp32 result is (near) subnormal, where the boundary sits at a different bit
    if (__builtin_expect(lo - 0x0ffffff8u <= 16u || fabsf(f) < 1e-30f, 0)) return (float)(d / den);
    return f;
}

struct NormGeom {
    int c;
    unsigned wh;            // pixels per plane
    unsigned per_frame;     // wh * c
    int stats_per_frame;
};

// ---- HWC.  grid = (ctas_per_frame, frames).  C = channel count (1..4).
template <int C>
__global__ void __launch_bounds__(256) normalize_u8_hwc_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                                const float* __restrict__ mean, const float* __restrict__ stddev,
                                                                NormGeom g) {
    __shared__ float lut[C * 256];
    const size_t frame = blockIdx.y;
    const float* mu = mean + (g.stats_per_frame ? frame * C : 0);
    const float* sd = stddev + (g.stats_per_frame ? frame * C : 0);
    for (int t = threadIdx.x; t < 256 * C; t += blockDim.x)
        lut[t] = normalize_one((float)(t & 255), mu[t >> 8], (double)sd[t >> 8] + 1e-6);
    __syncthreads();
    const uint8_t* s = src + frame * g.per_frame;
    float* d = dst + frame * g.per_frame;
    const unsigned n4 = g.per_frame >> 2, stride = gridDim.x * blockDim.x;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    // element 4*i has channel (4*i) mod C; advance it by (4*stride) mod C per iteration
    unsigned ph = (4u * (i % C)) % C;
    const unsigned dph = (4u * (stride % C)) % C;
    // kU independent 4-byte loads are in flight per thread before the first table lookup: with one load per iteration the
    // kernel was latency-bound (ncu: long_scoreboard 47 stall cycles per issue, DRAM at 57 %)
    constexpr int kU = 4;
    for (; i < n4; i += kU * stride) {
        uint32_t v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) v[u] = i + u * stride < n4 ? ld_stream4(s + 4 * (size_t)(i + u * stride)) : 0u;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (i + u * stride < n4) {
                float o[4];
                unsigned k = ph;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    o[j] = lut[k * 256 + ((v[u] >> (8 * j)) & 0xff)];
                    k = (k + 1 == C) ? 0 : k + 1;
                }
                st_stream16f(d + 4 * (size_t)(i + u * stride), make_float4(o[0], o[1], o[2], o[3]));
            }
            ph += dph;
            if (ph >= C) ph -= C;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < g.per_frame - 4 * n4) {
        const unsigned e = 4 * n4 + threadIdx.x;
        d[e] = lut[(e % C) * 256 + s[e]];
    }
}

template <int C>
__global__ void __launch_bounds__(256) normalize_f32_hwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                 const float* __restrict__ mean, const float* __restrict__ stddev,
                                                                 NormGeom g) {
    const size_t frame = blockIdx.y;
    float mu[C]; double den[C], rden[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        mu[k] = __ldg(mean + (g.stats_per_frame ? frame * C : 0) + k);
        den[k] = (double)__ldg(stddev + (g.stats_per_frame ? frame * C : 0) + k) + 1e-6;
        rden[k] = 1.0 / den[k];
    }
    const float* s = src + frame * g.per_frame;
    float* d = dst + frame * g.per_frame;
    const unsigned n4 = g.per_frame >> 2, stride = gridDim.x * blockDim.x;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned ph = (4u * (i % C)) % C;
    const unsigned dph = (4u * (stride % C)) % C;
    constexpr int kU = 2;   // two independent 16-byte loads in flight per thread
    for (; i < n4; i += kU * stride) {
        uint4 rr[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) rr[u] = i + u * stride < n4 ? ld_stream16(s + 4 * (size_t)(i + u * stride)) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (i + u * stride < n4) {
                const float x[4] = {__uint_as_float(rr[u].x), __uint_as_float(rr[u].y), __uint_as_float(rr[u].z), __uint_as_float(rr[u].w)};
                float o[4];
                unsigned k = ph;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // select the channel's constants without dynamic register indexing
                    float m = mu[0]; double dn = den[0], rd = rden[0];
#pragma unroll
                    for (int q = 1; q < C; ++q) if (k == q) { m = mu[q]; dn = den[q]; rd = rden[q]; }
                    o[j] = normalize_fast_exact(x[j], m, dn, rd);
                    k = (k + 1 == C) ? 0 : k + 1;
                }
                st_stream16f(d + 4 * (size_t)(i + u * stride), make_float4(o[0], o[1], o[2], o[3]));
            }
            ph += dph;
            if (ph >= C) ph -= C;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < g.per_frame - 4 * n4) {
        const unsigned e = 4 * n4 + threadIdx.x, k = e % C;
        d[e] = normalize_one(s[e], mean[(g.stats_per_frame ? frame * C : 0) + k], (double)stddev[(g.stats_per_frame ? frame * C : 0) + k] + 1e-6);
    }
}

// ---- planes (CHW, or c == 1).  grid = (ctas_per_plane, planes); plane p = frame p / c, channel p % c.
template <bool kU8>
__global__ void __launch_bounds__(256) normalize_plane_kernel(const void* __restrict__ src_, float* __restrict__ dst,
                                                               const float* __restrict__ mean, const float* __restrict__ stddev,
                                                               NormGeom g) {
    __shared__ float lut[256];
    const unsigned plane = blockIdx.y, frame = plane / g.c, k = plane - frame * g.c;
    const float mu = __ldg(mean + (g.stats_per_frame ? frame * g.c : 0) + k);
    const double den = (double)__ldg(stddev + (g.stats_per_frame ? frame * g.c : 0) + k) + 1e-6;
    const size_t base = (size_t)plane * g.wh;
    float* d = dst + base;
    // vector body needs the plane start 16-byte (f32) / 4-byte (u8) aligned: true when wh % 4 == 0, else scalar
    const bool vec = (g.wh & 3) == 0;
    const unsigned n4 = vec ? g.wh >> 2 : 0, stride = gridDim.x * blockDim.x;
    if (kU8) {
        const uint8_t* s = (const uint8_t*)src_ + base;
        for (int t = threadIdx.x; t < 256; t += blockDim.x) lut[t] = normalize_one((float)t, mu, den);
        __syncthreads();
        // four independent loads in flight per thread before the first table lookup (same reason as in the HWC kernel)
        constexpr int kU = 4;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += kU * stride) {
            uint32_t v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) v[u] = i + u * stride < n4 ? ld_stream4(s + 4 * (size_t)(i + u * stride)) : 0u;
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (i + u * stride < n4)
                    st_stream16f(d + 4 * (size_t)(i + u * stride),
                                 make_float4(lut[v[u] & 0xff], lut[(v[u] >> 8) & 0xff], lut[(v[u] >> 16) & 0xff], lut[v[u] >> 24]));
        }
        for (unsigned e = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; e < g.wh; e += stride) d[e] = lut[s[e]];
    } else {
        const float* s = (const float*)src_ + base;
        const double rden = 1.0 / den;
        constexpr int kU = 2;   // two 16-byte loads in flight per thread
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += kU * stride) {
            uint4 r[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) r[u] = i + u * stride < n4 ? ld_stream16(s + 4 * (size_t)(i + u * stride)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (i + u * stride < n4)
                    st_stream16f(d + 4 * (size_t)(i + u * stride),
                                 make_float4(normalize_fast_exact(__uint_as_float(r[u].x), mu, den, rden), normalize_fast_exact(__uint_as_float(r[u].y), mu, den, rden),
                                             normalize_fast_exact(__uint_as_float(r[u].z), mu, den, rden), normalize_fast_exact(__uint_as_float(r[u].w), mu, den, rden)));
        }
        for (unsigned e = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; e < g.wh; e += stride) d[e] = normalize_one(s[e], mu, den);
    }
}

}  // namespace vacv

using namespace vacv;

extern "C" int vacv_cuda_crop(const void* src, void* dst, int batch, int w, int h, int c, int dtype, int layout,
                              int left, int top, int cw, int ch, void* stream) {
    VACV_REQUIRE(src && dst, "crop: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0 && cw > 0 && ch > 0, "crop: non-positive size");
    if (dtype != VACV_INT8 && dtype != VACV_FP32) return set_error(VACV_ERR_UNSUPPORTED, "crop: dtype %d (reference: INT8/FP32 only, crop.cpp:133)", dtype);
    VACV_REQUIRE(left >= 0 && top >= 0 && left + cw <= w && top + ch <= h,
                 "crop: rect (%d,%d %dx%d) outside %dx%d frame", left, top, cw, ch, w, h);
    const size_t es = elem_size(dtype);
    CropGeom g;
    if (layout == VACV_NHWC) {
        g.rows_per_frame = ch; g.ch = ch; g.RB = (int)(cw * c * es);
        g.src_plane = 0; g.src_pitch = (size_t)w * c * es; g.src_ofs = ((size_t)top * w + left) * c * es;
    } else {
        g.rows_per_frame = c * ch; g.ch = ch; g.RB = (int)(cw * es);
        g.src_plane = (size_t)w * h * es; g.src_pitch = (size_t)w * es; g.src_ofs = ((size_t)top * w + left) * es;
    }
    g.src_frame = (size_t)w * h * c * es;
    const size_t rows = (size_t)batch * g.rows_per_frame;
    g.segs = std::max(1, (g.RB / 16 + kCropSeg - 1) / kCropSeg);
    VACV_REQUIRE(rows * g.segs < 0x7fffffffull / 32, "crop: too many rows for one launch");
    g.by_rows_per_frame.init((unsigned)g.rows_per_frame);
    g.by_ch.init((unsigned)g.ch);
    g.by_segs.init((unsigned)g.segs);
    if ((g.RB % 16) == 0 && ((uintptr_t)dst % 16) == 0 && rows * (size_t)(g.RB / 16) < 0x7fffffffull) {   // flat chunk sequence
        g.cpr = g.RB / 16;
        g.by_cpr.init((unsigned)g.cpr);
        const unsigned chunks = (unsigned)(rows * g.cpr);
        crop_flat_kernel<<<ceil_div(ceil_div(chunks, kCropSeg) * (size_t)32, 256), 256, 0, as_stream(stream)>>>((const uint8_t*)src, (uint8_t*)dst, g, chunks);
        return check_launch("crop");
    }
    const unsigned warps = (unsigned)(rows * g.segs);
    crop_rows_kernel<<<ceil_div((size_t)warps * 32, 256), 256, 0, as_stream(stream)>>>((const uint8_t*)src, (uint8_t*)dst, g, warps);
    return check_launch("crop");
}

template <typename T, int C>
static void launch_layout_cn(const void* src, void* dst, size_t groups, size_t gpf, bool to_chw, cudaStream_t s) {
    if (to_chw) layout_cn_kernel<T, C, true><<<ceil_div(groups, 256), 256, 0, s>>>((const uint4*)src, (uint4*)dst, groups, gpf);
    else layout_cn_kernel<T, C, false><<<ceil_div(groups, 256), 256, 0, s>>>((const uint4*)src, (uint4*)dst, groups, gpf);
}

template <typename T>
static void launch_layout(const void* src, void* dst, int batch, int wh, int c, bool to_chw, cudaStream_t s) {
    constexpr int V = 16 / sizeof(T);
    const bool vec = c >= 2 && c <= 4 && (wh % V) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    if (vec) {
        const size_t gpf = wh / V, groups = gpf * batch;
        if (c == 2) launch_layout_cn<T, 2>(src, dst, groups, gpf, to_chw, s);
        else if (c == 3) launch_layout_cn<T, 3>(src, dst, groups, gpf, to_chw, s);
        else launch_layout_cn<T, 4>(src, dst, groups, gpf, to_chw, s);
    } else {
        const size_t total = (size_t)batch * wh * c;
        if (to_chw) layout_generic_kernel<T, true><<<ceil_div(total, 256), 256, 0, s>>>((const T*)src, (T*)dst, wh, c, total);
        else layout_generic_kernel<T, false><<<ceil_div(total, 256), 256, 0, s>>>((const T*)src, (T*)dst, wh, c, total);
    }
}

extern "C" int vacv_cuda_layout_change(const void* src, void* dst, int batch, int w, int h, int c, int dtype,
                                       int from_layout, int to_layout, void* stream) {
    VACV_REQUIRE(src && dst, "layout_change: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0, "layout_change: non-positive size");
    const size_t es = elem_size(dtype);
    if (es != 1 && es != 2 && es != 4) return set_error(VACV_ERR_UNSUPPORTED, "layout_change: dtype %d", dtype);
    cudaStream_t s = as_stream(stream);
    if (c == 1 || from_layout == to_layout) {   // tensor.cpp:398-400: clone
        cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)batch * w * h * c * es, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "layout_change: %s", cudaGetErrorString(e));
        return VACV_OK;
    }
    const bool to_chw = to_layout == VACV_NCHW;
    if (es == 1) launch_layout<uint8_t>(src, dst, batch, w * h, c, to_chw, s);
    else if (es == 2) launch_layout<uint16_t>(src, dst, batch, w * h, c, to_chw, s);
    else launch_layout<uint32_t>(src, dst, batch, w * h, c, to_chw, s);
    return check_launch("layout_change");
}

// CTAs (of 256 threads) for a streaming kernel over n4 16-byte groups: MANY SHORT CTAs.  Round 1 sized these grids as ~16 resident CTAs per
// SM grid-striding over everything (19 CTAs per 4K frame, 2 432 in all for config 5's apply pass); measured on the same box, u8 HWC
// normalize 4K x128: 19 CTAs per frame 2.958 ms, 152: 2.660, 1216: 2.509 (6.35 TB/s, 0.97 of the copy peak), 3037: 2.52 -- a grid of two
// long waves leaves the machine half empty at the end and keeps every CTA of a frame in lockstep on the same DRAM pages.
static unsigned stream_ctas(size_t n4) {
    const int qpt = knob(kKnobStreamQpt) > 0 ? knob(kKnobStreamQpt) : 8;   // 16-byte groups per thread (sweep 2..256: 8 is the best or within 1 % of it for every kernel)
    return (unsigned)std::max<size_t>(1, std::min<size_t>((n4 + (size_t)256 * qpt - 1) / ((size_t)256 * qpt), 0x7fffffff));
}

extern "C" int vacv_cuda_dtype_change(const void* src, void* dst, size_t n, int from_dtype, int to_dtype, void* stream) {
    VACV_REQUIRE(src && dst, "dtype_change: null pointer");
    VACV_REQUIRE(n > 0, "dtype_change: empty");
    cudaStream_t s = as_stream(stream);
    if (from_dtype == to_dtype) {   // tensor.cpp:464-466: clone
        cudaError_t e = cudaMemcpyAsync(dst, src, n * elem_size(from_dtype), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "dtype_change: %s", cudaGetErrorString(e));
        return VACV_OK;
    }
    const bool to_f32 = from_dtype == VACV_INT8 && to_dtype == VACV_FP32;
    const bool to_u8 = from_dtype == VACV_FP32 && to_dtype == VACV_INT8;
    if (!to_f32 && !to_u8)   // the reference silently returns an uninitialised tensor here (tensor.cpp:494-499)
        return set_error(VACV_ERR_UNSUPPORTED, "dtype_change: %d -> %d (reference: INT8<->FP32 only)", from_dtype, to_dtype);
    const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    const size_t n4 = aligned ? n / 4 : 0;
    if (n4) {
        const unsigned blocks = stream_ctas(n4);
        if (to_f32) u8_to_f32_kernel<<<blocks, 256, 0, s>>>((const uint8_t*)src, (float*)dst, n4);
        else f32_to_u8_kernel<<<blocks, 256, 0, s>>>((const float*)src, (uint8_t*)dst, n4);
    }
    if (4 * n4 < n) dtype_tail_kernel<<<ceil_div(n - 4 * n4, 256), 256, 0, s>>>(src, dst, 4 * n4, n, to_f32 ? 1 : 0);
    return check_launch("dtype_change");
}

template <int C>
static void launch_norm_hwc(const void* src, float* dst, const float* mean, const float* stddev, const NormGeom& g, int src_dtype,
                            dim3 grid, cudaStream_t s) {
    if (src_dtype == VACV_INT8) normalize_u8_hwc_kernel<C><<<grid, 256, 0, s>>>((const uint8_t*)src, dst, mean, stddev, g);
    else normalize_f32_hwc_kernel<C><<<grid, 256, 0, s>>>((const float*)src, dst, mean, stddev, g);
}

extern "C" int vacv_cuda_normalize(const void* src, float* dst, int batch, int w, int h, int c, int src_dtype, int layout,
                                   const float* mean, const float* stddev, int stats_per_frame, void* stream) {
    VACV_REQUIRE(src && dst && mean && stddev, "normalize: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0, "normalize: non-positive size");
    if (src_dtype != VACV_INT8 && src_dtype != VACV_FP32) return set_error(VACV_ERR_UNSUPPORTED, "normalize: src dtype %d", src_dtype);
    VACV_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "normalize: buffers must be 16-byte aligned");
    VACV_REQUIRE((size_t)w * h * c < 0xffffffffull, "normalize: frame too large");
    cudaStream_t s = as_stream(stream);
    const int sms = current_sm_count();
    NormGeom g;
    g.c = c; g.wh = (unsigned)w * h; g.per_frame = g.wh * c; g.stats_per_frame = stats_per_frame;
    const size_t es = src_dtype == VACV_INT8 ? 1 : 4;
    if (layout == VACV_NHWC && c > 1) {
        if (c > kNormMaxC) return set_error(VACV_ERR_UNSUPPORTED, "normalize: HWC supports c <= %d", kNormMaxC);
        VACV_REQUIRE((g.per_frame % 4) == 0 || batch == 1, "normalize: HWC batch needs w*h*c %% 4 == 0");
        const unsigned ctas = stream_ctas(g.per_frame / 4);
        for (int f0 = 0; f0 < batch; f0 += 65535) {
            dim3 grid(ctas, min(batch - f0, 65535));
            const void* sp = (const uint8_t*)src + (size_t)f0 * g.per_frame * es;
            float* dp = dst + (size_t)f0 * g.per_frame;
            const float* mp = mean + (stats_per_frame ? (size_t)f0 * c : 0);
            const float* dvp = stddev + (stats_per_frame ? (size_t)f0 * c : 0);
            if (c == 2) launch_norm_hwc<2>(sp, dp, mp, dvp, g, src_dtype, grid, s);
            else if (c == 3) launch_norm_hwc<3>(sp, dp, mp, dvp, g, src_dtype, grid, s);
            else launch_norm_hwc<4>(sp, dp, mp, dvp, g, src_dtype, grid, s);
        }
    } else {
        const long long planes = (long long)batch * c;
        const int chunk = 65535 / c * c;
        const unsigned ctas = stream_ctas(g.wh / 4 + 1);
        for (long long p0 = 0; p0 < planes; p0 += chunk) {
            dim3 grid(ctas, (unsigned)min((long long)chunk, planes - p0));
            const void* sp = (const uint8_t*)src + (size_t)p0 * g.wh * es;
            float* dp = dst + (size_t)p0 * g.wh;
            const float* mp = mean + (stats_per_frame ? (size_t)(p0 / c) * c : 0);
            const float* dvp = stddev + (stats_per_frame ? (size_t)(p0 / c) * c : 0);
            if (src_dtype == VACV_INT8) normalize_plane_kernel<true><<<grid, 256, 0, s>>>(sp, dp, mp, dvp, g);
            else normalize_plane_kernel<false><<<grid, 256, 0, s>>>(sp, dp, mp, dvp, g);
        }
    }
    return check_launch("normalize");
}
