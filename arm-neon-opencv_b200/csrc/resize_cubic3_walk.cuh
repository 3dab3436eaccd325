// Bicubic resize of 3-channel interleaved images (the C4 shape): one thread walks down one output column.
//
// Reference arithmetic: u8 = OpenCV 2.4.13 cv::resize (the reference's only u8 cubic path,
// src/cv/resize.cpp:33-36); fp32 = ResizeNaive::resize_naive_inter_cubic_fp32_three_channel
// (src/cv/resize_naive.cpp:187-366).  Both are separable with a rolling window of four horizontally filtered rows.
// Here the window lives in REGISTERS: a thread owns one output column of a vertical segment, filters each source
// row of its column exactly once (12 contiguous tap bytes read as aligned 32-bit words, lanes read adjacent words
// -> coalesced; PRMT regroups them per channel, two IDP.2A form sum(tap * 16-bit coef)), shifts the window and
// emits an output pixel whenever the window covers its four tap rows.  No shared-memory ring, no CTA barriers in
// the loop: warps run independently, so one warp's DRAM latency hides behind the others' arithmetic; the next
// source row is also prefetched into registers one step ahead.
// Output pixels are staged per WARP in shared memory (4 rows x 32 columns) and stored as lane-contiguous 128-bit
// chunks when the destination rows are 16-byte aligned (direct element stores otherwise).
#pragma once
#include <climits>
#include <type_traits>

#include "resize_coeffs.cuh"
#include "vacv_common.cuh"

namespace vacv {

constexpr int kWalkThreads = 128;     // 4 warps = 128 adjacent output columns
constexpr int kWalkStageRows = 4;     // output rows staged per warp between stores
constexpr int kWalkMaxRows = 512;     // output rows per segment (vertical tables live in shared memory)

struct WalkGeom {
    int w, h, wo, ho;
    int strips, segs, rows_per_seg;
    int store16;                       // destination rows / images are 16-byte aligned
    int ring_pitch;                    // kAsync: bytes per ring row of a warp (multiple of 16, >= the widest warp span)
    double scale_x, scale_y;
    size_t src_image, dst_image;       // elements between images
};

struct __align__(32) WalkRow {   // per output row of the segment
    float b[4];     // the four vertical weights
    int last;       // last tap row
    int pad[3];
};

__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c) {   // a.lo16 * b.byte0 + a.hi16 * b.byte1 + c  (a signed halves, b unsigned bytes)
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c) {   // a.lo16 * b.byte2 + a.hi16 * b.byte3 + c
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr int kWalkRing = 8;      // kAsync: ring slots (source rows) per warp
constexpr int kWalkAhead = 6;     // kAsync: rows in flight ahead of the one being filtered (<= kWalkRing - 2)

// fp32, one output column per thread; C = 3: interleaved BGR (resize_naive.cpp:187-366), C = 1: one plane of a CHW tensor
// (resize_naive_inter_cubic_fp32_one_channel, :368-529; the CHW wrapper :547-569 calls it per plane).
// kAsync (source rows and base 16-byte aligned): each WARP streams the bytes its 32 columns need, row by row, into its own
// shared-memory ring with cp.async (LDGSTS) kWalkAhead rows ahead, so DRAM latency is covered without holding registers;
// otherwise the next row is prefetched into registers.
template <int C, bool kAsync>
__global__ void __launch_bounds__(kWalkThreads) resize_cubic3_walk_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, WalkGeom g) {
    constexpr int PX = 4 * C, NW = 4 * C;            // bytes per pixel, 32-bit words a thread reads per source row
    constexpr int kWarpRow = 32 * PX;                // bytes one warp produces per output row
    extern __shared__ __align__(16) uint8_t smem[];
    WalkRow* rows = reinterpret_cast<WalkRow*>(smem);                                   // [rows_per_seg + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* stage = smem + (g.rows_per_seg + 1) * (int)sizeof(WalkRow) + warp * (kWalkStageRows * kWarpRow);
    uint8_t* ring = smem + (g.rows_per_seg + 1) * (int)sizeof(WalkRow) + (kWalkThreads / 32) * (kWalkStageRows * kWarpRow) + warp * (kWalkRing * g.ring_pitch);
    const int strip = blockIdx.x % g.strips, seg = blockIdx.x / g.strips;
    const int dx_warp = strip * kWalkThreads + (tid & ~31);
    const int dx = dx_warp + lane;
    const int dy_begin = seg * g.rows_per_seg, nrows = min(g.ho, dy_begin + g.rows_per_seg) - dy_begin;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(src + blockIdx.y * g.src_image);
    uint8_t* out_img = reinterpret_cast<uint8_t*>(dst + blockIdx.y * g.dst_image);
    const unsigned row_bytes = (unsigned)g.w * PX, out_row_bytes = (unsigned)g.wo * PX;   // an image is < 4 GiB (checked by the launcher)

    // ---- once per CTA: vertical tables of the segment (entry nrows = sentinel); once per thread: x taps of its column
    for (int r = tid; r <= nrows; r += kWalkThreads) {
        WalkRow e;
        int ofs;
        cubic_naive_scaled(dy_begin + min(r, nrows - 1), g.h, g.scale_y, ofs, e.b);
        e.last = r < nrows ? ofs + 2 : INT_MAX;       // tap rows ofs-1 .. ofs+2 are in range by construction (border folding)
        e.pad[0] = e.pad[1] = e.pad[2] = 0;
        rows[r] = e;
    }
    float xa[4];
    int x_first;
    {
        int ofs;
        cubic_naive_scaled(min(dx, g.wo - 1), g.w, g.scale_x, ofs, xa);
        x_first = ofs - 1;
    }
    const int a0 = x_first * PX;                      // byte offset of the first tap in a source row
    const uint8_t* const colp = img + a0;
    // kAsync: the warp's byte span of a source row = [span0, span0 + 16 * nchunk), 16-byte aligned (x_first is monotone in dx)
    const int span0 = __shfl_sync(0xffffffffu, a0, 0) & ~15;
    const int nchunk = (__shfl_sync(0xffffffffu, a0 + 4 * PX, 31) - span0 + 15) >> 4;
    if (kAsync && nchunk * 16 > g.ring_pitch) __trap();   // the launcher's bound on the span is wrong: fail loudly
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t tap_s = ring_s + (uint32_t)(a0 - span0);
    const uint8_t* const spanp = img + span0 + 16 * lane;
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    __syncthreads();

    float pf[NW];                                      // the 4 x 3 tap values of the next source row
    auto prefetch = [&](int r) {                       // !kAsync: into registers one step ahead.  kAsync: row r + kWalkAhead into the ring
        if (kAsync) {
            const int rr = min(r + kWalkAhead, g.h - 1);
            const uint32_t slot = ring_s + (uint32_t)((r + kWalkAhead) & (kWalkRing - 1)) * g.ring_pitch + 16 * lane;
            const uint8_t* gp = spanp + (size_t)(unsigned)rr * row_bytes;
            if (lane < nchunk) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(gp) : "memory");
            if (lane + 32 < nchunk) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + 512), "l"(gp + 512) : "memory");   // ring_pitch <= 1024
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
        const float* wp = reinterpret_cast<const float*>(colp + (size_t)(unsigned)min(r, g.h - 1) * row_bytes);
#pragma unroll
        for (int i = 0; i < NW; ++i) pf[i] = __ldg(wp + i);
    };
    auto fetch = [&](int t) {                          // kAsync: row t has landed in the ring -> its tap values
        asm volatile("cp.async.wait_group %0;" ::"n"(kWalkAhead) : "memory");
        __syncwarp();                                  // every lane's copies of row t are visible; everyone is done with row t - 2's slot
        const uint32_t p = tap_s + (uint32_t)(t & (kWalkRing - 1)) * g.ring_pitch;
#pragma unroll
        for (int i = 0; i < NW; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pf[i]) : "r"(p + 4 * i));
    };
    auto hfilter = [&](float (&H)[C]) {               // horizontal pass of the fetched row, resize_naive.cpp:230 order
#pragma unroll
        for (int k = 0; k < C; ++k) H[k] = pf[k] * xa[0] + pf[C + k] * xa[1] + pf[2 * C + k] * xa[2] + pf[3 * C + k] * xa[3];
    };

    float H[4][C];                                     // filtered rows: row t of the walk lives in slot t & 3
    int staged = 0;
    const bool staged_store = g.store16 && dx_warp + 32 <= g.wo;
    const uint32_t stage_lane = (uint32_t)__cvta_generic_to_shared(stage) + PX * lane;
    uint32_t sp = stage_lane;
    constexpr int kChunks = kWarpRow / 16;             // 16-byte chunks per staged row
    constexpr int kFlushIters = (kWalkStageRows * kChunks + 31) / 32;
    uint8_t* gflush = out_img + (size_t)dy_begin * out_row_bytes + (size_t)dx_warp * PX;   // first staged row of this warp in global memory
    float* gdirect = reinterpret_cast<float*>(out_img + (size_t)dy_begin * out_row_bytes) + (size_t)dx * C;
    auto flush = [&]() {                               // the warp's staged rows -> global, 16 bytes per lane
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kFlushIters; ++j) {
            const int i = lane + 32 * j, ty = i / kChunks, q = i - ty * kChunks;
            if (i < staged * kChunks)
                st_stream16(gflush + (size_t)(unsigned)ty * out_row_bytes + 16 * q, *reinterpret_cast<const uint4*>(stage + ty * kWarpRow + 16 * q));
        }
        __syncwarp();
        gflush += (size_t)(unsigned)staged * out_row_bytes;
        staged = 0;
        sp = stage_lane;
    };
    // vertical pass + store of one output pixel from the window (h0 = oldest row)
    auto emit = [&](const float (&h0)[C], const float (&h1)[C], const float (&h2)[C], const float (&h3)[C], uint32_t entry) {
        float4 bw;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bw.x), "=f"(bw.y), "=f"(bw.z), "=f"(bw.w) : "r"(entry));
        float o[C];
#pragma unroll
        for (int k = 0; k < C; ++k) o[k] = h0[k] * bw.x + h1[k] * bw.y + h2[k] * bw.z + h3[k] * bw.w;   // resize_naive.cpp:345
        if (staged_store) {
#pragma unroll
            for (int k = 0; k < C; ++k) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sp + 4 * k), "f"(o[k]) : "memory");
            sp += kWarpRow;
            if (++staged == kWalkStageRows) flush();
        } else {
            if (dx < g.wo) {
#pragma unroll
                for (int k = 0; k < C; ++k) gdirect[k] = o[k];
            }
            gdirect += (size_t)g.wo * C;
        }
    };

    // ---- the walk: source rows t = first tap row of the segment's first output row, t+1, ...; row t goes to window slot
    // t & 3, so inside the 4x unrolled body every slot index is a compile-time constant and the window never moves between
    // registers.  An output row is emitted as soon as its last tap row is filtered.
    uint32_t entry = rows_s;                           // shared address of the next output row's table entry
    int next_last, t;
    asm volatile("ld.shared.s32 %0, [%1+16];" : "=r"(next_last) : "r"(entry));
    t = next_last - 3;
    const uint32_t entry_end = rows_s + nrows * (int)sizeof(WalkRow);
    if (kAsync) {
        for (int a = 0; a < kWalkAhead; ++a) prefetch(t - kWalkAhead + a);   // rows t .. t + kWalkAhead - 1 in flight
    } else {
        prefetch(t);
    }
    while ((t & 3) != 0) {                             // leading rows up to the first multiple of 4: no complete window yet
        float hv[C];
        if (kAsync) { prefetch(t); fetch(t); }
        hfilter(hv);
        const int slot = t & 3;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            if (slot == 1) H[1][k] = hv[k];
            if (slot == 2) H[2][k] = hv[k];
            if (slot == 3) H[3][k] = hv[k];
        }
        ++t;
        if (!kAsync) prefetch(t);
    }
    while (entry != entry_end) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {                  // t & 3 == u
            if (kAsync) { prefetch(t); fetch(t); }     // issue row t + kWalkAhead, wait for row t
            hfilter(H[u]);
            if (!kAsync) prefetch(t + 1);              // the next row of the walk
            while (next_last == t) {                   // uniform across the CTA; more than once per row only when upscaling
                emit(H[(u + 1) & 3], H[(u + 2) & 3], H[(u + 3) & 3], H[u], entry);
                entry += (int)sizeof(WalkRow);
                asm volatile("ld.shared.s32 %0, [%1+16];" : "=r"(next_last) : "r"(entry));   // sentinel INT_MAX after the last row
            }
            ++t;
        }
    }
    if (kAsync) asm volatile("cp.async.wait_all;" ::: "memory");
    if (staged_store && staged) flush();
}

// =====================================================================================================================
// u8, two output columns per thread (32 columns apart, so that a warp's tap reads stay bank-conflict free).  Same walk as
// above; the vertical pass runs on PACKED fp32 pairs (FFMA2 /
// FMUL2, sm_100): packed value = (column A, column B) of one channel, so OpenCV's mul/add chain costs 8 instructions per
// two output values instead of 16.  Each product and each sum is rounded separately, exactly like mulps / addps:
// ptxas contracts mul.f32x2 + add.f32x2 (even .rn ones) into one FFMA2, so both are written as fma.rn.f32x2 with the
// neutral operand (-0.0 resp. 1.0) coming from a KERNEL PARAMETER -- a value ptxas cannot see, hence cannot fold.
// The last (3*w_out & 7) elements of a row follow OpenCV's integer tail rule; resize_cubic3_tail_kernel rewrites those
// (at most 3) pixels per row afterwards, so this kernel has no per-pixel special case.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 pack2i(int lo, int hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void unpack2i(f32x2 v, int& lo, int& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

struct Walk2Geom {
    int w, h, wo, ho;
    int strips, segs, rows_per_seg;
    int store16;                       // destination rows / images are 16-byte aligned
    double scale_x, scale_y;
    size_t src_image, dst_image;       // bytes between images
    f32x2 one2, negzero2, magic2, negmagic2;   // (1,1), (-0,-0), (1.5*2^23)x2, (-1.5*2^23)x2 -- opaque to ptxas on purpose
    int ring_pitch;                    // kAsync: bytes per ring row of a warp (multiple of 16, >= the widest warp span)
};

struct __align__(16) Walk2Row {   // per output row of the segment: vertical weights duplicated into pairs, last tap row
    float b[8];     // (b0,b0) (b1,b1) (b2,b2) (b3,b3)
    int last;       // y0 + 3 (unclamped)
    int pad[3];
};

constexpr int kWalk2Cols = 256;   // output columns per CTA (128 threads x 2)
constexpr int kWalk2Ring = kWalkRing, kWalk2Ahead = kWalkAhead;

// kAsync (source rows and base 16-byte aligned): each WARP streams the bytes its 64 columns need, row by row, into its own
// shared-memory ring with cp.async (LDGSTS) kWalk2Ahead rows ahead -- DRAM latency is covered without holding registers
// (ncu: the register-prefetch variant ran at 4.8 warps/scheduler and stalled on long_scoreboard).  Otherwise the next row
// is prefetched into registers.
template <bool kAsync>
__global__ void __launch_bounds__(kWalkThreads, kAsync ? 6 : 1) resize_cubic3_walk2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, Walk2Geom g) {
    constexpr int kWarpRow = 64 * 3;                 // bytes one warp produces per output row
    extern __shared__ __align__(16) uint8_t smem[];
    Walk2Row* rows = reinterpret_cast<Walk2Row*>(smem);                                   // [rows_per_seg + 1]
    const int tid = threadIdx.x, lane = tid & 31;
    uint8_t* stage = smem + (g.rows_per_seg + 1) * (int)sizeof(Walk2Row) + (tid >> 5) * (kWalkStageRows * kWarpRow);
    uint8_t* ring = smem + (g.rows_per_seg + 1) * (int)sizeof(Walk2Row) + (kWalkThreads / 32) * (kWalkStageRows * kWarpRow) +
                    (tid >> 5) * (kWalk2Ring * g.ring_pitch);
    const int strip = blockIdx.x % g.strips, seg = blockIdx.x / g.strips;
    const int dx_warp = strip * kWalk2Cols + (tid >> 5) * 64;
    const int dx = dx_warp + lane;                   // columns dx, dx + 32: each tap-word read of a warp then covers one contiguous
                                                     // ~128-byte span (adjacent columns per lane cost a 2-way bank conflict on every read)
    const int dy_begin = seg * g.rows_per_seg, nrows = min(g.ho, dy_begin + g.rows_per_seg) - dy_begin;
    const uint8_t* img = src + blockIdx.y * g.src_image;
    uint8_t* out_img = dst + blockIdx.y * g.dst_image;
    const unsigned row_bytes = (unsigned)g.w * 3, out_row_bytes = (unsigned)g.wo * 3;

    for (int r = tid; r <= nrows; r += kWalkThreads) {   // entry nrows = sentinel that never matches
        Walk2Row e;
        int s, q[4];
        cubic_cv_coord_scaled(dy_begin + min(r, nrows - 1), g.h, g.scale_y, false, s, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) e.b[2 * j] = e.b[2 * j + 1] = (float)q[j] * (1.f / (2048 * 2048));
        e.last = r < nrows ? s + 2 : INT_MAX;
        e.pad[0] = e.pad[1] = e.pad[2] = 0;
        rows[r] = e;
    }
    // x taps of the two columns: four CONSECUTIVE source pixels from x_first each; taps OpenCV clamps onto the edge pixel
    // have their integer coefficients added up (identical sums)
    const uint8_t* colp[2];
    int sh[2], c01[2], c23[2], aw[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int dxc = min(dx + 32 * c, g.wo - 1);
        int s, q[4], xc[4] = {0, 0, 0, 0};
        cubic_cv_coord_scaled(dxc, g.w, g.scale_x, true, s, q);
        const int x_first = min(max(s - 1, 0), g.w - 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pos = min(max(s - 1 + j, 0), g.w - 1) - x_first;   // 0..3
#pragma unroll
            for (int t = 0; t < 4; ++t) xc[t] += pos == t ? q[j] : 0;
        }
        const int a0 = x_first * 3;
        colp[c] = img + (a0 & ~3);
        aw[c] = a0 & ~3;
        sh[c] = (a0 & 3) * 8;
        c01[c] = (xc[0] & 0xffff) | (xc[1] << 16);
        c23[c] = (xc[2] & 0xffff) | (xc[3] << 16);
    }
    // kAsync: the warp's byte span of a source row = [span0, span0 + 16 * nchunk), 16-byte aligned (x_first is monotone in dx)
    const int span0 = __shfl_sync(0xffffffffu, aw[0], 0) & ~15;
    const int nchunk = (__shfl_sync(0xffffffffu, aw[1] + (sh[1] ? 16 : 12), 31) - span0 + 15) >> 4;
    if (kAsync && nchunk * 16 + 16 > g.ring_pitch) __trap();   // the launcher's bound on the span is wrong: fail loudly
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t tap_s[2] = {ring_s + (uint32_t)(aw[0] - span0), ring_s + (uint32_t)(aw[1] - span0)};
    const uint8_t* const spanp = img + span0 + 16 * lane;
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    const uint32_t stage_lane = (uint32_t)__cvta_generic_to_shared(stage) + 3 * lane;
    const f32x2 one2 = g.one2, negzero2 = g.negzero2, magic2 = g.magic2, negmagic2 = g.negmagic2;
    __syncthreads();

    uint32_t pf[2][4];                                 // words of the next source row
    auto prefetch = [&](int r) {                       // !kAsync: into registers one step ahead.  kAsync: row r + kWalk2Ahead into the ring
        if (kAsync) {
            const int rr = min(max(r + kWalk2Ahead, 0), g.h - 1);   // OpenCV clamps tap rows
            const uint32_t slot = ring_s + (uint32_t)((r + kWalk2Ahead) & (kWalk2Ring - 1)) * g.ring_pitch + 16 * lane;
            const uint8_t* gp = spanp + (size_t)(unsigned)rr * row_bytes;
            if (lane < nchunk) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(gp) : "memory");
            if (lane + 32 < nchunk) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + 512), "l"(gp + 512) : "memory");   // ring_pitch <= 1024
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
        r = min(max(r, 0), g.h - 1);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(colp[c] + (size_t)(unsigned)r * row_bytes);
            pf[c][0] = __ldg(wp); pf[c][1] = __ldg(wp + 1); pf[c][2] = __ldg(wp + 2);
            pf[c][3] = sh[c] ? __ldg(wp + 3) : 0u;
        }
    };
    auto fetch = [&](int t) {                          // kAsync: row t has landed in the ring -> its tap words
        asm volatile("cp.async.wait_group %0;" ::"n"(kWalk2Ahead) : "memory");
        __syncwarp();                                  // every lane's copies of row t are visible; everyone is done with row t - 2's slot
        const uint32_t off = (uint32_t)(t & (kWalk2Ring - 1)) * g.ring_pitch;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pf[c][i]) : "r"(tap_s[c] + off + 4 * i));
        }
    };
    auto hfilter = [&](f32x2 (&H)[3]) {               // horizontal pass of the prefetched row: (column A, column B) per channel
        int hb[2], hg[2], hr[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t b0 = __funnelshift_r(pf[c][0], pf[c][1], sh[c]), b1 = __funnelshift_r(pf[c][1], pf[c][2], sh[c]),
                           b2 = __funnelshift_r(pf[c][2], pf[c][3], sh[c]);
            // 12 bytes  b0 = [t0.b t0.g t0.r t1.b]  b1 = [t1.g t1.r t2.b t2.g]  b2 = [t2.r t3.b t3.g t3.r]  -> one word per channel
            const unsigned cb = __byte_perm(__byte_perm(b0, b1, 0x0630), b2, 0x5210);   // t0.b t1.b t2.b t3.b
            const unsigned cg = __byte_perm(__byte_perm(b0, b1, 0x0741), b2, 0x6210);   // t0.g t1.g t2.g t3.g
            const unsigned cr = __byte_perm(__byte_perm(b0, b1, 0x0052), b2, 0x7410);   // t0.r t1.r t2.r t3.r
            // sum(tap * coef) on top of the bit pattern of 1.5*2^23 (exact int -> float for |H| < 2^22 after subtracting it)
            hb[c] = dp2a_hi_su(c23[c], cb, dp2a_lo_su(c01[c], cb, 0x4B400000));
            hg[c] = dp2a_hi_su(c23[c], cg, dp2a_lo_su(c01[c], cg, 0x4B400000));
            hr[c] = dp2a_hi_su(c23[c], cr, dp2a_lo_su(c01[c], cr, 0x4B400000));
        }
        H[0] = fma2(pack2i(hb[0], hb[1]), one2, negmagic2);
        H[1] = fma2(pack2i(hg[0], hg[1]), one2, negmagic2);
        H[2] = fma2(pack2i(hr[0], hr[1]), one2, negmagic2);
    };

    f32x2 H[4][3];                                     // filtered rows: row t of the walk lives in slot t & 3
    int staged = 0;
    const bool staged_store = g.store16 && dx_warp + 64 <= g.wo;
    uint32_t sp = stage_lane;
    constexpr int kChunks = kWarpRow / 16;             // 12 16-byte chunks per staged row
    constexpr int kFlushIters = (kWalkStageRows * kChunks + 31) / 32;
    uint8_t* gflush = out_img + (size_t)dy_begin * out_row_bytes + (size_t)dx_warp * 3;   // first staged row of this warp in global memory
    uint8_t* gdirect = out_img + (size_t)dy_begin * out_row_bytes + (size_t)dx * 3;
    auto flush = [&]() {                               // the warp's staged rows -> global, 16 bytes per lane
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kFlushIters; ++j) {
            const int i = lane + 32 * j, ty = i / kChunks, q = i - ty * kChunks;
            if (i < staged * kChunks)
                st_stream16(gflush + (size_t)(unsigned)ty * out_row_bytes + 16 * q, *reinterpret_cast<const uint4*>(stage + ty * kWarpRow + 16 * q));
        }
        __syncwarp();
        gflush += (size_t)(unsigned)staged * out_row_bytes;
        staged = 0;
        sp = stage_lane;
    };

    // vertical pass + store of the two output pixels from the window (h0 = oldest row); bw = row table entry
    auto emit = [&](const f32x2 (&h0)[3], const f32x2 (&h1)[3], const f32x2 (&h2)[3], const f32x2 (&h3)[3], uint32_t entry) {
        f32x2 w0, w1, w2, w3;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(entry));
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+16];" : "=l"(w2), "=l"(w3) : "r"(entry));
        int v[3][2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {   // OpenCV's SSE2 body: mulps, addps (one rounding each), cvtps2dq (half-even), packs, packus
            f32x2 f = fma2(h0[k], w0, negzero2);
            f = fma2(f, one2, fma2(h1[k], w1, negzero2));
            f = fma2(f, one2, fma2(h2[k], w2, negzero2));
            f = fma2(f, one2, fma2(h3[k], w3, negzero2));
            // |f| < 2^22: adding 1.5*2^23 rounds half-to-even at integer granularity; subtracting its bit pattern and clamping to
            // [0,255] is one DPX op per value (the intermediate s16 saturation of packs cannot change the result)
            f = fma2(f, one2, magic2);
            int lo, hi;
            unpack2i(f, lo, hi);
            v[k][0] = __viaddmin_s32_relu(lo, -0x4B400000, 255);
            v[k][1] = __viaddmin_s32_relu(hi, -0x4B400000, 255);
        }
        if (staged_store) {   // the two pixels are 32 columns = 96 bytes apart in the staged row
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                asm volatile("st.shared.u8 [%0], %1;" ::"r"(sp + k), "r"(v[k][0]) : "memory");
                asm volatile("st.shared.u8 [%0], %1;" ::"r"(sp + 96 + k), "r"(v[k][1]) : "memory");
            }
            sp += kWarpRow;
            if (++staged == kWalkStageRows) flush();
        } else {
            if (dx < g.wo) { gdirect[0] = (uint8_t)v[0][0]; gdirect[1] = (uint8_t)v[1][0]; gdirect[2] = (uint8_t)v[2][0]; }
            if (dx + 32 < g.wo) { gdirect[96] = (uint8_t)v[0][1]; gdirect[97] = (uint8_t)v[1][1]; gdirect[98] = (uint8_t)v[2][1]; }
            gdirect += out_row_bytes;
        }
    };

    // ---- the walk (see resize_cubic3_walk_kernel): row t -> slot t & 3, emit when t is the last tap row of the next output row
    uint32_t entry = rows_s;                           // shared address of the next output row's table entry
    int next_last, t;
    asm volatile("ld.shared.s32 %0, [%1+32];" : "=r"(next_last) : "r"(entry));
    t = next_last - 3;
    const uint32_t entry_end = rows_s + nrows * (int)sizeof(Walk2Row);
    if (kAsync) {
        for (int a = 0; a < kWalk2Ahead; ++a) prefetch(t - kWalk2Ahead + a);   // rows t .. t + kWalk2Ahead - 1 in flight
    } else {
        prefetch(t);
    }
    while ((t & 3) != 0) {                             // leading rows up to the first multiple of 4: no complete window yet
        f32x2 hv[3];
        if (kAsync) { prefetch(t); fetch(t); }
        hfilter(hv);
        const int slot = t & 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (slot == 1) H[1][k] = hv[k];
            if (slot == 2) H[2][k] = hv[k];
            if (slot == 3) H[3][k] = hv[k];
        }
        ++t;
        if (!kAsync) prefetch(t);
    }
    while (entry != entry_end) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {                  // t & 3 == u
            if (kAsync) { prefetch(t); fetch(t); }     // issue row t + kWalk2Ahead, wait for row t
            hfilter(H[u]);
            if (!kAsync) prefetch(t + 1);              // the next row of the walk
            while (next_last == t) {                   // uniform across the CTA; more than once per row only when upscaling
                emit(H[(u + 1) & 3], H[(u + 2) & 3], H[(u + 3) & 3], H[u], entry);
                entry += (int)sizeof(Walk2Row);
                asm volatile("ld.shared.s32 %0, [%1+32];" : "=r"(next_last) : "r"(entry));   // sentinel INT_MAX after the last row
            }
            ++t;
        }
    }
    if (kAsync) asm volatile("cp.async.wait_all;" ::: "memory");
    if (staged_store && staged) flush();
}

// Rewrites the pixels of every output row that reach into its last (3*w_out & 7) elements: OpenCV's SSE2 body stops at
// x < (width & ~7) and the remaining elements use FixedPtCast<int, uchar, 22> on integer sums.
__global__ void resize_cubic3_tail_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w, int h, int wo, int ho,
                                          double scale_x, double scale_y, size_t src_image, size_t dst_image, int first_px, int images) {
    const int npx = wo - first_px;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)images * ho * npx) return;
    const int dx = first_px + (int)(i % npx), dy = (int)((i / npx) % ho), im = (int)(i / ((long long)npx * ho));
    const uint8_t* img = src + (size_t)im * src_image;
    int sx, sy, qx[4], qy[4];
    cubic_cv_coord_scaled(dx, w, scale_x, true, sx, qx);
    cubic_cv_coord_scaled(dy, h, scale_y, false, sy, qy);
    int H[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint8_t* rowp = img + (size_t)min(max(sy - 1 + j, 0), h - 1) * w * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int acc = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) acc += rowp[min(max(sx - 1 + t, 0), w - 1) * 3 + k] * qx[t];
            H[j][k] = acc;
        }
    }
    const int vec_end = (wo * 3) & ~7;
    uint8_t* o = dst + (size_t)im * dst_image + ((size_t)dy * wo + dx) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int v;
        if (dx * 3 + k < vec_end) {
            float f = (float)H[0][k] * ((float)qy[0] * (1.f / (2048 * 2048)));
            f = f + (float)H[1][k] * ((float)qy[1] * (1.f / (2048 * 2048)));
            f = f + (float)H[2][k] * ((float)qy[2] * (1.f / (2048 * 2048)));
            f = f + (float)H[3][k] * ((float)qy[3] * (1.f / (2048 * 2048)));
            v = __float2int_rn(f);
        } else {
            v = (H[0][k] * qy[0] + H[1][k] * qy[1] + H[2][k] * qy[2] + H[3][k] * qy[3] + (1 << 21)) >> 22;
        }
        o[k] = (uint8_t)clamp255(v);
    }
}

}  // namespace vacv
