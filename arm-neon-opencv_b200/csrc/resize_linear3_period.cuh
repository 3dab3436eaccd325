// a5, u8 bilinear resize of interleaved BGR at RATIONAL horizontal scales (1920 -> 1280 = 3 source pixels per 2 output pixels): the
// construction of the periodic bicubic walker (resize_cubic3_period.cuh) applied to the reference's integer bilinear rule
// (src/cv/resize_naive.cpp:10-68: source index and weights per column / row, (p00*cx0*cy0 + p01*cx1*cy0 + p10*cx0*cy1 + p11*cx1*cy1) >> 22).
//
// Why: ncu of the persistent bilinear pipeline at 1080p -> 720p (profiles/r2_lin720_head_ncu_raw.txt): the shared-memory data pipe at 73 %
// of its wavefront peak (three tap words per column and source row at a lane stride of 4.5 bytes = two wavefronts each) and 55
// instructions per pixel -- a thread's columns lie a whole CTA apart, so every column fetches and aligns its own 6 bytes.  When
// w_in : w_out = P : Q the tap pattern repeats every Q output columns / P source pixels, and for P > Q the two taps of a thread's
// Q*KP ADJACENT columns lie inside its own P*KP source pixels: windows of neighbouring lanes do not even overlap.  A thread
//   * reads its window ONCE per source row (3:2 with KP = 4: nine 32-bit shared loads for eight columns, lane stride 36 bytes = 9
//     banks: conflict-free, one wavefront each),
//   * cuts every column's [L R] byte pairs out with PRMTs whose selectors are compile-time constants (two per column) and forms
//     the horizontal sums with three IDP.2A on the packed 16-bit weights,
//   * keeps the sums of the last two source rows in registers (by row parity, nothing is copied) and blends an output row as soon
//     as its lower tap row is there; the vertical weights carry a factor 4, so the blended byte is byte 3 of the 32-bit sum,
//   * source rows arrive by one cp.async.bulk per warp and row (8-slot ring, 6 rows ahead), output rows leave as lane-contiguous
//     16-byte chunks from a two-row staging buffer -- the periodic bicubic walker's plumbing.
// The weights stay per-lane registers computed with the reference's own float arithmetic (the fp32 scale makes them not exactly
// periodic); the launcher verifies on the host, with the device's arithmetic, that every column's taps sit where the pattern expects
// them and otherwise leaves the shape to the persistent pipeline.
#pragma once
#include <climits>

#include "period_common.cuh"   // pd:: helpers (static_for, tap0, mbarrier / bulk-copy wrappers), kPdRing, kPdAhead, kPdStageRows

namespace vacv {

struct LinPeriodGeom {
    int w, h, wo, ho;
    int warp_strips, cta_strips, segs, rows_per_seg;
    double scale_x, scale_y;
    size_t src_image, dst_image;       // bytes between images
};

struct LinRow { int cy0q, cy1q, last, pad; };   // 4 * cy0, 4 * cy1, walk step that completes the row

// C = 3: interleaved BGR; C = 1: one plane of a CHW tensor (resized plane by plane, resize.cpp:73-87) or a grey image
template <int P, int Q, int KP, int C = 3>
struct LinPeriodShape {
    static constexpr int NCOL = Q * KP;          // adjacent output columns per thread
    static constexpr int NPX = P * KP;           // source pixels in a thread's window = its own period pixels
    static constexpr int LS = C * NPX;           // window bytes = bytes between the windows of neighbouring lanes
    static constexpr int NW = LS / 4;            // window words
    static constexpr int NV = C * NCOL;          // output bytes per thread and row
    static constexpr int kWarpRow = 32 * NV;     // bytes one warp produces per output row
    static constexpr int kWarpSpan = 32 * LS;    // source bytes of a warp per row = ring slot size
    static_assert(P > Q, "down-scaling only: both taps of every column inside the thread's own pixels");
    static_assert(LS % 4 == 0 && NV % 4 == 0, "windows and output runs are whole words");
    static_assert(pd::tap0(P, Q, NCOL - 1) + 1 < NPX, "last column's right tap inside the window");
    static_assert(C == 3 || (C == 1 && NCOL % 2 == 0), "planes: columns are filtered in pairs");
};

// bytes B0..B3 of the window (any order, inside two neighbouring words) -> one word with ONE PRMT whose selector is an immediate
template <int B0, int B1, int B2, int B3, int N>
__device__ __forceinline__ uint32_t lin_pick(const uint32_t (&W)[N]) {
    constexpr int lo = (B0 < B1 ? B0 : B1) < (B2 < B3 ? B2 : B3) ? (B0 < B1 ? B0 : B1) : (B2 < B3 ? B2 : B3);
    constexpr int hi = (B0 > B1 ? B0 : B1) > (B2 > B3 ? B2 : B3) ? (B0 > B1 ? B0 : B1) : (B2 > B3 ? B2 : B3);
    constexpr int w0 = lo >> 2, w1 = (hi >> 2) > w0 ? (hi >> 2) : (w0 + 1 < N ? w0 + 1 : w0);
    static_assert((hi >> 2) <= w0 + 1 && w1 < N, "bytes must lie inside two neighbouring window words");
    constexpr unsigned sel = (unsigned)(B0 - 4 * w0) | (unsigned)(B1 - 4 * w0) << 4 | (unsigned)(B2 - 4 * w0) << 8 | (unsigned)(B3 - 4 * w0) << 12;
    return __byte_perm(W[w0], W[w1], sel);
}

template <int P, int Q, int KP, bool kSigned, bool kDown, int C = 3>
__global__ void __launch_bounds__(128) resize_linear3_period_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, LinPeriodGeom g) {
    using S = LinPeriodShape<P, Q, KP, C>;
    constexpr int NCOL = S::NCOL, LS = S::LS, NW = S::NW, NV = S::NV, kWarpRow = S::kWarpRow;
    constexpr unsigned kPitch = S::kWarpSpan;             // bytes per ring slot
    extern __shared__ __align__(16) uint8_t smem[];
    LinRow* rows = reinterpret_cast<LinRow*>(smem);       // [rows_per_seg + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rows_bytes = (g.rows_per_seg + 1) * (int)sizeof(LinRow);
    uint8_t* stage = smem + rows_bytes + warp * (kPdStageRows * kWarpRow);
    uint8_t* ring = smem + rows_bytes + nwarps * (kPdStageRows * kWarpRow) + warp * (kPdRing * kPitch);
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem + rows_bytes + nwarps * (kPdStageRows * kWarpRow + kPdRing * kPitch)) + warp * (kPdRing * 8);
    const int cta_strip = blockIdx.x % g.cta_strips, seg = blockIdx.x / g.cta_strips;
    const int wstrip = cta_strip * nwarps + warp;
    const int pt = wstrip * 32 + lane;                 // this thread's index along x: columns NCOL * pt ..
    const int dy_begin = seg * g.rows_per_seg, nrows = min(g.ho, dy_begin + g.rows_per_seg) - dy_begin;
    uint8_t* out_img = dst + blockIdx.y * g.dst_image;
    const unsigned row_bytes = (unsigned)g.w * C, out_row_bytes = (unsigned)g.wo * C;

    // walk steps: step n filters source row t_first + n; the last step is the last output row's lower tap row (all inside the image:
    // linear_coord clamps the index to [0, h - 2])
    int t_first, n_stop;
    {
        int s0, s1; float f;
        linear_coord(dy_begin, g.scale_y, g.h, s0, f);
        linear_coord(dy_begin + nrows - 1, g.scale_y, g.h, s1, f);
        t_first = __shfl_sync(0xffffffffu, s0, 0);
        n_stop = __shfl_sync(0xffffffffu, s1 + 1, 0) - t_first;
    }
    const bool active = wstrip < g.warp_strips;        // false: padding warp of the last CTA strip

    // the warp's bytes of a source row: [span0, span0 + kPitch) clipped to the row.  Everything the copy issue needs is warp-uniform;
    // the shuffles tell the compiler so (uniform registers, one UBLKCP per warp).
    const int span0 = S::kWarpSpan * wstrip;
    const int lo = __shfl_sync(0xffffffffu, min(span0, (int)row_bytes), 0), hi = __shfl_sync(0xffffffffu, min(span0 + (int)kPitch, (int)row_bytes), 0);
    const uint32_t copy_bytes = (uint32_t)(hi - lo);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t ring_dst = __shfl_sync(0xffffffffu, ring_s, 0);
    const uint32_t ubars = __shfl_sync(0xffffffffu, bars, 0);
    const uint32_t win_s = ring_s + (uint32_t)(LS * lane);
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);

    int n_pre = 0;                                     // walk step of the next copy
    const uint8_t* g_pre = src + blockIdx.y * g.src_image + (unsigned)lo + (size_t)(unsigned)t_first * row_bytes;
    auto issue = [&](const uint32_t slot) {            // warp-uniform; slot is a literal at every call site
        if (n_pre <= n_stop) {
            if (lane == 0) {
                pd::mbar_expect_tx(ubars + 8 * slot, copy_bytes);
                pd::bulk_g2s(ring_dst + slot * kPitch, g_pre, copy_bytes, ubars + 8 * slot);
            }
            g_pre += row_bytes;
        }
        ++n_pre;
    };
    // the first rows go in flight before the tables below are computed (their DRAM latency hides behind that arithmetic)
    if (active) {
        if (lane == 0) {
            for (int i = 0; i < kPdRing; ++i) pd::mbar_init(bars + 8 * i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        static_assert(kPdAhead == 6 && kPdRing == 8, "the unrolled walk below assumes 6 rows ahead in an 8-slot ring");
        issue(0); issue(1); issue(2); issue(3); issue(4); issue(5);
    }

    // per output row: vertical weights (times 4: 255 * 2049 * 2049 * 4 < 2^32, the blended byte is byte 3 of the sum) and the walk step
    // that completes it
    for (int r = tid; r <= nrows; r += blockDim.x) {   // entry nrows = sentinel that never matches
        int s; float f;
        linear_coord(dy_begin + min(r, nrows - 1), g.scale_y, g.h, s, f);
        LinRow e;
        e.cy0q = sat_short((1.f - f) * 2048.f) << 2;
        e.cy1q = sat_short(2048.f * f) << 2;
        e.last = r < nrows ? s + 1 - t_first : INT_MAX;
        e.pad = 0;
        rows[r] = e;
    }
    // x weights: column c's taps are window pixels tap0(c), tap0(c) + 1 (verified by the launcher)
    uint32_t cx[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
        int sx; float fx;
        linear_coord(min(NCOL * pt + c, g.wo - 1), g.scale_x, g.w, sx, fx);
        cx[c] = (uint32_t)(sat_short((1.f - fx) * 2048.f) & 0xffff) | ((uint32_t)sat_short(2048.f * fx) << 16);
    }
    __syncthreads();
    if (!active) return;                               // (no CTA barrier below)

    uint32_t entry = rows_s;                           // shared address of the next output row's table entry
    int next_last;
    asm volatile("ld.shared.s32 %0, [%1+8];" : "=r"(next_last) : "r"(entry));

    // horizontal pass of the walk step in ring slot `slot`: the thread's NV = 3 * NCOL sums L * cx0 + R * cx1
    auto hfilter = [&](const uint32_t slot, uint32_t parity, int (&H)[NV]) {
        pd::mbar_wait(ubars + 8 * slot, parity);
        const uint32_t p = win_s + slot * kPitch;
        // 32-bit loads.  Lane strides of 32 / 48 / 24 bytes (planes 4:3 and 2:1, BGR 4:3 and 2:1, planes 3:2) make them 8- / 4- / 2-way
        // bank conflicts, which 128- / 64-bit loads would avoid -- measured: planes unchanged (0.095 / 0.100 / 0.164 ms either way), BGR
        // 4:3 0.097 -> 0.096 ms, BGR 2:1 0.087 -> 0.128 ms (the PRMTs then wait for whole vectors): the shared-memory pipe is not what
        // bounds these kernels, 32-bit loads stay.
        uint32_t W[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W[i]) : "r"(p + 4 * i));
        if constexpr (C == 3) {
            pd::static_for<NCOL>([&](auto ic) {
                constexpr int c = decltype(ic)::value;
                constexpr int a = 3 * pd::tap0(P, Q, c);   // first byte of the left tap
                const uint32_t bg = lin_pick<a, a + 3, a + 1, a + 4>(W);   // [L.b R.b L.g R.g]
                const uint32_t rr = lin_pick<a + 2, a + 5, a + 2, a + 5>(W);   // [L.r R.r  .   . ]
                if (kSigned) {
                    H[3 * c] = __dp2a_lo((int)cx[c], (int)bg, 0); H[3 * c + 1] = __dp2a_hi((int)cx[c], (int)bg, 0); H[3 * c + 2] = __dp2a_lo((int)cx[c], (int)rr, 0);
                } else {
                    H[3 * c] = (int)__dp2a_lo(cx[c], bg, 0u); H[3 * c + 1] = (int)__dp2a_hi(cx[c], bg, 0u); H[3 * c + 2] = (int)__dp2a_lo(cx[c], rr, 0u);
                }
            });
        } else {   // planes: one PRMT per column PAIR ([L R] of both columns lie within four consecutive bytes), one IDP.2A per column
            pd::static_for<NCOL / 2>([&](auto ic) {
                constexpr int c = 2 * decltype(ic)::value;
                constexpr int a = pd::tap0(P, Q, c), b = pd::tap0(P, Q, c + 1);
                const uint32_t lr = lin_pick<a, a + 1, b, b + 1>(W);   // [L0 R0 L1 R1]
                if (kSigned) {
                    H[c] = __dp2a_lo((int)cx[c], (int)lr, 0); H[c + 1] = __dp2a_hi((int)cx[c + 1], (int)lr, 0);
                } else {
                    H[c] = (int)__dp2a_lo(cx[c], lr, 0u); H[c + 1] = (int)__dp2a_hi(cx[c + 1], lr, 0u);
                }
            });
        }
    };

    // ---- output: every row is staged in one of two kWarpRow-byte buffers of the warp and leaves at once as lane-contiguous 16-byte
    //      chunks.  Threads past the last column stage bytes that are never flushed.
    constexpr int NWD = NV / 4;                        // words per thread and row
    constexpr int kChunks = kWarpRow / 16;             // 16-byte chunks per row of the warp
    static_assert(kChunks <= 96, "flush: up to three chunks per lane");
    uint32_t st_w = stage_s + NV * lane;
    uint32_t st_f = stage_s + 16 * lane;                            // this lane's first chunk of the staged row
    int st_d = kWarpRow;                                            // distance to the other staging buffer
    const int valid_chunks = min(kWarpRow, max(0, (int)out_row_bytes - wstrip * kWarpRow)) >> 4;   // chunks of a warp row inside the image row
    const bool f0 = lane < min(valid_chunks, kChunks), f1 = kChunks > 32 && lane + 32 < min(valid_chunks, kChunks),
               f2 = kChunks > 64 && lane + 64 < min(valid_chunks, kChunks);
    uint8_t* gflush = out_img + (size_t)dy_begin * out_row_bytes + (size_t)wstrip * kWarpRow + 16 * lane;   // this lane's first chunk in global memory

    // vertical blend + store of the NV output bytes (resize_naive.cpp:60-65, the same integer regrouped row-wise, times 4)
    auto emit = [&](const int (&ht)[NV], const int (&hb)[NV]) {
        uint32_t q0, q1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(q0), "=r"(q1) : "r"(entry));
        uint32_t v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = (uint32_t)ht[i] * q0 + (uint32_t)hb[i] * q1;
#pragma unroll
        for (int j = 0; j < NWD; ++j) {
            const uint32_t x = __byte_perm(__byte_perm(v[4 * j], v[4 * j + 1], 0x0073), __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x0073), 0x5410);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(st_w + 4 * j), "r"(x) : "memory");
        }
        __syncwarp();                                  // the row is staged; the other buffer's readers passed this point a row ago
        if (f0) {
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(st_f));
            st_stream16(gflush, q);
        }
        if (kChunks > 32 && f1) {
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(st_f));
            st_stream16(gflush + 512, q);
        }
        if (kChunks > 64 && f2) {
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+1024];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(st_f));
            st_stream16(gflush + 1024, q);
        }
        gflush += out_row_bytes;
        st_w += st_d; st_f += st_d;                    // the other buffer
        st_d = -st_d;
        entry += (int)sizeof(LinRow);
        asm volatile("ld.shared.s32 %0, [%1+8];" : "=r"(next_last) : "r"(entry));   // sentinel INT_MAX after the last row
    };

    // ---- the walk: step n -> ring slot n & 7, sums of row parity n & 1 (compile-time inside the 8x unrolled body); an output row is
    //      emitted as soon as its lower tap row has been filtered (never at step 0).
    int H[2][NV];
    const uint32_t entry_end = rows_s + nrows * (int)sizeof(LinRow);
    int n = 0;
    uint32_t parity = 0;
#define VACV_LP_STEP(u)                                                                                        \
    if (entry != entry_end) {                                                                                  \
        __syncwarp(); /* every lane is done with step n - 2, whose ring slot the next copy overwrites */       \
        issue((uint32_t)(((u) + 6) & 7));                                                                      \
        hfilter((uint32_t)(u), parity, H[(u) & 1]);                                                            \
        if (kDown) { /* scale_y > 1: consecutive output rows end on different source rows */                   \
            if (next_last == n) emit(H[((u) + 1) & 1], H[(u) & 1]);                                            \
        } else {                                                                                               \
            while (next_last == n) emit(H[((u) + 1) & 1], H[(u) & 1]);                                         \
        }                                                                                                      \
        ++n;                                                                                                   \
    }
    while (entry != entry_end) {                       // n & 7 == u
        VACV_LP_STEP(0) VACV_LP_STEP(1) VACV_LP_STEP(2) VACV_LP_STEP(3) VACV_LP_STEP(4) VACV_LP_STEP(5) VACV_LP_STEP(6) VACV_LP_STEP(7)
        parity ^= 1u;
    }
#undef VACV_LP_STEP
}

}  // namespace vacv
