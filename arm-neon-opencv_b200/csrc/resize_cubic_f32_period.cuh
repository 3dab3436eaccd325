// fp32 bicubic resize at RATIONAL horizontal scales -- C = 1: single planes (CHW tensors are resized plane by plane), C = 3: interleaved
// BGR -- the periodic form of the column walker (resize_cubic3_walk.cuh), as resize_cubic3_period.cuh is for u8 BGR.
//
// Reference arithmetic: resize_naive_inter_cubic_fp32_one_channel / _three_channel (src/cv/resize_naive.cpp:368-529 / :187-366;
// coefficients and border folding :130-185, horizontal / vertical accumulation order :230 / :345), fp32 with one rounding per operation (built --fmad=false).
//
// Why.  ncu of resize_cubic3_walk_f32_kernel<1> at 1080p -> 1280x720 (profiles/r2_cubic_f32_chw_ncu_raw.txt): the shared-memory data
// pipe at 87 % of its wavefront peak, short-scoreboard the top stall -- one column per thread reads its four taps with four LDS.32
// at a lane stride of 1.5 floats (two wavefronts each), and pays a whole row's bookkeeping for four multiplies.  When
// w_in : w_out = P : Q the tap pattern repeats every Q columns / P source pixels: a thread that owns KP periods of ADJACENT columns
// reads its P*KP + 3 source floats once per row as 64-bit words (14 floats for 8 columns at 3 : 2 instead of 32), every tap is then a
// register with a compile-time index, and its Q*KP results leave as 16-byte stores straight from registers (no staging).
// Coefficients are per-lane registers computed with the reference's arithmetic; where the reference folds border taps
// (:154-181) they land on the window positions of the pixels they multiply.  The launcher verifies on the host that every tap with a
// non-zero coefficient lies inside the column's compile-time window and falls back to the generic walker otherwise.  Ring bytes
// outside the copied part of a row (left of pixel 0, right of pixel w-1) are zeroed once: their coefficients are zero, and 0 * 0
// must not be 0 * NaN.  Periods per thread: enough that a warp's bulk copy of a source row is >= 1.5 KB -- with 800-byte copies (3 : 2,
// two periods) the kernel ran at 3.0 TB/s whenever more than ~13 warps per SM were resident, with 1.5 KB copies at 5.1 - 6.3 TB/s.
#pragma once
#include "resize_cubic3_period.cuh"

namespace vacv {

struct PeriodF32Geom {
    int w, h, wo, ho;
    int warp_strips, cta_strips, segs, rows_per_seg;
    double scale_x, scale_y;
    size_t src_image, dst_image;       // floats between planes
};

template <int C, int P, int Q, int KP>
struct PeriodF32Shape {
    static constexpr int NCOL = Q * KP;                                  // adjacent output columns per thread
    static constexpr int NPX = pd::tap0(P, Q, NCOL - 1) + 4;              // source pixels in a thread's window
    static constexpr int NWF = C * NPX;                                  // ... as floats
    static constexpr int NV = C * NCOL;                                  // output floats per thread and row
    static constexpr int LS = 4 * C * P * KP;                            // bytes between the windows of neighbouring lanes
    static constexpr int kWarpSpan = 32 * LS;                            // source bytes between the windows of neighbouring warps
    static constexpr int kLaneOff = 16 - 4 * C;                          // ring offset of lane 0's window float 0 (source pixel P*KP*thread - 1; the ring row starts 16 bytes early)
    static constexpr int kNeed = (LS * 31 + kLaneOff + 4 * (NWF + 1) + 15) & ~15;   // ring bytes of one source row (the 64-bit loads may read one float past the window)
    static constexpr bool kDirect = NV % 4 == 0;                         // results leave as 16-byte stores straight from registers, else through a staged row
    static constexpr int kWarpRow = 32 * NV * 4;                         // bytes one warp produces per output row
    static_assert(C >= 1 && C <= 3, "window float 0 must lie inside the 16 bytes before the warp's span");
    static_assert(LS % 8 == 0 && kLaneOff % 8 == 4, "window float 0 is read alone, floats 1.. as 64-bit words");
    static_assert(kWarpRow % 16 == 0 && kWarpRow / 16 <= 96, "flush: up to three 16-byte chunks per lane");
};

template <int C, int P, int Q, int KP, bool kDown>
__global__ void __launch_bounds__(128) resize_cubic_f32_period_kernel(const float* __restrict__ src, float* __restrict__ dst, PeriodF32Geom g) {
    using S = PeriodF32Shape<C, P, Q, KP>;
    constexpr int NCOL = S::NCOL, NWF = S::NWF, NV = S::NV, LS = S::LS, kWarpRow = S::kWarpRow;
    constexpr unsigned kPitch = S::kNeed;
    extern __shared__ __align__(16) uint8_t smem[];
    WalkRow* rows = reinterpret_cast<WalkRow*>(smem);                                     // [rows_per_seg + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rows_bytes = (g.rows_per_seg + 1) * (int)sizeof(WalkRow);
    constexpr int kStage = S::kDirect ? 0 : 2 * kWarpRow;              // two staging buffers per warp (one output row each)
    uint8_t* ring = smem + rows_bytes + warp * (kPdRing * kPitch);
    uint8_t* stage = smem + rows_bytes + nwarps * (kPdRing * kPitch) + warp * kStage;
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem + rows_bytes + nwarps * (kPdRing * kPitch + kStage)) + warp * (kPdRing * 8);
    const int cta_strip = blockIdx.x % g.cta_strips, seg = blockIdx.x / g.cta_strips;
    const int wstrip = cta_strip * nwarps + warp;
    const int pt = wstrip * 32 + lane;                 // this thread's index along x: columns NCOL * pt ..
    const int dy_begin = seg * g.rows_per_seg, nrows = min(g.ho, dy_begin + g.rows_per_seg) - dy_begin;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(src + blockIdx.y * g.src_image);
    float* out_img = dst + blockIdx.y * g.dst_image;
    const unsigned row_bytes = (unsigned)g.w * 4u * C;
    const bool active = wstrip < g.warp_strips;        // false: padding warp of the last CTA strip

    // walk steps: step n filters source row t_first + n; the last step is the last output row's last tap row (ofs + 2)
    int t_first, n_stop;
    {
        int o0, o1; float a[4];
        cubic_naive_scaled(dy_begin, g.h, g.scale_y, o0, a);
        cubic_naive_scaled(dy_begin + nrows - 1, g.h, g.scale_y, o1, a);
        t_first = __shfl_sync(0xffffffffu, o0 - 1, 0);
        n_stop = __shfl_sync(0xffffffffu, o1 + 2, 0) - t_first;
    }
    // the warp's bytes of a source row: [span0, span0 + kNeed) clipped to the row; ring byte r <-> source byte span0 + r
    const int span0 = S::kWarpSpan * wstrip - 16;
    const int lo = __shfl_sync(0xffffffffu, max(span0, 0), 0), hi = __shfl_sync(0xffffffffu, min(span0 + S::kNeed, (int)row_bytes), 0);
    const uint32_t copy_bytes = (uint32_t)(hi - lo);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t ring_dst = __shfl_sync(0xffffffffu, ring_s - (uint32_t)span0, 0) + (uint32_t)lo;
    const uint32_t ubars = __shfl_sync(0xffffffffu, bars, 0);
    const uint32_t win_s = ring_s + (uint32_t)(LS * lane + S::kLaneOff);
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);

    int t_pre = t_first, n_pre = 0;
    const uint8_t* g_pre = img + (unsigned)lo + (size_t)(unsigned)min(max(t_first, 0), g.h - 1) * row_bytes;
    auto issue = [&](const uint32_t slot) {            // warp-uniform; slot is a literal at every call site
        if (n_pre <= n_stop) {
            if (lane == 0) {
                pd::mbar_expect_tx(ubars + 8 * slot, copy_bytes);
                pd::bulk_g2s(ring_dst + slot * kPitch, g_pre, copy_bytes, ubars + 8 * slot);
            }
            g_pre += (unsigned)t_pre < (unsigned)(g.h - 1) ? row_bytes : 0u;   // tap rows are in range by construction; the look-ahead is clamped
            ++t_pre;
        }
        ++n_pre;
    };
    if (active) {
        // zero what no copy ever writes (left of the row / right of it), then the first rows go in flight
        for (int sl = 0; sl < kPdRing; ++sl) {
            for (int b = 4 * lane; b < (int)kPitch; b += 128)
                if (b < lo - span0 || b >= hi - span0) *reinterpret_cast<uint32_t*>(ring + sl * kPitch + b) = 0u;
        }
        if (lane == 0) {
            for (int i = 0; i < kPdRing; ++i) pd::mbar_init(bars + 8 * i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zeros (generic proxy) before the copies (async proxy) next to them
        __syncwarp();
        static_assert(kPdAhead == 6 && kPdRing == 8, "the unrolled walk below assumes 6 rows ahead in an 8-slot ring");
        issue(0); issue(1); issue(2); issue(3); issue(4); issue(5);
    }
    // per output row: vertical weights and the walk step that completes it
    for (int r = tid; r <= nrows; r += blockDim.x) {   // entry nrows = sentinel that never matches
        WalkRow e;
        int ofs;
        cubic_naive_scaled(dy_begin + min(r, nrows - 1), g.h, g.scale_y, ofs, e.b);
        e.last = r < nrows ? ofs + 2 - t_first : INT_MAX;
        e.pad[0] = e.pad[1] = e.pad[2] = 0;
        rows[r] = e;
    }
    // x taps: column c's window = source floats base .. base + 3 with base = P*KP*pt - 1 + tap0(c); the reference's taps ofs-1 .. ofs+2
    // (ofs shifted and coefficients folded at the borders) land on those positions -- a tap outside has coefficient 0 (launcher check)
    const bool owner = NCOL * pt < g.wo;
    float xa[NCOL][4];
    pd::static_for<NCOL>([&](auto ic) {
        constexpr int c = decltype(ic)::value;
        int ofs; float a[4];
        cubic_naive_scaled(min(NCOL * pt + c, g.wo - 1), g.w, g.scale_x, ofs, a);
        const int base = P * KP * pt - 1 + pd::tap0(P, Q, c);
#pragma unroll
        for (int t = 0; t < 4; ++t) xa[c][t] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pos = ofs - 1 + j - base;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (owner && pos == t) xa[c][t] = a[j];
        }
    });
    __syncthreads();
    if (!active) return;                               // (no CTA barrier below)

    uint32_t entry = rows_s;                           // shared address of the next output row's table entry
    int next_last;
    asm volatile("ld.shared.s32 %0, [%1+16];" : "=r"(next_last) : "r"(entry));

    // horizontal pass of the walk step in ring slot `slot`: window floats -> the NV sums (value C*c + k), resize_naive.cpp:230 order
    auto hfilter = [&](const uint32_t slot, uint32_t parity, float (&H)[NV]) {
        pd::mbar_wait(ubars + 8 * slot, parity);
        const uint32_t p = win_s + slot * kPitch;
        float W[NWF + 1];
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(W[0]) : "r"(p));
#pragma unroll
        for (int i = 1; i + 1 <= NWF; i += 2) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(W[i]), "=f"(W[i + 1]) : "r"(p + 4 * i));
        pd::static_for<NCOL>([&](auto ic) {
            constexpr int c = decltype(ic)::value;
            constexpr int t0 = pd::tap0(P, Q, c);
#pragma unroll
            for (int k = 0; k < C; ++k)
                H[C * c + k] = W[C * t0 + k] * xa[c][0] + W[C * (t0 + 1) + k] * xa[c][1] + W[C * (t0 + 2) + k] * xa[c][2] + W[C * (t0 + 3) + k] * xa[c][3];
        });
    };
    // vertical pass (resize_naive.cpp:345 order) + store: NV adjacent floats per thread -- 16-byte stores straight from registers when NV
    // is a multiple of 4, else staged in one of two buffers of the warp and flushed at once as lane-contiguous 16-byte chunks
    float* orow = out_img + ((size_t)dy_begin * g.wo + (size_t)NCOL * pt) * C;
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    uint32_t st_w = stage_s + 4 * NV * lane, st_f = stage_s + 16 * lane;
    int st_d = kWarpRow;
    constexpr int kChunks = kWarpRow / 16;
    const int valid_chunks = min(kWarpRow, (int)((unsigned)g.wo * 4u * C) - wstrip * kWarpRow) >> 4;   // chunks of a warp row inside the image row
    uint8_t* gflush = reinterpret_cast<uint8_t*>(out_img + ((size_t)dy_begin * g.wo) * C) + (size_t)wstrip * kWarpRow + 16 * lane;
    auto emit = [&](const float (&h0)[NV], const float (&h1)[NV], const float (&h2)[NV], const float (&h3)[NV]) {
        float4 bw;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bw.x), "=f"(bw.y), "=f"(bw.z), "=f"(bw.w) : "r"(entry));
        float o[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) o[v] = h0[v] * bw.x + h1[v] * bw.y + h2[v] * bw.z + h3[v] * bw.w;
        if (S::kDirect) {
            if (owner) {
#pragma unroll
                for (int q = 0; q < NV / 4; ++q) st_stream16f(orow + 4 * q, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
            }
            orow += (size_t)g.wo * C;
        } else {                                       // threads past the last column stage zeros that are never flushed
#pragma unroll
            for (int v = 0; v < NV; ++v) asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_w + 4 * v), "f"(o[v]) : "memory");
            __syncwarp();                              // the row is staged; the other buffer's readers passed this point a row ago
#pragma unroll
            for (int q = 0; q < (kChunks + 31) / 32; ++q) {
                if (lane + 32 * q < min(valid_chunks, kChunks)) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(st_f + 512 * q));
                    st_stream16(gflush + 512 * q, v);
                }
            }
            gflush += (size_t)g.wo * 4u * C;
            st_w += st_d; st_f += st_d;
            st_d = -st_d;
        }
        entry += (int)sizeof(WalkRow);
        asm volatile("ld.shared.s32 %0, [%1+16];" : "=r"(next_last) : "r"(entry));   // sentinel INT_MAX after the last row
    };

    // ---- the walk: step n -> ring slot n & 7 and window slot n & 3 (compile-time inside the 8x unrolled body)
    float H[4][NV];
    const uint32_t entry_end = rows_s + nrows * (int)sizeof(WalkRow);
    int n = 0;
    uint32_t parity = 0;
#define VACV_PF_STEP(u)                                                                                        \
    if (entry != entry_end) {                                                                                  \
        __syncwarp(); /* every lane is done with step n - 2, whose ring slot the next copy overwrites */       \
        issue((uint32_t)(((u) + 6) & 7));                                                                      \
        hfilter((uint32_t)(u), parity, H[(u) & 3]);                                                            \
        if (kDown) { /* scale_y >= 1: consecutive output rows end on different source rows */                  \
            if (next_last == n) emit(H[((u) + 1) & 3], H[((u) + 2) & 3], H[((u) + 3) & 3], H[(u) & 3]);        \
        } else {                                                                                               \
            while (next_last == n) emit(H[((u) + 1) & 3], H[((u) + 2) & 3], H[((u) + 3) & 3], H[(u) & 3]);     \
        }                                                                                                      \
        ++n;                                                                                                   \
    }
    while (entry != entry_end) {                       // n & 7 == u
        VACV_PF_STEP(0) VACV_PF_STEP(1) VACV_PF_STEP(2) VACV_PF_STEP(3) VACV_PF_STEP(4) VACV_PF_STEP(5) VACV_PF_STEP(6) VACV_PF_STEP(7)
        parity ^= 1u;
    }
#undef VACV_PF_STEP
}

}  // namespace vacv
