// u8 bicubic resize of interleaved BGR at RATIONAL horizontal scales (config 4: 2560 -> 1920 = 4 source pixels per 3 output
// pixels), third generation of the column walker (resize_cubic3_walkn.cuh).
//
// Reference arithmetic: OpenCV 2.4.13 cv::resize(CV_8UC3, INTER_CUBIC) -- the reference's only u8 cubic path
// (src/cv/resize.cpp:33-36) -- bit for bit, exactly as in the other walkers.
//
// What is different, and why.  ncu of resize_cubic3_walkn_kernel<4> (profiles/r2_cubic_walkn_v2_ncu_raw.txt): 65 thread-instructions
// per output pixel, no pipe above 60 %, issue slots the limiter; more resident warps do not help (register caps of 112 / 104 gave
// 1.5 %).  Of the 65, the horizontal pass of a column and source row costs 4 LDS + 1 IMAD (tap words at a lane-dependent
// address), 3 SHF (lane-dependent byte alignment), 6 PRMT (de-interleave) and 6 IDP -- because a thread's columns lie 32 apart,
// every column fetches and aligns its own 12 bytes.  When w_in : w_out = P : Q the tap pattern repeats every Q output columns
// and P source pixels, so a thread that owns KP whole periods of ADJACENT columns
//   * reads its P*KP + 3 source pixels ONCE per source row (8 words for 6 columns instead of 24) -- at a lane stride of 3*P*KP
//     bytes, which the launcher requires to be a multiple of 4, so the byte alignment inside the words is the same in every lane;
//   * de-interleaves them once into per-channel words and cuts each column's four taps out with PRMTs whose selectors are
//     COMPILE-TIME constants (no funnel shifts, no per-column address arithmetic);
//   * source rows arrive by one cp.async.bulk per warp and row (one lane, mbarrier completion) instead of two LDGSTS per lane;
//   * NV = 3 * Q * KP output bytes per thread and row are packed in registers and staged as 32-bit words.
// The 11-bit coefficients stay per-lane registers computed with the reference's own float arithmetic (the rounding of the source
// coordinate to fp32 makes them NOT exactly periodic), and taps OpenCV clamps onto an edge pixel are folded into the in-range
// positions; the launcher verifies on the host that every column's taps sit where the compile-time pattern expects them and
// falls back to the generic walker otherwise.
#pragma once
#include <type_traits>

#include "period_common.cuh"
#include "resize_cubic3_walk.cuh"

namespace vacv {

struct PeriodGeom {
    int w, h, wo, ho;
    int warp_strips, cta_strips, segs, rows_per_seg;
    int ring_pitch;                    // bytes per ring row of a warp (multiple of 16)
    double scale_x, scale_y;
    size_t src_image, dst_image;       // bytes between images
    f32x2 one2, negzero2, magic2, negmagic2;   // (1,1), (-0,-0), (1.5*2^23)x2, (-1.5*2^23)x2 -- opaque to ptxas on purpose
};


// Geometry shared by the kernel and the launcher (host checks, shared-memory sizes).
template <int P, int Q, int KP>
struct PeriodShape {
    static constexpr int NCOL = Q * KP;                                  // adjacent output columns per thread
    static constexpr int NPX = pd::tap0(P, Q, NCOL - 1) + 4;              // source pixels in a thread's window
    static constexpr int LS = 3 * P * KP;                                // bytes between the windows of neighbouring lanes
    static constexpr int M0 = 1;                                         // the window starts at byte LS*thread - 3: byte 1 of an aligned word
    static constexpr int NW = (M0 + 3 * NPX + 3) / 4;                    // window words
    static constexpr int NPW = (NPX + 3) / 4;                            // words per de-interleaved channel
    static constexpr int NV = 3 * NCOL;                                  // output bytes per thread and row
    static constexpr int NP = NV / 2;                                    // packed fp32 pairs
    static constexpr int kWarpRow = 32 * NV;                             // bytes one warp produces per output row
    static constexpr int kWarpSpan = 32 * LS;                            // source bytes between the windows of neighbouring warps
    static constexpr int kLaneOff = 12;                                  // ring offset of lane 0's window word 0 (the ring row starts 16 bytes early)
    // LS % 4 == 2 (3 : 2 with two periods per thread: 18 bytes): the window of an odd lane starts two bytes later inside its word; odd
    // lanes read from the word before and every lane funnel-shifts its NW + 1 words by 0 or 16 bits -- after that the alignment is
    // lane-invariant again
    static constexpr bool kRealign = LS % 4 != 0;
    static constexpr int kNeed = (LS * 31 + kLaneOff + 4 * (NW + (kRealign ? 1 : 0)) + 15) & ~15;   // ring bytes of one source row
    static_assert(LS % 2 == 0, "lane stride must be even");
    static_assert(NV % 4 == 2 || NV % 4 == 0, "output bytes per thread must be even");
    static_assert((32 * NV) % 16 == 0 && (32 * LS) % 16 == 0, "warp spans must stay 16-byte aligned");
};

template <int P, int Q, int KP, bool kDown, int MAXREG, bool kVStat = false>
__global__ void __maxnreg__(MAXREG) resize_cubic3_period_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, PeriodGeom g) {
    using S = PeriodShape<P, Q, KP>;
    constexpr int NCOL = S::NCOL, NPX = S::NPX, LS = S::LS, M0 = S::M0, NW = S::NW, NPW = S::NPW, NV = S::NV, NP = S::NP;
    constexpr int kWarpRow = S::kWarpRow;
    extern __shared__ __align__(16) uint8_t smem[];
    Walk2Row* rows = reinterpret_cast<Walk2Row*>(smem);                                   // [rows_per_seg + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rows_bytes = (g.rows_per_seg + 1) * (int)sizeof(Walk2Row);
    uint8_t* stage = smem + rows_bytes + warp * (kPdStageRows * kWarpRow);
    uint8_t* ring = smem + rows_bytes + nwarps * (kPdStageRows * kWarpRow) + warp * (kPdRing * g.ring_pitch);
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem + rows_bytes + nwarps * (kPdStageRows * kWarpRow + kPdRing * g.ring_pitch)) + warp * (kPdRing * 8);
    const int cta_strip = blockIdx.x % g.cta_strips, seg = blockIdx.x / g.cta_strips;
    const int wstrip = cta_strip * nwarps + warp;
    const int pt = wstrip * 32 + lane;                 // this thread's index along x: columns NCOL * pt ..
    const int dy_begin = seg * g.rows_per_seg, nrows = min(g.ho, dy_begin + g.rows_per_seg) - dy_begin;
    const uint8_t* img = src + blockIdx.y * g.src_image;
    uint8_t* out_img = dst + blockIdx.y * g.dst_image;
    const unsigned row_bytes = (unsigned)g.w * 3, out_row_bytes = (unsigned)g.wo * 3;

    // walk steps: step n filters source row t_first + n; the last step is the last output row's last tap row
    int t_first, n_stop;
    {
        const float f0 = (float)(((double)dy_begin + 0.5) * g.scale_y - 0.5), f1 = (float)(((double)(dy_begin + nrows - 1) + 0.5) * g.scale_y - 0.5);
        t_first = __shfl_sync(0xffffffffu, (int)floorf(f0) - 1, 0);   // cubic_cv_coord_scaled's row index, without the coefficients
        n_stop = __shfl_sync(0xffffffffu, (int)floorf(f1) + 2, 0) - t_first;
    }
    const bool active = wstrip < g.warp_strips;        // false: padding warp of the last CTA strip

    // the warp's bytes of a source row: [span0, span0 + kNeed) clipped to the row; ring byte r <-> source byte span0 + r.
    // Everything the copy issue needs is warp-uniform; the shuffles tell the compiler so (uniform registers, one UBLKCP per warp).
    constexpr unsigned kPitch = S::kNeed;              // bytes per ring slot
    const int span0 = S::kWarpSpan * wstrip - 16;
    const int lo = __shfl_sync(0xffffffffu, max(span0, 0), 0), hi = __shfl_sync(0xffffffffu, min(span0 + S::kNeed, (int)row_bytes), 0);
    const uint32_t copy_bytes = (uint32_t)(hi - lo);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t ring_dst = __shfl_sync(0xffffffffu, ring_s - (uint32_t)span0, 0) + (uint32_t)lo;
    const uint32_t ubars = __shfl_sync(0xffffffffu, bars, 0);
    // this lane's window word 0 in ring slot 0 (kRealign: odd lanes start at the word before, see PeriodShape)
    const int realign_sh = S::kRealign && (lane & 1) ? 16 : 0;
    const uint32_t win_s = ring_s + (uint32_t)(LS * lane + S::kLaneOff - (S::kRealign && (lane & 1) ? 2 : 0));
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const f32x2 one2 = g.one2, negzero2 = g.negzero2, magic2 = g.magic2, negmagic2 = g.negmagic2;

    // ---- source rows: walk step n (row t_first + n, clamped to the image like OpenCV's tap rows) -> ring slot n & 7, one bulk copy
    //      per warp and row, kPdAhead steps ahead
    int t_pre = t_first;                               // (unclamped) row of the next copy; its walk step is n_pre
    int n_pre = 0;
    const uint8_t* g_pre = src + blockIdx.y * g.src_image + (unsigned)lo + (size_t)(unsigned)min(max(t_first, 0), g.h - 1) * row_bytes;
    auto issue = [&](const uint32_t slot) {            // warp-uniform; slot is a literal at every call site
        if (n_pre <= n_stop) {
            if (lane == 0) {
                pd::mbar_expect_tx(ubars + 8 * slot, copy_bytes);
                pd::bulk_g2s(ring_dst + slot * kPitch, g_pre, copy_bytes, ubars + 8 * slot);
            }
            g_pre += (unsigned)t_pre < (unsigned)(g.h - 1) ? row_bytes : 0u;   // rows below 0 / beyond h-1 repeat the edge row
            ++t_pre;
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

    // per output row: vertical weights, and the walk step that completes it
    for (int r = tid; r <= nrows; r += blockDim.x) {   // entry nrows = sentinel that never matches
        Walk2Row e;
        int s, q[4];
        cubic_cv_coord_scaled(dy_begin + min(r, nrows - 1), g.h, g.scale_y, false, s, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) e.b[2 * j] = e.b[2 * j + 1] = (float)q[j] * (1.f / (2048 * 2048));
        e.last = r < nrows ? s + 2 - t_first : INT_MAX;
        e.pad[0] = e.pad[1] = e.pad[2] = 0;
        rows[r] = e;
    }
    // x taps: column c's four taps sit at window pixels tap0(c) .. tap0(c) + 3 (verified by the launcher); taps OpenCV clamps onto an
    // edge pixel have their integer coefficients added up at that pixel's position (identical sums)
    const bool owner = NCOL * pt < g.wo;               // threads past the last column compute on zero coefficients; their bytes are never flushed
    int c01[NCOL], c23[NCOL];
    pd::static_for<NCOL>([&](auto ic) {
        constexpr int c = decltype(ic)::value;
        int s, q[4], xc[4] = {0, 0, 0, 0};
        cubic_cv_coord_scaled(min(NCOL * pt + c, g.wo - 1), g.w, g.scale_x, true, s, q);
        const int base = P * KP * pt - 1 + pd::tap0(P, Q, c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pos = min(max(s - 1 + j, 0), g.w - 1) - base;   // 0..3
#pragma unroll
            for (int t = 0; t < 4; ++t) xc[t] += (owner && pos == t) ? q[j] : 0;
        }
        c01[c] = (xc[0] & 0xffff) | (xc[1] << 16);
        c23[c] = (xc[2] & 0xffff) | (xc[3] << 16);
    });
    __syncthreads();
    if (!active) return;                               // (no CTA barrier below)

    uint32_t entry = rows_s;                           // shared address of the next output row's table entry
    int next_last;
    asm volatile("ld.shared.s32 %0, [%1+32];" : "=r"(next_last) : "r"(entry));

    // horizontal pass of the walk step in ring slot `slot`: (value 2i, value 2i+1) of the thread's NV = 3 * NCOL output bytes, exact fp32
    // hload: wait for the row in ring slot `slot` and read this lane's window words; hcompute: the arithmetic on them
    auto hload = [&](const uint32_t slot, uint32_t parity, uint32_t (&W)[NW]) {
        pd::mbar_wait(ubars + 8 * slot, parity);
        const uint32_t p = win_s + slot * kPitch;
        if constexpr (S::kRealign) {                   // NW + 1 words from the (lane-dependent) aligned address, shifted into place
            uint32_t X[NW + 1];
#pragma unroll
            for (int i = 0; i <= NW; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(X[i]) : "r"(p + 4 * i));
#pragma unroll
            for (int i = 0; i < NW; ++i) W[i] = __funnelshift_r(X[i], X[i + 1], realign_sh);
        } else if constexpr (LS % 8 == 0 && S::kLaneOff % 8 == 4) {   // word 0 alone, then 8-byte aligned pairs
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W[0]) : "r"(p));
#pragma unroll
            for (int i = 1; i + 1 < NW; i += 2) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(W[i]), "=r"(W[i + 1]) : "r"(p + 4 * i));
            if constexpr ((NW & 1) == 0) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W[NW - 1]) : "r"(p + 4 * (NW - 1)));
        } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W[i]) : "r"(p + 4 * i));
        }
    };
    auto hcompute = [&](const uint32_t (&W)[NW], f32x2 (&H)[NP]) {
        // de-interleave: channel k, pixels 4i .. 4i+3 of the window (pixels past the window repeat the last one: never used)
        uint32_t pl[3][NPW];
        pd::static_for<3>([&](auto ik) {
            constexpr int k = decltype(ik)::value;
            pd::static_for<NPW>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                constexpr int j0 = 4 * i, j1 = 4 * i + 1 < NPX ? 4 * i + 1 : NPX - 1, j2 = 4 * i + 2 < NPX ? 4 * i + 2 : NPX - 1, j3 = 4 * i + 3 < NPX ? 4 * i + 3 : NPX - 1;
                pl[k][i] = pd::gather4<M0 + 3 * j0 + k, M0 + 3 * j1 + k, M0 + 3 * j2 + k, M0 + 3 * j3 + k>(W);
            });
        });
        int hv[NV];
        pd::static_for<NCOL>([&](auto ic) {
            constexpr int c = decltype(ic)::value;
            constexpr int t0 = pd::tap0(P, Q, c);
            pd::static_for<3>([&](auto ik) {
                constexpr int k = decltype(ik)::value;
                const uint32_t taps = pd::gather4<t0, t0 + 1, t0 + 2, t0 + 3>(pl[k]);
                // sum(tap * coef) on top of the bit pattern of 1.5*2^23 (exact int -> float for |H| < 2^22 after subtracting it)
                hv[3 * c + k] = dp2a_hi_su(c23[c], taps, dp2a_lo_su(c01[c], taps, 0x4B400000));
            });
        });
#pragma unroll
        for (int i = 0; i < NP; ++i) H[i] = fma2(pack2i(hv[2 * i], hv[2 * i + 1]), one2, negmagic2);
    };
    auto hfilter = [&](const uint32_t slot, uint32_t parity, f32x2 (&H)[NP]) {
        uint32_t W[NW];
        hload(slot, parity, W);
        hcompute(W, H);
    };

    // ---- output: every row is staged in one of two kWarpRow-byte buffers of the warp and leaves at once as lane-contiguous 16-byte
    //      chunks (one warp-wide store of 512 bytes + the remainder).  A thread's NV bytes start at NV * lane: word aligned in even
    //      lanes, 2 bytes off in odd lanes (NV % 4 == 2) -- odd lanes store their first two bytes as a half word and the rest shifted by
    //      16 bits, so that everything else is a 32-bit store.  Threads past the last column stage zeros that are never flushed.
    constexpr bool kHalf = NV % 4 == 2;
    constexpr int NWD = NV / 4;                        // full words per thread and row
    constexpr int kChunks = kWarpRow / 16;             // 16-byte chunks per row of the warp
    static_assert(kChunks <= 64, "flush: one or two chunks per lane");
    const bool odd = kHalf && (lane & 1);
    const int sh16 = odd ? 16 : 0;
    uint32_t st_w = stage_s + NV * lane + (odd ? 2 : 0);            // the 32-bit stores
    uint32_t st_h = stage_s + NV * lane + (odd ? 0 : 4 * NWD);      // the 16-bit store
    uint32_t st_f = stage_s + 16 * lane;                            // this lane's first chunk of the staged row
    int st_d = kWarpRow;                                            // distance to the other staging buffer
    const int valid_chunks = min(kWarpRow, (int)out_row_bytes - wstrip * kWarpRow) >> 4;        // chunks of a warp row inside the image row
    const bool f0 = lane < min(valid_chunks, kChunks), f1 = kChunks > 32 && lane + 32 < min(valid_chunks, kChunks);
    uint8_t* gflush = out_img + (size_t)dy_begin * out_row_bytes + (size_t)wstrip * kWarpRow + 16 * lane;   // this lane's first chunk in global memory

    // vertical pass + store of the NV output bytes from the window (h0 = oldest row)
    auto emit = [&](const f32x2 (&h0)[NP], const f32x2 (&h1)[NP], const f32x2 (&h2)[NP], const f32x2 (&h3)[NP]) {
        f32x2 w0, w1, w2, w3;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(entry));
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+16];" : "=l"(w2), "=l"(w3) : "r"(entry));
        uint32_t pr[NP + 1];                           // two output bytes each (low half)
#pragma unroll
        for (int i = 0; i < NP; ++i) {   // OpenCV's SSE2 body: mulps, addps (one rounding each), cvtps2dq (half-even), packs, packus
            f32x2 f = fma2(h0[i], w0, negzero2);
            f = fma2(f, one2, fma2(h1[i], w1, negzero2));
            f = fma2(f, one2, fma2(h2[i], w2, negzero2));
            f = fma2(f, one2, fma2(h3[i], w3, negzero2));
            // |f| < 2^22: adding 1.5*2^23 rounds half-to-even at integer granularity; subtracting its bit pattern and clamping
            // to [0,255] is one DPX op per value (the intermediate s16 saturation of packs cannot change the result)
            f = fma2(f, one2, magic2);
            int lo, hi;
            unpack2i(f, lo, hi);
            pr[i] = __byte_perm(__viaddmin_s32_relu(lo, -0x4B400000, 255), __viaddmin_s32_relu(hi, -0x4B400000, 255), 0x0040);
        }
        pr[NP] = 0;
        uint32_t A[NWD + 1];
#pragma unroll
        for (int j = 0; j < NWD; ++j) A[j] = __byte_perm(pr[2 * j], pr[2 * j + 1], 0x5410);
        A[NWD] = pr[2 * NWD];                          // kHalf: the last two bytes
#pragma unroll
        for (int j = 0; j < NWD; ++j) {
            const uint32_t x = kHalf ? __funnelshift_r(A[j], A[j + 1], sh16) : A[j];
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(st_w + 4 * j), "r"(x) : "memory");
        }
        if (kHalf) {
            const uint32_t x = odd ? A[0] : A[NWD];
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(st_h), "h"((unsigned short)x) : "memory");
        }
        __syncwarp();                                  // the row is staged; the other buffer's readers passed this point a row ago
        if (f0) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(st_f));
            st_stream16(gflush, v);
        }
        if (kChunks > 32 && f1) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(st_f));
            st_stream16(gflush + 512, v);
        }
        gflush += out_row_bytes;
        st_w += st_d; st_h += st_d; st_f += st_d;                 // the other buffer
        st_d = -st_d;
        entry += (int)sizeof(Walk2Row);
        asm volatile("ld.shared.s32 %0, [%1+32];" : "=r"(next_last) : "r"(entry));   // sentinel INT_MAX after the last row
    };

    // ---- the walk: step n -> ring slot n & 7 and window slot n & 3 (compile-time inside the 8x unrolled body); an output row is emitted
    //      as soon as its last tap row has been filtered.  The first output row completes at step 3.
    f32x2 H[4][NP];
    const uint32_t entry_end = rows_s + nrows * (int)sizeof(Walk2Row);
#define VACV_PD_IC(k) (uint32_t)(k)
    int n = 0;
    uint32_t parity = 0;
#define VACV_PD_STEP(u)                                                                                        \
    if (entry != entry_end) {                                                                                  \
        __syncwarp(); /* every lane is done with step n - 2, whose ring slot the next copy overwrites */       \
        issue(VACV_PD_IC(((u) + 6) & 7));                                                                      \
        hfilter(VACV_PD_IC(u), parity, H[(u) & 3]);                                                            \
        if (kDown) { /* scale_y >= 1: consecutive output rows end on different source rows */                  \
            if (next_last == n) emit(H[((u) + 1) & 3], H[((u) + 2) & 3], H[((u) + 3) & 3], H[(u) & 3]);        \
        } else {                                                                                               \
            while (next_last == n) emit(H[((u) + 1) & 3], H[((u) + 2) & 3], H[((u) + 3) & 3], H[(u) & 3]);     \
        }                                                                                                      \
        ++n;                                                                                                   \
    }
    if (!kVStat) {
        while (entry != entry_end) {                   // n & 7 == u
            VACV_PD_STEP(0) VACV_PD_STEP(1) VACV_PD_STEP(2) VACV_PD_STEP(3) VACV_PD_STEP(4) VACV_PD_STEP(5) VACV_PD_STEP(6) VACV_PD_STEP(7)
            parity ^= 1u;
        }
    } else {
        // kVStat (the launcher verified it): output row r of the segment completes at step 3 + r + r / 3 -- the 4 : 3 pattern, every step
        // with n & 3 != 2 completes one row.  The row completed at step n - 1 is blended and stored DURING step n, after that step's
        // window words were requested and in the same basic block as its horizontal pass: the FFMA2 chain of one and the PRMT / IDP
        // work of the other are independent, and no data-dependent branch separates them any more.
        // slots at step n (u = n & 7): rows n-4 .. n-1 live in window slots u & 3, (u+1) & 3, (u+2) & 3, (u+3) & 3; row n goes to u & 3
#define VACV_PD_STEP_S(u, EMITPREV)                                                                            \
        if (n <= n_stop) {                                                                                     \
            __syncwarp();                                                                                      \
            issue(VACV_PD_IC(((u) + 6) & 7));                                                                  \
            uint32_t W[NW];                                                                                    \
            hload(VACV_PD_IC(u), parity, W);                                                                   \
            if (EMITPREV) emit(H[(u) & 3], H[((u) + 1) & 3], H[((u) + 2) & 3], H[((u) + 3) & 3]);             \
            hcompute(W, H[(u) & 3]);                                                                           \
            ++n;                                                                                               \
        }
        VACV_PD_STEP_S(0, false) VACV_PD_STEP_S(1, false) VACV_PD_STEP_S(2, false) VACV_PD_STEP_S(3, false)   // rows 0 .. 3 of the walk
        while (n <= n_stop) {                          // enters with n & 7 == 4; a step emits the previous step's row unless (n - 1) & 3 == 2
            VACV_PD_STEP_S(4, true) VACV_PD_STEP_S(5, true) VACV_PD_STEP_S(6, true) VACV_PD_STEP_S(7, false)
            parity ^= 1u;
            VACV_PD_STEP_S(0, true) VACV_PD_STEP_S(1, true) VACV_PD_STEP_S(2, true) VACV_PD_STEP_S(3, false)
        }
        // the row completed by the last step (n_stop & 3 != 2 by construction); n == n_stop + 1 here
        {
            const int u = n & 3;
            if (u == 0) emit(H[0], H[1], H[2], H[3]);
            else if (u == 1) emit(H[1], H[2], H[3], H[0]);
            else if (u == 2) emit(H[2], H[3], H[0], H[1]);
            else emit(H[3], H[0], H[1], H[2]);
        }
#undef VACV_PD_STEP_S
    }
#undef VACV_PD_STEP
#undef VACV_PD_IC
}

}  // namespace vacv
