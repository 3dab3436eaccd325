// u8 bicubic resize of interleaved BGR (config 4), second generation of the column walker (resize_cubic3_walk.cuh).
//
// Reference arithmetic: OpenCV 2.4.13 cv::resize(CV_8UC3, INTER_CUBIC) -- the reference's only u8 cubic path
// (src/cv/resize.cpp:33-36) -- bit for bit: int32 horizontal pass with 11-bit coefficients, fp32 vertical pass
// with one rounding per multiply and per add, round-half-even, saturate.
//
// What changed against resize_cubic3_walk2_kernel, and why (ncu of that kernel at round-2 HEAD,
// profiles/r2_cubic_walk2_base_ncu_raw.txt: 89 thread-instructions per output pixel, of which only ~48 are the arithmetic
// itself -- 15 ALU-pipe ops per column and source row for the horizontal pass, 16 per pixel for the vertical pass -- and the
// rest is per-row bookkeeping paid once per THREAD and source row: prefetch address arithmetic, ring and tap addressing,
// loop control, staging).  The per-row bookkeeping does not depend on how many columns a thread owns, so:
//   * a thread owns NC = 4 columns (32 apart, so that every tap read of a warp still covers one contiguous byte span) instead of
//     2: the bookkeeping per pixel halves; a warp produces 128 adjacent pixels = 384 contiguous bytes per output row;
//   * the prefetch pointer advances incrementally (one compare + add) instead of being rebuilt from the row index each time;
//   * down-scaling (every source row completes at most one output row) gets a straight `if` instead of the `while` loop;
//   * the lane -> (staged row, 16-byte chunk) map of the output flush is computed once, not per flush;
//   * a CTA has as many warps (2..4) as divide the image width with the least padding (1920 px = 15 warp strips = 5 CTAs x 3).
// Per-pixel arithmetic is untouched, so are the results.
#pragma once
#include "resize_cubic3_walk.cuh"

namespace vacv {

constexpr int kWnRing = 8, kWnAhead = 6;     // ring slots per warp / rows in flight ahead of the one being filtered
constexpr int kWnStageRows = 4;

struct WalkNGeom {
    int w, h, wo, ho;
    int warp_strips, cta_strips, segs, rows_per_seg;
    int store16;                       // destination rows / images are 16-byte aligned
    int ring_pitch;                    // bytes per ring row of a warp (multiple of 16, >= the widest warp span)
    double scale_x, scale_y;
    size_t src_image, dst_image;       // bytes between images
    f32x2 one2, negzero2, magic2, negmagic2;   // (1,1), (-0,-0), (1.5*2^23)x2, (-1.5*2^23)x2 -- opaque to ptxas on purpose
};

template <int NC, bool kDown, int MAXREG = (NC == 4 ? 128 : 80), bool kPre = true>
__global__ void __maxnreg__(MAXREG) resize_cubic3_walkn_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, WalkNGeom g) {
    constexpr int kWarpCols = 32 * NC;
    constexpr int kWarpRow = kWarpCols * 3;            // bytes one warp produces per output row
    constexpr int NP = NC / 2;                         // column pairs (packed fp32 lanes)
    extern __shared__ __align__(16) uint8_t smem[];
    Walk2Row* rows = reinterpret_cast<Walk2Row*>(smem);                                   // [rows_per_seg + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rows_bytes = (g.rows_per_seg + 1) * (int)sizeof(Walk2Row);
    uint8_t* stage = smem + rows_bytes + warp * (kWnStageRows * kWarpRow);
    uint8_t* ring = smem + rows_bytes + nwarps * (kWnStageRows * kWarpRow) + warp * (kWnRing * g.ring_pitch);
    const int cta_strip = blockIdx.x % g.cta_strips, seg = blockIdx.x / g.cta_strips;
    const int wstrip = cta_strip * nwarps + warp;
    const int dx_warp = wstrip * kWarpCols;
    const int dx = dx_warp + lane;                     // columns dx + 32 c
    const int dy_begin = seg * g.rows_per_seg, nrows = min(g.ho, dy_begin + g.rows_per_seg) - dy_begin;
    const uint8_t* img = src + blockIdx.y * g.src_image;
    uint8_t* out_img = dst + blockIdx.y * g.dst_image;
    const unsigned row_bytes = (unsigned)g.w * 3, out_row_bytes = (unsigned)g.wo * 3;

    for (int r = tid; r <= nrows; r += blockDim.x) {   // entry nrows = sentinel that never matches
        Walk2Row e;
        int s, q[4];
        cubic_cv_coord_scaled(dy_begin + min(r, nrows - 1), g.h, g.scale_y, false, s, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) e.b[2 * j] = e.b[2 * j + 1] = (float)q[j] * (1.f / (2048 * 2048));
        e.last = r < nrows ? s + 2 : INT_MAX;
        e.pad[0] = e.pad[1] = e.pad[2] = 0;
        rows[r] = e;
    }
    __syncthreads();
    if (wstrip >= g.warp_strips) return;               // padding warp of the last CTA strip (no CTA barrier below)

    // x taps: four CONSECUTIVE source pixels from x_first per column; taps OpenCV clamps onto the edge pixel have their integer
    // coefficients added up (identical sums)
    int sh[NC], c01[NC], c23[NC], aw[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
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
        aw[c] = a0 & ~3;
        sh[c] = (a0 & 3) * 8;
        c01[c] = (xc[0] & 0xffff) | (xc[1] << 16);
        c23[c] = (xc[2] & 0xffff) | (xc[3] << 16);
    }
    // the warp's byte span of a source row = [span0, span0 + 16 * nchunk), 16-byte aligned (x_first is monotone in dx)
    const int span0 = __shfl_sync(0xffffffffu, aw[0], 0) & ~15;
    const int nchunk = (__shfl_sync(0xffffffffu, aw[NC - 1] + (sh[NC - 1] ? 16 : 12), 31) - span0 + 15) >> 4;
    if (nchunk * 16 + 16 > g.ring_pitch || nchunk > 64) __trap();   // the launcher's bound on the span is wrong: fail loudly
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    uint32_t tap_s[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) tap_s[c] = ring_s + (uint32_t)(aw[c] - span0);
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const f32x2 one2 = g.one2, negzero2 = g.negzero2, magic2 = g.magic2, negmagic2 = g.negmagic2;
    const bool c0 = lane < nchunk, c1 = lane + 32 < nchunk;
    const uint32_t ring_lane = ring_s + 16 * lane;
    const unsigned pitch = (unsigned)g.ring_pitch;

    // ---- source rows: row t + kWnAhead goes into ring slot (t + kWnAhead) & 7 with cp.async; OpenCV clamps tap rows to the image
    uint32_t entry = rows_s;                           // shared address of the next output row's table entry
    int next_last, t;
    asm volatile("ld.shared.s32 %0, [%1+32];" : "=r"(next_last) : "r"(entry));
    t = next_last - 3;
    const uint8_t* gpre = img + span0 + 16 * lane + (size_t)(unsigned)min(max(t, 0), g.h - 1) * row_bytes;   // row being prefetched next
    int tpre = t;                                      // its (unclamped) row index
    auto prefetch = [&]() {
        const uint32_t slot = ring_lane + (uint32_t)(tpre & (kWnRing - 1)) * pitch;
        if (c0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(gpre) : "memory");
        if (c1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + 512), "l"(gpre + 512) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        gpre += (unsigned)tpre < (unsigned)(g.h - 1) ? row_bytes : 0u;   // rows below 0 / beyond h-1 repeat the edge row
        ++tpre;
    };
    uint32_t pf[NC][4];
    auto fetch = [&](int tt) {                         // row tt has landed in the ring -> its tap words
        asm volatile("cp.async.wait_group %0;" ::"n"(kWnAhead) : "memory");
        __syncwarp();                                  // every lane's copies of row tt are visible; everyone is done with row tt - 2's slot
        const uint32_t off = (uint32_t)(tt & (kWnRing - 1)) * pitch;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pf[c][i]) : "r"(tap_s[c] + off + 4 * i));
        }
    };
    auto hfilter = [&](f32x2 (&H)[3][NP]) {            // horizontal pass of the fetched row: (column 2p, column 2p+1) per channel
        int hb[NC], hg[NC], hr[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
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
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            H[0][p] = fma2(pack2i(hb[2 * p], hb[2 * p + 1]), one2, negmagic2);
            H[1][p] = fma2(pack2i(hg[2 * p], hg[2 * p + 1]), one2, negmagic2);
            H[2][p] = fma2(pack2i(hr[2 * p], hr[2 * p + 1]), one2, negmagic2);
        }
    };

    // ---- output staging: 4 rows x kWarpRow bytes per warp, flushed as lane-contiguous 16-byte chunks
    const bool staged_store = g.store16 && dx_warp + kWarpCols <= g.wo;
    constexpr int kChunks = kWarpRow / 16;             // 16-byte chunks per staged row (6 NC)
    constexpr int kFlushIters = (kWnStageRows * kChunks + 31) / 32;
    uint32_t fl_s[kFlushIters];                        // per lane and flush iteration: staged row, shared offset, global offset
    unsigned fl_g[kFlushIters];
    int fl_row[kFlushIters];
#pragma unroll
    for (int j = 0; j < kFlushIters; ++j) {
        const int i = lane + 32 * j, ty = i / kChunks, q = i - ty * kChunks;
        fl_row[j] = ty < kWnStageRows ? ty : INT_MAX;
        fl_s[j] = stage_s + ty * kWarpRow + 16 * q;
        fl_g[j] = (unsigned)ty * out_row_bytes + 16u * q;
    }
    int staged = 0;
    uint32_t sp = stage_s + 3 * lane;
    uint8_t* gflush = out_img + (size_t)dy_begin * out_row_bytes + (size_t)dx_warp * 3;   // first staged row of this warp in global memory
    uint8_t* gdirect = out_img + (size_t)dy_begin * out_row_bytes + (size_t)dx * 3;
    auto flush = [&]() {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kFlushIters; ++j)
            if (fl_row[j] < staged) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(fl_s[j]));
                st_stream16(gflush + fl_g[j], v);
            }
        __syncwarp();
        gflush += (size_t)(unsigned)staged * out_row_bytes;
        staged = 0;
        sp = stage_s + 3 * lane;
    };

    // vertical pass + store of the NC output pixels from the window (h0 = oldest row)
    auto emit = [&](const f32x2 (&h0)[3][NP], const f32x2 (&h1)[3][NP], const f32x2 (&h2)[3][NP], const f32x2 (&h3)[3][NP]) {
        f32x2 w0, w1, w2, w3;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(entry));
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+16];" : "=l"(w2), "=l"(w3) : "r"(entry));
        int v[3][NC];
#pragma unroll
        for (int k = 0; k < 3; ++k) {   // OpenCV's SSE2 body: mulps, addps (one rounding each), cvtps2dq (half-even), packs, packus
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                f32x2 f = fma2(h0[k][p], w0, negzero2);
                f = fma2(f, one2, fma2(h1[k][p], w1, negzero2));
                f = fma2(f, one2, fma2(h2[k][p], w2, negzero2));
                f = fma2(f, one2, fma2(h3[k][p], w3, negzero2));
                // |f| < 2^22: adding 1.5*2^23 rounds half-to-even at integer granularity; subtracting its bit pattern and clamping
                // to [0,255] is one DPX op per value (the intermediate s16 saturation of packs cannot change the result)
                f = fma2(f, one2, magic2);
                int lo, hi;
                unpack2i(f, lo, hi);
                v[k][2 * p] = __viaddmin_s32_relu(lo, -0x4B400000, 255);
                v[k][2 * p + 1] = __viaddmin_s32_relu(hi, -0x4B400000, 255);
            }
        }
        if (staged_store) {   // the NC pixels are 32 columns = 96 bytes apart in the staged row
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int k = 0; k < 3; ++k) asm volatile("st.shared.u8 [%0], %1;" ::"r"(sp + 96 * c + k), "r"(v[k][c]) : "memory");
            sp += kWarpRow;
            if (++staged == kWnStageRows) flush();
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (dx + 32 * c < g.wo) { gdirect[96 * c] = (uint8_t)v[0][c]; gdirect[96 * c + 1] = (uint8_t)v[1][c]; gdirect[96 * c + 2] = (uint8_t)v[2][c]; }
            gdirect += out_row_bytes;
        }
        entry += (int)sizeof(Walk2Row);
        asm volatile("ld.shared.s32 %0, [%1+32];" : "=r"(next_last) : "r"(entry));   // sentinel INT_MAX after the last row
    };

    // ---- the walk: row t -> window slot t & 3 (compile-time inside the 4x unrolled body); an output row is emitted as soon as its
    //      last tap row has been filtered
    f32x2 H[4][3][NP];
    const uint32_t entry_end = rows_s + nrows * (int)sizeof(Walk2Row);
    // put rows t .. t + kWnAhead - 1 in flight: tpre runs from t - 0; the ring slot of row r is r & 7
    for (int a = 0; a < kWnAhead; ++a) prefetch();
    // kPre: software pipeline -- the tap words of row t are fetched one step early, so their shared-memory latency hides behind the
    // previous row's vertical pass (16 more live registers); !kPre: fetched right before use, other warps cover the latency
    if (kPre) { prefetch(); fetch(t); }
    while ((t & 3) != 0) {
        f32x2 hv[3][NP];                               // leading rows up to the first multiple of 4: no complete window yet
        if (!kPre) { prefetch(); fetch(t); }
        hfilter(hv);
        const int slot = t & 3;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (slot == 1) H[1][k][p] = hv[k][p];
                if (slot == 2) H[2][k][p] = hv[k][p];
                if (slot == 3) H[3][k][p] = hv[k][p];
            }
        ++t;
        if (kPre) { prefetch(); fetch(t); }
    }
    while (entry != entry_end) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {                  // t & 3 == u
            if (!kPre) { prefetch(); fetch(t); }
            hfilter(H[u]);                             // consumes the tap words of row t
            const bool due = next_last == t;
            ++t;
            if (kPre) { prefetch(); fetch(t); }        // issue row t + kWnAhead, wait for row t, read its tap words
            if (kDown) {                               // scale_y >= 1: consecutive output rows end on different source rows
                if (due) emit(H[(u + 1) & 3], H[(u + 2) & 3], H[(u + 3) & 3], H[u]);
            } else {
                while (next_last == t - 1) emit(H[(u + 1) & 3], H[(u + 2) & 3], H[(u + 3) & 3], H[u]);
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (staged_store && staged) flush();
}

}  // namespace vacv
