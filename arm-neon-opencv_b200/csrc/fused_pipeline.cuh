// Persistent, TMA-fed version of the config-2 fused pipeline (see fused.cu for the arithmetic contract).
//
//   grid  = 148 SMs x kCtasPerSm persistent CTAs; CTA c processes tiles c, c+grid, c+2*grid, ...
//   tile  = TH full-width output rows of one frame.  Its source is one CONTIGUOUS byte range of the Y plane and
//           one of the chroma plane (dense frames, full rows), so each is fetched by a single 1-D bulk copy
//           (cp.async.bulk global->shared, completion on an mbarrier; SASS: UBLKCP) issued by one thread.
//   pipe  = two shared-memory stages: the copy for tile i+1 is in flight while all warps compute tile i;
//           a stage is handed back with one __syncthreads per tile.
//   once per CTA (amortised over ~70 tiles): the exact 3x256 normalisation table, all h_out row coefficients
//           (shared memory) and each thread's column coefficients (registers).
//   inner loop: a thread owns NCOL output columns and walks down the tile rows.  Per source row it forms the
//           horizontally interpolated BGR triple H = p_left*cx0 + p_right*cx1 from u8-clamped taps; chroma terms are
//           cached per chroma row (two luma rows share one), H of the lower row is carried to the next output row
//           when it starts there.  out = table[c][(H0*cy0 + H1*cy1) >> 22], stored 128 B / warp / plane, streaming.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is resolved at run time, tma_host.cuh)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "vacv_common.cuh"

namespace vacv {

constexpr int kPipeThreads = 384;    // max threads per CTA (2 CTAs / SM -> <= 85 registers per thread)
constexpr int kPipeMaxCols = 4;      // output columns per thread -> w_out <= 1536

struct PipeGeom {
    int w, h, wo, ho;
    int TH;                 // output rows per tile
    int tiles_per_frame;    // ceil(ho / TH)
    int total_tiles;        // tiles_per_frame * batch
    int ystage, cstage;     // bytes reserved per stage for the Y / chroma band (multiples of 128)
    int table_bytes;        // bytes of the two row tables at the start of dynamic shared memory (multiple of 128)
    // source layout (dense NV12/NV21: y_pitch = c_pitch = w, c_off = w*h, frame_stride = w*h*3/2)
    int y_pitch, c_pitch;   // bytes per luma row / per chroma row (planar formats: per U or V row)
    // staged rows: sy_pitch / sc_pitch bytes apart in shared memory.  Equal to the source pitches when a band is one contiguous copy;
    // with VACV_PIPE_ROWS=1 on padded surfaces (pitch >= row + 16) the row rounded up to 16 bytes -- the band is then copied row by
    // row and the padding stays in DRAM
    int sy_pitch, sc_pitch;
    int vstage_off;         // planar formats: byte offset of the V band inside the chroma stage
    size_t frame_stride;    // bytes between frames
    size_t c_off, c2_off;   // byte offset of the chroma plane (semi-planar) / of the U and V planes (planar) inside a frame
    // destination canvas (letterbox): the wo x ho result is written at (x0, y0) of a canvas_w x canvas_h plane; without
    // letterbox canvas = result and x0 = y0 = 0
    int canvas_w, canvas_h, x0, y0;
    int bf16;               // 16-bit outputs: table holds bfloat16 instead of half
    int tile_maps;          // padded surfaces staged through tensor maps (PipeMaps below)
};

// Padded decoder surfaces (pitch >= row + 16), default since round 2: a tile's luma / chroma band is ONE tensor-map box of
// [band rows] x [w bytes] (cp.async.bulk.tensor, SASS UTMALDG) -- only the rows' own bytes leave DRAM and land densely in the stage,
// one copy per plane like the dense path.  A box has a fixed row count, so there is one map per band height that occurs (the
// heights of a linear scale take two or three values); frames are the third tensor dimension.
constexpr int kPipeMapHeights = 4;
struct PipeMaps {
    CUtensorMap y[kPipeMapHeights], c[kPipeMapHeights], c2[kPipeMapHeights];   // luma, chroma (or U), V -- one per band height
    int yh[kPipeMapHeights], ch[kPipeMapHeights];                               // the band heights (unused slots 0)
};

enum { kFmtVU = 0, kFmtUV = 1, kFmtPlanar = 2 };   // interleaved chroma V-first (NV21), U-first (NV12), separate U and V planes (I420 / YV12)

// ---- mbarrier / bulk-copy PTX (sm_90+; SASS on sm_100a: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void pipe_tma_load_3d(uint32_t smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// ---- explicit shared-window accesses (32-bit addresses: no generic->shared conversion in the inner loop)
__device__ __forceinline__ int lds_u8(uint32_t a) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u16(uint32_t a) { unsigned v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }

struct ColState {
    int yo;        // byte offset of the left tap inside a Y row
    int ca, cb;    // byte offsets of the chroma pairs of the left / right tap inside a chroma row
    int cx0, cx1;
};

// Chroma terms straight from the interleaved 16-bit sample with IDP.4A: 179*(v-128) = 179*v - 22912 etc. (cvt_color.cpp:76-78),
// the constant rides in the accumulator, so no byte unpacking and no subtraction of 128: 2 instructions per term.
template <bool kVFirst>
__device__ __forceinline__ ChromaTerms chroma_terms_pair(unsigned pair) {   // pair = first chroma byte | second << 8
    constexpr unsigned kRa = kVFirst ? 179u : 179u << 8, kBa = kVFirst ? 227u << 8 : 227u, kGa = kVFirst ? (91u | 44u << 8) : (44u | 91u << 8);
    ChromaTerms t;
    t.ra = (int)__dp4a(pair, kRa, (unsigned)-22912) >> 7;
    t.ga = (int)__dp4a(pair, kGa, (unsigned)-17280) >> 7;
    t.ba = (int)__dp4a(pair, kBa, (unsigned)-29056) >> 7;
    return t;
}

// kDp4a: measured on B200 -- the IDP.4A form wins where the loop is issue / latency bound (right taps live: 1080p -> 608x608
// 0.419 -> 0.399 ms; 16-bit outputs) and costs ~1 % on the HBM-bound integer-ratio fp32 case, which keeps the IMAD form.
template <int FMT, bool kDp4a>
__device__ __forceinline__ ChromaTerms terms_at(uint32_t addr, int vstage_off) {
    if (kDp4a) {
        if (FMT == kFmtPlanar) return chroma_terms_pair<true>((unsigned)lds_u8(addr + vstage_off) | ((unsigned)lds_u8(addr) << 8));   // (v, u)
        return chroma_terms_pair<FMT == kFmtVU>(lds_u16(addr));
    }
    if (FMT == kFmtPlanar) return chroma_terms(lds_u8(addr + vstage_off), lds_u8(addr));   // (v, u)
    const unsigned p = lds_u16(addr);
    return FMT == kFmtVU ? chroma_terms(p & 0xff, p >> 8) : chroma_terms(p >> 8, p & 0xff);
}

// normalisation table entry -> global memory.  OutT = float: fp32 planes; __half: fp16 planes, one 16-bit store per element
// (odd w_out); __half2: fp16 planes, a thread owns column PAIRS and stores 32 bits (128 B per warp and store).
template <typename OutT> struct OutOps;
template <> struct OutOps<float> {
    typedef float Lut;
    enum { kElem = 4, kCols = 1 };
    static __device__ __forceinline__ float make(float v, int) { return v; }
    static __device__ __forceinline__ void copy(char* o, uint32_t lut, int k, unsigned v, unsigned) { st_stream4f(o, lds_f32(lut + k * 1024 + v * 4)); }
};
template <> struct OutOps<__half> {
    typedef unsigned short Lut;
    enum { kElem = 2, kCols = 1 };
    static __device__ __forceinline__ unsigned short make(float v, int bf16) {   // exact fp32 value rounded to nearest-even fp16 / bf16
        return bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
    }
    static __device__ __forceinline__ void copy(char* o, uint32_t lut, int k, unsigned v, unsigned) {
        const unsigned short h = (unsigned short)lds_u16(lut + k * 512 + v * 2);
        asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(o), "h"(h) : "memory");
    }
};
template <> struct OutOps<__half2> {
    typedef unsigned short Lut;
    enum { kElem = 2, kCols = 2 };
    static __device__ __forceinline__ unsigned short make(float v, int bf16) { return OutOps<__half>::make(v, bf16); }
    static __device__ __forceinline__ void copy(char* o, uint32_t lut, int k, unsigned v0, unsigned v1) {
        const unsigned pair = lds_u16(lut + k * 512 + v0 * 2) | (lds_u16(lut + k * 512 + v1 * 2) << 16);
        asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(o), "r"(pair) : "memory");
    }
};

template <bool kRightTap>
__device__ __forceinline__ void hrow_cached(uint32_t yaddr, const ColState& c, const ChromaTerms& ta,
                                            const ChromaTerms& tb, int (&H)[3]) {
    const int Y0 = lds_u8(yaddr);
    if (kRightTap) {
        const int Y1 = lds_u8(yaddr + 1);
        H[0] = add_clamp255(Y0, ta.ba) * c.cx0 + add_clamp255(Y1, tb.ba) * c.cx1;
        H[1] = add_clamp255(Y0, -ta.ga) * c.cx0 + add_clamp255(Y1, -tb.ga) * c.cx1;
        H[2] = add_clamp255(Y0, ta.ra) * c.cx0 + add_clamp255(Y1, tb.ra) * c.cx1;
    } else {   // every cx1 of this launch is 0 (odd integer x ratio): the right tap contributes p*0
        H[0] = add_clamp255(Y0, ta.ba) * c.cx0;
        H[1] = add_clamp255(Y0, -ta.ga) * c.cx0;
        H[2] = add_clamp255(Y0, ta.ra) * c.cx0;
    }
}

// ybuf/cbuf/tab_sy/tab_cy/lut are shared-window addresses.  out[j] points at (frame, plane 0, row dy0, column of j).
template <int FMT, typename OutT, bool kRightTap, int NCOL>
__device__ __forceinline__ void compute_tile(uint32_t ybuf, uint32_t cbuf, uint32_t tab_sy, uint32_t tab_cy, uint32_t lut,
                                             const ColState (&col)[NCOL], int dy0, int th, int y_pitch, int c_pitch, int vstage_off,
                                             char* (&out)[NCOL], size_t row_bytes, size_t plane_bytes) {
    const int y_first = lds_s32(tab_sy + 4 * dy0), c_first = y_first >> 1;
    int H0[NCOL][3], H1[NCOL][3];
    ChromaTerms ta[NCOL], tb[NCOL];
    int have = -2, have_c = -1;
    auto row = [&](int r, int (&H)[NCOL][3]) {
        const int cr = r >> 1;
        if (cr != have_c) {   // all threads walk the same rows: no divergence
            const uint32_t crow = cbuf + (cr - c_first) * c_pitch;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                ta[j] = terms_at<FMT, kRightTap>(crow + col[j].ca, vstage_off);
                if (kRightTap) tb[j] = terms_at<FMT, true>(crow + col[j].cb, vstage_off);
            }
            have_c = cr;
        }
        const uint32_t yrow = ybuf + (r - y_first) * y_pitch;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) hrow_cached<kRightTap>(yrow + col[j].yo, col[j], ta[j], tb[j], H[j]);
    };
    for (int ty = 0; ty < th; ++ty) {
        const int sy = lds_s32(tab_sy + 4 * (dy0 + ty));
        const int cy = lds_s32(tab_cy + 4 * (dy0 + ty));
        const int cy0 = cy & 0xffff, cy1 = cy >> 16;   // both in [0, 2048]
        if (sy == have) {
#pragma unroll
            for (int j = 0; j < NCOL; ++j) { H0[j][0] = H1[j][0]; H0[j][1] = H1[j][1]; H0[j][2] = H1[j][2]; }
        } else if (sy + 1 != have) {
            row(sy, H0);
        }
        if (sy + 1 != have) {
            row(sy + 1, H1);
            have = sy + 1;
        }
        constexpr int kCols = OutOps<OutT>::kCols;
#pragma unroll
        for (int j = 0; j < NCOL; j += kCols) {
            char* o = out[j];
            constexpr int j1 = kCols == 2 ? 1 : 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const unsigned v = (unsigned)(H0[j][k] * cy0 + H1[j][k] * cy1) >> 22;   // the u8 the unfused chain stores
                const unsigned v1 = kCols == 2 ? (unsigned)(H0[j + j1][k] * cy0 + H1[j + j1][k] * cy1) >> 22 : 0u;
                OutOps<OutT>::copy(o, lut, k, v, v1);
                o += plane_bytes;
            }
            out[j] += row_bytes;
        }
    }
}

// Second version of compute_tile: two register sets by source-row parity instead of upper / lower sets.
template <int FMT, typename OutT, bool kRightTap, int NCOL>
__device__ __forceinline__ void compute_tile_parity(uint32_t ybuf, uint32_t cbuf, uint32_t tab_sy, uint32_t tab_cy, uint32_t lut,
                                             const ColState (&col)[NCOL], int dy0, int th, int y_pitch, int c_pitch, int vstage_off,
                                             char* (&out)[NCOL], size_t row_bytes, size_t plane_bytes) {
    const int y_first = lds_s32(tab_sy + 4 * dy0), c_first = y_first >> 1;
    // Horizontal sums live in two register sets: Ha holds an EVEN source row, Hb an ODD one (have_a / have_b: which).  An output row
    // blends row sy with row sy + 1, one of each parity; when the next output row starts on the previous lower row, that row's sums
    // already sit where its parity puts them -- nothing is copied.  Used where it measured faster than the first version (compute_tile,
    // upper / lower sets with a copy when rows are shared) on one box, 1080p -> 640x640 x256, each tap rule in its own kernel:
    // every launch with right taps (608x608 0.395 -> 0.371 ms, 720p source 0.333 -> 0.317, 416x416 0.244 -> 0.236) and planar
    // chroma without (I420 0.379 -> 0.365); NV12 without right taps keeps the first version (dense 0.351 vs 0.376, pitch 2048
    // 0.380 vs 0.387, letterbox 0.381 vs 0.396).
    int Ha[NCOL][3] = {}, Hb[NCOL][3] = {};
    ChromaTerms ta[NCOL], tb[NCOL];
    int have_a = -2, have_b = -2, have_c = -1;
    auto row = [&](int r, int (&H)[NCOL][3]) {
        const int cr = r >> 1;
        if (cr != have_c) {   // all threads walk the same rows: no divergence
            const uint32_t crow = cbuf + (cr - c_first) * c_pitch;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                ta[j] = terms_at<FMT, kRightTap>(crow + col[j].ca, vstage_off);
                if (kRightTap) tb[j] = terms_at<FMT, true>(crow + col[j].cb, vstage_off);
            }
            have_c = cr;
        }
        const uint32_t yrow = ybuf + (r - y_first) * y_pitch;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) hrow_cached<kRightTap>(yrow + col[j].yo, col[j], ta[j], tb[j], H[j]);
    };
    auto emit = [&](const int (&Hu)[NCOL][3], const int (&Hl)[NCOL][3], int cy0, int cy1) {   // Hu / Hl: sums of the upper / lower tap row
        constexpr int kCols = OutOps<OutT>::kCols;
#pragma unroll
        for (int j = 0; j < NCOL; j += kCols) {
            char* o = out[j];
            constexpr int j1 = kCols == 2 ? 1 : 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const unsigned v = (unsigned)(Hu[j][k] * cy0 + Hl[j][k] * cy1) >> 22;   // the u8 the unfused chain stores
                const unsigned v1 = kCols == 2 ? (unsigned)(Hu[j + j1][k] * cy0 + Hl[j + j1][k] * cy1) >> 22 : 0u;
                OutOps<OutT>::copy(o, lut, k, v, v1);
                o += plane_bytes;
            }
            out[j] += row_bytes;
        }
    };
    for (int ty = 0; ty < th; ++ty) {
        const int sy = lds_s32(tab_sy + 4 * (dy0 + ty));
        const int cy = lds_s32(tab_cy + 4 * (dy0 + ty));
        const int cy0 = cy & 0xffff, cy1 = cy >> 16;   // both in [0, 2048]
        // a lower tap row of weight 0 is not filtered: whatever its register set holds is multiplied by 0
        if ((sy & 1) == 0) {
            if (have_a != sy) { row(sy, Ha); have_a = sy; }
            if (cy1 != 0 && have_b != sy + 1) { row(sy + 1, Hb); have_b = sy + 1; }
            emit(Ha, Hb, cy0, cy1);
        } else {
            if (have_b != sy) { row(sy, Hb); have_b = sy; }
            if (cy1 != 0 && have_a != sy + 1) { row(sy + 1, Ha); have_a = sy + 1; }
            emit(Hb, Ha, cy0, cy1);
        }
    }
}

// Packed variant for an even number of columns per thread: columns are handled in PAIRS on 16-bit lanes.
//   * add + clamp of a luma sample and a chroma term for two columns = one DPX VIADDMNMX.S16x2.RELU (was two VIADDMNMX),
//   * the horizontal blend pa*cx0 + pb*cx1 = one IDP.2A on the byte pair (pa, pb) and the packed 16-bit weights (was two
//     IMAD); when no right tap has weight, cx0 is exactly 2048 and the blend disappears: the clamped bytes themselves
//     are carried and the vertical blend (p0*cy0 + p1*cy1) >> 11 -- the same integer as ((p0*2048)*cy0 + (p1*2048)*cy1)
//     >> 22 -- is again one IDP.2A per value, with the row table's packed (cy0, cy1) entry as is.
template <int FMT, typename OutT, bool kRightTap, int NCOL>
__device__ __forceinline__ void compute_tile_packed(uint32_t ybuf, uint32_t cbuf, uint32_t tab_sy, uint32_t tab_cy, uint32_t lut,
                                                    const ColState (&col)[NCOL], int dy0, int th, int y_pitch, int c_pitch, int vstage_off,
                                                    char* (&out)[NCOL], size_t row_bytes, size_t plane_bytes) {
    constexpr int P = NCOL >= 2 ? NCOL / 2 : 1;   // (instantiated but never called for odd NCOL)
    constexpr int kCols = OutOps<OutT>::kCols;
    const int y_first = lds_s32(tab_sy + 4 * dy0), c_first = y_first >> 1;
    // row state: kRightTap: horizontally blended sums per column; else: clamped bytes of a column pair [p_j 0 p_j+1 0] per channel
    // two register sets by source-row parity (see compute_tile): set a holds an even row, set b an odd one
    int Ha[kRightTap ? NCOL : 1][3] = {}, Hb[kRightTap ? NCOL : 1][3] = {};
    uint32_t Ra[kRightTap ? 1 : P][3] = {}, Rb[kRightTap ? 1 : P][3] = {};
    uint32_t ta2[P][3], tb2[kRightTap ? P : 1][3];   // chroma terms (ba, -ga, ra) of the pair's two columns, one per 16-bit lane
    uint32_t cxp[kRightTap ? NCOL : 1];
    if (kRightTap) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j) cxp[j] = (uint32_t)col[j].cx0 | ((uint32_t)col[j].cx1 << 16);
    }
    int have_a = -2, have_b = -2, have_c = -1;
    auto pack_terms = [](const ChromaTerms& a, const ChromaTerms& b, uint32_t (&t)[3]) {
        t[0] = __byte_perm((uint32_t)a.ba, (uint32_t)b.ba, 0x5410);
        t[1] = __byte_perm((uint32_t)(-a.ga), (uint32_t)(-b.ga), 0x5410);
        t[2] = __byte_perm((uint32_t)a.ra, (uint32_t)b.ra, 0x5410);
    };
    auto row = [&](int r, int (&H)[kRightTap ? NCOL : 1][3], uint32_t (&R)[kRightTap ? 1 : P][3]) {
        const int cr = r >> 1;
        if (cr != have_c) {   // all threads walk the same rows: no divergence
            const uint32_t crow = cbuf + (cr - c_first) * c_pitch;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                pack_terms(terms_at<FMT, true>(crow + col[2 * p].ca, vstage_off), terms_at<FMT, true>(crow + col[2 * p + 1].ca, vstage_off), ta2[p]);
                if (kRightTap) pack_terms(terms_at<FMT, true>(crow + col[2 * p].cb, vstage_off), terms_at<FMT, true>(crow + col[2 * p + 1].cb, vstage_off), tb2[p]);
            }
            have_c = cr;
        }
        const uint32_t yrow = ybuf + (r - y_first) * y_pitch;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const uint32_t ya = (uint32_t)lds_u8(yrow + col[2 * p].yo) | ((uint32_t)lds_u8(yrow + col[2 * p + 1].yo) << 16);
            if (kRightTap) {
                const uint32_t yb = (uint32_t)lds_u8(yrow + col[2 * p].yo + 1) | ((uint32_t)lds_u8(yrow + col[2 * p + 1].yo + 1) << 16);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t a = __viaddmin_s16x2_relu(ya, ta2[p][k], 0x00ff00ffu), b = __viaddmin_s16x2_relu(yb, tb2[p][k], 0x00ff00ffu);
                    const uint32_t m = __byte_perm(a, b, 0x6240);   // [pa_j pb_j pa_j+1 pb_j+1]
                    H[2 * p][k] = (int)__dp2a_lo(cxp[2 * p], m, 0u);
                    H[2 * p + 1][k] = (int)__dp2a_hi(cxp[2 * p + 1], m, 0u);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) R[p][k] = __viaddmin_s16x2_relu(ya, ta2[p][k], 0x00ff00ffu);
            }
        }
    };
    // u = upper tap row's set, l = lower tap row's; cy = cy0 | cy1 << 16, both in [0, 2048]
    auto emit = [&](const int (&Hu)[kRightTap ? NCOL : 1][3], const int (&Hl)[kRightTap ? NCOL : 1][3], const uint32_t (&Ru)[kRightTap ? 1 : P][3],
                    const uint32_t (&Rl)[kRightTap ? 1 : P][3], uint32_t cy) {
        const int cy0 = cy & 0xffff, cy1 = cy >> 16;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            char* o0 = out[2 * p];
            char* o1 = out[2 * p + 1];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                unsigned v0, v1;   // the u8 the unfused chain stores
                if (kRightTap) {
                    v0 = (unsigned)(Hu[2 * p][k] * cy0 + Hl[2 * p][k] * cy1) >> 22;
                    v1 = (unsigned)(Hu[2 * p + 1][k] * cy0 + Hl[2 * p + 1][k] * cy1) >> 22;
                } else {
                    const uint32_t m = __byte_perm(Ru[p][k], Rl[p][k], 0x6240);   // [p0_j p1_j p0_j+1 p1_j+1]
                    v0 = __dp2a_lo(cy, m, 0u) >> 11;
                    v1 = __dp2a_hi(cy, m, 0u) >> 11;
                }
                if (kCols == 2) {
                    OutOps<OutT>::copy(o0, lut, k, v0, v1);
                } else {
                    OutOps<OutT>::copy(o0, lut, k, v0, 0u);
                    OutOps<OutT>::copy(o1, lut, k, v1, 0u);
                    o1 += plane_bytes;
                }
                o0 += plane_bytes;
            }
            out[2 * p] += row_bytes;
            if (kCols != 2) out[2 * p + 1] += row_bytes;
        }
    };
    for (int ty = 0; ty < th; ++ty) {
        const int sy = lds_s32(tab_sy + 4 * (dy0 + ty));
        const uint32_t cy = (uint32_t)lds_s32(tab_cy + 4 * (dy0 + ty));
        const bool lower = (cy >> 16) != 0;           // a lower tap row of weight 0 is not filtered: its set is multiplied by 0
        if ((sy & 1) == 0) {
            if (have_a != sy) { row(sy, Ha, Ra); have_a = sy; }
            if (lower && have_b != sy + 1) { row(sy + 1, Hb, Rb); have_b = sy + 1; }
            emit(Ha, Hb, Ra, Rb, cy);
        } else {
            if (have_b != sy) { row(sy, Hb, Rb); have_b = sy; }
            if (lower && have_a != sy + 1) { row(sy + 1, Ha, Ra); have_a = sy + 1; }
            emit(Hb, Ha, Rb, Ra, cy);
        }
    }
}

// Padded surfaces (pitch >= row + 16: decoder pools), opt-in variant VACV_PIPE_ROWS=1 (measured slower than whole bands, see
// build_pipe_plan in fused.cu): the band of a tile is copied ROW BY ROW, g.sy_pitch / g.sc_pitch bytes each (the
// row rounded up to 16 bytes), so the padding stays in DRAM; lane i of the calling warp issues rows i, i + 32, ...  Kept out of line:
// inlined, it changed the register allocation of the tile loop (fp16 outputs lost 8 %).
// (Its arguments are scalars on purpose: a `const PipeGeom&` made every kernel that can reach this call keep a copy of the
// geometry in local memory -- 112 bytes of stack in all non-dense instantiations.)
template <int FMT>
__device__ __noinline__ void issue_band_rows(int sy_pitch, int sc_pitch, int y_pitch, int c_pitch, int ystage, int vstage_off, size_t c_off, size_t c2_off,
                                             const uint8_t* f, uint8_t* st, uint64_t* bar, int y_first, int y_last, int lane) {
    const int c_first = y_first >> 1, c_last = y_last >> 1;
    const int yrows = y_last - y_first + 1, crows = c_last - c_first + 1;
    constexpr int kPlanes = FMT == kFmtPlanar ? 2 : 1;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)yrows * sy_pitch + (uint32_t)(kPlanes * crows) * sc_pitch);
    __syncwarp();
    for (int r = lane; r < yrows + kPlanes * crows; r += 32) {
        if (r < yrows) {
            bulk_g2s(st + (size_t)r * sy_pitch, f + (size_t)(y_first + r) * y_pitch, (uint32_t)sy_pitch, bar);
        } else {
            const int q = r - yrows, second = q >= crows ? 1 : 0, cr = q - second * crows;
            bulk_g2s(st + ystage + (second ? vstage_off : 0) + (size_t)cr * sc_pitch,
                     f + (second ? c2_off : c_off) + (size_t)(c_first + cr) * c_pitch, (uint32_t)sc_pitch, bar);
        }
    }
}

// Padded surfaces, default: the band of a tile is one tensor-map box per plane (PipeMaps), only the rows' own bytes leave DRAM.
// Out of line for the same reason as issue_band_rows.
template <int FMT>
__device__ __noinline__ void issue_band_maps(const PipeMaps* maps, uint32_t st, int ystage, int vstage_off, int sy_pitch, int sc_pitch,
                                             uint64_t* bar, int y_first, int y_last, int frame) {
    const int c_first = y_first >> 1, c_last = y_last >> 1;
    const int yrows = y_last - y_first + 1, crows = c_last - c_first + 1;
    int ky = 0, kc = 0;
#pragma unroll
    for (int k = 1; k < kPipeMapHeights; ++k) {
        if (maps->yh[k] == yrows) ky = k;
        if (maps->ch[k] == crows) kc = k;
    }
    mbar_expect_tx(bar, (uint32_t)yrows * sy_pitch + (uint32_t)((FMT == kFmtPlanar ? 2 : 1) * crows) * sc_pitch);
    pipe_tma_load_3d(st, &maps->y[ky], 0, y_first, frame, bar);
    pipe_tma_load_3d(st + ystage, &maps->c[kc], 0, c_first, frame, bar);
    if (FMT == kFmtPlanar) pipe_tma_load_3d(st + ystage + vstage_off, &maps->c2[kc], 0, c_first, frame, bar);
}

// kDense: the reference's own layout (tensor.cpp:524: no pitch; chroma right after luma) -- one pitch register, constants folded.
// RIGHT: -1 = whether any right tap has weight is found out at run time (both tile loops in the kernel), 0 / 1 = known to the launcher:
// one tap rule per kernel, so that the loop a launch never takes does not shape the register allocation and schedule of the one it
// does (config 2, same box: 0.364 -> 0.351 ms).
// kMaps (with kDense): the SOURCE is a padded surface staged through tensor maps -- the stage then looks exactly like the dense case
// (rows w bytes apart), so the dense tile loops and output addressing apply; only the copy issue differs.
template <int FMT, typename OutT, int NCOL, bool kDense, int RIGHT = -1, bool kMaps = false>
__global__ void __launch_bounds__(NCOL == 1 ? 640 : kPipeThreads, NCOL <= 2 ? 2 : 1)
nv_resize_normalize_chw_pipe_kernel(const uint8_t* __restrict__ src, OutT* __restrict__ dst, PipeGeom g,
                                    const float* __restrict__ mean, const float* __restrict__ stddev,
                                    const __grid_constant__ PipeMaps maps) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];     // [s_sy: ho][s_cy: ho][pad] 2 x (ystage + cstage)
    typedef OutOps<OutT> Ops;
    static_assert(NCOL % Ops::kCols == 0, "column pairs need an even column count");
    __shared__ typename Ops::Lut lut[768];
    int* s_sy = reinterpret_cast<int*>(dyn_smem);
    int* s_cy = s_sy + g.ho;
    uint8_t* stages = dyn_smem + g.table_bytes;
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ int s_any_right;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t plane_bytes = kDense ? (size_t)g.wo * g.ho * Ops::kElem : (size_t)g.canvas_w * g.canvas_h * Ops::kElem;
    const size_t row_bytes = kDense ? (size_t)g.wo * Ops::kElem : (size_t)g.canvas_w * Ops::kElem;
    const int y_pitch = kDense ? g.w : g.y_pitch, c_pitch = kDense ? g.w : g.c_pitch;
    const int sy_pitch = kDense ? g.w : g.sy_pitch, sc_pitch = kDense ? g.w : g.sc_pitch;      // row pitches of the staged bands
    const bool by_map = kDense ? kMaps : g.tile_maps != 0;
    const bool by_row = !kDense && !by_map && (sy_pitch != y_pitch || sc_pitch != c_pitch);
    const size_t frame_stride = kDense ? (size_t)g.w * g.h * 3 / 2 : g.frame_stride, c_off = kDense ? (size_t)g.w * g.h : g.c_off;
    const uint32_t stages_s = smem_u32(stages), sy_s = smem_u32(s_sy), cy_s = smem_u32(s_cy), lut_s = smem_u32(lut);

    // ---- once per CTA: barriers, normalisation table, row and column coefficients
    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        s_any_right = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = tid; t < 768; t += nthr)
        lut[t] = Ops::make(normalize_one((float)(t & 255), __ldg(mean + (t >> 8)), (double)__ldg(stddev + (t >> 8)) + 1e-6), g.bf16);
    const double scale_x = (double)((float)g.w / (float)g.wo), scale_y = (double)((float)g.h / (float)g.ho);
    for (int dy = tid; dy < g.ho; dy += nthr) {
        int s; float f;
        linear_coord(dy, scale_y, g.h, s, f);
        s_sy[dy] = s;
        s_cy[dy] = sat_short((1.f - f) * 2048.f) | (sat_short(2048.f * f) << 16);
    }
    __syncthreads();
    // Column ownership: thread t owns columns t, t+nthr, ...; a column index past the row end is clamped to the
    // last column (the thread then recomputes and rewrites that pixel with the identical value: no branch needed).
    ColState col[NCOL];
    int colx[NCOL];
    int any_right = 0;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        // column pairs (fp16x2 stores, w_out even): pair p of thread t = columns 2(t + p nthr), +1
        const int dx = Ops::kCols == 2 ? min(2 * (tid + (j >> 1) * nthr), g.wo - 2) + (j & 1) : min(tid + j * nthr, g.wo - 1);
        colx[j] = dx;
        int sx; float fx;
        linear_coord(dx, scale_x, g.w, sx, fx);
        col[j].cx0 = sat_short((1.f - fx) * 2048.f);
        col[j].cx1 = sat_short(2048.f * fx);
        col[j].yo = sx;
        col[j].ca = FMT == kFmtPlanar ? sx >> 1 : sx & ~1;
        col[j].cb = FMT == kFmtPlanar ? (sx + 1) >> 1 : (sx + 1) & ~1;
        any_right |= col[j].cx1;
    }
    if (any_right) s_any_right = 1;   // benign race: all writers store 1

    auto issue = [&](int tile, int b) {   // one thread: every band is contiguous in memory -> one copy each
        const int frame = tile / g.tiles_per_frame, dy0 = (tile - frame * g.tiles_per_frame) * g.TH;
        const int th = min(g.TH, g.ho - dy0);
        const int y_first = s_sy[dy0], y_last = s_sy[dy0 + th - 1] + 1;
        const int c_first = y_first >> 1, c_last = y_last >> 1;
        const uint32_t ybytes = (uint32_t)(y_last - y_first + 1) * y_pitch, cbytes = (uint32_t)(c_last - c_first + 1) * c_pitch;
        const uint8_t* f = src + (size_t)frame * frame_stride;
        uint8_t* st = stages + (size_t)b * (g.ystage + g.cstage);
        mbar_expect_tx(&full_bar[b], ybytes + (FMT == kFmtPlanar ? 2 * cbytes : cbytes));
        bulk_g2s(st, f + (size_t)y_first * y_pitch, ybytes, &full_bar[b]);
        bulk_g2s(st + g.ystage, f + c_off + (size_t)c_first * c_pitch, cbytes, &full_bar[b]);
        if (FMT == kFmtPlanar) bulk_g2s(st + g.ystage + g.vstage_off, f + g.c2_off + (size_t)c_first * c_pitch, cbytes, &full_bar[b]);
    };
    auto issue_maps = [&](int tile, int b) {   // padded surface, one thread
        const int frame = tile / g.tiles_per_frame, dy0 = (tile - frame * g.tiles_per_frame) * g.TH;
        const int th = min(g.TH, g.ho - dy0);
        issue_band_maps<FMT>(&maps, stages_s + (uint32_t)b * (g.ystage + g.cstage), g.ystage, g.vstage_off, sy_pitch, sc_pitch, &full_bar[b],
                             s_sy[dy0], s_sy[dy0 + th - 1] + 1, frame);
    };
    auto issue_rows = [&](int tile, int b) {   // padded surface; called by every lane of warp 0
        const int frame = tile / g.tiles_per_frame, dy0 = (tile - frame * g.tiles_per_frame) * g.TH;
        const int th = min(g.TH, g.ho - dy0);
        issue_band_rows<FMT>(sy_pitch, sc_pitch, y_pitch, c_pitch, g.ystage, g.vstage_off, c_off, g.c2_off, src + (size_t)frame * frame_stride,
                             stages + (size_t)b * (g.ystage + g.cstage), &full_bar[b], s_sy[dy0], s_sy[dy0 + th - 1] + 1, tid);
    };

    int tile = blockIdx.x;
    if (tile < g.total_tiles) {
        if (by_row) { if (tid < 32) issue_rows(tile, 0); }
        else if (by_map) { if (tid == 0) issue_maps(tile, 0); }
        else if (tid == 0) issue(tile, 0);
    }
    __syncthreads();
    const bool right = RIGHT < 0 ? s_any_right != 0 : RIGHT == 1;

    for (int it = 0; tile < g.total_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const int next = tile + gridDim.x;
        if (next < g.total_tiles) {   // stage b^1 was released by the sync below
            if (by_row) { if (tid < 32) issue_rows(next, b ^ 1); }
            else if (by_map) { if (tid == 0) issue_maps(next, b ^ 1); }
            else if (tid == 0) issue(next, b ^ 1);
        }
        mbar_wait(&full_bar[b], (it >> 1) & 1);
        const int frame = tile / g.tiles_per_frame, dy0 = (tile - frame * g.tiles_per_frame) * g.TH;
        const int th = min(g.TH, g.ho - dy0);
        const uint32_t ybuf = stages_s + b * (g.ystage + g.cstage);
        const uint32_t cbuf = ybuf + g.ystage;
        char* out[NCOL];
        char* const row0 = kDense ? reinterpret_cast<char*>(reinterpret_cast<typename Ops::Lut*>(dst) + ((size_t)frame * 3 * g.ho + dy0) * g.wo)
                                  : reinterpret_cast<char*>(reinterpret_cast<typename Ops::Lut*>(dst) + ((size_t)frame * 3 * g.canvas_h + g.y0 + dy0) * g.canvas_w + g.x0);
#pragma unroll
        for (int j = 0; j < NCOL; ++j) out[j] = row0 + (int)Ops::kElem * colx[j];
        // measured on B200 (bench_ops.py c2 / c2g / c2s): the packed variant wins for 16-bit outputs (issue-bound: 0.312 -> 0.282 ms)
        // and loses for fp32 (0.367 -> 0.384 ms: it moves the blend from the FMA pipe's IMADs onto the already busier ALU pipe)
        if (NCOL % 2 == 0 && sizeof(typename Ops::Lut) == 2) {
            if (right) compute_tile_packed<FMT, OutT, true, NCOL>(ybuf, cbuf, sy_s, cy_s, lut_s, col, dy0, th, sy_pitch, sc_pitch, g.vstage_off, out, row_bytes, plane_bytes);
            else compute_tile_packed<FMT, OutT, false, NCOL>(ybuf, cbuf, sy_s, cy_s, lut_s, col, dy0, th, sy_pitch, sc_pitch, g.vstage_off, out, row_bytes, plane_bytes);
        } else {
            if (right) compute_tile_parity<FMT, OutT, true, NCOL>(ybuf, cbuf, sy_s, cy_s, lut_s, col, dy0, th, sy_pitch, sc_pitch, g.vstage_off, out, row_bytes, plane_bytes);
            else if (FMT == kFmtPlanar) compute_tile_parity<FMT, OutT, false, NCOL>(ybuf, cbuf, sy_s, cy_s, lut_s, col, dy0, th, sy_pitch, sc_pitch, g.vstage_off, out, row_bytes, plane_bytes);
            else compute_tile<FMT, OutT, false, NCOL>(ybuf, cbuf, sy_s, cy_s, lut_s, col, dy0, th, sy_pitch, sc_pitch, g.vstage_off, out, row_bytes, plane_bytes);
        }
        __syncthreads();   // all reads of stage b done -> it may be refilled by the next iteration's issue
    }
}

// Letterbox border: every canvas pixel outside the content rectangle gets the normalised pad colour of its plane.
// blockIdx.y = frame * 3 + plane; the CTAs of a plane stride over its border elements in the order top rows, bottom rows
// (both contiguous spans), then the strips left and right of the content.
template <typename T>
__global__ void letterbox_pad_kernel(T* __restrict__ dst, int canvas_w, int canvas_h, int x0, int y0, int cw, int ch, T pad0, T pad1, T pad2) {
    const int plane = blockIdx.y % 3;
    T* p = dst + (size_t)blockIdx.y * canvas_w * canvas_h;
    const T v = plane == 0 ? pad0 : plane == 1 ? pad1 : pad2;
    const int top = y0 * canvas_w, full = (canvas_h - ch) * canvas_w, side_w = canvas_w - cw;
    const int total = full + ch * side_w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int pos;
        if (i < top) pos = i;
        else if (i < full) pos = i + ch * canvas_w;
        else {
            const int j = i - full, r = j / side_w, xx = j - r * side_w;
            pos = (y0 + r) * canvas_w + (xx < x0 ? xx : xx + cw);
        }
        p[pos] = v;
    }
}

}  // namespace vacv
