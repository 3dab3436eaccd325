// a5 for interleaved 3-channel u8: persistent CTAs + TMA row staging, the structure of the fused NV12 pipeline
// (fused_pipeline.cuh) applied to plain BGR resize.
//
// The gather kernel (resize_linear_u8c3_kernel) costs ~100 instructions per output pixel because every pixel fetches and
// blends its four taps from scratch.  Here a CTA owns tiles of TH output rows x the full width of one frame; exactly the
// source rows the tile's outputs give weight to (tile_rows() below: a zero-weight lower tap row is NOT fetched, so an integer
// ratio such as 1080 -> 360 reads one row in three) arrive by bulk copies (cp.async.bulk, UBLKCP), one per run of consecutive
// rows, into one of two shared-memory stages while the previous tile is computed.  A thread owns NCOL output columns and
// walks down the rows:
//   * the horizontally blended sums of a source row (two aligned 32-bit shared loads + funnel shift -> PRMT -> three
//     IDP.2A with the packed 16-bit weights) are computed ONCE per source row and column; when the next output row starts
//     on the previous lower row, its sums are carried over instead of recomputed;
//   * vertical blend = the reference's integer expression regrouped row-wise, (Ht*cy0 + Hb*cy1) >> 22
//     (resize_naive.cpp:60-65; same integer, < 2^31);
//   * a warp's 32 consecutive output pixels are re-chunked through a private shared-memory line into one lane-contiguous
//     96-byte store.
#pragma once
#include "fused_pipeline.cuh"
#include "gather_u8c3.cuh"

namespace vacv {

constexpr int kRpThreads = 384;   // max threads per CTA
constexpr int kRpMaxCols = 4;     // output columns per thread

constexpr int kRpMaxTH = 8;

// The source rows a tile of output rows [dy0, dy0 + th) needs, in increasing order: sy of every output row, and sy + 1 where
// the lower tap has weight (cy1 != 0).  slot[ty] (optional) = position of output row ty's upper tap row in that list; its
// lower tap row, when needed, is the next entry.  Shared by the launcher (stage size), the kernel's slot table and the copy
// issue, so all three agree by construction.  sy / cy: the per-output-row tables (cy = cy0 | cy1 << 16).
__host__ __device__ inline int tile_rows(const int* sy, const int* cy, int dy0, int th, int* rows, int* slot) {
    int n = 0;
    for (int ty = 0; ty < th; ++ty) {
        const int r = sy[dy0 + ty];
        int pos = n;
        for (int i = n - 1; i >= 0 && rows[i] >= r; --i) if (rows[i] == r) pos = i;   // at most two steps back
        if (pos == n) rows[n++] = r;
        if (slot) slot[ty] = pos;
        if ((cy[dy0 + ty] >> 16) != 0 && (pos + 1 >= n || rows[pos + 1] != r + 1)) rows[n++] = r + 1;   // rows stay sorted: r + 1 > everything before
    }
    return n;
}

struct ResizePipeGeom {
    int w, h, wo, ho;
    int TH, tiles_per_frame, total_tiles;
    int stage_bytes;        // bytes reserved per stage (multiple of 128)
    int table_bytes;        // row tables at the start of dynamic shared memory (multiple of 128)
    size_t src_image, dst_image;   // bytes between images
};

// kBand: every source row between a tile's first and last tap row is needed (small ratios): the stage is that contiguous band,
// fetched with ONE bulk copy, row slot = row - first row.  !kBand (integer ratios, where no lower tap has weight): only the
// listed rows are fetched, one bulk copy per run of consecutive rows.
// kPoint (with !kBand): additionally no right tap has weight (odd integer ratio in x as well, e.g. 1920x1080 -> 640x360): every
// weight pair is (2048, 0), the reference's integer blend (p * 2048 * 2048) >> 22 returns the tap itself and the kernel only
// moves bytes.
// OUT: kRpOutU8 = the resized u8 BGR image (a5); kRpOutF32CHW / kRpOutF32HWC = resize_normalize (a13): every blended byte goes
// through the exact 3 x 256 normalisation table (built once per persistent CTA) and leaves as fp32 planes (lane-contiguous
// 128-byte stores per plane) or as interleaved fp32 (re-chunked per warp like the u8 bytes).
enum { kRpOutU8 = 0, kRpOutF32CHW = 1, kRpOutF32HWC = 2 };

template <bool kSigned, int NCOL, bool kBand, bool kPoint = false, int OUT = kRpOutU8>
__global__ void __launch_bounds__(kRpThreads, NCOL <= 2 || !kBand ? 2 : 1)
resize_linear_u8c3_pipe_kernel(const uint8_t* __restrict__ src, void* __restrict__ dst_, ResizePipeGeom g,
                               const float* __restrict__ mean, const float* __restrict__ stddev) {
    static_assert(!kPoint || OUT == kRpOutU8, "the byte-moving variant exists for u8 output only");
    uint8_t* const dst = reinterpret_cast<uint8_t*>(dst_);
    float* const dstf = reinterpret_cast<float*>(dst_);
    __shared__ float lut[OUT == kRpOutU8 ? 1 : 768];
    extern __shared__ __align__(128) uint8_t dyn_smem[];     // [s_sy: ho][s_cy: ho][pad] 2 x stage, then per-warp output lines
    int* s_sy = reinterpret_cast<int*>(dyn_smem);
    int* s_cy = s_sy + g.ho;
    int* s_slot = s_cy + g.ho;
    int* s_tile = s_slot + g.ho;                             // [tiles_per_frame][1 + 2 * kRpMaxTH]: row count, rows
    uint8_t* stages = dyn_smem + g.table_bytes;
    __shared__ __align__(8) uint64_t full_bar[2];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kLineWords = OUT == kRpOutF32HWC ? 96 : 24;   // per warp and column group: 32 pixels x 3 channels as bytes / floats
    uint32_t* line = reinterpret_cast<uint32_t*>(stages + 2 * (size_t)g.stage_bytes) + warp * (NCOL * kLineWords);
    const unsigned row_bytes = (unsigned)g.w * 3u;
    const uint32_t stages_s = smem_u32(stages), sy_s = smem_u32(s_sy), cy_s = smem_u32(s_cy), slot_s = smem_u32(s_slot);

    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (OUT != kRpOutU8)
        for (int t = tid; t < 768; t += nthr)
            lut[t] = normalize_one((float)(t & 255), __ldg(mean + (t >> 8)), (double)__ldg(stddev + (t >> 8)) + 1e-6);
    const double scale_x = (double)((float)g.w / (float)g.wo), scale_y = (double)((float)g.h / (float)g.ho);
    for (int dy = tid; dy < g.ho; dy += nthr) {
        int s; float f;
        linear_coord(dy, scale_y, g.h, s, f);
        s_sy[dy] = s;
        s_cy[dy] = sat_short((1.f - f) * 2048.f) | (sat_short(2048.f * f) << 16);
    }
    __syncthreads();
    for (int tt = tid; tt < (kBand ? 0 : g.tiles_per_frame); tt += nthr) {   // stage slot of every output row's upper tap row
        int rows[2 * kRpMaxTH], slot[kRpMaxTH];
        const int dy0 = tt * g.TH, th = min(g.TH, g.ho - dy0);
        const int n = tile_rows(s_sy, s_cy, dy0, th, rows, slot);
        for (int ty = 0; ty < th; ++ty) s_slot[dy0 + ty] = slot[ty];
        int* tl = s_tile + tt * (1 + 2 * kRpMaxTH);            // the copy issue of every tile just replays this list
        tl[0] = n;
        for (int i = 0; i < n; ++i) tl[1 + i] = rows[i];
    }
    __syncthreads();
    // thread t owns columns t, t + nthr, ...; columns past the row end are clamped (computed, not stored)
    unsigned aw[NCOL];      // byte offset of the aligned word holding the first tap byte, inside a source row
    int sh[NCOL];           // bit shift of the first tap byte inside that word
    uint32_t cx[NCOL];      // cx0 | cx1 << 16
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int dx = min(tid + j * nthr, g.wo - 1);
        int sx; float fx;
        linear_coord(dx, scale_x, g.w, sx, fx);
        cx[j] = (uint32_t)(sat_short((1.f - fx) * 2048.f) & 0xffff) | ((uint32_t)sat_short(2048.f * fx) << 16);
        aw[j] = ((unsigned)sx * 3u) & ~3u;
        sh[j] = (int)(((unsigned)sx * 3u) & 3u) * 8;
    }

    // Copy issue by the whole first warp: lane 0 arms the barrier with the tile's byte count; band staging is ONE copy (lane 0),
    // row-list staging one copy per listed row, lane i issuing the i-th (the rows are not consecutive in memory, and one thread
    // issuing them one after the other was slower than the tile's arithmetic).
    auto issue = [&](int tile, int b) {   // called by every lane of warp 0
        const int frame = tile / g.tiles_per_frame;
        if (kBand) {
            if (lane == 0) {
                const int dy0 = (tile - frame * g.tiles_per_frame) * g.TH, th = min(g.TH, g.ho - dy0);
                const int y_first = s_sy[dy0], y_last = s_sy[dy0 + th - 1] + 1;
                const uint32_t bytes = (uint32_t)(y_last - y_first + 1) * row_bytes;
                mbar_expect_tx(&full_bar[b], bytes);
                bulk_g2s(stages + (size_t)b * g.stage_bytes, src + (size_t)frame * g.src_image + (size_t)y_first * row_bytes, bytes, &full_bar[b]);
            }
            return;
        }
        const int* rows = s_tile + (tile - frame * g.tiles_per_frame) * (1 + 2 * kRpMaxTH) + 1;
        const int n = rows[-1];   // <= 2 * kRpMaxTH = 16
        if (lane == 0) mbar_expect_tx(&full_bar[b], (uint32_t)n * row_bytes);
        __syncwarp();
        if (lane < n)
            bulk_g2s(stages + (size_t)b * g.stage_bytes + (size_t)lane * row_bytes, src + (size_t)frame * g.src_image + (size_t)rows[lane] * row_bytes,
                     row_bytes, &full_bar[b]);
    };
    // horizontal sums of one source row for this thread's columns
    auto hrow = [&](uint32_t rowaddr, int (&H)[NCOL][3]) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const uint32_t p = rowaddr + aw[j];
            uint32_t w0, w1, w2;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(p));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(p));
            asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(p));   // inside the stage (rows are followed by rows or slack)
            const uint32_t b0 = __funnelshift_r(w0, w1, sh[j]), b1 = __funnelshift_r(w1, w2, sh[j]);
            hsum_u8c3<kSigned>(b0, b1, cx[j], H[j]);
        }
    };

    int tile = blockIdx.x;
    if (tid < 32 && tile < g.total_tiles) issue(tile, 0);
    __syncthreads();
    // u8 output: all of this warp's column groups are full runs of 32 pixels and every output row starts word-aligned
    const bool all_fast = (tid & ~31) + (NCOL - 1) * nthr + 32 <= g.wo && ((g.wo * 3) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0;

    for (int it = 0; tile < g.total_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const int next = tile + gridDim.x;
        if (tid < 32 && next < g.total_tiles) issue(next, b ^ 1);   // stage b^1 was released by the sync below
        mbar_wait(&full_bar[b], (it >> 1) & 1);
        const int frame = tile / g.tiles_per_frame, dy0 = (tile - frame * g.tiles_per_frame) * g.TH;
        const int th = min(g.TH, g.ho - dy0);
        const uint32_t buf = stages_s + b * g.stage_bytes;
        const int y_first = kBand ? lds_s32(sy_s + 4 * dy0) : 0;
        uint8_t* orow = dst + (size_t)frame * g.dst_image + (size_t)dy0 * g.wo * 3;
        // Horizontal sums live in two register sets: Ha holds an EVEN source row, Hb an ODD one (have_a / have_b: which).  An output
        // row blends row sy (upper) with row sy + 1 (lower), one of each parity; when the next output row starts on the previous lower
        // row, that row's sums are already where its parity puts them -- nothing is copied (the round-1 loop moved H1 into H0: 8 of
        // its 63 instructions per pixel were register moves, profiles/r2_lin720_ncu_raw.txt).  All threads walk the same rows.
        int Ha[NCOL][3] = {}, Hb[NCOL][3] = {};
        int have_a = -2, have_b = -2;
        // vertical blend + store of one output row (Hu = sums of the upper tap row, Hl = of the lower one)
        auto blend_store = [&](const int (&Hu)[NCOL][3], const int (&Hl)[NCOL][3], int cy0, int cy1, int ty, uint8_t* orow) {
            if (OUT == kRpOutU8) {
                uint8_t* lb = reinterpret_cast<uint8_t*>(line);
                if (!kPoint) {
#pragma unroll
                    for (int j = 0; j < NCOL; ++j) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) lb[j * 96 + 3 * lane + k] = (uint8_t)((Hu[j][k] * cy0 + Hl[j][k] * cy1) >> 22);   // resize_naive.cpp:60-65
                    }
                }
                __syncwarp();
                if (all_fast) {   // every group of this warp is a full, word-aligned run of 32 pixels: 24 lanes x 4 bytes each
#pragma unroll
                    for (int j = 0; j < NCOL; ++j)
                        if (lane < 24) st_stream4(orow + (size_t)((tid & ~31) + j * nthr) * 3 + 4 * lane, line[j * 24 + lane]);
                } else {
#pragma unroll
                    for (int j = 0; j < NCOL; ++j) {
                        const int c0 = (tid & ~31) + j * nthr;             // first column of this warp's j-th group
                        if (c0 >= g.wo) continue;                          // warp-uniform
                        uint8_t* o = orow + (size_t)c0 * 3;
                        const int n = min(32, g.wo - c0);
                        if (n == 32 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) { if (lane < 24) st_stream4(o + 4 * lane, line[j * 24 + lane]); }
                        else for (int bb = lane; bb < 3 * n; bb += 32) o[bb] = lb[j * 96 + bb];
                    }
                }
                __syncwarp();
            } else {
                // the u8 value the unfused resize would have stored, then the exact table
                const size_t plane = (size_t)g.wo * g.ho;
                const size_t prow = (size_t)(dy0 + ty) * g.wo;
#pragma unroll
                for (int j = 0; j < NCOL; ++j) {
                    const int c0 = (tid & ~31) + j * nthr, col = c0 + lane;
                    if (c0 >= g.wo) continue;                          // warp-uniform
                    float r[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) r[k] = lut[k * 256 + (((Hu[j][k] * cy0 + Hl[j][k] * cy1) >> 22) & 0xff)];
                    if (OUT == kRpOutF32CHW) {
                        float* o = dstf + (size_t)frame * 3 * plane + prow + col;
                        if (col < g.wo) { st_stream4f(o, r[0]); st_stream4f(o + plane, r[1]); st_stream4f(o + 2 * plane, r[2]); }
                    } else {
                        float* sf = reinterpret_cast<float*>(line) + j * 96;
                        sf[3 * lane] = r[0]; sf[3 * lane + 1] = r[1]; sf[3 * lane + 2] = r[2];
                        __syncwarp();
                        float* o = dstf + ((size_t)frame * plane + prow + c0) * 3;
                        const int n = min(32, g.wo - c0);
                        if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                            if (4 * lane < 3 * n) st_stream16f(o + 4 * lane, *reinterpret_cast<const float4*>(sf + 4 * lane));
                        } else {
                            for (int e = lane; e < 3 * n; e += 32) st_stream4f(o + e, sf[e]);
                        }
                        __syncwarp();
                    }
                }
            }
        };
        for (int ty = 0; ty < th; ++ty, orow += (size_t)g.wo * 3) {
            const int sy = lds_s32(sy_s + 4 * (dy0 + ty));
            const int cy = lds_s32(cy_s + 4 * (dy0 + ty));
            const int cy0 = (short)(cy & 0xffff), cy1 = cy >> 16;
            if (kPoint) {  // all weights are (2048, 0) x (2048, 0): (p << 22) >> 22 = p, signed or not
                const uint32_t upper = buf + (unsigned)lds_s32(slot_s + 4 * (dy0 + ty)) * row_bytes;
                uint8_t* lbp = reinterpret_cast<uint8_t*>(line);
#pragma unroll
                for (int j = 0; j < NCOL; ++j) {
                    uint32_t w0, w1;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(upper + aw[j]));
                    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(upper + aw[j]));
                    const uint32_t px = __funnelshift_r(w0, w1, sh[j]);
                    lbp[j * 96 + 3 * lane] = (uint8_t)px; lbp[j * 96 + 3 * lane + 1] = (uint8_t)(px >> 8); lbp[j * 96 + 3 * lane + 2] = (uint8_t)(px >> 16);
                }
                blend_store(Ha, Hb, cy0, cy1, ty, orow);
                continue;
            }
            // staged address of the upper tap row; the lower one follows it (band: slot = row - first row; row list: the slot table).
            // A lower tap row of weight 0 is not filtered (row list: not even staged): whatever its register set holds is multiplied by 0.
            const uint32_t upper = kBand ? buf + (unsigned)(sy - y_first) * row_bytes : buf + (unsigned)lds_s32(slot_s + 4 * (dy0 + ty)) * row_bytes;
            if ((sy & 1) == 0) {
                if (have_a != sy) { hrow(upper, Ha); have_a = sy; }
                if (cy1 != 0 && have_b != sy + 1) { hrow(upper + row_bytes, Hb); have_b = sy + 1; }
                blend_store(Ha, Hb, cy0, cy1, ty, orow);
            } else {
                if (have_b != sy) { hrow(upper, Hb); have_b = sy; }
                if (cy1 != 0 && have_a != sy + 1) { hrow(upper + row_bytes, Ha); have_a = sy + 1; }
                blend_store(Hb, Ha, cy0, cy1, ty, orow);
            }
        }
        __syncthreads();   // all reads of stage b done -> it may be refilled by the next iteration's issue
    }
}

}  // namespace vacv
