// a5 for single-channel u8 planes (grey images, and CHW tensors, which the reference resizes plane by plane: resize.cpp:73-87):
// the persistent TMA pipeline of resize_pipe_u8c3.cuh with one byte per pixel.  Same tiles, same row lists / bands, same
// stage layout (ResizePipeGeom, tile_rows); per column the two tap bytes come from one or two aligned 32-bit shared-memory
// words (+ funnel shift) and one IDP.2A forms L*cx0 + R*cx1; a warp's 32 output bytes leave as 8 words.
#pragma once
#include "resize_pipe_u8c3.cuh"

namespace vacv {

template <bool kSigned, int NCOL, bool kBand>
__global__ void __launch_bounds__(kRpThreads, 2)
resize_linear_u8c1_pipe_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, ResizePipeGeom g) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];     // [s_sy: ho][s_cy: ho][s_slot: ho][s_tile][pad] 2 x stage, then per-warp output lines
    int* s_sy = reinterpret_cast<int*>(dyn_smem);
    int* s_cy = s_sy + g.ho;
    int* s_slot = s_cy + g.ho;
    int* s_tile = s_slot + g.ho;
    uint8_t* stages = dyn_smem + g.table_bytes;
    __shared__ __align__(8) uint64_t full_bar[2];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* line = reinterpret_cast<uint32_t*>(stages + 2 * (size_t)g.stage_bytes) + warp * (NCOL * 8);   // per warp: NCOL x 32 bytes
    const unsigned row_bytes = (unsigned)g.w;
    const uint32_t stages_s = smem_u32(stages), sy_s = smem_u32(s_sy), cy_s = smem_u32(s_cy), slot_s = smem_u32(s_slot);

    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        int* tl = s_tile + tt * (1 + 2 * kRpMaxTH);
        tl[0] = n;
        for (int i = 0; i < n; ++i) tl[1 + i] = rows[i];
    }
    __syncthreads();
    unsigned aw[NCOL];      // byte offset of the aligned word holding the left tap, inside a source row
    int sh[NCOL];           // bit shift of the left tap inside that word
    uint32_t cx[NCOL];      // cx0 | cx1 << 16
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int dx = min(tid + j * nthr, g.wo - 1);
        int sx; float fx;
        linear_coord(dx, scale_x, g.w, sx, fx);
        cx[j] = (uint32_t)(sat_short((1.f - fx) * 2048.f) & 0xffff) | ((uint32_t)sat_short(2048.f * fx) << 16);
        aw[j] = (unsigned)sx & ~3u;
        sh[j] = (int)((unsigned)sx & 3u) * 8;
    }

    // Copy issue by the whole first warp (see resize_pipe_u8c3.cuh): lane 0 arms the barrier with the tile's byte count; band staging is ONE copy (lane 0),
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
    // horizontal sums of one source row for this thread's columns: L * cx0 + R * cx1
    auto hrow = [&](uint32_t rowaddr, int (&H)[NCOL]) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const uint32_t p = rowaddr + aw[j];
            uint32_t w0, w1;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(p));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(p));   // inside the stage (rows are followed by rows or slack)
            const uint32_t b = __funnelshift_r(w0, w1, sh[j]);               // [L R . .]
            H[j] = kSigned ? __dp2a_lo((int)cx[j], (int)b, 0) : (int)__dp2a_lo(cx[j], b, 0u);
        }
    };

    int tile = blockIdx.x;
    if (tid < 32 && tile < g.total_tiles) issue(tile, 0);
    __syncthreads();

    for (int it = 0; tile < g.total_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const int next = tile + gridDim.x;
        if (tid < 32 && next < g.total_tiles) issue(next, b ^ 1);   // stage b^1 was released by the sync below
        mbar_wait(&full_bar[b], (it >> 1) & 1);
        const int frame = tile / g.tiles_per_frame, dy0 = (tile - frame * g.tiles_per_frame) * g.TH;
        const int th = min(g.TH, g.ho - dy0);
        const uint32_t buf = stages_s + b * g.stage_bytes;
        const int y_first = kBand ? lds_s32(sy_s + 4 * dy0) : 0;
        uint8_t* orow = dst + (size_t)frame * g.dst_image + (size_t)dy0 * g.wo;
        int H0[NCOL], H1[NCOL];
        int have = -2;   // source row whose sums H1 holds (-2: none).  All threads walk the same rows: no divergence.
        for (int ty = 0; ty < th; ++ty, orow += (size_t)g.wo) {
            const int sy = lds_s32(sy_s + 4 * (dy0 + ty));
            const int cy = lds_s32(cy_s + 4 * (dy0 + ty));
            const int cy0 = (short)(cy & 0xffff), cy1 = cy >> 16;
            if (kBand) {   // contiguous band: slot = row - first row; a zero-weight lower row is staged anyway and contributes H1 * 0
                if (sy == have) {
#pragma unroll
                    for (int j = 0; j < NCOL; ++j) H0[j] = H1[j];
                } else if (sy + 1 != have) {
                    hrow(buf + (unsigned)(sy - y_first) * row_bytes, H0);
                }
                if (sy + 1 != have) {
                    hrow(buf + (unsigned)(sy + 1 - y_first) * row_bytes, H1);
                    have = sy + 1;
                }
            } else {
                const uint32_t upper = buf + (unsigned)lds_s32(slot_s + 4 * (dy0 + ty)) * row_bytes;   // staged upper tap row; the lower one follows it
                if (sy == have) {
#pragma unroll
                    for (int j = 0; j < NCOL; ++j) H0[j] = H1[j];
                    have = -2;
                } else if (sy + 1 != have || cy1 == 0) {
                    hrow(upper, H0);
                }
                if (cy1 != 0) {
                    if (sy + 1 != have) { hrow(upper + row_bytes, H1); have = sy + 1; }
                } else {   // zero-weight lower tap: not staged, contributes H1 * 0
#pragma unroll
                    for (int j = 0; j < NCOL; ++j) H1[j] = 0;
                    have = -2;
                }
            }
            uint8_t* lb = reinterpret_cast<uint8_t*>(line);
#pragma unroll
            for (int j = 0; j < NCOL; ++j) lb[j * 32 + lane] = (uint8_t)((H0[j] * cy0 + H1[j] * cy1) >> 22);   // resize_naive.cpp:60-65
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                const int c0 = (tid & ~31) + j * nthr;             // first column of this warp's j-th group
                if (c0 >= g.wo) continue;                          // warp-uniform
                uint8_t* o = orow + c0;
                const int n = min(32, g.wo - c0);
                if (n == 32 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) { if (lane < 8) st_stream4(o + 4 * lane, line[j * 8 + lane]); }
                else if (lane < n) o[lane] = lb[j * 32 + lane];
            }
            __syncwarp();
        }
        __syncthreads();   // all reads of stage b done -> it may be refilled by the next iteration's issue
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Second generation for planes: a thread owns QUADS of four adjacent output columns.
// ncu of the kernel above at 1080p -> 640x360 planes (profiles/r2_linchw_ncu_raw.txt): 52 thread-instructions per output BYTE and
// stalled on CTA barriers -- with one byte per pixel the per-row bookkeeping (row tables, per-warp staging line, two warp
// syncs, 32-byte stores by 8 lanes) and the per-tile barrier are paid for very little payload.  Here
//   * the four blended bytes of a quad are packed with three PRMT and leave as ONE 32-bit streaming store, lane-contiguous
//     (a warp writes 128 contiguous bytes): no staging line, no warp sync;
//   * the CTA's threads form RY row groups of TX threads; a group walks its own run of consecutive output rows of the tile (the
//     carry-over of a source row's horizontal sums works inside a run), so tiles can be taller (up to 16 rows) and the per-tile
//     barrier is amortised over more output;
//   * horizontal sums as before: two aligned 32-bit shared loads + funnel shift + one IDP.2A per column and source row.
// Arithmetic unchanged (resize_naive.cpp:17-64).  Needs w_out % 4 == 0 and a 4-byte aligned destination.
constexpr int kQuadMaxTH = 16;
constexpr int kQuadStages = 2;     // shared-memory stages: two tiles in flight while one is computed (tiles of planes are small)

template <bool kSigned, bool kBand, int NQ>
__global__ void __launch_bounds__(kRpThreads, 2)
resize_linear_u8c1_quad_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, ResizePipeGeom g, int TX) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];     // [s_sy: ho][s_cy: ho][s_slot: ho][s_tile][pad] 2 x stage
    int* s_sy = reinterpret_cast<int*>(dyn_smem);
    int* s_cy = s_sy + g.ho;
    int* s_slot = s_cy + g.ho;
    int* s_tile = s_slot + g.ho;
    uint8_t* stages = dyn_smem + g.table_bytes;
    __shared__ __align__(8) uint64_t full_bar[kQuadStages];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int ry = tid / TX, tx = tid - ry * TX, RY = nthr / TX;
    const int rows_per_group = (g.TH + RY - 1) / RY;
    const unsigned row_bytes = (unsigned)g.w;
    const uint32_t stages_s = smem_u32(stages), sy_s = smem_u32(s_sy), cy_s = smem_u32(s_cy), slot_s = smem_u32(s_slot);

    if (tid == 0) {
        for (int i = 0; i < kQuadStages; ++i) mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const double scale_x = (double)((float)g.w / (float)g.wo), scale_y = (double)((float)g.h / (float)g.ho);
    for (int dy = tid; dy < g.ho; dy += nthr) {
        int s; float f;
        linear_coord(dy, scale_y, g.h, s, f);
        s_sy[dy] = s;
        s_cy[dy] = sat_short((1.f - f) * 2048.f) | (sat_short(2048.f * f) << 16);
    }
    __syncthreads();
    for (int tt = tid; tt < (kBand ? 0 : g.tiles_per_frame); tt += nthr) {   // stage slot of every output row's upper tap row
        int rows[2 * kQuadMaxTH], slot[kQuadMaxTH];
        const int dy0 = tt * g.TH, th = min(g.TH, g.ho - dy0);
        const int n = tile_rows(s_sy, s_cy, dy0, th, rows, slot);
        for (int ty = 0; ty < th; ++ty) s_slot[dy0 + ty] = slot[ty];
        int* tl = s_tile + tt * (1 + 2 * kQuadMaxTH);
        tl[0] = n;
        for (int i = 0; i < n; ++i) tl[1 + i] = rows[i];
    }
    __syncthreads();
    const int quads = g.wo >> 2;
    unsigned aw[NQ][4];     // byte offset of the aligned word holding the left tap, inside a source row
    int sh[NQ][4];          // bit shift of the left tap inside that word
    uint32_t cx[NQ][4];     // cx0 | cx1 << 16
#pragma unroll
    for (int j = 0; j < NQ; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int dx = min(4 * (tx + j * TX) + i, g.wo - 1);
            int sx; float fx;
            linear_coord(dx, scale_x, g.w, sx, fx);
            cx[j][i] = (uint32_t)(sat_short((1.f - fx) * 2048.f) & 0xffff) | ((uint32_t)sat_short(2048.f * fx) << 16);
            aw[j][i] = (unsigned)sx & ~3u;
            sh[j][i] = (int)((unsigned)sx & 3u) * 8;
        }

    // Copy issue by the whole first warp: lane 0 arms the barrier with the tile's byte count, then lane i issues the bulk copy of the
    // tile's i-th source row (row-list staging: up to 32 rows that are NOT consecutive in memory -- one thread issuing them one
    // after the other took longer than the tile's arithmetic: 1.8 TB/s at 1080p -> 640x360 planes).  Band staging stays one copy.
    auto issue = [&](int tile, int b) {   // called by every lane of warp 0
        const int lane = tid & 31;
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
        const int* rows = s_tile + (tile - frame * g.tiles_per_frame) * (1 + 2 * kQuadMaxTH) + 1;
        const int n = rows[-1];
        if (lane == 0) mbar_expect_tx(&full_bar[b], (uint32_t)n * row_bytes);
        __syncwarp();
        if (lane < n)
            bulk_g2s(stages + (size_t)b * g.stage_bytes + (size_t)lane * row_bytes, src + (size_t)frame * g.src_image + (size_t)rows[lane] * row_bytes,
                     row_bytes, &full_bar[b]);
    };
    // horizontal sums of one source row for this thread's columns: L * cx0 + R * cx1
    auto hrow = [&](uint32_t rowaddr, int (&H)[NQ][4]) {
#pragma unroll
        for (int j = 0; j < NQ; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t p = rowaddr + aw[j][i];
                uint32_t w0, w1;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(p));
                asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(p));   // inside the stage (rows are followed by rows or slack)
                const uint32_t b = __funnelshift_r(w0, w1, sh[j][i]);           // [L R . .]
                H[j][i] = kSigned ? __dp2a_lo((int)cx[j][i], (int)b, 0) : (int)__dp2a_lo(cx[j][i], b, 0u);
            }
    };

    int tile = blockIdx.x;
    if (tid < 32)
        for (int i = 0; i < kQuadStages - 1; ++i)
            if (tile + i * (int)gridDim.x < g.total_tiles) issue(tile + i * (int)gridDim.x, i);
    __syncthreads();

    for (int it = 0, b = 0, par = 0; tile < g.total_tiles; tile += gridDim.x, ++it) {
        const int next = tile + (kQuadStages - 1) * (int)gridDim.x;
        const int bn = b == 0 ? kQuadStages - 1 : b - 1;             // stage of tile it - 1: released by the sync at the end of that iteration
        if (tid < 32 && next < g.total_tiles) issue(next, bn);
        mbar_wait(&full_bar[b], par);
        const int frame = tile / g.tiles_per_frame, dy0 = (tile - frame * g.tiles_per_frame) * g.TH;
        const int th = min(g.TH, g.ho - dy0);
        const uint32_t buf = stages_s + b * g.stage_bytes;
        const int y_first = kBand ? lds_s32(sy_s + 4 * dy0) : 0;
        const int ty0 = ry * rows_per_group, ty1 = min(th, ty0 + rows_per_group);
        uint8_t* orow = dst + (size_t)frame * g.dst_image + (size_t)(dy0 + ty0) * g.wo;
        int H0[NQ][4], H1[NQ][4];
        int have = -2;   // source row whose sums H1 holds (-2: none).  All threads of a row group walk the same rows.
        for (int ty = ty0; ty < ty1; ++ty, orow += (size_t)g.wo) {
            const int sy = lds_s32(sy_s + 4 * (dy0 + ty));
            const int cy = lds_s32(cy_s + 4 * (dy0 + ty));
            const int cy0 = (short)(cy & 0xffff), cy1 = cy >> 16;
            if (kBand) {   // contiguous band: slot = row - first row; a zero-weight lower row is staged anyway and contributes H1 * 0
                if (sy == have) {
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) H0[j][i] = H1[j][i];
                } else if (sy + 1 != have) {
                    hrow(buf + (unsigned)(sy - y_first) * row_bytes, H0);
                }
                if (sy + 1 != have) {
                    hrow(buf + (unsigned)(sy + 1 - y_first) * row_bytes, H1);
                    have = sy + 1;
                }
            } else {
                const uint32_t upper = buf + (unsigned)lds_s32(slot_s + 4 * (dy0 + ty)) * row_bytes;   // staged upper tap row; the lower one follows it
                if (sy == have) {
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) H0[j][i] = H1[j][i];
                    have = -2;
                } else if (sy + 1 != have || cy1 == 0) {
                    hrow(upper, H0);
                }
                if (cy1 != 0) {
                    if (sy + 1 != have) { hrow(upper + row_bytes, H1); have = sy + 1; }
                } else {   // zero-weight lower tap: not staged, contributes H1 * 0
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) H1[j][i] = 0;
                    have = -2;
                }
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int q = tx + j * TX;
                int v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = (H0[j][i] * cy0 + H1[j][i] * cy1) >> 22;   // resize_naive.cpp:60-65
                const uint32_t word = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
                if (q < quads) st_stream4(orow + 4 * q, word);
            }
        }
        __syncthreads();   // all reads of stage b done -> it may be refilled by the next iteration's issue
        if (++b == kQuadStages) { b = 0; par ^= 1; }
    }
}

}  // namespace vacv
