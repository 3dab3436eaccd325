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

    auto issue = [&](int tile, int b) {   // one thread: one bulk copy per run of consecutive source rows
        const int frame = tile / g.tiles_per_frame;
        if (kBand) {
            const int dy0 = (tile - frame * g.tiles_per_frame) * g.TH, th = min(g.TH, g.ho - dy0);
            const int y_first = s_sy[dy0], y_last = s_sy[dy0 + th - 1] + 1;
            const uint32_t bytes = (uint32_t)(y_last - y_first + 1) * row_bytes;
            mbar_expect_tx(&full_bar[b], bytes);
            bulk_g2s(stages + (size_t)b * g.stage_bytes, src + (size_t)frame * g.src_image + (size_t)y_first * row_bytes, bytes, &full_bar[b]);
            return;
        }
        const int* rows = s_tile + (tile - frame * g.tiles_per_frame) * (1 + 2 * kRpMaxTH) + 1;
        const int n = rows[-1];
        mbar_expect_tx(&full_bar[b], (uint32_t)n * row_bytes);
        const uint8_t* f = src + (size_t)frame * g.src_image;
        uint8_t* st = stages + (size_t)b * g.stage_bytes;
        for (int i = 0; i < n;) {
            int j = i + 1;
            while (j < n && rows[j] == rows[j - 1] + 1) ++j;
            bulk_g2s(st + (size_t)i * row_bytes, f + (size_t)rows[i] * row_bytes, (uint32_t)(j - i) * row_bytes, &full_bar[b]);
            i = j;
        }
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
    if (tid == 0 && tile < g.total_tiles) issue(tile, 0);
    __syncthreads();

    for (int it = 0; tile < g.total_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const int next = tile + gridDim.x;
        if (tid == 0 && next < g.total_tiles) issue(next, b ^ 1);   // stage b^1 was released by the sync below
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

}  // namespace vacv
