// a10 / a13: warp_affine (+ normalize) of interleaved u8 BGR frames with the SOURCE WINDOW STAGED IN SHARED MEMORY BY TMA.
// Reference arithmetic: src/cv/warp_affine_naive.cpp:9-58 (identical to warp.cu, whose gather kernel stays the fallback).
//
//   work item = one output tile (TW x TH pixels, about 28 x 28: square tiles keep the bounding box of a rotated
//               footprint small) of one crop; persistent CTAs stride over the items.
//   producer  = the last warp.  Its lanes evaluate the affine map at the tile's four corners with the per-pixel fp32
//               expression (each rounding step is monotone in dx and in dy, so the extremes over the tile sit on the
//               corners): that is the exact range of tap columns / rows of the tile.  The window is fetched by 3-D tiled
//               TMA copies (cp.async.bulk.tensor, SASS UTMALDG) of 16 rows each from a tensor map over the frame pool
//               {w*3/4 words, h, frames}; the box width is picked per tile from a small family of maps (144 ... 528
//               bytes, odd multiples of 16 B so consecutive rows start in different banks).  Parts of a box outside the
//               frame are zero-filled by the hardware and never read.  The geometry of four tiles is computed at once
//               (8 lanes each), one pass ahead of the copies.  The windows live in a shared-memory RING: a tile
//               takes exactly the bytes of its boxes, so several windows (typically 3-5, at most 8) are in flight while
//               the consumer warps work on the oldest one (full / empty mbarriers per tile slot, no CTA barrier).
//   consumers = a lane owns a pixel (flat order inside the tile), forms coordinates and weights exactly like the gather
//               kernel, reads its 2 x 6 tap bytes from the staged window as aligned 32-bit shared-memory words
//               (+ funnel shift), PRMT + IDP.2A horizontal sums, the reference's integer blend, exact 3 x 256
//               normalisation table (built once per persistent CTA), outputs re-chunked per warp for coalesced stores.
//   Tiles whose window does not fit the ring (tiny scales, extreme shear) or whose coordinates are not finite
//   run the direct global gather inside the same kernel (CTA-uniform branch).
#pragma once
#include <type_traits>

#include <cuda.h>   // CUtensorMap (type only; the encoder is resolved at run time through cudaGetDriverEntryPoint)

#include "fused_pipeline.cuh"   // mbarrier / shared-window helpers
#include "gather_u8c3.cuh"
#include "vacv_common.cuh"

namespace vacv {

constexpr int kWsMaps = 7;             // box widths 144 + 64 k bytes
constexpr int kWsBoxRows = 16;
constexpr int kWsRingBytes = 100 * 1024;   // shared-memory ring of source windows per CTA
constexpr int kWsCtasPerSm = 2;            // 2 x 100 KB measured best (ring depth beats occupancy)
constexpr int kWsSlots = 8;                // tiles in flight per CTA (descriptor + full / empty barrier each)
__host__ __device__ constexpr int ws_box_bytes(int k) { return 144 + 64 * k; }

struct WarpStagedMaps { CUtensorMap m[kWsMaps]; };

struct WarpStagedGeom {
    int w, h, wo, ho;
    int tw, th, tiles_x, tiles_per_crop;   // tw <= 32, a multiple of 4
    int total_tiles;                       // tiles_per_crop * crops < 2^31 (host check)
    float inv_tiles_x;                     // t / tiles_x == (int)((t + 0.5f) * inv_tiles_x) for t < 2^20
    size_t frame_bytes;
};

struct __align__(16) WsDesc {          // written by the producer, read by the consumers after the full barrier
    int crop, frame, x0, y0;
    int tw, npx, sy_lo, bx_lo;
    int pitch, staged, magic, step_x;  // i / tw == (i * magic) >> 16 for i < 1024; (step_y, step_x) = divmod(consumer threads, tw)
    int step_y, pos, pad1, pad2;       // pos: byte offset of the window in the ring
    float m[8];
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
template <int OFF> __device__ __forceinline__ float lds_f32_at(uint32_t a) {
    float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF)); return v;
}

template <int NW> constexpr int ws_smem_bytes() { return kWsRingBytes + 768 * 4 + NW * 96 * 4 + kWsSlots * (int)sizeof(WsDesc) + 2 * kWsSlots * 8; }

// NW consumer warps + 1 producer warp
template <int OUT, bool kSigned, int NW>
__global__ void __launch_bounds__(32 * NW + 32, kWsCtasPerSm)
warp_affine_u8c3_staged_kernel(const __grid_constant__ WarpStagedMaps maps, const uint8_t* __restrict__ frames,
                               const int* __restrict__ frame_idx, const float* __restrict__ minv, void* __restrict__ dst_,
                               const WarpStagedGeom g, const float* __restrict__ mean, const float* __restrict__ stddev) {
    constexpr int NT = 32 * NW;
    extern __shared__ __align__(128) uint8_t ws_smem[];
    uint8_t* stages = ws_smem;                                                  // window ring, 128-byte aligned
    float* lut = reinterpret_cast<float*>(ws_smem + kWsRingBytes);              // 3 x 256 (unused for u8 output)
    uint32_t* ostage = reinterpret_cast<uint32_t*>(lut + 768);                  // NW x 96 words
    WsDesc* desc = reinterpret_cast<WsDesc*>(ostage + NW * 96);                 // kWsSlots
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(desc + kWsSlots);          // kWsSlots
    uint64_t* empty_bar = full_bar + kWsSlots;                                  // kWsSlots

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int j = 0; j < kWsSlots; ++j) { mbar_init(&full_bar[j], 1); mbar_init(&empty_bar[j], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (OUT != kWarpOutU8)
        for (int t = tid; t < 768; t += blockDim.x)
            lut[t] = normalize_one((float)(t & 255), __ldg(mean + (t >> 8)), (double)__ldg(stddev + (t >> 8)) + 1e-6);
    __syncthreads();
    const uint32_t stages_s = smem_u32(stages);

    if (warp == NW) {
        // ------------------------------------------------------------------ producer
        // Windows live in a ring: a tile takes exactly the bytes of its TMA boxes, so several windows (up to kWsSlots) are in
        // flight.  Tiles are released in order (consumers walk them in order); lane j keeps the ring interval of slot j, the
        // overlap test against all tiles in flight is one ballot.
        // The geometry of kG = 4 tiles is worked out at once, 8 lanes per tile (matrix load, corners, bounding box), one pass
        // ahead of the copies.  (Prefetching that pass's windows into L2 with cp.async.bulk.prefetch.tensor was measured: slower
        // in every configuration, up to 45 % for fp32 output -- 444 CTAs x 4..8 windows ahead is tens of MB that compete with the
        // output stream for L2.)
        constexpr int kG = 4, kGL = 32 / kG;   // tiles per pass, lanes per tile
        const int grp = lane / kGL, sub = lane % kGL;
        const int n_my = (int)blockIdx.x < g.total_tiles ? (g.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        struct TileInfo {
            int crop, frame, x0, y0, tw, th, sy_lo, bx_lo, k, pitch, nops, need, staged;
            float m[6];
        };
        auto geometry = [&](int it, TileInfo& ti) {   // it: index in this CTA's tile sequence (one value per 8-lane group)
            const bool valid = it < n_my;
            const int tile = (int)blockIdx.x + (valid ? it : 0) * (int)gridDim.x;
            ti.crop = tile / g.tiles_per_crop;
            const int t = tile - ti.crop * g.tiles_per_crop;
            const int tyi = (int)(((float)t + 0.5f) * g.inv_tiles_x), txi = t - tyi * g.tiles_x;
            ti.x0 = txi * g.tw; ti.y0 = tyi * g.th;
            ti.tw = min(g.tw, g.wo - ti.x0); ti.th = min(g.th, g.ho - ti.y0);
#pragma unroll
            for (int q = 0; q < 6; ++q) ti.m[q] = __ldg(minv + 6 * (size_t)ti.crop + q);
            ti.frame = frame_idx ? __ldg(frame_idx + ti.crop) : ti.crop;
            // lanes 0..3 of the group: the four corners, with the per-pixel expression (warp_affine_naive.cpp:23-24); 4..7 repeat them
            const int dx = (sub & 1) ? ti.x0 + ti.tw - 1 : ti.x0, dy = (sub & 2) ? ti.y0 + ti.th - 1 : ti.y0;
            const float fx = ti.m[0] * (float)dx + ti.m[1] * (float)dy + ti.m[2];
            const float fy = ti.m[3] * (float)dx + ti.m[4] * (float)dy + ti.m[5];
            const float kBig = 16777216.f;
            const bool finite = fx > -kBig && fx < kBig && fy > -kBig && fy < kBig;   // false for NaN too
            const int sx = finite ? (int)floorf(fx) : 0, sy = finite ? (int)floorf(fy) : 0;
            int sx_lo = sx, sx_hi = sx, sy_lo = sy, sy_hi = sy;
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                sx_lo = min(sx_lo, __shfl_xor_sync(0xffffffffu, sx_lo, o)); sx_hi = max(sx_hi, __shfl_xor_sync(0xffffffffu, sx_hi, o));
                sy_lo = min(sy_lo, __shfl_xor_sync(0xffffffffu, sy_lo, o)); sy_hi = max(sy_hi, __shfl_xor_sync(0xffffffffu, sy_hi, o));
            }
            const unsigned gmask = (kGL == 32 ? 0xffffffffu : ((1u << kGL) - 1u)) << (kGL * grp);
            const bool all_finite = (__ballot_sync(0xffffffffu, finite) & gmask) == gmask;
            // taps of the tile: columns sx_lo .. sx_hi + 1, rows sy_lo .. sy_hi + 1
            ti.sy_lo = sy_lo;
            ti.bx_lo = ((sx_lo * 3) >> 4) << 4;                              // floor to 16 bytes (arithmetic shift: negative too)
            const int bx_hi = (((sx_hi + 2) * 3 + 15) >> 4) << 4;
            const int bw = bx_hi - ti.bx_lo, bh = sy_hi + 2 - sy_lo;
            ti.k = bw <= 144 ? 0 : (bw - 144 + 63) >> 6;
            ti.pitch = ws_box_bytes(ti.k);
            ti.nops = (bh + kWsBoxRows - 1) / kWsBoxRows;
            ti.need = ti.nops * kWsBoxRows * ti.pitch;                       // a multiple of 256 bytes
            ti.staged = valid && all_finite && ti.k < kWsMaps && ti.nops <= 64 && ti.need <= kWsRingBytes;
        };
        auto tmap_of = [&](int k) { return reinterpret_cast<const uint8_t*>(&maps) + (size_t)k * sizeof(CUtensorMap); };
        int my_start = 0, my_end = 0;
        bool my_active = false;
        int head = 0, tail_it = 0;
        auto wait_release = [&]() {   // oldest tile in flight
            mbar_wait(&empty_bar[tail_it & (kWsSlots - 1)], (tail_it / kWsSlots) & 1);
            if (lane == (tail_it & (kWsSlots - 1))) my_active = false;
            ++tail_it;
        };
        const int full_magic = 65536 / g.tw + 1, full_step_y = NT / g.tw, full_step_x = NT - full_step_y * g.tw;
        TileInfo cur, nxt;
        geometry(grp, cur);
        for (int it0 = 0; it0 < n_my; it0 += kG) {
            geometry(it0 + kG + grp, nxt);   // next pass: its matrix loads are in flight while this pass is handed out
            for (int gg = 0; gg < kG && it0 + gg < n_my; ++gg) {   // warp-uniform: ring space is handed out in tile order
                const int it = it0 + gg, slot = it & (kWsSlots - 1);
                const int need = __shfl_sync(0xffffffffu, cur.need, kGL * gg);
                const bool staged = __shfl_sync(0xffffffffu, cur.staged, kGL * gg) != 0;
                if (it >= kWsSlots && tail_it <= it - kWsSlots) wait_release();   // the slot itself (tile it - kWsSlots)
                int pos = 0;
                if (staged) {
                    for (;;) {
                        pos = head + need > kWsRingBytes ? 0 : head;
                        if (!__ballot_sync(0xffffffffu, my_active && pos < my_end && my_start < pos + need)) break;
                        wait_release();
                    }
                    head = pos + need;
                    if (lane == slot) { my_start = pos; my_end = pos + need; my_active = true; }
                }
                if (grp == gg) {   // the tile's own 8 lanes
                    if (sub == 0) {
                        WsDesc& d = desc[slot];
                        d.crop = cur.crop; d.frame = cur.frame; d.x0 = cur.x0; d.y0 = cur.y0; d.tw = cur.tw; d.npx = cur.tw * cur.th;
                        d.sy_lo = cur.sy_lo; d.bx_lo = cur.bx_lo; d.pitch = cur.pitch; d.staged = cur.staged; d.pos = pos;
                        if (cur.tw == g.tw) { d.magic = full_magic; d.step_y = full_step_y; d.step_x = full_step_x; }
                        else { d.magic = 65536 / cur.tw + 1; d.step_y = NT / cur.tw; d.step_x = NT - d.step_y * cur.tw; }
#pragma unroll
                        for (int q = 0; q < 6; ++q) d.m[q] = cur.m[q];
                        if (cur.staged) mbar_expect_tx(&full_bar[slot], (uint32_t)cur.need);
                        else mbar_arrive(&full_bar[slot]);
                    }
                    if (cur.staged)
                        for (int j = sub; j < cur.nops; j += kGL)
                            tma_load_3d(stages_s + pos + j * kWsBoxRows * cur.pitch, tmap_of(cur.k), cur.bx_lo >> 2, cur.sy_lo + j * kWsBoxRows,
                                        cur.frame, &full_bar[slot]);
                }
                __syncwarp();
            }
            cur = nxt;
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    uint32_t* my_stage = ostage + warp * 96;
    const uint32_t lut_s = smem_u32(lut);
    const size_t crop_px = (size_t)g.wo * g.ho;
    const int row = g.w * 3;
    // tile rows are whole 32-bit words (u8) / whole 16-byte groups (fp32 HWC) of dst when w_out is a multiple of 4 (tw and x0 are)
    const bool vec_ok = (g.wo & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_) & (OUT == kWarpOutU8 ? 3 : 15)) == 0;
    const int p4 = (4 * lane) / 3, r4 = 4 * lane - 3 * p4;   // group `lane` of 4 output values starts inside pixel p4 of the warp's run
    // The 32-pixel runs of a tile are dealt to the warps starting at a different warp for every tile, so the leftover run of a
    // tile (784 pixels = 3 x 256 + 16) does not always land on the same warp.
    int it = 0, vw = warp;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it, vw = vw + 1 == NW ? 0 : vw + 1) {
        const int b = it & (kWsSlots - 1);
        mbar_wait(&full_bar[b], (it / kWsSlots) & 1);
        const WsDesc& d = desc[b];
        const float mr[6] = {d.m[0], d.m[1], d.m[2], d.m[3], d.m[4], d.m[5]};
        const int npx = d.npx, tw = d.tw, x_end = d.x0 + tw, step_x = d.step_x, step_y = d.step_y;
        const int staged = d.staged, pitch = d.pitch;
        // window address of source byte (sy, 3 sx): win + sy * pitch + 3 sx
        const uint32_t win_s = stages_s + d.pos - (uint32_t)(d.sy_lo * pitch + d.bx_lo);
        const uint8_t* img = frames + (size_t)d.frame * g.frame_bytes;
        const size_t out_base = (size_t)d.crop * crop_px * 3;
        const uint32_t safe_s = stages_s + d.pos;   // any valid window address, for lanes without a pixel
        int dx, dy;
        {
            const int i = vw * 32 + lane, ty = (i * d.magic) >> 16;
            dx = d.x0 + i - ty * tw; dy = d.y0 + ty;
        }
        // blended values << 22 of the pixel (px, py); branch-free so that two pixels of a lane can be in flight together
        auto blend = [&](auto staged_c, int px, int py, bool have, int (&H)[3]) {
            constexpr bool kStaged = decltype(staged_c)::value;
            const Taps t = warp_taps_fast(mr, px, py, g.w, g.h);
            const bool ok = have && t.in;
            uint32_t t0, t1, u0, u1;
            if (kStaged) {
                const uint32_t a = ok ? win_s + (uint32_t)(t.sy * pitch + t.sx * 3) : safe_s;
                const uint32_t wa = a & ~3u;
                const int r = (int)(a & 3u), sh = r * 8;    // window bases are 128-byte aligned: same phase as in the frame row
                const uint32_t w0 = lds_u32(wa), w1 = lds_u32(wa + 4), w2 = r == 3 ? lds_u32(wa + 8) : 0u;
                const uint32_t wb = wa + pitch;
                const uint32_t x0w = lds_u32(wb), x1w = lds_u32(wb + 4), x2w = r == 3 ? lds_u32(wb + 8) : 0u;
                t0 = __funnelshift_r(w0, w1, sh); t1 = __funnelshift_r(w1, w2, sh);
                u0 = __funnelshift_r(x0w, x1w, sh); u1 = __funnelshift_r(x1w, x2w, sh);
            } else {
                const unsigned ofs = ok ? (unsigned)t.ofs * 3u : 0u;
                linear_taps_u8c3(img, ofs, t0, t1);
                linear_taps_u8c3(img, ofs + (unsigned)row, u0, u1);
            }
            const uint32_t cx = (uint32_t)t.cx0 | ((uint32_t)t.cx1 << 16);
            int Ht[3], Hb[3];
            hsum_u8c3<kSigned>(t0, t1, cx, Ht);
            hsum_u8c3<kSigned>(u0, u1, cx, Hb);
#pragma unroll
            for (int k = 0; k < 3; ++k) H[k] = ok ? Ht[k] * t.cy0 + Hb[k] * t.cy1 : 0;   // warp_affine_naive.cpp:50-54 regrouped row-wise
        };
        auto emit = [&](const int (&H)[3], int px, int py, int i0) {
            const int n = min(32, npx - i0);   // valid pixels of this warp
            const int gpx = py * g.wo + px;    // pixel index inside the crop
            if (OUT == kWarpOutU8) {
                uint8_t* o = reinterpret_cast<uint8_t*>(dst_) + out_base;
                const int v0 = (H[0] >> 22) & 0xff, v1 = (H[1] >> 22) & 0xff, v2 = (H[2] >> 22) & 0xff;
                if (vec_ok) {
                    uint8_t* sb = reinterpret_cast<uint8_t*>(my_stage);
                    sb[3 * lane] = (uint8_t)v0; sb[3 * lane + 1] = (uint8_t)v1; sb[3 * lane + 2] = (uint8_t)v2;
                    __syncwarp();
                    const int goff = __shfl_sync(0xffffffffu, gpx, p4 & 31) * 3 + r4;
                    if (lane < 24 && 4 * lane < 3 * n) st_stream4(o + goff, my_stage[lane]);
                    __syncwarp();
                } else if (lane < n) {
                    o[(size_t)gpx * 3] = (uint8_t)v0; o[(size_t)gpx * 3 + 1] = (uint8_t)v1; o[(size_t)gpx * 3 + 2] = (uint8_t)v2;
                }
            } else {
                // table offsets in bytes: ((H >> 22) & 255) * 4
                const float r0 = lds_f32_at<0>(lut_s + ((H[0] >> 20) & 0x3fc)), r1 = lds_f32_at<1024>(lut_s + ((H[1] >> 20) & 0x3fc)),
                            r2 = lds_f32_at<2048>(lut_s + ((H[2] >> 20) & 0x3fc));
                float* out = reinterpret_cast<float*>(dst_) + out_base;
                if (OUT == kWarpOutF32CHW) {
                    if (lane < n) { st_stream4f(out + gpx, r0); st_stream4f(out + crop_px + gpx, r1); st_stream4f(out + 2 * crop_px + gpx, r2); }
                } else {
                    float* sf = reinterpret_cast<float*>(my_stage);
                    sf[3 * lane] = r0; sf[3 * lane + 1] = r1; sf[3 * lane + 2] = r2;
                    __syncwarp();
                    if (vec_ok) {
                        const int goff = __shfl_sync(0xffffffffu, gpx, p4 & 31) * 3 + r4;
                        if (lane < 24 && 4 * lane < 3 * n) st_stream16f(out + goff, *reinterpret_cast<const float4*>(sf + 4 * lane));
                    } else {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int e = 32 * j + lane, p = e / 3;
                            const int goff = __shfl_sync(0xffffffffu, gpx, p) * 3 + (e - 3 * p);
                            if (e < 3 * n) st_stream4f(out + goff, sf[e]);
                        }
                    }
                    __syncwarp();
                }
            }
        };
        auto advance = [&](int& px, int& py) { px += step_x; py += step_y; if (px >= x_end) { px -= tw; ++py; } };
        if (staged) {
            // two 32-pixel runs of the warp at a time: both pixels' window reads and blends are independent instruction streams
            int i0 = vw * 32;
            for (; i0 + NT < npx; i0 += 2 * NT) {   // warp-uniform
                int dxb = dx, dyb = dy;
                advance(dxb, dyb);
                int HA[3], HB[3];
                blend(std::true_type{}, dx, dy, true, HA);
                blend(std::true_type{}, dxb, dyb, i0 + NT + lane < npx, HB);
                emit(HA, dx, dy, i0);
                emit(HB, dxb, dyb, i0 + NT);
                dx = dxb; dy = dyb;
                advance(dx, dy);
            }
            if (i0 < npx) {
                int HA[3];
                blend(std::true_type{}, dx, dy, i0 + lane < npx, HA);
                emit(HA, dx, dy, i0);
            }
        } else {
            for (int i0 = vw * 32; i0 < npx; i0 += NT) {   // warp-uniform
                int HA[3];
                blend(std::false_type{}, dx, dy, i0 + lane < npx, HA);
                emit(HA, dx, dy, i0);
                advance(dx, dy);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[b]);   // this warp no longer reads the window / desc[b]
    }
}

}  // namespace vacv
