// a10: warp_affine (reference: src/cv/warp_affine.cpp:76-169, src/cv/warp_affine_naive.cpp:9-106) and the fused
// warp_affine_normalize (API src/cv/warp_affine_normalize.cpp:13-45; semantics = composition).
//
// Matrix inversion / rotation-matrix construction stay on the host (vacv_host.cpp) with the reference's mixed
// float/double arithmetic; the device evaluates fx = (m0*dx + m1*dy) + m2 in fp32 with no FMA contraction.
// One launch covers a whole batch of crops; each crop reads its own frame of a device-resident frame pool.
#include <algorithm>
#include <cstdlib>

#include "gather_u8c3.cuh"
#include "host_util.cuh"
#include "vacv_common.cuh"

namespace vacv {

struct WarpGeom {
    int w, h, c, wo, ho;
    size_t frame_elems;   // w*h*c
    int planar;           // 1: CHW (c planes, same geometry; warp_affine.cpp:152-168)
};

struct Taps { int ofs, sx, sy; int cx0, cx1, cy0, cy1; float fx, fy; bool in; };

// warp_affine_naive.cpp:23-44
__device__ __forceinline__ Taps warp_taps(const float* __restrict__ m, int dx, int dy, int w, int h) {
    Taps t;
    float fx = m[0] * (float)dx + m[1] * (float)dy + m[2];
    float fy = m[3] * (float)dx + m[4] * (float)dy + m[5];
    const int sy = (int)floorf(fy);
    fy -= (float)sy;
    const int sx = (int)floorf(fx);
    fx -= (float)sx;
    t.in = !(sy < 0 || sy >= h - 1 || sx < 0 || sx >= w - 1);
    t.cy0 = sat_short((1.f - fy) * 2048.f);
    t.cy1 = sat_short((float)(2048 - t.cy0));   // sums to exactly 2048 (:32)
    t.cx0 = sat_short((1.f - fx) * 2048.f);
    t.cx1 = sat_short((float)(2048 - t.cx0));
    t.fx = fx; t.fy = fy;
    t.sx = sx; t.sy = sy;
    t.ofs = sy * w + sx;
    return t;
}

// Same values with the saturating casts resolved: fy - floor(fy) lies in [0, 1], so (1 - fy) * 2048 lies in [0, 2048]:
// SATURATE_CAST_SHORT (macro.h:25-30) reduces to trunc(x + 0.5f) and never clamps, and the cast of the integer 2048 - c0
// returns it unchanged.
__device__ __forceinline__ Taps warp_taps_fast(const float (&m)[6], int dx, int dy, int w, int h) {
    Taps t;
    float fx = m[0] * (float)dx + m[1] * (float)dy + m[2];
    float fy = m[3] * (float)dx + m[4] * (float)dy + m[5];
    const float flx = floorf(fx), fly = floorf(fy);
    const int sy = (int)fly, sx = (int)flx;
    fy -= fly;
    fx -= flx;
    t.in = !(sy < 0 || sy >= h - 1 || sx < 0 || sx >= w - 1);
    t.cy0 = (int)((1.f - fy) * 2048.f + 0.5f);
    t.cy1 = 2048 - t.cy0;
    t.cx0 = (int)((1.f - fx) * 2048.f + 0.5f);
    t.cx1 = 2048 - t.cx0;
    t.fx = fx; t.fy = fy;
    t.sx = sx; t.sy = sy;
    t.ofs = sy * w + sx;
    return t;
}

template <bool kSigned>
__device__ __forceinline__ int warp_u8(const uint8_t* __restrict__ lt, int row, int c, const Taps& t) {
    const int p00 = pix<kSigned>(__ldg(lt)), p01 = pix<kSigned>(__ldg(lt + c));
    const int p10 = pix<kSigned>(__ldg(lt + row)), p11 = pix<kSigned>(__ldg(lt + row + c));
    return (p00 * t.cx0 * t.cy0 + p10 * t.cx0 * t.cy1 + p01 * t.cx1 * t.cy0 + p11 * t.cx1 * t.cy1) >> 22;
}

// grid = (ceil(wo/32), ceil(ho/8), crops)
template <typename T, bool kSigned>
__global__ void __launch_bounds__(256) warp_affine_kernel(const T* __restrict__ frames, const int* __restrict__ frame_idx,
                                                           const float* __restrict__ minv, T* __restrict__ dst,
                                                           WarpGeom g, int crop0) {
    __shared__ float m[6];
    const int crop = crop0 + blockIdx.z;
    if (threadIdx.y == 0 && threadIdx.x < 6) m[threadIdx.x] = __ldg(minv + 6 * (size_t)crop + threadIdx.x);
    __syncthreads();
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    if (dx >= g.wo || dy >= g.ho) return;
    const size_t f = frame_idx ? (size_t)__ldg(frame_idx + crop) : (size_t)crop;
    const T* img = frames + f * g.frame_elems;
    T* out = dst + (size_t)crop * g.wo * g.ho * g.c;
    const Taps t = warp_taps(m, dx, dy, g.w, g.h);
    const int step = g.planar ? 1 : g.c;                          // element stride between x-neighbours
    const size_t plane_in = g.planar ? (size_t)g.w * g.h : 1, plane_out = g.planar ? (size_t)g.wo * g.ho : 1;
    const T* lt = img + (size_t)t.ofs * step;
    T* o = out + ((size_t)dy * g.wo + dx) * step;
    const int row = g.w * step;
    for (int k = 0; k < g.c; ++k) {
        T v;
        if (!t.in) v = (T)0;   // reference leaves these untouched; defined as 0 (App. C-5)
        else if (sizeof(T) == 1) v = (T)warp_u8<kSigned>((const uint8_t*)lt + k * plane_in, row, step, t);
        else {
            const float* p = (const float*)lt + k * plane_in;
            const float cx0 = 1.f - t.fx, cx1 = t.fx, cy0 = 1.f - t.fy, cy1 = t.fy;
            v = (T)(__ldg(p) * cx0 * cy0 + __ldg(p + row) * cx0 * cy1 + __ldg(p + step) * cx1 * cy0 +
                    __ldg(p + row + step) * cx1 * cy1);
        }
        o[k * plane_out] = v;
    }
}

// Fused: warp u8 HWC -> (u8 value) -> table[c][256] = normalised fp32.  One CTA per crop (grid.x) x row band
// (grid.y); the exact 256 x c table is built once per CTA.  out_layout HWC (reference) or CHW planes.
template <int C>
__global__ void __launch_bounds__(256) warp_affine_normalize_kernel(const uint8_t* __restrict__ frames,
                                                                     const int* __restrict__ frame_idx,
                                                                     const float* __restrict__ minv,
                                                                     float* __restrict__ dst, WarpGeom g,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ stddev,
                                                                     int out_layout, int rows_per_cta) {
    __shared__ float lut[C][256];
    __shared__ float m[6];
    const int crop = blockIdx.x;
    if (threadIdx.x < 6) m[threadIdx.x] = __ldg(minv + 6 * (size_t)crop + threadIdx.x);
    for (int t = threadIdx.x; t < 256 * C; t += blockDim.x)
        lut[t >> 8][t & 255] = normalize_one((float)(t & 255), __ldg(mean + (t >> 8)), (double)__ldg(stddev + (t >> 8)) + 1e-6);
    __syncthreads();
    const size_t f = frame_idx ? (size_t)__ldg(frame_idx + crop) : (size_t)crop;
    const uint8_t* img = frames + f * g.frame_elems;
    float* out = dst + (size_t)crop * g.wo * g.ho * C;
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(y0 + rows_per_cta, g.ho);
    const int row = g.w * C;
    const size_t plane = (size_t)g.wo * g.ho;
    for (int i = y0 * g.wo + threadIdx.x; i < y1 * g.wo; i += blockDim.x) {
        const int dy = i / g.wo, dx = i - dy * g.wo;
        const Taps t = warp_taps(m, dx, dy, g.w, g.h);
        const uint8_t* lt = img + (size_t)t.ofs * C;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int v = t.in ? (warp_u8<false>(lt + k, row, C, t) & 0xff) : 0;
            const float r = lut[k][v];
            if (out_layout == VACV_NHWC) out[(size_t)i * C + k] = r;
            else st_stream4f(out + k * plane + i, r);
        }
    }
}


// ----------------------------------------------------------------------------------------------------
// u8, 3 interleaved channels (the C3 shape), word-granular taps and coalesced output.
// ncu on the byte-granular kernels above: L1/LSU wavefronts at 80 % of peak (12 byte loads per pixel), i.e. LSU-bound.
// Here the 6 contiguous bytes of a tap row (2 pixels x BGR) come from 2-3 aligned 32-bit words + funnel shifts, and
// every warp re-chunks its 32 output pixels through a private shared-memory line so that global stores are
// lane-contiguous (96 B of u8, or 3 x 128 B of fp32) instead of 3 strided partial-sector stores per lane.
enum { kWarpOutU8 = 0, kWarpOutF32HWC = 1, kWarpOutF32CHW = 2 };

// grid = (crops, bands); each CTA walks pixels [y0*wo, y1*wo) of its crop in flat order (a warp = 32 consecutive pixels)
template <int OUT, bool kSigned>
__global__ void __launch_bounds__(256) warp_affine_u8c3_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ frame_idx,
                                                                const float* __restrict__ minv, void* __restrict__ dst_, WarpGeom g,
                                                                const float* __restrict__ mean, const float* __restrict__ stddev,
                                                                int rows_per_cta, int crop0) {
    __shared__ float lut[OUT == kWarpOutU8 ? 1 : 3 * 256];
    __shared__ float m[6];
    __shared__ __align__(16) uint32_t stage[8][OUT == kWarpOutU8 ? 24 : 96];   // per warp: 32 px x 3 ch
    const int crop = crop0 + blockIdx.x;
    if (threadIdx.x < 6) m[threadIdx.x] = __ldg(minv + 6 * (size_t)crop + threadIdx.x);
    if (OUT != kWarpOutU8)
        for (int t = threadIdx.x; t < 768; t += blockDim.x)
            lut[t] = normalize_one((float)(t & 255), __ldg(mean + (t >> 8)), (double)__ldg(stddev + (t >> 8)) + 1e-6);
    __syncthreads();
    const size_t f = frame_idx ? (size_t)__ldg(frame_idx + crop) : (size_t)crop;
    const uint8_t* img = frames + f * g.frame_elems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i_begin = blockIdx.y * rows_per_cta * g.wo, i_end = min((blockIdx.y + 1) * rows_per_cta, g.ho) * g.wo;
    const size_t crop_px = (size_t)g.wo * g.ho;
    const int row = g.w * 3;
    const float mr[6] = {m[0], m[1], m[2], m[3], m[4], m[5]};
    // fp32 HWC: a warp's 96 values are one contiguous 384-byte run of dst -> 24 lanes x 16 bytes when everything is 16-byte aligned
    const bool vec_ok = OUT == kWarpOutF32HWC && (crop_px & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_) & 15) == 0;
    // (dx, dy) of this lane's pixel, advanced by 256 pixels per iteration without a division
    const int step_y = 256 / g.wo, step_x = 256 - step_y * g.wo;
    int dy = (i_begin + warp * 32 + lane) / g.wo, dx = (i_begin + warp * 32 + lane) - dy * g.wo;
    for (int i0 = i_begin + warp * 32; i0 < i_end; i0 += 256, dx += step_x, dy += step_y) {   // warp-uniform loop
        const int i = i0 + lane;
        if (dx >= g.wo) { dx -= g.wo; ++dy; }
        int v[3] = {0, 0, 0};
        if (i < i_end) {
            const Taps t = warp_taps_fast(mr, dx, dy, g.w, g.h);
            if (t.in) {
                uint32_t t0, t1, u0, u1;
                linear_taps_u8c3(img, (unsigned)t.ofs * 3u, t0, t1);
                linear_taps_u8c3(img, (unsigned)t.ofs * 3u + (unsigned)row, u0, u1);
                const uint32_t cx = (uint32_t)t.cx0 | ((uint32_t)t.cx1 << 16);
                int Ht[3], Hb[3];
                hsum_u8c3<kSigned>(t0, t1, cx, Ht);   // p00*cx0 + p01*cx1 (two PRMT + three IDP.2A)
                hsum_u8c3<kSigned>(u0, u1, cx, Hb);   // p10*cx0 + p11*cx1
#pragma unroll
                for (int k = 0; k < 3; ++k)   // warp_affine_naive.cpp:50-54: the same integer regrouped row-wise
                    v[k] = ((Ht[k] * t.cy0 + Hb[k] * t.cy1) >> 22) & 0xff;
            }
        }
        const int n = min(32, i_end - i0);   // valid pixels of this warp
        if (OUT == kWarpOutU8) {
            uint8_t* sb = reinterpret_cast<uint8_t*>(stage[warp]);
            sb[3 * lane] = (uint8_t)v[0]; sb[3 * lane + 1] = (uint8_t)v[1]; sb[3 * lane + 2] = (uint8_t)v[2];
            __syncwarp();
            uint8_t* o = reinterpret_cast<uint8_t*>(dst_) + ((size_t)crop * crop_px + i0) * 3;   // 96 B per warp, 4-byte aligned when dst is
            if ((reinterpret_cast<uintptr_t>(o) & 3) == 0 && n == 32) { if (lane < 24) st_stream4(o + 4 * lane, stage[warp][lane]); }
            else for (int b = lane; b < 3 * n; b += 32) o[b] = sb[b];
            __syncwarp();
        } else {
            const float r0 = lut[v[0]], r1 = lut[256 + v[1]], r2 = lut[512 + v[2]];
            float* out = reinterpret_cast<float*>(dst_) + (size_t)crop * crop_px * 3;
            if (OUT == kWarpOutF32CHW) {
                if (i < i_end) { st_stream4f(out + i, r0); st_stream4f(out + crop_px + i, r1); st_stream4f(out + 2 * crop_px + i, r2); }
            } else {
                float* sf = reinterpret_cast<float*>(stage[warp]);
                sf[3 * lane] = r0; sf[3 * lane + 1] = r1; sf[3 * lane + 2] = r2;
                __syncwarp();
                float* o = out + (size_t)i0 * 3;
                if (vec_ok && (n & 3) == 0) {
                    if (4 * lane < 3 * n) st_stream16f(o + 4 * lane, *reinterpret_cast<const float4*>(sf + 4 * lane));
                } else {
#pragma unroll
                    for (int j = 0; j < 3; ++j) if (32 * j + lane < 3 * n) st_stream4f(o + 32 * j + lane, sf[32 * j + lane]);
                }
                __syncwarp();
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// u8, one channel (grey frames / one plane of a CHW frame), same structure: the two tap bytes of a row come from one aligned
// 32-bit word (two when they straddle it), one IDP.2A forms L*cx0 + R*cx1; OUT = u8 (a warp's 32 bytes leave as 8 words) or
// normalised fp32 (lane-contiguous 128-byte store, HWC == CHW for one channel).
template <int OUT, bool kSigned>
__global__ void __launch_bounds__(256) warp_affine_u8c1_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ frame_idx,
                                                                const float* __restrict__ minv, void* __restrict__ dst_, WarpGeom g,
                                                                const float* __restrict__ mean, const float* __restrict__ stddev,
                                                                int rows_per_cta, int crop0, size_t plane_elems, int planes) {
    __shared__ float lut[OUT == kWarpOutU8 ? 1 : 256];
    __shared__ float m[6];
    __shared__ __align__(16) uint32_t stage[8][8];   // per warp: 32 px
    // blockIdx.x = crop * planes + plane (planes > 1: the c planes of a CHW frame share the crop's matrix)
    const int unit = crop0 + blockIdx.x, crop = unit / planes, plane = unit - crop * planes;
    if (threadIdx.x < 6) m[threadIdx.x] = __ldg(minv + 6 * (size_t)crop + threadIdx.x);
    if (OUT != kWarpOutU8)
        for (int t = threadIdx.x; t < 256; t += blockDim.x) lut[t] = normalize_one((float)t, __ldg(mean), (double)__ldg(stddev) + 1e-6);
    __syncthreads();
    const size_t f = frame_idx ? (size_t)__ldg(frame_idx + crop) : (size_t)crop;
    const uint8_t* img = frames + f * g.frame_elems + (size_t)plane * plane_elems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i_begin = blockIdx.y * rows_per_cta * g.wo, i_end = min((blockIdx.y + 1) * rows_per_cta, g.ho) * g.wo;
    const size_t crop_px = (size_t)g.wo * g.ho;
    const float mr[6] = {m[0], m[1], m[2], m[3], m[4], m[5]};
    const int step_y = 256 / g.wo, step_x = 256 - step_y * g.wo;
    int dy = (i_begin + warp * 32 + lane) / g.wo, dx = (i_begin + warp * 32 + lane) - dy * g.wo;
    auto taps = [&](unsigned a) -> uint32_t {   // bytes a, a+1 of the plane in the low half-word
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (a & ~3u));
        const int r = (int)(a & 3u);
        const uint32_t w0 = __ldg(wp), w1 = r == 3 ? __ldg(wp + 1) : 0u;
        return __funnelshift_r(w0, w1, r * 8);
    };
    for (int i0 = i_begin + warp * 32; i0 < i_end; i0 += 256, dx += step_x, dy += step_y) {   // warp-uniform loop
        const int i = i0 + lane;
        if (dx >= g.wo) { dx -= g.wo; ++dy; }
        int v = 0;
        if (i < i_end) {
            const Taps t = warp_taps_fast(mr, dx, dy, g.w, g.h);
            if (t.in) {
                const uint32_t p0 = taps((unsigned)t.ofs), p1 = taps((unsigned)t.ofs + (unsigned)g.w);
                const uint32_t cx = (uint32_t)t.cx0 | ((uint32_t)t.cx1 << 16);
                const int Ht = kSigned ? __dp2a_lo((int)cx, (int)p0, 0) : (int)__dp2a_lo(cx, p0, 0u);
                const int Hb = kSigned ? __dp2a_lo((int)cx, (int)p1, 0) : (int)__dp2a_lo(cx, p1, 0u);
                v = ((Ht * t.cy0 + Hb * t.cy1) >> 22) & 0xff;   // warp_affine_naive.cpp:50-54 regrouped row-wise
            }
        }
        const int n = min(32, i_end - i0);   // valid pixels of this warp
        if (OUT == kWarpOutU8) {
            uint8_t* sb = reinterpret_cast<uint8_t*>(stage[warp]);
            sb[lane] = (uint8_t)v;
            __syncwarp();
            uint8_t* o = reinterpret_cast<uint8_t*>(dst_) + (size_t)unit * crop_px + i0;
            if ((reinterpret_cast<uintptr_t>(o) & 3) == 0 && n == 32) { if (lane < 8) st_stream4(o + 4 * lane, stage[warp][lane]); }
            else if (lane < n) o[lane] = sb[lane];
            __syncwarp();
        } else if (i < i_end) {
            st_stream4f(reinterpret_cast<float*>(dst_) + (size_t)unit * crop_px + i, lut[v]);
        }
    }
}

}  // namespace vacv

#include "tma_host.cuh"
#include "warp_pack_u8c3.cuh"     // column-owning form of the u8 BGR gather kernel (64-bit tap loads, shuffle-packed stores)
#include "warp_staged_u8c3.cuh"   // TMA-staged variant of the 3-channel kernel (uses Taps / warp_taps_fast / kWarpOut*)

using namespace vacv;

// ---- host side of the staged kernel: tensor maps over the frame pool (one per box width), cached per host thread
namespace {
struct StagedPlan {
    const void* frames = nullptr;
    int n_frames = 0, w = 0, h = 0, device = -1;
    bool ok = false;
    WarpStagedMaps maps;
};

// true when the frame pool can be described by the tensor maps (dense rows that are multiples of 16 bytes)
bool staged_plan(const void* frames, int n_frames, int w, int h, const WarpStagedMaps** maps) {
    static thread_local PlanCache<StagedPlan, 8> cache;   // a tensor map binds the pool's address: key = (device, pool, shape)
    const int dev = current_device();
    StagedPlan* pp = cache.find([&](const StagedPlan& p) { return p.device == dev && p.frames == frames && p.n_frames == n_frames && p.w == w && p.h == h; });
    if (!pp) {
        pp = cache.claim();
        StagedPlan& plan = *pp;
        plan.device = dev; plan.frames = frames; plan.n_frames = n_frames; plan.w = w; plan.h = h; plan.ok = false;
        const bool shape_ok = (w % 16) == 0 && w * 3 >= ws_box_bytes(kWsMaps - 1) && h >= kWsBoxRows &&
                              ((uintptr_t)frames % 16) == 0 && (size_t)w * h * 3 < 0xfffffff0ull;
        if (shape_ok) {
            plan.ok = true;
            for (int k = 0; k < kWsMaps && plan.ok; ++k)
                plan.ok = encode_map_3d(&plan.maps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT32, frames, (cuuint64_t)w * 3 / 4, h, n_frames,
                                        (cuuint64_t)w * 3, (cuuint64_t)w * 3 * h, ws_box_bytes(k) / 4, kWsBoxRows, 1);
        }
        cache.commit();
    }
    *maps = &pp->maps;
    return pp->ok;
}

// tile counters of the staged kernel are 32-bit
bool staged_fits(int n_crops, int w_out, int h_out) {
    return (long long)n_crops * ceil_div(w_out, 4) * ceil_div(h_out, 4) < 0x40000000LL;
}

template <int OUT, bool kSigned, int NW>
int launch_staged_nw(const WarpStagedMaps& maps, const uint8_t* frames, const int* frame_idx, const float* minv, void* dst, int w, int h,
                     int n_crops, int w_out, int h_out, const float* mean, const float* stddev, cudaStream_t s) {
    constexpr int kSmem = ws_smem_bytes<NW>();
    static thread_local unsigned long long attr_devices = 0;   // the opt-in is per device: one bit per device already done
    const int device = current_device();
    if (device >= 64 || !((attr_devices >> device) & 1)) {
        if (cudaFuncSetAttribute(warp_affine_u8c3_staged_kernel<OUT, kSigned, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
            return check_launch("warp_affine (staged kernel attribute)");
        if (device < 64) attr_devices |= 1ull << device;
    }
    WarpStagedGeom g;
    g.w = w; g.h = h; g.wo = w_out; g.ho = h_out;
    const int nx = ceil_div(w_out, 32), ny = ceil_div(h_out, 28);
    g.tw = (ceil_div(w_out, nx) + 3) & ~3;                 // multiple of 4 pixels: u8 tile rows are whole 32-bit words
    g.th = ceil_div(h_out, ny);
    g.tiles_x = ceil_div(w_out, g.tw);
    g.tiles_per_crop = g.tiles_x * ceil_div(h_out, g.th);
    g.total_tiles = g.tiles_per_crop * n_crops;           // < 2^31: checked by staged_fits()
    g.inv_tiles_x = 1.f / (float)g.tiles_x;
    g.frame_bytes = (size_t)w * h * 3;
    const int grid = std::min(g.total_tiles, kWsCtasPerSm * sm_count(device));
    warp_affine_u8c3_staged_kernel<OUT, kSigned, NW><<<grid, 32 * NW + 32, kSmem, s>>>(maps, frames, frame_idx, minv, dst, g, mean, stddev);
    return check_launch("warp_affine (staged)");
}
template <int OUT, bool kSigned>
int launch_staged(const WarpStagedMaps& maps, const uint8_t* frames, const int* frame_idx, const float* minv, void* dst, int w, int h,
                  int n_crops, int w_out, int h_out, const float* mean, const float* stddev, cudaStream_t s) {
    // 8 consumer warps: measured faster than 10 or 12 (config 3: 0.318 vs 0.332 / 0.331 ms; the per-tile prologue is paid per warp)
    return launch_staged_nw<OUT, kSigned, 8>(maps, frames, frame_idx, minv, dst, w, h, n_crops, w_out, h_out, mean, stddev, s);
}
}  // namespace

static int check_warp_args(const void* frames, const float* minv, const void* dst, int n_frames, int w, int h, int c,
                           int n_crops, int w_out, int h_out) {
    VACV_REQUIRE(frames && minv && dst, "warp_affine: null pointer");
    VACV_REQUIRE(n_frames > 0 && n_crops > 0 && w > 1 && h > 1 && c > 0 && w_out > 0 && h_out > 0, "warp_affine: bad size");
    return VACV_OK;
}

extern "C" int vacv_cuda_warp_affine(const void* frames, int n_frames, int w, int h, int c, int dtype, int layout,
                                     const int* frame_idx, const float* minv, int n_crops,
                                     void* dst, int w_out, int h_out, int flags, void* stream) {
    if (int e = check_warp_args(frames, minv, dst, n_frames, w, h, c, n_crops, w_out, h_out)) return e;
    if (dtype != VACV_INT8 && dtype != VACV_FP32) return set_error(VACV_ERR_UNSUPPORTED, "warp_affine: dtype %d", dtype);
    VACV_REQUIRE(frame_idx || n_crops <= n_frames, "warp_affine: frame_idx == NULL needs n_crops <= n_frames");
    WarpGeom g;
    g.w = w; g.h = h; g.c = c; g.wo = w_out; g.ho = h_out; g.frame_elems = (size_t)w * h * c; g.planar = layout == VACV_NCHW;
    cudaStream_t s = as_stream(stream);
    const bool words_ok = (((size_t)w * h * 3) % 4) == 0 && ((uintptr_t)frames % 4) == 0 && (size_t)w * h * 3 < 0xfffffff0ull;
    // u8 output: the direct gather kernel is the faster one (config-3 shape: 0.275 vs 0.30 ms); the TMA-staged kernel runs on request
    if (dtype == VACV_INT8 && c == 3 && layout == VACV_NHWC && words_ok && (flags & VACV_FLAG_TILED)) {
        const WarpStagedMaps* maps;
        if (staged_fits(n_crops, w_out, h_out) && staged_plan(frames, n_frames, w, h, &maps)) {
            if (flags & VACV_FLAG_SIGNED_CHAR)
                return launch_staged<kWarpOutU8, true>(*maps, (const uint8_t*)frames, frame_idx, minv, dst, w, h, n_crops, w_out, h_out, nullptr, nullptr, s);
            return launch_staged<kWarpOutU8, false>(*maps, (const uint8_t*)frames, frame_idx, minv, dst, w, h, n_crops, w_out, h_out, nullptr, nullptr, s);
        }
    }
    if (dtype == VACV_INT8 && c == 3 && layout == VACV_NHWC && words_ok && knob(kKnobWarpV) != 1 && (w % 8) == 0 && ((uintptr_t)frames % 8) == 0 &&
        (w_out % 4) == 0 && w_out <= 256 && ((uintptr_t)dst % 4) == 0 && n_crops <= 0x7fffffff / 6) {
        // pack kernel: a CTA is `rows` whole output rows wide; taken when its last warp is not mostly idle
        const int rows = 256 / w_out, threads = (rows * w_out + 31) & ~31;
        if (rows * w_out * 100 >= threads * 85) {
            const int bands = ceil_div(h_out, (4096 + w_out - 1) / w_out);
            const int rows_per_cta = (ceil_div(h_out, bands) + rows - 1) / rows * rows;
            dim3 grid(n_crops, ceil_div(h_out, rows_per_cta));
            const bool sc = (flags & VACV_FLAG_SIGNED_CHAR) != 0;
            // one pixel per thread and pass; two (U = 2, tap loads of both in flight together) measured the same to 1.5 % slower -- the
            // kernel is bound by L1 wavefronts, not by load latency -- and stay as the A/B variant VACV_WARP_V=2
            if (knob(kKnobWarpV) == 3) {   // A/B: 32-bit tap loads
                auto k3 = sc ? warp_affine_u8c3_pack_kernel<true, 1, false> : warp_affine_u8c3_pack_kernel<false, 1, false>;
                k3<<<grid, threads, 0, s>>>((const uint8_t*)frames, frame_idx, minv, (uint8_t*)dst, w, h, w_out, h_out, g.frame_elems, rows, rows_per_cta, 0);
                return check_launch("warp_affine");
            }
            auto kern = knob(kKnobWarpV) == 2 ? (sc ? warp_affine_u8c3_pack_kernel<true, 2> : warp_affine_u8c3_pack_kernel<false, 2>)
                                              : (sc ? warp_affine_u8c3_pack_kernel<true, 1> : warp_affine_u8c3_pack_kernel<false, 1>);
            kern<<<grid, threads, 0, s>>>((const uint8_t*)frames, frame_idx, minv, (uint8_t*)dst, w, h, w_out, h_out, g.frame_elems, rows, rows_per_cta, 0);
            return check_launch("warp_affine");
        }
    }
    if (dtype == VACV_INT8 && c == 3 && layout == VACV_NHWC && words_ok) {
        const int rows_per_cta = max(1, min(h_out, (4096 + w_out - 1) / w_out));
        dim3 grid(n_crops, ceil_div(h_out, rows_per_cta));
        if (flags & VACV_FLAG_SIGNED_CHAR)
            warp_affine_u8c3_kernel<kWarpOutU8, true><<<grid, 256, 0, s>>>((const uint8_t*)frames, frame_idx, minv, dst, g, nullptr, nullptr, rows_per_cta, 0);
        else
            warp_affine_u8c3_kernel<kWarpOutU8, false><<<grid, 256, 0, s>>>((const uint8_t*)frames, frame_idx, minv, dst, g, nullptr, nullptr, rows_per_cta, 0);
        return check_launch("warp_affine");
    }
    const bool planes_ok = (((size_t)w * h) % 4) == 0 && ((uintptr_t)frames % 4) == 0 && (size_t)w * h < 0xfffffff0ull;
    if (dtype == VACV_INT8 && (c == 1 || layout == VACV_NCHW) && planes_ok && (long long)n_crops * c <= 0x7fffffffLL) {
        // grey frames, or a CHW frame = c planes warped with the same matrix (warp_affine.cpp:150-166)
        const int rows_per_cta = max(1, min(h_out, (4096 + w_out - 1) / w_out));
        const int planes = c;
        const long long units = (long long)n_crops * planes;
        for (long long u0 = 0; u0 < units; u0 += 0x7ffffff0LL / planes * planes) {
            const int nu = (int)min(units - u0, 0x7ffffff0LL / planes * planes);
            dim3 grid(nu, ceil_div(h_out, rows_per_cta));
            if (flags & VACV_FLAG_SIGNED_CHAR)
                warp_affine_u8c1_kernel<kWarpOutU8, true><<<grid, 256, 0, s>>>((const uint8_t*)frames, frame_idx, minv, dst, g, nullptr, nullptr, rows_per_cta, (int)u0, (size_t)w * h, planes);
            else
                warp_affine_u8c1_kernel<kWarpOutU8, false><<<grid, 256, 0, s>>>((const uint8_t*)frames, frame_idx, minv, dst, g, nullptr, nullptr, rows_per_cta, (int)u0, (size_t)w * h, planes);
        }
        return check_launch("warp_affine");
    }
    dim3 block(32, 8);
    for (int c0 = 0; c0 < n_crops; c0 += 65535) {
        dim3 grid(ceil_div(w_out, 32), ceil_div(h_out, 8), min(n_crops - c0, 65535));
        if (dtype == VACV_FP32)
            warp_affine_kernel<float, false><<<grid, block, 0, s>>>((const float*)frames, frame_idx, minv, (float*)dst, g, c0);
        else if (flags & VACV_FLAG_SIGNED_CHAR)
            warp_affine_kernel<uint8_t, true><<<grid, block, 0, s>>>((const uint8_t*)frames, frame_idx, minv, (uint8_t*)dst, g, c0);
        else
            warp_affine_kernel<uint8_t, false><<<grid, block, 0, s>>>((const uint8_t*)frames, frame_idx, minv, (uint8_t*)dst, g, c0);
    }
    return check_launch("warp_affine");
}

extern "C" int vacv_cuda_warp_affine_normalize(const uint8_t* frames, int n_frames, int w, int h, int c,
                                               const int* frame_idx, const float* minv, int n_crops,
                                               float* dst, int w_out, int h_out, const float* mean, const float* stddev,
                                               int out_layout, void* stream) {
    if (int e = check_warp_args(frames, minv, dst, n_frames, w, h, c, n_crops, w_out, h_out)) return e;
    VACV_REQUIRE(mean && stddev, "warp_affine_normalize: null statistics");
    VACV_REQUIRE(frame_idx || n_crops <= n_frames, "warp_affine_normalize: frame_idx == NULL needs n_crops <= n_frames");
    if (c != 1 && c != 3) return set_error(VACV_ERR_UNSUPPORTED, "warp_affine_normalize: c must be 1 or 3 (got %d)", c);
    WarpGeom g;
    g.w = w; g.h = h; g.c = c; g.wo = w_out; g.ho = h_out; g.frame_elems = (size_t)w * h * c; g.planar = 0;
    // >= ~8k output pixels per CTA so the 256*c-entry table build is amortised
    const int rows_per_cta = max(1, min(h_out, (8192 + w_out - 1) / w_out));
    dim3 grid(n_crops, ceil_div(h_out, rows_per_cta));
    cudaStream_t s = as_stream(stream);
    // fp32 output: the TMA-staged kernel is the default (config 3: 0.335 vs 0.373 ms); VACV_WARP_GATHER=1 forces the direct gather kernel
    const WarpStagedMaps* maps;
    if (c == 3 && !knob(kKnobWarpGather) && staged_fits(n_crops, w_out, h_out) && staged_plan(frames, n_frames, w, h, &maps)) {
        if (out_layout == VACV_NHWC) return launch_staged<kWarpOutF32HWC, false>(*maps, frames, frame_idx, minv, dst, w, h, n_crops, w_out, h_out, mean, stddev, s);
        return launch_staged<kWarpOutF32CHW, false>(*maps, frames, frame_idx, minv, dst, w, h, n_crops, w_out, h_out, mean, stddev, s);
    }
    if (c == 3 && (((size_t)w * h * 3) % 4) == 0 && ((uintptr_t)frames % 4) == 0 && (size_t)w * h * 3 < 0xfffffff0ull) {
        if (out_layout == VACV_NHWC) warp_affine_u8c3_kernel<kWarpOutF32HWC, false><<<grid, 256, 0, s>>>(frames, frame_idx, minv, dst, g, mean, stddev, rows_per_cta, 0);
        else warp_affine_u8c3_kernel<kWarpOutF32CHW, false><<<grid, 256, 0, s>>>(frames, frame_idx, minv, dst, g, mean, stddev, rows_per_cta, 0);
        return check_launch("warp_affine_normalize");
    }
    if (c == 1 && (((size_t)w * h) % 4) == 0 && ((uintptr_t)frames % 4) == 0 && (size_t)w * h < 0xfffffff0ull) {
        warp_affine_u8c1_kernel<kWarpOutF32HWC, false><<<grid, 256, 0, s>>>(frames, frame_idx, minv, dst, g, mean, stddev, rows_per_cta, 0, (size_t)w * h, 1);
        return check_launch("warp_affine_normalize");
    }
    if (c == 3) warp_affine_normalize_kernel<3><<<grid, 256, 0, s>>>(frames, frame_idx, minv, dst, g, mean, stddev, out_layout, rows_per_cta);
    else warp_affine_normalize_kernel<1><<<grid, 256, 0, s>>>(frames, frame_idx, minv, dst, g, mean, stddev, out_layout, rows_per_cta);
    return check_launch("warp_affine_normalize");
}
