// Host side of the periodic u8 bilinear walker (resize_linear3_period.cuh): eligibility checks with the device's own arithmetic,
// grid shape, launch.  Called from vacv_cuda_resize (resize.cu) ahead of the persistent bilinear pipeline.
#include <algorithm>
#include <climits>
#include <cmath>

#include "resize_linear3_period.cuh"
#include "host_util.cuh"
#include "vacv_common.cuh"

using namespace vacv;

namespace {

// host twin of linear_coord's source index without the clamps (identical IEEE arithmetic on x86-64: no FMA, same rounding)
inline int host_linear_floor(int d, double scale) { return (int)floorf((float)(((double)d + 0.5) * scale - 0.5)); }

struct LinPeriodPlan { int w, wo, h, ho, P, Q, KP; bool ok, down; };

// ok: every column's left tap is window pixel tap0(c) of its thread and no column is edge-clamped; down: output rows end on strictly
// increasing source rows (the kernel's kDown variant emits at most one row per walk step)
template <int P, int Q, int KP>
const LinPeriodPlan* lin_period_plan(int w, int h, int wo, int ho, double scale_x, double scale_y) {
    static thread_local PlanCache<LinPeriodPlan, 8> cache;
    if (const LinPeriodPlan* p = cache.find([&](const LinPeriodPlan& q) { return q.w == w && q.wo == wo && q.h == h && q.ho == ho && q.P == P && q.Q == Q && q.KP == KP; }))
        return p;
    constexpr int NCOL = Q * KP;
    bool ok = true;
    for (int dx = 0; dx < wo && ok; ++dx) {
        const int sx = host_linear_floor(dx, scale_x);
        ok = sx >= 0 && sx < w - 1 && sx == P * KP * (dx / NCOL) + pd::tap0(P, Q, dx % NCOL);
    }
    bool down = true;
    int prev = INT_MIN;
    for (int d = 0; d < ho && down; ++d) {
        const int sy = std::min(std::max(host_linear_floor(d, scale_y), 0), h - 2);
        down = sy > prev;
        prev = sy;
    }
    LinPeriodPlan* p = cache.claim();
    p->w = w; p->wo = wo; p->h = h; p->ho = ho; p->P = P; p->Q = Q; p->KP = KP; p->ok = ok; p->down = down;
    cache.commit();
    return p;
}

template <int P, int Q, int KP, bool kSigned, int C = 3>
int launch_linear3_period(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    using S = LinPeriodShape<P, Q, KP, C>;
    if ((long long)w * Q != (long long)wo * P || wo % S::NCOL != 0) return 0;
    // bulk copies and the staged flush move aligned 16-byte chunks
    if (((size_t)wo * C) % 16 != 0 || ((size_t)w * C) % 16 != 0 || ((uintptr_t)dst % 16) != 0 || ((uintptr_t)src % 16) != 0) return 0;
    // the walk filters every source row between a segment's first and last tap row: up to a vertical ratio of 3 (where one row in three
    // is copied without being used) that is still faster than the gather kernel (1080p -> 768x432: see profiles/r2f_bench_ops.txt)
    if (h > 3 * ho || h < 2 || ho < 1) return 0;
    LinPeriodGeom g;
    g.w = w; g.h = h; g.wo = wo; g.ho = ho;
    g.src_image = (size_t)w * h * C; g.dst_image = (size_t)wo * ho * C;
    g.scale_x = (double)((float)w / (float)wo); g.scale_y = (double)((float)h / (float)ho);   // resize_naive.cpp:14-15: fp32 scales
    const LinPeriodPlan* plan = lin_period_plan<P, Q, KP>(w, h, wo, ho, g.scale_x, g.scale_y);
    if (!plan->ok) return 0;
    g.warp_strips = (wo + 32 * S::NCOL - 1) / (32 * S::NCOL);
    int warps = 4, best_pad = INT_MAX;
    for (int wv = 4; wv >= 2; --wv) {
        const int pad = (g.warp_strips + wv - 1) / wv * wv - g.warp_strips;
        if (pad < best_pad) { best_pad = pad; warps = wv; }
    }
    g.cta_strips = (g.warp_strips + warps - 1) / warps;
    const size_t per_warp = (size_t)kPdStageRows * S::kWarpRow + (size_t)kPdRing * S::kWarpSpan + kPdRing * 8;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(64 / warps, (200 * 1024) / (per_warp * warps + 2048)));   // resident CTAs per SM: warp slots, shared memory (4 - 6 in practice)
    // waves of CTAs to aim for and the shortest segment, by measurement (profiles/_lin_segs.py): BGR is best at ~12 waves of short segments
    // (1080p -> 720p x64: 0.088 ms at 11 rows per segment, 0.092 at 22, 0.102 at 45), planes -- two to three times the columns per thread,
    // hence a heavier prologue -- at ~4 waves of >= 16 rows (CHW 1080p -> 720p x64: 0.093 ms at 45 rows, 0.105 at 11)
    // (planes at 2 : 1 -- four source bytes per output byte, waiting on its row copies -- behave like BGR: 16 4K frames 0.147 ms at 18 rows,
    // 0.103 at 8)
    constexpr bool kShort = C == 3 || (P == 2 && Q == 1);
    constexpr int kWaves = kShort ? 12 : 4, kMinRows = kShort ? 8 : 16;
    const long long want = (long long)kWaves * per_sm * sm_count(current_device());
    const long long per_seg = (long long)g.cta_strips * images;
    // a segment re-reads one or two source rows and pays the table prologue, but small batches need the CTAs.  Sweep of segments per strip on B200 (profiles/_lin_segs.py): one 1080p frame per call 0.0151 ms at >= 32 rows per
    // segment (46 CTAs), 0.0069 at 5 rows; 16 4K frames 0.0878 -> 0.0837 ms; large batches are within 5 % from 11 to 45 rows.
    const long long segs = std::max<long long>(1, std::min<long long>((want + per_seg - 1) / per_seg, (ho + kMinRows - 1) / kMinRows));
    int rps = (int)((ho + segs - 1) / segs);
    if (const int v = knob(kKnobWalkSegs)) rps = (ho + std::max(1, v) - 1) / std::max(1, v);   // tuning knob: vertical segments per column strip
    rps = std::min(512, std::max(rps, 1));
    g.rows_per_seg = rps;
    g.segs = (ho + rps - 1) / rps;
    const size_t smem = (size_t)(rps + 1) * sizeof(LinRow) + warps * per_warp;
    auto kern = plan->down ? resize_linear3_period_kernel<P, Q, KP, kSigned, true, C> : resize_linear3_period_kernel<P, Q, KP, kSigned, false, C>;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return 0;
        static thread_local unsigned long long attr_devices = 0;   // opt-in per device (the maximum, so any later shape fits)
        const int device = current_device();
        if (device >= 64 || !((attr_devices >> device) & 1)) {
            for (auto k : {resize_linear3_period_kernel<P, Q, KP, kSigned, true, C>, resize_linear3_period_kernel<P, Q, KP, kSigned, false, C>}) {
                cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "resize: %s", cudaGetErrorString(e));
            }
            if (device < 64) attr_devices |= 1ull << device;
        }
    }
    for (int i0 = 0; i0 < images; i0 += 65535) {
        const int n = std::min(images - i0, 65535);
        dim3 grid(g.cta_strips * g.segs, n);
        kern<<<grid, 32 * warps, smem, s>>>(src + (size_t)i0 * g.src_image, dst + (size_t)i0 * g.dst_image, g);
    }
    return check_launch("resize (periodic bilinear walker)") == VACV_OK ? 1 : -1;
}

template <bool kSigned>
int launch_any(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, cudaStream_t s) {
    int rc = launch_linear3_period<3, 2, 4, kSigned>(src, dst, images, w, h, wo, ho, s);            // 3 : 2 (1920 -> 1280): 8 columns per thread
    if (rc == 0) rc = launch_linear3_period<4, 3, 4, kSigned>(src, dst, images, w, h, wo, ho, s);   // 4 : 3 (2560 -> 1920): 12 columns
    if (rc == 0) rc = launch_linear3_period<5, 2, 4, kSigned>(src, dst, images, w, h, wo, ho, s);   // 5 : 2 (1920 -> 768, 1280 -> 512): 8 columns
    if (rc == 0) rc = launch_linear3_period<5, 3, 4, kSigned>(src, dst, images, w, h, wo, ho, s);   // 5 : 3 (1920 -> 1152, 1280 -> 768): 12 columns
    if (rc == 0 && knob(kKnobLinearV) != 2) rc = launch_linear3_period<2, 1, 8, kSigned>(src, dst, images, w, h, wo, ho, s);   // 2 : 1 (3840 -> 1920): 8 columns
    if (rc == 0) rc = launch_linear3_period<2, 1, 4, kSigned>(src, dst, images, w, h, wo, ho, s);   // 2 : 1, widths that are not multiples of 8 columns (LINEAR_V=2: always): 4 columns
    return rc;
}

// planes (CHW tensors plane by plane, grey images): more periods per thread so that a warp's bulk copy of a source row stays >= 768 bytes
template <bool kSigned>
int launch_any_planes(const uint8_t* src, uint8_t* dst, int planes, int w, int h, int wo, int ho, cudaStream_t s) {
    int rc = launch_linear3_period<3, 2, 8, kSigned, 1>(src, dst, planes, w, h, wo, ho, s);            // 3 : 2: 16 columns per thread
    if (rc == 0) rc = launch_linear3_period<4, 3, 8, kSigned, 1>(src, dst, planes, w, h, wo, ho, s);   // 4 : 3: 24 columns
    if (rc == 0) rc = launch_linear3_period<2, 1, 16, kSigned, 1>(src, dst, planes, w, h, wo, ho, s);  // 2 : 1: 16 columns
    if (rc == 0) rc = launch_linear3_period<5, 2, 8, kSigned, 1>(src, dst, planes, w, h, wo, ho, s);   // 5 : 2: 16 columns
    if (rc == 0) rc = launch_linear3_period<5, 3, 8, kSigned, 1>(src, dst, planes, w, h, wo, ho, s);   // 5 : 3: 24 columns
    return rc;
}

}  // namespace

int vacv::try_launch_resize_linear1_period(const uint8_t* src, uint8_t* dst, int planes, int w, int h, int wo, int ho, bool signed_char, cudaStream_t s) {
    if (knob(kKnobLinearV) == 1) return 0;
    return signed_char ? launch_any_planes<true>(src, dst, planes, w, h, wo, ho, s) : launch_any_planes<false>(src, dst, planes, w, h, wo, ho, s);
}

// 1 = launched, 0 = shape not eligible (the caller goes on to the persistent pipeline), < 0 = error
int vacv::try_launch_resize_linear3_period(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, bool signed_char, cudaStream_t s) {
    if (knob(kKnobLinearV) == 1) return 0;   // A/B: VACV_LINEAR_V=1 keeps every shape on the persistent pipeline
    return signed_char ? launch_any<true>(src, dst, images, w, h, wo, ho, s) : launch_any<false>(src, dst, images, w, h, wo, ho, s);
}
