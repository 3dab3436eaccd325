// a11: per-channel mean / population stddev (reference: src/cv/normalize_naive.cpp:7-72).
//
// The reference accumulates sequentially in fp32 (relative error up to 1e-2 at 4K); no parallel
// reduction can or should reproduce that.  The statistic is therefore carried as EXACT integer sums -- per channel
// sum(x) and sum(x^2) in u64 -- which is deterministic, order-independent, and lets a multi-GPU batch statistic be
// one 2*c-element u64 all-reduce between this kernel and vacv_cuda_finalize_mean_stddev.
//
// HBM-bound: one read of the frame, 16-byte loads.  Bytes are regrouped per channel with PRMT and summed with
// dp4a (sum x: dot with 0x01010101; sum x^2: dot with itself), u32 partials per thread, warp-shuffle tree,
// one u64 atomicAdd per counter per CTA.
#include "host_util.cuh"
#include "vacv_common.cuh"

namespace vacv {

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Reduce per-thread partials acc[0..N) over the CTA and add them to out[0..N).
template <int N>
__device__ __forceinline__ void cta_reduce_add(unsigned long long (&acc)[N], unsigned long long* out) {
    __shared__ unsigned long long part[8][N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        unsigned long long v = warp_sum(acc[i]);
        if (lane == 0) part[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        unsigned long long v = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += part[wv][threadIdx.x];
        atomicAdd(out + threadIdx.x, v);
    }
}

// HWC, c == 3.  grid = (ctas_per_frame, frames).  frame bytes = 3*wh; 48-byte groups (16 px) + scalar tail.
__global__ void __launch_bounds__(256) sums_hwc3_kernel(const uint8_t* __restrict__ src, size_t frame_bytes,
                                                         unsigned long long* __restrict__ sums, int per_frame) {
    const uint8_t* f = src + (size_t)blockIdx.y * frame_bytes;
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};   // S0 Q0 S1 Q1 S2 Q2
    unsigned s0 = 0, s1 = 0, s2 = 0, q0 = 0, q1 = 0, q2 = 0;
    const size_t groups = frame_bytes / 48;
    int since_flush = 0;
    for (size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += (size_t)gridDim.x * blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(f + gidx * 48);
        const uint4 v0 = ld_stream16(p), v1 = ld_stream16(p + 1), v2 = ld_stream16(p + 2);
        const unsigned wd[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {   // 12 bytes = 4 pixels: [c0 c1 c2 c0][c1 c2 c0 c1][c2 c0 c1 c2]
            const unsigned a = wd[3 * t], b = wd[3 * t + 1], c = wd[3 * t + 2];
            const unsigned ab0 = __byte_perm(a, b, 0x0630), ch0 = __byte_perm(ab0, c, 0x5210) ;   // a0 a3 b2 c1
            const unsigned ab1 = __byte_perm(a, b, 0x0741), ch1 = __byte_perm(ab1, c, 0x6210);    // a1 b0 b3 c2
            const unsigned ab2 = __byte_perm(a, b, 0x0052), ch2 = __byte_perm(ab2, c, 0x7410);    // a2 b1 c0 c3
            s0 = __dp4a(ch0, 0x01010101u, s0); q0 = __dp4a(ch0, ch0, q0);
            s1 = __dp4a(ch1, 0x01010101u, s1); q1 = __dp4a(ch1, ch1, q1);
            s2 = __dp4a(ch2, 0x01010101u, s2); q2 = __dp4a(ch2, ch2, q2);
        }
        if (++since_flush == 2048) {   // q grows by <= 16*255^2 per iteration: flush well before 2^32
            acc[0] += s0; acc[1] += q0; acc[2] += s1; acc[3] += q1; acc[4] += s2; acc[5] += q2;
            s0 = s1 = s2 = q0 = q1 = q2 = 0; since_flush = 0;
        }
    }
    acc[0] += s0; acc[1] += q0; acc[2] += s1; acc[3] += q1; acc[4] += s2; acc[5] += q2;
    // tail bytes (< 48) of the frame
    if (blockIdx.x == 0) {
        for (size_t i = groups * 48 + threadIdx.x; i < frame_bytes; i += blockDim.x) {
            const unsigned v = f[i];
            const int k = (int)(i % 3);
            acc[2 * k] += v; acc[2 * k + 1] += v * v;
        }
    }
    cta_reduce_add<6>(acc, sums + (per_frame ? (size_t)blockIdx.y * 6 : 0));
}

// single-channel planes (c == 1, or each plane of a CHW tensor).  grid = (ctas_per_plane, planes).
// plane i belongs to frame i / c, channel i % c.
__global__ void __launch_bounds__(256) sums_plane_kernel(const uint8_t* __restrict__ src, size_t plane_bytes, int c,
                                                          unsigned long long* __restrict__ sums, int per_frame) {
    const uint8_t* f = src + (size_t)blockIdx.y * plane_bytes;
    unsigned long long acc[2] = {0, 0};
    unsigned s = 0, q = 0;
    const size_t head = min(plane_bytes, (size_t)((16 - ((uintptr_t)f & 15)) & 15));
    const size_t chunks = (plane_bytes - head) / 16;
    int since_flush = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = ld_stream16(f + head + 16 * i);
        s = __dp4a(v.x, 0x01010101u, s); q = __dp4a(v.x, v.x, q);
        s = __dp4a(v.y, 0x01010101u, s); q = __dp4a(v.y, v.y, q);
        s = __dp4a(v.z, 0x01010101u, s); q = __dp4a(v.z, v.z, q);
        s = __dp4a(v.w, 0x01010101u, s); q = __dp4a(v.w, v.w, q);
        if (++since_flush == 2048) { acc[0] += s; acc[1] += q; s = q = 0; since_flush = 0; }
    }
    acc[0] += s; acc[1] += q;
    if (blockIdx.x == 0) {   // unaligned head and < 16-byte tail of the plane
        const size_t tail0 = head + 16 * chunks, ntail = head + (plane_bytes - tail0);
        for (size_t j = threadIdx.x; j < ntail; j += blockDim.x) {
            const unsigned v = f[j < head ? j : tail0 + (j - head)];
            acc[0] += v; acc[1] += v * v;
        }
    }
    const int frame = blockIdx.y / c, k = blockIdx.y % c;
    cta_reduce_add<2>(acc, sums + (per_frame ? (size_t)frame * 2 * c : 0) + 2 * k);
}

// any c, HWC: one thread per pixel (correct, not fast; the reference's HWC statistics are 3-channel only)
__global__ void sums_hwc_generic_kernel(const uint8_t* __restrict__ src, size_t wh, int c,
                                        unsigned long long* __restrict__ sums, int per_frame) {
    const uint8_t* f = src + (size_t)blockIdx.y * wh * c;
    unsigned long long* out = sums + (per_frame ? (size_t)blockIdx.y * 2 * c : 0);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < wh; p += (size_t)gridDim.x * blockDim.x)
        for (int k = 0; k < c; ++k) {
            const unsigned long long v = f[p * c + k];
            atomicAdd(out + 2 * k, v);
            atomicAdd(out + 2 * k + 1, v * v);
        }
}

// fp32 pixels, any layout, c <= 4.  grid = (ctas, images); HWC: image = frame, channel = element index mod c;
// CHW: image = plane, one channel per image.  fp64 accumulation, warp-shuffle tree, one atomicAdd(double) per CTA.
template <int C>
__global__ void __launch_bounds__(256) sums_f32_kernel(const float* __restrict__ src, size_t image_elems, int c_total,
                                                        double* __restrict__ sums, int per_frame) {
    const float* f = src + (size_t)blockIdx.y * image_elems;
    double s[C], q[C];
#pragma unroll
    for (int k = 0; k < C; ++k) s[k] = q[k] = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned ph = (unsigned)(i % C);
    const unsigned dph = (unsigned)(stride % C);
    for (; i < image_elems; i += stride) {
        const double v = (double)__ldg(f + i);
#pragma unroll
        for (int k = 0; k < C; ++k) if (ph == (unsigned)k) { s[k] += v; q[k] += v * v; }
        ph += dph;
        if (ph >= C) ph -= C;
    }
    __shared__ double part[8][2 * C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        double a = s[k], b = q[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
        if (lane == 0) { part[warp][2 * k] = a; part[warp][2 * k + 1] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * C) {
        double v = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += part[wv][threadIdx.x];
        // HWC (C == c_total): set = frame; planes (C == 1): image = frame * c_total + channel
        const size_t frame = C == 1 ? blockIdx.y / c_total : blockIdx.y;
        const int k0 = C == 1 ? 2 * (int)(blockIdx.y % c_total) : 0;
        atomicAdd(sums + (per_frame ? frame * 2 * c_total : 0) + k0 + threadIdx.x, v);
    }
}

__global__ void finalize_mean_stddev_f64_kernel(const double* __restrict__ sums, int total, double n,
                                                float* __restrict__ mean, float* __restrict__ stddev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double m = sums[2 * i] / n;
    double var = sums[2 * i + 1] / n - m * m;
    if (var < 0) var = 0;
    mean[i] = (float)m;
    stddev[i] = (float)sqrt(var);
}

__global__ void finalize_mean_stddev_kernel(const unsigned long long* __restrict__ sums, int total, double n,
                                            float* __restrict__ mean, float* __restrict__ stddev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (set, channel) flattened
    if (i >= total) return;
    const double m = (double)sums[2 * i] / n;
    double var = (double)sums[2 * i + 1] / n - m * m;
    if (var < 0) var = 0;
    mean[i] = (float)m;
    stddev[i] = (float)sqrt(var);
}

}  // namespace vacv

using namespace vacv;

extern "C" int vacv_cuda_sums_u8(const uint8_t* src, int batch, int w, int h, int c, int layout,
                                 unsigned long long* sums, int per_frame, void* stream) {
    VACV_REQUIRE(src && sums, "sums_u8: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0, "sums_u8: non-positive size");
    cudaStream_t s = as_stream(stream);
    const size_t wh = (size_t)w * h;
    // enough CTAs to fill the machine several times over, few enough that atomics stay negligible
    const int sms = current_sm_count();
    auto ctas_for = [sms](size_t bytes, int images) {
        size_t want = (size_t)sms * 16 / (size_t)max(1, images) + 1;
        return (unsigned)max((size_t)1, min(want, (bytes + 256 * 48 - 1) / (256 * 48)));
    };
    if (layout == VACV_NHWC && c == 3 && (((uintptr_t)src & 15) == 0) && ((wh * 3) % 16 == 0 || batch == 1)) {
        for (int f0 = 0; f0 < batch; f0 += 65535) {
            const int nf = min(batch - f0, 65535);
            dim3 grid(ctas_for(wh * 3, nf), nf);
            sums_hwc3_kernel<<<grid, 256, 0, s>>>(src + (size_t)f0 * wh * 3, wh * 3, sums + (per_frame ? (size_t)f0 * 6 : 0), per_frame);
        }
    } else if (layout == VACV_NCHW || c == 1) {
        const int planes = batch * c;
        VACV_REQUIRE(planes <= 65535 * 64, "sums_u8: too many planes");
        for (int p0 = 0; p0 < planes; p0 += 65535 / c * c) {
            const int np = min(planes - p0, 65535 / c * c);
            dim3 grid(ctas_for(wh, np), np);
            sums_plane_kernel<<<grid, 256, 0, s>>>(src + (size_t)p0 * wh, wh, c, sums + (per_frame ? (size_t)(p0 / c) * 2 * c : 0), per_frame);
        }
    } else {
        VACV_REQUIRE(batch <= 65535, "sums_u8: generic path supports batch <= 65535");
        dim3 grid((unsigned)min((size_t)1024, (wh + 255) / 256), batch);
        sums_hwc_generic_kernel<<<grid, 256, 0, s>>>(src, wh, c, sums, per_frame);
    }
    return check_launch("sums_u8");
}

extern "C" int vacv_cuda_finalize_mean_stddev(const unsigned long long* sums, int n_sets, int c,
                                              unsigned long long n_per_channel, float* mean, float* stddev, void* stream) {
    VACV_REQUIRE(sums && mean && stddev, "finalize_mean_stddev: null pointer");
    VACV_REQUIRE(n_sets > 0 && c > 0 && n_per_channel > 0, "finalize_mean_stddev: bad size");
    const int total = n_sets * c;
    finalize_mean_stddev_kernel<<<ceil_div(total, 128), 128, 0, as_stream(stream)>>>(sums, total, (double)n_per_channel, mean, stddev);
    return check_launch("finalize_mean_stddev");
}

extern "C" int vacv_cuda_sums_f32(const float* src, int batch, int w, int h, int c, int layout, double* sums, int per_frame, void* stream) {
    VACV_REQUIRE(src && sums, "sums_f32: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0, "sums_f32: non-positive size");
    if (c > 4) return set_error(VACV_ERR_UNSUPPORTED, "sums_f32: c <= 4");
    cudaStream_t s = as_stream(stream);
    const size_t wh = (size_t)w * h;
    const bool planes = layout == VACV_NCHW || c == 1;
    const long long images = planes ? (long long)batch * c : batch;
    VACV_REQUIRE(images <= 65535, "sums_f32: at most 65535 frames (planes) per call");
    const size_t elems = planes ? wh : wh * c;
    const int sms = current_sm_count();
    const unsigned ctas = (unsigned)max((size_t)1, min((elems + 256 * 16 - 1) / (256 * 16), (size_t)(sms * 16 / (size_t)images + 1)));
    dim3 grid(ctas, (unsigned)images);
    if (planes) sums_f32_kernel<1><<<grid, 256, 0, s>>>(src, elems, c, sums, per_frame);
    else if (c == 2) sums_f32_kernel<2><<<grid, 256, 0, s>>>(src, elems, c, sums, per_frame);
    else if (c == 3) sums_f32_kernel<3><<<grid, 256, 0, s>>>(src, elems, c, sums, per_frame);
    else sums_f32_kernel<4><<<grid, 256, 0, s>>>(src, elems, c, sums, per_frame);
    return check_launch("sums_f32");
}

extern "C" int vacv_cuda_finalize_mean_stddev_f64(const double* sums, int n_sets, int c, unsigned long long n_per_channel,
                                                  float* mean, float* stddev, void* stream) {
    VACV_REQUIRE(sums && mean && stddev, "finalize_mean_stddev_f64: null pointer");
    VACV_REQUIRE(n_sets > 0 && c > 0 && n_per_channel > 0, "finalize_mean_stddev_f64: bad size");
    const int total = n_sets * c;
    finalize_mean_stddev_f64_kernel<<<ceil_div(total, 128), 128, 0, as_stream(stream)>>>(sums, total, (double)n_per_channel, mean, stddev);
    return check_launch("finalize_mean_stddev_f64");
}
