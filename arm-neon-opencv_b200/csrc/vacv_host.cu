// Host side of the C-ABI: error reporting and the matrix helpers of warp_affine, which the reference evaluates on
// the host in mixed float/double (src/cv/warp_affine.cpp:76-133).  Built with -ffp-contract=off for baseline
// x86-64, like the reference (no FMA), so the results are bit-identical to its build.
#include <cmath>
#include <cstdarg>
#include <cstdio>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "host_util.cuh"
#include "vacv_common.cuh"

namespace vacv {

static thread_local char g_err[512] = "";

// ---- tuning knobs: environment read once, then plain atomics (host_util.cuh)
static std::atomic<int> g_knobs[kKnobCount];
static std::atomic<int> g_knob_gen{0};
static std::once_flag g_knob_once;
static const char* const kKnobNames[kKnobCount] = {"PIPE_NCOL", "RPIPE_NCOL", "WARP_GATHER", "NO_RPIPE", "RESIZE_NORMALIZE_GATHER",
                                                    "WALK_SEGS", "WALK_SYNC", "WALK2_SYNC", "CUBIC3", "CUBIC_V", "PIPE_ROWS", "STREAM_QPT", "WARP_V", "LINEAR_V"};
static void init_knobs() {
    for (int k = 0; k < kKnobCount; ++k) {
        char name[64];
        snprintf(name, sizeof(name), "VACV_%s", kKnobNames[k]);
        const char* e = getenv(name);
        int v = 0;
        if (e) v = k == kKnobCubic3Roll ? (strcmp(e, "roll") == 0) : (*e >= '0' && *e <= '9') ? atoi(e) : 1;
        g_knobs[k].store(v, std::memory_order_relaxed);
    }
}
int knob(Knob k) {
    std::call_once(g_knob_once, init_knobs);
    return g_knobs[k].load(std::memory_order_relaxed);
}
int knob_generation() { return g_knob_gen.load(std::memory_order_relaxed); }

int sm_count(int device) {
    static std::atomic<int> cached[64];
    if (device < 0 || device >= 64) device = 0;
    int n = cached[device].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) { cudaGetLastError(); n = kNumSMs; }
        cached[device].store(n, std::memory_order_relaxed);
    }
    return n;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) return VACV_OK;
    cudaGetLastError();   // clear the (non-sticky) launch error
    return set_error(VACV_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

}  // namespace vacv

extern "C" int vacv_cuda_abi_version(void) { return 1; }

extern "C" int vacv_cuda_set_tuning(const char* name, int value) {
    VACV_REQUIRE(name, "set_tuning: null name");
    for (int k = 0; k < vacv::kKnobCount; ++k)
        if (strcmp(name, vacv::kKnobNames[k]) == 0) {
            vacv::knob((vacv::Knob)k);   // make sure the environment defaults are in before overriding one
            vacv::g_knobs[k].store(value, std::memory_order_relaxed);
            vacv::g_knob_gen.fetch_add(1, std::memory_order_relaxed);
            return VACV_OK;
        }
    return vacv::set_error(VACV_ERR_INVALID_ARG, "set_tuning: unknown knob '%s'", name);
}
extern "C" const char* vacv_cuda_last_error(void) { return vacv::g_err; }
extern "C" int vacv_cuda_set_last_error(int code, const char* message) { return vacv::set_error(code, "%s", message ? message : ""); }

// warp_affine.cpp:121-133, types exactly as written there (float products widened afterwards; m[1] *= -D is
// float*double rounded back to float; b1/b2 use the UPDATED m[0], m[1], m[3], m[4] in float arithmetic).
extern "C" void vacv_invert_affine(float* m) {
    double D = m[0] * m[4] - m[1] * m[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = m[4] * D;
    double A22 = m[0] * D;
    m[0] = (float)A11;
    m[1] = (float)(m[1] * -D);
    m[3] = (float)(m[3] * -D);
    m[4] = (float)A22;
    double b1 = -m[0] * m[2] - m[1] * m[5];
    double b2 = -m[3] * m[2] - m[4] * m[5];
    m[2] = (float)b1;
    m[5] = (float)b2;
}

// warp_affine.cpp:76-94 (get_rotation_matrix_2D about (0,0)) followed by the aux translation of :105-106.
extern "C" void vacv_rotation_matrix(float scale, float rot_deg, const double* aux, float* m) {
    float angle = rot_deg;
    angle *= M_PI / 180;
    const double alpha = scale * cos(angle);
    const double beta = scale * sin(angle);
    m[0] = (float)alpha;
    m[1] = (float)beta;
    m[3] = (float)-beta;
    m[4] = (float)alpha;
    m[2] = (float)(aux[2] - m[0] * aux[0] - m[1] * aux[1]);
    m[5] = (float)(aux[3] - m[3] * aux[0] - m[4] * aux[1]);
}

// ---- runtime helpers (the host layer's only door to CUDA) -------------------------------------------------------
#define VACV_RT(call, what)                                                              \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return vacv::set_error(VACV_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e_)); \
        return VACV_OK;                                                                  \
    } while (0)

extern "C" int vacv_cuda_device_count(int* count) {
    VACV_REQUIRE(count, "device_count: null pointer");
    *count = 0;
    VACV_RT(cudaGetDeviceCount(count), "device_count");
}
extern "C" int vacv_cuda_set_device(int device) { VACV_RT(cudaSetDevice(device), "set_device"); }
extern "C" int vacv_cuda_get_device(int* device) {
    VACV_REQUIRE(device, "get_device: null pointer");
    VACV_RT(cudaGetDevice(device), "get_device");
}
extern "C" int vacv_cuda_malloc(void** dptr, size_t bytes) {
    VACV_REQUIRE(dptr, "malloc: null pointer");
    VACV_RT(cudaMalloc(dptr, bytes), "malloc");
}
extern "C" int vacv_cuda_free(void* dptr) { VACV_RT(cudaFree(dptr), "free"); }
extern "C" int vacv_cuda_host_alloc(void** h_ptr, size_t bytes) {
    VACV_REQUIRE(h_ptr, "host_alloc: null pointer");
    VACV_RT(cudaHostAlloc(h_ptr, bytes, cudaHostAllocDefault), "host_alloc");
}
extern "C" int vacv_cuda_host_alloc_flags(void** h_ptr, size_t bytes, int write_combined) {
    VACV_REQUIRE(h_ptr, "host_alloc_flags: null pointer");
    VACV_RT(cudaHostAlloc(h_ptr, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault), "host_alloc_flags");
}
extern "C" int vacv_cuda_host_free(void* h_ptr) { VACV_RT(cudaFreeHost(h_ptr), "host_free"); }
extern "C" int vacv_cuda_memcpy_h2d(void* dptr, const void* h_ptr, size_t bytes, void* stream) {
    VACV_RT(cudaMemcpyAsync(dptr, h_ptr, bytes, cudaMemcpyHostToDevice, vacv::as_stream(stream)), "memcpy_h2d");
}
extern "C" int vacv_cuda_memcpy_d2h(void* h_ptr, const void* dptr, size_t bytes, void* stream) {
    VACV_RT(cudaMemcpyAsync(h_ptr, dptr, bytes, cudaMemcpyDeviceToHost, vacv::as_stream(stream)), "memcpy_d2h");
}
extern "C" int vacv_cuda_memcpy2d_h2d(void* dptr, size_t dst_pitch, const void* h_ptr, size_t src_pitch, size_t row_bytes, size_t rows, void* stream) {
    VACV_RT(cudaMemcpy2DAsync(dptr, dst_pitch, h_ptr, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, vacv::as_stream(stream)), "memcpy2d_h2d");
}
extern "C" int vacv_cuda_memset(void* dptr, int value, size_t bytes, void* stream) {
    VACV_RT(cudaMemsetAsync(dptr, value, bytes, vacv::as_stream(stream)), "memset");
}
extern "C" int vacv_cuda_stream_create(void** stream) {
    VACV_REQUIRE(stream, "stream_create: null pointer");
    VACV_RT(cudaStreamCreateWithFlags(reinterpret_cast<cudaStream_t*>(stream), cudaStreamNonBlocking), "stream_create");
}
extern "C" int vacv_cuda_stream_destroy(void* stream) { VACV_RT(cudaStreamDestroy(vacv::as_stream(stream)), "stream_destroy"); }
extern "C" int vacv_cuda_stream_sync(void* stream) { VACV_RT(cudaStreamSynchronize(vacv::as_stream(stream)), "stream_sync"); }

// ---- end-to-end host-buffer pipeline ------------------------------------------------------------------------------
namespace {
// Streams, events and double-buffered device staging of the host-buffer entry points.  One instance per (host thread,
// device): everything in it belongs to the device that was current when it was created, so a thread that alternates
// vacv_cuda_set_device() gets a separate pipeline per GPU (the reference's multi-device hook is
// CudaDevice::set_device, src/cv/cuda_device.cu:10-18).
struct HostPipe {
    int device = -1;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t in_done[2] = {}, k_done[2] = {}, out_done[2] = {};
    uint8_t* d_in[2] = {};
    float* d_out[2] = {};
    float* d_stats = nullptr;
    size_t in_cap = 0, out_cap = 0;
    bool ready = false;
    cudaError_t init() {
        if (ready) return cudaSuccess;
        cudaError_t e;
        if (!s_in && (e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if (!s_k && (e = cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if (!s_out && (e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int b = 0; b < 2; ++b) {
            if (!in_done[b] && (e = cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming)) != cudaSuccess) return e;
            if (!k_done[b] && (e = cudaEventCreateWithFlags(&k_done[b], cudaEventDisableTiming)) != cudaSuccess) return e;
            if (!out_done[b] && (e = cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if (!d_stats && (e = cudaMalloc(&d_stats, 6 * sizeof(float))) != cudaSuccess) return e;
        ready = true;
        return cudaSuccess;
    }
    // grow one pair of staging buffers; the capacity is only raised once BOTH allocations exist
    template <typename T>
    cudaError_t grow(T* (&buf)[2], size_t& cap, size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        cudaError_t e = drain();
        if (e != cudaSuccess) return e;
        cap = 0;
        for (int b = 0; b < 2; ++b) { if (buf[b]) cudaFree(buf[b]); buf[b] = nullptr; }
        for (int b = 0; b < 2; ++b)
            if ((e = cudaMalloc(&buf[b], bytes)) != cudaSuccess) {
                for (int q = 0; q < 2; ++q) { if (buf[q]) cudaFree(buf[q]); buf[q] = nullptr; }
                return e;
            }
        cap = bytes;
        return cudaSuccess;
    }
    cudaError_t reserve(size_t in_bytes, size_t out_bytes) {
        cudaError_t e = grow(d_in, in_cap, in_bytes);
        return e != cudaSuccess ? e : grow(d_out, out_cap, out_bytes);
    }
    // nothing of this pipeline may still be in flight on the caller's buffers when a "synchronous" call returns
    cudaError_t drain() {
        cudaError_t first = cudaSuccess;
        for (cudaStream_t st : {s_in, s_k, s_out})
            if (st) { const cudaError_t e = cudaStreamSynchronize(st); if (first == cudaSuccess) first = e; }
        return first;
    }
    void destroy() {   // called with `device` current
        drain();
        for (int b = 0; b < 2; ++b) {
            if (d_in[b]) cudaFree(d_in[b]);
            if (d_out[b]) cudaFree(d_out[b]);
            if (in_done[b]) cudaEventDestroy(in_done[b]);
            if (k_done[b]) cudaEventDestroy(k_done[b]);
            if (out_done[b]) cudaEventDestroy(out_done[b]);
        }
        if (d_stats) cudaFree(d_stats);
        for (cudaStream_t st : {s_in, s_k, s_out}) if (st) cudaStreamDestroy(st);
        *this = HostPipe();
    }
};

constexpr int kMaxPipeDevices = 16;
struct HostPipes {
    HostPipe pipe[kMaxPipeDevices];
    ~HostPipes() {   // thread exit: release what this thread created (best effort: the context may already be gone at process exit)
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return; }
        for (HostPipe& p : pipe)
            if (p.device >= 0 && cudaSetDevice(p.device) == cudaSuccess) p.destroy();
        cudaSetDevice(prev);
        cudaGetLastError();
    }
};
thread_local HostPipes g_pipes;

// the pipeline of the CURRENT device (nullptr + error message if it cannot be determined)
HostPipe* current_pipe(const char* who) {
    int dev = -1;
    const cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= kMaxPipeDevices) {
        vacv::set_error(VACV_ERR_CUDA, "%s: %s", who, e != cudaSuccess ? cudaGetErrorString(e) : "device index beyond the per-thread pipeline table");
        return nullptr;
    }
    HostPipe& p = g_pipes.pipe[dev];
    p.device = dev;
    return &p;
}
}  // namespace

// Chunked, three-stream pipeline shared by the host-buffer entry points: chunk i's H2D copy, chunk i-1's kernel and chunk
// i-2's D2H copy overlap.  `launch(d_in, d_out, frames, stream)` enqueues the kernel(s) of one chunk and returns a status.
// in_frame = distance between frames in h_src; last_frame = bytes of a frame that are actually defined (a pitched decoder
// pool may end right after the last surface: the tail of the final frame_stride is never read).
template <typename Launch>
static int run_host_pipeline(const char* who, const uint8_t* h_src, void* h_dst, int batch, size_t in_frame, size_t last_frame, size_t out_frame,
                             int chunk_frames, const float* h_mean, const float* h_stddev, float** d_stats_out, Launch launch) {
    HostPipe* pp = current_pipe(who);
    if (!pp) return VACV_ERR_CUDA;
    HostPipe& p = *pp;
    // on any failure: wait for whatever was already queued on the caller's buffers, then report
#define VACV_CUP(call)                                                                                          \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) { p.drain(); cudaGetLastError(); return vacv::set_error(VACV_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e_)); } \
    } while (0)
    VACV_CUP(p.init());
    const int chunk = chunk_frames < batch ? chunk_frames : batch;
    VACV_CUP(p.reserve(in_frame * chunk, out_frame * chunk));
    float stats[6] = {h_mean[0], h_mean[1], h_mean[2], h_stddev[0], h_stddev[1], h_stddev[2]};
    VACV_CUP(cudaMemcpyAsync(p.d_stats, stats, sizeof(stats), cudaMemcpyHostToDevice, p.s_k));
    VACV_CUP(cudaStreamSynchronize(p.s_k));   // `stats` lives on this stack frame
    *d_stats_out = p.d_stats;
    int i = 0;
    for (int f0 = 0; f0 < batch; f0 += chunk, ++i) {
        const int b = i & 1, n = (batch - f0 < chunk) ? batch - f0 : chunk;
        const size_t in_bytes = f0 + n == batch ? in_frame * (n - 1) + last_frame : in_frame * n;
        // H2D of chunk i may start once the kernel that last read d_in[b] (chunk i-2) is done
        if (i >= 2) VACV_CUP(cudaStreamWaitEvent(p.s_in, p.k_done[b], 0));
        VACV_CUP(cudaMemcpyAsync(p.d_in[b], h_src + (size_t)f0 * in_frame, in_bytes, cudaMemcpyHostToDevice, p.s_in));
        VACV_CUP(cudaEventRecord(p.in_done[b], p.s_in));
        // kernel of chunk i: needs its input, and d_out[b] drained by the D2H of chunk i-2
        VACV_CUP(cudaStreamWaitEvent(p.s_k, p.in_done[b], 0));
        if (i >= 2) VACV_CUP(cudaStreamWaitEvent(p.s_k, p.out_done[b], 0));
        const int rc = launch(p.d_in[b], (void*)p.d_out[b], n, (void*)p.s_k);
        if (rc != VACV_OK) { p.drain(); return rc; }
        VACV_CUP(cudaEventRecord(p.k_done[b], p.s_k));
        VACV_CUP(cudaStreamWaitEvent(p.s_out, p.k_done[b], 0));
        VACV_CUP(cudaMemcpyAsync((uint8_t*)h_dst + (size_t)f0 * out_frame, p.d_out[b], out_frame * n, cudaMemcpyDeviceToHost, p.s_out));
        VACV_CUP(cudaEventRecord(p.out_done[b], p.s_out));
    }
    VACV_CUP(cudaStreamSynchronize(p.s_out));
    VACV_CUP(cudaStreamSynchronize(p.s_k));
    return VACV_OK;
#undef VACV_CUP
}

extern "C" int vacv_cuda_nv_resize_normalize_chw_host(const uint8_t* h_src, float* h_dst, int batch, int w, int h, int v_first,
                                                      int w_out, int h_out, const float* h_mean, const float* h_stddev,
                                                      int chunk_frames) {
    VACV_REQUIRE(h_src && h_dst && h_mean && h_stddev, "nv_resize_normalize_chw_host: null pointer");
    VACV_REQUIRE(batch > 0 && chunk_frames > 0, "nv_resize_normalize_chw_host: bad batch / chunk");
    float* d_stats = nullptr;
    float** ds = &d_stats;
    return run_host_pipeline("nv_resize_normalize_chw_host", h_src, h_dst, batch, (size_t)w * h * 3 / 2, (size_t)w * h * 3 / 2, (size_t)w_out * h_out * 3 * sizeof(float),
                             chunk_frames, h_mean, h_stddev, ds, [=](const uint8_t* d_in, void* d_out, int n, void* stream) {
                                 return vacv_cuda_nv_resize_normalize_chw(d_in, (float*)d_out, n, w, h, v_first, w_out, h_out, *ds, *ds + 3, stream);
                             });
}

// Host-buffer form of the decoder-surface entries: pitched / planar frames in host memory in, fp32 / fp16 / bf16 planes back in
// host memory; content == NULL: plain resize to canvas_w x canvas_h, else letterbox placement (see the device entries).
extern "C" int vacv_cuda_yuv_normalize_chw_host(const uint8_t* h_src, const vacv_yuv_layout* layout, void* h_dst, int out_dtype, int batch,
                                                int canvas_w, int canvas_h, const vacv_rect* content, const uint8_t* pad_bgr,
                                                const float* h_mean, const float* h_stddev, int chunk_frames) {
    VACV_REQUIRE(h_src && layout && h_dst && h_mean && h_stddev, "yuv_normalize_chw_host: null pointer");
    VACV_REQUIRE(batch > 0 && chunk_frames > 0 && canvas_w > 0 && canvas_h > 0, "yuv_normalize_chw_host: bad batch / chunk / size");
    VACV_REQUIRE(!content || pad_bgr, "yuv_normalize_chw_host: letterbox needs a pad colour");
    if (out_dtype != VACV_FP32 && out_dtype != VACV_FP16 && out_dtype != VACV_BF16)
        return vacv::set_error(VACV_ERR_UNSUPPORTED, "yuv_normalize_chw_host: out dtype %d (FP32, FP16 or BF16)", out_dtype);
    const bool planar = layout->format == VACV_YUV_I420 || layout->format == VACV_YUV_YV12;
    const size_t yp = layout->y_pitch ? layout->y_pitch : layout->w, cp = layout->c_pitch ? layout->c_pitch : (planar ? layout->w / 2 : layout->w);
    const size_t surface = yp * layout->h + cp * (layout->h / 2) * (planar ? 2 : 1);   // bytes of one surface that are defined
    const size_t in_frame = layout->frame_stride ? layout->frame_stride : surface;
    const size_t out_frame = (size_t)canvas_w * canvas_h * 3 * (out_dtype == VACV_FP32 ? 4 : 2);
    vacv_yuv_layout lay = *layout;
    lay.frame_stride = in_frame;
    vacv_rect box = content ? *content : vacv_rect{0, 0, canvas_w, canvas_h};
    uint8_t pad[3] = {pad_bgr ? pad_bgr[0] : (uint8_t)0, pad_bgr ? pad_bgr[1] : (uint8_t)0, pad_bgr ? pad_bgr[2] : (uint8_t)0};
    float* d_stats = nullptr;
    float** ds = &d_stats;
    const bool letterbox = content != nullptr;
    return run_host_pipeline("yuv_normalize_chw_host", h_src, h_dst, batch, in_frame, surface < in_frame ? surface : in_frame, out_frame, chunk_frames, h_mean, h_stddev, ds,
                             [=](const uint8_t* d_in, void* d_out, int n, void* stream) {
                                 if (letterbox)
                                     return vacv_cuda_yuv_letterbox_normalize_chw(d_in, &lay, d_out, out_dtype, n, canvas_w, canvas_h, &box, pad, *ds, *ds + 3,
                                                                                  h_mean, h_stddev, stream);
                                 return vacv_cuda_yuv_resize_normalize_chw(d_in, &lay, d_out, out_dtype, n, canvas_w, canvas_h, *ds, *ds + 3, stream);
                             });
}
