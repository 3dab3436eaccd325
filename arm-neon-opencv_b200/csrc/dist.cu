// Config 5: batch-global mean / stddev + normalise over a frame batch sharded across GPUs.
//
// The reference computes the statistic of ONE image inside Normalize::normalize_naive (src/cv/normalize.cpp:98-108);
// here the same statistic is taken over every frame of every rank.  Because stats.cu carries it as exact u64 sums, the
// only exchange on the whole vacv path is a sum all-reduce of 2*c+1 integers between the reduction kernel and the
// normalise kernel -- 56 bytes: pure latency, no bandwidth.  Two transports live here (NCCL is in dist_nccl.cpp):
//
//  * callback -- the caller's all-reduce;
//  * peer memory -- every rank owns a small slot buffer that all peers map (CUDA IPC).  One 1-CTA kernel per exchange:
//    lane r STORES this rank's values straight into rank r's buffer over NVLink / NVSwitch (fire and forget), fences, raises
//    the flag there, then polls the flag rank r raised in ITS OWN buffer (local L2, nobody spins over the fabric), and sums
//    the slots in rank order.  No library call, no host involvement: the epoch counter lives on the device, so the kernel
//    replays unchanged inside a CUDA graph.  Two slot sets alternate by epoch parity: a peer can only be one exchange ahead
//    (it needs MY flag of epoch e+1 before it can start e+2, and I raise that only after my epoch-e kernel has finished).
#include <cstring>

#include "vacv_common.cuh"

namespace vacv {

constexpr int kXchgWords = 16;   // 15 values + flag; one 128-byte line per (parity, rank)

struct XchgPeers { unsigned long long* buf[VACV_P2P_MAX_RANKS]; };

// layout of a rank's buffer (u64 words): [2][nranks][16] slots, then {epoch, timeouts}
__device__ __forceinline__ size_t slot_ofs(int parity, int nranks, int r) { return ((size_t)parity * nranks + r) * kXchgWords; }
static inline size_t xchg_words(int nranks) { return (size_t)2 * nranks * kXchgWords + 2; }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(32) p2p_allreduce_u64_kernel(XchgPeers peers, int nranks, int rank, unsigned long long* data, int count) {
    unsigned long long* mine = peers.buf[rank];
    unsigned long long* ctl = mine + (size_t)2 * nranks * kXchgWords;
    const int lane = threadIdx.x;
    unsigned long long epoch = 0;
    if (lane == 0) { epoch = ctl[0] + 1; ctl[0] = epoch; }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    const int parity = (int)(epoch & 1);
    bool ok = true;
    if (lane < nranks) {
        // publish: my values into MY slot of rank `lane`'s buffer (own buffer included), then the flag
        unsigned long long* dst = peers.buf[lane] + slot_ofs(parity, nranks, rank);
        for (int i = 0; i < count; ++i) dst[i] = data[i];
        __threadfence_system();
        st_release_sys(dst + kXchgWords - 1, epoch);
        // wait for rank `lane`'s flag in my own buffer
        const unsigned long long* flag = mine + slot_ofs(parity, nranks, lane) + kXchgWords - 1;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(flag) != epoch) {
            if (global_timer_ns() - t0 > 20000000000ull) { ok = false; break; }
            __nanosleep(200);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok && lane == 0) ctl[1] += 1;
    if (lane < count) {
        unsigned long long v = 0;
        for (int r = 0; r < nranks; ++r) v += __ldcg(mine + slot_ofs(parity, nranks, r) + lane);
        data[lane] = v;
    }
}

// d_work[0 .. 2c) = 0, d_work[2c] = pixels per channel of this rank's shard
__global__ void batch_global_init_kernel(unsigned long long* work, int c, unsigned long long n) {
    const int i = threadIdx.x;
    if (i < 2 * c) work[i] = 0;
    else if (i == 2 * c) work[i] = n;
}

// mean / population stddev from the all-reduced counters; the pixel count is the reduced work[2c]
__global__ void batch_global_finalize_kernel(const unsigned long long* __restrict__ work, int c, float* __restrict__ mean_std) {
    const int k = threadIdx.x;
    if (k >= c) return;
    const double n = (double)work[2 * c];
    const double m = (double)work[2 * k] / n;
    double var = (double)work[2 * k + 1] / n - m * m;
    if (var < 0) var = 0;
    mean_std[k] = (float)m;
    mean_std[c + k] = (float)sqrt(var);
}

struct Xchg {
    int nranks = 0, rank = 0, device = 0;
    bool connected = false;
    unsigned long long* mine = nullptr;
    XchgPeers peers = {};
};

static int p2p_allreduce_cb(void* ctx, unsigned long long* d_buf, int count, void* stream) {
    return vacv_cuda_p2p_allreduce_u64(ctx, d_buf, count, stream);
}

}  // namespace vacv

using namespace vacv;

#define VACV_DCU(call, who)                                                                                    \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) return set_error(VACV_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e_));         \
    } while (0)

extern "C" int vacv_cuda_normalize_batch_global_cb(vacv_allreduce_u64_fn allreduce, void* ctx, const uint8_t* src, float* dst,
                                                   int batch, int w, int h, int c, int layout, unsigned long long* d_work,
                                                   float* d_mean_std, void* ev_sums_done, void* ev_stats_ready, void* stream) {
    VACV_REQUIRE(allreduce && src && dst && d_work && d_mean_std, "normalize_batch_global: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0 && c <= 7, "normalize_batch_global: bad size (1 <= c <= 7)");
    cudaStream_t s = as_stream(stream);
    batch_global_init_kernel<<<1, 32, 0, s>>>(d_work, c, (unsigned long long)batch * w * h);
    int rc = check_launch("normalize_batch_global(init)");
    if (rc != VACV_OK) return rc;
    if ((rc = vacv_cuda_sums_u8(src, batch, w, h, c, layout, d_work, 0, stream)) != VACV_OK) return rc;
    if (ev_sums_done) VACV_DCU(cudaEventRecord((cudaEvent_t)ev_sums_done, s), "normalize_batch_global");
    if ((rc = allreduce(ctx, d_work, 2 * c + 1, stream)) != VACV_OK) {
        if (rc > 0 || rc < VACV_ERR_CUDA) return set_error(VACV_ERR_CUDA, "normalize_batch_global: the all-reduce callback returned %d", rc);
        return rc;
    }
    batch_global_finalize_kernel<<<1, 32, 0, s>>>(d_work, c, d_mean_std);
    if ((rc = check_launch("normalize_batch_global(finalize)")) != VACV_OK) return rc;
    if (ev_stats_ready) VACV_DCU(cudaEventRecord((cudaEvent_t)ev_stats_ready, s), "normalize_batch_global");
    return vacv_cuda_normalize(src, dst, batch, w, h, c, VACV_INT8, layout, d_mean_std, d_mean_std + c, 0, stream);
}

extern "C" int vacv_cuda_normalize_batch_global_p2p(void* xchg, const uint8_t* src, float* dst, int batch, int w, int h, int c,
                                                    int layout, unsigned long long* d_work, float* d_mean_std,
                                                    void* ev_sums_done, void* ev_stats_ready, void* stream) {
    VACV_REQUIRE(xchg, "normalize_batch_global_p2p: null exchange");
    return vacv_cuda_normalize_batch_global_cb(p2p_allreduce_cb, xchg, src, dst, batch, w, h, c, layout, d_work, d_mean_std,
                                               ev_sums_done, ev_stats_ready, stream);
}

extern "C" int vacv_cuda_p2p_create(void** xchg, int nranks, int rank, void* h_handle) {
    VACV_REQUIRE(xchg && h_handle, "p2p_create: null pointer");
    VACV_REQUIRE(nranks >= 1 && nranks <= VACV_P2P_MAX_RANKS && rank >= 0 && rank < nranks, "p2p_create: bad rank %d of %d (max %d ranks)",
                 rank, nranks, VACV_P2P_MAX_RANKS);
    static_assert(sizeof(cudaIpcMemHandle_t) == VACV_P2P_HANDLE_BYTES, "IPC handle size");
    Xchg* x = new Xchg;
    x->nranks = nranks; x->rank = rank;
    const size_t bytes = xchg_words(nranks) * sizeof(unsigned long long);
    cudaError_t e = cudaGetDevice(&x->device);
    if (e == cudaSuccess) e = cudaMalloc(&x->mine, bytes);
    if (e == cudaSuccess) e = cudaMemset(x->mine, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // zeroed before any peer can learn the handle
    cudaIpcMemHandle_t hnd;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hnd, x->mine);
    if (e != cudaSuccess) {
        if (x->mine) cudaFree(x->mine);
        delete x;
        return set_error(VACV_ERR_CUDA, "p2p_create: %s", cudaGetErrorString(e));
    }
    memcpy(h_handle, &hnd, sizeof(hnd));
    x->peers.buf[rank] = x->mine;
    *xchg = x;
    return VACV_OK;
}

extern "C" int vacv_cuda_p2p_connect(void* xchg, const void* h_handles) {
    VACV_REQUIRE(xchg && h_handles, "p2p_connect: null pointer");
    Xchg* x = static_cast<Xchg*>(xchg);
    VACV_REQUIRE(!x->connected, "p2p_connect: already connected");
    int dev = -1;
    VACV_DCU(cudaGetDevice(&dev), "p2p_connect");
    VACV_REQUIRE(dev == x->device, "p2p_connect: device %d is current, the exchange lives on device %d", dev, x->device);
    for (int r = 0; r < x->nranks; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, static_cast<const char*>(h_handles) + (size_t)r * VACV_P2P_HANDLE_BYTES, sizeof(hnd));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q) if (q != x->rank && x->peers.buf[q]) { cudaIpcCloseMemHandle(x->peers.buf[q]); x->peers.buf[q] = nullptr; }
            return set_error(VACV_ERR_CUDA, "p2p_connect: opening rank %d's buffer: %s", r, cudaGetErrorString(e));
        }
        x->peers.buf[r] = static_cast<unsigned long long*>(p);
    }
    x->connected = true;
    return VACV_OK;
}

extern "C" int vacv_cuda_p2p_destroy(void* xchg) {
    if (!xchg) return VACV_OK;
    Xchg* x = static_cast<Xchg*>(xchg);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->nranks; ++r)
        if (r != x->rank && x->peers.buf[r]) cudaIpcCloseMemHandle(x->peers.buf[r]);
    cudaError_t e = cudaFree(x->mine);
    delete x;
    if (e != cudaSuccess) return set_error(VACV_ERR_CUDA, "p2p_destroy: %s", cudaGetErrorString(e));
    return VACV_OK;
}

extern "C" int vacv_cuda_p2p_allreduce_u64(void* xchg, unsigned long long* d_buf, int count, void* stream) {
    VACV_REQUIRE(xchg && d_buf, "p2p_allreduce_u64: null pointer");
    Xchg* x = static_cast<Xchg*>(xchg);
    VACV_REQUIRE(x->connected || x->nranks == 1, "p2p_allreduce_u64: exchange not connected");
    VACV_REQUIRE(count >= 1 && count < kXchgWords, "p2p_allreduce_u64: 1 <= count <= %d", kXchgWords - 1);
    p2p_allreduce_u64_kernel<<<1, 32, 0, as_stream(stream)>>>(x->peers, x->nranks, x->rank, d_buf, count);
    return check_launch("p2p_allreduce_u64");
}

extern "C" int vacv_cuda_p2p_status(void* xchg, int* timed_out) {
    VACV_REQUIRE(xchg && timed_out, "p2p_status: null pointer");
    Xchg* x = static_cast<Xchg*>(xchg);
    unsigned long long v = 0;
    VACV_DCU(cudaDeviceSynchronize(), "p2p_status");
    VACV_DCU(cudaMemcpy(&v, x->mine + (size_t)2 * x->nranks * kXchgWords + 1, sizeof(v), cudaMemcpyDeviceToHost), "p2p_status");
    *timed_out = (int)v;
    return VACV_OK;
}
