// Config 5: batch-global mean / stddev + normalise over a frame batch sharded across GPUs.
//
// The reference computes the statistic of ONE image inside Normalize::normalize_naive (src/cv/normalize.cpp:98-108);
// here the same statistic is taken over every frame of every rank.  Because stats.cu carries it as exact u64 sums, the
// only exchange on the whole vacv path is a sum all-reduce of 2*c+1 integers between the reduction kernel and the
// normalise kernel -- 56 bytes: pure latency, no bandwidth.  Two transports live here (NCCL is in dist_nccl.cpp):
//
//  * callback -- the caller's all-reduce;
//  * peer memory -- every rank owns a small slot buffer that all peers map (CUDA IPC).  One 1-CTA kernel per exchange: a
//    thread STORES one 8-byte word {32 payload bits, 32-bit epoch} straight into a peer's buffer over NVLink / NVSwitch (fire
//    and forget; the word validates itself, so there is no fence and no flag) and then polls the matching word that peer
//    stored into THIS rank's buffer (local L2, nobody spins over the fabric); the slots are summed in rank order and, for
//    config 5, mean / stddev are finalised in the same kernel.  No library call, no host involvement: the epoch counter lives
//    on the device, so the kernel replays unchanged inside a CUDA graph.  Two slot sets alternate by epoch parity: a peer can
//    only be one exchange ahead (it needs MY words of epoch e+1 before it can start e+2, and I send those only after my
//    epoch-e kernel has finished reading).
#include <algorithm>
#include <cstring>

#include "vacv_common.cuh"

namespace vacv {

constexpr int kXchgMaxCount = 15;            // u64 values per exchange
constexpr int kXchgWords = 2 * 16;           // 8-byte words per (parity, rank) slot: one per 32-bit half of a value

struct XchgPeers { unsigned long long* buf[VACV_P2P_MAX_RANKS]; };

// layout of a rank's buffer (8-byte words): [2][nranks][32] slots, then {epoch, timeouts}
__device__ __forceinline__ size_t slot_ofs(int parity, int nranks, int r) { return ((size_t)parity * nranks + r) * kXchgWords; }
static inline size_t xchg_words(int nranks) { return (size_t)2 * nranks * kXchgWords + 2; }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The exchange, "low latency" style: every 8-byte word on the wire carries 32 bits of payload AND the 32-bit epoch, so a word is
// valid the moment its epoch matches -- 8-byte stores are single transactions, no fence and no separate flag are needed, and
// the cost is one store, its flight over NVLink, and one poll.  Thread i handles (peer r, half-word j): it stores its own half j
// into ITS slot of peer r's buffer and then polls the half that peer r stored into this rank's buffer (local memory).
// mean_std != nullptr: the values are config 5's {Sx, Sxx per channel, pixel count}; the statistics are finalised right here
// (one launch less on the critical path between the sums kernel and the normalise kernel).
__global__ void __launch_bounds__(512) p2p_allreduce_u64_kernel(XchgPeers peers, int nranks, int rank, unsigned long long* data, int count,
                                                                 float* mean_std, int c) {
    __shared__ unsigned halves[VACV_P2P_MAX_RANKS][2 * kXchgMaxCount];
    __shared__ unsigned s_epoch;
    __shared__ int s_bad;
    unsigned long long* mine = peers.buf[rank];
    unsigned long long* ctl = mine + (size_t)2 * nranks * kXchgWords;
    if (threadIdx.x == 0) {
        unsigned e = (unsigned)ctl[0] + 1;
        if (e == 0) e = 1;                       // 0 is what fresh (zeroed) slots hold
        ctl[0] = e;
        s_epoch = e;
        s_bad = 0;
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int parity = (int)(epoch & 1);
    const int nhalf = 2 * count;
    for (int i = threadIdx.x; i < nranks * nhalf; i += blockDim.x) {
        const int r = i / nhalf, j = i - r * nhalf;
        const unsigned half = (unsigned)(data[j >> 1] >> (32 * (j & 1)));
        st_volatile_u64(peers.buf[r] + slot_ofs(parity, nranks, rank) + j, ((unsigned long long)epoch << 32) | half);
        const unsigned long long* in = mine + slot_ofs(parity, nranks, r) + j;
        unsigned long long v = ld_volatile_u64(in);
        if ((unsigned)(v >> 32) != epoch) {
            const unsigned long long t0 = global_timer_ns();
            while ((unsigned)((v = ld_volatile_u64(in)) >> 32) != epoch) {
                if (global_timer_ns() - t0 > 20000000000ull) { s_bad = 1; break; }
            }
        }
        halves[r][j] = (unsigned)v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_bad) ctl[1] += 1;
    __shared__ unsigned long long total[kXchgMaxCount];
    if ((int)threadIdx.x < count) {
        unsigned long long v = 0;
        for (int r = 0; r < nranks; ++r) v += ((unsigned long long)halves[r][2 * threadIdx.x + 1] << 32) | halves[r][2 * threadIdx.x];
        data[threadIdx.x] = v;
        total[threadIdx.x] = v;
    }
    if (mean_std) {
        __syncthreads();
        const int k = threadIdx.x;
        if (k < c) {
            const double n = (double)total[2 * c];
            const double m = (double)total[2 * k] / n;
            double var = (double)total[2 * k + 1] / n - m * m;
            if (var < 0) var = 0;
            mean_std[k] = (float)m;
            mean_std[c + k] = (float)sqrt(var);
        }
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

static int launch_p2p(void* xchg, unsigned long long* d_buf, int count, float* mean_std, int c, void* stream, const char* who) {
    Xchg* x = static_cast<Xchg*>(xchg);
    VACV_REQUIRE(x->connected || x->nranks == 1, "%s: exchange not connected", who);
    VACV_REQUIRE(count >= 1 && count <= kXchgMaxCount, "%s: 1 <= count <= %d", who, kXchgMaxCount);
    const int items = x->nranks * 2 * count;
    const int threads = std::min(512, (items + 31) & ~31);
    p2p_allreduce_u64_kernel<<<1, std::max(threads, 32), 0, as_stream(stream)>>>(x->peers, x->nranks, x->rank, d_buf, count, mean_std, c);
    return check_launch(who);
}

// The peer-memory form of config 5 has its own sequence: the exchange kernel also finalises mean / stddev, so only ONE tiny kernel
// sits between the sums kernel and the normalise kernel.
extern "C" int vacv_cuda_normalize_batch_global_p2p(void* xchg, const uint8_t* src, float* dst, int batch, int w, int h, int c,
                                                    int layout, unsigned long long* d_work, float* d_mean_std,
                                                    void* ev_sums_done, void* ev_stats_ready, void* stream) {
    VACV_REQUIRE(xchg && src && dst && d_work && d_mean_std, "normalize_batch_global_p2p: null pointer");
    VACV_REQUIRE(batch > 0 && w > 0 && h > 0 && c > 0 && c <= 7, "normalize_batch_global_p2p: bad size (1 <= c <= 7)");
    cudaStream_t s = as_stream(stream);
    batch_global_init_kernel<<<1, 32, 0, s>>>(d_work, c, (unsigned long long)batch * w * h);
    int rc = check_launch("normalize_batch_global_p2p(init)");
    if (rc != VACV_OK) return rc;
    if ((rc = vacv_cuda_sums_u8(src, batch, w, h, c, layout, d_work, 0, stream)) != VACV_OK) return rc;
    if (ev_sums_done) VACV_DCU(cudaEventRecord((cudaEvent_t)ev_sums_done, s), "normalize_batch_global_p2p");
    if ((rc = launch_p2p(xchg, d_work, 2 * c + 1, d_mean_std, c, stream, "normalize_batch_global_p2p")) != VACV_OK) return rc;
    if (ev_stats_ready) VACV_DCU(cudaEventRecord((cudaEvent_t)ev_stats_ready, s), "normalize_batch_global_p2p");
    return vacv_cuda_normalize(src, dst, batch, w, h, c, VACV_INT8, layout, d_mean_std, d_mean_std + c, 0, stream);
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
    return launch_p2p(xchg, d_buf, count, nullptr, 0, stream, "p2p_allreduce_u64");
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
