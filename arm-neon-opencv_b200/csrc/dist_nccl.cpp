// libvacv_dist.so: the NCCL transport of config 5 (include/vacv_dist.h).  Host code only -- the kernels are in
// libvacv_cuda.so (dist.cu, stats.cu, stream_ops.cu); this file contributes the one collective on the vacv path:
//     ncclAllReduce(2*c+1 x ncclUint64, ncclSum)   between the sums kernel and the normalise kernel, on the caller's stream.
// Reference semantics: Normalize::normalize_naive, auto-statistics branch (src/cv/normalize.cpp:84-121), over the batch.
#include <cuda_runtime_api.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>

#include "../../include/vacv_dist.h"

namespace {

int nccl_fail(const char* who, ncclResult_t r) {
    char msg[384];
    snprintf(msg, sizeof(msg), "%s: NCCL: %s", who, ncclGetErrorString(r));
    return vacv_cuda_set_last_error(VACV_ERR_CUDA, msg);
}

int nccl_allreduce_cb(void* comm, unsigned long long* d_buf, int count, void* stream) {
    return vacv_dist_allreduce_u64(comm, d_buf, count, stream);
}

}  // namespace

extern "C" int vacv_dist_allreduce_u64(void* nccl_comm, unsigned long long* d_buf, int count, void* stream) {
    if (!nccl_comm || !d_buf || count <= 0) return vacv_cuda_set_last_error(VACV_ERR_INVALID_ARG, "dist_allreduce_u64: null communicator / buffer or empty");
    const ncclResult_t r = ncclAllReduce(d_buf, d_buf, (size_t)count, ncclUint64, ncclSum, static_cast<ncclComm_t>(nccl_comm),
                                         static_cast<cudaStream_t>(stream));
    return r == ncclSuccess ? VACV_OK : nccl_fail("dist_allreduce_u64", r);
}

extern "C" int vacv_cuda_normalize_batch_global(void* nccl_comm, const uint8_t* src, float* dst, int batch, int w, int h, int c,
                                                int layout, unsigned long long* d_work, float* d_mean_std,
                                                void* ev_sums_done, void* ev_stats_ready, void* stream) {
    if (!nccl_comm) return vacv_cuda_set_last_error(VACV_ERR_INVALID_ARG, "normalize_batch_global: null NCCL communicator");
    return vacv_cuda_normalize_batch_global_cb(nccl_allreduce_cb, nccl_comm, src, dst, batch, w, h, c, layout, d_work, d_mean_std,
                                               ev_sums_done, ev_stats_ready, stream);
}

extern "C" int vacv_dist_nccl_version(int* version) {
    if (!version) return vacv_cuda_set_last_error(VACV_ERR_INVALID_ARG, "dist_nccl_version: null pointer");
    const ncclResult_t r = ncclGetVersion(version);
    return r == ncclSuccess ? VACV_OK : nccl_fail("dist_nccl_version", r);
}

extern "C" int vacv_dist_nccl_unique_id(void* h_id) {
    if (!h_id) return vacv_cuda_set_last_error(VACV_ERR_INVALID_ARG, "dist_nccl_unique_id: null pointer");
    static_assert(sizeof(ncclUniqueId) == VACV_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    const ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail("dist_nccl_unique_id", r);
    memcpy(h_id, &id, sizeof(id));
    return VACV_OK;
}

extern "C" int vacv_dist_nccl_comm_create(void** nccl_comm, int nranks, int rank, const void* h_id) {
    if (!nccl_comm || !h_id || nranks < 1 || rank < 0 || rank >= nranks)
        return vacv_cuda_set_last_error(VACV_ERR_INVALID_ARG, "dist_nccl_comm_create: bad arguments");
    ncclUniqueId id;
    memcpy(&id, h_id, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = ncclCommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return nccl_fail("dist_nccl_comm_create", r);
    *nccl_comm = comm;
    return VACV_OK;
}

extern "C" int vacv_dist_nccl_comm_destroy(void* nccl_comm) {
    if (!nccl_comm) return VACV_OK;
    const ncclResult_t r = ncclCommDestroy(static_cast<ncclComm_t>(nccl_comm));
    return r == ncclSuccess ? VACV_OK : nccl_fail("dist_nccl_comm_destroy", r);
}
