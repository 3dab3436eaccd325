/* vacv_dist.h -- the NCCL transport of config 5 (libvacv_dist.so = libvacv_cuda.so + NCCL).
 *
 * north_star: "Multi-GPU runs shard the frame batch across the 8 GPUs of one box with no communication, except a single
 * NCCL allreduce of per-GPU sum and sum-of-squares when mean_stddev is computed over the whole batch."  The reference has
 * no multi-device code (its only hook is CudaDevice::set_device, src/cv/cuda_device.cu:10-18); the operator semantics are
 * Normalize::normalize_naive's auto-statistics branch (src/cv/normalize.cpp:84-121) extended to the batch -- see
 * vacv_cuda_normalize_batch_global_cb in vacv_cuda.h for the exact definition and the scratch / event parameters.
 *
 * Plain C.  `nccl_comm` is the caller's ncclComm_t passed as void*; it must have been created by the libnccl.so.2 that is
 * loaded in the process (the helpers below call that same library).  One process per GPU.
 */
#ifndef VACV_DIST_H
#define VACV_DIST_H

#include "vacv_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VACV_NCCL_UNIQUE_ID_BYTES 128

/* sums_u8 -> ncclAllReduce(2*c+1 x ncclUint64, ncclSum) -> finalize -> normalize, all on `stream`. */
VACV_API int vacv_cuda_normalize_batch_global(void* nccl_comm, const uint8_t* src, float* dst, int batch, int w, int h, int c,
                                              int layout, unsigned long long* d_work, float* d_mean_std,
                                              void* ev_sums_done, void* ev_stats_ready, void* stream);
/* The exchange alone: in-place ncclAllReduce(count x ncclUint64, ncclSum) on `stream`. */
VACV_API int vacv_dist_allreduce_u64(void* nccl_comm, unsigned long long* d_buf, int count, void* stream);

/* Communicator helpers for callers that have no ncclComm_t yet (C / C++ hosts without their own NCCL setup):
 * rank 0 calls unique_id and ships the 128 bytes to the other ranks by any means; every rank then calls comm_create with the
 * CUDA device it will use already current (vacv_cuda_set_device). */
VACV_API int vacv_dist_nccl_version(int* version);
VACV_API int vacv_dist_nccl_unique_id(void* h_id);
VACV_API int vacv_dist_nccl_comm_create(void** nccl_comm, int nranks, int rank, const void* h_id);
VACV_API int vacv_dist_nccl_comm_destroy(void* nccl_comm);

#ifdef __cplusplus
}
#endif
#endif /* VACV_DIST_H */
