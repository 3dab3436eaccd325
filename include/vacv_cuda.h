/* vacv_cuda.h -- C-ABI of the B200-native vacv operator set (libvacv_cuda.so).
 *
 * This is the drop-in boundary beneath the reference's C++ API (reference: src/cv/cv.h:85-239 and
 * src/common/tensor.h:53-54).  The reference has no FFI of its own -- its dispatchers pick a backend at
 * compile time (`#if USE_NEON / USE_CUDA / else naive`, e.g. src/cv/resize.cpp:19-27) and call raw-pointer
 * kernels (e.g. src/cv/resize_naive.h:9-33).  Each entry point below replaces one of those raw-pointer
 * kernel families; the reference-side binding is shown in INTEGRATION.md and implemented by the C++
 * shim library libvacv.so (include/vacv/cv.h keeps the `va_cv::*` signatures).
 *
 * Conventions
 *   - plain C: pointers, ints, no C++/torch types.  Every function returns 0 on success, a negative
 *     vacv_status otherwise; vacv_cuda_last_error() gives the message (thread-local).
 *   - all data pointers are DEVICE pointers unless the parameter name starts with `h_`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All launches are asynchronous
 *     and stream-ordered; the operator entry points never synchronise or allocate (only the runtime
 *     helpers at the end of this header do).
 *   - tensors are dense, no row pitch, exactly like vision::Tensor (src/common/tensor.cpp:524);
 *     a batch is `batch` frames back to back.
 *   - dtype / layout codes are the reference's: vision::DType (tensor.h:12-18), vision::DLayout (:21-24).
 *   - u8 pixels have unsigned-char semantics (ARM ABI) unless VACV_FLAG_SIGNED_CHAR.
 */
#ifndef VACV_CUDA_H
#define VACV_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VACV_API
#else
#define VACV_API __attribute__((visibility("default")))
#endif

typedef enum {
    VACV_OK = 0,
    VACV_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, odd NV12 size ... */
    VACV_ERR_UNSUPPORTED = -2,   /* dtype/layout/interpolation combination the reference has no native path for */
    VACV_ERR_CUDA = -3           /* a CUDA runtime call or launch failed; message has the CUDA error string */
} vacv_status;

enum { VACV_FP32 = 0, VACV_FP16 = 1, VACV_INT8 = 2, VACV_FP64 = 3,        /* vision::DType (tensor.h:12-17) */
       VACV_BF16 = 16 };                                                   /* extension: output type of the fused pipeline only */
enum { VACV_NCHW = 0, VACV_NHWC = 1 };                                    /* vision::DLayout */
enum { VACV_INTER_LINEAR = 1, VACV_INTER_CUBIC = 2 };                     /* va_cv::VInterMode (cv.h:28-36) */
enum {
    VACV_FLAG_NONE = 0,
    VACV_FLAG_NEON_RULE = 1,     /* u8 bilinear with the rounding of resize_neon.cpp (aarch64 builds of the reference) */
    VACV_FLAG_SIGNED_CHAR = 2,   /* reproduce an x86 default-signed-char build of resize_naive/warp_affine_naive */
    VACV_FLAG_DIRECT_GATHER = 0x100, /* resize: force the direct global-gather kernels (default for bilinear) */
    VACV_FLAG_TILED = 0x200          /* resize: force the shared-memory tiled kernels (default for bicubic);
                                        warp_affine u8 BGR: run the TMA-staged kernel (default: direct gather) */
};

VACV_API int vacv_cuda_abi_version(void);
VACV_API const char* vacv_cuda_last_error(void);
/* for companion libraries (libvacv_dist.so) that report through the same thread-local message; returns `code` */
VACV_API int vacv_cuda_set_last_error(int code, const char* message);

/* ---- a1: CvtColor::nv_to_bgr_naive (src/cv/cvt_color.cpp:39-135) ------------------------------------------
 * src: batch x [Y plane w*h | interleaved chroma w*h/2];  dst: batch x (h x w x 3) BGR.  w, h even.
 * v_first = 1: chroma byte 0 is V (NV21 -- and what the reference does for BOTH codes 91 and 93);
 * v_first = 0: true NV12. */
VACV_API int vacv_cuda_cvt_nv2bgr(const uint8_t* src, uint8_t* dst, int batch, int w, int h, int v_first, void* stream);

/* ---- a2: Crop::crop_naive_{hwc_rgb,chw} (src/cv/crop.cpp:44-142) -------------------------------------------
 * ROI copy; dst = batch x (ch x cw x c) in the source layout.  dtype INT8 or FP32.  The rect must lie inside
 * the frame (the reference does not clip; here it is rejected with VACV_ERR_INVALID_ARG). */
VACV_API int vacv_cuda_crop(const void* src, void* dst, int batch, int w, int h, int c, int dtype, int layout,
                            int left, int top, int cw, int ch, void* stream);

/* ---- a3: Tensor::change_layout (src/common/tensor.cpp:393-457) --------------------------------------------
 * HWC <-> CHW permutation per frame; dtype INT8, FP16 (moved as 16-bit) or FP32. */
VACV_API int vacv_cuda_layout_change(const void* src, void* dst, int batch, int w, int h, int c, int dtype,
                                     int from_layout, int to_layout, void* stream);

/* ---- a4: Tensor::change_dtype (src/common/tensor.cpp:459-502) ---------------------------------------------
 * INT8->FP32 (unsigned) or FP32->INT8 (truncate toward zero, domain [0,256)); n = element count. */
VACV_API int vacv_cuda_dtype_change(const void* src, void* dst, size_t n, int from_dtype, int to_dtype, void* stream);

/* ---- a5-a9: Resize::resize_naive (src/cv/resize.cpp:42-100) -----------------------------------------------
 *   INTER_LINEAR INT8 : ResizeNaive::resize_naive_inter_linear_u8   (resize_naive.cpp:10-68)   bit-exact
 *                       (+VACV_FLAG_NEON_RULE: ResizeNeon::resize_neon_inter_linear_* resize_neon.cpp:12-347)
 *   INTER_LINEAR FP32 : resize_naive_inter_linear_fp32              (resize_naive.cpp:70-128)
 *   INTER_CUBIC  FP32 : resize_naive_inter_cubic_fp32_{hwc,chw}     (resize_naive.cpp:130-569), intended buffers
 *   INTER_CUBIC  INT8 : OpenCV 2.4.13 cv::resize (the reference's only path, resize.cpp:33-36), HWC only
 * Same-size input is copied (resize.cpp:58-61, full length). */
VACV_API int vacv_cuda_resize(const void* src, void* dst, int batch, int w, int h, int c, int dtype, int layout,
                              int w_out, int h_out, int interpolation, int flags, void* stream);

/* ---- a10: WarpAffine (src/cv/warp_affine.cpp:76-169, warp_affine_naive.cpp:9-106) --------------------------
 * Host helpers reproduce the reference's mixed float/double matrix arithmetic exactly. */
VACV_API void vacv_invert_affine(float h_m[6]);                                         /* warp_affine.cpp:121-133 */
VACV_API void vacv_rotation_matrix(float scale, float rot_deg, const double h_aux[4], float h_m[6]); /* :76-109 */
/* Crop i samples frame frame_idx[i] (frame i if frame_idx == NULL) of `frames` (n_frames dense frames) with the
 * INVERTED 2x3 matrix minv[6*i .. 6*i+5]; dst = n_crops x (h_out x w_out x c) in the source layout.
 * Destination pixels that map outside the source are written as 0 (== the reference with a zeroed dst). */
VACV_API int vacv_cuda_warp_affine(const void* frames, int n_frames, int w, int h, int c, int dtype, int layout,
                                   const int* frame_idx, const float* minv, int n_crops,
                                   void* dst, int w_out, int h_out, int flags, void* stream);

/* ---- a11: mean / stddev (src/cv/normalize_naive.cpp:7-72; exact-sum decision SURVEY App. C-4) -------------
 * Accumulates per-channel sum(x) and sum(x^2) of u8 pixels into sums[set][2*c] (u64; [2k]=Sx, [2k+1]=Sxx).
 * per_frame = 0: one set for the whole batch; 1: one set per frame.  The caller zeroes `sums`.  For a
 * batch-global statistic over several GPUs, all-reduce `sums` (sum, 2*c u64) between the two calls. */
VACV_API int vacv_cuda_sums_u8(const uint8_t* src, int batch, int w, int h, int c, int layout,
                               unsigned long long* sums, int per_frame, void* stream);
/* mean = Sx/n, stddev = sqrt(max(Sxx/n - mean^2, 0)) in fp64, stored fp32 (population stddev). */
VACV_API int vacv_cuda_finalize_mean_stddev(const unsigned long long* sums, int n_sets, int c,
                                            unsigned long long n_per_channel, float* mean, float* stddev, void* stream);

/* fp32 pixels (the reference computes statistics on the fp32 copy, normalize.cpp:92-108): sums in fp64, [2k]=Sx, [2k+1]=Sxx.
 * Exact -- and therefore identical to the u8 path -- whenever the data are integer-valued (pixels converted from u8);
 * for general fp32 data the cross-CTA summation order is not fixed (last-bit differences in fp64). */
VACV_API int vacv_cuda_sums_f32(const float* src, int batch, int w, int h, int c, int layout, double* sums, int per_frame, void* stream);
VACV_API int vacv_cuda_finalize_mean_stddev_f64(const double* sums, int n_sets, int c, unsigned long long n_per_channel,
                                                float* mean, float* stddev, void* stream);

/* ---- a12: NormalizeNaive::normalize_naive_{hwc_bgr,chw} (src/cv/normalize_naive.cpp:74-90) -----------------
 * dst = (float)((double)(x - mean[k]) / ((double)stddev[k] + 1e-6)), x converted from u8 first when
 * src_dtype == INT8 (normalize.cpp:92-95).  mean/stddev: c floats (stats_per_frame = 0) or batch x c. */
VACV_API int vacv_cuda_normalize(const void* src, float* dst, int batch, int w, int h, int c, int src_dtype, int layout,
                                 const float* mean, const float* stddev, int stats_per_frame, void* stream);

/* ---- a13: fused entry points (API: cv.h:154-201; semantics = composition) ----------------------
 * Config 2: nv->bgr -> resize INTER_LINEAR u8 -> u8->fp32 -> normalize -> HWC->CHW in ONE pass; intermediates
 * never touch HBM.  dst = batch x (3 x h_out x w_out) fp32 planes, bit-identical to the unfused chain. */
VACV_API int vacv_cuda_nv_resize_normalize_chw(const uint8_t* src, float* dst, int batch, int w, int h, int v_first,
                                               int w_out, int h_out, const float* mean, const float* stddev, void* stream);
/* ResizeNormalize::resize_normalize (resize_normalize.cpp:15-31): resize INTER_LINEAR u8 HWC -> fp32 normalised;
 * out_layout chooses HWC (reference) or CHW (fuses the change_layout that usually follows). */
VACV_API int vacv_cuda_resize_normalize(const uint8_t* src, float* dst, int batch, int w, int h, int c,
                                        int w_out, int h_out, const float* mean, const float* stddev,
                                        int out_layout, void* stream);
/* WarpAffineNormalize::warp_affine_normalize (warp_affine_normalize.cpp:13-45): warp u8 HWC -> fp32 normalised. */
VACV_API int vacv_cuda_warp_affine_normalize(const uint8_t* frames, int n_frames, int w, int h, int c,
                                             const int* frame_idx, const float* minv, int n_crops,
                                             float* dst, int w_out, int h_out, const float* mean, const float* stddev,
                                             int out_layout, void* stream);

/* ---- next rows: decoder surfaces in, fp32 / fp16 planes out -----------------------------------------
 * The fused pipeline on frames as a hardware decoder delivers them: row pitch, semi-planar (NV12 / NV21) or planar
 * (I420 = Y,U,V; YV12 = Y,V,U; enum slot COLOR_YUV2BGR_YV12 of cv.h:73 has no reference implementation) chroma, same
 * integer colour matrix (cvt_color.cpp:76-78).  Plane order inside a frame: Y (h rows of y_pitch bytes), then chroma
 * (semi-planar: h/2 rows of c_pitch; planar: two planes of h/2 rows of c_pitch each).  0 = dense defaults.
 * dst: batch x 3 x h_out x w_out planes of out_dtype VACV_FP32, VACV_FP16 or VACV_BF16 (16-bit = the fp32 result rounded to
 * nearest even).
 * Pitches, plane sizes and frame_stride that are multiples of 16 bytes run on the persistent TMA pipeline; anything else on
 * the tiled kernel (same results, register-staged loads). */
enum { VACV_YUV_NV21 = 0, VACV_YUV_NV12 = 1, VACV_YUV_I420 = 2, VACV_YUV_YV12 = 3 };
typedef struct {
    int format;            /* VACV_YUV_* */
    int w, h;              /* luma size, both even */
    int y_pitch, c_pitch;  /* bytes per luma / chroma row; 0 = dense (w, and w or w/2) */
    size_t frame_stride;   /* bytes between frames; 0 = dense */
} vacv_yuv_layout;
VACV_API int vacv_cuda_yuv_resize_normalize_chw(const uint8_t* src, const vacv_yuv_layout* layout, void* dst, int out_dtype,
                                                int batch, int w_out, int h_out, const float* mean, const float* stddev, void* stream);
/* The colour conversion alone on such surfaces (cvt_color.cpp:39-135 generalised; gives va_cv::cvt_color its
 * COLOR_YUV2BGR_YV12 case).  dst: batch dense HWC BGR frames. */
VACV_API int vacv_cuda_cvt_yuv2bgr(const uint8_t* src, const vacv_yuv_layout* layout, uint8_t* dst, int batch, void* stream);

/* Letterbox (no reference implementation -- the reference resizes without preserving aspect, 8d C2):
 * frame -> bilinear resize (the reference's rule, resize_naive.cpp:10-68) to content->w x content->h -> placed at
 * (content->x, content->y) of a canvas_w x canvas_h canvas filled with pad_bgr -> normalise -> 3 CHW planes of
 * out_dtype (VACV_FP32 / VACV_FP16 / VACV_BF16).  Bit-identical to running the unfused operators on the padded u8 canvas.
 * mean / stddev: device pointers (3 floats); mean_host / stddev_host: the same values in host memory (for the pad colour).
 * vacv_letterbox_rect computes the usual centred aspect-preserving content rectangle. */
typedef struct { int x, y, w, h; } vacv_rect;
VACV_API void vacv_letterbox_rect(int w, int h, int canvas_w, int canvas_h, vacv_rect* content);
VACV_API int vacv_cuda_yuv_letterbox_normalize_chw(const uint8_t* src, const vacv_yuv_layout* layout, void* dst, int out_dtype, int batch,
                                                   int canvas_w, int canvas_h, const vacv_rect* content, const uint8_t* pad_bgr,
                                                   const float* mean, const float* stddev, const float* mean_host, const float* stddev_host,
                                                   void* stream);

/* Host-buffer form of the two entries above (same chunked H2D / kernel / D2H pipeline as
 * vacv_cuda_nv_resize_normalize_chw_host below): h_src holds `batch` surfaces, h_dst receives batch x 3 x canvas_h x canvas_w
 * planes of out_dtype.  content == NULL: plain resize to the canvas; else letterbox (pad_bgr required).  Synchronous. */
VACV_API int vacv_cuda_yuv_normalize_chw_host(const uint8_t* h_src, const vacv_yuv_layout* layout, void* h_dst, int out_dtype, int batch,
                                              int canvas_w, int canvas_h, const vacv_rect* content, const uint8_t* pad_bgr,
                                              const float* h_mean, const float* h_stddev, int chunk_frames);

/* ---- config 5: batch-global statistics over a frame batch sharded across GPUs (north_star: "a single
 * allreduce of per-GPU sum and sum-of-squares") ------------------------------------------------------------------------
 * Semantics = Normalize::normalize_naive's auto-statistics branch (src/cv/normalize.cpp:84-121, :98-108) with the
 * statistic taken over EVERY frame of EVERY rank instead of one image: per-channel exact u64 sums of this rank's shard ->
 * one sum all-reduce of 2*c+1 u64 values ([2k] = Sx, [2k+1] = Sxx, [2c] = pixels per channel; 56 bytes for BGR, so ragged
 * shards need no side channel) -> mean / population stddev in fp64, stored fp32 (identical bits on every rank, independent of
 * the rank count) -> dst = (float)((double)(x - mean) / ((double)stddev + 1e-6)) over this rank's shard.
 * Everything is enqueued on `stream`; nothing synchronises or allocates.
 *   d_work      device scratch, >= 2*c+1 u64 (16-byte aligned), owned by the call while it is in flight
 *   d_mean_std  device, 2*c floats: receives mean[0..c) then stddev[0..c) (also the normalise kernel's input)
 *   ev_sums_done / ev_stats_ready   optional cudaEvent_t (NULL = none) recorded after the sums kernel and right before the
 *               normalise kernel: their distance is the cost of the exchange + finalize (bench.py reports it)
 * Three transports for the exchange:
 *   _cb    the caller's own all-reduce (MPI, a framework collective ...): allreduce(ctx, d_buf, count, stream) must sum
 *          `count` u64 values in place across ranks, stream-ordered, and return 0
 *   _p2p   this library's peer-memory exchange (below): every rank stores its 56 bytes straight into every peer's slot over
 *          NVLink / NVSwitch and polls its OWN memory -- one tiny kernel, no library call, CUDA-graph capturable
 *   NCCL   vacv_cuda_normalize_batch_global(ncclComm_t, ...) in libvacv_dist.so (include/vacv_dist.h), which keeps
 *          libvacv_cuda.so free of an NCCL dependency */
typedef int (*vacv_allreduce_u64_fn)(void* ctx, unsigned long long* d_buf, int count, void* stream);
VACV_API int vacv_cuda_normalize_batch_global_cb(vacv_allreduce_u64_fn allreduce, void* ctx, const uint8_t* src, float* dst,
                                                 int batch, int w, int h, int c, int layout, unsigned long long* d_work,
                                                 float* d_mean_std, void* ev_sums_done, void* ev_stats_ready, void* stream);

/* Peer-memory exchange between the ranks of ONE node (one process per GPU; several ranks may share a GPU).
 *   create   allocates this rank's slot buffer on the current device and returns its 64-byte IPC handle in h_handle
 *   connect  h_handles = the nranks handles in rank order (the caller all-gathers them with whatever it has: MPI,
 *            torch.distributed, a pipe); opens every peer's buffer (cudaIpcOpenMemHandle, peer access enabled lazily)
 *   allreduce_u64   in-place sum of count <= 15 u64 values across ranks, stream-ordered.  Collective: every rank must issue
 *            the same sequence of calls.  A rank that does not see all peers within ~20 s gives up and raises the status word
 *   status   0 = ok, else the number of timed-out exchanges (synchronises the device) */
#define VACV_P2P_HANDLE_BYTES 64
#define VACV_P2P_MAX_RANKS 16
VACV_API int vacv_cuda_p2p_create(void** xchg, int nranks, int rank, void* h_handle);
VACV_API int vacv_cuda_p2p_connect(void* xchg, const void* h_handles);
VACV_API int vacv_cuda_p2p_destroy(void* xchg);
VACV_API int vacv_cuda_p2p_allreduce_u64(void* xchg, unsigned long long* d_buf, int count, void* stream);
VACV_API int vacv_cuda_p2p_status(void* xchg, int* timed_out);
VACV_API int vacv_cuda_normalize_batch_global_p2p(void* xchg, const uint8_t* src, float* dst, int batch, int w, int h, int c,
                                                  int layout, unsigned long long* d_work, float* d_mean_std,
                                                  void* ev_sums_done, void* ev_stats_ready, void* stream);

/* ---- host-buffer entry point of the fused pipeline (the end-to-end path) ------------------------------------------
 * Same operation as vacv_cuda_nv_resize_normalize_chw, but `h_src` / `h_dst` / `h_mean` / `h_stddev` are HOST pointers
 * (pinned memory -- vacv_cuda_host_alloc -- for full PCIe speed; pageable works, slower).  The batch is cut into chunks of
 * `chunk_frames` frames that are pipelined over three internal streams (H2D copy | kernel | D2H copy, double-buffered
 * device staging, grown on demand and cached per host thread AND device: a thread that alternates vacv_cuda_set_device gets one
 * pipeline per GPU).  h_src must span (batch-1) frames + the defined bytes of the last one.  Synchronous: returns when h_dst is
 * complete; on an error nothing is left in flight on the caller's buffers. */
VACV_API int vacv_cuda_nv_resize_normalize_chw_host(const uint8_t* h_src, float* h_dst, int batch, int w, int h, int v_first,
                                                    int w_out, int h_out, const float* h_mean, const float* h_stddev,
                                                    int chunk_frames);

/* ---- runtime helpers: the host layer (libvacv.so) reaches CUDA only through this C-ABI -----------
 * Thin wrappers over the CUDA runtime so that C/C++ host code needs no CUDA headers: device memory, pinned staging
 * memory (the reference's USE_CUDA allocator, src/common/va_cuda_allocator.cu:8-34, made checkable), copies,
 * streams.  Copies are asynchronous on `stream`; vacv_cuda_stream_sync waits. */
/* Diagnostic switches ("WARP_GATHER", "NO_RPIPE", "PIPE_NCOL", ... -- DESIGN.md section 5): process-wide, initialised once from the
 * environment (VACV_<NAME>), changeable here at run time; the operator entry points never call getenv. */
VACV_API int vacv_cuda_set_tuning(const char* name, int value);
VACV_API int vacv_cuda_device_count(int* count);
VACV_API int vacv_cuda_set_device(int device);      /* the reference's hook: CudaDevice::set_device (src/cv/cuda_device.cu:15-18) */
VACV_API int vacv_cuda_get_device(int* device);
VACV_API int vacv_cuda_malloc(void** dptr, size_t bytes);
VACV_API int vacv_cuda_free(void* dptr);
VACV_API int vacv_cuda_host_alloc(void** h_ptr, size_t bytes);      /* pinned */
/* write_combined = 1: cudaHostAllocWriteCombined (the reference's USE_CUDA allocator flag, va_cuda_allocator.cu:28-30): fast for
 * the device to read, very slow for the CPU to read -- only for buffers the host fills once and never reads back */
VACV_API int vacv_cuda_host_alloc_flags(void** h_ptr, size_t bytes, int write_combined);
VACV_API int vacv_cuda_host_free(void* h_ptr);
VACV_API int vacv_cuda_memcpy_h2d(void* dptr, const void* h_ptr, size_t bytes, void* stream);
VACV_API int vacv_cuda_memcpy_d2h(void* h_ptr, const void* dptr, size_t bytes, void* stream);
/* `rows` rows of `row_bytes`, src_pitch apart in host memory -> dst_pitch apart on the device (ROI uploads: va_cv::crop) */
VACV_API int vacv_cuda_memcpy2d_h2d(void* dptr, size_t dst_pitch, const void* h_ptr, size_t src_pitch, size_t row_bytes, size_t rows, void* stream);
VACV_API int vacv_cuda_memset(void* dptr, int value, size_t bytes, void* stream);
VACV_API int vacv_cuda_stream_create(void** stream);
VACV_API int vacv_cuda_stream_destroy(void* stream);
VACV_API int vacv_cuda_stream_sync(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VACV_CUDA_H */
