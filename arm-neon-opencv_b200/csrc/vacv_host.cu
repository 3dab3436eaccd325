// Host side of the C-ABI: error reporting and the matrix helpers of warp_affine, which the reference evaluates on
// the host in mixed float/double (src/cv/warp_affine.cpp:76-133).  Built with -ffp-contract=off for baseline
// x86-64, like the reference (no FMA), so the results are bit-identical to its build.
#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "vacv_common.cuh"

namespace vacv {

static thread_local char g_err[512] = "";

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
extern "C" const char* vacv_cuda_last_error(void) { return vacv::g_err; }

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
extern "C" int vacv_cuda_malloc(void** dptr, size_t bytes) {
    VACV_REQUIRE(dptr, "malloc: null pointer");
    VACV_RT(cudaMalloc(dptr, bytes), "malloc");
}
extern "C" int vacv_cuda_free(void* dptr) { VACV_RT(cudaFree(dptr), "free"); }
extern "C" int vacv_cuda_host_alloc(void** h_ptr, size_t bytes) {
    VACV_REQUIRE(h_ptr, "host_alloc: null pointer");
    VACV_RT(cudaHostAlloc(h_ptr, bytes, cudaHostAllocDefault), "host_alloc");
}
extern "C" int vacv_cuda_host_free(void* h_ptr) { VACV_RT(cudaFreeHost(h_ptr), "host_free"); }
extern "C" int vacv_cuda_memcpy_h2d(void* dptr, const void* h_ptr, size_t bytes, void* stream) {
    VACV_RT(cudaMemcpyAsync(dptr, h_ptr, bytes, cudaMemcpyHostToDevice, vacv::as_stream(stream)), "memcpy_h2d");
}
extern "C" int vacv_cuda_memcpy_d2h(void* h_ptr, const void* dptr, size_t bytes, void* stream) {
    VACV_RT(cudaMemcpyAsync(h_ptr, dptr, bytes, cudaMemcpyDeviceToHost, vacv::as_stream(stream)), "memcpy_d2h");
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
