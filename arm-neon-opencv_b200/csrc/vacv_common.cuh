// Shared device/host helpers for the vacv sm_100a kernels.
//
// Build contract: every .cu here is compiled with --fmad=false.  The reference's oracle build has no FMA
// (SURVEY 8c), so every float expression must be evaluated operation by operation; with contraction off a
// plain `a*b+c` is two IEEE roundings, exactly as on the x86 reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vacv_cuda.h"

namespace vacv {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
// resize.cu: persistent TMA kernel for u8 BGR bilinear (resize_pipe_u8c3.cuh); out_mode 0 = u8 BGR, 1 = normalised fp32 CHW planes,
// 2 = normalised fp32 HWC.  Returns 1 = launched, 0 = shape not eligible, < 0 = error.
int try_launch_resize_pipe_u8c3(const uint8_t* src, void* dst, int images, int w, int h, int wo, int ho, bool signed_char, int out_mode,
                                const float* mean, const float* stddev, cudaStream_t s, int c = 3);   // c = 1: single planes (u8 output only)

// resize_linear_period.cu: periodic walker for u8 BGR bilinear at rational horizontal scales (resize_linear3_period.cuh).
// Returns 1 = launched, 0 = shape not eligible, < 0 = error.
int try_launch_resize_linear3_period(const uint8_t* src, uint8_t* dst, int images, int w, int h, int wo, int ho, bool signed_char, cudaStream_t s);
int try_launch_resize_linear1_period(const uint8_t* src, uint8_t* dst, int planes, int w, int h, int wo, int ho, bool signed_char, cudaStream_t s);   // single planes

#define VACV_REQUIRE(cond, ...)                                                 \
    do {                                                                        \
        if (!(cond)) return ::vacv::set_error(VACV_ERR_INVALID_ARG, __VA_ARGS__); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int elem_size(int dtype) {
    return dtype == VACV_INT8 ? 1 : dtype == VACV_FP16 ? 2 : dtype == VACV_FP32 ? 4 : dtype == VACV_FP64 ? 8 : 0;
}
static inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

constexpr int kNumSMs = 148;   // B200

// ---------------------------------------------------------------------------------------------------
// memory access: 128-bit streaming loads/stores.  Inputs are read once -> bypass L1 allocation;
// outputs are write-once -> streaming (evict-first) stores.
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream4(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream16f(void* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream4(void* p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream4f(void* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// arithmetic shared by several operators

// SATURATE_CAST_SHORT (src/common/macro.h:25-30): fp32 round-half-away, truncating int cast, clamp to s16.
__device__ __forceinline__ int sat_short(float x) {
    int v = (int)(x + (x >= 0.f ? 0.5f : -0.5f));
    return max(min(v, 32767), -32768);
}

// Bilinear source coordinate (resize_naive.cpp:21-32,38-50): the product is formed in double from an fp32
// (NAIVE rule) or fp64 (NEON rule / cubic) scale, rounded to fp32, floored, edge-clamped.
__device__ __forceinline__ void linear_coord(int d, double scale, int n_in, int& s, float& f) {
    float fx = (float)(((double)d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= n_in - 1) { sx = n_in - 2; fx = 1.f; }
    s = sx; f = fx;
}

// u8 pixel fetch with the reference's `char` typing (App. C-1).
template <bool kSigned>
__device__ __forceinline__ int pix(uint8_t v) { return kSigned ? (int)(int8_t)v : (int)v; }

// (float)((double)(x - mean) / ((double)std + 1e-6))  -- normalize_naive.cpp:76-78
__device__ __forceinline__ float normalize_one(float x, float mean, double den) {
    return (float)((double)(x - mean) / den);
}

// YUV -> BGR chroma terms (cvt_color.cpp:76-78)
struct ChromaTerms { int ra, ga, ba; };
__device__ __forceinline__ ChromaTerms chroma_terms(int v, int u) {
    v -= 128; u -= 128;
    ChromaTerms t;
    t.ra = (179 * v) >> 7;
    t.ga = (44 * u + 91 * v) >> 7;
    t.ba = (227 * u) >> 7;
    return t;
}
__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }
// clamp(a + b, 0, 255) in one DPX instruction (SASS: VIADDMNMX.RELU)
__device__ __forceinline__ int add_clamp255(int a, int b) { return __viaddmin_s32_relu(a, b, 255); }

}  // namespace vacv
