// Host helper shared by the TMA-based operators: cuTensorMapEncodeTiled resolved at run time through the CUDA runtime
// (cudaGetDriverEntryPoint), so the library links against libcudart only.
#pragma once
#include <cuda.h>   // CUtensorMap and enums (types only)
#include <cuda_runtime.h>

namespace vacv {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// rank-3 tiled map without swizzle / interleave; strides in bytes for dimensions 1 and 2
inline bool encode_map_3d(CUtensorMap* map, CUtensorMapDataType type, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2,
                          cuuint64_t stride1, cuuint64_t stride2, cuuint32_t b0, cuuint32_t b1, cuuint32_t b2) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {stride1, stride2};
    const cuuint32_t box[3] = {b0, b1, b2};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, type, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace vacv
