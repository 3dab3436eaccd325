// Word-granular tap fetch for interleaved 3-channel u8 images: the 6 contiguous bytes of two horizontally adjacent
// pixels at byte offset `a` of an image whose base is 4-byte aligned, from 2 (3 when a % 4 == 3) aligned 32-bit words.
#pragma once
#include "vacv_common.cuh"

namespace vacv {

__device__ __forceinline__ void linear_taps_u8c3(const uint8_t* __restrict__ img, size_t a, uint32_t& b0, uint32_t& b1) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (a & ~(size_t)3));
    const int r = (int)(a & 3), sh = r * 8;
    const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = r == 3 ? __ldg(wp + 2) : 0u;
    b0 = __funnelshift_r(w0, w1, sh);   // [L.b L.g L.r R.b]
    b1 = __funnelshift_r(w1, w2, sh);   // [R.g R.r  .   . ]
}


}  // namespace vacv
