// Word-granular tap fetch for interleaved 3-channel u8 images: the 6 contiguous bytes of two horizontally adjacent
// pixels at byte offset `a` of an image whose base is 4-byte aligned, from 2 (3 when a % 4 == 3) aligned 32-bit words.
#pragma once
#include "vacv_common.cuh"

namespace vacv {

__device__ __forceinline__ void linear_taps_u8c3(const uint8_t* __restrict__ img, unsigned a, uint32_t& b0, uint32_t& b1) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (a & ~3u));
    const int r = (int)(a & 3u), sh = r * 8;
    const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = r == 3 ? __ldg(wp + 2) : 0u;
    b0 = __funnelshift_r(w0, w1, sh);   // [L.b L.g L.r R.b]
    b1 = __funnelshift_r(w1, w2, sh);   // [R.g R.r  .   . ]
}



// Horizontal bilinear sums of the three channels of one tap row with two PRMT + three IDP.2A:
//   b0 = [L.b L.g L.r R.b], b1 = [R.g R.r . .] (from linear_taps_u8c3), cx = cx0 | cx1 << 16 (two 16-bit weights)
//   H[k] = L[k]*cx0 + R[k]*cx1.   kSigned: pixels are signed chars (App. C-1 compat), weights are non-negative either way.
template <bool kSigned>
__device__ __forceinline__ void hsum_u8c3(uint32_t b0, uint32_t b1, uint32_t cx, int (&H)[3]) {
    const uint32_t bg = __byte_perm(b0, b1, 0x4130);   // [L.b R.b L.g R.g]
    const uint32_t r = __byte_perm(b0, b1, 0x0052);    // [L.r R.r  .   . ]
    if (kSigned) {
        H[0] = __dp2a_lo((int)cx, (int)bg, 0); H[1] = __dp2a_hi((int)cx, (int)bg, 0); H[2] = __dp2a_lo((int)cx, (int)r, 0);
    } else {
        H[0] = (int)__dp2a_lo(cx, bg, 0u); H[1] = (int)__dp2a_hi(cx, bg, 0u); H[2] = (int)__dp2a_lo(cx, r, 0u);
    }
}

}  // namespace vacv
