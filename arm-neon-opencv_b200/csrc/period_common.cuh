// Plumbing shared by the periodic walkers (resize_cubic3_period.cuh, resize_linear3_period.cuh): ring / staging constants, compile-time
// helpers, mbarrier and bulk-copy wrappers on 32-bit shared addresses.
#pragma once
#include <type_traits>

#include "vacv_common.cuh"

namespace vacv {

constexpr int kPdRing = 8, kPdAhead = 6;     // ring slots per warp / rows in flight ahead of the one being filtered
constexpr int kPdStageRows = 2;      // staging buffers per warp (one output row each)

namespace pd {

template <int N, int I = 0, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
}
__host__ __device__ constexpr int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
// window index of the first tap (source pixel sx - 1) of a thread's column c; the thread's window starts at source pixel P*KP*thread - 1
__host__ __device__ constexpr int tap0(int P, int Q, int c) { return floordiv((2 * c + 1) * P - Q, 2 * Q); }

// bytes G0 <= G1 <= G2 <= G3 (indices into the byte string of W[]) -> one word, with 0..3 PRMTs whose selectors are immediates
template <int G0, int G1, int G2, int G3, int N>
__device__ __forceinline__ uint32_t gather4(const uint32_t (&W)[N]) {
    constexpr int w0 = G0 >> 2, w1 = G1 >> 2, w2 = G2 >> 2, w3 = G3 >> 2;
    static_assert(w0 <= w1 && w1 <= w2 && w2 <= w3 && w3 < N, "byte indices must ascend inside W");
    constexpr int d0 = w0;
    constexpr int d1 = w1 != d0 ? w1 : w2 != d0 ? w2 : w3 != d0 ? w3 : d0;
    constexpr int d2 = (w2 != d0 && w2 != d1) ? w2 : (w3 != d0 && w3 != d1) ? w3 : -1;
    constexpr int d3 = (d2 >= 0 && w3 != d0 && w3 != d1 && w3 != d2) ? w3 : -1;
    constexpr auto nib1 = [](int G) constexpr { return (G >> 2) == d0 ? (G & 3) : (G >> 2) == d1 ? 4 + (G & 3) : 0; };
    constexpr unsigned sel1 = nib1(G0) | nib1(G1) << 4 | nib1(G2) << 8 | nib1(G3) << 12;
    uint32_t r;
    if constexpr (d1 == d0 && sel1 == 0x3210u) r = W[d0];
    else r = __byte_perm(W[d0], W[d1], sel1);
    if constexpr (d2 >= 0) {
        constexpr auto nib2 = [](int G, int j) constexpr { return (G >> 2) == d2 ? 4 + (G & 3) : j; };
        constexpr unsigned sel2 = nib2(G0, 0) | nib2(G1, 1) << 4 | nib2(G2, 2) << 8 | nib2(G3, 3) << 12;
        r = __byte_perm(r, W[d2], sel2);
    }
    if constexpr (d3 >= 0) {
        constexpr auto nib3 = [](int G, int j) constexpr { return (G >> 2) == d3 ? 4 + (G & 3) : j; };
        constexpr unsigned sel3 = nib3(G0, 0) | nib3(G1, 1) << 4 | nib3(G2, 2) << 8 | nib3(G3, 3) << 12;
        r = __byte_perm(r, W[d3], sel3);
    }
    return r;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace pd

}  // namespace vacv
