/* TEST INFRASTRUCTURE ONLY -- scalar emulation of the handful of Arm NEON intrinsics that the reference's
 * src/cv/resize_neon.cpp uses, so that this file (aarch64-only in the reference: `#if defined(USE_NEON) and __ARM_NEON`)
 * can be compiled UNMODIFIED on the x86 build host and its results used to pin the oracle's restatement of the NEON
 * bilinear rule (orc_resize_linear_u8_neon_rule).
 *
 * Semantics follow the Arm C Language Extensions / Armv8-A ISA:
 *   vld1_s16       load 4 x int16                      vdup_n_s16 / vdupq_n_s32   broadcast
 *   vmull_s16      4 x (int16 * int16) -> int32        vsraq_n_s32(a, b, n)       a + (b >> n)   (arithmetic shift)
 *   vshrn_n_s32    (int32 >> n) truncated to int16     vcombine_s16               concatenate two 4-lane halves
 *   vqmovun_s16    signed int16 -> unsigned 8-bit with saturation to [0, 255]
 *   vst1_u8        store 8 x uint8
 * Nothing else is provided on purpose: a new intrinsic in the reference fails the build loudly. */
#ifndef VACV_ORACLE_NEON_EMUL_H
#define VACV_ORACLE_NEON_EMUL_H
#include <stdint.h>

struct int16x4_t { int16_t v[4]; };
struct int16x8_t { int16_t v[8]; };
struct int32x4_t { int32_t v[4]; };
struct uint8x8_t { uint8_t v[8]; };

static inline int16x4_t vld1_s16(const int16_t* p) { int16x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = p[i]; return r; }
static inline int16x4_t vdup_n_s16(int16_t x) { int16x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = x; return r; }
static inline int32x4_t vdupq_n_s32(int32_t x) { int32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = x; return r; }
static inline int32x4_t vmull_s16(int16x4_t a, int16x4_t b) {
    int32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (int32_t)a.v[i] * (int32_t)b.v[i]; return r;
}
#define vsraq_n_s32(a, b, n) vacv_emul_vsraq_n_s32((a), (b), (n))
static inline int32x4_t vacv_emul_vsraq_n_s32(int32x4_t a, int32x4_t b, int n) {
    int32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (int32_t)((uint32_t)a.v[i] + (uint32_t)(b.v[i] >> n)); return r;   /* wraps like the hardware */
}
#define vshrn_n_s32(a, n) vacv_emul_vshrn_n_s32((a), (n))
static inline int16x4_t vacv_emul_vshrn_n_s32(int32x4_t a, int n) {
    int16x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (int16_t)(a.v[i] >> n); return r;   /* narrowing: low 16 bits */
}
static inline int16x8_t vcombine_s16(int16x4_t lo, int16x4_t hi) {
    int16x8_t r; for (int i = 0; i < 4; ++i) { r.v[i] = lo.v[i]; r.v[4 + i] = hi.v[i]; } return r;
}
static inline uint8x8_t vqmovun_s16(int16x8_t a) {
    uint8x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint8_t)(a.v[i] < 0 ? 0 : a.v[i] > 255 ? 255 : a.v[i]); return r;
}
static inline void vst1_u8(uint8_t* p, uint8x8_t a) { for (int i = 0; i < 8; ++i) p[i] = a.v[i]; }
#endif
