/* TEST INFRASTRUCTURE ONLY -- see vacv_oracle.h.  Build: gcc -O2 -ffp-contract=off (no FMA contraction,
 * no fast-math), so every float expression below is evaluated operation by operation in IEEE
 * binary32/binary64, exactly as the reference's -O3 baseline-x86-64 build does. */
#include "vacv_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* macro.h:25-30 SATURATE_CAST_SHORT: round half away from zero in fp32, then clamp to int16. */
static inline short sat_short(float x) {
    int v = (int)(x + (x >= 0.f ? 0.5f : -0.5f));
    if (v < SHRT_MIN) v = SHRT_MIN;
    if (v > SHRT_MAX) v = SHRT_MAX;
    return (short)v;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int pix(const uint8_t* p, int signed_char) { return signed_char ? (int)(int8_t)*p : (int)*p; }

/* ------------------------------------------------------------------ colour ------------------------ */
/* cvt_color.cpp:58-131: 2x2 quads share one chroma pair; integer BT.601 (JPEG range), arithmetic >>7. */
void orc_nv_to_bgr(const uint8_t* src, int w, int h, int v_first, uint8_t* dst) {
    const uint8_t* yp = src;
    const uint8_t* cp = src + (size_t)w * h;
    const int vi = v_first ? 0 : 1, ui = v_first ? 1 : 0;   /* cvt_color.cpp:72-73 (x_num,y_num) */
    for (int r = 0; r < h; ++r) {
        const uint8_t* crow = cp + (size_t)(r / 2) * w;
        for (int x = 0; x < w; ++x) {
            int v = crow[(x & ~1) + vi] - 128, u = crow[(x & ~1) + ui] - 128;
            int ra = (179 * v) >> 7;                          /* :76 */
            int ga = (44 * u + 91 * v) >> 7;                  /* :77 */
            int ba = (227 * u) >> 7;                          /* :78 */
            int Y = yp[(size_t)r * w + x];
            uint8_t* o = dst + ((size_t)r * w + x) * 3;
            o[0] = (uint8_t)clampi(Y + ba, 0, 255);           /* B :82,:100 */
            o[1] = (uint8_t)clampi(Y - ga, 0, 255);           /* G */
            o[2] = (uint8_t)clampi(Y + ra, 0, 255);           /* R */
        }
    }
}

/* Next-row extension: the same colour matrix on pitched / planar surfaces.  format 0 = NV21 (VU pairs),
 * 1 = NV12 (UV pairs), 2 = I420 (Y,U,V planes), 3 = YV12 (Y,V,U planes).  Plane order: Y (h x y_pitch), chroma.
 * Done by re-packing to a dense VU surface and running the pinned routine above. */
void orc_yuv_to_bgr(const uint8_t* src, int format, int w, int h, int y_pitch, int c_pitch, uint8_t* dst) {
    uint8_t* dense = (uint8_t*)malloc((size_t)w * h * 3 / 2);
    const uint8_t* cbase = src + (size_t)y_pitch * h;
    for (int r = 0; r < h; ++r) memcpy(dense + (size_t)r * w, src + (size_t)r * y_pitch, (size_t)w);
    for (int r = 0; r < h / 2; ++r) {
        uint8_t* o = dense + (size_t)w * h + (size_t)r * w;
        if (format <= 1) {
            const uint8_t* c = cbase + (size_t)r * c_pitch;
            for (int x = 0; x < w; x += 2) { o[x] = c[x + (format == 1)]; o[x + 1] = c[x + (format != 1)]; }
        } else {
            const uint8_t* p0 = cbase + (size_t)r * c_pitch;                            /* first plane  */
            const uint8_t* p1 = cbase + (size_t)c_pitch * (h / 2) + (size_t)r * c_pitch; /* second plane */
            const uint8_t* up = format == 2 ? p0 : p1;
            const uint8_t* vp = format == 2 ? p1 : p0;
            for (int x = 0; x < w; x += 2) { o[x] = vp[x / 2]; o[x + 1] = up[x / 2]; }
        }
    }
    orc_nv_to_bgr(dense, w, h, 1, dst);
    free(dense);
}

/* image_util.cpp:9-40.  NB the U/V expressions wrap through unsigned int (no clamp). */
void orc_bgr_to_nv21(const uint8_t* src, int w, int h, uint8_t* dst) {
    uint8_t* yp = dst;
    uint8_t* vu = dst + (size_t)w * h;
    const unsigned shift = 14, off = 128u << 14;
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c, src += 3) {
            int Y = (int)((unsigned)(src[0] * 1868u + src[1] * 9617u + src[2] * 4899u) >> shift);
            *yp++ = (uint8_t)Y;
            if (r % 2 == 0 && c % 2 == 0) {
                int U = (int)((unsigned)((src[0] - Y) * 9241u + off) >> shift);
                int V = (int)((unsigned)((src[2] - Y) * 11682u + off) >> shift);
                vu[0] = (uint8_t)V;
                vu[1] = (uint8_t)U;
                vu += 2;
            }
        }
}

/* ------------------------------------------------------------------ copies ------------------------ */
void orc_crop(const void* src_, int w, int h, int c, int elem, int layout,
              int left, int top, int cw, int ch, void* dst_) {
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dst = (uint8_t*)dst_;
    if (layout == 1) { /* crop.cpp:84-125 */
        size_t px = (size_t)c * elem;
        for (int y = 0; y < ch; ++y)
            memcpy(dst + (size_t)y * cw * px, src + ((size_t)(top + y) * w + left) * px, (size_t)cw * px);
    } else {           /* crop.cpp:44-82 */
        for (int k = 0; k < c; ++k)
            for (int y = 0; y < ch; ++y)
                memcpy(dst + ((size_t)k * cw * ch + (size_t)y * cw) * elem,
                       src + ((size_t)k * w * h + (size_t)(top + y) * w + left) * elem, (size_t)cw * elem);
    }
}

void orc_hwc_to_chw(const void* src_, int w, int h, int c, int elem, void* dst_) { /* tensor.cpp:160-170 */
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dst = (uint8_t*)dst_;
    size_t n = (size_t)w * h;
    for (int k = 0; k < c; ++k)
        for (size_t j = 0; j < n; ++j) memcpy(dst + (k * n + j) * elem, src + (j * c + k) * elem, elem);
}
void orc_chw_to_hwc(const void* src_, int w, int h, int c, int elem, void* dst_) { /* tensor.cpp:172-182 */
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dst = (uint8_t*)dst_;
    size_t n = (size_t)w * h;
    for (size_t j = 0; j < n; ++j)
        for (int k = 0; k < c; ++k) memcpy(dst + (j * c + k) * elem, src + (k * n + j) * elem, elem);
}

void orc_u8_to_f32(const uint8_t* src, size_t n, float* dst) { /* tensor.cpp:477-481 */
    for (size_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}
void orc_f32_to_u8(const float* src, size_t n, uint8_t* dst) { /* tensor.cpp:488-492: static_cast<char> = cvttss2si, low byte */
    for (size_t i = 0; i < n; ++i) dst[i] = (uint8_t)(int)src[i];
}

/* ------------------------------------------------------------------ bilinear resize --------------- */
/* resize_naive.cpp:21-32 / :38-50: source index + fraction with edge clamp. */
static inline void lin_coord(int d, float scale, int n_in, int* s, float* f) {
    float fx = (float)((d + 0.5) * scale - 0.5);   /* double arithmetic on a float scale, then rounded */
    int sx = (int)floor(fx);
    fx -= sx;
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= n_in - 1) { sx = n_in - 2; fx = 1.f; }
    *s = sx; *f = fx;
}

static void resize_linear_u8_plane(const uint8_t* src, int w, int h, int c, uint8_t* dst, int wo, int ho, int sc) {
    float scale_x = (float)w / wo, scale_y = (float)h / ho;   /* :17-18 fp32 division */
    for (int dy = 0; dy < ho; ++dy) {
        int sy; float fy;
        lin_coord(dy, scale_y, h, &sy, &fy);
        short cy0 = sat_short((1.f - fy) * 2048), cy1 = sat_short(2048 * fy);   /* :35-36 (independently rounded) */
        for (int dx = 0; dx < wo; ++dx) {
            int sx; float fx;
            lin_coord(dx, scale_x, w, &sx, &fx);
            short cx0 = sat_short((1.f - fx) * 2048), cx1 = sat_short(2048 * fx);
            const uint8_t* lt = src + ((size_t)sy * w + sx) * c;
            const uint8_t* lb = lt + (size_t)w * c;
            uint8_t* o = dst + ((size_t)dy * wo + dx) * c;
            for (int k = 0; k < c; ++k) {   /* :60-65 int32 MAC, truncating >>22, low byte stored */
                int v = (pix(lt + k, sc) * cx0 * cy0 + pix(lb + k, sc) * cx0 * cy1 +
                         pix(lt + c + k, sc) * cx1 * cy0 + pix(lb + c + k, sc) * cx1 * cy1) >> 22;
                o[k] = (uint8_t)v;
            }
        }
    }
}
void orc_resize_linear_u8(const uint8_t* src, int w, int h, int c, int layout, uint8_t* dst, int wo, int ho, int sc) {
    if (layout == 1) resize_linear_u8_plane(src, w, h, c, dst, wo, ho, sc);
    else for (int k = 0; k < c; ++k)   /* resize.cpp:73-87 */
        resize_linear_u8_plane(src + (size_t)w * h * k, w, h, 1, dst + (size_t)wo * ho * k, wo, ho, sc);
}

static void resize_linear_f32_plane(const float* src, int w, int h, int c, float* dst, int wo, int ho) {
    float scale_x = (float)w / wo, scale_y = (float)h / ho;
    for (int dy = 0; dy < ho; ++dy) {
        int sy; float fy;
        lin_coord(dy, scale_y, h, &sy, &fy);
        float cy0 = 1.f - fy, cy1 = fy;
        for (int dx = 0; dx < wo; ++dx) {
            int sx; float fx;
            lin_coord(dx, scale_x, w, &sx, &fx);
            float cx0 = 1.f - fx, cx1 = fx;
            const float* lt = src + ((size_t)sy * w + sx) * c;
            const float* lb = lt + (size_t)w * c;
            float* o = dst + ((size_t)dy * wo + dx) * c;
            for (int k = 0; k < c; ++k)   /* resize_naive.cpp:121-124 order */
                o[k] = lt[k] * cx0 * cy0 + lb[k] * cx0 * cy1 + lt[c + k] * cx1 * cy0 + lb[c + k] * cx1 * cy1;
        }
    }
}
void orc_resize_linear_f32(const float* src, int w, int h, int c, int layout, float* dst, int wo, int ho) {
    if (layout == 1) resize_linear_f32_plane(src, w, h, c, dst, wo, ho);
    else for (int k = 0; k < c; ++k)
        resize_linear_f32_plane(src + (size_t)w * h * k, w, h, 1, dst + (size_t)wo * ho * k, wo, ho);
}

/* resize_neon.cpp:17-183 (one channel) / :190-347 (three channel, called with 3*w): separable, scale in
 * double, rows = (S0*a0+S1*a1)>>4 as int16, D = ((b0*rows0>>16)+(b1*rows1>>16)+2)>>2.
 * The three-channel variant compares sx against the *tripled* width (:220), so its right-edge clamp never
 * fires; it only matters when up-scaling (reads the next row).  The intended pixel-unit clamp is used. */
static void resize_neon_rule_plane(const uint8_t* src, int w, int h, int c, uint8_t* dst, int wo, int ho) {
    double scale_x = (double)w / wo, scale_y = (double)h / ho;
    for (int dy = 0; dy < ho; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floor(fy); fy -= sy;
        if (sy < 0) { sy = 0; fy = 0.f; }
        if (sy >= h - 1) { sy = h - 2; fy = 1.f; }
        short b0 = sat_short((1.f - fy) * 2048), b1 = sat_short(fy * 2048);
        for (int dx = 0; dx < wo; ++dx) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)floor(fx); fx -= sx;
            if (sx < 0) { sx = 0; fx = 0.f; }
            if (sx >= w - 1) { sx = w - 2; fx = 1.f; }
            short a0 = sat_short((1.f - fx) * 2048), a1 = sat_short(fx * 2048);
            const uint8_t* S0 = src + ((size_t)sy * w + sx) * c;
            const uint8_t* S1 = S0 + (size_t)w * c;
            for (int k = 0; k < c; ++k) {
                short r0 = (short)((S0[k] * a0 + S0[c + k] * a1) >> 4);
                short r1 = (short)((S1[k] * a0 + S1[c + k] * a1) >> 4);
                int v = ((short)((b0 * r0) >> 16) + (short)((b1 * r1) >> 16) + 2) >> 2;
                dst[((size_t)dy * wo + dx) * c + k] = (uint8_t)clampi(v, 0, 255);   /* vqmovun_s16 */
            }
        }
    }
}
void orc_resize_linear_u8_neon_rule(const uint8_t* src, int w, int h, int c, int layout, uint8_t* dst, int wo, int ho) {
    if (layout == 1) resize_neon_rule_plane(src, w, h, c, dst, wo, ho);
    else for (int k = 0; k < c; ++k)
        resize_neon_rule_plane(src + (size_t)w * h * k, w, h, 1, dst + (size_t)wo * ho * k, wo, ho);
}

/* ------------------------------------------------------------------ bicubic fp32 (naive) ---------- */
/* resize_naive.cpp:130-141: Keys cubic, A=-0.75, expanded (non-Horner) form, left-to-right fp32. */
static void cubic_w(float fx, float* k) {
    const float A = -0.75f;
    float fx0 = fx + 1, fx1 = fx, fx2 = 1 - fx;
    k[0] = A * fx0 * fx0 * fx0 - 5 * A * fx0 * fx0 + 8 * A * fx0 - 4 * A;
    k[1] = (A + 2) * fx1 * fx1 * fx1 - (A + 3) * fx1 * fx1 + 1;
    k[2] = (A + 2) * fx2 * fx2 * fx2 - (A + 3) * fx2 * fx2 + 1;
    k[3] = 1.f - k[0] - k[1] - k[2];
}
/* resize_naive.cpp:143-185: border handling by folding coefficients (the four ifs run sequentially). */
static void cubic_table(int n_in, int n_out, int* ofs, float* al) {
    double scale = (double)n_in / n_out;
    for (int d = 0; d < n_out; ++d) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int sx = (int)floor(fx);
        fx -= sx;
        float* a = al + d * 4;
        cubic_w(fx, a);
        if (sx <= -1) { sx = 1; a[0] = 1.f - a[3]; a[1] = a[3]; a[2] = 0.f; a[3] = 0.f; }
        if (sx == 0) { sx = 1; a[0] = a[0] + a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = 0.f; }
        if (sx == n_in - 2) { sx = n_in - 3; a[3] = a[2] + a[3]; a[2] = a[1]; a[1] = a[0]; a[0] = 0.f; }
        if (sx >= n_in - 1) { sx = n_in - 3; a[3] = 1.f - a[0]; a[2] = a[0]; a[1] = 0.f; a[0] = 0.f; }
        ofs[d] = sx;
    }
}
/* The rolling 4-row cache of resize_naive.cpp:187-366 only avoids recomputation; every horizontal value is
 * S[-1]*a0 + S[0]*a1 + S[1]*a2 + S[2]*a3 and every output r0*b0 + r1*b1 + r2*b2 + r3*b3 (:230,:345). */
static void resize_cubic_f32_plane(const float* src, int w, int h, int c, float* dst, int wo, int ho,
                                   const int* xofs, const float* al, const int* yofs, const float* be) {
    (void)h;
    for (int dy = 0; dy < ho; ++dy) {
        int sy = yofs[dy];
        const float* b = be + dy * 4;
        for (int dx = 0; dx < wo; ++dx) {
            int sx = xofs[dx];
            const float* a = al + dx * 4;
            for (int k = 0; k < c; ++k) {
                float r[4];
                for (int j = 0; j < 4; ++j) {
                    const float* S = src + ((size_t)(sy - 1 + j) * w + sx) * c + k;
                    r[j] = S[-c] * a[0] + S[0] * a[1] + S[c] * a[2] + S[2 * c] * a[3];
                }
                dst[((size_t)dy * wo + dx) * c + k] = r[0] * b[0] + r[1] * b[1] + r[2] * b[2] + r[3] * b[3];
            }
        }
    }
}
void orc_resize_cubic_f32(const float* src, int w, int h, int c, int layout, float* dst, int wo, int ho) {
    int* xofs = (int*)malloc(sizeof(int) * (wo + ho));
    int* yofs = xofs + wo;
    float* al = (float*)malloc(sizeof(float) * 4 * (wo + ho));
    float* be = al + 4 * wo;
    cubic_table(w, wo, xofs, al);
    cubic_table(h, ho, yofs, be);
    if (layout == 1) resize_cubic_f32_plane(src, w, h, c, dst, wo, ho, xofs, al, yofs, be);
    else for (int k = 0; k < c; ++k)
        resize_cubic_f32_plane(src + (size_t)w * h * k, w, h, 1, dst + (size_t)wo * ho * k, wo, ho, xofs, al, yofs, be);
    free(xofs); free(al);
}

/* ------------------------------------------------------------------ bicubic u8 (OpenCV 2.4.13) ---- */
/* cvRound / saturate_cast: SSE2 cvtss2si = round half to even in the current (default) rounding mode. */
static inline int round_half_even_f(float v) { return (int)lrintf(v); }
static inline short sat_short_rhe(float v) { return (short)clampi(round_half_even_f(v), SHRT_MIN, SHRT_MAX); }
/* imgwarp.cpp interpolateCubic (Horner form, A = -0.75). */
static void cv_cubic_w(float x, float* k) {
    const float A = -0.75f;
    k[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    k[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    k[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    k[3] = 1.f - k[0] - k[1] - k[2];
}
void orc_resize_cubic_u8_cv24(const uint8_t* src, int w, int h, int c, uint8_t* dst, int wo, int ho) {
    /* cv::resize: inv_scale = dsize/ssize (double), scale = 1/inv_scale. */
    double scale_x = 1. / ((double)wo / w), scale_y = 1. / ((double)ho / h);
    int* xofs = (int*)malloc(sizeof(int) * wo);
    short* ia = (short*)malloc(sizeof(short) * 4 * wo);
    for (int dx = 0; dx < wo; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }                 /* 2.4: applied for every interpolation, x only */
        if (sx >= w - 1) { fx = 0; sx = w - 1; }
        float k[4];
        cv_cubic_w(fx, k);
        for (int j = 0; j < 4; ++j) ia[dx * 4 + j] = sat_short_rhe(k[j] * 2048);   /* INTER_RESIZE_COEF_SCALE, no sum fix-up */
        xofs[dx] = sx;
    }
    int* H = (int*)malloc(sizeof(int) * 4 * (size_t)wo * c);
    const int wide = wo * c, vec_end = wide & ~7;         /* VResizeCubicVec_32s8u handles x < (width & ~7) */
    for (int dy = 0; dy < ho; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floor(fy);
        fy -= sy;
        float kb[4];
        cv_cubic_w(fy, kb);
        short ib[4];
        for (int j = 0; j < 4; ++j) ib[j] = sat_short_rhe(kb[j] * 2048);
        for (int j = 0; j < 4; ++j) {                     /* HResizeCubic on rows clip(sy-1+j) */
            const uint8_t* S = src + (size_t)clampi(sy - 1 + j, 0, h - 1) * w * c;
            int* Hr = H + (size_t)j * wide;
            for (int dx = 0; dx < wo; ++dx)
                for (int k = 0; k < c; ++k) {
                    int v = 0;
                    for (int t = 0; t < 4; ++t) v += S[clampi(xofs[dx] - 1 + t, 0, w - 1) * c + k] * ia[dx * 4 + t];
                    Hr[dx * c + k] = v;
                }
        }
        uint8_t* D = dst + (size_t)dy * wide;
        const float s = 1.f / (2048 * 2048);
        const float b0 = ib[0] * s, b1 = ib[1] * s, b2 = ib[2] * s, b3 = ib[3] * s;
        for (int x = 0; x < wide; ++x) {
            int h0 = H[x], h1 = H[wide + x], h2 = H[2 * wide + x], h3 = H[3 * wide + x];
            int v;
            if (x < vec_end) {   /* SSE2 path: fp32 mul/add chain, cvtps2dq (half-even), packs (s16) + packus (u8) */
                float f = (float)h0 * b0;
                f = f + (float)h1 * b1;
                f = f + (float)h2 * b2;
                f = f + (float)h3 * b3;
                v = clampi(round_half_even_f(f), SHRT_MIN, SHRT_MAX);
            } else {             /* scalar tail: FixedPtCast<int,uchar,22> */
                v = (h0 * ib[0] + h1 * ib[1] + h2 * ib[2] + h3 * ib[3] + (1 << 21)) >> 22;
            }
            D[x] = (uint8_t)clampi(v, 0, 255);
        }
    }
    free(H); free(ia); free(xofs);
}

/* ------------------------------------------------------------------ warp affine ------------------- */
void orc_invert_affine(float* m) {  /* warp_affine.cpp:121-133, types exactly as written */
    double D = m[0] * m[4] - m[1] * m[3];    /* float arithmetic, widened afterwards */
    D = D != 0 ? 1. / D : 0;
    double A11 = m[4] * D;
    double A22 = m[0] * D;
    m[0] = (float)A11;
    m[1] = (float)(m[1] * -D);               /* m[1] *= -D : float*double -> double -> float */
    m[3] = (float)(m[3] * -D);
    m[4] = (float)A22;
    double b1 = -m[0] * m[2] - m[1] * m[5];  /* float arithmetic with the UPDATED m[0], m[1] */
    double b2 = -m[3] * m[2] - m[4] * m[5];
    m[2] = (float)b1;
    m[5] = (float)b2;
}

void orc_rotation_matrix(float scale, float rot_deg, const double* aux, float* m) {
    /* warp_affine.cpp:76-94 with point (0,0) */
    float angle = rot_deg;
    angle *= M_PI / 180;                      /* float * double -> double -> float */
    double alpha = scale * cos(angle);
    double beta = scale * sin(angle);
    m[0] = (float)alpha;
    m[1] = (float)beta;
    m[2] = (float)((1 - alpha) * 0.f - beta * 0.f);
    m[3] = (float)-beta;
    m[4] = (float)alpha;
    m[5] = (float)(beta * 0.f + (1 - alpha) * 0.f);
    /* warp_affine.cpp:105-106 (double arithmetic: VScalar is double) */
    m[2] = (float)(aux[2] - m[0] * aux[0] - m[1] * aux[1]);
    m[5] = (float)(aux[3] - m[3] * aux[0] - m[4] * aux[1]);
}

static void warp_u8_plane(const uint8_t* src, int w, int h, int c, uint8_t* dst, int wo, int ho, const float* m, int sc) {
    for (int dy = 0; dy < ho; ++dy)
        for (int dx = 0; dx < wo; ++dx) {
            float fx = m[0] * dx + m[1] * dy + m[2];   /* warp_affine_naive.cpp:23-24, fp32 left-to-right */
            float fy = m[3] * dx + m[4] * dy + m[5];
            int sy = (int)floor(fy);
            fy -= sy;
            if (sy < 0 || sy >= h - 1) continue;
            short cy0 = sat_short((1.f - fy) * 2048);
            short cy1 = sat_short(2048 - cy0);          /* sums to exactly 2048 (:32) */
            int sx = (int)floor(fx);
            fx -= sx;
            if (sx < 0 || sx >= w - 1) continue;
            short cx0 = sat_short((1.f - fx) * 2048);
            short cx1 = sat_short(2048 - cx0);
            const uint8_t* lt = src + ((size_t)sy * w + sx) * c;
            const uint8_t* lb = lt + (size_t)w * c;
            uint8_t* o = dst + ((size_t)dy * wo + dx) * c;
            for (int k = 0; k < c; ++k) {
                int v = (pix(lt + k, sc) * cx0 * cy0 + pix(lb + k, sc) * cx0 * cy1 +
                         pix(lt + c + k, sc) * cx1 * cy0 + pix(lb + c + k, sc) * cx1 * cy1) >> 22;
                o[k] = (uint8_t)v;
            }
        }
}
void orc_warp_affine_u8(const uint8_t* src, int w, int h, int c, int layout, uint8_t* dst, int wo, int ho,
                        const float* m, int sc) {
    if (layout == 1) warp_u8_plane(src, w, h, c, dst, wo, ho, m, sc);
    else for (int k = 0; k < c; ++k)   /* warp_affine.cpp:152-168 */
        warp_u8_plane(src + (size_t)w * h * k, w, h, 1, dst + (size_t)wo * ho * k, wo, ho, m, sc);
}

static void warp_f32_plane(const float* src, int w, int h, int c, float* dst, int wo, int ho, const float* m) {
    for (int dy = 0; dy < ho; ++dy)
        for (int dx = 0; dx < wo; ++dx) {
            float fx = m[0] * dx + m[1] * dy + m[2];
            float fy = m[3] * dx + m[4] * dy + m[5];
            int sy = (int)floor(fy);
            fy -= sy;
            if (sy < 0 || sy >= h - 1) continue;
            float cy0 = 1.f - fy, cy1 = fy;
            int sx = (int)floor(fx);
            fx -= sx;
            if (sx < 0 || sx >= w - 1) continue;
            float cx0 = 1.f - fx, cx1 = fx;
            const float* lt = src + ((size_t)sy * w + sx) * c;
            const float* lb = lt + (size_t)w * c;
            float* o = dst + ((size_t)dy * wo + dx) * c;
            for (int k = 0; k < c; ++k)
                o[k] = lt[k] * cx0 * cy0 + lb[k] * cx0 * cy1 + lt[c + k] * cx1 * cy0 + lb[c + k] * cx1 * cy1;
        }
}
void orc_warp_affine_f32(const float* src, int w, int h, int c, int layout, float* dst, int wo, int ho, const float* m) {
    if (layout == 1) warp_f32_plane(src, w, h, c, dst, wo, ho, m);
    else for (int k = 0; k < c; ++k)
        warp_f32_plane(src + (size_t)w * h * k, w, h, 1, dst + (size_t)wo * ho * k, wo, ho, m);
}

/* ------------------------------------------------------------------ statistics + normalize -------- */
void orc_sums_u8(const uint8_t* src, size_t pixels, int c, int layout, uint64_t* sums) {
    for (int k = 0; k < c; ++k) {
        uint64_t sx = 0, sxx = 0;
        for (size_t i = 0; i < pixels; ++i) {
            uint64_t v = layout == 1 ? src[i * c + k] : src[(size_t)k * pixels + i];
            sx += v; sxx += v * v;
        }
        sums[2 * k] += sx; sums[2 * k + 1] += sxx;
    }
}
void orc_sums_f32(const float* src, size_t pixels, int c, int layout, double* sums) {
    for (int k = 0; k < c; ++k) {
        double sx = 0, sxx = 0;
        for (size_t i = 0; i < pixels; ++i) {
            double v = layout == 1 ? src[i * c + k] : src[(size_t)k * pixels + i];
            sx += v; sxx += v * v;
        }
        sums[2 * k] += sx; sums[2 * k + 1] += sxx;
    }
}
void orc_finalize_mean_stddev_f64(const double* sums, int c, uint64_t n, float* mean, float* stddev) {
    for (int k = 0; k < c; ++k) {
        double m = sums[2 * k] / (double)n;
        double var = sums[2 * k + 1] / (double)n - m * m;
        if (var < 0) var = 0;
        mean[k] = (float)m;
        stddev[k] = (float)sqrt(var);
    }
}
void orc_finalize_mean_stddev(const uint64_t* sums, int c, uint64_t n, float* mean, float* stddev) {
    for (int k = 0; k < c; ++k) {
        double m = (double)sums[2 * k] / (double)n;
        double var = (double)sums[2 * k + 1] / (double)n - m * m;
        if (var < 0) var = 0;
        mean[k] = (float)m;
        stddev[k] = (float)sqrt(var);
    }
}
void orc_mean_stddev_f32_sequential(const float* src, size_t pixels, int c, int layout, float* mean, float* stddev) {
    /* normalize_naive.cpp:7-72: sum in fp32, mean = sum/N (float/int), then sum of ((x-mean)^2 / N) in fp32. */
    int stride = (int)pixels;
    for (int k = 0; k < c; ++k) {
        float s = 0.f;
        for (size_t i = 0; i < pixels; ++i) s += layout == 1 ? src[i * c + k] : src[(size_t)k * pixels + i];
        mean[k] = s / stride;
    }
    for (int k = 0; k < c; ++k) {
        float acc = 0.f;
        for (size_t i = 0; i < pixels; ++i) {
            float p = layout == 1 ? src[i * c + k] : src[(size_t)k * pixels + i];
            p -= mean[k];
            p = p * p;
            acc += p / stride;
        }
        stddev[k] = (float)sqrt(acc);   /* sqrt(double) of a float, stored to float: correctly rounded == sqrtf */
    }
}

void orc_normalize_f32(const float* src, size_t pixels, int c, int layout, const float* mean, const float* stddev, float* dst) {
    for (int k = 0; k < c; ++k) {   /* normalize_naive.cpp:74-90 */
        double den = stddev[k] + 1e-6;
        for (size_t i = 0; i < pixels; ++i) {
            size_t j = layout == 1 ? i * c + k : (size_t)k * pixels + i;
            dst[j] = (float)((src[j] - mean[k]) / den);
        }
    }
}
void orc_normalize_u8(const uint8_t* src, size_t pixels, int c, int layout, const float* mean, const float* stddev, float* dst) {
    for (int k = 0; k < c; ++k) {   /* normalize.cpp:92-95 (change_dtype) then normalize_naive.cpp:74-90 */
        double den = stddev[k] + 1e-6;
        for (size_t i = 0; i < pixels; ++i) {
            size_t j = layout == 1 ? i * c + k : (size_t)k * pixels + i;
            dst[j] = (float)(((float)src[j] - mean[k]) / den);
        }
    }
}

/* ------------------------------------------------------------------ compositions ----- */
void orc_nv_resize_normalize_chw(const uint8_t* src, int w, int h, int v_first, int wo, int ho,
                                 const float* mean, const float* stddev, float* dst) {
    uint8_t* bgr = (uint8_t*)malloc((size_t)w * h * 3);
    uint8_t* sm = (uint8_t*)malloc((size_t)wo * ho * 3);
    float* nf = (float*)malloc(sizeof(float) * (size_t)wo * ho * 3);
    orc_nv_to_bgr(src, w, h, v_first, bgr);
    if (wo == w && ho == h) memcpy(sm, bgr, (size_t)w * h * 3);   /* resize.cpp:58-61 */
    else orc_resize_linear_u8(bgr, w, h, 3, 1, sm, wo, ho, 0);
    orc_normalize_u8(sm, (size_t)wo * ho, 3, 1, mean, stddev, nf);
    orc_hwc_to_chw(nf, wo, ho, 3, 4, dst);
    free(bgr); free(sm); free(nf);
}

typedef struct {
    const uint8_t* src; int n, w, h, v_first, wo, ho; const float* mean; const float* stddev; float* dst;
    int* next; pthread_mutex_t* mu;
} batch_job;
static void* batch_worker(void* p) {
    batch_job* j = (batch_job*)p;
    for (;;) {
        pthread_mutex_lock(j->mu);
        int i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->n) break;
        orc_nv_resize_normalize_chw(j->src + (size_t)i * j->w * j->h * 3 / 2, j->w, j->h, j->v_first, j->wo, j->ho,
                                    j->mean, j->stddev, j->dst + (size_t)i * j->wo * j->ho * 3);
    }
    return NULL;
}
void orc_nv_resize_normalize_chw_batch(const uint8_t* src, int n, int w, int h, int v_first, int wo, int ho,
                                       const float* mean, const float* stddev, float* dst, int threads) {
    int next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    batch_job j = {src, n, w, h, v_first, wo, ho, mean, stddev, dst, &next, &mu};
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, batch_worker, &j);
    batch_worker(&j);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    free(th);
}

void orc_warp_affine_normalize(const uint8_t* src, int w, int h, int c, const float* m, int wo, int ho,
                               const float* mean, const float* stddev, float* dst) {
    uint8_t* tmp = (uint8_t*)calloc((size_t)wo * ho * c, 1);   /* zero-filled dst (App. C-5) */
    orc_warp_affine_u8(src, w, h, c, 1, tmp, wo, ho, m, 0);
    orc_normalize_u8(tmp, (size_t)wo * ho, c, 1, mean, stddev, dst);
    free(tmp);
}
