"""Pin the C restatement (oracle/vacv_oracle.c) against the UNMODIFIED reference compiled from its own
sources (oracle/_ref) on seeded inputs, the reference's fixtures and the reference tests' known-answer
parameters.  CPU only.  Integer/u8 ops: bit-exact.  fp32 ops: bit-exact as well (same IEEE operation order;
no FMA on either side), which is stricter than the 1e-5 the north star asks for."""
import numpy as np
import pytest

from oracle_lib import (COLOR_YUV2BGR_NV12, COLOR_YUV2BGR_NV21, FP32, INT8, INTER_CUBIC, INTER_LINEAR, NCHW, NHWC,
                        load_fixture)


def rng(seed):
    return np.random.default_rng(seed)


def u8(seed, *shape):
    return rng(seed).integers(0, 256, shape, dtype=np.uint8)


def f32(seed, *shape):
    return (rng(seed).random(shape, dtype=np.float32) * 255).astype(np.float32)


# ---------------------------------------------------------------- yuv -> bgr (a1)
@pytest.mark.parametrize("w,h", [(2, 2), (16, 8), (176, 144), (642, 362), (1920, 1080)])
@pytest.mark.parametrize("code", [COLOR_YUV2BGR_NV21, COLOR_YUV2BGR_NV12])
def test_nv_to_bgr(oracle, ref, w, h, code):
    src = u8(w * 31 + h, w * h * 3 // 2)
    want = ref.cvt_color(src, w, h, code)
    # the reference decodes BOTH codes with V-first chroma (cvt_color.cpp:139-149)
    got = oracle.nv_to_bgr(src, w, h, v_first=1)
    assert np.array_equal(got, want)
    # true NV12 == V-first decode of the chroma-swapped frame
    sw = src.copy()
    sw[w * h::2], sw[w * h + 1::2] = src[w * h + 1::2], src[w * h::2]
    assert np.array_equal(oracle.nv_to_bgr(sw, w, h, v_first=0), want)


def make_yuv_surface(seed, fmt, w, h, y_pitch, c_pitch):
    """Random pitched surface (padding bytes random too) + the dense NV21 frame holding the same samples."""
    r = rng(seed)
    planar = fmt >= 2
    yp = r.integers(0, 256, (h, y_pitch), dtype=np.uint8)
    u = r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    v = r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    if planar:
        up = r.integers(0, 256, (h // 2, c_pitch), dtype=np.uint8)
        vp = r.integers(0, 256, (h // 2, c_pitch), dtype=np.uint8)
        up[:, :w // 2], vp[:, :w // 2] = u, v
        chroma = np.concatenate([up.ravel(), vp.ravel()] if fmt == 2 else [vp.ravel(), up.ravel()])
    else:
        cp = r.integers(0, 256, (h // 2, c_pitch), dtype=np.uint8)
        cp[:, 0:w:2], cp[:, 1:w:2] = (v, u) if fmt == 0 else (u, v)
        chroma = cp.ravel()
    dense = np.empty(w * h * 3 // 2, np.uint8)
    dense[:w * h] = yp[:, :w].ravel()
    dense[w * h::2], dense[w * h + 1::2] = v.ravel(), u.ravel()
    return np.concatenate([yp.ravel(), chroma]), dense


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("w,h,yp,cp", [(16, 8, 0, 0), (176, 144, 192, 0), (642, 362, 656, 336), (1920, 1080, 2048, 0)])
def test_yuv_surface_to_bgr(oracle, ref, fmt, w, h, yp, cp):
    """Next-row extension (pitch, I420/YV12): same matrix as the reference on the re-packed frame."""
    y_pitch = yp or w
    c_pitch = cp or (w // 2 if fmt >= 2 else w)
    if fmt < 2:
        c_pitch = max(c_pitch, w)
    surf, dense = make_yuv_surface(w + fmt, fmt, w, h, y_pitch, c_pitch)
    want = ref.cvt_color(dense, w, h, COLOR_YUV2BGR_NV21)
    assert np.array_equal(oracle.yuv_to_bgr(surf, fmt, w, h, y_pitch, c_pitch), want)


def test_bgr2nv21_roundtrip_fixture(oracle, ref):
    img = load_fixture("t640x360")
    if img is None:
        pytest.skip("fixtures not staged")
    nv = ref.bgr2nv21(img)
    assert np.array_equal(oracle.bgr_to_nv21(img), nv)
    assert np.array_equal(oracle.nv_to_bgr(nv, 640, 360), ref.cvt_color(nv, 640, 360))


# ---------------------------------------------------------------- crop (a2)
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("dt", ["u8", "f32"])
@pytest.mark.parametrize("rect", [(0, 0, 5, 5), (0, 0, 320, 180), (7.9, 3.2, 200.5, 99.7), (33, 17, 640, 360)])
def test_crop(oracle, ref, layout, dt, rect):
    w, h, c = 640, 360, 3
    src = u8(1, h * w * c) if dt == "u8" else f32(1, h * w * c)
    l, t, r, b = rect
    want = ref.crop(src, w, h, c, layout, l, t, r, b)
    cw, ch = int(np.float32(r) - np.float32(l)), int(np.float32(b) - np.float32(t))   # crop.cpp:128-131
    got = oracle.crop(src, w, h, c, layout, int(l), int(t), cw, ch)
    assert got.shape == want.shape and np.array_equal(got, want)


# ---------------------------------------------------------------- layout / dtype (a3, a4)
@pytest.mark.parametrize("dt", ["u8", "f32"])
@pytest.mark.parametrize("w,h,c", [(176, 144, 3), (33, 7, 4), (5, 3, 2)])
def test_layout(oracle, ref, dt, w, h, c):
    src = u8(2, h, w, c) if dt == "u8" else f32(2, h, w, c)
    chw = ref.change_layout(src, w, h, c, NHWC, NCHW)
    assert np.array_equal(oracle.hwc_to_chw(src, w, h, c), chw)
    assert np.array_equal(oracle.chw_to_hwc(chw, w, h, c), ref.change_layout(chw, w, h, c, NCHW, NHWC))
    assert np.array_equal(oracle.chw_to_hwc(chw, w, h, c), src)


def test_dtype(oracle, ref):
    w, h, c = 176, 144, 3
    src = u8(3, h, w, c)
    f = ref.change_dtype(src, w, h, c, NHWC, FP32)
    assert np.array_equal(oracle.u8_to_f32(src), f)
    g = (rng(4).random((h, w, c), dtype=np.float32) * 255.999).astype(np.float32)   # domain [0,256)
    assert np.array_equal(oracle.f32_to_u8(g), ref.change_dtype(g, w, h, c, NHWC, INT8))


# ---------------------------------------------------------------- bilinear (a5, a6)
SIZES = [((64, 48), (20, 16)), ((64, 48), (200, 111)), ((1920, 1080), (640, 360)), ((1920, 1080), (640, 640)),
         ((333, 211), (500, 300)), ((2, 2), (7, 5)), ((640, 360), (639, 359))]


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", SIZES)
def test_resize_linear_u8(oracle, ref, layout, sz):
    (w, h), (wo, ho) = sz
    c = 3
    src = u8(5, c * h * w)
    want = ref.resize(src, w, h, c, layout, wo, ho, INTER_LINEAR)
    got = oracle.resize_linear(src, w, h, c, layout, wo, ho)
    assert np.array_equal(got, want)


def test_resize_linear_u8_signed_char_compat(oracle, ref_schar):
    w, h, c, wo, ho = 333, 211, 3, 200, 100
    src = u8(6, h * w * c)
    want = ref_schar.resize(src, w, h, c, NHWC, wo, ho, INTER_LINEAR)
    assert np.array_equal(oracle.resize_linear(src, w, h, c, NHWC, wo, ho, signed_char=1), want)


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", [((64, 48), (20, 16)), ((640, 360), (213, 120)), ((1920, 1080), (640, 360)), ((333, 211), (200, 100)),
                                ((1280, 720), (500, 300)), ((64, 48), (200, 111)), ((176, 144), (640, 640))])
def test_resize_linear_u8_neon_rule_vs_reference_neon_source(oracle, ref, layout, sz):
    """Row a7: the NEON rounding rule (resize_neon.cpp:12-347) is aarch64-only in the reference; its source is compiled
    unmodified against a scalar emulation of the nine intrinsics it uses (oracle/neon_emul/arm_neon.h) and compared with
    the oracle's restatement.  CHW (one-channel kernel): identical everywhere.  HWC (three-channel kernel, called with
    tripled widths, resize.cpp:133-134): identical wherever the right-edge clamp cannot trigger, i.e. for down-scaling;
    when up-scaling the reference compares sx with the tripled width (:220), never clamps and reads past the row end --
    there the oracle keeps the pixel-unit clamp of the one-channel kernel and only the columns left of the clamp agree."""
    (w, h), (wo, ho) = sz
    shape = (h, w, 3) if layout == NHWC else (3, h, w)
    src = u8(w + 7 * ho, *shape)
    want = ref.resize_neon(src, w, h, layout, wo, ho)
    got = oracle.resize_linear_neon_rule(src, w, h, 3, layout, wo, ho)
    if layout == NCHW or wo <= w:
        assert np.array_equal(got, want)
    else:
        scale = w / wo
        safe = int(np.floor((w - 1 + 0.5) / scale - 0.5))      # columns whose left tap is <= w - 2
        assert np.array_equal(got[:, :safe], want[:, :safe])


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", SIZES[:5])
def test_resize_linear_f32(oracle, ref, layout, sz):
    (w, h), (wo, ho) = sz
    c = 3
    src = f32(7, c * h * w)
    want = ref.resize(src, w, h, c, layout, wo, ho, INTER_LINEAR)
    got = oracle.resize_linear(src, w, h, c, layout, wo, ho)
    assert np.array_equal(got, want)


def test_resize_fixture_config1(oracle, ref):
    img = load_fixture("universe1920x1080")
    if img is None:
        pytest.skip("fixtures not staged")
    want = ref.resize(img, 1920, 1080, 3, NHWC, 640, 360, INTER_LINEAR)
    assert np.array_equal(oracle.resize_linear(img, 1920, 1080, 3, NHWC, 640, 360), want)
    # at the exact 3:1 ratio bundled OpenCV agrees byte for byte
    assert np.array_equal(ref.cv_resize(img, 1920, 1080, 3, 640, 360, INTER_LINEAR), want)


# ---------------------------------------------------------------- bicubic fp32 (a8)
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", [((64, 48), (20, 20)), ((64, 48), (100, 100)), ((320, 180), (240, 240))])
def test_resize_cubic_f32_square_vs_reference_api(oracle, ref, layout, sz):
    """w_out == h_out: the reference's public entry point is correct (App. C-2)."""
    (w, h), (wo, ho) = sz
    src = f32(8, 3 * h * w)
    want = ref.resize(src, w, h, 3, layout, wo, ho, INTER_CUBIC)
    got = oracle.resize_cubic_f32(src, w, h, 3, layout, wo, ho)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", [((64, 48), (37, 20)), ((64, 48), (20, 37)), ((2560, 1440), (1920, 1080)),
                                ((320, 180), (640, 360)), ((16, 16), (5, 9))])
def test_resize_cubic_f32_vs_reference_blocks(oracle, ref, layout, sz):
    """any shape: the reference's own building blocks driven with non-aliased buffers."""
    (w, h), (wo, ho) = sz
    src = f32(9, 3 * h * w)
    want = ref.resize_cubic_f32_fixed(src, w, h, 3, layout, wo, ho)
    got = oracle.resize_cubic_f32(src, w, h, 3, layout, wo, ho)
    assert np.array_equal(got, want)
    if layout == NHWC and wo >= 8:
        # sanity: away from the borders (the reference folds coefficients, OpenCV replicates pixels) it agrees
        # with bundled OpenCV's fp32 cubic
        cvr = ref.cv_resize(src, w, h, 3, wo, ho, INTER_CUBIC)
        b = 2 * max(1, -(-wo // w), -(-ho // h)) + 2
        assert np.abs(cvr - got)[b:-b, b:-b].max() < 2e-3 * 255


# ---------------------------------------------------------------- bicubic u8 = OpenCV 2.4.13 (a9)
@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("sz", [((2560, 1440), (1920, 1080)), ((256, 144), (100, 70)), ((176, 144), (640, 640)),
                                ((257, 145), (300, 171)), ((64, 48), (333, 77)), ((64, 48), (21, 13)),
                                ((8, 8), (3, 3)), ((5, 4), (13, 11))])
def test_resize_cubic_u8_cv24(oracle, ref, c, sz):
    (w, h), (wo, ho) = sz
    if c != 3 and w > 1000:
        pytest.skip("big case only for c=3")
    src = u8(10 + c, h * w * c)
    want = ref.cv_resize(src, w, h, c, wo, ho, INTER_CUBIC)
    got = oracle.resize_cubic_u8(src, w, h, c, wo, ho)
    assert np.array_equal(got, want)


def test_resize_cubic_u8_fixture_config4(oracle, ref):
    img = load_fixture("lakers2560x1440")
    if img is None:
        pytest.skip("fixtures not staged")
    want = ref.cv_resize(img, 2560, 1440, 3, 1920, 1080, INTER_CUBIC)
    assert np.array_equal(oracle.resize_cubic_u8(img, 2560, 1440, 3, 1920, 1080), want)


# ---------------------------------------------------------------- warp affine (a10)
M_TEST = [0.849158, 0.012257, -474.827, -0.01225, 0.849158, -379.18]           # test_warp_affine.cpp:31-32
ROT_TEST = dict(scale=1.073914, rot=-3.314525, aux=[738.518372, 537.672852, 204.766998, 73.329681])  # :198-205


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("dt", ["u8", "f32"])
def test_warp_affine_matrix(oracle, ref, layout, dt):
    w, h, c, wo, ho = 1280, 720, 3, 240, 240
    img = load_fixture("t1280x720")
    src = img if img is not None else u8(11, h, w, c)
    if layout == NCHW:
        src = np.ascontiguousarray(src.transpose(2, 0, 1))
    if dt == "f32":
        src = src.astype(np.float32)
    want, m_inv = ref.warp_affine(src, w, h, c, layout, wo, ho, M_TEST)
    assert np.array_equal(oracle.invert_affine(M_TEST).view(np.uint32), m_inv.view(np.uint32))
    got = oracle.warp_affine(src, w, h, c, layout, wo, ho, m_inv)
    assert np.array_equal(got, want)


def test_warp_affine_rotation(oracle, ref):
    w, h, wo, ho = 1280, 720, 140, 210
    img = load_fixture("t1280x720_grey")
    src = img if img is not None else u8(12, h, w, 1)
    want = ref.warp_affine_rot(src, w, h, 1, NHWC, wo, ho, ROT_TEST["scale"], ROT_TEST["rot"], ROT_TEST["aux"])
    m = oracle.invert_affine(oracle.rotation_matrix(ROT_TEST["scale"], ROT_TEST["rot"], ROT_TEST["aux"]))
    got = oracle.warp_affine(src, w, h, 1, NHWC, wo, ho, m)
    assert np.array_equal(got, want)
    assert want.any()


@pytest.mark.parametrize("seed", range(8))
def test_warp_affine_random(oracle, ref, seed):
    r = rng(100 + seed)
    w, h, c, wo, ho = 320, 200, 3, 112, 112
    s = r.uniform(0.3, 0.9)
    a = np.deg2rad(r.uniform(-30, 30))
    m = [s * np.cos(a), s * np.sin(a), r.uniform(-60, 10), -s * np.sin(a), s * np.cos(a), r.uniform(-60, 10)]
    src = u8(seed, h, w, c)
    want, m_inv = ref.warp_affine(src, w, h, c, NHWC, wo, ho, m)
    assert np.array_equal(oracle.invert_affine(m).view(np.uint32), m_inv.view(np.uint32))
    assert np.array_equal(oracle.warp_affine(src, w, h, c, NHWC, wo, ho, m_inv), want)


def test_warp_affine_signed_char_compat(oracle, ref_schar):
    w, h, c, wo, ho = 320, 200, 3, 100, 80
    src = u8(13, h, w, c)
    m = [0.5, 0.05, -10, -0.05, 0.5, -5]
    want, m_inv = ref_schar.warp_affine(src, w, h, c, NHWC, wo, ho, m)
    assert np.array_equal(oracle.warp_affine(src, w, h, c, NHWC, wo, ho, m_inv, signed_char=1), want)


# ---------------------------------------------------------------- statistics + normalize (a11, a12)
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("name,w,h", [("t176x144", 176, 144), ("t284x214", 284, 214)])
def test_normalize_given_stats(oracle, ref, layout, name, w, h):
    img = load_fixture(name)
    src = img if img is not None else u8(14, h, w, 3)
    if layout == NCHW:
        src = np.ascontiguousarray(src.transpose(2, 0, 1))
    mean = np.array([104.5, 117.25, 123.125], np.float32)
    std = np.array([57.375, 57.12, 58.395], np.float32)
    want = ref.normalize(src, w, h, 3, layout, mean, std)
    assert np.array_equal(oracle.normalize(src, w * h, 3, layout, mean, std), want)
    wantf = ref.normalize(src.astype(np.float32), w, h, 3, layout, mean, std)
    assert np.array_equal(oracle.normalize(src.astype(np.float32), w * h, 3, layout, mean, std), wantf)


@pytest.mark.parametrize("layout", [NHWC, NCHW])
def test_mean_stddev(oracle, ref, layout):
    w, h, c = 284, 214, 3
    img = load_fixture("t284x214")
    hwc = img if img is not None else u8(15, h, w, c)
    src = hwc if layout == NHWC else np.ascontiguousarray(hwc.transpose(2, 0, 1))
    # (i) the verbatim sequential-fp32 restatement equals the reference bit for bit
    m_ref, s_ref = ref.mean_stddev_f32(src.astype(np.float32), w, h, c, layout)
    m_seq, s_seq = oracle.mean_stddev_f32_sequential(src.astype(np.float32), w * h, c, layout)
    assert np.array_equal(m_ref, m_seq) and np.array_equal(s_ref, s_seq)
    # (ii) the exact-sum statistic (the parity pin, App. C-4) equals bundled cv::meanStdDev, the truth the
    #      reference's own test uses (test_normalize.cpp:31)
    sums = oracle.sums_u8(src, w * h, c, layout)
    m, s = oracle.finalize_mean_stddev(sums, c, w * h)
    m_cv, s_cv = ref.cv_mean_stddev(hwc, w, h, c)
    assert np.allclose(m, m_cv, rtol=1e-6) and np.allclose(s, s_cv, rtol=1e-6)
    # (iii) and is close to the reference's own value at this small size
    assert np.allclose(m, m_ref, rtol=1e-4) and np.allclose(s, s_ref, rtol=1e-3)


# ---------------------------------------------------------------- fused composition (a13 / config 2)
@pytest.mark.parametrize("w,h,wo,ho", [(1920, 1080, 640, 640), (640, 360, 224, 224), (64, 48, 100, 37)])
def test_pipeline_config2(oracle, ref, w, h, wo, ho):
    src = u8(16, w * h * 3 // 2)
    mean = np.array([103.53, 116.28, 123.675], np.float32)
    std = np.array([57.375, 57.12, 58.395], np.float32)
    want = ref.pipeline(src, w, h, COLOR_YUV2BGR_NV21, wo, ho, mean, std)
    got = oracle.nv_resize_normalize_chw(src, w, h, 1, wo, ho, mean, std)
    assert np.array_equal(got, want)
    gotb = oracle.nv_resize_normalize_chw(np.concatenate([src, src]), w, h, 1, wo, ho, mean, std, batch=2, threads=2)
    assert np.array_equal(gotb[0], want) and np.array_equal(gotb[1], want)
