"""GPU parity: every CUDA operator, called through the C-ABI (libvacv_cuda.so via ctypes), against the oracle
(oracle/vacv_oracle.c) and -- when oracle/_ref travelled to the box -- the unmodified reference itself, on the same
seeded inputs.  Integer / u8 outputs: bit-exact.  fp32 outputs: bit-exact where the operation order is reproduced
(everything here), which implies the north star's 1e-5 relative / cosine >= 0.99999 bounds; those bounds are also
asserted explicitly for the fused pipeline."""
import numpy as np
import pytest

from oracle_lib import (COLOR_YUV2BGR_NV21, INTER_CUBIC as R_CUBIC, INTER_LINEAR as R_LINEAR, NCHW, NHWC, Ref,
                        load_fixture, ref_available)

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def vacv():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vacv_b200
    return vacv_b200


def rng(seed):
    return np.random.default_rng(seed)


def u8(seed, *shape):
    return rng(seed).integers(0, 256, shape, dtype=np.uint8)


def f32(seed, *shape):
    return (rng(seed).random(shape, dtype=np.float32) * 255).astype(np.float32)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_same(got, want):
    assert got.shape == want.shape, (got.shape, want.shape)
    if not np.array_equal(bits(got), bits(want)):
        bad = np.flatnonzero(bits(got).ravel() != bits(want).ravel())
        raise AssertionError(f"{bad.size} of {got.size} elements differ; first at {bad[0]}: got {got.ravel()[bad[0]]} want {want.ravel()[bad[0]]}")


def _gpu_rand_u8(seed, *shape):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)


MEAN = np.array([103.53, 116.28, 123.675], np.float32)
STD = np.array([57.375, 57.12, 58.395], np.float32)


# ------------------------------------------------------------------ a1 yuv -> bgr
@pytest.mark.parametrize("w,h,b", [(1920, 1080, 2), (640, 360, 3), (16, 2, 1), (2, 2, 1), (642, 362, 2), (4096, 16, 1)])
@pytest.mark.parametrize("v_first", [True, False])
def test_cvt_nv2bgr(vacv, oracle, w, h, b, v_first):
    src = u8(w + h, b, w * h * 3 // 2)
    got = host(vacv.cvt_nv2bgr(dev(src), w, h, v_first))
    want = np.stack([oracle.nv_to_bgr(src[i], w, h, int(v_first)) for i in range(b)])
    assert_same(got, want)


def test_cvt_nv2bgr_vs_reference_roundtrip_fixture(vacv):
    """The reference's own test: BGR fixture -> bgr2nv21 -> cvt_color (test_cvt_color.cpp:23-77)."""
    img = load_fixture("universe1920x1080")
    if img is None or not ref_available():
        pytest.skip("oracle/_ref not staged")
    ref = Ref()
    nv = ref.bgr2nv21(img)
    want = ref.cvt_color(nv, 1920, 1080, COLOR_YUV2BGR_NV21)
    got = host(vacv.cvt_nv2bgr(dev(nv[None]), 1920, 1080, True))[0]
    assert_same(got, want)


def test_cvt_nv2bgr_rejects_odd(vacv):
    with pytest.raises(vacv.VacvError):
        vacv.cvt_nv2bgr(dev(u8(0, 1, 15 * 10 * 3 // 2 + 8)), 15, 10)


# ------------------------------------------------------------------ a2 crop
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("dt", ["u8", "f32"])
@pytest.mark.parametrize("rect", [(0, 0, 5, 5), (0, 0, 320, 180), (7, 3, 193, 96), (33, 17, 607, 343), (1, 0, 639, 360)])
def test_crop(vacv, oracle, layout, dt, rect):
    w, h, c, b = 640, 360, 3, 2
    shape = (b, h, w, c) if layout == NHWC else (b, c, h, w)
    src = u8(1, *shape) if dt == "u8" else f32(1, *shape)
    l, t, cw, ch = rect
    got = host(vacv.crop(dev(src), layout, l, t, cw, ch))
    want = np.stack([oracle.crop(src[i], w, h, c, layout, l, t, cw, ch) for i in range(b)])
    assert_same(got, want)


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("dt", ["u8", "f32"])
@pytest.mark.parametrize("rect", [(7, 3, 192, 100), (33, 17, 400, 300), (5, 1, 336, 359), (239, 0, 400, 65), (1, 295, 512, 64)])
def test_crop_16_byte_rows(vacv, oracle, layout, dt, rect):
    """Rects whose rows are multiples of 16 bytes (no head / tail bytes), with left edges at every kind of byte alignment."""
    w, h, c, b = 640, 360, 3, 3
    shape = (b, h, w, c) if layout == NHWC else (b, c, h, w)
    src = u8(5, *shape) if dt == "u8" else f32(5, *shape)
    l, t, cw, ch = rect
    got = host(vacv.crop(dev(src), layout, l, t, cw, ch))
    want = np.stack([oracle.crop(src[i], w, h, c, layout, l, t, cw, ch) for i in range(b)])
    assert_same(got, want)


def test_crop_large_batch_equals_slicing(vacv):
    """64 x 1080p -> 1280x720 at (321, 181) and at a 16-byte aligned left edge; equals plain slicing."""
    src = _gpu_rand_u8(6, 64, 1080, 1920, 3)
    got = vacv.crop(src, NHWC, 321, 181, 1280, 720)
    assert torch.equal(got, src[:, 181:901, 321:1601, :])
    assert torch.equal(vacv.crop(src, NHWC, 320, 0, 1280, 1080), src[:, :, 320:1600, :])
    chw = _gpu_rand_u8(7, 16, 3, 1080, 1920)
    got = vacv.crop(chw, NCHW, 321, 181, 1280, 720)
    assert torch.equal(got, chw[:, :, 181:901, 321:1601])


def test_crop_reference_sizes(vacv, oracle):
    """test_crop.cpp:16-20 rects at the origin of the 2560x1440 fixture shape."""
    src = u8(2, 1, 1440, 2560, 3)
    for cw, ch in [(5, 5), (320, 180), (640, 360), (1280, 720), (1920, 1080)]:
        got = host(vacv.crop(dev(src), NHWC, 0, 0, cw, ch))[0]
        assert_same(got, oracle.crop(src[0], 2560, 1440, 3, NHWC, 0, 0, cw, ch))


def test_crop_rejects_out_of_frame(vacv):
    with pytest.raises(vacv.VacvError):
        vacv.crop(dev(u8(0, 1, 10, 10, 3)), NHWC, 5, 5, 10, 10)


# ------------------------------------------------------------------ a3 / a4 layout, dtype
@pytest.mark.parametrize("dt", ["u8", "f32"])
@pytest.mark.parametrize("w,h,c", [(176, 144, 3), (1920, 1080, 3), (33, 7, 3), (33, 7, 4), (5, 3, 2), (64, 64, 1), (640, 360, 4),
                                   (640, 360, 2), (64, 48, 5)])
def test_layout_roundtrip(vacv, oracle, dt, w, h, c):
    b = 2
    src = u8(3, b, h, w, c) if dt == "u8" else f32(3, b, h, w, c)
    chw = host(vacv.layout_change(dev(src), NHWC, NCHW))
    want = np.stack([oracle.hwc_to_chw(src[i], w, h, c) for i in range(b)])
    assert_same(chw, want)
    back = host(vacv.layout_change(dev(chw), NCHW, NHWC))
    assert_same(back, src)


def test_dtype_change(vacv, oracle):
    for n in [176 * 144 * 3, 1920 * 1080 * 3, 1001, 3]:
        src = u8(4, n)
        assert_same(host(vacv.dtype_change(dev(src), vacv.FP32)), oracle.u8_to_f32(src))
        g = (rng(5).random(n, dtype=np.float32) * 255.999).astype(np.float32)
        assert_same(host(vacv.dtype_change(dev(g), vacv.INT8)), oracle.f32_to_u8(g))
    with pytest.raises(vacv.VacvError):   # the reference silently returns garbage here (tensor.cpp:494-499)
        vacv.dtype_change(dev(np.zeros(8, np.float16)), vacv.FP32)


@pytest.mark.parametrize("qpt", [1, 3, 8, 1000])
def test_streaming_kernels_any_grid(vacv, oracle, qpt):
    """The streaming kernels (dtype_change, normalize u8 / fp32, HWC and planes) give every thread STREAM_QPT 16-byte groups of work
    (8 by default); the results must not depend on the grid that follows from it -- one group per thread, a handful of CTAs for the
    whole frame, sizes that are not multiples of anything."""
    assert vacv.lib.vacv_cuda_set_tuning(b"STREAM_QPT", qpt) == 0
    try:
        for n in [176 * 144 * 3, 1920 * 1080 * 3, 4 * 1001, 20]:
            src = u8(4, n)
            assert_same(host(vacv.dtype_change(dev(src), vacv.FP32)), oracle.u8_to_f32(src))
            g = (rng(5).random(n, dtype=np.float32) * 255.999).astype(np.float32)
            assert_same(host(vacv.dtype_change(dev(g), vacv.INT8)), oracle.f32_to_u8(g))
        for (w, h) in [(640, 360), (333, 212), (1920, 1080)]:
            img = u8(17, 2, h, w, 3)
            got = host(vacv.normalize(dev(img), NHWC, dev(MEAN), dev(STD)))
            for i in range(2):
                assert_same(got[i], oracle.normalize(img[i], w * h, 3, NHWC, MEAN, STD))
            planes = np.ascontiguousarray(img.transpose(0, 3, 1, 2))
            gotp = host(vacv.normalize(dev(planes), NCHW, dev(MEAN), dev(STD)))
            for i in range(2):
                assert_same(gotp[i], oracle.normalize(planes[i], w * h, 3, NCHW, MEAN, STD))
            f = img.astype(np.float32)
            gotf = host(vacv.normalize(dev(f), NHWC, dev(MEAN), dev(STD)))
            assert_same(gotf[0], oracle.normalize(f[0], w * h, 3, NHWC, MEAN, STD))
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"STREAM_QPT", 0)


# ------------------------------------------------------------------ a5-a7 bilinear
LIN_SIZES = [((64, 48), (20, 16)), ((64, 48), (200, 111)), ((1920, 1080), (640, 360)), ((1920, 1080), (640, 640)),
             ((333, 211), (500, 300)), ((2, 2), (7, 5)), ((640, 360), (639, 359)), ((2560, 1440), (320, 180))]


# every resize case runs through both implementations: shared-memory tiled (default) and direct global gather
PATHS = [0x200, 0x100]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", LIN_SIZES)
def test_resize_linear_u8(vacv, oracle, layout, sz, path):
    (w, h), (wo, ho) = sz
    b, c = 2, 3
    shape = (b, h, w, c) if layout == NHWC else (b, c, h, w)
    src = u8(6, *shape)
    got = host(vacv.resize(dev(src), layout, wo, ho, vacv.INTER_LINEAR, path))
    want = np.stack([oracle.resize_linear(src[i], w, h, c, layout, wo, ho) for i in range(b)])
    assert_same(got, want)


@pytest.mark.parametrize("signed", [False, True])
@pytest.mark.parametrize("sz", [((1920, 1080), (1280, 720)), ((1920, 1080), (640, 640)), ((1280, 720), (1000, 500)), ((64, 48), (200, 111)),
                                ((640, 360), (639, 359)), ((640, 360), (1536, 700)), ((1280, 720), (1279, 361)), ((3840, 2160), (1920, 1080)),
                                ((960, 540), (320, 180)), ((320, 200), (64, 40)), ((1920, 1080), (640, 216)), ((1920, 1080), (384, 360))])
def test_resize_linear_u8_default_path(vacv, oracle, sz, signed):
    """Default dispatch (no path flag): small vertical ratios run on the persistent TMA kernel (resize_pipe_u8c3.cuh), the rest on
    the gather kernel; both must equal the oracle."""
    (w, h), (wo, ho) = sz
    b = 3
    src = u8(26, b, h, w, 3)
    flags = vacv.FLAG_SIGNED_CHAR if signed else 0
    got = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_LINEAR, flags))
    for i in range(b):
        assert_same(got[i], oracle.resize_linear(src[i], w, h, 3, NHWC, wo, ho, signed_char=int(signed)))


LINEAR_PERIOD_CASES = [
    # rational horizontal scales the periodic walker takes (3:2 / 4:3 / 2:1, rows of whole 16-byte chunks); any vertical scale <= 2
    ((192, 96), (128, 64)),      # 3:2, one partly filled warp strip
    ((480, 100), (320, 77)),     # 3:2, two strips (the second with 8 owning lanes), generic vertical scale
    ((1920, 270), (1280, 180)),  # 3:2, five strips
    ((1920, 100), (1280, 200)),  # 3:2, vertical up-scaling (several output rows per walk step)
    ((384, 64), (256, 64)),      # 3:2, vertical scale 1 (the last two rows share a source row)
    ((384, 128), (256, 64)),     # 3:2, vertical 2:1 (the limit: every source row still used)
    ((384, 700), (256, 600)),    # 3:2, more than one vertical segment per column strip
    ((128, 60), (96, 45)),       # 4:3 both ways
    ((2560, 90), (1920, 77)),    # 4:3, five strips
    ((128, 50), (64, 25)),       # 2:1 both ways
    ((3840, 64), (1920, 40)),    # 2:1, fifteen strips
    ((640, 100), (256, 40)),     # 5:2 both ways (vertical ratio 2.5: one source row in five is walked without being used)
    ((1920, 270), (768, 108)),   # 5:2, three strips
    ((320, 90), (192, 54)),      # 5:3 both ways
    ((1920, 100), (1152, 60)),   # 5:3, three strips
    ((384, 180), (256, 60)),     # 3:2 with vertical 3:1 (the limit of the walk)
]


@pytest.mark.parametrize("signed", [False, True])
@pytest.mark.parametrize("case", range(len(LINEAR_PERIOD_CASES)))
def test_resize_linear_u8_periodic_walker(vacv, oracle, case, signed):
    """u8 BGR bilinear at rational horizontal scales: the periodic walker (adjacent columns per thread, compile-time tap positions,
    bulk-copy ring) against the oracle and against the persistent pipeline (VACV_LINEAR_V=1), both char signedness rules."""
    (w, h), (wo, ho) = LINEAR_PERIOD_CASES[case]
    b = 3
    src = u8(60 + case, b, h, w, 3)
    flags = vacv.FLAG_SIGNED_CHAR if signed else 0
    got = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_LINEAR, flags))
    assert vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 1) == 0
    try:
        pipe = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_LINEAR, flags))
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 0)
    assert_same(got, pipe)
    for i in range(b):
        assert_same(got[i], oracle.resize_linear(src[i], w, h, 3, NHWC, wo, ho, signed_char=int(signed)))


LINEAR_PERIOD_PLANE_CASES = [
    # CHW tensors (plane by plane) at rational horizontal scales: rows of whole 16-byte chunks on both sides
    ((384, 96), (256, 64)),       # 3:2, half a warp strip
    ((1920, 270), (1280, 180)),   # 3:2, 2.5 strips
    ((1920, 100), (1280, 200)),   # 3:2, vertical up-scaling
    ((768, 64), (512, 64)),       # 3:2, vertical scale 1
    ((768, 700), (512, 600)),     # 3:2, several vertical segments
    ((256, 60), (192, 45)),       # 4:3 both ways (24 columns per thread: wo must be a multiple of 48)
    ((2560, 90), (1920, 77)),     # 4:3, wide planes (no pipeline kernel takes 1920 columns)
    ((128, 50), (64, 25)),        # 2:1 both ways
    ((3840, 64), (1920, 40)),     # 2:1, wide planes
    ((640, 100), (256, 40)),      # 5:2 both ways
    ((1920, 90), (1152, 54)),     # 5:3 both ways, 1.5 strips
]


@pytest.mark.parametrize("signed", [False, True])
@pytest.mark.parametrize("case", range(len(LINEAR_PERIOD_PLANE_CASES)))
def test_resize_linear_u8_planes_periodic_walker(vacv, oracle, case, signed):
    """CHW u8 bilinear at rational horizontal scales: the periodic walker with one channel (16 / 24 adjacent columns per thread, one
    PRMT per column pair) against the oracle and against the persistent planes pipeline / gather kernel (VACV_LINEAR_V=1)."""
    (w, h), (wo, ho) = LINEAR_PERIOD_PLANE_CASES[case]
    src = u8(80 + case, 2, 3, h, w)
    flags = vacv.FLAG_SIGNED_CHAR if signed else 0
    got = host(vacv.resize(dev(src), NCHW, wo, ho, vacv.INTER_LINEAR, flags))
    assert vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 1) == 0
    try:
        pipe = host(vacv.resize(dev(src), NCHW, wo, ho, vacv.INTER_LINEAR, flags))
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 0)
    assert_same(got, pipe)
    for i in range(2):
        assert_same(got[i], oracle.resize_linear(src[i], w, h, 3, NCHW, wo, ho, signed_char=int(signed)))
    grey = np.ascontiguousarray(src[:, :1].transpose(0, 2, 3, 1))   # [2, h, w, 1]: single-channel HWC takes the same kernel
    got_g = host(vacv.resize(dev(grey), NHWC, wo, ho, vacv.INTER_LINEAR, flags))
    assert_same(got_g[1, :, :, 0], got[1, 0])


def test_resize_linear_u8_config1_fixture_vs_reference(vacv):
    img = load_fixture("universe1920x1080")
    if img is None or not ref_available():
        pytest.skip("oracle/_ref not staged")
    want = Ref().resize(img, 1920, 1080, 3, NHWC, 640, 360, R_LINEAR)
    got = host(vacv.resize(dev(img[None]), NHWC, 640, 360))[0]
    assert_same(got, want)


@pytest.mark.parametrize("path", PATHS)
def test_resize_linear_u8_flags(vacv, oracle, path):
    w, h, c, wo, ho = 333, 211, 3, 200, 100
    src = u8(7, 1, h, w, c)
    got = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_LINEAR, vacv.FLAG_SIGNED_CHAR | path))[0]
    assert_same(got, oracle.resize_linear(src[0], w, h, c, NHWC, wo, ho, signed_char=1))
    got = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_LINEAR, vacv.FLAG_NEON_RULE | path))[0]
    assert_same(got, oracle.resize_linear_neon_rule(src[0], w, h, c, NHWC, wo, ho))
    chw = np.ascontiguousarray(src.transpose(0, 3, 1, 2))
    got = host(vacv.resize(dev(chw), NCHW, wo, ho, vacv.INTER_LINEAR, vacv.FLAG_NEON_RULE | path))[0]
    assert_same(got, oracle.resize_linear_neon_rule(chw[0], w, h, c, NCHW, wo, ho))


@pytest.mark.parametrize("sz", [((64, 48), (200, 111)), ((640, 360), (213, 120)), ((320, 240), (300, 97))])
def test_resize_linear_u8_planar_flags(vacv, oracle, sz):
    """CHW planes (c = 1 per call, resize.cpp:73-87) through the word-granular single-channel kernel, all three rules."""
    (w, h), (wo, ho) = sz
    chw = u8(12, 1, 3, h, w)
    for flag, want in [(0, oracle.resize_linear(chw[0], w, h, 3, NCHW, wo, ho)),
                       (vacv.FLAG_SIGNED_CHAR, oracle.resize_linear(chw[0], w, h, 3, NCHW, wo, ho, signed_char=1)),
                       (vacv.FLAG_NEON_RULE, oracle.resize_linear_neon_rule(chw[0], w, h, 3, NCHW, wo, ho))]:
        got = host(vacv.resize(dev(chw), NCHW, wo, ho, vacv.INTER_LINEAR, flag | 0x100))[0]
        assert_same(got, want)


@pytest.mark.parametrize("signed", [False, True])
@pytest.mark.parametrize("sz", [((1920, 1080), (640, 360)), ((1920, 1080), (1280, 720)), ((3840, 2160), (1920, 1080)), ((640, 360), (1536, 700)),
                                ((1280, 720), (1276, 361)), ((960, 540), (320, 180)), ((320, 200), (64, 40)), ((1920, 1080), (384, 360)),
                                ((640, 480), (632, 479)), ((3840, 1080), (3072, 1000)), ((1920, 1080), (638, 360))])
def test_resize_linear_u8_planes_default_path(vacv, oracle, sz, signed):
    """CHW tensors (resized plane by plane, resize.cpp:73-87) on the default dispatch: the persistent TMA pipeline for planes --
    quads of four adjacent columns per thread with packed 32-bit stores where w_out % 4 == 0 (one and two quads per thread, one to
    several row groups per CTA, band and row-list staging), the byte-per-lane kernel otherwise -- and the gather kernel for the
    rest.  Batch of 2 frames x 3 planes; both char signedness rules."""
    (w, h), (wo, ho) = sz
    src = u8(27, 2, 3, h, w)
    flags = vacv.FLAG_SIGNED_CHAR if signed else 0
    got = host(vacv.resize(dev(src), NCHW, wo, ho, vacv.INTER_LINEAR, flags))
    for i in range(2):
        assert_same(got[i], oracle.resize_linear(src[i], w, h, 3, NCHW, wo, ho, signed_char=int(signed)))
    # a destination that is not 4-byte aligned must take the byte-per-lane path and still be right
    if wo % 4 == 0 and not signed:
        import torch
        buf = torch.empty(2 * 3 * wo * ho + 8, dtype=torch.uint8, device="cuda")
        dst = buf[1:1 + 2 * 3 * wo * ho]
        d = dev(src)
        rc = vacv.lib.vacv_cuda_resize(d.data_ptr(), dst.data_ptr(), 2, w, h, 3, vacv.INT8, NCHW, wo, ho, vacv.INTER_LINEAR, 0,
                                       torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        assert_same(dst.cpu().numpy().reshape(2, 3, ho, wo), got)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", LIN_SIZES[:5] + [((2560, 1440), (320, 180))])
def test_resize_linear_f32(vacv, oracle, layout, sz, path):
    (w, h), (wo, ho) = sz
    b, c = 1, 3
    shape = (b, h, w, c) if layout == NHWC else (b, c, h, w)
    src = f32(8, *shape)
    got = host(vacv.resize(dev(src), layout, wo, ho, vacv.INTER_LINEAR, path))
    want = np.stack([oracle.resize_linear(src[i], w, h, c, layout, wo, ho) for i in range(b)])
    assert_same(got, want)


def test_resize_same_size_is_copy(vacv):
    src = f32(9, 2, 30, 40, 3)
    assert_same(host(vacv.resize(dev(src), NHWC, 40, 30)), src)


# ------------------------------------------------------------------ a8 / a9 bicubic
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("sz", [((64, 48), (37, 20)), ((64, 48), (20, 37)), ((2560, 1440), (1920, 1080)),
                                ((320, 180), (640, 360)), ((16, 16), (5, 9)), ((64, 48), (100, 100)),
                                ((50, 40), (33, 27)), ((640, 480), (100, 60)), ((1000, 200), (1400, 150))])
@pytest.mark.parametrize("path", PATHS)
def test_resize_cubic_f32(vacv, oracle, layout, sz, path):
    (w, h), (wo, ho) = sz
    c = 3
    shape = (1, h, w, c) if layout == NHWC else (1, c, h, w)
    src = f32(10, *shape)
    got = host(vacv.resize(dev(src), layout, wo, ho, vacv.INTER_CUBIC, path))[0]
    want = oracle.resize_cubic_f32(src[0], w, h, c, layout, wo, ho)
    assert_same(got, want)


@pytest.mark.parametrize("sz", [((1920, 1080), (1280, 720)), ((96, 40), (64, 30)), ((48, 9), (32, 20)), ((1536, 33), (1024, 77)), ((12, 12), (8, 8)),
                                ((2560, 360), (1920, 270)), ((64, 64), (48, 48)), ((1024, 20), (768, 50)), ((16, 16), (12, 12)),
                                ((3840, 540), (1920, 270)), ((64, 17), (32, 40)), ((8, 8), (4, 4)), ((1296, 50), (648, 25))])
@pytest.mark.parametrize("layout", [NCHW, NHWC])
def test_resize_cubic_f32_planes_periodic_walker(vacv, oracle, sz, layout):
    """fp32 bicubic at rational horizontal scales (resize_cubic_f32_period.cuh; planes: 3 : 2 with eight adjacent columns per thread,
    4 : 3 with twelve; interleaved BGR: 4 : 3 with three -- staged stores -- and 3 : 2 with four; the 2 : 1 shapes take the one-column
    walker): full and partial warp strips, a handful of threads per row, the reference's border folding at both image edges (taps that
    move onto other window positions), up- and down-scaling along y, a batch of 2 -- against the oracle (resize_naive.cpp:187-529) and
    against the one-column walker bit for bit."""
    (w, h), (wo, ho) = sz
    src = f32(71, 2, 3, h, w) if layout == NCHW else f32(71, 2, h, w, 3)
    got = host(vacv.resize(dev(src), layout, wo, ho, vacv.INTER_CUBIC))
    for i in range(2):
        assert_same(got[i], oracle.resize_cubic_f32(src[i], w, h, 3, layout, wo, ho))
    assert vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 1) == 0
    try:
        first = host(vacv.resize(dev(src), layout, wo, ho, vacv.INTER_CUBIC))
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 0)
    assert_same(got, first)


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("sz", [((2560, 1440), (1920, 1080)), ((256, 144), (100, 70)), ((176, 144), (640, 640)),
                                ((257, 145), (300, 171)), ((64, 48), (333, 77)), ((64, 48), (21, 13)),
                                ((8, 8), (3, 3)), ((5, 4), (13, 11)), ((100, 60), (75, 45)), ((640, 480), (100, 60)),
                                ((1000, 200), (1401, 150)), ((1920, 1080), (2560, 1440))])
@pytest.mark.parametrize("path", PATHS)
def test_resize_cubic_u8(vacv, oracle, c, sz, path):
    (w, h), (wo, ho) = sz
    if c != 3 and w >= 1000:
        pytest.skip("big case only for c=3")
    src = u8(11 + c, 2, h, w, c)
    got = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_CUBIC, path))
    want = np.stack([oracle.resize_cubic_u8(src[i], w, h, c, wo, ho) for i in range(2)])
    assert_same(got, want)


@pytest.mark.parametrize("variant", [0, 1, 2, 4, 21])
@pytest.mark.parametrize("sz", [((1280, 720), (960, 540)), ((1280, 720), (400, 300)), ((1024, 96), (1000, 90)), ((512, 700), (384, 1000)),
                                ((2048, 64), (1536, 48)), ((768, 300), (1000, 200)), ((256, 256), (129, 255)), ((1920, 1080), (1280, 720))])
def test_resize_cubic_u8_column_walker_generations(vacv, oracle, sz, variant):
    """The u8 bicubic walkers (resize_cubic3_walk.cuh: 2 columns per thread; resize_cubic3_walkn.cuh: 4 or 2 columns per thread, 2..4
    warps per CTA, down-scaling fast path) against the oracle's OpenCV-2.4 rule: partial warp strips, several vertical segments,
    up- and down-scaling in either axis, a batch of 3.  variant = the CUBIC_V tuning switch (0 automatic, 1 first generation)."""
    (w, h), (wo, ho) = sz
    src = u8(40 + variant, 3, h, w, 3)
    assert vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", variant) == 0
    try:
        got = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_CUBIC))
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 0)
    for i in range(3):
        assert_same(got[i], oracle.resize_cubic_u8(src[i], w, h, 3, wo, ho))


@pytest.mark.parametrize("sz", [((2560, 1440), (1920, 1080)), ((1040, 300), (780, 225)), ((1040, 90), (780, 131)), ((64, 40), (48, 30)),
                                ((16, 16), (12, 12)), ((32, 9), (24, 20)), ((4096, 40), (3072, 30)), ((1280, 64), (960, 64)),
                                ((3840, 2160), (1920, 1080)), ((2560, 100), (1280, 77)), ((64, 64), (32, 32)), ((32, 12), (16, 30)),
                                ((1296, 50), (648, 25)), ((1920, 1080), (1280, 720)), ((96, 40), (64, 30)), ((1536, 100), (1024, 77)),
                                ((48, 9), (32, 20)), ((3072, 24), (2048, 16))])
def test_resize_cubic_u8_periodic_walker(vacv, oracle, sz):
    """u8 bicubic at rational horizontal scales (resize_cubic3_period.cuh: 4 : 3 with six adjacent columns per thread, 2 : 1 and 3 : 2 with four -- the latter with the
    per-lane realignment stage):
    full and partial warp strips, a handful of threads per row, the clamped taps at both image edges, up- and down-scaling along y
    (the walk's two emit rules), several vertical segments, a batch of 3 -- against the oracle's OpenCV-2.4 rule and against the
    first-generation walker byte for byte."""
    (w, h), (wo, ho) = sz
    src = u8(63, 3, h, w, 3)
    got = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_CUBIC))
    for i in range(3):
        assert_same(got[i], oracle.resize_cubic_u8(src[i], w, h, 3, wo, ho))
    assert vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 1) == 0
    try:
        first = host(vacv.resize(dev(src), NHWC, wo, ho, vacv.INTER_CUBIC))
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 0)
    assert_same(got, first)


def test_resize_cubic_u8_config4_fixture_vs_bundled_opencv(vacv):
    img = load_fixture("lakers2560x1440")
    if img is None or not ref_available():
        pytest.skip("oracle/_ref not staged")
    want = Ref().cv_resize(img, 2560, 1440, 3, 1920, 1080, R_CUBIC)
    got = host(vacv.resize(dev(img[None]), NHWC, 1920, 1080, vacv.INTER_CUBIC))[0]
    assert_same(got, want)


# ------------------------------------------------------------------ a10 warp affine
M_TEST = [0.849158, 0.012257, -474.827, -0.01225, 0.849158, -379.18]           # test_warp_affine.cpp:31-32
ROT_TEST = dict(scale=1.073914, rot=-3.314525, aux=[738.518372, 537.672852, 204.766998, 73.329681])


def test_host_matrix_helpers(vacv, oracle):
    assert np.array_equal(np.array(vacv.invert_affine(M_TEST), np.float32).view(np.uint32),
                          oracle.invert_affine(M_TEST).view(np.uint32))
    m = np.array(vacv.rotation_matrix(ROT_TEST["scale"], ROT_TEST["rot"], ROT_TEST["aux"]), np.float32)
    assert np.array_equal(m.view(np.uint32), oracle.rotation_matrix(ROT_TEST["scale"], ROT_TEST["rot"], ROT_TEST["aux"]).view(np.uint32))


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("dt", ["u8", "f32"])
def test_warp_affine_reference_matrix(vacv, oracle, layout, dt):
    w, h, c, wo, ho = 1280, 720, 3, 240, 240
    img = load_fixture("t1280x720")
    hwc = img if img is not None else u8(12, h, w, c)
    src = hwc if layout == NHWC else np.ascontiguousarray(hwc.transpose(2, 0, 1))
    if dt == "f32":
        src = src.astype(np.float32)
    minv = np.array(vacv.invert_affine(M_TEST), np.float32)
    got = host(vacv.warp_affine(dev(src[None]), layout, dev(minv[None]), wo, ho))[0]
    want = oracle.warp_affine(src, w, h, c, layout, wo, ho, minv)
    assert_same(got, want)
    if ref_available() and img is not None:
        want_ref, _ = Ref().warp_affine(src, w, h, c, layout, wo, ho, M_TEST)
        assert_same(got, want_ref)


def test_warp_affine_rotation_variant(vacv, oracle):
    w, h, wo, ho = 1280, 720, 140, 210
    img = load_fixture("t1280x720_grey")
    src = img if img is not None else u8(13, h, w, 1)
    minv = np.array(vacv.invert_affine(vacv.rotation_matrix(ROT_TEST["scale"], ROT_TEST["rot"], ROT_TEST["aux"])), np.float32)
    got = host(vacv.warp_affine(dev(src[None]), NHWC, dev(minv[None]), wo, ho))[0]
    assert_same(got, oracle.warp_affine(src, w, h, 1, NHWC, wo, ho, minv))
    if ref_available() and img is not None:
        assert_same(got, Ref().warp_affine_rot(src, w, h, 1, NHWC, wo, ho, ROT_TEST["scale"], ROT_TEST["rot"], ROT_TEST["aux"]))


def random_face_matrices(n, w, h, wo, seed):
    r = rng(seed)
    ms = []
    for _ in range(n):
        s = r.uniform(0.3, 0.6)
        a = np.deg2rad(r.uniform(-15, 15))
        cx, cy = r.uniform(0.3 * w, 0.7 * w), r.uniform(0.3 * h, 0.7 * h)
        al, be = s * np.cos(a), s * np.sin(a)
        ms.append([al, be, wo / 2 - al * cx - be * cy, -be, al, wo / 2 + be * cx - al * cy])
    return ms


def test_warp_affine_batch_with_frame_pool(vacv, oracle):
    w, h, c, wo, ho, nf, n = 320, 200, 3, 112, 112, 4, 24
    frames = u8(14, nf, h, w, c)
    fwd = random_face_matrices(n, w, h, wo, 7) + [[5, 0, 1000, 0, 5, 1000]]   # last one maps fully outside -> zeros
    n += 1
    minv = np.array([vacv.invert_affine(m) for m in fwd], np.float32)
    idx = (np.arange(n) % nf).astype(np.int32)
    got = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx)))
    for i in range(n):
        assert_same(got[i], oracle.warp_affine(frames[idx[i]], w, h, c, NHWC, wo, ho, minv[i]))
    assert not got[-1].any()
    got_sc = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx), vacv.FLAG_SIGNED_CHAR))
    assert_same(got_sc[3], oracle.warp_affine(frames[idx[3]], w, h, c, NHWC, wo, ho, minv[3], signed_char=1))


def test_warp_affine_normalize_fused(vacv, oracle):
    w, h, c, wo, ho, nf, n = 320, 200, 3, 112, 112, 3, 10
    frames = u8(15, nf, h, w, c)
    minv = np.array([vacv.invert_affine(m) for m in random_face_matrices(n, w, h, wo, 8)], np.float32)
    idx = (np.arange(n) % nf).astype(np.int32)
    got = host(vacv.warp_affine_normalize(dev(frames), dev(minv), wo, ho, dev(MEAN), dev(STD), dev(idx)))
    for i in range(n):
        assert_same(got[i], oracle.warp_affine_normalize(frames[idx[i]], w, h, c, minv[i], wo, ho, MEAN, STD))
    got_chw = host(vacv.warp_affine_normalize(dev(frames), dev(minv), wo, ho, dev(MEAN), dev(STD), dev(idx), out_layout=NCHW))
    assert_same(got_chw, np.ascontiguousarray(got.transpose(0, 3, 1, 2)))


def _similarity(s, deg, cx, cy, wo, ho):
    """forward matrix of a crop centred on (cx, cy) of the source: scale s, rotation deg"""
    a = np.deg2rad(deg)
    al, be = s * np.cos(a), s * np.sin(a)
    return [al, be, wo / 2 - al * cx - be * cy, -be, al, ho / 2 + be * cx - al * cy]


STAGED_CASES = [
    # (w, h, wo, ho, forward matrices) -- frames whose rows are multiples of 16 bytes run the TMA-staged kernel
    (320, 200, 112, 112, [_similarity(0.45, 7, 160, 100, 112, 112), _similarity(0.5, -15, 20, 10, 112, 112),        # window leaves the frame
                          _similarity(0.5, 12, 310, 195, 112, 112), _similarity(0.35, 0, 160, 100, 112, 112)]),     # top-left / bottom-right
    (320, 200, 100, 57, [_similarity(0.6, 3, 150, 90, 100, 57), _similarity(0.9, -8, 100, 100, 100, 57)]),           # partial tiles in x and y
    (320, 200, 97, 33, [_similarity(0.7, 5, 150, 90, 97, 33)]),                                                       # odd width: byte stores
    (640, 480, 112, 112, [_similarity(0.1, 30, 320, 240, 112, 112), _similarity(0.25, 45, 320, 240, 112, 112),        # windows too large for a stage
                          _similarity(0.3, 90, 320, 240, 112, 112), _similarity(2.5, 170, 320, 240, 112, 112),        #   -> direct gather inside the kernel
                          [1e-3, 0, 5, 0, 1e-3, 5], [1e30, 0, 0, 0, 1e30, 0]]),                                        # huge / overflowing coordinates
    (1280, 720, 240, 240, [M_TEST, _similarity(1.0, 0, 640, 360, 240, 240), _similarity(1.7, -20, 600, 300, 240, 240)]),
    (336, 64, 64, 40, [_similarity(0.5, 4, 168, 32, 64, 40), [0.5, 0, 0, 0, 0.5, 0]]),                                 # narrow frame, exact half scale
    (328, 200, 112, 112, [_similarity(0.45, 7, 160, 100, 112, 112), _similarity(0.6, -11, 30, 170, 112, 112)]),        # rows not a multiple of 16 bytes: no tensor map, gather kernel
]


@pytest.mark.parametrize("case", range(len(STAGED_CASES)))
def test_warp_affine_staged_windows(vacv, oracle, case):
    """The TMA-staged 3-channel kernel: windows that leave the frame, partial tiles, widths without word-aligned rows, windows
    too large for a stage and non-finite coordinates (both fall back to the gather path per tile) -- u8, fp32 HWC, fp32 CHW,
    and identical to the direct gather kernel."""
    w, h, wo, ho, fwd = STAGED_CASES[case]
    nf = 3
    frames = u8(40 + case, nf, h, w, 3)
    with np.errstate(all="ignore"):
        minv = np.array([vacv.invert_affine(m) for m in fwd], np.float32)
    n = len(fwd)
    idx = (np.arange(n) % nf).astype(np.int32)
    got = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx), vacv.FLAG_TILED))   # staged kernel
    direct = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx)))                # gather kernel (u8 default)
    assert_same(got, direct)
    f32hwc = host(vacv.warp_affine_normalize(dev(frames), dev(minv), wo, ho, dev(MEAN), dev(STD), dev(idx)))
    f32chw = host(vacv.warp_affine_normalize(dev(frames), dev(minv), wo, ho, dev(MEAN), dev(STD), dev(idx), out_layout=NCHW))
    assert_same(f32chw, np.ascontiguousarray(f32hwc.transpose(0, 3, 1, 2)))
    for i in range(n):
        if not np.all(np.isfinite(minv[i])):
            continue   # the oracle's behaviour on non-finite matrices is not defined; the two CUDA kernels agree (above)
        assert_same(got[i], oracle.warp_affine(frames[idx[i]], w, h, 3, NHWC, wo, ho, minv[i]))
        assert_same(f32hwc[i], oracle.warp_affine_normalize(frames[idx[i]], w, h, 3, minv[i], wo, ho, MEAN, STD))
    got_sc = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx), vacv.FLAG_SIGNED_CHAR | vacv.FLAG_TILED))
    assert_same(got_sc[0], oracle.warp_affine(frames[idx[0]], w, h, 3, NHWC, wo, ho, minv[0], signed_char=1))


def test_warp_affine_staged_many_tiles_per_cta(vacv, oracle):
    """More tiles than persistent CTAs (both stages and both barrier parities are reused many times); staged == gather on every crop."""
    w, h, wo, ho, nf, n = 640, 368, 112, 112, 8, 600
    frames = _gpu_rand_u8(41, nf, h, w, 3)
    minv = np.array([vacv.invert_affine(m) for m in random_face_matrices(n, w, h, wo, 12)], np.float32)
    idx = (np.arange(n) % nf).astype(np.int32)
    a = vacv.warp_affine(frames, NHWC, dev(minv), wo, ho, dev(idx), vacv.FLAG_TILED)
    b = vacv.warp_affine(frames, NHWC, dev(minv), wo, ho, dev(idx))
    assert torch.equal(a, b)
    f = vacv.warp_affine_normalize(frames, dev(minv), wo, ho, dev(MEAN), dev(STD), dev(idx))       # staged by default
    for i in (1, 300, 598):
        assert_same(host(f[i]), oracle.warp_affine_normalize(host(frames[idx[i]]), w, h, 3, minv[i], wo, ho, MEAN, STD))
    for i in (0, 299, 599):
        assert_same(host(a[i]), oracle.warp_affine(host(frames[idx[i]]), w, h, 3, NHWC, wo, ho, minv[i]))


PACK_CASES = [
    # (w, h, wo, ho): frames with rows of whole 8-byte words, output widths that are multiples of 4 -> the column-owning pack kernel
    (320, 200, 112, 112),    # two output rows per pass, 7 whole warps
    (320, 200, 100, 57),     # 200 threads in 7 warps (padding lanes), odd row count (half-filled last pass)
    (1280, 720, 240, 240),   # one row per pass, 16 padding lanes
    (336, 64, 64, 40),       # four rows per pass
    (640, 368, 28, 9),       # nine rows per pass, one band
    (328, 200, 112, 112),    # rows of 984 bytes: whole 8-byte words, but no 16-byte rows (no tensor map)
]


@pytest.mark.parametrize("case", range(len(PACK_CASES)))
def test_warp_affine_u8_pack_kernel(vacv, oracle, case):
    """u8 BGR warp_affine, column-owning kernel (64-bit tap loads, shuffle-packed stores): oracle, first-generation gather kernel
    (VACV_WARP_V=1) and signed-char compat on crops that leave the frame on every side."""
    w, h, wo, ho = PACK_CASES[case]
    nf, n = 3, 10
    frames = u8(70 + case, nf, h, w, 3)
    r = np.random.default_rng(700 + case)
    fwd = [_similarity(r.uniform(0.3, 1.4), r.uniform(-40, 40), r.uniform(0, w), r.uniform(0, h), wo, ho) for _ in range(n - 2)]
    fwd += [[1, 0, 0, 0, 1, 0], [0.5, 0, 0, 0, 0.5, 0]]   # identity (every tap on a pixel centre), exact half scale
    minv = np.array([vacv.invert_affine(m) for m in fwd], np.float32)
    idx = (np.arange(n) % nf).astype(np.int32)
    got = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx)))
    got_sc = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx), vacv.FLAG_SIGNED_CHAR))
    for variant in (1, 2):   # 1: first-generation gather kernel; 2: pack kernel with two pixels per thread in flight (default: one)
        assert vacv.lib.vacv_cuda_set_tuning(b"WARP_V", variant) == 0
        try:
            other = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx)))
            other_sc = host(vacv.warp_affine(dev(frames), NHWC, dev(minv), wo, ho, dev(idx), vacv.FLAG_SIGNED_CHAR))
        finally:
            vacv.lib.vacv_cuda_set_tuning(b"WARP_V", 0)
        assert_same(got, other)
        assert_same(got_sc, other_sc)
    for i in range(n):
        assert_same(got[i], oracle.warp_affine(frames[idx[i]], w, h, 3, NHWC, wo, ho, minv[i]))
    for i in (0, n - 1):
        assert_same(got_sc[i], oracle.warp_affine(frames[idx[i]], w, h, 3, NHWC, wo, ho, minv[i], signed_char=1))


def test_warp_affine_grey_and_planar_batches(vacv, oracle):
    """Single-channel frames and CHW frames (c planes, one matrix) through the word-granular single-channel kernel."""
    w, h, wo, ho, nf, n = 320, 200, 112, 100, 3, 9
    minv = np.array([vacv.invert_affine(m) for m in random_face_matrices(n, w, h, wo, 9)], np.float32)
    idx = (np.arange(n) % nf).astype(np.int32)
    grey = u8(24, nf, h, w, 1)
    got = host(vacv.warp_affine(dev(grey), NHWC, dev(minv), wo, ho, dev(idx)))
    for i in range(n):
        assert_same(got[i], oracle.warp_affine(grey[idx[i]], w, h, 1, NHWC, wo, ho, minv[i]))
    got_sc = host(vacv.warp_affine(dev(grey), NHWC, dev(minv), wo, ho, dev(idx), vacv.FLAG_SIGNED_CHAR))
    assert_same(got_sc[2], oracle.warp_affine(grey[idx[2]], w, h, 1, NHWC, wo, ho, minv[2], signed_char=1))
    chw = u8(25, nf, 3, h, w)
    got = host(vacv.warp_affine(dev(chw), NCHW, dev(minv), wo, ho, dev(idx)))
    for i in range(n):
        assert_same(got[i], oracle.warp_affine(chw[idx[i]], w, h, 3, NCHW, wo, ho, minv[i]))
    m1, s1 = np.array([117.0], np.float32), np.array([57.5], np.float32)
    got = host(vacv.warp_affine_normalize(dev(grey), dev(minv), wo, ho, dev(m1), dev(s1), dev(idx)))
    for i in range(n):
        assert_same(got[i], oracle.warp_affine_normalize(grey[idx[i]], w, h, 1, minv[i], wo, ho, m1, s1))


# ------------------------------------------------------------------ a11 / a12 statistics, normalize
@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("w,h,b", [(284, 214, 3), (3840, 2160, 2), (176, 144, 1), (33, 7, 2)])
def test_sums_and_mean_stddev(vacv, oracle, layout, w, h, b):
    c = 3
    shape = (b, h, w, c) if layout == NHWC else (b, c, h, w)
    src = u8(16, *shape)
    want = np.stack([oracle.sums_u8(src[i], w * h, c, layout) for i in range(b)]).astype(np.int64).reshape(b, c, 2)
    got_pf = host(vacv.sums_u8(dev(src), layout, per_frame=True))
    assert_same(got_pf, want)
    got_all = host(vacv.sums_u8(dev(src), layout, per_frame=False))
    assert_same(got_all, want.sum(0, keepdims=True))
    mean, std = vacv.finalize_mean_stddev(dev(got_all), b * w * h)
    m_o, s_o = oracle.finalize_mean_stddev(want.sum(0).astype(np.uint64).ravel(), c, b * w * h)
    assert_same(host(mean)[0], m_o)
    assert_same(host(std)[0], s_o)
    if ref_available() and layout == NHWC and b == 1:
        m_cv, s_cv = Ref().cv_mean_stddev(src[0], w, h, c)   # the truth the reference's own test uses
        assert np.allclose(host(mean)[0], m_cv, rtol=1e-6) and np.allclose(host(std)[0], s_cv, rtol=1e-6)


@pytest.mark.parametrize("layout", [NHWC, NCHW])
@pytest.mark.parametrize("dt", ["u8", "f32"])
@pytest.mark.parametrize("w,h,b", [(176, 144, 2), (284, 214, 1), (33, 7, 1)])
def test_normalize(vacv, oracle, layout, dt, w, h, b):
    c = 3
    shape = (b, h, w, c) if layout == NHWC else (b, c, h, w)
    if dt == "u8" and b > 1 and (w * h * c) % 4:
        pytest.skip("u8 batch needs 4-byte frames")
    src = u8(17, *shape) if dt == "u8" else f32(17, *shape)
    got = host(vacv.normalize(dev(src), layout, dev(MEAN), dev(STD)))
    want = np.stack([oracle.normalize(src[i], w * h, c, layout, MEAN, STD) for i in range(b)])
    assert_same(got, want)
    # per-frame statistics
    means = np.stack([MEAN + i for i in range(b)]).astype(np.float32)
    stds = np.stack([STD + 0.5 * i for i in range(b)]).astype(np.float32)
    got = host(vacv.normalize(dev(src), layout, dev(means), dev(stds), stats_per_frame=True))
    want = np.stack([oracle.normalize(src[i], w * h, c, layout, means[i], stds[i]) for i in range(b)])
    assert_same(got, want)


def test_auto_stats_normalize_matches_reference_semantics(vacv, oracle):
    """va_cv::normalize with empty mean/stddev (normalize.cpp:98-108) = stats pass + apply pass."""
    img = load_fixture("t284x214")
    src = img if img is not None else u8(18, 214, 284, 3)
    d = dev(src[None])
    mean, std = vacv.finalize_mean_stddev(vacv.sums_u8(d, NHWC), 284 * 214)
    got = host(vacv.normalize(d, NHWC, mean[0], std[0]))[0]
    sums = oracle.sums_u8(src, 284 * 214, 3, NHWC)
    m, s = oracle.finalize_mean_stddev(sums, 3, 284 * 214)
    assert_same(got, oracle.normalize(src, 284 * 214, 3, NHWC, m, s))
    if ref_available() and img is not None:
        want = Ref().normalize(src, 284, 214, 3, NHWC)      # reference: sequential-fp32 statistics
        assert np.abs(got - want).max() < 5e-3               # documented deviation (App. C-4), tiny at this size


# ------------------------------------------------------------------ a13 fused pipelines
@pytest.mark.parametrize("v_first", [True, False])
@pytest.mark.parametrize("w,h,wo,ho,b", [(1920, 1080, 640, 640, 2), (640, 360, 224, 224, 3), (64, 48, 100, 37, 2),
                                         (1280, 720, 640, 384, 1), (642, 362, 300, 200, 2), (3840, 2160, 640, 640, 1),
                                         (320, 240, 1000, 700, 1), (1920, 1080, 1919, 1079, 1), (7680, 4320, 1280, 720, 1),
                                         (7680, 4320, 1536, 864, 1), (2560, 1440, 1408, 792, 1)])
def test_fused_pipeline_config2(vacv, oracle, v_first, w, h, wo, ho, b):
    src = u8(19 + w, b, w * h * 3 // 2)
    got = host(vacv.nv_resize_normalize_chw(dev(src), w, h, wo, ho, dev(MEAN), dev(STD), v_first))
    want = oracle.nv_resize_normalize_chw(src, w, h, int(v_first), wo, ho, MEAN, STD, batch=b, threads=4)
    assert_same(got, want)
    # the north star's stated bounds, explicitly
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    assert rel.max() <= 1e-5
    a, bb = got.astype(np.float64).ravel(), want.astype(np.float64).ravel()
    assert a @ bb / np.sqrt((a @ a) * (bb @ bb)) >= 0.99999


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("w,h,yp,cp", [(16, 8, 0, 0), (176, 144, 192, 0), (642, 362, 656, 336), (1920, 1080, 2048, 0), (1920, 1080, 0, 0),
                                       (100, 60, 0, 0)])
def test_cvt_yuv2bgr_surfaces(vacv, oracle, fmt, w, h, yp, cp):
    """Next row 8f-1: pitched NV21 / NV12 and planar I420 / YV12 surfaces -> BGR with the reference's matrix."""
    from test_oracle_vs_ref import make_yuv_surface
    planar = fmt >= 2
    y_pitch = yp or w
    c_pitch = cp or (y_pitch // 2 if planar else y_pitch)
    if not planar:
        c_pitch = max(c_pitch, w)
    b = 2
    per = y_pitch * h + c_pitch * (h // 2) * (2 if planar else 1)
    stride = per + 48
    buf = u8(94, b * stride)
    want = np.empty((b, h, w, 3), np.uint8)
    for i in range(b):
        surf, _ = make_yuv_surface(400 * i + w + fmt, fmt, w, h, y_pitch, c_pitch)
        buf[i * stride:i * stride + per] = surf
        want[i] = oracle.yuv_to_bgr(surf, fmt, w, h, y_pitch, c_pitch)
    got = host(vacv.cvt_yuv2bgr(dev(buf), fmt, w, h, y_pitch=y_pitch, c_pitch=c_pitch, frame_stride=stride, batch=b))
    assert_same(got, want)


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("w,h,yp,cp,wo,ho,b", [(1920, 1080, 2048, 0, 640, 640, 2), (1920, 1080, 0, 0, 640, 640, 1),
                                              (640, 480, 768, 512, 224, 224, 3), (1280, 720, 1280, 0, 1000, 500, 1),
                                              (64, 32, 0, 0, 100, 37, 2), (64, 32, 0, 0, 101, 37, 2), (1920, 1080, 0, 0, 1100, 360, 1),
                                              (3840, 2160, 4096, 0, 640, 384, 1)])
def test_fused_pipeline_yuv_surfaces(vacv, oracle, fmt, half, w, h, yp, cp, wo, ho, b):
    """Next rows 8f-1 / 8f-3: pitched NV21/NV12 and planar I420/YV12 surfaces in, fp32 / fp16 planes out."""
    from test_oracle_vs_ref import make_yuv_surface
    planar = fmt >= 2
    y_pitch = yp or w
    c_pitch = cp or (w // 2 if planar else w)
    if not planar:
        c_pitch = max(c_pitch, y_pitch)
    elif not cp:
        c_pitch = y_pitch // 2
    per = y_pitch * h + c_pitch * (h // 2) * (2 if planar else 1)
    stride = per + 256                                # gap between frames
    buf = u8(91, b * stride)
    want = np.empty((b, 3, ho, wo), np.float32)
    for i in range(b):
        surf, _ = make_yuv_surface(100 * i + w + fmt, fmt, w, h, y_pitch, c_pitch)
        buf[i * stride:i * stride + per] = surf
        bgr = oracle.yuv_to_bgr(surf, fmt, w, h, y_pitch, c_pitch)
        small = oracle.resize_linear(bgr, w, h, 3, NHWC, wo, ho)
        want[i] = oracle.hwc_to_chw(oracle.normalize(small, wo * ho, 3, NHWC, MEAN, STD), wo, ho, 3)
    got = host(vacv.yuv_resize_normalize_chw(dev(buf), fmt, w, h, wo, ho, dev(MEAN), dev(STD), y_pitch=y_pitch,
                                             c_pitch=c_pitch, frame_stride=stride, batch=b, half=half))
    if half:
        assert got.dtype == np.float16
        assert_same(got, want.astype(np.float16))    # fp16 = the exact fp32 result rounded to nearest even
    else:
        assert_same(got, want)
    if y_pitch != w:   # padded surfaces: tensor-map boxes by default (where eligible); one bulk copy per row (VACV_PIPE_ROWS=1) and
        for mode in (1, 2):   # whole bands with their padding (VACV_PIPE_ROWS=2, the round-1 path) must give the same bytes
            assert vacv.lib.vacv_cuda_set_tuning(b"PIPE_ROWS", mode) == 0
            try:
                other = host(vacv.yuv_resize_normalize_chw(dev(buf), fmt, w, h, wo, ho, dev(MEAN), dev(STD), y_pitch=y_pitch,
                                                           c_pitch=c_pitch, frame_stride=stride, batch=b, half=half))
            finally:
                vacv.lib.vacv_cuda_set_tuning(b"PIPE_ROWS", 0)
            assert_same(other, got)


def _bf16_round(x):
    """fp32 -> bfloat16 (round to nearest even) -> fp32, in numpy."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    return r.view(np.float32)


@pytest.mark.parametrize("out_dtype", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("fmt,w,h,cw,ch,content", [(1, 1920, 1080, 640, 640, None), (0, 1280, 720, 640, 384, (10, 6, 600, 338)),
                                                  (2, 640, 480, 416, 416, None), (3, 64, 32, 101, 77, (7, 9, 51, 40)),
                                                  (1, 1080, 1920, 640, 640, None), (0, 1920, 1080, 640, 360, (0, 0, 640, 360))])
def test_fused_pipeline_letterbox(vacv, oracle, out_dtype, fmt, w, h, cw, ch, content):
    """Next row 8f-3: aspect-preserving placement on a padded canvas + fp16 / bf16 planes == unfused chain on the padded canvas."""
    from test_oracle_vs_ref import make_yuv_surface
    b, pad = 2, (114, 100, 7)
    planar = fmt >= 2
    y_pitch = (w + 31) & ~31 if fmt != 1 else w   # aligned decoder pitch (TMA pipeline); dense NV12 1080-wide -> tiled kernel
    c_pitch = y_pitch // 2 if planar else y_pitch
    per = y_pitch * h * 3 // 2
    buf = np.empty(b * per, np.uint8)
    x0, y0, rw, rh = content if content is not None else vacv.letterbox_rect(w, h, cw, ch)
    want = np.empty((b, 3, ch, cw), np.float32)
    for i in range(b):
        surf, _ = make_yuv_surface(200 * i + w + fmt, fmt, w, h, y_pitch, c_pitch)
        buf[i * per:(i + 1) * per] = surf
        bgr = oracle.yuv_to_bgr(surf, fmt, w, h, y_pitch, c_pitch)
        canvas = np.empty((ch, cw, 3), np.uint8)
        canvas[:] = np.array(pad, np.uint8)
        canvas[y0:y0 + rh, x0:x0 + rw] = oracle.resize_linear(bgr, w, h, 3, NHWC, rw, rh)
        want[i] = oracle.hwc_to_chw(oracle.normalize(canvas, cw * ch, 3, NHWC, MEAN, STD), cw, ch, 3)
    dt = {"f32": vacv.FP32, "f16": vacv.FP16, "bf16": vacv.BF16}[out_dtype]
    got = vacv.yuv_letterbox_normalize_chw(dev(buf), fmt, w, h, cw, ch, MEAN, STD, pad_bgr=pad, content=content, y_pitch=y_pitch,
                                           c_pitch=c_pitch, batch=b, out_dtype=dt)
    if out_dtype == "bf16":
        assert_same(got.float().cpu().numpy(), _bf16_round(want))
    elif out_dtype == "f16":
        assert_same(host(got), want.astype(np.float16))
    else:
        assert_same(host(got), want)


def test_letterbox_rect_is_centred_and_aspect_preserving(vacv):
    assert vacv.letterbox_rect(1920, 1080, 640, 640) == (0, 140, 640, 360)
    assert vacv.letterbox_rect(1080, 1920, 640, 640) == (140, 0, 360, 640)
    assert vacv.letterbox_rect(640, 480, 416, 416) == (0, 52, 416, 312)


def test_fused_pipeline_yuv_dense_nv21_equals_base_entry(vacv):
    w, h, wo, ho, b = 1920, 1080, 640, 640, 2
    src = dev(u8(92, b, w * h * 3 // 2))
    a = vacv.nv_resize_normalize_chw(src, w, h, wo, ho, dev(MEAN), dev(STD), True)
    c = vacv.yuv_resize_normalize_chw(src, 0, w, h, wo, ho, dev(MEAN), dev(STD))
    assert_same(host(a), host(c))


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("w,h,yp,cp,wo,ho", [(64, 32, 72, 0, 20, 20), (1080, 1920, 0, 0, 360, 640), (100, 60, 104, 60, 333, 77),
                                            (1920, 1080, 0, 0, 1600, 900)])
def test_fused_pipeline_yuv_tiled_kernel_shapes(vacv, oracle, fmt, half, w, h, yp, cp, wo, ho):
    """Shapes the TMA pipeline does not take (pitch / width not a multiple of 16, w_out > 1536) run on the tiled kernel."""
    from test_oracle_vs_ref import make_yuv_surface
    planar = fmt >= 2
    y_pitch = yp or w
    c_pitch = cp or (y_pitch // 2 if planar else y_pitch)
    if not planar:
        c_pitch = max(c_pitch, w)
    surf, _ = make_yuv_surface(300 + w + fmt, fmt, w, h, y_pitch, c_pitch)
    bgr = oracle.yuv_to_bgr(surf, fmt, w, h, y_pitch, c_pitch)
    small = oracle.resize_linear(bgr, w, h, 3, NHWC, wo, ho)
    want = oracle.hwc_to_chw(oracle.normalize(small, wo * ho, 3, NHWC, MEAN, STD), wo, ho, 3)[None]
    got = host(vacv.yuv_resize_normalize_chw(dev(surf), fmt, w, h, wo, ho, dev(MEAN), dev(STD), y_pitch=y_pitch, c_pitch=c_pitch,
                                             batch=1, half=half))
    assert_same(got, want.astype(np.float16) if half else want)


def test_fused_pipeline_equals_unfused_cuda_chain(vacv):
    w, h, wo, ho, b = 1920, 1080, 640, 640, 2
    src = dev(u8(20, b, w * h * 3 // 2))
    fused = vacv.nv_resize_normalize_chw(src, w, h, wo, ho, dev(MEAN), dev(STD))
    bgr = vacv.cvt_nv2bgr(src, w, h)
    small = vacv.resize(bgr, NHWC, wo, ho)
    norm = vacv.normalize(small, NHWC, dev(MEAN), dev(STD))
    chain = vacv.layout_change(norm, NHWC, NCHW)
    assert_same(host(fused), host(chain))


def test_fused_pipeline_vs_reference_chain(vacv):
    if not ref_available():
        pytest.skip("oracle/_ref not staged")
    w, h, wo, ho = 1920, 1080, 640, 640
    src = u8(21, 1, w * h * 3 // 2)
    want = Ref().pipeline(src[0], w, h, COLOR_YUV2BGR_NV21, wo, ho, MEAN, STD)
    got = host(vacv.nv_resize_normalize_chw(dev(src), w, h, wo, ho, dev(MEAN), dev(STD)))[0]
    assert_same(got, want)


@pytest.mark.parametrize("out_layout", [NHWC, NCHW])
@pytest.mark.parametrize("w,h,wo,ho", [(640, 360, 224, 200), (1920, 1080, 640, 360), (333, 211, 500, 300), (64, 48, 37, 45),
                                       (1920, 1080, 640, 640), (1280, 720, 416, 416), (1920, 1080, 1536, 864), (640, 360, 321, 181),
                                       (1920, 1080, 500, 300)])
def test_resize_normalize_fused(vacv, oracle, out_layout, w, h, wo, ho):
    c, b = 3, 2
    src = u8(22, b, h, w, c)
    got = host(vacv.resize_normalize(dev(src), wo, ho, dev(MEAN), dev(STD), out_layout))
    for i in range(b):
        small = oracle.resize_linear(src[i], w, h, c, NHWC, wo, ho)
        want = oracle.normalize(small, wo * ho, c, NHWC, MEAN, STD)
        if out_layout == NCHW:
            want = oracle.hwc_to_chw(want, wo, ho, c)
        assert_same(got[i], want)


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_config2_properties(vacv):
    """256 x 1080p is too slow for the CPU oracle; check (i) batch independence: every frame equals the same
    frame processed alone, (ii) a flat grey frame maps to the single table value."""
    w, h, wo, ho, b = 1920, 1080, 640, 640, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    src = torch.randint(0, 256, (b, w * h * 3 // 2), dtype=torch.uint8, device="cuda", generator=g)
    src[5] = 128   # Y=128, U=V=128 -> BGR (128,128,128)
    mean, std = dev(MEAN), dev(STD)
    out = vacv.nv_resize_normalize_chw(src, w, h, wo, ho, mean, std)
    for i in (0, 5, 63):
        single = vacv.nv_resize_normalize_chw(src[i:i + 1], w, h, wo, ho, mean, std)
        assert torch.equal(out[i], single[0])
    flat = host(out[5])
    for k in range(3):
        want = np.float32((np.float32(128.0) - MEAN[k]).astype(np.float64) / (np.float64(STD[k]) + 1e-6))
        assert np.all(flat[k] == want)


def test_full_size_config2_batch256_sampled_vs_oracle(vacv, oracle):
    """BASELINE config 2 at its full size (256 x 1080p NV12 -> 640x640 CHW fp32): sampled frames against the oracle."""
    w, h, wo, ho, b = 1920, 1080, 640, 640, 256
    src = _gpu_rand_u8(1, b, w * h * 3 // 2)
    out = vacv.nv_resize_normalize_chw(src, w, h, wo, ho, dev(MEAN), dev(STD))
    for i in (0, 131, 255):
        want = oracle.nv_resize_normalize_chw(host(src[i]), w, h, 1, wo, ho, MEAN, STD)
        assert_same(host(out[i]), want)


def test_full_size_config1_batch256_sampled_vs_oracle(vacv, oracle):
    """BASELINE config 1 shape, batch 256: sampled frames against the oracle; the batch equals frame-by-frame processing."""
    w, h, wo, ho, b = 1920, 1080, 640, 360, 256
    src = _gpu_rand_u8(2, b, h, w, 3)
    out = vacv.resize(src, NHWC, wo, ho)
    for i in (0, 100, 255):
        assert_same(host(out[i]), oracle.resize_linear(host(src[i]), w, h, 3, NHWC, wo, ho))
        assert torch.equal(out[i], vacv.resize(src[i:i + 1], NHWC, wo, ho)[0])


def test_full_size_config3_4096_crops_sampled_vs_oracle(vacv, oracle):
    """BASELINE config 3 at its full size: 4096 face crops from a pool of 720p frames; sampled crops against the oracle."""
    w, h, wo, ho, nf, n = 1280, 720, 112, 112, 64, 4096
    frames = _gpu_rand_u8(3, nf, h, w, 3)
    minv = np.array([vacv.invert_affine(m) for m in random_face_matrices(n, w, h, wo, 11)], np.float32)
    idx = (np.arange(n) % nf).astype(np.int32)
    out = vacv.warp_affine_normalize(frames, dev(minv), wo, ho, dev(MEAN), dev(STD), dev(idx))
    assert tuple(out.shape) == (n, ho, wo, 3)
    for i in (0, 777, 2048, 4095):
        want = oracle.warp_affine_normalize(host(frames[idx[i]]), w, h, 3, minv[i], wo, ho, MEAN, STD)
        assert_same(host(out[i]), want)


def test_full_size_u8_warp_pack_kernel_4096_crops(vacv, oracle):
    """Config 3's crops with u8 output at full size: the column-owning pack kernel equals the first-generation gather kernel on every
    crop (device-side comparison) and the oracle on sampled crops."""
    w, h, wo, ho, nf, n = 1280, 720, 112, 112, 64, 4096
    frames = _gpu_rand_u8(13, nf, h, w, 3)
    minv = np.array([vacv.invert_affine(m) for m in random_face_matrices(n, w, h, wo, 17)], np.float32)
    idx = (np.arange(n) % nf).astype(np.int32)
    out = vacv.warp_affine(frames, NHWC, dev(minv), wo, ho, dev(idx))
    assert vacv.lib.vacv_cuda_set_tuning(b"WARP_V", 1) == 0
    try:
        first = vacv.warp_affine(frames, NHWC, dev(minv), wo, ho, dev(idx))
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"WARP_V", 0)
    assert torch.equal(out, first)
    for i in (0, 1234, 4095):
        assert_same(host(out[i]), oracle.warp_affine(host(frames[idx[i]]), w, h, 3, NHWC, wo, ho, minv[i]))


@pytest.mark.parametrize("sz", [((1920, 1080), (1280, 720), 64), ((2560, 1440), (1920, 1080), 32), ((3840, 2160), (1920, 1080), 16)])
def test_full_size_u8_bilinear_periodic_walker(vacv, oracle, sz):
    """bench_ops' rational-scale bilinear shapes at full size: the periodic walker equals the persistent pipeline on the whole batch
    (device-side comparison), one frame equals the oracle, and the batch equals frame-by-frame processing."""
    (w, h), (wo, ho), b = sz
    src = _gpu_rand_u8(14, b, h, w, 3)
    out = vacv.resize(src, NHWC, wo, ho)
    assert vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 1) == 0
    try:
        pipe = vacv.resize(src, NHWC, wo, ho)
    finally:
        vacv.lib.vacv_cuda_set_tuning(b"LINEAR_V", 0)
    assert torch.equal(out, pipe)
    assert_same(host(out[b - 1]), oracle.resize_linear(host(src[b - 1]), w, h, 3, NHWC, wo, ho))
    assert torch.equal(out[3], vacv.resize(src[3:4], NHWC, wo, ho)[0])


def test_fused_pipeline_padded_surfaces_rotating_pools(vacv):
    """Tensor-map staging binds the maps to a surface pool's address: a caller that rotates through more pools than the plan keeps
    encoded (4) must get every pool's own pixels, in any order and with changing batch sizes."""
    w, h, wo, ho, pitch = 1280, 720, 416, 416, 1536
    mean, std = dev(MEAN), dev(STD)
    pools, wants = [], []
    for i in range(6):
        b = 1 + i % 3
        dense = _gpu_rand_u8(200 + i, b, h * 3 // 2, w)
        padded = torch.full((b, h * 3 // 2, pitch), 0x5A, dtype=torch.uint8, device="cuda")
        padded[:, :, :w] = dense
        pools.append((padded, b))
        wants.append(vacv.nv_resize_normalize_chw(dense.reshape(b, -1), w, h, wo, ho, mean, std, False))
    for i in (0, 1, 2, 3, 4, 5, 0, 3, 5, 1, 4, 2, 2, 0):
        padded, b = pools[i]
        got = vacv.yuv_resize_normalize_chw(padded.reshape(-1), vacv.YUV_NV12, w, h, wo, ho, mean, std, y_pitch=pitch, c_pitch=pitch, batch=b)
        assert torch.equal(got, wants[i])
    got = vacv.yuv_resize_normalize_chw(pools[2][0].reshape(-1), vacv.YUV_NV12, w, h, wo, ho, mean, std, y_pitch=pitch, c_pitch=pitch, batch=1)
    assert torch.equal(got[0], wants[2][0])   # same pool, smaller batch: the maps' frame extent follows


def test_full_size_config2_padded_surfaces_three_staging_forms(vacv):
    """Config 2 on 256 pitch-2048 NV12 surfaces: tensor-map boxes (default), whole bands with their padding and per-row copies give the
    same bytes, and those equal the dense path on the same frames re-packed without padding."""
    w, h, wo, ho, b, pitch = 1920, 1080, 640, 640, 256, 2048
    dense = _gpu_rand_u8(15, b, h * 3 // 2, w)
    padded = torch.zeros((b, h * 3 // 2, pitch), dtype=torch.uint8, device="cuda")
    padded[:, :, :w] = dense
    padded[:, :, w:] = 0xA5   # the padding must never reach the result
    mean, std = dev(MEAN), dev(STD)
    want = vacv.nv_resize_normalize_chw(dense.reshape(b, -1), w, h, wo, ho, mean, std, False)   # v_first = False: NV12
    outs = []
    for mode in (0, 2, 1):
        assert vacv.lib.vacv_cuda_set_tuning(b"PIPE_ROWS", mode) == 0
        try:
            outs.append(vacv.yuv_resize_normalize_chw(padded.reshape(-1), vacv.YUV_NV12, w, h, wo, ho, mean, std, y_pitch=pitch, c_pitch=pitch, batch=b))
        finally:
            vacv.lib.vacv_cuda_set_tuning(b"PIPE_ROWS", 0)
    for o in outs:
        assert torch.equal(o, want)


def test_full_size_config4_batch128_properties(vacv, oracle):
    """BASELINE config 4 at its full size (128 x 2560x1440 -> 1920x1080 u8 bicubic): one frame against the oracle (OpenCV-2.4
    rule), and the batch result is independent of the batch position (the same frame replicated gives identical outputs)."""
    w, h, wo, ho, b = 2560, 1440, 1920, 1080, 128
    one = _gpu_rand_u8(4, 1, h, w, 3)
    src = one.expand(b, h, w, 3).contiguous()
    src[77] = _gpu_rand_u8(5, h, w, 3)
    out = vacv.resize(src, NHWC, wo, ho, vacv.INTER_CUBIC)
    assert_same(host(out[0]), oracle.resize_cubic_u8(host(one[0]), w, h, 3, wo, ho))
    for i in (1, 64, 127):
        assert torch.equal(out[i], out[0])
    assert_same(host(out[77]), oracle.resize_cubic_u8(host(src[77]), w, h, 3, wo, ho))


def test_full_size_config5_per_gpu_shard_properties(vacv, oracle):
    """BASELINE config 5, one GPU's shard at full size (128 x 3840x2160 BGR u8): the exact sums are additive over frames
    (sum of per-frame sums == batch sums == numpy's exact integer sums of sampled frames), and sampled frames of the
    normalised output equal the oracle's normalise with the batch-global statistics."""
    w, h, b = 3840, 2160, 128
    src = _gpu_rand_u8(6, b, h, w, 3)
    per_frame = vacv.sums_u8(src, NHWC, per_frame=True)            # [b, 3, 2] int64
    total = vacv.sums_u8(src, NHWC, per_frame=False)               # [1, 3, 2]
    assert torch.equal(per_frame.sum(0, keepdim=True), total)
    for i in (0, 127):
        f = host(src[i]).astype(np.int64).reshape(-1, 3)
        assert np.array_equal(host(per_frame[i])[:, 0], f.sum(0)) and np.array_equal(host(per_frame[i])[:, 1], (f * f).sum(0))
    mean, std = vacv.finalize_mean_stddev(total, b * w * h)
    m, s = oracle.finalize_mean_stddev(host(total)[0].astype(np.uint64).ravel(), 3, b * w * h)
    assert_same(host(mean)[0], m)
    assert_same(host(std)[0], s)
    out = vacv.normalize(src, NHWC, mean[0], std[0])
    for i in (3, 90):
        assert_same(host(out[i]), oracle.normalize(host(src[i]), w * h, 3, NHWC, m, s))


# ------------------------------------------------------------------ the C++ drop-in (libvacv.so) end to end
def test_cpp_dropin_binary(vacv):
    """tests/cpp/test_dropin: a reference-style C++ caller linked against libvacv.so, checked against the oracle."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "test_dropin")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr


# ------------------------------------------------------------------ CUDA path vs the committed golden digests
def test_cuda_vs_golden_digests(vacv):
    """sha256 of CUDA outputs == digests the compiled reference produced (tests/golden/golden.json)."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import cases as gc
    golden = json.load(open(os.path.join(here, "golden", "golden.json")))
    img, imgf = gc.u8(2, 360, 640, 3), gc.f32(3, 360, 640, 3)
    big, grey = gc.u8(4, 720, 1280, 3), gc.u8(5, 720, 1280, 1)
    nv, nv2 = gc.u8(1, 640 * 360 * 3 // 2), gc.u8(6, 1920 * 1080 * 3 // 2)
    minv = dev(np.array(vacv.invert_affine(gc.M_TEST), np.float32)[None])
    mrot = dev(np.array(vacv.invert_affine(vacv.rotation_matrix(gc.ROT["scale"], gc.ROT["rot"], gc.ROT["aux"])), np.float32)[None])
    mean, std = dev(gc.MEAN), dev(gc.STD)
    got = {
        "nv21_to_bgr_640x360": vacv.cvt_nv2bgr(dev(nv[None]), 640, 360, True),
        "crop_hwc_u8": vacv.crop(dev(img[None]), NHWC, 7, 3, 192, 96),
        "hwc_to_chw_u8": vacv.layout_change(dev(img[None]), NHWC, NCHW),
        "u8_to_f32": vacv.dtype_change(dev(img[None]), vacv.FP32),
        "f32_to_u8": vacv.dtype_change(dev(imgf[None]), vacv.INT8),
        "resize_linear_u8_hwc_320x180": vacv.resize(dev(img[None]), NHWC, 320, 180),
        "resize_linear_u8_chw_213x97": vacv.resize(dev(np.ascontiguousarray(img.transpose(2, 0, 1))[None]), NCHW, 213, 97),
        "resize_linear_u8_neon_source_chw_213x97": vacv.resize(dev(np.ascontiguousarray(img.transpose(2, 0, 1))[None]), NCHW, 213, 97,
                                                               vacv.INTER_LINEAR, vacv.FLAG_NEON_RULE),
        "resize_linear_u8_neon_source_hwc_320x180": vacv.resize(dev(img[None]), NHWC, 320, 180, vacv.INTER_LINEAR, vacv.FLAG_NEON_RULE),
        "resize_linear_f32_hwc_500x300": vacv.resize(dev(imgf[None]), NHWC, 500, 300),
        "resize_cubic_f32_hwc_300x300": vacv.resize(dev(imgf[None]), NHWC, 300, 300, vacv.INTER_CUBIC),
        "resize_cubic_f32_hwc_480x270_fixed": vacv.resize(dev(imgf[None]), NHWC, 480, 270, vacv.INTER_CUBIC),
        "resize_cubic_u8_cv24_480x270": vacv.resize(dev(img[None]), NHWC, 480, 270, vacv.INTER_CUBIC),
        "resize_cubic_u8_cv24_up_803x451": vacv.resize(dev(img[None]), NHWC, 803, 451, vacv.INTER_CUBIC),
        "warp_affine_u8_240x240": vacv.warp_affine(dev(big[None]), NHWC, minv, 240, 240),
        "warp_affine_f32_240x240": vacv.warp_affine(dev(big.astype(np.float32)[None]), NHWC, minv, 240, 240),
        "warp_affine_rot_grey_140x210": vacv.warp_affine(dev(grey[None]), NHWC, mrot, 140, 210),
        "normalize_u8_hwc": vacv.normalize(dev(img[None]), NHWC, mean, std),
        "pipeline_c2_1080p_to_640x640": vacv.nv_resize_normalize_chw(dev(nv2[None]), 1920, 1080, 640, 640, mean, std),
    }
    bad = [k for k, v in got.items() if gc.digest(host(v)[0]) != golden[k]]
    assert not bad, f"CUDA output differs from the reference digest for: {bad}"


def test_fused_pipeline_is_cuda_graph_capturable(vacv):
    """After the first call of a shape the entry point only launches (the plan is cached): it can be captured and replayed."""
    w, h, wo, ho, b = 1280, 720, 640, 384, 4
    mean, std = dev(MEAN), dev(STD)
    src = _gpu_rand_u8(31, b, w * h * 3 // 2)
    out = torch.empty((b, 3, ho, wo), dtype=torch.float32, device="cuda")
    vacv.nv_resize_normalize_chw(src, w, h, wo, ho, mean, std, True, out=out)          # builds the plan
    want = out.clone()
    out.zero_()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            vacv.nv_resize_normalize_chw(src, w, h, wo, ho, mean, std, True, out=out)
    torch.cuda.current_stream().wait_stream(side)
    src2 = _gpu_rand_u8(32, b, w * h * 3 // 2)
    src.copy_(src2)                                                                   # new input, same buffers
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, vacv.nv_resize_normalize_chw(src2, w, h, wo, ho, mean, std, True))
    assert not torch.equal(out, want)


def test_host_buffer_pipeline_entry(vacv, oracle):
    """vacv_cuda_nv_resize_normalize_chw_host: host pointers in/out, chunked + pipelined inside the library."""
    w, h, wo, ho, b = 640, 360, 224, 224, 7
    src = u8(23, b, w * h * 3 // 2)
    h_in = torch.from_numpy(src).pin_memory()
    h_out = torch.empty((b, 3, ho, wo), dtype=torch.float32).pin_memory()
    for chunk in (3, 2, 16):   # ragged last chunk, more than two chunks (buffer reuse), single chunk
        h_out.zero_()
        vacv.nv_resize_normalize_chw_host(h_in, h_out, w, h, wo, ho, MEAN, STD, True, chunk)
        want = oracle.nv_resize_normalize_chw(src, w, h, 1, wo, ho, MEAN, STD, batch=b, threads=4)
        assert_same(h_out.numpy(), want)
    pageable = torch.from_numpy(src.copy())          # pageable host memory also works
    out2 = torch.empty((b, 3, ho, wo), dtype=torch.float32)
    vacv.nv_resize_normalize_chw_host(pageable, out2, w, h, wo, ho, MEAN, STD, True, 4)
    assert_same(out2.numpy(), want)


def test_host_buffer_yuv_entry(vacv, oracle):
    """vacv_cuda_yuv_normalize_chw_host: pitched NV12 surfaces in host memory -> fp16 letterboxed planes in host memory."""
    from test_oracle_vs_ref import make_yuv_surface
    w, h, cw, ch, b, yp = 640, 360, 416, 416, 5, 704
    per = yp * h * 3 // 2
    buf = np.empty(b * per, np.uint8)
    x0, y0, rw, rh = vacv.letterbox_rect(w, h, cw, ch)
    want = np.empty((b, 3, ch, cw), np.float32)
    want_plain = np.empty((b, 3, ch, cw), np.float32)
    for i in range(b):
        surf, _ = make_yuv_surface(500 + i, 1, w, h, yp, yp)
        buf[i * per:(i + 1) * per] = surf
        bgr = oracle.yuv_to_bgr(surf, 1, w, h, yp, yp)
        canvas = np.full((ch, cw, 3), 114, np.uint8)
        canvas[y0:y0 + rh, x0:x0 + rw] = oracle.resize_linear(bgr, w, h, 3, NHWC, rw, rh)
        want[i] = oracle.hwc_to_chw(oracle.normalize(canvas, cw * ch, 3, NHWC, MEAN, STD), cw, ch, 3)
        want_plain[i] = oracle.hwc_to_chw(oracle.normalize(oracle.resize_linear(bgr, w, h, 3, NHWC, cw, ch), cw * ch, 3, NHWC, MEAN, STD), cw, ch, 3)
    h_in = torch.from_numpy(buf).pin_memory()
    h_out = torch.empty((b, 3, ch, cw), dtype=torch.float16).pin_memory()
    vacv.yuv_normalize_chw_host(h_in, h_out, vacv.YUV_NV12, w, h, cw, ch, MEAN, STD, content=(x0, y0, rw, rh), y_pitch=yp, c_pitch=yp,
                                batch=b, out_dtype=vacv.FP16, chunk_frames=2)
    assert_same(h_out.numpy(), want.astype(np.float16))
    h_out32 = torch.empty((b, 3, ch, cw), dtype=torch.float32)
    vacv.yuv_normalize_chw_host(h_in, h_out32, vacv.YUV_NV12, w, h, cw, ch, MEAN, STD, y_pitch=yp, c_pitch=yp, batch=b, chunk_frames=3)
    assert_same(h_out32.numpy(), want_plain)


# ------------------------------------------------------------------ the reference's OWN test-suite on the drop-in
def test_reference_test_suite_links_and_passes_on_dropin(vacv):
    """oracle/_ref/va_cv_ut_b200 = the reference's src/test sources, compiled unmodified, linked against libvacv.so.
    Every case must score at least what the same suite scores against the reference itself (va_cv_ut_ref)."""
    import os
    import subprocess
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    b200, ref = os.path.join(ref_dir, "va_cv_ut_b200"), os.path.join(ref_dir, "va_cv_ut_ref")
    if not (os.path.exists(b200) and os.path.exists(ref) and os.path.isdir(os.path.join(ref_dir, "res"))):
        pytest.skip("oracle/_ref test-suite binaries not staged (make -C oracle ref)")

    def run(exe):
        out = subprocess.run([exe], cwd=ref_dir, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return {l.split()[1]: float(l.split()[2]) for l in out.stdout.splitlines() if l.startswith("CASE ")}

    ours, theirs = run(b200), run(ref)
    assert len(ours) == 36 and ours.keys() == theirs.keys()
    passed = lambda d: sum(abs(v - 1.0) <= 5e-4 for v in d.values())      # the reference's criterion (cv_profile.cpp:10,108)
    worse = {k: (ours[k], theirs[k]) for k in ours if abs(ours[k] - 1.0) > abs(theirs[k] - 1.0) + 2e-4}
    print(f"reference suite: drop-in passes {passed(ours)}/36, reference itself passes {passed(theirs)}/36")
    assert not worse, f"cases where the drop-in scores worse than the reference: {worse}"
    assert passed(ours) >= passed(theirs)


@pytest.mark.parametrize("layout", [NHWC, NCHW])
def test_sums_f32_statistics(vacv, oracle, layout):
    """fp32 pixels (integer-valued, as after change_dtype): fp64 sums are exact -> same statistics as the u8 path."""
    w, h, c, b = 284, 214, 3, 2
    shape = (b, h, w, c) if layout == NHWC else (b, c, h, w)
    src8 = u8(24, *shape)
    src = src8.astype(np.float32)
    sums = host(vacv.sums_f32(dev(src), layout, per_frame=True))
    for i in range(b):
        want = oracle.sums_f32(src[i], w * h, c, layout)
        assert np.array_equal(sums[i].ravel(), want)
        assert np.array_equal(want, oracle.sums_u8(src8[i], w * h, c, layout).astype(np.float64))
    mean, std = vacv.finalize_mean_stddev_f64(vacv.sums_f32(dev(src), layout), b * w * h)
    tot = sum(oracle.sums_f32(src[i], w * h, c, layout) for i in range(b))
    m_o, s_o = oracle.finalize_mean_stddev_f64(tot, c, b * w * h)
    assert_same(host(mean)[0], m_o)
    assert_same(host(std)[0], s_o)
