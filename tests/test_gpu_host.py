"""Host-layer behaviour of libvacv_cuda.so that real callers depend on: one thread alternating shapes (LRU plan caches) and
GPUs (per-device pipelines; the reference's hook is CudaDevice::set_device, src/cv/cuda_device.cu:15-18), the tuning switches,
and the host-buffer entry on a surface pool that ends right after its last surface."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MEAN = [103.53, 116.28, 123.675]
STD = [57.375, 57.12, 58.395]


def _want(oracle, src, w, h, wo, ho):
    return oracle.nv_resize_normalize_chw(src, w, h, 1, wo, ho, np.array(MEAN, np.float32), np.array(STD, np.float32), batch=src.shape[0])


def test_one_thread_alternates_shapes_and_devices(oracle):
    """Fused device entry + host-buffer entry, two source shapes x two output shapes, alternating cuda:0 / cuda:1 when present."""
    import vacv_b200 as vacv
    n_dev = torch.cuda.device_count()
    rng = np.random.default_rng(3)
    cases = []
    for (w, h, wo, ho) in [(640, 360, 224, 224), (320, 240, 160, 96), (640, 480, 320, 320)]:
        src = rng.integers(0, 256, (2, w * h * 3 // 2), dtype=np.uint8)
        cases.append((w, h, wo, ho, src, _want(oracle, src, w, h, wo, ho)))
    for it in range(4 * len(cases)):
        dev = it % n_dev if n_dev >= 2 else 0
        assert vacv.lib.vacv_cuda_set_device(dev) == 0
        torch.cuda.set_device(dev)
        w, h, wo, ho, src, want = cases[it % len(cases)]
        mean = torch.tensor(MEAN, device=f"cuda:{dev}")
        std = torch.tensor(STD, device=f"cuda:{dev}")
        out = vacv.nv_resize_normalize_chw(torch.from_numpy(src).to(f"cuda:{dev}"), w, h, wo, ho, mean, std)
        assert out.device.index == dev
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), (it, dev)
        h_out = torch.empty((2, 3, ho, wo), dtype=torch.float32).pin_memory()
        vacv.nv_resize_normalize_chw_host(torch.from_numpy(src).pin_memory(), h_out, w, h, wo, ho, MEAN, STD, True, 1)
        assert np.array_equal(h_out.numpy().view(np.uint32), want.view(np.uint32)), (it, dev, "host entry")
    vacv.lib.vacv_cuda_set_device(0)
    torch.cuda.set_device(0)


def test_resize_and_warp_plans_survive_alternation(oracle):
    """Bilinear pipeline and TMA-staged warp: results stay right when shapes / frame pools alternate on one thread."""
    import vacv_b200 as vacv
    from oracle_lib import NHWC
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (2, 360, 640, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (1, 240, 320, 3), dtype=np.uint8)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    minv = np.array([[2.0, 0.1, 5.0, -0.1, 2.0, 7.0]], np.float32)
    tm = torch.from_numpy(minv).cuda()
    mean, std = torch.tensor(MEAN, device="cuda"), torch.tensor(STD, device="cuda")
    for it in range(6):
        src, t, (h, w) = (a, ta, (360, 640)) if it % 2 == 0 else (b, tb, (240, 320))
        wo, ho = (w // 2, h // 2) if it % 3 else (w * 2 // 3, h * 2 // 3)
        got = vacv.resize(t, vacv.NHWC, wo, ho).cpu().numpy()
        want = oracle.resize_linear(src[0], w, h, 3, NHWC, wo, ho)
        assert np.array_equal(got[0].ravel(), want.ravel()), it
        gw = vacv.warp_affine_normalize(t, tm, 64, 64, mean, std).cpu().numpy()
        ww = oracle.warp_affine_normalize(src[0], w, h, 3, minv[0], 64, 64, np.array(MEAN, np.float32), np.array(STD, np.float32))
        assert np.array_equal(gw[0].view(np.uint32).ravel(), ww.view(np.uint32).ravel()), it


def test_tuning_switch_changes_the_kernel_not_the_result():
    """vacv_cuda_set_tuning replaces the getenv calls of round 1: same bits from both warp kernels, unknown names rejected."""
    import vacv_b200 as vacv
    frames = torch.randint(0, 256, (2, 720, 1280, 3), dtype=torch.uint8, device="cuda")
    minv = torch.tensor([[2.2, 0.2, 100.0, -0.2, 2.2, 50.0], [1.7, -0.3, 300.0, 0.3, 1.7, 90.0]], device="cuda")
    mean, std = torch.tensor(MEAN, device="cuda"), torch.tensor(STD, device="cuda")
    staged = vacv.warp_affine_normalize(frames, minv, 112, 112, mean, std)
    assert vacv.lib.vacv_cuda_set_tuning(b"WARP_GATHER", 1) == 0
    try:
        gather = vacv.warp_affine_normalize(frames, minv, 112, 112, mean, std)
    finally:
        assert vacv.lib.vacv_cuda_set_tuning(b"WARP_GATHER", 0) == 0
    assert torch.equal(staged, gather)
    assert vacv.lib.vacv_cuda_set_tuning(b"NO_SUCH_KNOB", 1) == -1


def test_host_entry_reads_no_byte_past_the_last_surface(oracle):
    """A pitched decoder pool of (batch-1) * frame_stride + surface bytes is legal: the last chunk's H2D copy must stop at the
    end of the last surface.  The pool is placed at the very end of a pinned allocation."""
    import vacv_b200 as vacv
    w, h, wo, ho, batch = 320, 240, 160, 128, 3
    pitch = 384
    surface = pitch * h + pitch * (h // 2)
    stride = surface + 4096
    total = (batch - 1) * stride + surface
    rng = np.random.default_rng(2)
    pool = torch.empty(total, dtype=torch.uint8).pin_memory()
    pool.numpy()[:] = rng.integers(0, 256, total, dtype=np.uint8)
    dense = np.empty((batch, w * h * 3 // 2), np.uint8)
    p = pool.numpy()
    for i in range(batch):
        s = p[i * stride:i * stride + surface]
        y = s[:pitch * h].reshape(h, pitch)[:, :w]
        c = s[pitch * h:].reshape(h // 2, pitch)[:, :w]
        dense[i] = np.concatenate([y.ravel(), c.ravel()])
    h_out = torch.empty((batch, 3, ho, wo), dtype=torch.float32).pin_memory()
    vacv.yuv_normalize_chw_host(pool, h_out, vacv.YUV_NV21, w, h, wo, ho, MEAN, STD, y_pitch=pitch, c_pitch=pitch, frame_stride=stride,
                                batch=batch, chunk_frames=2)
    want = _want(oracle, dense, w, h, wo, ho)
    assert np.array_equal(h_out.numpy().view(np.uint32), want.view(np.uint32))
