#!/usr/bin/env python
"""Times the u8 bicubic kernel variants (tuning knob CUBIC_V) on config 4 and checks each against the first-generation walker
(CUBIC_V=1, the kernel the parity tests pinned first) byte for byte.   python profiles/_c4_variants.py 0 5 6 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vacv_b200 as vacv  # noqa: E402
from bench_ops import rand_u8, timeit  # noqa: E402

variants = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(0,)]   # CUBIC_V[:WALK_SEGS]
shapes = [((2560, 1440), (1920, 1080), 128), ((1920, 1080), (1280, 720), 64), ((1920, 1080), (1000, 700), 32), ((3840, 2160), (1920, 1080), 32)][:int(os.environ.get("C4_SHAPES", "4"))]
for (w, h), (wo, ho), b in shapes:
    src = rand_u8(b, h, w, 3)
    vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 1)
    want = vacv.resize(src, vacv.NHWC, wo, ho, vacv.INTER_CUBIC).clone()
    for var in variants:
        v, segs = var[0], (var[1] if len(var) > 1 else 0)
        vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", v)
        vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", segs)
        got = vacv.resize(src, vacv.NHWC, wo, ho, vacv.INTER_CUBIC)
        same = bool(torch.equal(got, want))
        ms, mn = timeit(lambda: vacv.resize(src, vacv.NHWC, wo, ho, vacv.INTER_CUBIC), 30)
        gbs = b * (w * h * 3 + wo * ho * 3) / (ms * 1e-3) / 1e9
        print(f"{w}x{h}->{wo}x{ho} x{b} CUBIC_V={v} WALK_SEGS={segs}: {ms:.4f} ms (min {mn:.4f})  {gbs:.0f} GB/s  bit-exact vs V=1: {same}", flush=True)
    vacv.lib.vacv_cuda_set_tuning(b"CUBIC_V", 0)
    vacv.lib.vacv_cuda_set_tuning(b"WALK_SEGS", 0)
