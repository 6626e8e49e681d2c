"""ncu / timing target (not a test): the CTA-pair implicit-GEMM kernel on the four heaviest shapes of a sampling step
(profiles/README.md) plus the 8x8 level that runs 20-40 tiles on 74 CTA pairs.
    python tests/prof_igemm_shapes.py
    ncu --set full --clock-control none --import-source on -k regex:igemm -c 6 -o gpurun_out/r2_igemm2_full \
        python tests/prof_igemm_shapes.py --once
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpo_b200 import ops  # noqa: E402

dev = "cuda"
once = "--once" in sys.argv
g = torch.Generator(device="cpu").manual_seed(0)


def conv(b, h, c0, c1, n, stats=True, bn=0, pair=0):
    x0 = torch.randn(b, h, h, c0, generator=g).to(dev).to(torch.bfloat16)
    x1 = torch.randn(b, h, h, c1, generator=g).to(dev).to(torch.bfloat16) if c1 else None
    w = (torch.randn(n, 9 * (c0 + c1), generator=g) / 50).to(dev).to(torch.bfloat16)
    bias = torch.randn(n, generator=g).to(dev)
    m = b * h * h
    res = torch.randn(m, n, generator=g).to(dev)
    out = torch.empty(m, n, device=dev)
    st = torch.empty(ops.gn_stats_shape(m, n), device=dev) if stats else None
    fn = lambda: ops.igemm(a0=x0, a1=x1, wt=w, n=n, c0=c0, c1=c1, conv=(b, h, h), taps=9, bias=bias, residual=res, out_f32=out,
                           gn_stats=st, bn=bn, pair=pair)
    for _ in range(0 if once else 3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 1 if once else 10
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    fl = 2.0 * m * n * 9 * (c0 + c1)
    print(f"conv3x3 B{b} {h}x{h} C{c0}+{c1} -> {n} (M{m} N{n} K{9 * (c0 + c1)}) stats={stats} bn={bn} pair={pair}: {t * 1e6:8.1f} us  "
          f"{fl / t / 1e12:7.1f} TFLOP/s", flush=True)


conv(16, 64, 320, 0, 320)        # M65536 N320 K2880   x7 per application
conv(16, 32, 640, 0, 640)        # M16384 N640 K5760   x6
conv(16, 16, 1280, 0, 1280)      # M4096  N1280 K11520 x7
conv(16, 64, 320, 320, 320)      # M65536 N320 K5760 (skip concat)
if not once:
    conv(16, 64, 320, 0, 320, stats=False)
    conv(16, 16, 1280, 0, 1280, stats=False)
    conv(16, 8, 1280, 0, 1280)       # M1024: 8x8 level
    conv(16, 8, 1280, 1280, 1280)    # K23040
    conv(40, 64, 320, 0, 320)
    conv(40, 8, 1280, 0, 1280)
if "--widths" in sys.argv:   # tile-width A/B: 320 = two 160-column sub-tiles per CTA pair (one activation fetch)
    for args in [(16, 64, 320, 0, 320), (16, 32, 640, 0, 640), (16, 16, 1280, 0, 1280), (16, 64, 320, 320, 320),
                 (16, 32, 640, 640, 640), (16, 16, 1280, 1280, 1280), (16, 8, 1280, 0, 1280), (40, 64, 320, 0, 320),
                 (40, 32, 640, 0, 640), (40, 16, 1280, 0, 1280), (40, 8, 1280, 0, 1280)]:
        for bn in (160, 256, 320):
            if args[4] % bn == 0:
                conv(*args, bn=bn)
if "--lowres" in sys.argv:   # 8x8 level: CTA pairs (256-row tiles) vs single CTAs (128-row tiles)
    for args in [(16, 8, 1280, 0, 1280), (16, 8, 1280, 1280, 1280), (16, 8, 1280, 640, 1280)]:
        for pair in (1, 2):
            for bn in (64, 128, 160, 256):
                conv(*args, bn=bn, pair=pair)
print("done")
