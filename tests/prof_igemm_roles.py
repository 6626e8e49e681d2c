"""Timing target (not a test): what each role of the CTA-pair implicit GEMM costs on its own.
DDPO_IGEMM_DEBUG bit 0 = no operand loads (the producer only arrives), bit 1 = no MMAs, bit 2 = no epilogue body; the
barrier protocol and the tile schedule are unchanged, results are garbage.
    python tests/prof_igemm_roles.py
"""
import os
import sys

import torch

sys.path.insert(0, os.environ.get("DDPO_PROF_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ddpo_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)


def run(b, h, c, n, taps=9, res=True, stats=True, f32=True, bn=0, tag="", epi=2):
    m = b * h * h
    if taps == 9:
        x = torch.randn(b, h, h, c, generator=g).to(dev).to(torch.bfloat16)
    else:
        x = torch.randn(m, c, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(n, taps * c, generator=g) / 50).to(dev).to(torch.bfloat16)
    bias = torch.randn(n, generator=g).to(dev)
    r = torch.randn(m, n, generator=g).to(dev) if res else None
    out = torch.empty(m, n, device=dev) if f32 else None
    ob = None if f32 else torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    st = torch.empty(ops.gn_stats_shape(m, n), device=dev) if (stats and f32) else None
    kw = dict(conv=(b, h, h), taps=9) if taps == 9 else dict(m=m)
    fn = lambda: ops.igemm(a0=x, wt=w, n=n, c0=c, bias=bias, residual=r, out_f32=out, out_bf16=ob, gn_stats=st, bn=bn, epi=epi,
                           pair=1, **kw)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e-3
    fl = 2.0 * m * n * taps * c
    print(f"{tag:34s} M{m} N{n} K{taps * c} taps{taps} res={int(res)} stats={int(stats)} f32={int(f32)} bn={bn} epi={epi}: {t * 1e6:8.1f} us "
          f"{fl / t / 1e12:7.1f} TFLOP/s", flush=True)


NAMES = {0: "all roles", 1: "no loads", 2: "no MMA", 3: "epilogue only", 4: "no epilogue", 5: "MMA only", 6: "loads only", 7: "schedule only"}
names = NAMES
if "--shortk" in sys.argv:   # short-K layers: TMA-staged epilogue (epi=1) vs coalescing register epilogue (epi=2)
    for (m, c, n, res, f32) in [(65536, 320, 320, True, True), (65536, 1280, 320, True, True), (16384, 640, 640, True, True),
                                (4096, 1280, 1280, True, True), (65536, 320, 960, False, False), (16384, 640, 1920, False, False),
                                (163840, 320, 320, True, True), (163840, 320, 320, False, False), (40960, 640, 640, True, True)]:
        for epi in (1, 2):
            for bn in ((0, 320) if n % 320 == 0 else (0,)):
                if epi == 1 and bn == 320:
                    continue
                run(m, 1, c, n, taps=1, res=res, stats=False, f32=f32, bn=bn, epi=epi, tag="linear")
    print("done")
    sys.exit(0)
quick = "--ab" in sys.argv
if "--epi" in sys.argv:   # register epilogue (all stages for operands) vs TMA-staged epilogue (128 KB of staging: 2-3 stages left)
    for shape in [(16, 64, 320, 320), (16, 32, 640, 640), (16, 16, 1280, 1280), (40, 64, 320, 320)]:
        for bn in (160, 320):
            for epi in (2, 1):
                for dbg in (0, 4):
                    os.environ["DDPO_IGEMM_DEBUG"] = str(dbg)
                    run(*shape, bn=bn, epi=epi, tag=names[dbg])
    os.environ["DDPO_IGEMM_DEBUG"] = "0"
    print("done")
    sys.exit(0)
for shape in [(16, 64, 320, 320), (16, 16, 1280, 1280)]:
    for bn in ((160,) if quick else (160, 320)):
        for dbg in ((0,) if quick else (0, 1, 2, 4, 5, 6)):   # 3 / 7 (neither loads nor MMAs) fault: not investigated
            os.environ["DDPO_IGEMM_DEBUG"] = str(dbg)
            run(*shape, bn=bn, tag=names[dbg])
os.environ["DDPO_IGEMM_DEBUG"] = "0"
if not quick:
    for bn in (160, 320):
        run(16, 64, 320, 320, res=False, bn=bn, tag="no residual")
        run(16, 64, 320, 320, stats=False, bn=bn, tag="no statistics")
        run(16, 64, 320, 320, res=False, stats=False, bn=bn, tag="no residual, no statistics")
        run(16, 64, 320, 320, res=False, f32=False, bn=bn, tag="bf16 output only")
        run(16, 64, 2880, 320, taps=1, bn=bn, tag="linear with the same K")
        os.environ["DDPO_IGEMM_PREFETCH"] = "0"
        run(16, 64, 320, 320, bn=bn, tag="no residual prefetch")
        os.environ["DDPO_IGEMM_PREFETCH"] = "1"
print("done")
