"""ncu / timing target (not a test): attention forward at the U-Net's self-attention (64x64 tokens, 5 heads) and
cross-attention (Nq 4096, Nk 77) shapes of a sampling batch (16), and both backward kernels at the training batch.
    python tests/prof_attention_shapes.py
    ncu --set full --clock-control none --import-source on -k regex:attention -c 6 -o gpurun_out/r2_attention_full \
        python tests/prof_attention_shapes.py --once
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpo_b200 import ops  # noqa: E402

dev = "cuda"
once = "--once" in sys.argv
g = torch.Generator(device="cpu").manual_seed(0)


def timeit(fn, n=10):
    for _ in range(0 if once else 2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(1 if once else n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (1 if once else n) * 1e-3


def run(B, H, Nq, Nk, bwd):
    c = H * 64
    q = (torch.randn(B * Nq, c, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    kv = (torch.randn(B * Nk, 2 * c, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    out = torch.empty(B * Nq, c, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, Nq, device=dev)
    f = 4.0 * B * H * Nq * Nk * 64
    tf = timeit(lambda: ops.attention_fwd(q, kv, kv[:, c:], out, B, H, Nq, Nk, c, 2 * c, 2 * c, c, lse=lse))
    # algorithmic HBM bytes of the forward: Q in + O out (bf16) + K/V once per (sample, head)
    byt = 2.0 * B * Nq * c * 2 + 2.0 * B * Nk * c * 2
    line = (f"attention fwd B{B} H{H} Nq{Nq} Nk{Nk}: {tf * 1e6:8.1f} us  {f / tf / 1e12:7.1f} TFLOP/s  "
            f"{byt / tf / 1e9:6.0f} GB/s (Q,K,V in + O out)")
    if bwd:
        do = (torch.randn(B * Nq, c, generator=g) * 0.1).to(dev).to(torch.bfloat16)
        dq = torch.empty(B * Nq, c, dtype=torch.bfloat16, device=dev)
        dkv = torch.empty(B * Nk, 2 * c, dtype=torch.bfloat16, device=dev)
        delta = torch.empty(B, H, Nq, device=dev)
        tb = timeit(lambda: ops.attention_bwd(q, kv, kv[:, c:], out, do, lse, delta, dq, dkv, dkv[:, c:], B, H, Nq, Nk, c, 2 * c,
                                              2 * c, c, c, c, 2 * c, 2 * c), 5)
        line += f" | bwd {tb * 1e6:8.1f} us {3.5 * f / tb / 1e12:7.1f} TFLOP/s (14 N^2 d)"
    print(line, flush=True)


run(16, 5, 4096, 4096, "--bwd" in sys.argv)     # self-attention, sampling batch (--bwd: both backward kernels too)
run(16, 5, 4096, 77, False)       # cross-attention, sampling batch
if not once:
    run(16, 10, 1024, 1024, False)
    run(16, 10, 1024, 77, False)
    run(40, 5, 4096, 4096, True)  # training batch
    run(40, 5, 4096, 77, True)
print("done")
