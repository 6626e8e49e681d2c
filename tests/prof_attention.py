"""ncu target (not a test): self-attention forward + backward at the U-Net's dominant shape (64x64 tokens, 5 heads).
    ncu --set full --clock-control none --import-source on -k regex:attention -o gpurun_out/prof_attn python tests/prof_attention.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpo_b200 import ops

dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
B, H, N = int(os.environ.get("PROF_B", 8)), 5, 4096
c = H * 64
qkv = (torch.randn(B * N, 3 * c, generator=g) * 0.5).to(dev).to(torch.bfloat16)
out = torch.empty(B * N, c, dtype=torch.bfloat16, device=dev)
lse = torch.empty(B, H, N, device=dev)
dout = (torch.randn(B * N, c, generator=g) * 0.1).to(dev).to(torch.bfloat16)
dqkv = torch.empty(B * N, 3 * c, dtype=torch.bfloat16, device=dev)
delta = torch.empty(B, H, N, device=dev)
for _ in range(2):
    ops.attention_fwd(qkv, qkv[:, c:], qkv[:, 2 * c:], out, B, H, N, N, 3 * c, 3 * c, 3 * c, c, lse=lse)
    ops.attention_bwd(qkv, qkv[:, c:], qkv[:, 2 * c:], out, dout, lse, delta, dqkv, dqkv[:, c:], dqkv[:, 2 * c:], B, H, N, N,
                      3 * c, 3 * c, 3 * c, c, c, 3 * c, 3 * c, 3 * c)
torch.cuda.synchronize()
print("done")
if os.environ.get("PROF_TIME", "1") == "1":
    def t(fn, n=10):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    f = 4.0 * B * H * N * N * 64
    tf = t(lambda: ops.attention_fwd(qkv, qkv[:, c:], qkv[:, 2 * c:], out, B, H, N, N, 3 * c, 3 * c, 3 * c, c, lse=lse))
    tb = t(lambda: ops.attention_bwd(qkv, qkv[:, c:], qkv[:, 2 * c:], out, dout, lse, delta, dqkv, dqkv[:, c:], dqkv[:, 2 * c:],
                                     B, H, N, N, 3 * c, 3 * c, 3 * c, c, c, 3 * c, 3 * c, 3 * c))
    print(f"attention fwd {tf * 1e3:.1f} us {f / tf / 1e9:.1f} TF/s | bwd {tb * 1e3:.1f} us {3.5 * f / tb / 1e9:.1f} TF/s (14 N^2 d)")
