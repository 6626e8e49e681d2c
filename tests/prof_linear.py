"""ncu target (not a test): the memory-/epilogue-bound small-K linear shapes of the U-Net, two launches each.
    ncu --set full --clock-control none --import-source on -k regex:igemm -o gpurun_out/prof_linear python tests/prof_linear.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpo_b200 import ops

dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
M = int(os.environ.get("PROF_M", 65536))


def lin(k, n, res=False, f32=False, bf16=False, geglu=False):
    a = torch.randn(M, k, generator=g).to(dev).to(torch.bfloat16)
    w = torch.randn(n, k, generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn(n, generator=g).to(dev)
    r = torch.randn(M, n, device=dev) if res else None
    no = n // 2 if geglu else n
    out = torch.zeros(M, no, device=dev) if f32 else None
    outb = torch.zeros(M, no, dtype=torch.bfloat16, device=dev) if bf16 else None
    for _ in range(2):
        ops.igemm(a0=a, wt=w, n=n, c0=k, m=M, bias=bias, residual=r, ld_res=n if res else 0, out_f32=out, out_bf16=outb,
                  ld_out=no, geglu=geglu, bn=256 if geglu else 0)
    torch.cuda.synchronize()


lin(320, 320, res=True, f32=True)
lin(320, 320, bf16=True)
lin(320, 2560, bf16=True, geglu=True)
lin(1280, 320, res=True, f32=True, bf16=True)
lin(640, 640, res=True, f32=True)
print("done")
