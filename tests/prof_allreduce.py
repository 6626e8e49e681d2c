"""Timing target (not a test): the one data-path collective of the PPO update -- sum all-reduce of the flat fp32 U-Net
gradient (865.9 M floats = 3.46 GB) -- under different NCCL settings and bucketings.  Launch with torchrun:
    NCCL_ALGO=... python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/prof_allreduce.py [tag]
Prints (rank 0) one line per variant: ms (max over ranks, CUDA events) and bus bandwidth 2 (N-1)/N bytes / t."""
import os
import sys

import torch
import torch.distributed as dist

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
n = 865_910_724 // 64 * 64
g = torch.ones(n, device="cuda")
side = torch.cuda.Stream()


def run(kind, buckets):
    if buckets == 1:
        dist.all_reduce(g)
        return
    per = (n // buckets + 63) // 64 * 64
    works = []
    for i in range(buckets):
        works.append(dist.all_reduce(g[i * per:min(n, (i + 1) * per)], async_op=(kind == "async")))
    if kind == "async":
        for w in works:
            w.wait()


for kind, buckets in (("sync", 1), ("sync", 4), ("sync", 8), ("async", 8), ("sync", 32)):
    for _ in range(2):
        run(kind, buckets)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(kind, buckets)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    g.fill_(1.0)
    if rank == 0:
        ms = t.item()
        print(f"[allreduce {tag}] N={world} {kind} buckets={buckets}: {ms:7.2f} ms  bus {2 * (world - 1) / world * n * 4 / ms / 1e6:7.1f} GB/s",
              flush=True)
# bf16 wire format (halves the bytes; NOT the default: the reference reduces in fp32)
gb = torch.ones(n, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    dist.all_reduce(gb)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dist.all_reduce(gb)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"[allreduce {tag}] N={world} bf16 single: {t.item():7.2f} ms", flush=True)
dist.destroy_process_group()
