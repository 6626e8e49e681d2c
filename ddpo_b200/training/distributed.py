"""Data-parallel plumbing of the PPO update: one process per GPU, `torch.distributed` (NCCL on the GPU box,
gloo in the CPU tests).  The reference does this with `jax.pmap` + `lax.pmean` over the "batch" axis
(reference pipeline/policy_gradient.py:137-142: `grads = lax.pmean(grads, "batch")`, `info = lax.pmean(info)`),
and shards the per-device sample / train batches by the leading device axis (reference
pipeline/finetune.py:214-230 `shard`/`unshard`).

The only data-path collective is ONE sum all-reduce of the flat fp32 gradient per optimizer update; the mean
(1 / (accumulated micro-batches x world size)) is folded into the fused clip+AdamW kernel's `grad_scale`, so no
extra pass over the 3.46 GB gradient buffer is made.  These helpers are device-agnostic on purpose: the same code
runs under gloo with world_size 2 in `tests/test_distributed_cpu.py`.
"""
from __future__ import annotations

from typing import Tuple

import torch


def is_distributed() -> bool:
    return torch.distributed.is_available() and torch.distributed.is_initialized()


def world_size() -> int:
    return torch.distributed.get_world_size() if is_distributed() else 1


def rank() -> int:
    return torch.distributed.get_rank() if is_distributed() else 0


def allreduce_sum_(flat: torch.Tensor) -> int:
    """In-place sum of `flat` over ranks; returns the world size (the divisor the caller folds into its scale)."""
    w = world_size()
    if w > 1:
        torch.distributed.all_reduce(flat)
    return w


def pmean_(t: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks (`lax.pmean`, reference pipeline/policy_gradient.py:142)."""
    w = world_size()
    if w > 1:
        torch.distributed.all_reduce(t)
        t /= w
    return t


def grad_scale(n_acc: int, world: int) -> float:
    """Scale that turns the rank-summed accumulated gradient into the mean over all micro-batches of all ranks
    (reference :33-41 divides by the accumulation count; pmean divides by the device count)."""
    if n_acc <= 0 or world <= 0:
        raise ValueError(f"grad_scale needs positive counts, got n_acc={n_acc}, world={world}")
    return 1.0 / (n_acc * world)


def shard_bounds(n_units: int, rank_: int | None = None, world: int | None = None) -> Tuple[int, int]:
    """[lo, hi) of the units (prompts / samples) this rank owns.  The reference requires the global batch to divide
    by the device count (`shard` reshapes to [n_devices, -1, ...]); same rule here, loudly."""
    w = world_size() if world is None else world
    r = rank() if rank_ is None else rank_
    if n_units % w != 0:
        raise ValueError(f"{n_units} units do not divide over {w} ranks")
    per = n_units // w
    return r * per, (r + 1) * per


def max_over_ranks(x: float, device=None) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    if world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())
