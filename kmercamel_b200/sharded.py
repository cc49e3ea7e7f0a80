"""Thin test / bench driver of the multi-GPU construction in its one-process-per-GPU form (torchrun).

The product logic — slicing, hash-range ownership, the peer-memory exchange, the device-side synchronisation, the fallback
to the exact construction — lives in libkcgpu (kmercamel_b200/csrc/group.cuh, include/kcgpu.h "multi-GPU").  What is left
for the caller is set-up plumbing: all-gather the ranks' 64-byte CUDA IPC heap handles once (torch.distributed here; any
transport works) and then make the same kc_group_compute_device call on every rank.  No collective runs on the data path.
"""
from __future__ import annotations

import numpy as np

from .api import Context, ComputeResult, group_plan


def gather_handles(handle: np.ndarray, world: int, device=None) -> np.ndarray:
    """All-gather of the per-rank IPC handles -> [world * 64] bytes in rank order (gloo: CPU tensors, nccl: on `device`)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(handle, dtype=np.uint8))
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    if world > 1:
        dist.all_gather_into_tensor(out, t)
    else:
        out.copy_(t)
    return out.cpu().numpy()


def attach(ctx: Context, rank: int, world: int, *, k: int, n_bytes_cap: int, device=None):
    """Make `ctx` rank `rank` of a group of `world` processes sized for inputs of up to n_bytes_cap bytes."""
    import torch.distributed as dist
    handle = ctx.group_alloc(world, rank, k, n_bytes_cap)
    ctx.group_open(gather_handles(handle, world, device))
    if world > 1:
        dist.barrier()  # nobody stores into a heap that its owner has not mapped and zeroed yet


def sharded_compute(ctx: Context, seq_dev_ptr: int, n_bytes: int, *, k: int, complements: bool = True, min_frequency: int = 1) -> ComputeResult:
    """One sharded job; every rank calls this with the same arguments (its own copy of the whole framed sequence)."""
    return ctx.group_compute_device(seq_dev_ptr, n_bytes, k=k, complements=complements, min_frequency=min_frequency)


def check_plans(world: int, k: int, n_bytes: int):
    """Host-side invariants of the geometry every rank derives on its own (kc_group_plan): the slices tile the input in whole
    level-0 tiles, the hash ranges partition the level-0 digits, every rank sizes the same heap."""
    plans = [group_plan(world, r, k, n_bytes) for r in range(world)]
    assert plans[0]["pos_begin"] == 0 and plans[-1]["pos_end"] == n_bytes
    assert plans[0]["digit_begin"] == 0 and plans[-1]["digit_end"] == plans[0]["n_digits"]
    for a, b in zip(plans, plans[1:]):
        assert a["pos_end"] == b["pos_begin"] and a["digit_end"] == b["digit_begin"]
    assert len({(p["heap_bytes"], p["cap_sub"], p["n_digits"], p["fixed_slots"]) for p in plans}) == 1
    return plans
