"""Multi-GPU sharding of the PFNL forward: one process per GPU (torchrun), clips split across
ranks, weights replicated, no collective inside the forward (clips are independent: the
non-local block is per clip, utils.py:44-53).  The only exchange is an all-gather of the
per-clip MSE vector for the PSNR reduction (model/pfnl.py:90,139-141), optionally of the SR
frames.  Works with backend "nccl" (GPU) and "gloo" (CPU tests of the host logic)."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* if launched by torchrun.
    Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend)
    return rank, world, local


def shard_range(n_total, rank, world):
    """Contiguous block of clips for `rank`: the first n_total % world ranks get one extra."""
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_clips(local, n_total, group=None):
    """local: tensor [n_local, ...] holding this rank's shard (shard_range order).
    Returns the full [n_total, ...] tensor in clip order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def sharded_eval_mse(model, lr_local, hr_local, n_total, group=None):
    """Per-clip MSE of this rank's shard -> all-gathered [n_total,1] MSE and PSNR on every rank
    (model/pfnl.py:90,139-141)."""
    sr = model.forward(lr_local)
    mse_local = model.engine.mse(sr, hr_local)
    mse = all_gather_clips(mse_local, n_total, group)[:, None]
    psnr = 10.0 * torch.log10(1.0 / mse.double())
    return mse, psnr


def max_over_ranks(value, device=None, group=None):
    """MAX all-reduce of a python float (timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def psnr_np(mse):
    return 10.0 * np.log10(1.0 / np.asarray(mse, np.float64))
