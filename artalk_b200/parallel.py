"""Multi-GPU scheme of the path: clips are independent (no cross-clip term anywhere in app/models.py:62-121), so a
job of N clips is sharded into contiguous ranges, one process per GPU, weights replicated; the only collective is the
gather of the output motion tensors after generation (never inside the chunk / scale / block loops). The reference has
no distributed code at all; this module is new.

Backend-agnostic (``nccl`` over NVLink on the B200 box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``rank``; the first ``n_items % world`` ranks hold one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank %d / world %d" % (rank, world))
    base, extra = divmod(max(0, n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: str, device: Optional[torch.device] = None) -> Tuple[int, int, int]:
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def gather_motion(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of per-rank motion ``(n_local, T, D)`` into ``(n_total, T, D)`` in clip order on every rank.
    Ragged shards (n_total not divisible by the world size) are padded to the largest shard for the collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local.shape[0] != n_total:
            raise ValueError("single process must hold all %d clips" % n_total)
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_total, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d clips, expected %d" % (rank, local.shape[0], hi - lo))
    n_max = shard_bounds(n_total, world, 0)[1]
    buf = local
    if local.shape[0] < n_max:
        buf = torch.cat([local, local.new_zeros((n_max - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((world * n_max,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if n_total == world * n_max:
        return out
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_total, world, r)
        parts.append(out[r * n_max:r * n_max + (b - a)])
    return torch.cat(parts, dim=0)


def sharded_inference(engine, make_audio, make_style, n_clips: int, clip_length=None) -> torch.Tensor:
    """Run ``engine.inference_batch`` on this rank's shard and gather: ``make_audio(lo, hi)`` / ``make_style(lo, hi)``
    produce the shard's inputs (so a large job is never materialised on one host). Returns ``(n_clips, T, 106)``."""
    rank, world, _ = env_rank()
    lo, hi = shard_bounds(n_clips, world, rank)
    style = make_style(lo, hi) if make_style is not None else None
    local = engine.inference_batch(make_audio(lo, hi), style, clip_length=clip_length)
    return gather_motion(local, n_clips)
