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


class MotionGather:
    """The path's one collective, off the compute stream: ``start`` enqueues the all-gather of a step's motion on a side
    stream (it waits for the producer through an event), so the collective of step i overlaps the compute of step i+1;
    ``wait`` makes the current stream wait for it and returns the gathered tensor. CUDA events around the collective give
    its device time (``gather_ms``). Falls back to a synchronous gather on CPU process groups (gloo tests)."""

    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else None
        self.cuda = self.device is not None and self.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._pending = []
        self.times_ms = []

    def start(self, local: torch.Tensor, n_total: int, group=None):
        if not self.cuda:
            h = {"out": gather_motion(local, n_total, group), "done": None}
            self._pending.append(h)
            return h
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            t0.record(self.stream)
            out = gather_motion(local, n_total, group)
            t1.record(self.stream)
        local.record_stream(self.stream)
        h = {"out": out, "done": t1, "t0": t0}
        self._pending.append(h)
        return h

    def wait(self, h=None):
        """Current stream waits for ``h`` (default: every pending gather); returns the gathered motion of ``h`` / the last one."""
        hs = [h] if h is not None else list(self._pending)
        out = None
        for x in hs:
            if x["done"] is not None:
                torch.cuda.current_stream(self.device).wait_event(x["done"])
                x["out"].record_stream(torch.cuda.current_stream(self.device))
            out = x["out"]
            if x in self._pending:
                self._pending.remove(x)
                if x["done"] is not None:
                    self.times_ms.append((x["t0"], x["done"]))
        return out

    def mean_ms(self) -> float:
        """Mean device time of the completed gathers (call after a synchronize)."""
        if not self.times_ms:
            return 0.0
        return sum(a.elapsed_time(b) for a, b in self.times_ms) / len(self.times_ms)


def sharded_inference(engine, make_audio, make_style, n_clips: int, clip_length=None) -> torch.Tensor:
    """Run ``engine.inference_batch`` on this rank's shard and gather: ``make_audio(lo, hi)`` / ``make_style(lo, hi)``
    produce the shard's inputs (so a large job is never materialised on one host). Returns ``(n_clips, T, 106)``."""
    rank, world, _ = env_rank()
    lo, hi = shard_bounds(n_clips, world, rank)
    style = make_style(lo, hi) if make_style is not None else None
    local = engine.inference_batch(make_audio(lo, hi), style, clip_length=clip_length)
    return gather_motion(local, n_clips)
