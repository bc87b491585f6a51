"""Shared helpers for parity tests: golden fixture loading, bit packing, cached seeded weights."""
import functools
import os

import numpy as np
import torch

from artalk_b200 import synthetic
from oracle.cases import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COND_STRIDE = 32


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def unpack_bits(words) -> torch.Tensor:
    w = torch.as_tensor(np.asarray(words).astype(np.int64))
    return ((w[..., None] >> torch.arange(32)) & 1).to(torch.int32)


def pack_bits(bits: torch.Tensor) -> torch.Tensor:
    return (bits.to(torch.int64) << torch.arange(32, dtype=torch.int64)).sum(dim=-1)


@functools.lru_cache(maxsize=2)
def state_dict(cfg_name: str, seed: int = 0):
    from artalk_b200 import config
    return synthetic.make_state_dict(getattr(config, cfg_name), seed)


def margins(logits) -> torch.Tensor:
    l = torch.as_tensor(logits)
    l = l.reshape(*l.shape[:-1], 32, 2)
    return (l[..., 1] - l[..., 0]).abs()
