"""ORACLE tooling — seeded parity cases shared by ``make_golden.py`` and ``tests/``.

Each case names a config, a weight seed and the synthetic inputs; everything is
regenerated from seeds on whatever box runs the tests (weights are ~2 GB fp32 for the
full config and are never stored), only the reference's *outputs* live in tests/golden/.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from artalk_b200 import config as _cfg
from artalk_b200 import synthetic


@dataclass(frozen=True)
class Case:
    name: str
    cfg_name: str               # "TINY" | "FULL"
    n_clips: int
    n_samples: int              # per clip, 16 kHz
    with_style: bool
    weight_seed: int = 0
    audio_fixture: Optional[str] = None     # tests/golden/<name>.npz holding 'audio' (n_clips, n_samples) fp32 instead of seeded noise

    @property
    def cfg(self):
        return getattr(_cfg, self.cfg_name)

    def audio(self) -> torch.Tensor:
        if self.audio_fixture:
            z = np.load(os.path.join(GOLD, self.audio_fixture + ".npz"))
            a = torch.from_numpy(z["audio"].astype(np.float32))
            assert tuple(a.shape) == (self.n_clips, self.n_samples), a.shape
            return a
        return synthetic.make_audio(self.n_clips, self.n_samples)

    def style(self) -> Optional[torch.Tensor]:
        return synthetic.make_style_motion(self.n_clips) if self.with_style else None


GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # 5.2 s -> 130 frames -> 2 chunks (second one zero padded: quirk 4), with style
    "tiny_style": Case("tiny_style", "TINY", 2, 83200, True),
    # exactly one chunk, null style (app/models.py:71-73)
    "tiny_null": Case("tiny_null", "TINY", 1, 64000, False),
    # ragged tail: 4.01 s -> 101 frames -> 2 chunks, 1 useful frame in the second
    "tiny_ragged": Case("tiny_ragged", "TINY", 1, 64160, True),
    # full depth, 10 s clip (BASELINE configs[1] per-clip unit): 250 frames, 3 chunks
    "full_10s": Case("full_10s", "FULL", 1, 160000, True),
    # full depth, 30 s clip = clip_length 750 (BASELINE configs[2]/[3] per-clip unit): 8 chunks of KV-cached recurrence
    "full_30s": Case("full_30s", "FULL", 1, 480000, True),
    # BASELINE configs[0]: the reference's own demo clip demo/eng1.wav after inference.py:230-231 (48 kHz stereo ->
    # Resample(48000, 16000) -> channel mean): 217 088 samples -> 340 frames -> 4 chunks, last one zero padded. The
    # resampled mono clip is stored as a fixture (tests/golden/eng1_audio.npz, written by make_golden.py) because
    # /root/reference does not exist on the GPU box
    "full_eng1": Case("full_eng1", "FULL", 1, 217088, True, audio_fixture="eng1_audio"),
}

#: cases additionally run through ARTAvatarInferEngine.inference (savgol + clip + zeroing): name -> clip_length
ENGINE_CASES = {"tiny_style": 120, "full_eng1": 750}

FLAME_CASE = dict(seed=7, n_frames=6)


def flame_inputs():
    g = torch.Generator().manual_seed(FLAME_CASE["seed"])
    n = FLAME_CASE["n_frames"]
    shape = 0.5 * torch.randn(n, 300, generator=g)
    motion = 0.3 * torch.randn(n, 106, generator=g)
    motion[0, 100:] = 0.0           # exact zero rotation exercises the ||r+1e-8|| quirk
    return shape, motion


GAGA_CASE = dict(seed=11, n_frames=12)


def gaga_inputs():
    """Motion codes of consecutive frames + the tracked avatar's shape code for the GAGAvatar point builder (f4)."""
    g = torch.Generator().manual_seed(GAGA_CASE["seed"])
    motion = 0.3 * torch.randn(GAGA_CASE["n_frames"], 106, generator=g)
    shape = 0.5 * torch.randn(1, 300, generator=g)
    return motion, shape
