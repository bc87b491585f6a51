"""Model configuration for the audio->motion path.

Mirrors the reference's only config file (``assets/config.json:1-15``) plus the
constants the reference hard-codes in Python (``app/models.py:19,22,27,37,41-42``:
embed 768, cond 1024, style 128) and the XLS-R-300m ``Wav2Vec2Config`` that the
reference fetches over the network (``app/models.py:25``).
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field, asdict
from typing import Tuple


@dataclass(frozen=True)
class Wav2VecConfig:
    """Subset of HF ``Wav2Vec2Config`` that the path reads (XLS-R-300m values)."""
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    ffn: int = 4096
    conv_dim: int = 512
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    pos_conv_kernel: int = 128
    pos_conv_groups: int = 16
    ln_eps: float = 1e-5

    def conv_lengths(self, n_samples: int):
        """Output length after each feature-extractor conv (no padding)."""
        out, L = [], n_samples
        for k, s in zip(self.conv_kernel, self.conv_stride):
            L = (L - k) // s + 1
            out.append(L)
        return out


@dataclass(frozen=True)
class ModelConfig:
    # AR_CONFIG
    ar_depth: int = 12
    ar_heads: int = 12
    prev_ratio: int = 1
    # VAE_CONFIG
    motion_dim: int = 106
    code_dim: int = 32
    vae_depth: int = 8
    vae_heads: int = 8
    vae_hidden: int = 512
    patch_nums: Tuple[int, ...] = (1, 5, 25, 50, 100)
    # hard-coded in the reference
    embed_dim: int = 768
    cond_dim: int = 1024
    style_dim: int = 128
    style_layers: int = 4
    style_heads: int = 4
    style_ffn: int = 512
    style_len: int = 50
    style_pe_len: int = 600
    sample_rate: int = 16000
    fps: float = 25.0
    wav2vec: Wav2VecConfig = field(default_factory=Wav2VecConfig)

    # ---- derived -------------------------------------------------------
    @property
    def seq_tokens(self) -> int:          # 181
        return sum(self.patch_nums)

    @property
    def chunk_frames(self) -> int:        # 100
        return self.patch_nums[-1]

    @property
    def chunk_samples(self) -> int:       # 64000
        return int(self.chunk_frames / self.fps * self.sample_rate)

    @property
    def audio_frames(self) -> int:        # 199 wav2vec frames per chunk
        return self.wav2vec.conv_lengths(self.chunk_samples)[-1]

    def frames_for_samples(self, n_samples: int) -> int:
        """``seq_length`` of app/models.py:66."""
        return math.ceil(n_samples / self.sample_rate * self.fps)

    def chunks_for_samples(self, n_samples: int) -> int:
        return max(1, math.ceil(self.frames_for_samples(n_samples) / self.chunk_frames))

    # ---- (de)serialisation in the reference's config.json schema -------
    @classmethod
    def from_reference_json(cls, cfg: dict, wav2vec: Wav2VecConfig | None = None) -> "ModelConfig":
        ar, vae = cfg["AR_CONFIG"], cfg["VAE_CONFIG"]
        enc = ar.get("AUDIO_ENCODER", "wav2vec")
        if enc != "wav2vec":
            raise ValueError("Invalid audio encoder: {}".format(enc))
        if vae.get("MOTION_DIM", 106) != 106:
            raise ValueError("MOTION_DIM must be 106 (FLAME exp 100 + pose 6)")
        return cls(
            ar_depth=ar["T_DEPTH"], ar_heads=ar["T_NUM_HEADS"], prev_ratio=ar["PREV_RATIO"],
            motion_dim=106, code_dim=vae["V_CODE_DIM"], vae_depth=vae["T_DEPTH"],
            vae_heads=vae["T_NUM_HEADS"], vae_hidden=vae["T_HIDDEN_DIM"],
            patch_nums=tuple(vae["V_PATCH_NUMS"]),
            wav2vec=wav2vec or Wav2VecConfig(),
        )

    @classmethod
    def from_json_file(cls, path: str, **kw) -> "ModelConfig":
        with open(path) as f:
            return cls.from_reference_json(json.load(f), **kw)

    def to_reference_json(self) -> dict:
        return {
            "AR_CONFIG": {"T_DEPTH": self.ar_depth, "T_NUM_HEADS": self.ar_heads,
                          "PREV_RATIO": self.prev_ratio},
            "VAE_CONFIG": {"MOTION_DIM": self.motion_dim, "V_CODE_DIM": self.code_dim,
                           "T_DEPTH": self.vae_depth, "T_NUM_HEADS": self.vae_heads,
                           "T_HIDDEN_DIM": self.vae_hidden, "V_PATCH_NUMS": list(self.patch_nums)},
        }

    def validate(self) -> None:
        if self.prev_ratio != 1:
            raise ValueError("only PREV_RATIO=1 (assets/config.json:5) is supported")
        if self.embed_dim % self.ar_heads or self.embed_dim // self.ar_heads != 64:
            raise ValueError("AR head_dim must be 64")
        if self.vae_hidden // self.vae_heads != 64 or self.vae_hidden % self.vae_heads:
            raise ValueError("VAE head_dim must be 64")
        if self.code_dim != 32:
            raise ValueError("V_CODE_DIM must be 32 (bits are packed one word per token)")
        if len(self.patch_nums) > 8 or list(self.patch_nums) != sorted(self.patch_nums):
            raise ValueError("V_PATCH_NUMS must be increasing, at most 8 levels")


FULL = ModelConfig()
#: reduced-depth config used by the fast parity tests (same widths, fewer layers)
TINY = ModelConfig(ar_depth=2, vae_depth=2, wav2vec=Wav2VecConfig(layers=2))


def as_dict(cfg: ModelConfig) -> dict:
    return asdict(cfg)
