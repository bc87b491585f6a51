"""Audio front-end of the path's callers (inference.py:112-113,230-231):

    audio, sr = torchaudio.load(path)
    audio = torchaudio.transforms.Resample(sr, 16000)(audio).mean(dim=0)

restated without torchaudio: a stdlib WAV reader (torchaudio.load needs torchcodec, absent here) and the polyphase
windowed-sinc resampler + channel mean as one CUDA kernel (``artalk_resample_mono``). The filter bank follows
torchaudio/functional/functional.py::_get_sinc_resample_kernel (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99),
built in fp64 and rounded to fp32 exactly like ``transforms.Resample`` does. Host logic only; the arithmetic is in
csrc/frontend.cu and there is no CPU path.
"""
from __future__ import annotations

import math
import wave
from typing import Tuple

import numpy as np
import torch

from . import _lib

TARGET_SR = 16000
_bank_cache = {}


def sinc_resample_bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99
                       ) -> Tuple[np.ndarray, int, int, int]:
    """-> (bank [new][taps] float32, orig, new, width) for the gcd-reduced rates (taps = 2*width + orig)."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("sample rates must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = np.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / orig)
    return np.ascontiguousarray(k.astype(np.float32)), orig, new, width


def read_wav(path: str) -> Tuple[torch.Tensor, int]:
    """PCM WAV -> ((channels, samples) float32 in [-1, 1), sample_rate) like ``torchaudio.load`` (normalize=True)."""
    with wave.open(path, "rb") as w:
        ch, sw, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if sw == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif sw == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) / 8388608.0
    else:
        raise ValueError("unsupported WAV sample width %d" % sw)
    return torch.from_numpy(x.reshape(-1, ch).T.copy()), sr


def resample_mono(waveform: torch.Tensor, sr: int, new_sr: int = TARGET_SR, device="cuda") -> torch.Tensor:
    """(channels, S) or (S,) fp32 at ``sr`` -> (ceil(new_sr * S / sr),) fp32 mono at ``new_sr`` on ``device``
    (== ``torchaudio.transforms.Resample(sr, new_sr)(waveform).mean(dim=0)``; with sr == new_sr only the channel mean)."""
    dev = _lib.require_cuda(device)
    x = waveform if waveform.dim() == 2 else waveform[None]
    if x.dim() != 2 or x.shape[1] < 1:
        raise ValueError("waveform must be (channels, samples) or (samples,)")
    x = x.to(dev, torch.float32).contiguous()
    C_, S = x.shape
    key = (int(sr), int(new_sr), str(dev))
    if key not in _bank_cache:
        bank, orig, new, width = sinc_resample_bank(sr, new_sr) if sr != new_sr else (np.ones((1, 1), np.float32), 1, 1, 0)
        _bank_cache[key] = (torch.from_numpy(bank).to(dev), orig, new, width)
    bank, orig, new, width = _bank_cache[key]
    if sr == new_sr:                       # torchaudio returns the input unchanged; only the mean remains
        return x.mean(dim=0) if C_ > 1 else x[0].clone()
    out_len = -((-new * S) // orig)        # ceil(new * S / orig)
    out = torch.empty(out_len, device=dev, dtype=torch.float32)
    _lib.call(dev, _lib.lib().artalk_resample_mono, x.data_ptr(), C_, x.stride(0), S, bank.data_ptr(), orig, new, bank.shape[1], width,
                                               out.data_ptr(), out_len, _lib.stream_ptr(dev))
    return out


def load_audio(path: str, device="cuda") -> torch.Tensor:
    """inference.py:230-231 for a WAV file: 16 kHz mono fp32 on ``device``."""
    wav, sr = read_wav(path)
    return resample_mono(wav, sr, TARGET_SR, device)
