"""TEST INFRASTRUCTURE ONLY (imported by tests/ and never by the product path).

CPU restatement of the audio front-end the reference's callers run before the hot path (inference.py:112-113,230-231):
``torchaudio.transforms.Resample(sr, 16000)(audio).mean(dim=0)``. torchaudio is a third-party dependency of the reference
(pinned 2.4.1 in environment.yml:218; 2.11.0 installed here); its published algorithm
(torchaudio/functional/functional.py: _get_sinc_resample_kernel + _apply_sinc_resample_kernel) is restated below in numpy
fp64 and pinned against the installed torchaudio by oracle/make_golden_audio.py -> tests/golden/resample.npz.
"""
import math

import numpy as np


def sinc_kernel(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """functional.py::_get_sinc_resample_kernel, sinc_interp_hann; returns (kernel [new][taps] fp64, orig, new, width)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    out = np.zeros((new, 2 * width + orig))
    for p in range(new):
        for k in range(2 * width + orig):
            t = (-p / new + (k - width) / orig) * base
            t = min(max(t, -lowpass_filter_width), lowpass_filter_width)
            win = math.cos(t * math.pi / lowpass_filter_width / 2) ** 2
            t *= math.pi
            out[p, k] = (1.0 if t == 0 else math.sin(t) / t) * win * (base / orig)
    return out, orig, new, width


def resample_mean(wave, sr, new_sr=16000):
    """(channels, S) -> (ceil(new*S/orig),): pad (width, width+orig), stride-orig correlation with every phase, interleave
    the phases, cut to the target length (functional.py::_apply_sinc_resample_kernel), then mean over channels."""
    wave = np.asarray(wave, dtype=np.float64)
    if wave.ndim == 1:
        wave = wave[None]
    if sr == new_sr:
        return wave.mean(axis=0)
    k, orig, new, width = sinc_kernel(sr, new_sr)
    k = k.astype(np.float32).astype(np.float64)            # transforms.Resample stores the bank in fp32
    C, S = wave.shape
    pad = np.concatenate([np.zeros((C, width)), wave, np.zeros((C, width + orig))], axis=1)
    n_frames = (pad.shape[1] - k.shape[1]) // orig + 1
    idx = np.arange(n_frames)[:, None] * orig + np.arange(k.shape[1])[None, :]
    frames = pad[:, idx]                                    # (C, n_frames, taps)
    res = np.einsum("cft,pt->cfp", frames, k).reshape(C, -1)
    target = int(math.ceil(new * S / orig))
    return res[:, :target].mean(axis=0)
