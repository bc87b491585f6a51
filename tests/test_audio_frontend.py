"""Audio front-end (inference.py:112-113,230-231): the numpy oracle and the host-built filter bank
against outputs of the installed torchaudio (tests/golden/resample.npz, oracle/make_golden_audio.py), and the CUDA
resample + channel-mean kernel against both. Tolerance: fp32 FIR sums of <= 475 taps, 1e-5 absolute on |x| ~ 0.1."""
import os
import struct
import wave

import numpy as np
import pytest
import torch

from artalk_b200 import audio
from oracle import audio_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample.npz")
CASES = ["48k_stereo", "44k1_mono", "22k05_stereo", "8k_mono", "16k_stereo"]
TOL = 1e-5


def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_torchaudio_golden(name):
    g = gold()
    y = audio_oracle.resample_mean(g[name + "_in"], int(g[name + "_sr"]))
    ref = g[name + "_out"]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < TOL


@pytest.mark.parametrize("sr", [48000, 44100, 22050, 8000, 32000])
def test_filter_bank_matches_oracle_kernel(sr):
    bank, orig, new, width = audio.sinc_resample_bank(sr, 16000)
    k, o2, n2, w2 = audio_oracle.sinc_kernel(sr, 16000)
    assert (orig, new, width) == (o2, n2, w2) and bank.shape == k.shape == (new, 2 * width + orig)
    assert np.abs(bank.astype(np.float64) - k).max() < 1e-7


def test_read_wav_pcm16_stereo(tmp_path):
    p = str(tmp_path / "t.wav")
    data = [(-32768, 32767), (0, 1), (12345, -12345)]
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes(b"".join(struct.pack("<hh", a, b) for a, b in data))
    x, sr = audio.read_wav(p)
    assert sr == 48000 and x.shape == (2, 3) and x.dtype == torch.float32
    assert torch.allclose(x[0], torch.tensor([-1.0, 0.0, 12345 / 32768.0]))
    assert torch.allclose(x[1], torch.tensor([32767 / 32768.0, 1 / 32768.0, -12345 / 32768.0]))


def test_resample_requires_cuda_device():
    from artalk_b200._lib import ArtalkError
    with pytest.raises(ArtalkError):
        audio.resample_mono(torch.zeros(2, 100), 48000, 16000, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_resample_matches_golden_and_oracle(name):
    g = gold()
    x, sr = torch.from_numpy(g[name + "_in"]), int(g[name + "_sr"])
    y = audio.resample_mono(x, sr, 16000, device="cuda:0").cpu().numpy()
    assert y.shape == g[name + "_out"].shape
    assert np.abs(y - g[name + "_out"]).max() < TOL
    assert np.abs(y - audio_oracle.resample_mean(g[name + "_in"], sr)).max() < TOL


@pytest.mark.gpu
def test_cuda_resample_long_clip_properties():
    """30 s of 48 kHz stereo (BASELINE clip length): length rule, linearity and the DC gain of the filter bank."""
    g = torch.Generator().manual_seed(7)
    S = 48000 * 30 + 17
    a, b = 0.1 * torch.randn(2, S, generator=g), 0.1 * torch.randn(2, S, generator=g)
    ya, yb = audio.resample_mono(a, 48000, device="cuda:0"), audio.resample_mono(b, 48000, device="cuda:0")
    yab = audio.resample_mono(2.0 * a - 3.0 * b, 48000, device="cuda:0")
    assert ya.shape[0] == -((-S) // 3)
    assert (yab - (2.0 * ya - 3.0 * yb)).abs().max().item() < 2e-5
    dc = audio.resample_mono(torch.ones(1, 48000), 48000, device="cuda:0")
    assert (dc[100:-100] - 1.0).abs().max().item() < 2e-3       # windowed sinc with rolloff 0.99: unit DC gain to ~1e-3
