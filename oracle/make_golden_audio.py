"""Generates tests/golden/resample.npz: outputs of the INSTALLED torchaudio (the reference's third-party resampler) for
seeded inputs, used to pin oracle/audio_oracle.py and the CUDA kernel. Run here (CPU):  python oracle/make_golden_audio.py"""
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
cases = [("48k_stereo", 48000, 2, 9601), ("44k1_mono", 44100, 1, 5000), ("22k05_stereo", 22050, 2, 4410), ("8k_mono", 8000, 1, 777),
         ("16k_stereo", 16000, 2, 1000)]
for name, sr, ch, n in cases:
    g = torch.Generator().manual_seed(sr + ch)
    x = (0.1 * torch.randn(ch, n, generator=g)).float()
    y = torchaudio.transforms.Resample(sr, 16000)(x).mean(dim=0)          # inference.py:231
    out[name + "_in"] = x.numpy()
    out[name + "_out"] = y.numpy()
    out[name + "_sr"] = np.int64(sr)
out["torchaudio_version"] = np.array(torchaudio.__version__)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resample.npz"), **out)
print("wrote tests/golden/resample.npz", {k: v.shape for k, v in out.items() if k.endswith("_out")})
