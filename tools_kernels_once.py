"""Profiling aid: one launch each of the kernels that bench.py's step does not reach (audio front-end resampler, operand
splitter of the parity-grade modes, vertex normals, forehead EMA) at representative sizes, for `ncu -k <name>` captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import _lib, synthetic, audio as fe
from artalk_b200.flame import FLAMEModel
from artalk_b200.gaga import GagaPointBuilder

dev = torch.device("cuda:0")
lib = _lib.lib()
# 30 s of 48 kHz stereo -> 16 kHz mono (inference.py:230-231)
wav = 0.1 * torch.randn(2, 48000 * 30, device=dev)
for _ in range(2):
    mono = fe.resample_mono(wav, 48000, 16000, device=dev)
# operand split of a wav2vec FFN2 A operand of a 96-chunk sub-batch (19104 x 4096 fp32 -> 3 / 6 slots)
x = torch.randn(19104 * 4096, device=dev)
for slots in (3, 6):
    out = torch.empty(x.numel() * slots, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        _lib.check(lib.artalk_op_split_bf16(x.data_ptr(), out.data_ptr(), x.numel(), slots, 0, _lib.stream_ptr(dev)))
# vertex normals + forehead EMA of 16 000 decoded frames
asset = synthetic.make_flame_asset(0)
fm = FLAMEModel(n_shape=300, n_exp=100, scale=5.0, no_lmks=True, asset=asset, device=str(dev))
motion = (0.3 * torch.randn(16000, 106)).to(dev)
b = GagaPointBuilder(fm, 0.5 * torch.randn(1, 300))
for _ in range(2):
    b.reset()
    pts = b.t_points(motion)
    n = fm.vertex_normals(pts)
torch.cuda.synchronize()
print("ok", tuple(mono.shape), tuple(n.shape))
