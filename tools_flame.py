"""FLAME blend+LBS kernel timing: frames/s and achieved HBM GB/s (algorithmic bytes = 60 276 B out + 424 B in per frame).
  python tools_flame.py [--frames 16000]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import synthetic
from artalk_b200.flame import FLAMEModel

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=16000); a = ap.parse_args()
dev = "cuda:0"
fm = FLAMEModel(n_shape=300, n_exp=100, scale=1.0, no_lmks=True, asset=synthetic.make_flame_asset(0), device=dev)
N = a.frames
g = torch.Generator().manual_seed(0)
motion = (0.3 * torch.randn(N, 106, generator=g)).to(dev)
res = {}
for name, shape in (("shared_shape", torch.zeros(1, 300, device=dev).expand(N, -1)),
                    ("per_frame_shape", (0.5 * torch.randn(N, 300, generator=g)).to(dev))):
    for _ in range(3):
        v = fm(shape_params=shape, expression_params=motion[:, :100], pose_params=motion[:, 100:])
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(5):
        flush.fill_(1)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); v = fm(shape_params=shape, expression_params=motion[:, :100], pose_params=motion[:, 100:]); t1.record()
        torch.cuda.synchronize(); ts.append(t0.elapsed_time(t1))
    ms = sorted(ts)[len(ts) // 2]
    byt = N * (5023 * 3 * 4 + (106 + (300 if name != "shared_shape" else 0)) * 4)
    flops = N * 5023 * 3 * 2 * (136 if name == "shared_shape" else 436)
    res[name] = {"ms": ms, "frames_per_s": N / ms * 1e3, "GBps": byt / ms / 1e6, "frac_of_6554": byt / ms / 1e6 / 6554.2,
                 "fp32_TFLOPs": flops / ms / 1e9}
print(json.dumps({"frames": N, **res}))
