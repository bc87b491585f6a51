"""FLAME blend+LBS kernel timing: frames/s and achieved HBM GB/s (algorithmic bytes = 60 276 B out + 424 B in per frame),
for the tensor-core (split-bf16) and the fp32 CUDA-core kernels, plus their max deviation from each other.
  python tools_flame.py [--frames 16000]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import synthetic
from artalk_b200.flame import FLAMEModel

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=16000); a = ap.parse_args()
dev = "cuda:0"
asset = synthetic.make_flame_asset(0)
N = a.frames
g = torch.Generator().manual_seed(0)
motion = (0.3 * torch.randn(N, 106, generator=g)).to(dev)
shapes = {"shared_shape": torch.zeros(1, 300, device=dev).expand(N, -1),
          "per_frame_shape": (0.5 * torch.randn(N, 300, generator=g)).to(dev)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res, outs = {}, {}
for prec in ("tc", "fp32"):
    fm = FLAMEModel(n_shape=300, n_exp=100, scale=1.0, no_lmks=True, asset=asset, device=dev, precision=prec)
    for name, shape in shapes.items():
        for _ in range(3):
            v = fm(shape_params=shape, expression_params=motion[:, :100], pose_params=motion[:, 100:])
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.fill_(1)
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record(); v = fm(shape_params=shape, expression_params=motion[:, :100], pose_params=motion[:, 100:]); t1.record()
            torch.cuda.synchronize(); ts.append(t0.elapsed_time(t1))
        ms = sorted(ts)[len(ts) // 2]
        byt = N * (5023 * 3 * 4 + (106 + (300 if name != "shared_shape" else 0)) * 4)
        res["%s/%s" % (prec, name)] = {"ms": round(ms, 4), "frames_per_s": round(N / ms * 1e3), "GBps": round(byt / ms / 1e6, 1),
                                       "frac_of_hbm_peak_6554": round(byt / ms / 1e6 / 6554.2, 4)}
        outs[(prec, name)] = v[: min(N, 512)].clone()
for name in shapes:
    res["max_abs_diff_tc_vs_fp32/" + name] = float((outs[("tc", name)] - outs[("fp32", name)]).abs().max())
# per-kernel split of one tensor-core call (library launch trace: CUDA event after every launch)
import ctypes as C
from artalk_b200 import _lib
lib = _lib.lib(); st = _lib.stream_ptr(torch.device(dev))
fm = FLAMEModel(n_shape=300, n_exp=100, scale=1.0, no_lmks=True, asset=asset, device=dev, precision="tc")
fm(shape_params=shapes["shared_shape"], expression_params=motion[:, :100], pose_params=motion[:, 100:])
torch.cuda.synchronize()
_lib.check(lib.artalk_trace_begin(st))
fm(shape_params=shapes["shared_shape"], expression_params=motion[:, :100], pose_params=motion[:, 100:])
buf = C.create_string_buffer(1 << 16)
lib.artalk_trace_end(buf, len(buf), st)
res["tc_shared_shape_kernel_us"] = [l.split(",")[0] + ":" + l.split(",")[-1] for l in buf.value.decode().strip().split("\n")]
print(json.dumps({"frames": N, **res}))
