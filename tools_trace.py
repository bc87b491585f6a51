"""Developer tool: per-launcher / per-shape time table of one bench step from the library's launch trace
(CUDA events, no ncu replay).  python tools_trace.py [--clips 64 --seconds 10 --precision bf16 --top 40]"""
import argparse, collections, ctypes as C, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import config, synthetic, _lib
from artalk_b200.engine import ARTAvatarInferEngine

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64); ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--precision", default="bf16"); ap.add_argument("--top", type=int, default=45)
ap.add_argument("--config", default="FULL"); ap.add_argument("--raw", default="")
a = ap.parse_args()
cfg = getattr(config, a.config)
dev = "cuda:0"
eng = ARTAvatarInferEngine(load_gaga=False, device=dev, precision=a.precision, state_dict=synthetic.make_state_dict(cfg, 0),
                           config=cfg.to_reference_json(), flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec,
                           make_output_dir=False)
audio = synthetic.make_audio(a.clips, int(a.seconds * 16000)).to(dev)
style = synthetic.make_style_motion(a.clips).to(dev)
for _ in range(2):
    eng.inference_batch(audio, style)
torch.cuda.synchronize()
lib = _lib.lib()
st = _lib.stream_ptr(torch.device(dev))
_lib.check(lib.artalk_trace_begin(st))
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); eng.inference_batch(audio, style); t1.record()
buf = C.create_string_buffer(4 << 20)
n = lib.artalk_trace_end(buf, len(buf), st)
torch.cuda.synchronize()
text = buf.value.decode()
if a.raw:
    open(a.raw, "w").write(text)
agg = collections.defaultdict(lambda: [0, 0.0])
for line in text.strip().split("\n"):
    f, d0, d1, d2, us = line.split(",")
    agg[(f, int(d0), int(d1), int(d2))][0] += 1
    agg[(f, int(d0), int(d1), int(d2))][1] += float(us)
tot = sum(v[1] for v in agg.values())
print("step %.2f ms by events, %.2f ms traced, %d launches" % (t0.elapsed_time(t1), tot / 1e3, sum(v[0] for v in agg.values())))
for (f, d0, d1, d2), (c, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:a.top]:
    fl = ""
    if "gemm" in f and d0:
        fl = "%7.0f TFLOP/s" % (2.0 * d0 * d1 * d2 * c / us / 1e6)
    elif "attention" in f and d0:
        fl = "%7.0f TFLOP/s" % (4.0 * d0 * d1 * d2 * 64 * c / us / 1e6)
    print("%-28s %8d %6d %6d  n=%4d %9.1f us %5.1f%%  avg %8.1f %s" % (f, d0, d1, d2, c, us, 100 * us / tot, us / c, fl))
