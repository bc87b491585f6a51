"""Developer tool: where the e2e - resident gap of the bench step goes (pinned-host upload, D2H, grouping)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import config, synthetic
from artalk_b200.engine import ARTAvatarInferEngine
cfg = config.FULL
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
eng = ARTAvatarInferEngine(load_gaga=False, clip_length=750, device=str(dev), precision="bf16", state_dict=synthetic.make_state_dict(cfg, 0),
                           config=cfg.to_reference_json(), flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec, make_output_dir=False)
B, S = 256, 480000
ah = synthetic.make_audio(B, S).pin_memory(); sh = synthetic.make_style_motion(B).pin_memory()
ad, sd = ah.to(dev), sh.to(dev)
out_host = torch.empty(B, 750, 106).pin_memory()
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = ev(), ev(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
res = {}
tmp = torch.empty(B, S, device=dev)
res["h2d_497MB_ms"] = timeit(lambda: tmp.copy_(ah, non_blocking=True))
tmp2 = torch.empty(B, 512000, device=dev)
res["h2d_strided_ms"] = timeit(lambda: tmp2[:, :S].copy_(ah, non_blocking=True))
m = eng.inference_batch(ad, sd)
res["d2h_motion_ms"] = timeit(lambda: out_host.copy_(m, non_blocking=True))
res["resident_ms"] = timeit(lambda: eng.inference_batch(ad, sd))
def e2e():
    mm = eng.inference_batch(ah, sh); out_host.copy_(mm, non_blocking=True)
for g in ("1", "0"):
    os.environ["ARTALK_UPLOAD_GROUPS"] = g
    res["e2e_groups%s_ms" % g] = timeit(e2e)
# host-side cost of the call (enqueue time): wall clock until the call returns
torch.cuda.synchronize(); t0 = time.perf_counter(); eng.inference_batch(ah, sh); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
res["enqueue_wall_ms"] = (t1 - t0) * 1e3; res["total_wall_ms"] = (t2 - t0) * 1e3
print(json.dumps(res))
