"""Developer tool: same-process, interleaved A/B of process-wide options on the bench step (no mesh, device-resident inputs).
  python tools_ab.py --clips 64 --seconds 10 --rounds 3 base attn_bound=0 attn_blk=0
Every option spec is `name=value[,name=value...]`; the first spec is the baseline. Options are reset to the library defaults
(DEFAULTS below) before each spec is applied. Chunk graphs are re-captured after an option change (option epoch)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import _lib, config, synthetic
from artalk_b200.engine import ARTAvatarInferEngine

DEFAULTS = {"attn_bound": 1, "w2v_graph_chunks": 4, "skinny_tokens": 1, "pdl_mask": 3, "gemm_pair": 1, "gemm_tma_out": 2, "gemm_band_mb": 32, "posconv4": 1, "conv0_fold": 1, "attn_blk": 1, "attn_split": 1, "gemm_pair_split": 1, "gemm_pair_min_waves10": 18, "gemm_pair_qkv": 1, "gemm_epi_warps": 12}

ap = argparse.ArgumentParser()
ap.add_argument("specs", nargs="+")
ap.add_argument("--clips", type=int, default=64); ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--steps", type=int, default=4); ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--precision", default="bf16"); ap.add_argument("--config", default="FULL")
a = ap.parse_args()
cfg = getattr(config, a.config)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
eng = ARTAvatarInferEngine(load_gaga=False, device=str(dev), precision=a.precision, state_dict=synthetic.make_state_dict(cfg, 0),
                           config=cfg.to_reference_json(), flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec,
                           make_output_dir=False)
lib = _lib.lib()
n = int(a.seconds * cfg.sample_rate)
audio = synthetic.make_audio(a.clips, n).to(dev)
style = synthetic.make_style_motion(a.clips).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def apply(spec):
    for k, v in DEFAULTS.items():
        _lib.check(lib.artalk_set_option(k.encode(), v))
    for kv in spec.split(","):
        if kv and kv != "base":
            k, v = kv.split("=")
            _lib.check(lib.artalk_set_option(k.encode(), int(v)))


def run(spec):
    apply(spec)
    for _ in range(3):                       # eager warm-up, capture, replay
        out = eng.inference_batch(audio, style)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = eng.inference_batch(audio, style); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], out


res = {s: [] for s in a.specs}
ref_out = None
diff = {}
for r in range(a.rounds):
    for s in a.specs:
        ms, out = run(s)
        res[s].append(ms)
        if ref_out is None:
            ref_out = out.clone()
        elif r == 0:
            diff[s] = float((out - ref_out).abs().max())
frames = a.clips * min(750, cfg.frames_for_samples(n))
base = sorted(res[a.specs[0]])[len(res[a.specs[0]]) // 2]
print(json.dumps({"workload": "%d clips x %g s, %s" % (a.clips, a.seconds, a.precision),
                  "results": {s: {"ms": [round(x, 3) for x in v], "median_ms": sorted(v)[len(v) // 2],
                                  "frames_per_s": frames / (sorted(v)[len(v) // 2] / 1e3),
                                  "vs_first": sorted(v)[len(v) // 2] / base, "max_abs_diff_vs_first": diff.get(s)} for s, v in res.items()}}))
