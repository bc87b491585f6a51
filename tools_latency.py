"""Latency mode (BASELINE configs[4]): streaming batch-1 inference, one 4 s / 100-frame chunk at a time
(wav2vec2 on the chunk -> AR chunk -> savgol on the running clip -> FLAME LBS vertices for the chunk's frames).
Prints p50/p90 per chunk (the reference's chunk is 4 s; per-2-s figure = /2).
  python tools_latency.py [--chunks 40 --precision bf16 --clips 1]"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import config, synthetic
from artalk_b200.engine import ARTAvatarInferEngine

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=40); ap.add_argument("--precision", default="bf16")
ap.add_argument("--clips", type=int, default=1); ap.add_argument("--config", default="FULL")
ap.add_argument("--no-mesh", action="store_true")
ap.add_argument("--latency-rows", type=int, default=128, help="row cap of the latency kernels (artalk_set_latency_mode)")
ap.add_argument("--throughput-mode", action="store_true", help="keep the batch-invariant throughput kernels (no latency mode)")
a = ap.parse_args()
cfg = getattr(config, a.config)
dev = torch.device("cuda:0")
eng = ARTAvatarInferEngine(load_gaga=False, device=str(dev), precision=a.precision, state_dict=synthetic.make_state_dict(cfg, 0),
                           config=cfg.to_reference_json(), flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec,
                           make_output_dir=False, latency_mode=not a.throughput_mode)
m = eng.ARTalk
if not a.throughput_mode:
    m.set_latency_mode(True, a.latency_rows)
B = a.clips
audio = synthetic.make_audio(B, cfg.chunk_samples * a.chunks).pin_memory()
style = m.style_cond(synthetic.make_style_motion(B), B)
prev = m.initial_words(B)
out = torch.empty(B, 100, 106, device=dev)
shape = torch.zeros(1, 300, device=dev).expand(B * 100, -1)
lat, parts = [], []
for c in range(a.chunks):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chunk = audio[:, c * cfg.chunk_samples:(c + 1) * cfg.chunk_samples].to(dev, non_blocking=True)
    cond = m.audio_cond(chunk)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    m.ar_chunk(cond, style, prev, out)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if not a.no_mesh:
        verts = m.basic_vae.get_flame_verts(eng.flame_model, shape, out.view(B * 100, 106), with_global=True)
    host = out.cpu()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if c >= 5:
        lat.append((t3 - t0) * 1e3); parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
lat.sort()
p = lambda q: lat[min(len(lat) - 1, int(q * len(lat)))]
med = lambda i: sorted(x[i] for x in parts)[len(parts) // 2]
print(json.dumps({"mode": "latency", "latency_kernels": not a.throughput_mode, "clips": B, "precision": a.precision, "chunks_timed": len(lat),
                  "p50_ms_per_4s_chunk": p(0.5), "p90_ms_per_4s_chunk": p(0.9), "p50_ms_per_2s": p(0.5) / 2,
                  "p50_parts_ms": {"h2d+wav2vec": med(0), "ar+vae": med(1), "flame+d2h": med(2)},
                  "frames_per_sec": B * 100 / (p(0.5) / 1e3), "mesh": not a.no_mesh}))
