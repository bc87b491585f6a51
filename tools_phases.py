"""Developer tool: CUDA-event time of each phase of one bench step (wav2vec for all chunks, style encoder, every AR chunk,
smoothing) with the production launch mode (CUDA graphs, PDL), plus SM clocks / power sampled while one phase is looped.
  python tools_phases.py [--clips 64 --seconds 10 --reps 5]"""
import argparse, os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import config, synthetic
from artalk_b200.engine import ARTAvatarInferEngine, smooth_motion

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64); ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--reps", type=int, default=5); ap.add_argument("--loop-s", type=float, default=1.5)
a = ap.parse_args()
cfg = config.FULL
dev = "cuda:0"
eng = ARTAvatarInferEngine(load_gaga=False, device=dev, precision="bf16", state_dict=synthetic.make_state_dict(cfg, 0),
                           config=cfg.to_reference_json(), flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec,
                           make_output_dir=False)
m = eng.ARTalk
if os.environ.get("ARTALK_WS_LIMIT_MB"):          # developer switch: smaller wav2vec sub-batches (activations closer to L2 size)
    m.set_workspace_limit(int(os.environ["ARTALK_WS_LIMIT_MB"]) << 20)
B = a.clips
S = int(a.seconds * 16000)
n_chunks = cfg.chunks_for_samples(S)
audio = synthetic.make_audio(B, S).to(dev)
style_m = synthetic.make_style_motion(B).to(dev)
pad = n_chunks * cfg.chunk_samples - S
audio_p = torch.cat([audio, audio.new_zeros(B, pad)], dim=-1).reshape(B * n_chunks, cfg.chunk_samples).contiguous()
for _ in range(3):
    eng.inference_batch(audio, style_m)
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


acc = {}
for rep in range(a.reps):
    marks = [("start", ev())]
    marks[0][1].record()
    cond = m.audio_cond(audio_p).view(B, n_chunks, cfg.seq_tokens, cfg.cond_dim)
    e = ev(); e.record(); marks.append(("wav2vec (all %d chunks)" % (B * n_chunks), e))
    style = m.style_cond(style_m, B)
    e = ev(); e.record(); marks.append(("style encoder", e))
    prev = m.initial_words(B)
    out = torch.empty(B, cfg.chunk_frames, cfg.motion_dim, device=dev)
    motion = torch.empty(B, n_chunks, cfg.chunk_frames, cfg.motion_dim, device=dev)
    for c in range(n_chunks):
        m.ar_chunk(cond[:, c], style, prev, out)
        motion[:, c].copy_(out)
        e = ev(); e.record(); marks.append(("AR chunk %d" % c, e))
    smooth_motion(motion.view(B, -1, cfg.motion_dim)[:, :cfg.frames_for_samples(S)], 750, False)
    e = ev(); e.record(); marks.append(("savgol + post", e))
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        acc.setdefault(n1, []).append(e0.elapsed_time(e1))
tot = 0.0
for k, v in acc.items():
    v.sort(); med = v[len(v) // 2]; tot += med
    print("%-32s %8.3f ms" % (k, med))
print("%-32s %8.3f ms" % ("sum", tot))


def sample_loop(name, fn):
    rows = []
    stop = threading.Event()
    def smp():
        while not stop.is_set():
            o = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True).stdout.strip()
            if o:
                rows.append([float(x) for x in o.split(",")])
            stop.wait(0.05)
    t = threading.Thread(target=smp, daemon=True); t.start()
    t0 = time.time(); n = 0
    e0, e1 = ev(), ev()
    e0.record()
    while time.time() - t0 < a.loop_s:
        fn(); n += 1
        if n % 4 == 0:
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); t.join()
    rows = rows[len(rows) // 3:]
    sm = sorted(r[0] for r in rows); pw = sorted(r[1] for r in rows)
    print("loop %-24s %8.3f ms/iter  sm clock median %4.0f MHz (min %4.0f)  power median %4.0f W" %
          (name, e0.elapsed_time(e1) / n, sm[len(sm) // 2] if sm else 0, sm[0] if sm else 0, pw[len(pw) // 2] if pw else 0))


sample_loop("wav2vec", lambda: m.audio_cond(audio_p))
cond = m.audio_cond(audio_p).view(B, n_chunks, cfg.seq_tokens, cfg.cond_dim)
style = m.style_cond(style_m, B)
prev = m.initial_words(B)
out = torch.empty(B, cfg.chunk_frames, cfg.motion_dim, device=dev)
sample_loop("AR chunk", lambda: m.ar_chunk(cond[:, 0], style, prev, out))
sample_loop("full step", lambda: eng.inference_batch(audio, style_m))
m.enable_graphs(False)
sample_loop("AR chunk (eager launches)", lambda: m.ar_chunk(cond[:, 0], style, prev, out))
m.enable_graphs(True)
