"""Benchmark of the audio->motion hot path (BASELINE.json metric: motion frames/sec; p50 latency per chunk).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation on the host cores

A step = one pass of the path (style encoder -> wav2vec2 -> AR scale loop with KV cache -> VAE decode / re-encode ->
Savitzky-Golay post-ops -> FLAME blendshape + LBS mesh decode of every frame) over one batch of synthetic clips.
Default workload = BASELINE.json configs[2]: 256 synthetic 30 s 16 kHz clips (clip_length 750, 8 chunks each) per GPU, bf16
(weak scaling: every rank gets its own 256 clips; one NCCL all-gather of the motion tensors per step, on a side stream).
Sub-records of the same JSON line: `latency` (configs[4]: batch-1 streaming, p50/p90 per 4 s chunk incl. FLAME mesh + D2H;
N = 1 only), `strong_4096` (configs[3]: 4096 clips sharded over the ranks), `parity` (flip rate of this precision against
the live reference's stored outputs, teacher-forced), `roofline`, `cpu_baseline`. Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from artalk_b200 import config, synthetic  # noqa: E402

GFLOP_PER_CHUNK = 216.0          # algorithmic 2*MAC per 100-frame chunk per clip
FLAME_BYTES_PER_FRAME = 60276 + 424      # vertices out + codes in
METRIC = "motion_frames_per_sec"


_SD = {}


def state_dict(cfg):
    """Seeded synthetic checkpoint (2 GB fp32 for FULL), generated once per process."""
    if id(cfg) not in _SD:
        _SD[id(cfg)] = synthetic.make_state_dict(cfg, 0)
    return _SD[id(cfg)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d["bf16_tflops_sustained"], "MEASURED_PEAKS.json"
    return 6650.0, 1590.0, 1400.0, "B200_PROFILING.md fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
def reference_runner(cfg):
    """The reference's own CPU implementation of the path. The unmodified reference tree is imported when it is present
    ($ARTALK_REFERENCE, baseline/_ref/ or /root/reference: kind "reference"); it is a Python script tree that cannot travel
    to the GPU box, so there the oracle port of the same (un-cached) schedule runs instead (kind "port"). Returns
    (kind, run) where run(audio (S,), style (50,106)) -> (frames, 106) smoothed motion, fp32, all host threads."""
    torch.set_num_threads(os.cpu_count() or 1)
    sd = state_dict(cfg)
    for root in (os.environ.get("ARTALK_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, "app", "models.py")):
            os.environ["ARTALK_REFERENCE"] = root
            from oracle import reference_live as live
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                eng = live.load_engine(cfg, sd, synthetic.make_flame_asset(0), clip_length=750)

            def run_live(audio, style):
                eng.set_style_motion(style)
                return eng.inference(audio)
            return "reference", run_live
    from oracle.artalk_oracle import Oracle
    orc = Oracle(sd, cfg)

    def run_port(audio, style):
        with torch.no_grad():
            return orc.engine_inference(audio, style[None], clip_length=750)
    return "port", run_port


def cpu_reference_run(cfg, n_samples, reps, warmup=1):
    """Bounded sample of the workload: one clip per repetition (the reference only supports batch 1, app/models.py:65)."""
    kind, run = reference_runner(cfg)
    audio = synthetic.make_audio(1, n_samples)[0]
    style = synthetic.make_style_motion(1)[0]
    frames = min(750, cfg.frames_for_samples(n_samples))
    import contextlib
    import io
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints progress lines
            out = run(audio, style)
        dt = time.perf_counter() - t0
        assert out.shape[0] == frames
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return frames / med, med, torch.get_num_threads(), kind


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------ sub-records
def parity_record(eng, precision, dev):
    """This build's bits against the live reference's stored outputs (tests/golden/full_10s.npz: 10 s clip, full-depth model,
    the fixture's inputs are synthetic.make_audio(1, 160000) / make_style_motion(1)), teacher-forced so one flip does not
    cascade: flip rate over all 17 376 sampled bits, the largest reference margin among flipped bits, logits error."""
    path = os.path.join(ROOT, "tests", "golden", "full_10s.npz")
    if not os.path.exists(path):
        return None
    from artalk_b200.model import unpack_words
    g = np.load(path)
    logits = torch.from_numpy(g["logits"])
    l2 = logits.reshape(*logits.shape[:-1], 32, 2)
    margin = (l2[..., 1] - l2[..., 0]).abs()
    gold_words = torch.from_numpy(g["bits"].view(np.int32).copy())
    gold_prev = torch.from_numpy(g["prev_bits"].view(np.int32).copy())
    gb = unpack_words(gold_words)
    tr = {}
    out = eng.ARTalk.inference({"audio": synthetic.make_audio(1, 160000), "style_motion": synthetic.make_style_motion(1)},
                               trace=tr, teacher_words=gold_words, teacher_prev_words=gold_prev)
    torch.cuda.synchronize(dev)
    bits = unpack_words(tr["words"]).cpu()
    flips = bits != gb
    lerr = (tr["logits"].cpu() - logits).abs()
    worst = float(margin[flips].max()) if bool(flips.any()) else 0.0
    return {"mode": precision, "fixture": "tests/golden/full_10s.npz (live reference, fp32 CPU)", "teacher_forced": True,
            "bits": int(flips.numel()), "flips": int(flips.sum()), "flip_rate_teacher_forced": float(flips.float().mean()),
            "margin_exact_above": worst, "flips_above_margin_1e-3": int((flips & (margin > 1e-3)).sum()),
            "logits_max_err": float(lerr.max()), "logits_mean_err": float(lerr.mean()),
            "motion_max_err": float((out.cpu() - torch.from_numpy(g["motion"])).abs().max()),
            "contract": "bits exact where the reference margin > 1e-3, motion within 1e-3 (fp32 / bf16x3 / bf16x6); 2e-2 motion (bf16)"}


def latency_record(cfg, dev, precision, n_chunks=45, skip=5):
    """BASELINE configs[4]: streaming batch-1 inference, one 4 s / 100-frame chunk at a time, host audio in, motion on the
    host + FLAME mesh vertices on the device out: H2D of the chunk -> wav2vec2 -> AR chunk (KV cache, CUDA graph) -> VAE decode
    / re-encode -> FLAME LBS for the chunk's 100 frames -> D2H of the motion, wall clock around a device synchronize."""
    from artalk_b200.engine import ARTAvatarInferEngine
    eng = ARTAvatarInferEngine(load_gaga=False, device=str(dev), precision=precision, state_dict=state_dict(cfg),
                               config=cfg.to_reference_json(), flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec,
                               make_output_dir=False, latency_mode=(precision == "bf16"))
    m = eng.ARTalk
    audio = synthetic.make_audio(1, cfg.chunk_samples * n_chunks).pin_memory()
    style = m.style_cond(synthetic.make_style_motion(1), 1)
    prev = m.initial_words(1)
    out = torch.empty(1, 100, 106, device=dev)
    host = torch.empty(1, 100, 106).pin_memory()
    shape = torch.zeros(1, 300, device=dev).expand(100, -1)
    verts = torch.empty(100, 5023, 3, device=dev)
    lat, parts = [], []
    for c in range(n_chunks):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        chunk = audio[:, c * cfg.chunk_samples:(c + 1) * cfg.chunk_samples].to(dev, non_blocking=True)
        cond = m.audio_cond(chunk)
        torch.cuda.synchronize(dev); t1 = time.perf_counter()
        m.ar_chunk(cond, style, prev, out)
        torch.cuda.synchronize(dev); t2 = time.perf_counter()
        eng.mesh_vertices(out.view(100, 106), out=verts)
        host.copy_(out, non_blocking=True)
        torch.cuda.synchronize(dev); t3 = time.perf_counter()
        if c >= skip:
            lat.append((t3 - t0) * 1e3); parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    n_graphs, n_replays, failure = m.graph_status()
    m.close()
    lat.sort()
    p = lambda q: lat[min(len(lat) - 1, int(q * len(lat)))]
    med = lambda i: sorted(x[i] for x in parts)[len(parts) // 2]
    return {"workload": "configs[4]: streaming batch 1, %d chunks of 4 s / 100 frames timed after %d warm-up chunks, FLAME mesh "
                        "(100 x 5023 x 3 vertices, left on the device) + motion D2H inside" % (len(lat), skip),
            "precision": precision, "p50_ms_per_4s_chunk": p(0.5), "p90_ms_per_4s_chunk": p(0.9), "p50_ms_per_2s": p(0.5) / 2,
            "p90_ms_per_2s": p(0.9) / 2, "p50_parts_ms": {"h2d+wav2vec": med(0), "ar+vae": med(1), "flame+d2h": med(2)},
            "frames_per_sec_stream": 100 / (p(0.5) / 1e3), "graph_replays": n_replays, "graph_failure": failure,
            "note": "the reference's chunk is 4 s / 100 frames (app/models.py:19,76-81); the metric's 'per 2 s' figure is half a chunk"}


def wav2vec_gemm_bytes(M, N, K):
    """Algorithmic HBM bytes of one launch of the wav2vec layer GEMMs: bf16 A and W; QKV / FFN1 write a
    bf16 activation, out-projection / FFN2 read-modify-write the fp32 residual stream."""
    a, w = 2.0 * M * K, 2.0 * N * K
    if (N, K) in ((3072, 1024), (4096, 1024)):
        return a + w + 2.0 * M * N
    if (N, K) in ((1024, 1024), (1024, 4096)):
        return a + w + 8.0 * M * N
    return None


def measured_traffic(M, N, K):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this build's
    kernels (profiles/r2_ncu_dominant.json, written by tools_ncu_summary.py from the .ncu-rep); None if no capture matches."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_dominant.json")
    if not os.path.exists(path):
        return None, None
    try:
        d = json.load(open(path))
        for e in d.get("launches", []):
            if (e["M"], e["N"], e["K"]) == (M, N, K):
                return e["dram_bytes"], "profiles/r2_ncu_dominant.json (%s)" % d.get("source", "ncu --set full")
    except Exception:
        pass
    return None, None


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3", "bf16x6"],
                    help="bf16: tcgen05 (headline); bf16x3 / bf16x6: tcgen05 at the reference's bit-exact contract; fp32: CUDA cores")
    ap.add_argument("--clips", type=int, default=256, help="clips per GPU")
    ap.add_argument("--seconds", type=float, default=30.0, help="clip length")
    ap.add_argument("--config", default="FULL", choices=["FULL", "TINY"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mesh", action="store_true", help="leave the FLAME mesh decode out of the step")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the configs[3] strong-scaling sub-record")
    ap.add_argument("--strong-clips", type=int, default=4096)
    ap.add_argument("--profile-step", action="store_true",
                    help="profiling aid: run only the warm-up and timed resident steps (no e2e / instrumented / CPU legs), "
                         "so an ncu launch list of the command covers whole steps and nothing else")
    args = ap.parse_args()
    cfg = getattr(config, args.config)
    n_samples = int(args.seconds * cfg.sample_rate)
    frames = min(750, cfg.frames_for_samples(n_samples))
    n_chunks = cfg.chunks_for_samples(n_samples)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    which = {(256, 30.0): "BASELINE configs[2]: ", (64, 10.0): "BASELINE configs[1]: "}.get((args.clips, args.seconds), "")
    workload = "%s%d synthetic %g s 16 kHz clips per GPU (%d frames, %d chunks each), %s, config %s" % (
        which, args.clips, args.seconds, frames, n_chunks, args.precision, args.config)

    mesh = not args.no_mesh
    B = args.clips
    # identical for both arms (the driver compares it): the reference arm's step is a bounded sample of this workload
    cfg_rec = {"workload": workload, "l2": "256 MiB flush write between timed iterations",
               "weights": "seeded random init (no checkpoints ship)",
               "step": "style encoder + wav2vec2 + AR + VAE + savgol post-ops" +
                       (" + FLAME mesh decode of all %d frames (vertices stay on the device)" % (B * frames) if mesh else " (FLAME mesh decode excluded)"),
               "parallelism": "clip-sharded replicas, 1 process/GPU, one NCCL all-gather of motion per step on a side stream" if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        fps, med, cores, kind = cpu_reference_run(cfg, n_samples, max(1, args.steps), max(1, args.warmup))
        sample = "1 clip x %g s (%d frames) per step, median of %d after %d warm-up; %s" % (
            args.seconds, frames, max(1, args.steps), max(1, args.warmup), cpu_model_name())
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "fp32", "data": "synthetic", "gpu_launches": 0,
                "config": cfg_rec,
                "reference_note": (("the unmodified reference (ARTAvatarInferEngine.inference, CPU, fp32) imported from %s" %
                                    os.environ.get("ARTALK_REFERENCE")) if kind == "reference" else
                                   "reference algorithm (oracle port of the un-cached schedule, incl. savgol) on the host CPU: the "
                                   "reference is a Python script tree that is not present on this box") +
                                  "; each step = 1 clip (bounded sample of the batch; the reference only supports batch 1, "
                                  "app/models.py:65) without the FLAME mesh decode",
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product has no CPU path)"
    import torch.distributed as dist
    from artalk_b200 import _lib, parallel
    from artalk_b200.engine import ARTAvatarInferEngine
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    eng = ARTAvatarInferEngine(load_gaga=False, clip_length=750, device=str(dev), precision=args.precision,
                               state_dict=state_dict(cfg), config=cfg.to_reference_json(),
                               flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec, make_output_dir=False)
    lib = _lib.lib()
    audio_host = synthetic.make_audio(B, n_samples, first_clip=rank * B).pin_memory()
    style_host = synthetic.make_style_motion(B, first_clip=rank * B).pin_memory()
    audio_dev, style_dev = audio_host.to(dev), style_host.to(dev)
    out_host = torch.empty(B, frames, 106).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    verts = torch.empty(B * frames, 5023, 3, device=dev) if mesh else None          # 60 KB per frame, stays on the device
    gather = parallel.MotionGather(dev) if world > 1 else None
    d2h_stream = torch.cuda.Stream(device=dev)
    flame_ev = []

    def decode_mesh(m, timed_events=False):
        if not mesh:
            return
        if timed_events:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        eng.mesh_vertices(m.reshape(-1, 106), out=verts)
        if timed_events:
            b.record()
            flame_ev.append((a, b))

    def step_resident(timed_events=False):
        m = eng.inference_batch(audio_dev, style_dev)
        decode_mesh(m, timed_events)
        if gather is not None:
            gather.start(m, world * B)           # the path's only collective: side stream, overlaps the next step
        return m

    def step_e2e(timed_events=False):
        # pinned host buffers straight into the public call: it uploads the style clips, runs the style encoder while the
        # audio upload proceeds on its copy stream, then wav2vec waits for the upload; the result (motion) goes back to the host
        m = eng.inference_batch(audio_host, style_host)
        # the motion goes back to the host on a side stream while the mesh decode (which leaves its vertices on the device) runs
        main = torch.cuda.current_stream(dev)
        d2h_stream.wait_stream(main)
        with torch.cuda.stream(d2h_stream):
            out_host.copy_(m, non_blocking=True)
        m.record_stream(d2h_stream)
        decode_mesh(m)
        if gather is not None:
            gather.start(m, world * B)
        main.wait_stream(d2h_stream)                     # the step ends when the result is on the host
        return m

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if gather is not None:
            gather.wait()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        if gather is not None:
            gather.times_ms.clear()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i, (s0, s1) in enumerate(ev):
            flush.fill_(1)                                   # L2 flush between timed iterations (outside the events)
            s0.record()
            fn(True)
            if gather is not None and i == steps - 1:
                gather.wait()                                # the last step's collective ends inside the timed region
            s1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        mine = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([mine, -mine], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[0].item() + t[1].item())        # max over ranks, max - min over ranks

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.artalk_launch_count()
    total_ms, spread_ms = timed(step_resident, args.steps, args.warmup)
    launches = (lib.artalk_launch_count() - l0) * args.steps // (args.steps + args.warmup)
    gather_ms = gather.mean_ms() if gather is not None else 0.0
    flame_ms = (sum(a.elapsed_time(b) for a, b in flame_ev) / len(flame_ev)) if flame_ev else 0.0
    n_graphs, n_replays, graph_failure = eng.ARTalk.graph_status()
    if args.profile_step:
        if sampler:
            sampler.stop_flag.set(); sampler.join()
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": world * B * frames * args.steps / (total_ms / 1e3), "unit": "frames/s",
                              "ms_per_step": total_ms / args.steps, "gpu_launches": int(launches), "profile_step": True}))
        if world > 1:
            dist.destroy_process_group()
        return
    flame_ev.clear()
    e2e_ms, _ = timed(step_e2e, args.steps, 1)
    if sampler:
        sampler.stop_flag.set()
        sampler.join()
    value = world * B * frames * args.steps / (total_ms / 1e3)
    e2e_value = world * B * frames * args.steps / (e2e_ms / 1e3)

    # instrumented pass: CUDA events around every GEMM / attention launch of one step (not part of the timed numbers)
    h = eng.ARTalk._handle()
    _lib.check(lib.artalk_profile_enable(h, 1))
    torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); eng.inference_batch(audio_dev, style_dev); t1.record()
    prof = (C.c_double * 16)()
    _lib.check(lib.artalk_profile_read(h, prof, _lib.stream_ptr(dev)))
    _lib.check(lib.artalk_profile_enable(h, 0))
    torch.cuda.synchronize(dev)
    inst_ms = t0.elapsed_time(t1)
    hbm, tf_burst, tf_sus, src = peaks()
    n_g, ms_g, fl_g, n_a, ms_a, fl_a = [prof[i] for i in range(6)]
    n_top, ms_top, fl_top = prof[6], prof[7], prof[8]          # dominant GEMM shape (largest summed duration)
    Mt, Nt, Kt = int(prof[9]), int(prof[10]), int(prof[11])
    gemm_tflops = fl_g / (ms_g * 1e-3) / 1e12 if ms_g > 0 else 0.0
    top_tflops = fl_top * n_top / (ms_top * 1e-3) / 1e12 if ms_top > 0 else 0.0
    tc = args.precision != "fp32"
    peak = tf_sus if tc else 72.0
    traffic, traffic_src = measured_traffic(Mt, Nt, Kt)
    alg_bytes = wav2vec_gemm_bytes(Mt, Nt, Kt) if args.precision == "bf16" else None
    step_ms = total_ms / args.steps
    roofline = {"bound": "tensor",
                "kernel": ("gemm_tc2_kernel (tcgen05 cta_group::2 bf16 GEMM, 256x256 tiles over CTA pairs)%s, dominant shape of the step: "
                           "M=%d N=%d K=%d, %.1f GFLOP per launch, %d launches" %
                           ("" if args.precision == "bf16" else " on %s piece blocks (flops counted once: algorithmic)" % args.precision,
                            Mt, Nt, Kt, fl_top / 1e9, int(n_top))) if tc else "gemm_simt_kernel (fp32 CUDA-core GEMM)",
                "achieved": top_tflops, "peak": peak, "unit": "TFLOP/s", "frac": top_tflops / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": ("%s bf16_tflops_sustained (kernel timed inside a long step)" % src) if tc
                else "nominal fp32 FMA peak 148 SMs x 128 lanes x 2 x 1.9 GHz",
                "flops_per_launch": fl_top, "ms_per_launch": ms_top / max(n_top, 1),
                "all_gemm_launches": {"launches": int(n_g), "achieved": gemm_tflops, "frac": gemm_tflops / peak,
                                      "flops_per_launch": fl_g / max(n_g, 1), "ms_per_launch": ms_g / max(n_g, 1)},
                "share_of_step": {"gemm": ms_g / inst_ms, "gemm_dominant_shape": ms_top / inst_ms, "attention": ms_a / inst_ms,
                                  "other": max(0.0, 1 - (ms_g + ms_a) / inst_ms), "flame_mesh_of_timed_step": flame_ms / step_ms},
                "attention_tflops": fl_a / (ms_a * 1e-3) / 1e12 if ms_a > 0 else 0.0,
                "flame_mesh": {"ms_per_step": flame_ms, "frames": B * frames, "bound": "hbm",
                               "achieved_gbs": B * frames * FLAME_BYTES_PER_FRAME / (flame_ms * 1e-3) / 1e9 if flame_ms > 0 else None,
                               "peak_gbs": hbm, "frac": (B * frames * FLAME_BYTES_PER_FRAME / (flame_ms * 1e-3) / 1e9 / hbm) if flame_ms > 0 else None},
                "path_tflops": B * n_chunks * GFLOP_PER_CHUNK * 1e9 / (step_ms * 1e-3) / 1e12,
                "path_frac_of_peak": B * n_chunks * GFLOP_PER_CHUNK * 1e9 / (step_ms * 1e-3) / 1e12 / tf_sus}

    # ---- configs[3]: 4096 clips sharded over the ranks (strong scaling), same 30 s clips, sub-batches of `clips`
    strong = None
    if not args.no_strong and args.strong_clips > 0:
        lo, hi = parallel.shard_bounds(args.strong_clips, world, rank)
        n_sub = -(-(hi - lo) // B)
        passes = 1 if world == 1 else 2
        g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
        sg = parallel.MotionGather(dev) if world > 1 else None
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(passes):
            outs = []
            for j in range(n_sub):
                nb = min(B, hi - lo - j * B)
                a = 0.1 * torch.randn(nb, n_samples, device=dev, generator=g)       # generated per rank on the device (7.9 GB job)
                s = style_dev[:nb]
                m = eng.inference_batch(a, s)
                if mesh:
                    eng.mesh_vertices(m.reshape(-1, 106), out=verts[:nb * frames])
                outs.append(m)
            local = torch.cat(outs, 0)
            if sg is not None:
                sg.start(local, args.strong_clips)
        if sg is not None:
            sg.wait()
        s1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong_ms = float(t.item()) / passes
        strong = {"workload": "BASELINE configs[3]: %d synthetic 30 s clips sharded over %d GPU(s) (%d per rank, sub-batches of %d), "
                              "NCCL all-gather of the (clips, 750, 106) motion only" % (args.strong_clips, world, hi - lo, B),
                  "scaling": "strong", "value": args.strong_clips * frames / (strong_ms / 1e3), "unit": "frames/s",
                  "ms_per_pass": strong_ms, "passes_timed": passes, "gather_ms": sg.mean_ms() if sg is not None else 0.0,
                  "gathered_bytes": args.strong_clips * frames * 106 * 4,
                  "note": "audio generation on the device (torch.randn) is inside the timed pass"}
        del outs, local

    # ---- sub-records measured on rank 0 only
    parity = latency = None
    if rank == 0 and args.config == "FULL":
        parity = parity_record(eng, args.precision, dev)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world == 1 and not args.no_latency and args.config == "FULL":
        del verts, flush
        torch.cuda.empty_cache()
        latency = latency_record(cfg, dev, args.precision)
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": cfg_rec,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": audio_host.numel() * 4 + style_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                    "api": "ARTAvatarInferEngine.inference_batch(pinned host audio, pinned host style) -> motion copied to pinned host memory"},
            "gpu_launches": int(launches), "graphs": {"instantiated": n_graphs, "replays": n_replays, "failure": graph_failure},
            "clocks": sampler.summary() if sampler else None, "roofline": roofline,
            "multi_gpu": {"gather_ms": gather_ms, "rank_spread_ms": spread_ms / args.steps,
                          "note": "gather_ms: CUDA events around the all-gather on its side stream (mean per step); rank_spread_ms: "
                                  "max - min over ranks of the per-rank step time"} if world > 1 else None,
            "parity": parity, "latency": latency, "strong_4096": strong}
    if not args.no_cpu_baseline and world == 1:
        fps, med, cores, kind = cpu_reference_run(cfg, n_samples, 3)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                                "sample": "1 clip x %g s (%d frames), median of 3 after 1 warm-up, fp32, all host threads, %s" %
                                          (args.seconds, frames, cpu_model_name())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
