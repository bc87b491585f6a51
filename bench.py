"""Benchmark of the audio->motion hot path (BASELINE.json metric: motion frames/sec).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

A step = one pass of the path (wav2vec2 -> AR scale loop with KV cache -> VAE decode/re-encode -> savgol post-ops)
over one batch of synthetic clips. Default workload = BASELINE.json configs[1]: 64 synthetic 10 s 16 kHz clips, bf16,
one B200 (weak scaling: every rank gets its own 64 clips; one NCCL all-gather of the motion tensors per step).
Prints ONE JSON line (rank 0).
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

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from artalk_b200 import config, synthetic  # noqa: E402

GFLOP_PER_CHUNK = 216.0          # algorithmic 2*MAC per 100-frame chunk per clip
METRIC = "motion_frames_per_sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


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


def cpu_reference_run(cfg, n_samples, reps, with_style=True):
    """The oracle port (reference schedule, fp32, all host threads) on a bounded sample: one clip per repetition."""
    from oracle.artalk_oracle import Oracle
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.make_state_dict(cfg, 0)
    orc = Oracle(sd, cfg)
    audio = synthetic.make_audio(1, n_samples)
    style = synthetic.make_style_motion(1) if with_style else None
    frames = cfg.frames_for_samples(n_samples)
    with torch.no_grad():
        orc.inference(audio, style)                      # warm-up
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.smooth_savgol(orc.inference(audio, style)[0])
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return frames / med, med, torch.get_num_threads()


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--clips", type=int, default=64, help="clips per GPU")
    ap.add_argument("--seconds", type=float, default=10.0, help="clip length")
    ap.add_argument("--config", default="FULL", choices=["FULL", "TINY"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=1, help="independent clip lanes (streams) per GPU")
    ap.add_argument("--profile-step", action="store_true",
                    help="profiling aid: run only the warm-up and timed resident steps (no e2e / instrumented / CPU legs), "
                         "so an ncu launch list of the command covers whole steps and nothing else")
    args = ap.parse_args()
    cfg = getattr(config, args.config)
    n_samples = int(args.seconds * cfg.sample_rate)
    frames = cfg.frames_for_samples(n_samples)
    n_chunks = cfg.chunks_for_samples(n_samples)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "%d synthetic %g s 16 kHz clips per GPU (%d frames, %d chunks each), %s, config %s" % (
        args.clips, args.seconds, frames, n_chunks, args.precision, args.config)

    if args.impl == "reference":
        if rank != 0:
            return
        reps = max(1, args.steps)
        for _ in range(max(0, args.warmup - 1)):
            pass                                          # cpu_reference_run does its own warm-up pass
        fps, med, cores = cpu_reference_run(cfg, n_samples, reps)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "fp32", "data": "synthetic", "gpu_launches": 0,
                "config": {"workload": workload, "note": "reference algorithm (oracle port of the un-cached schedule) on host CPU; "
                           "each step = 1 clip (bounded sample of the batch; the reference only supports batch 1)"},
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": "1 clip x %g s per step, median of %d; %s" % (args.seconds, reps, cpu_model_name())},
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
                               state_dict=synthetic.make_state_dict(cfg, 0), config=cfg.to_reference_json(),
                               flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec, make_output_dir=False, lanes=args.lanes)
    lib = _lib.lib()
    B = args.clips
    audio_host = synthetic.make_audio(B, n_samples, first_clip=rank * B).pin_memory()
    style_host = synthetic.make_style_motion(B, first_clip=rank * B).pin_memory()
    audio_dev, style_dev = audio_host.to(dev), style_host.to(dev)
    out_host = torch.empty(B, frames, 106).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step_resident():
        m = eng.inference_batch(audio_dev, style_dev)
        return parallel.gather_motion(m, world * B) if world > 1 else m      # the path's only collective

    def step_e2e():
        # pinned host buffers straight into the public call: it uploads the style clips, runs the style encoder while the
        # 41 MB audio upload proceeds on its copy stream, then wav2vec waits for the upload
        m = eng.inference_batch(audio_host, style_host)
        if world > 1:
            parallel.gather_motion(m, world * B)
        out_host.copy_(m, non_blocking=True)
        return m

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s0, s1 in ev:
            flush.fill_(1)                                   # L2 flush between timed iterations (outside the events)
            s0.record()
            fn()
            s1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.artalk_launch_count()
    total_ms = timed(step_resident, args.steps, args.warmup)
    launches = (lib.artalk_launch_count() - l0) * args.steps // (args.steps + args.warmup)
    if args.profile_step:
        if sampler:
            sampler.stop_flag.set(); sampler.join()
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": world * B * frames * args.steps / (total_ms / 1e3), "unit": "frames/s",
                              "ms_per_step": total_ms / args.steps, "gpu_launches": int(launches), "profile_step": True}))
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_ms = timed(step_e2e, args.steps, 1)
    if sampler:
        sampler.stop_flag.set()
        sampler.join()
    value = world * B * frames * args.steps / (total_ms / 1e3)
    e2e_value = world * B * frames * args.steps / (e2e_ms / 1e3)

    # instrumented pass: CUDA events around every GEMM / attention launch of one step (not part of the timed numbers)
    eng.ARTalk.lanes = 1                                 # serial launches on one engine handle for the attribution pass
    h = eng.ARTalk._handle()
    _lib.check(lib.artalk_profile_enable(h, 1))
    torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); step_resident(); t1.record()
    prof = (C.c_double * 12)()
    _lib.check(lib.artalk_profile_read(h, prof, _lib.stream_ptr(dev)))
    _lib.check(lib.artalk_profile_enable(h, 0))
    torch.cuda.synchronize(dev)
    inst_ms = t0.elapsed_time(t1)
    hbm, tf_burst, tf_sus, src = peaks()
    n_g, ms_g, fl_g, n_a, ms_a, fl_a = [prof[i] for i in range(6)]
    n_top, ms_top, fl_top = prof[6], prof[7], prof[8]          # dominant GEMM shape (largest summed duration)
    gemm_tflops = fl_g / (ms_g * 1e-3) / 1e12 if ms_g > 0 else 0.0
    top_tflops = fl_top * n_top / (ms_top * 1e-3) / 1e12 if ms_top > 0 else 0.0
    peak = tf_sus if args.precision == "bf16" else 72.0
    # ncu --set full capture of the dominant launches (profiles/r1h_ncu_full_summary.md): dram read + write per launch;
    # only valid for the default workload's wav2vec FFN GEMMs (M = clips*chunks*199 = 38208, N x K = 4096 x 1024 / 1024 x 4096)
    is_default_top = args.precision == "bf16" and abs(fl_top - 2.0 * 38208 * 4096 * 1024) < 1.0
    roofline = {"bound": "tensor",
                "kernel": ("gemm_tc2_kernel (tcgen05 cta_group::2 bf16 GEMM, 256x256 tiles over CTA pairs), dominant shape of the "
                           "step: %.1f GFLOP per launch, %d launches" % (fl_top / 1e9, int(n_top))) if args.precision == "bf16"
                else "gemm_simt_kernel (fp32 CUDA-core GEMM)",
                "achieved": top_tflops, "peak": peak, "unit": "TFLOP/s", "frac": top_tflops / peak,
                "traffic": 506.2e6 if is_default_top else None,
                "traffic_note": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the two wav2vec FFN GEMMs that "
                                "share this flop count: FFN1 38208x4096x1024 86.8+263.9 MB (algorithmic 78 A + 8 W + 313 out), "
                                "FFN2 38208x1024x4096 533.2+128.5 MB (algorithmic 313 A + 8 W + 157 resid read + 157 write)"
                if is_default_top else None,
                "peak_source": ("%s bf16_tflops_sustained (kernel timed inside a long step)" % src) if args.precision == "bf16"
                else "nominal fp32 FMA peak 148 SMs x 128 lanes x 2 x 1.9 GHz",
                "flops_per_launch": fl_top, "ms_per_launch": ms_top / max(n_top, 1),
                "all_gemm_launches": {"launches": int(n_g), "achieved": gemm_tflops, "frac": gemm_tflops / peak,
                                      "flops_per_launch": fl_g / max(n_g, 1), "ms_per_launch": ms_g / max(n_g, 1)},
                "share_of_step": {"gemm": ms_g / inst_ms, "gemm_dominant_shape": ms_top / inst_ms, "attention": ms_a / inst_ms,
                                  "other": max(0.0, 1 - (ms_g + ms_a) / inst_ms)},
                "attention_tflops": fl_a / (ms_a * 1e-3) / 1e12 if ms_a > 0 else 0.0,
                "path_tflops": B * n_chunks * GFLOP_PER_CHUNK * 1e9 / (total_ms / args.steps * 1e-3) / 1e12,
                "path_frac_of_peak": B * n_chunks * GFLOP_PER_CHUNK * 1e9 / (total_ms / args.steps * 1e-3) / 1e12 / tf_sus}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload, "l2": "256 MiB flush write between timed iterations", "weights": "seeded random init (no checkpoints ship)", "lanes": args.lanes,
                       "parallelism": "clip-sharded replicas, 1 process/GPU, one NCCL all-gather of motion per step" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": audio_host.numel() * 4 + style_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "clocks": sampler.summary() if sampler else None, "roofline": roofline}
    if not args.no_cpu_baseline and world >= 1:
        fps, med, cores = cpu_reference_run(cfg, n_samples, 3)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "1 clip x %g s (%d frames), median of 3 after 1 warm-up, oracle port of the reference "
                                          "schedule, fp32, %s" % (args.seconds, frames, cpu_model_name())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
