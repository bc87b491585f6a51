"""ORACLE tooling — pin ``artalk_oracle.Oracle`` against the live, unmodified reference and
write the fixtures under tests/golden/.

Run in the build container (needs /root/reference):  ``python -m oracle.make_golden``

For every case in ``oracle/cases.py`` the seeded synthetic checkpoint is loaded *strictly*
into the reference ``BitwiseARModel`` (verifies the 814-key wire format), the reference is run
clip by clip (it asserts batch 1, app/models.py:65) and its outputs are stored:
final motion, per-chunk final-step logits (forward hook on ``logits_head``), predicted bits,
re-encoded prev bits, VAE encoder output, a strided slice of the audio conditioning and the
style token. The restatement is run on the same inputs and the max deviations are written to
tests/golden/PIN_REPORT.json (the test-suite re-checks them from the fixtures on any box).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from artalk_b200 import synthetic                      # noqa: E402
from oracle import reference_live as live              # noqa: E402
from oracle.artalk_oracle import Oracle, get_flame_verts  # noqa: E402
from oracle.cases import CASES, flame_inputs, gaga_inputs          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
COND_STRIDE = 32


def pack_bits(bits: torch.Tensor) -> np.ndarray:
    """(...,32) {0,1} -> (...) uint32, bit j of the word = code dim j."""
    w = (bits.to(torch.int64) << torch.arange(32, dtype=torch.int64)).sum(dim=-1)
    return w.numpy().astype(np.uint32)


def run_case(case, report):
    cfg = case.cfg
    t0 = time.time()
    sd = synthetic.make_state_dict(cfg, case.weight_seed)
    ref = live.load_model(cfg, sd)
    orc = Oracle(sd, cfg)
    audio, style = case.audio(), case.style()
    T = cfg.chunk_frames
    gold, dev = {}, {}
    motions, logits_all, bits_all, prev_all, enc_all, cond_all, style_all, margins = [], [], [], [], [], [], [], []
    for b in range(case.n_clips):
        logs = []
        h = ref.logits_head.register_forward_hook(lambda m, i, o: logs.append(o.detach().clone()))
        sm = None if style is None else style[b:b + 1]
        with torch.no_grad():
            motion = ref.inference({"audio": audio[b:b + 1], "style_motion": sm})
        h.remove()
        n_chunks = len(logs) // len(cfg.patch_nums)
        final = torch.stack([logs[(c + 1) * len(cfg.patch_nums) - 1][0] for c in range(n_chunks)])   # (n,181,64)
        bits = final.view(n_chunks, -1, 32, 2).argmax(-1)
        mg = (final.view(n_chunks, -1, 32, 2)[..., 1] - final.view(n_chunks, -1, 32, 2)[..., 0]).abs()
        # stage-wise reference values through its own sub-modules
        seq, chunks = orc.split_chunks(audio[b:b + 1])
        with torch.no_grad():
            conds = []
            for ch in chunks:
                f = ref.audio_encoder(ch).permute(0, 2, 1)
                conds.append(torch.cat([F.interpolate(f, size=pn, mode="area").permute(0, 2, 1)
                                        for pn in cfg.patch_nums], dim=1)[0])
            cond = torch.stack(conds)
            if sm is not None:
                st = ref.style_cond_embed(ref.style_encoder(sm))[:, None] * 1.1 - ref.null_style_cond * 0.1
            else:
                st = ref.null_style_cond
            # un-truncated per-chunk motion is not returned by the reference; redo the VAE legs
            prev_bits, encs, prevs = None, [], []
            pm = torch.zeros(1, T, cfg.motion_dim)
            pb, _ = ref.basic_vae.quant_to_vqidx(pm, this_motion=None)
            for c in range(n_chunks):
                _, mo = ref.basic_vae.vqidx_to_motion(pb, bits[c:c + 1])
                x = ref.basic_vae.norm_with_stats(mo) + ref.basic_vae.enc_pos_embed[:, :T]
                encs.append(ref.basic_vae.encoder(x, attn_mask=None)[0])
                pb, _ = ref.basic_vae.quant_to_vqidx(mo, this_motion=None)
                prevs.append(pb[0])
        motions.append(motion[0]); logits_all.append(final); bits_all.append(bits)
        prev_all.append(torch.stack(prevs)); enc_all.append(torch.stack(encs))
        cond_all.append(cond); style_all.append(st.detach().reshape(-1)); margins.append(mg)
    gold["motion"] = torch.stack(motions).numpy()
    gold["logits"] = torch.stack(logits_all).numpy()
    gold["bits"] = pack_bits(torch.stack(bits_all))
    gold["prev_bits"] = pack_bits(torch.stack(prev_all))
    gold["enc_out"] = torch.stack(enc_all).numpy()
    cond_full = torch.stack(cond_all)
    gold["cond_slice"] = cond_full[..., ::COND_STRIDE].contiguous().numpy()
    gold["cond_rowsum"] = cond_full.sum(dim=-1).numpy()
    gold["style"] = torch.stack(style_all).numpy()
    # the restatement on the same inputs
    tr = {}
    t1 = time.time()
    with torch.no_grad():
        om = orc.inference(audio, style, trace=tr)
    t_orc = time.time() - t1
    o_logits = torch.stack(tr["logits"], dim=1)
    o_bits = torch.stack(tr["bits"], dim=1)
    mg_all = torch.stack(margins)
    safe = mg_all > 1e-3
    dev["motion_maxabs"] = float((om - torch.stack(motions)).abs().max())
    dev["logits_maxabs"] = float((o_logits - torch.stack(logits_all)).abs().max())
    dev["cond_maxabs"] = float((torch.stack(tr["cond"], dim=1) - cond_full).abs().max())
    dev["style_maxabs"] = float((tr["style"].reshape(case.n_clips, -1) - torch.stack(style_all)).abs().max())
    dev["enc_out_maxabs"] = float((torch.stack(tr["enc_out"], dim=1) - torch.stack(enc_all)).abs().max())
    dev["bit_mismatch_margin_gt_1e-3"] = int(((o_bits != torch.stack(bits_all)) & safe).sum())
    dev["bit_mismatch_total"] = int((o_bits != torch.stack(bits_all)).sum())
    dev["prev_bit_mismatch_total"] = int((torch.stack(tr["prev_bits"], dim=1) != torch.stack(prev_all)).sum())
    dev["frac_margin_lt_1e-3"] = float((~safe).float().mean())
    dev["median_margin"] = float(mg_all.median())
    dev["n_state_dict_keys"] = len(sd)
    dev["oracle_seconds"] = round(t_orc, 2)
    dev["total_seconds"] = round(time.time() - t0, 2)
    report[case.name] = dev
    np.savez_compressed(os.path.join(GOLD, case.name + ".npz"), **gold)
    print(case.name, json.dumps(dev))
    return sd, ref


def write_eng1_audio():
    """demo/eng1.wav exactly as the reference's CLI feeds it (inference.py:229-231): torchaudio.load is unusable here (needs
    torchcodec), so the PCM is read with the stdlib and scaled like torchaudio does (int16 / 32768), then
    ``torchaudio.transforms.Resample(sr, 16000)(audio).mean(dim=0)``."""
    import wave
    import torchaudio
    w = wave.open(os.path.join(live.REF_ROOT, "demo", "eng1.wav"))
    assert w.getsampwidth() == 2
    pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).reshape(-1, w.getnchannels()).T
    audio = torch.from_numpy(pcm.astype(np.float32) / 32768.0)
    audio = torchaudio.transforms.Resample(w.getframerate(), 16000)(audio).mean(dim=0)
    np.savez_compressed(os.path.join(GOLD, "eng1_audio.npz"), audio=audio[None].numpy())
    print("eng1_audio", tuple(audio.shape))


def run_engine_case(report, name="tiny_style", clip_length=120):
    """ARTAvatarInferEngine.inference incl. savgol + zeroing (inference.py:47-57)."""
    if name != "tiny_style":
        return run_engine_case_named(report, name, clip_length)
    case = CASES["tiny_style"]
    cfg = case.cfg
    sd = synthetic.make_state_dict(cfg, case.weight_seed)
    asset = synthetic.make_flame_asset(0)
    eng = live.load_engine(cfg, sd, asset, clip_length=120)
    eng.set_style_motion(case.style()[0])
    audio = case.audio()[0]
    out = eng.inference(audio)
    orc = Oracle(sd, cfg)
    with torch.no_grad():
        o = orc.engine_inference(audio, case.style()[0:1], clip_length=120)
    np.savez_compressed(os.path.join(GOLD, "engine_tiny.npz"), motion=out.numpy())
    report["engine_tiny"] = {"maxabs": float((o - out).abs().max()), "shape": list(out.shape)}
    print("engine_tiny", report["engine_tiny"])
    # mesh branch: get_flame_verts through the engine's own FLAME model (inference.py:62-69)
    shape_code = torch.zeros(1, 300).expand(out.shape[0], -1)
    verts = eng.ARTalk.basic_vae.get_flame_verts(eng.flame_model, shape_code, out, with_global=True)
    ov = get_flame_verts(asset, shape_code, out, with_global=True)
    report["engine_tiny"]["verts_maxabs"] = float((ov - verts).abs().max())
    np.savez_compressed(os.path.join(GOLD, "engine_tiny_verts.npz"), verts=verts[:3].numpy())


def run_engine_case_named(report, name, clip_length):
    case = CASES[name]
    cfg = case.cfg
    sd = synthetic.make_state_dict(cfg, case.weight_seed)
    asset = synthetic.make_flame_asset(0)
    eng = live.load_engine(cfg, sd, asset, clip_length=clip_length)
    eng.set_style_motion(case.style()[0])
    audio = case.audio()[0]
    out = eng.inference(audio)
    orc = Oracle(sd, cfg)
    with torch.no_grad():
        o = orc.engine_inference(audio, case.style()[0:1], clip_length=clip_length)
    np.savez_compressed(os.path.join(GOLD, "engine_%s.npz" % name), motion=out.numpy())
    report["engine_" + name] = {"maxabs": float((o - out).abs().max()), "shape": list(out.shape)}
    print("engine_" + name, report["engine_" + name])


def run_flame(report):
    asset = synthetic.make_flame_asset(0)
    shape, motion = flame_inputs()
    out = {}
    dev = {}
    from types import SimpleNamespace
    for scale in (1.0, 5.0):
        fm = live.load_flame(asset, scale=scale)
        sys.path.insert(0, live.REF_ROOT)
        from app.modules.bitwise_vae import BITWISE_VAE
        for wg in (False, True):
            v = BITWISE_VAE.get_flame_verts(SimpleNamespace(), fm, shape, motion, with_global=wg)
            o = get_flame_verts(asset, shape, motion, with_global=wg, scale=scale)
            key = "verts_scale%g_global%d" % (scale, int(wg))
            out[key] = v.numpy()
            dev[key + "_maxabs"] = float((o - v).abs().max())
    np.savez_compressed(os.path.join(GOLD, "flame.npz"), **out)
    report["flame"] = dev
    print("flame", dev)


def run_gaga(report):
    """``GAGAvatar.build_forward_batch`` (app/GAGAvatar/models.py:98-128) called frame by frame like the render loop
    (inference.py:78-84) on a stand-in ``self`` holding a synthetic tracked avatar: FLAME(scale 5) + forehead EMA. The
    rasteriser extension and pytorch3d are stubbed (never executed on this leg); ``transform_emoca_to_p3d`` (camera matrix from
    the global rotation, pytorch3d) is replaced because it does not touch ``t_points``."""
    import types
    from types import SimpleNamespace
    live._stub_modules()
    if live.REF_ROOT not in sys.path:
        sys.path.insert(0, live.REF_ROOT)

    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Any()
        def __getattr__(self, n): return _Any()
    m = types.ModuleType("diff_gaussian_rasterization_32d")
    m.GaussianRasterizationSettings = _Any
    m.GaussianRasterizer = _Any
    sys.modules.setdefault("diff_gaussian_rasterization_32d", m)
    import app.GAGAvatar.models as gm
    asset = synthetic.make_flame_asset(0)
    fm = live.load_flame(asset, scale=5.0)
    gm.transform_emoca_to_p3d = lambda r: torch.eye(4)[None, :3].repeat(r.shape[0], 1, 1)
    motion, shape = gaga_inputs()
    me = SimpleNamespace(_tracked_id={"image": torch.zeros(3, 64, 64), "transform_matrix": torch.eye(4)[:3], "shapecode": shape[0]})
    pts = torch.stack([gm.GAGAvatar.build_forward_batch(me, motion[i:i + 1].clone(), fm)["t_points"][0] for i in range(motion.shape[0])])
    from oracle.artalk_oracle import gaga_t_points
    o = gaga_t_points(asset, shape, motion, gm.forehead_indices)
    idx = torch.as_tensor(gm.forehead_indices)
    np.savez_compressed(os.path.join(GOLD, "gaga.npz"), forehead=pts[:, idx].numpy(), strided=pts[:, ::8].numpy(),
                        forehead_indices=np.asarray(gm.forehead_indices, dtype=np.int32))
    report["gaga"] = {"t_points_maxabs": float((o - pts).abs().max()), "n_frames": int(motion.shape[0])}
    print("gaga", report["gaga"])


def main():
    if not live.available():
        raise SystemExit("reference not found at %s" % live.REF_ROOT)
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    report = {"torch": torch.__version__, "reference": live.REF_ROOT}
    only = sys.argv[1:]
    if not only or "eng1_audio" in only or not os.path.exists(os.path.join(GOLD, "eng1_audio.npz")):
        write_eng1_audio()
    for name, case in CASES.items():
        if only and name not in only:
            continue
        run_case(case, report)
    if not only or "engine" in only:
        run_engine_case(report)
    if not only or "engine_full_eng1" in only:
        run_engine_case(report, "full_eng1", 750)
    if not only or "flame" in only:
        run_flame(report)
    if not only or "gaga" in only:
        run_gaga(report)
    path = os.path.join(GOLD, "PIN_REPORT.json")
    old = {}
    if only and os.path.exists(path):
        old = json.load(open(path))
    old.update(report)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
