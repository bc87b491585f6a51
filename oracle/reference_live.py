"""ORACLE tooling — imports the *unmodified* reference from /root/reference (build container
only; the path does not exist on the GPU box) so that ``make_golden.py`` can pin
``artalk_oracle.Oracle`` against it and emit fixtures. Recipe from SURVEY.md §8c:

* ``sys.modules`` stubs for gradio / gtts / av / pytorch3d (imported at module top level by
  inference.py:10-11, app/utils_videos.py:4, app/flame_model/renderer_utils.py:8-20; never
  executed on the motion path);
* ``Wav2Vec2Config.from_pretrained`` (network call, app/models.py:25) replaced by the
  XLS-R-300m config built from our ``Wav2VecConfig``;
* ``torch.load`` shimmed for the two absent asset files.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
import tempfile
import types

import torch

REF_ROOT = os.environ.get("ARTALK_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "app", "models.py"))


def _stub_modules():
    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Any()
        def __getattr__(self, n): return _Any()

    def mk(name):
        m = types.ModuleType(name)
        def _ga(n):
            if n.startswith("__"):                 # inspect / importlib probe __file__, __spec__ ...: behave like a plain module
                raise AttributeError(n)
            return _Any
        m.__getattr__ = _ga                      # type: ignore[attr-defined]
        m.__path__ = []                          # behave like a package
        sys.modules.setdefault(name, m)
    for n in ("gradio", "gtts", "av", "pytorch3d", "pytorch3d.io", "pytorch3d.structures",
              "pytorch3d.renderer", "pytorch3d.transforms", "pytorch3d.renderer.implicit",
              "pytorch3d.renderer.implicit.harmonic_embedding", "pytorch3d.renderer.mesh",
              "pytorch3d.renderer.mesh.shader", "pytorch3d.renderer.cameras", "torchaudio"):
        if n == "torchaudio":
            try:
                import torchaudio  # noqa: F401
                continue
            except Exception:
                pass
        mk(n)


def hf_config(w):
    from transformers import Wav2Vec2Config
    return Wav2Vec2Config(
        hidden_size=w.hidden, num_hidden_layers=w.layers, num_attention_heads=w.heads,
        intermediate_size=w.ffn, hidden_act="gelu", feat_extract_norm="layer",
        feat_extract_activation="gelu", conv_dim=(w.conv_dim,) * len(w.conv_kernel),
        conv_stride=tuple(w.conv_stride), conv_kernel=tuple(w.conv_kernel), conv_bias=True,
        num_conv_pos_embeddings=w.pos_conv_kernel, num_conv_pos_embedding_groups=w.pos_conv_groups,
        do_stable_layer_norm=True, layer_norm_eps=w.ln_eps, mask_time_prob=0.075)


@contextlib.contextmanager
def _patched(cfg, state_dict, flame_asset):
    from transformers import Wav2Vec2Config
    orig_fp, orig_load = Wav2Vec2Config.from_pretrained, torch.load
    Wav2Vec2Config.from_pretrained = classmethod(lambda cls, *a, **k: hf_config(cfg.wav2vec))

    def load(f, *a, **k):
        name = os.path.basename(str(f))
        if name.startswith("ARTalk_") and name.endswith(".pt"):
            return state_dict
        if name == "FLAME_with_eye.pt":
            return flame_asset
        return orig_load(f, *a, **k)
    torch.load = load
    try:
        yield
    finally:
        Wav2Vec2Config.from_pretrained, torch.load = orig_fp, orig_load


def load_model(cfg, state_dict):
    """Live ``BitwiseARModel`` with ``state_dict`` strictly loaded (app/models.py:13-56)."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with _patched(cfg, state_dict, None):
        from app import BitwiseARModel
        j = cfg.to_reference_json()
        j["AR_CONFIG"]["AUDIO_ENCODER"] = "wav2vec"
        m = BitwiseARModel(j).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


def load_flame(flame_asset, scale=1.0):
    """Live ``FLAMEModel(n_shape=300, n_exp=100, no_lmks=True)`` (inference.py:29)."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _stub_modules()
    with _flame_patch(flame_asset):
        from app.flame_model import FLAMEModel
        return FLAMEModel(n_shape=300, n_exp=100, scale=scale, no_lmks=True)


@contextlib.contextmanager
def _flame_patch(flame_asset):
    import copy
    orig_load = torch.load

    def load(f, *a, **k):
        if os.path.basename(str(f)) == "FLAME_with_eye.pt":
            return copy.deepcopy(flame_asset)      # FLAME.py:41 mutates kintree_table in place
        return orig_load(f, *a, **k)
    torch.load = load
    try:
        yield
    finally:
        torch.load = orig_load


def load_engine(cfg, state_dict, flame_asset, **engine_kw):
    """Live ``ARTAvatarInferEngine(load_gaga=False, device='cpu')`` (inference.py:18-39), built in
    a temp cwd holding ``assets/config.json`` because the engine uses cwd-relative paths."""
    import copy
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _stub_modules()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="artalk_ref_")
    os.makedirs(os.path.join(tmp, "assets"))
    with open(os.path.join(tmp, "assets", "config.json"), "w") as f:
        json.dump(cfg.to_reference_json(), f)
    os.chdir(tmp)
    try:
        with _patched(cfg, state_dict, copy.deepcopy(flame_asset)):
            import inference as ref_inference
            eng = ref_inference.ARTAvatarInferEngine(load_gaga=False, device="cpu", **engine_kw)
    finally:
        os.chdir(cwd)
    return eng
