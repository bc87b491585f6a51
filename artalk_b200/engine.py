"""Drop-in for ``inference.ARTAvatarInferEngine`` on the audio->motion(->mesh vertices) path (inference.py:18-95).

Same constructor arguments, attributes (``ARTalk``, ``flame_model``, ``style_motion``, ``device``, ``fix_pose``,
``clip_length``, ``output_dir``), ``set_style_motion`` and ``inference`` semantics: cwd-relative ``./assets`` files,
shape assertion on the style clip, Savitzky-Golay smoothing, ``[:clip_length]``, pose / dims 104:106 zeroing.
Image rendering (pytorch3d mesh renderer, GAGAvatar, video mux) is outside the path: ``rendering`` raises and
``mesh_vertices`` returns what the mesh branch feeds to its renderer (inference.py:62-69).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch

from . import _lib
from .flame import FLAMEModel
from .model import BitwiseARModel
from .weights import savgol_hat

_savgol_ready = False


def _ensure_savgol_tables():
    global _savgol_ready
    if not _savgol_ready:
        h5 = savgol_hat(5, 2).reshape(-1).copy()
        h9 = savgol_hat(9, 3).reshape(-1).copy()
        _lib.check(_lib.lib().artalk_set_savgol_tables(h5.ctypes.data, h9.ctypes.data))
        _savgol_ready = True


def smooth_motion(motion: torch.Tensor, clip_length: Optional[int] = None, fix_pose: bool = False,
                  zero_tail: bool = True) -> torch.Tensor:
    """(B,T,106) or (T,106) device tensor -> savgol-smoothed, clipped, post-processed copy (inference.py:52-56,89-95)."""
    _ensure_savgol_tables()
    squeeze = motion.dim() == 2
    m = motion[None] if squeeze else motion
    m = m.to(torch.float32).contiguous()
    _lib.require_cuda(m.device)
    B, T, D = m.shape
    if T < 9:
        # scipy: "If mode is 'interp', window_length must be less than or equal to the size of x."
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
    T_out = T if clip_length is None else max(0, min(T, int(clip_length)))
    out = torch.empty(B, T_out, D, device=m.device)
    _lib.check(_lib.lib().artalk_smooth_motion(m.data_ptr(), out.data_ptr(), B, T, T_out, int(bool(fix_pose)),
                                               int(bool(zero_tail)), _lib.stream_ptr(m.device)))
    return out[0] if squeeze else out


class ARTAvatarInferEngine:
    def __init__(self, load_gaga=False, fix_pose=False, clip_length=750, device="cuda", *, precision="bf16",
                 state_dict=None, config=None, flame_asset=None, wav2vec=None, make_output_dir=True, lanes=1):
        if load_gaga:
            raise NotImplementedError("GAGAvatar rendering is outside the audio->motion path (use load_gaga=False)")
        self.device = device
        self.fix_pose = fix_pose
        self.clip_length = clip_length
        audio_encoder = "wav2vec"
        ckpt = state_dict if state_dict is not None else torch.load(
            "./assets/ARTalk_{}.pt".format(audio_encoder), map_location="cpu", weights_only=True)
        configs = config if config is not None else json.load(open("./assets/config.json"))
        configs = json.loads(json.dumps(configs))
        configs["AR_CONFIG"]["AUDIO_ENCODER"] = audio_encoder
        self.ARTalk = BitwiseARModel(configs, device=device, precision=precision, wav2vec=wav2vec, lanes=lanes).eval().to(device)
        self.ARTalk.load_state_dict(ckpt, strict=True)
        self.flame_model = FLAMEModel(n_shape=300, n_exp=100, scale=1.0, no_lmks=True, asset=flame_asset, device=device)
        self.mesh_renderer = None                      # pytorch3d RenderMesh: rendering is out of scope
        self.output_dir = "render_results/ARTAvatar_{}".format(audio_encoder)
        if make_output_dir:
            os.makedirs(self.output_dir, exist_ok=True)
        self.style_motion = None

    def set_style_motion(self, style_motion):
        if isinstance(style_motion, str):
            style_motion = torch.load("assets/style_motion/{}.pt".format(style_motion), map_location="cpu", weights_only=True)
        assert style_motion.shape == (50, 106), f"Invalid style_motion shape: {style_motion.shape}."
        self.style_motion = style_motion[None].to(self.device)

    def inference(self, audio, clip_length=None):
        """audio (S,) fp32 16 kHz mono -> (min(ceil(S/640), clip_length), 106) fp32 on ``device``."""
        audio_batch = {"audio": audio[None].to(self.device), "style_motion": self.style_motion}
        pred_motions = self.ARTalk.inference(audio_batch, with_gtmotion=False)[0]
        clip_length = clip_length if clip_length is not None else self.clip_length
        return smooth_motion(pred_motions, clip_length, self.fix_pose)

    def inference_batch(self, audio, style_motion=None, clip_length=None):
        """Batched form: audio (B,S), style_motion (B,50,106) or None -> (B, T, 106); equals the per-clip loop."""
        pred = self.ARTalk.inference({"audio": audio.to(self.device), "style_motion": style_motion}, with_gtmotion=False)
        clip_length = clip_length if clip_length is not None else self.clip_length
        return smooth_motion(pred, clip_length, self.fix_pose)

    def mesh_vertices(self, pred_motions, shape_code=None):
        """The vertices the mesh branch of ``rendering`` computes (inference.py:62-69): (N,106) -> (N,5023,3)."""
        if shape_code is None:
            shape_code = pred_motions.new_zeros(1, 300).to(self.device).expand(pred_motions.shape[0], -1)
        else:
            assert shape_code.dim() == 2, f"Invalid shape_code dim: {shape_code.dim()}."
            assert shape_code.shape[0] == 1, f"Invalid shape_code shape: {shape_code.shape}."
            shape_code = shape_code.to(self.device).expand(pred_motions.shape[0], -1)
        return self.ARTalk.basic_vae.get_flame_verts(self.flame_model, shape_code, pred_motions, with_global=True)

    def rendering(self, audio, pred_motions, shape_id="mesh", shape_code=None, save_name="ARTAvatar.mp4"):
        raise NotImplementedError("image rendering / video muxing is out of scope; use mesh_vertices() for the FLAME "
                                  "vertices the mesh branch renders")

    @staticmethod
    def smooth_motion_savgol(motion_codes):
        """inference.py:89-95 alone (no clipping / zeroing), on the device instead of the scipy host round trip."""
        return smooth_motion(motion_codes, None, False, zero_tail=False)
