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
    _lib.call(m.device, _lib.lib().artalk_smooth_motion, m.data_ptr(), out.data_ptr(), B, T, T_out, int(bool(fix_pose)),
              int(bool(zero_tail)), _lib.stream_ptr(m.device))
    return out[0] if squeeze else out


class ARTAvatarInferEngine:
    def __init__(self, load_gaga=False, fix_pose=False, clip_length=750, device="cuda", *, precision="bf16",
                 state_dict=None, config=None, flame_asset=None, wav2vec=None, make_output_dir=True, latency_mode=False):
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
        self.ARTalk = BitwiseARModel(configs, device=device, precision=precision, wav2vec=wav2vec).eval().to(device)
        self.ARTalk.load_state_dict(ckpt, strict=True)
        if latency_mode:                               # batch-1 / few-clip streaming: see BitwiseARModel.set_latency_mode
            self.ARTalk.set_latency_mode(True)
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
        if not (audio.device.type == "cpu" and audio.is_pinned()):      # pinned host audio is uploaded by the model itself,
            audio = audio.to(self.device)                               # overlapped with the style encoder
        pred = self.ARTalk.inference({"audio": audio, "style_motion": style_motion}, with_gtmotion=False)
        clip_length = clip_length if clip_length is not None else self.clip_length
        return smooth_motion(pred, clip_length, self.fix_pose)

    def inference_file(self, audio_path, clip_length=None):
        """inference.py:230-235 for a WAV file: load, mix to mono, resample to 16 kHz on the device, run ``inference``."""
        from .audio import load_audio
        return self.inference(load_audio(audio_path, self.device), clip_length)

    def save_motions(self, pred_motions, save_name):
        """inference.py:124: ``torch.save(pred_motions.float().cpu(), <output_dir>/<save_name>_motions.pt)``; returns the path."""
        path = os.path.join(self.output_dir, "{}_motions.pt".format(save_name))
        os.makedirs(self.output_dir, exist_ok=True)
        torch.save(pred_motions.float().cpu(), path)
        return path

    @staticmethod
    def load_motions(path):
        """(T,106) fp32 motion file written by the reference (inference.py:124) or by ``save_motions``."""
        m = torch.load(path, map_location="cpu", weights_only=True)
        if m.dim() != 2 or m.shape[1] != 106:
            raise ValueError("expected a (T, 106) motion tensor, got {}".format(tuple(m.shape)))
        return m.float()

    def stream(self, batch=1):
        """Streaming ingestion: audio arrives in pieces, every completed 4 s / 100-frame chunk is encoded and
        decoded immediately with the AR state carried across chunks; see ``StreamingSession``."""
        return StreamingSession(self, batch)

    def mesh_vertices(self, pred_motions, shape_code=None, out=None):
        """The vertices the mesh branch of ``rendering`` computes (inference.py:62-69): (N,106) -> (N,5023,3). ``out``: optional
        preallocated (N,5023,3) fp32 device tensor (batch jobs decode 10^5 frames per call: 60 KB per frame)."""
        if shape_code is None:
            shape_code = pred_motions.new_zeros(1, 300).to(self.device).expand(pred_motions.shape[0], -1)
        else:
            assert shape_code.dim() == 2, f"Invalid shape_code dim: {shape_code.dim()}."
            assert shape_code.shape[0] == 1, f"Invalid shape_code shape: {shape_code.shape}."
            shape_code = shape_code.to(self.device).expand(pred_motions.shape[0], -1)
        if out is not None:                     # get_flame_verts(with_global=True) for a 2-D shape code, writing in place
            return self.flame_model(shape_params=shape_code, expression_params=pred_motions[..., :100],
                                    pose_params=pred_motions[..., 100:], out=out)
        return self.ARTalk.basic_vae.get_flame_verts(self.flame_model, shape_code, pred_motions, with_global=True)

    def rendering(self, audio, pred_motions, shape_id="mesh", shape_code=None, save_name="ARTAvatar.mp4"):
        raise NotImplementedError("image rendering / video muxing is out of scope; use mesh_vertices() for the FLAME "
                                  "vertices the mesh branch renders")

    @staticmethod
    def smooth_motion_savgol(motion_codes):
        """inference.py:89-95 alone (no clipping / zeroing), on the device instead of the scipy host round trip."""
        return smooth_motion(motion_codes, None, False, zero_tail=False)


class StreamingSession:
    """Chunk-at-a-time form of ``ARTAvatarInferEngine.inference`` (app/models.py:76-115 run incrementally):

        sess = engine.stream()
        for piece in microphone:                 # (S_piece,) or (B, S_piece) fp32 16 kHz
            for frames in sess.push(piece):      # (B, 100, 106) raw motion of every chunk completed by this piece
                ...
        tail = sess.flush()                      # last partial chunk (zero padded like app/models.py:79-80), trimmed to ceil(S/640)
        motion = sess.result()                   # == engine.inference(all audio): smoothed, clipped, post-processed

    The concatenated raw frames equal the whole-clip call bit for bit (same kernels on the same per-chunk operands); the
    Savitzky-Golay post-filter needs neighbours on both sides, so smoothed output is only available from ``result()``.
    """

    def __init__(self, engine: ARTAvatarInferEngine, batch: int = 1):
        self.engine, self.batch = engine, int(batch)
        m = engine.ARTalk
        self._cfg = m.cfg
        style = engine.style_motion
        if style is not None and style.shape[0] != self.batch:
            style = style.expand(self.batch, -1, -1)
        self._style = m.style_cond(style, self.batch)
        self._prev = m.initial_words(self.batch)
        self._buf = torch.empty(self.batch, 0, device=m.device)
        self._frames = []
        self._samples = 0
        self._closed = False

    def _run_chunk(self, chunk):
        m = self.engine.ARTalk
        cond = m.audio_cond(chunk)
        out = torch.empty(self.batch, self._cfg.chunk_frames, self._cfg.motion_dim, device=m.device)
        m.ar_chunk(cond, self._style, self._prev, out)
        self._frames.append(out)
        return out

    def push(self, audio):
        if self._closed:
            raise RuntimeError("stream already flushed")
        a = audio if audio.dim() == 2 else audio[None]
        if a.shape[0] != self.batch:
            raise ValueError("expected %d audio rows, got %d" % (self.batch, a.shape[0]))
        a = a.to(self._buf.device, torch.float32)
        self._samples += a.shape[1]
        self._buf = torch.cat([self._buf, a], dim=1)
        cs, done = self._cfg.chunk_samples, []
        while self._buf.shape[1] >= cs:
            done.append(self._run_chunk(self._buf[:, :cs].contiguous()))
            self._buf = self._buf[:, cs:]
        return done

    def flush(self):
        """Runs the zero-padded last chunk (if any audio is pending) and returns its useful frames (B, <=100, 106) or None."""
        if self._closed:
            return None
        self._closed = True
        total = self._cfg.frames_for_samples(self._samples)
        have = len(self._frames) * self._cfg.chunk_frames
        if self._buf.shape[1] == 0 or total <= have:
            return None
        pad = self._cfg.chunk_samples - self._buf.shape[1]
        chunk = torch.cat([self._buf, self._buf.new_zeros(self.batch, pad)], dim=1)
        out = self._run_chunk(chunk)
        return out[:, :total - have]

    def raw_motion(self):
        """(B, ceil(S/640), 106) un-smoothed motion of everything pushed so far (app/models.py:115)."""
        total = self._cfg.frames_for_samples(self._samples)
        if not self._frames:
            return torch.zeros(self.batch, 0, self._cfg.motion_dim, device=self._buf.device)
        return torch.cat(self._frames, dim=1)[:, :total]

    def result(self, clip_length=None):
        """Smoothed / clipped / post-processed motion of the whole stream, as ``engine.inference`` returns it."""
        self.flush()
        clip_length = clip_length if clip_length is not None else self.engine.clip_length
        out = smooth_motion(self.raw_motion(), clip_length, self.engine.fix_pose)
        return out[0] if self.batch == 1 else out
