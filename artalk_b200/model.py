"""Drop-in for ``app.models.BitwiseARModel`` on the inference path (app/models.py:13-135) and for the parts of
``BITWISE_VAE`` the engine touches (app/modules/bitwise_vae.py:43-57,78-113). Python only sequences chunks; all
arithmetic runs in libartalk_b200.so. Differences from the reference that are part of the design:

* batches of clips are accepted (the reference asserts batch 1, app/models.py:65); batched == per-clip loop;
* wav2vec2 runs for every chunk of every clip up front; the scale loop uses a KV cache, hoisted AdaLN and
  once-per-chunk previous-chunk K/V — outputs equal the un-cached reference schedule;
* ``precision``: "bf16" (tcgen05, default), "fp32" (CUDA cores), "bf16x3" / "bf16x6" (tcgen05 at fp32 grade: every GEMM
  operand split into 2 / 3 bf16 pieces, fp32 data flow; the tensor-core mode that meets the reference's bit-exact contract).

Every C-ABI call is made with the model's device current (``torch.cuda.device``), so an engine on ``cuda:1`` works while the
process's current device is 0.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from .config import ModelConfig, Wav2VecConfig
from .weights import repack


def unpack_words(words: torch.Tensor) -> torch.Tensor:
    """(...,) int32 words -> (..., 32) int32 bits (reference layout of ``bit_indices``)."""
    sh = torch.arange(32, device=words.device, dtype=torch.int64)
    return ((words.to(torch.int64)[..., None] >> sh) & 1).to(torch.int32)


def pack_words(bits: torch.Tensor) -> torch.Tensor:
    sh = torch.arange(32, device=bits.device, dtype=torch.int64)
    w = (bits.to(torch.int64) << sh).sum(dim=-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)


class BitwiseVAE:
    """The ``basic_vae`` attribute: bit <-> motion legs and the FLAME adaptor."""

    def __init__(self, owner: "BitwiseARModel"):
        self._o = owner
        self.motion_dim = owner.cfg.motion_dim
        self.code_dim = owner.cfg.code_dim
        self.patch_nums = list(owner.cfg.patch_nums)

    def get_flame_verts(self, flame_model, shape_params, motion_params, with_global=False):
        # app/modules/bitwise_vae.py:43-57
        exp_code, pose_code = motion_params[..., :100], motion_params[..., 100:]
        if not with_global:
            pose_code = torch.cat([torch.zeros_like(pose_code[..., :3]), pose_code[..., 3:]], dim=-1)
        if shape_params.dim() == 2:
            return flame_model(shape_params=shape_params, expression_params=exp_code, pose_params=pose_code)
        if shape_params.dim() == 3:
            return torch.stack([flame_model(shape_params=shape_params[b], expression_params=exp_code[b],
                                            pose_params=pose_code[b]) for b in range(shape_params.shape[0])], dim=0)
        raise ValueError("Invalid shape of shape_params: {}".format(shape_params.shape))

    @torch.no_grad()
    def quant_to_vqidx(self, prev_motion, this_motion=None):
        """(B,100,106) -> ((B,181,32) int bits, None); the two-motion form is training-only and not on the path."""
        if this_motion is not None:
            raise NotImplementedError("quant_to_vqidx(prev, this) is only used in training")
        words = self._o.motion_to_words(prev_motion)
        return unpack_words(words), None

    @torch.no_grad()
    def vqidx_to_motion(self, prev_code_idx, this_code_idx):
        """Returns (None, new_half): the reference discards the prev half on this path (app/models.py:108)."""
        return None, self._o.words_to_motion(pack_words(prev_code_idx), pack_words(this_code_idx))


class BitwiseARModel:
    def __init__(self, model_cfg=None, *, device="cuda", precision="bf16", wav2vec: Optional[Wav2VecConfig] = None,
                 max_clips: int = 256, **kwargs):
        if isinstance(model_cfg, ModelConfig):
            self.cfg = model_cfg
        else:
            self.cfg = ModelConfig.from_reference_json(model_cfg, wav2vec=wav2vec)
        self.cfg.validate()
        if precision not in _lib.PRECISION:
            raise ValueError("precision must be one of %s" % sorted(_lib.PRECISION))
        self.precision = precision
        self._device = torch.device(device)
        self.max_clips = int(max_clips)
        self.patch_nums = list(self.cfg.patch_nums)
        self.attn_depth = self.cfg.ar_depth
        self.prev_ratio = self.cfg.prev_ratio
        self.audio_feature_dim = self.cfg.cond_dim
        self.basic_vae = BitwiseVAE(self)
        self._h: Optional[C.c_void_p] = None
        self._tensors: Dict[str, torch.Tensor] = {}
        self._init_words = None
        self.latency_rows = 0
        self._copy_stream = None
        self._graph_warned = False

    # ---- nn.Module look-alikes (inference.py:27-28) -------------------------------------------------
    def eval(self):
        return self

    def to(self, device):
        if self._h is not None and torch.device(device) != self._device:
            raise _lib.ArtalkError("weights already live on %s" % self._device)
        self._device = torch.device(device)
        return self

    @property
    def device(self):
        return self._device

    def _call(self, fn, *args):
        """One C-ABI call with this model's device current (the library allocates / launches on the current device)."""
        with torch.cuda.device(self._device):
            _lib.check(fn(*args))

    def load_state_dict(self, state_dict, strict=True):
        if not strict:
            raise ValueError("only strict loading is supported (inference.py:28)")
        dev = _lib.require_cuda(self._device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device = dev
        lib = _lib.lib()
        self.close()
        with torch.cuda.device(dev):
            self._tensors = repack(state_dict, self.cfg, dev, self.precision)
            ccfg = _lib.make_config(self.cfg, self.precision)
            code = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.int32: _lib.I32}
            h = C.c_void_p()
            _lib.check(lib.artalk_create(C.byref(ccfg), C.byref(h)))
            self._h = h
            for name, t in self._tensors.items():
                _lib.check(lib.artalk_set_tensor(h, name.encode(), t.data_ptr(), code[t.dtype], t.numel()))
            _lib.check(lib.artalk_finalize(h))
            if self.latency_rows:
                _lib.check(lib.artalk_set_latency_mode(h, self.latency_rows))
            torch.cuda.synchronize(dev)
        self._init_words = None
        return self

    def close(self):
        if self._h is not None:
            with torch.cuda.device(self._device):
                torch.cuda.synchronize(self._device)
                _lib.lib().artalk_destroy(self._h)
        self._h, self._tensors = None, {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_workspace_limit(self, n_bytes: int):
        self._call(_lib.lib().artalk_set_workspace_limit, self._handle(), n_bytes)

    def set_latency_mode(self, on=True, max_rows: int = 128):
        """Batch-1 / few-clip streaming: GEMMs with at most ``max_rows`` rows take the latency kernel (include/artalk_b200.h,
        artalk_set_latency_mode). Off (default) = throughput mode, whose arithmetic does not depend on the batch size."""
        self.latency_rows = int(max_rows) if on else 0
        if self._h is not None:
            self._call(_lib.lib().artalk_set_latency_mode, self._handle(), self.latency_rows)

    def enable_graphs(self, on: bool = True):
        self._call(_lib.lib().artalk_enable_graphs, self._handle(), int(on))

    def graph_status(self):
        """(graphs instantiated, replays so far, failure text or None): a failed capture makes the engine launch eagerly."""
        n, r = C.c_int(0), C.c_int(0)
        rc = _lib.lib().artalk_graph_status(self._handle(), C.byref(n), C.byref(r))
        msg = None
        if rc != 0:
            m = _lib.lib().artalk_last_error()
            msg = m.decode() if m else "unknown"
        return n.value, r.value, msg

    def _handle(self):
        if self._h is None:
            raise _lib.ArtalkError("no weights loaded: call load_state_dict first")
        return self._h

    # ---- stage-level calls (each is one C-ABI call on the current stream) ------------------------------
    def _dev(self, t, dtype=torch.float32):
        pinned = t.device.type == "cpu" and t.is_pinned()
        return t.to(self._device, dtype, non_blocking=pinned).contiguous()

    def audio_cond(self, chunks: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(N, 64000) audio chunks -> (N, 181, 1024) conditioning (wav2vec2 + area pooling); ``out``: optional contiguous
        (N, 181, 1024) fp32 device tensor to write into."""
        x = self._dev(chunks)
        if x.dim() != 2 or x.shape[1] != self.cfg.chunk_samples:
            raise ValueError("expected (N, %d) chunks, got %s" % (self.cfg.chunk_samples, tuple(x.shape)))
        cond = out if out is not None else torch.empty(x.shape[0], self.cfg.seq_tokens, self.cfg.cond_dim, device=self._device)
        assert cond.is_contiguous() and cond.dtype == torch.float32 and cond.numel() == x.shape[0] * self.cfg.seq_tokens * self.cfg.cond_dim
        self._call(_lib.lib().artalk_audio_encode, self._handle(), x.data_ptr(), x.shape[0], cond.data_ptr(),
                   _lib.stream_ptr(self._device))
        return cond

    def style_cond(self, style_motion: Optional[torch.Tensor], batch: int) -> torch.Tensor:
        """(B,50,106) or None -> (B,768) style token (app/models.py:67-73)."""
        if style_motion is None:
            return self._tensors["style.null"][None].expand(batch, -1).contiguous()
        s = self._dev(style_motion)
        if s.dim() != 3 or s.shape[0] != batch or s.shape[1] != self.cfg.style_len or s.shape[2] != self.cfg.motion_dim:
            raise ValueError("style_motion must be (%d, %d, %d), got %s" % (batch, self.cfg.style_len, self.cfg.motion_dim, tuple(s.shape)))
        out = torch.empty(batch, self.cfg.embed_dim, device=self._device)
        self._call(_lib.lib().artalk_style_encode, self._handle(), s.data_ptr(), batch, out.data_ptr(), _lib.stream_ptr(self._device))
        return out

    def motion_to_words(self, motion: torch.Tensor, enc_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        m = self._dev(motion)
        words = torch.empty(m.shape[0], self.cfg.seq_tokens, dtype=torch.int32, device=self._device)
        self._call(_lib.lib().artalk_motion_to_bits, self._handle(), m.data_ptr(), m.shape[0], words.data_ptr(), _lib.ptr(enc_out),
                   _lib.stream_ptr(self._device))
        return words

    def words_to_motion(self, prev_words: torch.Tensor, words: torch.Tensor) -> torch.Tensor:
        pw, w = self._dev(prev_words, torch.int32), self._dev(words, torch.int32)
        out = torch.empty(w.shape[0], self.cfg.chunk_frames, self.cfg.motion_dim, device=self._device)
        self._call(_lib.lib().artalk_bits_to_motion, self._handle(), pw.data_ptr(), w.data_ptr(), w.shape[0], out.data_ptr(),
                   _lib.stream_ptr(self._device))
        return out

    def initial_words(self, batch: int) -> torch.Tensor:
        """Bits of the all-zero previous motion (app/models.py:86-87); input independent, cached per weight set."""
        if self._init_words is None:
            z = torch.zeros(1, self.cfg.chunk_frames, self.cfg.motion_dim, device=self._device)
            self._init_words = self.motion_to_words(z)
            torch.cuda.current_stream(self._device).synchronize()
        return self._init_words.repeat(batch, 1)          # fresh copy: ar_chunk updates prev_words in place

    def ar_chunk(self, cond: torch.Tensor, style: torch.Tensor, prev_words: torch.Tensor, motion_out: torch.Tensor,
                 words_out=None, logits_out=None, forced_words=None, enc_out=None):
        """One chunk for all clips; ``cond`` is (B,181,1024) possibly a strided view over clips; prev_words updated in place."""
        B = style.shape[0]
        assert cond.stride(2) == 1 and cond.stride(1) == self.cfg.cond_dim
        self._call(_lib.lib().artalk_ar_chunk,
                   self._handle(), B, cond.data_ptr(), cond.stride(0), style.data_ptr(), prev_words.data_ptr(), motion_out.data_ptr(),
                   _lib.ptr(words_out), _lib.ptr(logits_out), _lib.ptr(forced_words), _lib.ptr(enc_out), _lib.stream_ptr(self._device))

    # ---- app/models.py:62-121 -------------------------------------------------------------------------
    def _run_clips(self, audio, style_motion, motion, b0, b1, trace, teacher_words, teacher_prev_words, audio_ready=None):
        """Full pipeline for clips [b0, b1) on the current stream; writes motion[b0:b1]."""
        cfg = self.cfg
        T, n_chunks = cfg.chunk_frames, motion.shape[1]
        nb = b1 - b0
        style = self.style_cond(None if style_motion is None else style_motion[b0:b1], nb)
        if isinstance(audio_ready, list):
            # host input uploaded in clip groups on the copy stream (see inference): wav2vec runs group by group as the
            # uploads land, so only the first group's transfer is exposed
            assert b0 == 0 and b1 == audio.shape[0]
            cond = torch.empty(nb, n_chunks, cfg.seq_tokens, cfg.cond_dim, device=self._device)
            for g0, g1, ev in audio_ready:
                torch.cuda.current_stream(self._device).wait_event(ev)
                self.audio_cond(audio[g0:g1].reshape((g1 - g0) * n_chunks, cfg.chunk_samples), out=cond[g0:g1])
        else:
            if audio_ready is not None:                   # audio upload in flight on the copy stream (see inference)
                torch.cuda.current_stream(self._device).wait_event(audio_ready)
            cond = self.audio_cond(audio[b0:b1].reshape(nb * n_chunks, cfg.chunk_samples)).view(
                nb, n_chunks, cfg.seq_tokens, cfg.cond_dim)
        if trace is not None:
            trace["cond"][b0:b1].copy_(cond); trace["style"][b0:b1].copy_(style)
        for g0 in range(0, nb, self.max_clips):
            g1 = min(nb, g0 + self.max_clips)
            ng = g1 - g0
            prev_words = self.initial_words(ng)
            chunk_out = torch.empty(ng, T, cfg.motion_dim, device=self._device)
            tr = None
            if trace is not None:
                tr = dict(logits=torch.empty(ng, cfg.seq_tokens, 2 * cfg.code_dim, device=self._device),
                          words=torch.empty(ng, cfg.seq_tokens, dtype=torch.int32, device=self._device),
                          enc=torch.empty(ng, T, cfg.code_dim, device=self._device))
            for c in range(n_chunks):
                forced = None
                if teacher_words is not None:
                    forced = teacher_words[b0 + g0:b0 + g1, c].to(self._device, torch.int32).contiguous()
                self.ar_chunk(cond[g0:g1, c], style[g0:g1], prev_words, chunk_out,
                              words_out=None if tr is None else tr["words"], logits_out=None if tr is None else tr["logits"],
                              forced_words=forced, enc_out=None if tr is None else tr["enc"])
                motion[b0 + g0:b0 + g1, c].copy_(chunk_out)
                if tr is not None:
                    s = slice(b0 + g0, b0 + g1)
                    trace["logits"][s, c].copy_(tr["logits"]); trace["words"][s, c].copy_(tr["words"])
                    trace["prev_words"][s, c].copy_(prev_words); trace["enc_out"][s, c].copy_(tr["enc"])
                if teacher_prev_words is not None:
                    prev_words.copy_(teacher_prev_words[b0 + g0:b0 + g1, c].to(self._device, torch.int32))

    @torch.no_grad()
    def inference(self, batch, with_gtmotion=False, trace: Optional[dict] = None, teacher_words: Optional[torch.Tensor] = None,
                  teacher_prev_words: Optional[torch.Tensor] = None):
        """``trace`` (dict) collects per-chunk intermediates; ``teacher_words`` / ``teacher_prev_words`` (B, n_chunks, 181)
        int32 force the AR bits fed to the next scale + decoder and the re-encoded bits carried to the next chunk
        (parity tests: one flipped bit must not cascade through the recurrence)."""
        cfg = self.cfg
        audio = batch["audio"]
        if audio.dim() != 2:
            raise ValueError("batch['audio'] must be (B, S)")
        B, S = audio.shape
        style_motion = batch.get("style_motion", None)
        seq_length = cfg.frames_for_samples(S)
        T = cfg.chunk_frames
        n_chunks = math.ceil(seq_length / T)
        if n_chunks == 0:
            return torch.zeros(B, 0, cfg.motion_dim, device=self._device)
        self._handle()
        pad = n_chunks * cfg.chunk_samples - S
        audio_ready = None
        if audio.device.type == "cpu" and audio.is_pinned():
            # host input: the audio upload runs on a copy stream while the style encoder (which needs only the 21 KB style
            # clips) runs on the caller's stream; wav2vec waits for the event
            main = torch.cuda.current_stream(self._device)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self._device)
            self._copy_stream.wait_stream(main)
            with torch.cuda.stream(self._copy_stream):
                if audio.dtype == torch.float32 and B * n_chunks > 768 and os.environ.get("ARTALK_UPLOAD_GROUPS", "1") != "0":
                    # large batches: upload in groups of ~512 chunks, one event per group; wav2vec (which is sub-batched at
                    # about that size anyway) starts on group 0 while the later groups are still crossing PCIe
                    gsz = max(1, 512 // n_chunks)
                    host = audio
                    audio = torch.empty(B, n_chunks * cfg.chunk_samples, device=self._device, dtype=torch.float32)
                    if pad:
                        audio[:, S:].zero_()
                    audio_ready = []
                    for g0 in range(0, B, gsz):
                        g1 = min(B, g0 + gsz)
                        audio[g0:g1, :S].copy_(host[g0:g1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self._copy_stream)
                        audio_ready.append((g0, g1, ev))
                else:
                    audio = audio.to(self._device, torch.float32, non_blocking=True).contiguous()
                    if pad:
                        audio = torch.cat([audio, audio.new_zeros(B, pad)], dim=-1)
                    audio_ready = torch.cuda.Event()
                    audio_ready.record(self._copy_stream)
            audio.record_stream(main)
        else:
            audio = self._dev(audio)
            if pad:
                audio = torch.cat([audio, audio.new_zeros(B, pad)], dim=-1)
        if style_motion is not None:
            style_motion = self._dev(style_motion)
        motion = torch.empty(B, n_chunks, T, cfg.motion_dim, device=self._device)
        if trace is not None:
            trace.update(cond=torch.empty(B, n_chunks, cfg.seq_tokens, cfg.cond_dim, device=self._device),
                         style=torch.empty(B, cfg.embed_dim, device=self._device),
                         logits=torch.empty(B, n_chunks, cfg.seq_tokens, 2 * cfg.code_dim, device=self._device),
                         words=torch.empty(B, n_chunks, cfg.seq_tokens, dtype=torch.int32, device=self._device),
                         prev_words=torch.empty(B, n_chunks, cfg.seq_tokens, dtype=torch.int32, device=self._device),
                         enc_out=torch.empty(B, n_chunks, T, cfg.code_dim, device=self._device))
        with torch.cuda.device(self._device):
            self._run_clips(audio, style_motion, motion, 0, B, trace, teacher_words, teacher_prev_words, audio_ready)
        if not self._graph_warned:
            _, _, failure = self.graph_status()
            if failure:
                import warnings
                warnings.warn("artalk_b200: %s — the chunk body is launched eagerly (slower)" % failure, RuntimeWarning)
                self._graph_warned = True
        pred_motions = motion.view(B, n_chunks * T, cfg.motion_dim)[:, :seq_length]
        if with_gtmotion:
            min_length = min(batch["motion"].shape[1], pred_motions.shape[1])
            shape_code = batch["shape"].expand(-1, min_length, -1)
            return pred_motions[:, :min_length], batch["motion"][:, :min_length], shape_code
        return pred_motions
