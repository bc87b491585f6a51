"""ctypes binding of libartalk_b200.so (include/artalk_b200.h). There is no CPU or PyTorch fallback: if the
library is missing or a call fails, the host raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

from .config import ModelConfig

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARTALK_LIB") or os.path.join(HERE, "lib", "libartalk_b200.so")      # ARTALK_LIB: A/B builds

F32, BF16, I32 = 0, 1, 2
PRECISION = {"fp32": 0, "bf16": 1, "bf16x3": 2, "bf16x6": 3}


class ArtalkError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("precision", C.c_int),
        ("ar_depth", C.c_int), ("ar_heads", C.c_int), ("embed_dim", C.c_int), ("cond_dim", C.c_int),
        ("vae_depth", C.c_int), ("vae_heads", C.c_int), ("vae_hidden", C.c_int), ("code_dim", C.c_int),
        ("motion_dim", C.c_int),
        ("n_levels", C.c_int), ("patch_nums", C.c_int * 8),
        ("w2v_layers", C.c_int), ("w2v_heads", C.c_int), ("w2v_hidden", C.c_int), ("w2v_ffn", C.c_int),
        ("w2v_conv_dim", C.c_int), ("w2v_n_conv", C.c_int),
        ("w2v_conv_kernel", C.c_int * 8), ("w2v_conv_stride", C.c_int * 8),
        ("w2v_pos_kernel", C.c_int), ("w2v_pos_groups", C.c_int),
        ("style_dim", C.c_int), ("style_layers", C.c_int), ("style_heads", C.c_int), ("style_ffn", C.c_int),
        ("style_len", C.c_int),
        ("chunk_samples", C.c_int),
        ("w2v_ln_eps", C.c_float),
    ]


class RowMap(C.Structure):
    _fields_ = [("rpb", C.c_int), ("bs", C.c_int64), ("rs", C.c_int64)]


class Gemm(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_map", RowMap), ("W", C.c_void_p), ("ldw", C.c_int64),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("tap_w", C.c_int), ("tap_pad", C.c_int),
        ("groups", C.c_int), ("a_gs", C.c_int64), ("w_gs", C.c_int64), ("c_gs", C.c_int64), ("bias_gs", C.c_int),
        ("bias", C.c_void_p), ("act", C.c_int),
        ("gate", C.c_void_p), ("gate_dt", C.c_int), ("gate_map", RowMap),
        ("resid", C.c_void_p), ("resid_map", RowMap),
        ("out32", C.c_void_p), ("out_act", C.c_void_p), ("out_act_dt", C.c_int), ("c_map", RowMap),
        ("tap_slots", C.c_int), ("exact", C.c_int), ("split_acc", C.c_int),
    ]


class Attn(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("dt", C.c_int), ("n_seq", C.c_int), ("n_heads", C.c_int), ("head_dim", C.c_int), ("lq", C.c_int), ("lk", C.c_int),
        ("q_ss", C.c_int64), ("q_rs", C.c_int64), ("k_ss", C.c_int64), ("k_rs", C.c_int64),
        ("v_ss", C.c_int64), ("v_rs", C.c_int64), ("o_ss", C.c_int64), ("o_rs", C.c_int64),
        ("scale", C.c_float), ("split", C.c_int), ("key_bound", C.c_void_p),
    ]


class FlameModelC(C.Structure):
    _fields_ = [
        ("n_verts", C.c_int), ("n_shape", C.c_int), ("n_exp", C.c_int),
        ("v_template", C.c_void_p), ("dirs", C.c_void_p), ("j_template", C.c_void_p), ("j_dirs", C.c_void_p),
        ("lbs_weights", C.c_void_p), ("parents", C.c_int * 5), ("scale", C.c_float),
        ("bsplit_full", C.c_void_p), ("ks_full", C.c_int), ("bsplit_expr", C.c_void_p), ("ks_expr", C.c_int),
    ]


#: every symbol include/artalk_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "artalk_last_error": (C.c_char_p, []),
    "artalk_abi_version": (C.c_int, []),
    "artalk_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "artalk_destroy": (C.c_int, [C.c_void_p]),
    "artalk_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64]),
    "artalk_finalize": (C.c_int, [C.c_void_p]),
    "artalk_set_workspace_limit": (C.c_int, [C.c_void_p, C.c_size_t]),
    "artalk_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "artalk_enable_graphs": (C.c_int, [C.c_void_p, C.c_int]),
    "artalk_set_latency_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "artalk_graph_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "artalk_audio_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "artalk_style_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "artalk_motion_to_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "artalk_bits_to_motion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "artalk_ar_chunk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "artalk_flame_workspace_floats": (C.c_size_t, [C.POINTER(FlameModelC), C.c_int]),
    "artalk_flame_vertices": (C.c_int, [C.POINTER(FlameModelC), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "artalk_set_savgol_tables": (C.c_int, [C.c_void_p, C.c_void_p]),
    "artalk_smooth_motion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "artalk_resample_mono": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int64, C.c_void_p]),
    "artalk_ema_scan": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "artalk_vertex_normals": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "artalk_launch_count": (C.c_ulonglong, []),
    "artalk_enable_pdl": (C.c_int, [C.c_int]),
    "artalk_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "artalk_trace_begin": (C.c_int, [C.c_void_p]),
    "artalk_trace_end": (C.c_long, [C.c_char_p, C.c_long, C.c_void_p]),
    "artalk_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "artalk_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "artalk_op_gemm": (C.c_int, [C.POINTER(Gemm), C.c_int, C.c_void_p]),
    "artalk_op_split_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "artalk_op_attention": (C.c_int, [C.POINTER(Attn), C.c_void_p]),
    "artalk_op_attention_split": (C.c_int, [C.POINTER(Attn), C.c_void_p, C.c_size_t, C.c_void_p]),
    "artalk_op_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_float, C.c_int, C.c_void_p]),
    "artalk_op_conv0": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_float, C.c_void_p]),
    "artalk_op_posconv4": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (once). Raises ArtalkError if it was not built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ArtalkError("%s not found: build it with `python -m artalk_b200.build` "
                              "(there is no CPU/PyTorch fallback for this path)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)         # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        if os.environ.get("ARTALK_PDL", "1") == "0":      # developer switch: plain stream-order launches
            l.artalk_enable_pdl(0)
        if os.environ.get("ARTALK_PDL_MASK"):
            l.artalk_set_option(b"pdl_mask", int(os.environ["ARTALK_PDL_MASK"]))
        if os.environ.get("ARTALK_PDL_W2V_MAX_CHUNKS"):
            l.artalk_set_option(b"pdl_w2v_max_chunks", int(os.environ["ARTALK_PDL_W2V_MAX_CHUNKS"]))
        if os.environ.get("ARTALK_GEMM_TMA_RESID"):
            l.artalk_set_option(b"gemm_tma_resid", int(os.environ["ARTALK_GEMM_TMA_RESID"]))
        if os.environ.get("ARTALK_GEMM_RESID_DEEP"):
            l.artalk_set_option(b"gemm_resid_deep", int(os.environ["ARTALK_GEMM_RESID_DEEP"]))
        if os.environ.get("ARTALK_GEMM_TMA_OUT"):
            l.artalk_set_option(b"gemm_tma_out", int(os.environ["ARTALK_GEMM_TMA_OUT"]))
        if os.environ.get("ARTALK_GEMM_BAND_MB"):
            l.artalk_set_option(b"gemm_band_mb", int(os.environ["ARTALK_GEMM_BAND_MB"]))
        if os.environ.get("ARTALK_GEMM_FORCE_BN"):
            l.artalk_set_option(b"gemm_force_bn", int(os.environ["ARTALK_GEMM_FORCE_BN"]))
        if os.environ.get("ARTALK_ATTN_SIMT_MAX_LQ"):
            l.artalk_set_option(b"attn_simt_max_lq", int(os.environ["ARTALK_ATTN_SIMT_MAX_LQ"]))
        if os.environ.get("ARTALK_SKINNY_MAX_M"):         # 0: the latency-path kernels (skinny.cu) are never taken
            l.artalk_set_option(b"skinny_max_m", int(os.environ["ARTALK_SKINNY_MAX_M"]))
        if os.environ.get("ARTALK_ATTN_BOUND"):            # 0: AR attention keeps the max pass (A/B switch)
            l.artalk_set_option(b"attn_bound", int(os.environ["ARTALK_ATTN_BOUND"]))
        if os.environ.get("ARTALK_CONV0_FOLD"):            # 0: conv layer 0 in its direct form (A/B switch)
            l.artalk_set_option(b"conv0_fold", int(os.environ["ARTALK_CONV0_FOLD"]))
        if os.environ.get("ARTALK_POSCONV4"):              # 0: positional conv as the N = 64 tap-mode GEMM (A/B switch)
            l.artalk_set_option(b"posconv4", int(os.environ["ARTALK_POSCONV4"]))
        if os.environ.get("ARTALK_ATTN_BLK"):              # 0: 257..384-key launches take the split-key mode of attn_tc_kernel (A/B switch)
            l.artalk_set_option(b"attn_blk", int(os.environ["ARTALK_ATTN_BLK"]))
        if os.environ.get("ARTALK_SKINNY_TOKENS"):
            l.artalk_set_option(b"skinny_tokens", int(os.environ["ARTALK_SKINNY_TOKENS"]))
        if os.environ.get("ARTALK_GEMM_PAIR", "1") == "0":
            l.artalk_set_option(b"gemm_pair", 0)
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().artalk_last_error()
        raise ArtalkError("artalk_b200 call failed (code %d): %s" % (status, msg.decode() if msg else "?"))


def call(device, fn, *args) -> None:
    """One C-ABI call with ``device`` current: the library allocates and launches on the current device, so an engine on
    ``cuda:1`` must not run with device 0 current (a stream handle of another device is an invalid resource handle)."""
    with torch.cuda.device(device):
        check(fn(*args))


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(device) -> torch.device:
    d = torch.device(device)
    if d.type != "cuda":
        raise ArtalkError("artalk_b200 runs on CUDA devices only (got %r); the product has no CPU path" % (device,))
    if not torch.cuda.is_available():
        raise ArtalkError("CUDA is not available: artalk_b200 has no CPU fallback")
    if d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def make_config(cfg: ModelConfig, precision: str) -> Config:
    cfg.validate()
    w = cfg.wav2vec
    c = Config()
    c.precision = PRECISION[precision]
    c.ar_depth, c.ar_heads, c.embed_dim, c.cond_dim = cfg.ar_depth, cfg.ar_heads, cfg.embed_dim, cfg.cond_dim
    c.vae_depth, c.vae_heads, c.vae_hidden = cfg.vae_depth, cfg.vae_heads, cfg.vae_hidden
    c.code_dim, c.motion_dim = cfg.code_dim, cfg.motion_dim
    c.n_levels = len(cfg.patch_nums)
    for i, p in enumerate(cfg.patch_nums):
        c.patch_nums[i] = p
    c.w2v_layers, c.w2v_heads, c.w2v_hidden, c.w2v_ffn = w.layers, w.heads, w.hidden, w.ffn
    c.w2v_conv_dim, c.w2v_n_conv = w.conv_dim, len(w.conv_kernel)
    for i, (k, s) in enumerate(zip(w.conv_kernel, w.conv_stride)):
        c.w2v_conv_kernel[i], c.w2v_conv_stride[i] = k, s
    c.w2v_pos_kernel, c.w2v_pos_groups = w.pos_conv_kernel, w.pos_conv_groups
    c.style_dim, c.style_layers, c.style_heads, c.style_ffn = cfg.style_dim, cfg.style_layers, cfg.style_heads, cfg.style_ffn
    c.style_len = cfg.style_len
    c.chunk_samples = cfg.chunk_samples
    c.w2v_ln_eps = w.ln_eps
    return c
