"""Weight wire format of the path + seeded synthetic weights / assets.

The reference ships no checkpoints (``build_resources.sh:14-35`` downloads them;
FLAME is licence gated), so parity and benchmarks run on seeded random weights.
``state_dict_spec`` is the reference ``BitwiseARModel.state_dict()`` key layout
(814 entries for assets/config.json; names/shapes verified by a strict
``load_state_dict`` into the live reference in ``oracle/make_golden.py``). The same
spec drives the strict checkpoint validation in ``artalk_b200.weights``.

Every tensor is drawn from its own ``torch.Generator`` seeded by
``crc32(key) ^ seed`` so a tensor's value does not depend on the depth of the
config or on the iteration order.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import ModelConfig

# init kinds -----------------------------------------------------------------
LINEAR_W, BIAS, LN_W, LN_B, EMBED, NULL_STYLE, SCALE_MUL, STATS_MEAN, STATS_STD, \
    PE, ATTN_MASK_AR, LVL_IDX, ATTN_MASK_VAE, POSCONV_G, POSCONV_V, CONV0_W, UNIT, OUT_W = range(18)


def state_dict_spec(cfg: ModelConfig) -> "OrderedDict[str, Tuple[tuple, torch.dtype, int]]":
    """name -> (shape, dtype, init kind), in the reference's registration order."""
    C, D, L = cfg.embed_dim, cfg.cond_dim, cfg.seq_tokens
    H, code, md = cfg.vae_hidden, cfg.code_dim, cfg.motion_dim
    w = cfg.wav2vec
    f32 = torch.float32
    s: "OrderedDict[str, Tuple[tuple, torch.dtype, int]]" = OrderedDict()

    def add(name, shape, kind, dtype=f32):
        s[name] = (tuple(shape), dtype, kind)

    def linear(prefix, out_f, in_f, bias=True):
        # decoder out_mapping: small init like the reference's xavier(gain=0.05) (bitwise_vae.py:168-169) so that the
        # decoded (normalised) motion is O(1) and the absolute parity tolerances on FLAME codes are meaningful
        add(prefix + ".weight", (out_f, in_f), OUT_W if prefix.endswith("decoder.out_mapping") else LINEAR_W)
        if bias:
            add(prefix + ".bias", (out_f,), BIAS)

    def ln(prefix, n):
        add(prefix + ".weight", (n,), LN_W)
        add(prefix + ".bias", (n,), LN_B)

    add("null_style_cond", (1, 1, C), NULL_STYLE)
    add("pos_embed", (1, L, C), EMBED)
    add("prev_pos_embed", (1, L * cfg.prev_ratio, C), EMBED)
    add("attn_bias_for_masking", (1, 1, L, L * (1 + cfg.prev_ratio)), ATTN_MASK_AR)
    add("lvl_idx", (1, L), LVL_IDX, torch.int64)
    # basic_vae (app/modules/bitwise_vae.py:15-41)
    T2 = cfg.chunk_frames * 2
    add("basic_vae.enc_pos_embed", (1, T2, md), EMBED)
    add("basic_vae.dec_pos_embed", (1, T2, code), EMBED)
    add("basic_vae.attn_mask", (1, 1, T2, T2), ATTN_MASK_VAE)
    add("basic_vae.motion_mean", (md,), STATS_MEAN)
    add("basic_vae.motion_std", (md,), STATS_STD)
    for side, stack, inp, outname, outdim in (
            ("encoder", "encoder_transformer", md, "code_mapping", code),
            ("decoder", "decoder_transformer", code, "out_mapping", md)):
        p = "basic_vae.%s" % side
        linear(p + ".inp_mapping.0", H, inp)
        linear(p + "." + outname, outdim, H)
        for i in range(cfg.vae_depth):
            a, m = "%s.%s.%d" % (p, stack, 2 * i), "%s.%s.%d" % (p, stack, 2 * i + 1)
            ln(a + ".norm", H)
            linear(a + ".to_qkv", 3 * H, H, bias=False)
            linear(a + ".to_out", H, H)
            linear(m + ".0", int(1.5 * H), H)
            linear(m + ".2", H, int(1.5 * H))
    linear("vqfeat_embed", C, code)
    # style encoder (app/modules/style_encoder.py:10-24)
    sd_ = cfg.style_dim
    add("style_encoder.motion_mean", (md,), STATS_MEAN)
    add("style_encoder.motion_std", (md,), STATS_STD)
    add("style_encoder.PE.pe", (1, cfg.style_pe_len, sd_), PE)
    linear("style_encoder.encoder.motion_proj", sd_, md)
    for i in range(cfg.style_layers):
        p = "style_encoder.encoder.transformer.layers.%d" % i
        add(p + ".self_attn.in_proj_weight", (3 * sd_, sd_), LINEAR_W)
        add(p + ".self_attn.in_proj_bias", (3 * sd_,), BIAS)
        linear(p + ".self_attn.out_proj", sd_, sd_)
        linear(p + ".linear1", cfg.style_ffn, sd_)
        linear(p + ".linear2", sd_, cfg.style_ffn)
        ln(p + ".norm1", sd_)
        ln(p + ".norm2", sd_)
    linear("style_cond_embed", C, sd_)
    # wav2vec2 (HF Wav2Vec2Model, do_stable_layer_norm, feat_extract_norm="layer")
    a = "audio_encoder"
    add(a + ".masked_spec_embed", (w.hidden,), UNIT)
    cin = 1
    for i, k in enumerate(w.conv_kernel):
        p = "%s.feature_extractor.conv_layers.%d" % (a, i)
        add(p + ".conv.weight", (w.conv_dim, cin, k), CONV0_W if i == 0 else LINEAR_W)
        add(p + ".conv.bias", (w.conv_dim,), BIAS)
        ln(p + ".layer_norm", w.conv_dim)
        cin = w.conv_dim
    ln(a + ".feature_projection.layer_norm", w.conv_dim)
    linear(a + ".feature_projection.projection", w.hidden, w.conv_dim)
    add(a + ".encoder.pos_conv_embed.conv.bias", (w.hidden,), BIAS)
    add(a + ".encoder.pos_conv_embed.conv.parametrizations.weight.original0",
        (1, 1, w.pos_conv_kernel), POSCONV_G)
    add(a + ".encoder.pos_conv_embed.conv.parametrizations.weight.original1",
        (w.hidden, w.hidden // w.pos_conv_groups, w.pos_conv_kernel), POSCONV_V)
    ln(a + ".encoder.layer_norm", w.hidden)
    for i in range(w.layers):
        p = "%s.encoder.layers.%d" % (a, i)
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            linear(p + ".attention." + nm, w.hidden, w.hidden)
        ln(p + ".layer_norm", w.hidden)
        linear(p + ".feed_forward.intermediate_dense", w.ffn, w.hidden)
        linear(p + ".feed_forward.output_dense", w.hidden, w.ffn)
        ln(p + ".final_layer_norm", w.hidden)
    # AR blocks (app/transformer.py:12-63)
    for i in range(cfg.ar_depth):
        p = "attn_blocks.%d" % i
        add(p + ".attn.scale_mul_1H11", (1, cfg.ar_heads, 1, 1), SCALE_MUL)
        linear(p + ".attn.query", C, C)
        linear(p + ".attn.key", C, C, bias=False)
        linear(p + ".attn.value", C, C)
        linear(p + ".attn.proj", C, C)
        linear(p + ".ffn.0", 4 * C, C)
        linear(p + ".ffn.2", C, 4 * C)
        linear(p + ".ada_lin.1", 6 * C, D)
    linear("cond_logits_head.ada_lin.1", 2 * C, D)
    linear("logits_head", 2 * code, C)
    add("lvl_embed.weight", (len(cfg.patch_nums), C), EMBED)
    return s


def level_index(cfg: ModelConfig) -> torch.Tensor:
    """Level of each of the 181 token positions (app/models.py:126-128)."""
    return torch.cat([torch.full((pn,), i, dtype=torch.int64)
                      for i, pn in enumerate(cfg.patch_nums)])


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _make(name: str, shape, dtype, kind: int, cfg: ModelConfig, seed: int) -> torch.Tensor:
    g = _gen(name, seed)
    rn = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float32)
    ru = lambda *sh: torch.rand(*sh, generator=g, dtype=torch.float32)
    if kind == LINEAR_W:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return rn(*shape) * (1.0 / math.sqrt(fan_in))
    if kind == OUT_W:
        return rn(*shape) * (0.15 / math.sqrt(shape[1]))
    if kind == CONV0_W:
        return rn(*shape) * math.sqrt(2.0 / shape[-1])
    if kind == BIAS:
        return rn(*shape) * 0.02
    if kind == LN_W:        # non-trivial affine on purpose (torch default ones/zeros hides bugs)
        return 1.0 + 0.1 * rn(*shape)
    if kind == LN_B:
        return 0.05 * rn(*shape)
    if kind == EMBED:
        return (rn(*shape) * math.sqrt(1.0 / shape[-1] / 3.0)).clamp_(-2, 2)
    if kind == NULL_STYLE:
        return rn(*shape) * 0.5
    if kind == SCALE_MUL:   # init ln 4; one head pushed over the ln 100 clamp (transformer.py:56,72)
        t = math.log(4.0) + 0.3 * rn(*shape)
        t.view(-1)[0] = 5.0
        return t
    if kind == STATS_MEAN:
        return 0.25 * rn(*shape)
    if kind == STATS_STD:
        return 0.05 + 0.45 * ru(*shape)
    if kind == UNIT:
        return ru(*shape)
    if kind == POSCONV_G:
        return 0.5 + 1.5 * ru(*shape)
    if kind == POSCONV_V:
        return rn(*shape) * 0.01
    if kind == PE:          # app/modules/style_encoder.py:49-56 (sinusoidal table)
        n, d = shape[1], shape[2]
        pos = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
        pe = torch.zeros(n, d)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        return pe.unsqueeze(0)
    if kind == LVL_IDX:
        return level_index(cfg).view(1, -1)
    if kind == ATTN_MASK_AR:  # app/models.py:123-135
        lv = level_index(cfg)
        cur = torch.where(lv.view(-1, 1) >= lv.view(1, -1), 0.0, -math.inf)
        prev = torch.zeros(cur.shape[0], cur.shape[1] * cfg.prev_ratio)
        return torch.cat([prev, cur], dim=-1).view(shape).contiguous()
    if kind == ATTN_MASK_VAE:  # app/modules/bitwise_vae.py:67-76
        T = cfg.chunk_frames
        m = torch.zeros(2 * T, 2 * T)
        m[:T, T:] = -math.inf
        return m.view(shape)
    raise ValueError(kind)


def make_state_dict(cfg: ModelConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic checkpoint in the reference's ``state_dict`` layout (fp32, CPU)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, (shape, dtype, kind) in state_dict_spec(cfg).items():
        sd[name] = _make(name, shape, dtype, kind, cfg, seed).to(dtype).contiguous()
    return sd


# FLAME ----------------------------------------------------------------------
FLAME_VERTS, FLAME_FACES, FLAME_JOINTS = 5023, 9976, 5


def make_flame_asset(seed: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    """Synthetic stand-in for ``assets/FLAME_with_eye.pt`` with the keys/shapes read at
    ``app/flame_model/FLAME.py:27-57`` (the real file is licence-gated). Geometry is a
    jittered unit-ish point cloud; landmark embeddings are zeros (``no_lmks=True`` path)."""
    V, F, J = FLAME_VERTS, FLAME_FACES, FLAME_JOINTS
    g = _gen("flame", seed)
    rn = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float32)
    ru = lambda *sh: torch.rand(*sh, generator=g, dtype=torch.float32)
    v_template = 0.1 * rn(V, 3)
    shapedirs = 0.002 * rn(V, 3, 400)
    posedirs = 0.01 * rn(V * 3, 36)
    jr = ru(J, V) ** 8
    jr = jr / jr.sum(dim=1, keepdim=True)
    wts = ru(V, J) ** 4
    wts = wts / wts.sum(dim=1, keepdim=True)
    faces = torch.randint(0, V, (F, 3), generator=g, dtype=torch.int64)
    kintree = torch.tensor([[2 ** 32 - 1, 0, 1, 1, 1], [0, 1, 2, 3, 4]], dtype=torch.int64)
    z = lambda *sh, dt=torch.float32: torch.zeros(*sh, dtype=dt)
    return {
        "flame_model": {"f": faces, "v_template": v_template, "shapedirs": shapedirs,
                        "posedirs": posedirs, "J_regressor": jr, "kintree_table": kintree,
                        "weights": wts},
        "lmk_embeddings": {
            "static_lmk_faces_idx": z(51, dt=torch.int64), "static_lmk_bary_coords": z(51, 3),
            "dynamic_lmk_faces_idx": z(79, 17, dt=torch.int64), "dynamic_lmk_bary_coords": z(79, 17, 3),
            "full_lmk_faces_idx_with_eye": z(1, 70, dt=torch.int64),
            "full_lmk_bary_coords_with_eye": z(1, 70, 3)},
        "lmk_embeddings_mediapipe": {"lmk_face_idx": z(105, dt=torch.int64), "lmk_b_coords": z(105, 3)},
    }


def make_audio(n_clips: int, n_samples: int, seed: int = 1234, first_clip: int = 0) -> torch.Tensor:
    """Synthetic 16 kHz audio, ``0.1*N(0,1)``, one generator per clip index."""
    out = torch.empty(n_clips, n_samples, dtype=torch.float32)
    for i in range(n_clips):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed + first_clip + i)
        out[i] = 0.1 * torch.randn(n_samples, generator=g)
    return out


def make_style_motion(n_clips: int, seed: int = 4321, first_clip: int = 0) -> torch.Tensor:
    out = torch.empty(n_clips, 50, 106, dtype=torch.float32)
    for i in range(n_clips):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed + first_clip + i)
        out[i] = 0.1 * torch.randn(50, 106, generator=g)
    return out
