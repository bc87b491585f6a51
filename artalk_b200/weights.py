"""Strict ingestion of the reference ``BitwiseARModel.state_dict()`` (wire format in ``synthetic.state_dict_spec``,
loaded with ``strict=True`` at inference.py:28) and repacking into the device layouts the sm_100a kernels consume.

Repacking done here once per weight set (never in the timed path):
  * q/k/v projections concatenated into one [3C, C] operand (AR key has no bias -> zero bias rows);
  * AdaLN ``ada_lin`` of all 12 blocks + the head concatenated into one [56832, 1024] operand (hoisted GEMM);
  * previous-chunk K/V projections of all blocks concatenated into one [12*1536, 768] operand;
  * wav2vec conv weights to channels-last implicit-GEMM form [Cout, k*Cin]; pos-conv weight-norm
    ``g * v / ||v||`` (recomputed every forward by the reference) folded and laid out per group [g][Cout/g][k*Cin/g];
  * ``motion_std`` / ``motion_mean`` folded into the decoder's out_mapping; style mix 1.1*s - 0.1*null folded into
    style_cond_embed; the (buggy, reproduced) single positional-encoding row pe[50] folded into the style projection bias;
  * level + position embedding tables pre-summed; per-head ``exp(min(scale_mul, ln 100))`` precomputed;
  * interpolation operators of the multi-scale quantiser as index/weight tables (ATen's align_corners=False formula).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import ModelConfig
from .synthetic import state_dict_spec, level_index, ATTN_MASK_AR, ATTN_MASK_VAE, LVL_IDX, _make


class CheckpointError(RuntimeError):
    pass


def validate_state_dict(sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> None:
    """Same contract as ``load_state_dict(strict=True)``: no missing or unexpected keys, exact shapes. Also checks
    that the mask / level-index buffers have the block-causal structure the KV-cached schedule relies on."""
    spec = state_dict_spec(cfg)
    missing = [k for k in spec if k not in sd]
    unexpected = [k for k in sd if k not in spec]
    if missing or unexpected:
        raise CheckpointError("Error(s) in loading state_dict: Missing key(s): %s. Unexpected key(s): %s." %
                              (missing[:8], unexpected[:8]))
    for k, (shape, dtype, kind) in spec.items():
        if tuple(sd[k].shape) != tuple(shape):
            raise CheckpointError("size mismatch for %s: checkpoint %s, model %s" % (k, tuple(sd[k].shape), tuple(shape)))
        if kind in (ATTN_MASK_AR, ATTN_MASK_VAE, LVL_IDX):
            want = _make(k, shape, dtype, kind, cfg, 0)
            if not torch.equal(sd[k].detach().cpu().to(want.dtype), want):
                raise CheckpointError("buffer %s does not have the expected block-causal structure" % k)


def interp_tables(cfg: ModelConfig):
    """Index/weight tables of F.interpolate(mode='linear', align_corners=False) pn[k] -> T and of mode='area'
    (adaptive average pooling) T -> pn[k], following ATen's area_pixel_compute_source_index in fp32."""
    T, n = cfg.chunk_frames, len(cfg.patch_nums)
    i0 = np.zeros((n, T), np.int32); i1 = np.zeros((n, T), np.int32); w1 = np.zeros((n, T), np.float32)
    ps = np.zeros((n, T), np.int32); pe = np.zeros((n, T), np.int32)
    for k, p in enumerate(cfg.patch_nums):
        scale = np.float32(p) / np.float32(T)
        for d in range(T):
            src = scale * np.float32(d + 0.5) - np.float32(0.5)
            if src < 0:
                src = np.float32(0.0)
            a = int(src)
            i0[k, d] = a
            i1[k, d] = a + (1 if a < p - 1 else 0)
            w1[k, d] = min(max(np.float32(src) - np.float32(a), np.float32(0)), np.float32(1))
        for i in range(p):
            ps[k, i] = (i * T) // p
            pe[k, i] = -((-(i + 1) * T) // p)
    return i0, i1, w1, ps, pe


def savgol_hat(window: int, poly: int) -> np.ndarray:
    """Hat matrix of the degree-``poly`` least-squares fit over ``window`` samples: row i gives the fitted value at
    window position i. Centre row = scipy's savgol_coeffs (interior); outer rows = mode='interp' edge fits."""
    x = np.arange(window, dtype=np.float64) - window // 2
    V = np.vander(x, poly + 1, increasing=True)
    return (V @ np.linalg.pinv(V)).astype(np.float32)


def conv0_fold(cw: torch.Tensor, b: torch.Tensor, g: torch.Tensor):
    """Conv layer 0 (512, 10) + LayerNorm(512) folded (csrc/conv0_fold.cu): with wc = w - mean_c(w), bc = b - mean_c(b) the
    centred conv output is wc.x + bc and its variance over channels is [x;1]^T Q [x;1], Q = Z^T Z / 512, Z = [wc | bc].
    Returns wq [10][512] = (wc * gamma)^T, bq [512] = bc * gamma and qf [11][12]: rows sqrt(lambda_i) v_i of Q's
    eigen-decomposition (fp64), so that var = sum_i (qf_i . [x;1])^2; column 11 is padding."""
    cw64, b64, g64 = cw.double(), b.double(), g.double()
    wc = cw64 - cw64.mean(0, keepdim=True)
    bc = b64 - b64.mean()
    z = torch.cat([wc, bc[:, None]], 1)                                         # (512, 11)
    lam, vec = torch.linalg.eigh(z.t() @ z / z.shape[0])
    qf = torch.zeros(11, 12, dtype=torch.float64)
    qf[:, :11] = (vec * lam.clamp_min(0).sqrt()[None, :]).t()
    return (wc * g64[:, None]).t().contiguous().float(), (bc * g64).float(), qf.float()


def posconv_shift4(pw: torch.Tensor, groups: int) -> torch.Tensor:
    """Positional-conv weight (hidden, hidden / groups, taps) -> the four-frames-per-row operand of csrc/posconv_tc.cu:
    B'[g][s*gw + co][j'*gw + ci] = W[g*gw + co][ci][j' - s] for the four frame shifts s, j' in [0, taps + 3), zero where j' - s
    falls outside the kernel (out[4t'+s] = sum_j' B'[s][j'] x[4t' + j' - taps/2])."""
    H, gw, K = pw.shape
    src = pw.view(groups, H // groups, gw, K).permute(0, 1, 3, 2)                   # [g][co][j][ci]
    w4 = torch.zeros(groups, 4, H // groups, K + 3, gw, dtype=pw.dtype, device=pw.device)
    for s in range(4):
        w4[:, s, :, s:s + K, :] = src
    return w4.reshape(groups, 4 * (H // groups), (K + 3) * gw)


def repack(sd: Dict[str, torch.Tensor], cfg: ModelConfig, device, precision: str) -> Dict[str, torch.Tensor]:
    """reference state_dict -> {canonical name: contiguous device tensor}."""
    validate_state_dict(sd, cfg)
    wt = torch.bfloat16 if precision == "bf16" else torch.float32       # "bf16x3" / "bf16x6": the engine splits fp32 weights itself
    f32 = torch.float32
    g = lambda k: sd[k].detach().to("cpu", f32)
    out: Dict[str, torch.Tensor] = {}

    def put(name, t, dtype=f32):
        out[name] = t.to(dtype).contiguous().to(device)

    w = cfg.wav2vec
    C, H, CD = cfg.embed_dim, w.hidden, w.conv_dim
    a = "audio_encoder."
    # ---- wav2vec2 ----
    for i, k in enumerate(w.conv_kernel):
        p = a + "feature_extractor.conv_layers.%d." % i
        cw = g(p + "conv.weight")
        if i == 0:
            put("w2v.conv0.w", cw[:, 0, :].t())                                   # [k][512]
        else:
            put("w2v.conv%d.w" % i, cw.permute(0, 2, 1).reshape(CD, k * CD), wt)      # [Cout][tap*Cin]
        if i == 0 and precision == "bf16" and k == 10 and w.conv_stride[0] == 5 and CD == 512:
            wq, bq, qf = conv0_fold(cw[:, 0, :], g(p + "conv.bias"), g(p + "layer_norm.weight"))
            put("w2v.conv0.wq", wq); put("w2v.conv0.bq", bq); put("w2v.conv0.qf", qf)
        put("w2v.conv%d.b" % i, g(p + "conv.bias"))
        put("w2v.conv%d.ln_g" % i, g(p + "layer_norm.weight"))
        put("w2v.conv%d.ln_b" % i, g(p + "layer_norm.bias"))
    put("w2v.proj.ln_g", g(a + "feature_projection.layer_norm.weight"))
    put("w2v.proj.ln_b", g(a + "feature_projection.layer_norm.bias"))
    put("w2v.proj.w", g(a + "feature_projection.projection.weight"), wt)
    put("w2v.proj.b", g(a + "feature_projection.projection.bias"))
    pg = g(a + "encoder.pos_conv_embed.conv.parametrizations.weight.original0")
    pv = g(a + "encoder.pos_conv_embed.conv.parametrizations.weight.original1")
    pw = pg * pv / pv.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()                  # (1024, 64, 128)
    G, gw, K = w.pos_conv_groups, H // w.pos_conv_groups, w.pos_conv_kernel
    put("w2v.pos.w", pw.view(G, gw, gw, K).permute(0, 1, 3, 2).reshape(G, gw, K * gw), wt)   # [g][out][tap*in]
    if precision == "bf16" and gw == 64 and K == 128:
        put("w2v.pos.w4", posconv_shift4(pw, G), wt)
    put("w2v.pos.b", g(a + "encoder.pos_conv_embed.conv.bias"))
    put("w2v.enc_ln_g", g(a + "encoder.layer_norm.weight"))
    put("w2v.enc_ln_b", g(a + "encoder.layer_norm.bias"))
    for l in range(w.layers):
        p = a + "encoder.layers.%d." % l
        put("w2v.l%d.ln1_g" % l, g(p + "layer_norm.weight")); put("w2v.l%d.ln1_b" % l, g(p + "layer_norm.bias"))
        put("w2v.l%d.qkv.w" % l, torch.cat([g(p + "attention.%s_proj.weight" % n) for n in "qkv"], 0), wt)
        put("w2v.l%d.qkv.b" % l, torch.cat([g(p + "attention.%s_proj.bias" % n) for n in "qkv"], 0))
        put("w2v.l%d.out.w" % l, g(p + "attention.out_proj.weight"), wt); put("w2v.l%d.out.b" % l, g(p + "attention.out_proj.bias"))
        put("w2v.l%d.ln2_g" % l, g(p + "final_layer_norm.weight")); put("w2v.l%d.ln2_b" % l, g(p + "final_layer_norm.bias"))
        put("w2v.l%d.ff1.w" % l, g(p + "feed_forward.intermediate_dense.weight"), wt)
        put("w2v.l%d.ff1.b" % l, g(p + "feed_forward.intermediate_dense.bias"))
        put("w2v.l%d.ff2.w" % l, g(p + "feed_forward.output_dense.weight"), wt)
        put("w2v.l%d.ff2.b" % l, g(p + "feed_forward.output_dense.bias"))
    # ---- AR ----
    ada_w = [g("attn_blocks.%d.ada_lin.1.weight" % l) for l in range(cfg.ar_depth)] + [g("cond_logits_head.ada_lin.1.weight")]
    ada_b = [g("attn_blocks.%d.ada_lin.1.bias" % l) for l in range(cfg.ar_depth)] + [g("cond_logits_head.ada_lin.1.bias")]
    put("ar.ada.w", torch.cat(ada_w, 0), wt); put("ar.ada.b", torch.cat(ada_b, 0))
    zC = torch.zeros(C)
    pkw, pkb = [], []
    for l in range(cfg.ar_depth):
        p = "attn_blocks.%d." % l
        qw, kw, vw = g(p + "attn.query.weight"), g(p + "attn.key.weight"), g(p + "attn.value.weight")
        qb, vb = g(p + "attn.query.bias"), g(p + "attn.value.bias")
        put("ar.l%d.qkv.w" % l, torch.cat([qw, kw, vw], 0), wt); put("ar.l%d.qkv.b" % l, torch.cat([qb, zC, vb], 0))
        pkw += [kw, vw]; pkb += [zC, vb]
        put("ar.l%d.head_scale" % l, g(p + "attn.scale_mul_1H11").reshape(-1).clamp_max(math.log(100)).exp())
        put("ar.l%d.proj.w" % l, g(p + "attn.proj.weight"), wt); put("ar.l%d.proj.b" % l, g(p + "attn.proj.bias"))
        put("ar.l%d.ff1.w" % l, g(p + "ffn.0.weight"), wt); put("ar.l%d.ff1.b" % l, g(p + "ffn.0.bias"))
        put("ar.l%d.ff2.w" % l, g(p + "ffn.2.weight"), wt); put("ar.l%d.ff2.b" % l, g(p + "ffn.2.bias"))
    put("ar.prevkv.w", torch.cat(pkw, 0), wt); put("ar.prevkv.b", torch.cat(pkb, 0))
    put("ar.head.w", g("logits_head.weight"), wt); put("ar.head.b", g("logits_head.bias"))
    put("ar.embed.w", g("vqfeat_embed.weight")); put("ar.embed.b", g("vqfeat_embed.bias"))
    lv = g("lvl_embed.weight")[level_index(cfg)]
    put("ar.lvl_pos", lv + g("pos_embed")[0]); put("ar.prev_lvl_pos", lv + g("prev_pos_embed")[0])
    # ---- VAE ----
    mean, std = g("basic_vae.motion_mean"), g("basic_vae.motion_std")
    T = cfg.chunk_frames
    for side, stack in (("dec", "decoder.decoder_transformer"), ("enc", "encoder.encoder_transformer")):
        p = "basic_vae." + ("decoder." if side == "dec" else "encoder.")
        iw = g(p + "inp_mapping.0.weight")
        if side == "enc":
            iw = torch.cat([iw, torch.zeros(iw.shape[0], 128 - iw.shape[1])], 1)      # K 106 -> 128
        put("vae.%s.in.w" % side, iw, wt); put("vae.%s.in.b" % side, g(p + "inp_mapping.0.bias"))
        for l in range(cfg.vae_depth):
            at, ml = "basic_vae.%s.%d." % (stack, 2 * l), "basic_vae.%s.%d." % (stack, 2 * l + 1)
            put("vae.%s.l%d.ln_g" % (side, l), g(at + "norm.weight")); put("vae.%s.l%d.ln_b" % (side, l), g(at + "norm.bias"))
            put("vae.%s.l%d.qkv.w" % (side, l), g(at + "to_qkv.weight"), wt)
            put("vae.%s.l%d.out.w" % (side, l), g(at + "to_out.weight"), wt); put("vae.%s.l%d.out.b" % (side, l), g(at + "to_out.bias"))
            put("vae.%s.l%d.ff1.w" % (side, l), g(ml + "0.weight"), wt); put("vae.%s.l%d.ff1.b" % (side, l), g(ml + "0.bias"))
            put("vae.%s.l%d.ff2.w" % (side, l), g(ml + "2.weight"), wt); put("vae.%s.l%d.ff2.b" % (side, l), g(ml + "2.bias"))
    ow, ob = g("basic_vae.decoder.out_mapping.weight"), g("basic_vae.decoder.out_mapping.bias")
    put("vae.dec.out.w", ow * std[:, None], wt); put("vae.dec.out.b", ob * std + mean)        # unnorm_with_stats folded
    put("vae.enc.out.w", g("basic_vae.encoder.code_mapping.weight"), wt); put("vae.enc.out.b", g("basic_vae.encoder.code_mapping.bias"))
    put("vae.dec_pos", g("basic_vae.dec_pos_embed")[0]); put("vae.enc_pos", g("basic_vae.enc_pos_embed")[0, :T])
    put("vae.mean", mean); put("vae.std", std)
    # ---- style encoder (fp32) ----
    s = "style_encoder."
    put("style.mean", g(s + "motion_mean")); put("style.std", g(s + "motion_std"))
    put("style.zero_pos", torch.zeros(cfg.style_len, cfg.motion_dim))
    pw_ = g(s + "encoder.motion_proj.weight")
    put("style.proj.w", torch.cat([pw_, torch.zeros(pw_.shape[0], 112 - pw_.shape[1])], 1))
    put("style.proj.b", g(s + "encoder.motion_proj.bias") + g(s + "PE.pe")[0, cfg.style_len])   # quirk 1
    for l in range(cfg.style_layers):
        p = s + "encoder.transformer.layers.%d." % l
        put("style.l%d.qkv.w" % l, g(p + "self_attn.in_proj_weight")); put("style.l%d.qkv.b" % l, g(p + "self_attn.in_proj_bias"))
        put("style.l%d.out.w" % l, g(p + "self_attn.out_proj.weight")); put("style.l%d.out.b" % l, g(p + "self_attn.out_proj.bias"))
        put("style.l%d.ln1_g" % l, g(p + "norm1.weight")); put("style.l%d.ln1_b" % l, g(p + "norm1.bias"))
        put("style.l%d.ff1.w" % l, g(p + "linear1.weight")); put("style.l%d.ff1.b" % l, g(p + "linear1.bias"))
        put("style.l%d.ff2.w" % l, g(p + "linear2.weight")); put("style.l%d.ff2.b" % l, g(p + "linear2.bias"))
        put("style.l%d.ln2_g" % l, g(p + "norm2.weight")); put("style.l%d.ln2_b" % l, g(p + "norm2.bias"))
    null = g("null_style_cond").reshape(-1)
    put("style.embed.w", 1.1 * g("style_cond_embed.weight")); put("style.embed.b", 1.1 * g("style_cond_embed.bias") - 0.1 * null)
    put("style.null", null)
    # ---- operator tables ----
    i0, i1, w1, ps, pe = interp_tables(cfg)
    put("tb.up_i0", torch.from_numpy(i0), torch.int32); put("tb.up_i1", torch.from_numpy(i1), torch.int32)
    put("tb.up_w1", torch.from_numpy(w1)); put("tb.pool_start", torch.from_numpy(ps), torch.int32)
    put("tb.pool_end", torch.from_numpy(pe), torch.int32)
    return out
