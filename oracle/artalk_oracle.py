"""ORACLE — test infrastructure, not product code.

CPU fp32 restatement (torch tensor ops on the host) of the reference's audio->motion
path, written from the formula-level spec in SURVEY.md Appendix A and citing the
reference file:line each function follows. It consumes the reference ``state_dict``
layout directly and executes the reference's *schedule* (no KV cache, AdaLN recomputed
every step), so it is the un-cached semantics the CUDA path must reproduce.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module. The product path
(``artalk_b200``) never does and fails loudly without its CUDA library.

Parity pin: the reference has no tests or golden vectors, so this oracle is
pinned against the *live* reference imported from /root/reference by
``oracle/make_golden.py`` (run in the build container), which also writes the fixtures in
``tests/golden/`` that ``tests/test_oracle_golden.py`` re-checks anywhere.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


class Oracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg):
        self.sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        self.cfg = cfg
        self.pn = list(cfg.patch_nums)
        self.L = sum(self.pn)
        self.T = self.pn[-1]
        self.cum = [sum(self.pn[:i + 1]) for i in range(len(self.pn))]

    # ---- helpers ---------------------------------------------------------
    def w(self, name):
        return self.sd[name]

    def lin(self, x, prefix, bias=True):
        return F.linear(x, self.sd[prefix + ".weight"], self.sd[prefix + ".bias"] if bias else None)

    # ---- A.1 wav2vec2 ----------------------------------------------------
    # app/modules/wav2vec.py:11-27 ; transformers modeling_wav2vec2.py:275-299,422-434,
    # 326-368,612-655,730-803
    def audio_encode(self, chunks: torch.Tensor) -> torch.Tensor:
        """chunks (N, 64000) -> (N, 199, 1024)."""
        wc, p = self.cfg.wav2vec, "audio_encoder."
        x = chunks.float()
        mean = x.mean(dim=-1, keepdim=True)
        std = x.std(dim=-1, keepdim=True)                       # unbiased (wav2vec.py:25)
        x = (x - mean) / (std + 1e-6)
        h = x[:, None, :]
        for i, (k, s) in enumerate(zip(wc.conv_kernel, wc.conv_stride)):
            q = p + "feature_extractor.conv_layers.%d." % i
            h = F.conv1d(h, self.w(q + "conv.weight"), self.w(q + "conv.bias"), stride=s)
            h = _ln(h.transpose(1, 2), self.w(q + "layer_norm.weight"), self.w(q + "layer_norm.bias"),
                    wc.ln_eps).transpose(1, 2)
            h = F.gelu(h)
        h = h.transpose(1, 2)                                   # (N, 199, 512)
        h = _ln(h, self.w(p + "feature_projection.layer_norm.weight"),
                self.w(p + "feature_projection.layer_norm.bias"), wc.ln_eps)
        h = self.lin(h, p + "feature_projection.projection")
        # positional conv: weight_norm over dims (0,1) -> g * v / ||v|| (modeling :338-368)
        g = self.w(p + "encoder.pos_conv_embed.conv.parametrizations.weight.original0")
        v = self.w(p + "encoder.pos_conv_embed.conv.parametrizations.weight.original1")
        wpos = g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
        pc = F.conv1d(h.transpose(1, 2), wpos, self.w(p + "encoder.pos_conv_embed.conv.bias"),
                      padding=wc.pos_conv_kernel // 2, groups=wc.pos_conv_groups)
        if wc.pos_conv_kernel % 2 == 0:
            pc = pc[:, :, :-1]
        h = h + F.gelu(pc).transpose(1, 2)
        nh, hd = wc.heads, wc.hidden // wc.heads
        for l in range(wc.layers):
            q = p + "encoder.layers.%d." % l
            a = _ln(h, self.w(q + "layer_norm.weight"), self.w(q + "layer_norm.bias"), wc.ln_eps)
            N, S, _ = a.shape
            qq = self.lin(a, q + "attention.q_proj").view(N, S, nh, hd).transpose(1, 2)
            kk = self.lin(a, q + "attention.k_proj").view(N, S, nh, hd).transpose(1, 2)
            vv = self.lin(a, q + "attention.v_proj").view(N, S, nh, hd).transpose(1, 2)
            att = torch.softmax(qq @ kk.transpose(-1, -2) * (hd ** -0.5), dim=-1) @ vv
            h = h + self.lin(att.transpose(1, 2).reshape(N, S, nh * hd), q + "attention.out_proj")
            f = _ln(h, self.w(q + "final_layer_norm.weight"), self.w(q + "final_layer_norm.bias"), wc.ln_eps)
            f = self.lin(F.gelu(self.lin(f, q + "feed_forward.intermediate_dense")),
                         q + "feed_forward.output_dense")
            h = h + f
        return _ln(h, self.w(p + "encoder.layer_norm.weight"), self.w(p + "encoder.layer_norm.bias"), wc.ln_eps)

    # ---- A.2 audio conditioning (app/models.py:93-95) ----------------------
    def audio_cond(self, feat: torch.Tensor) -> torch.Tensor:
        """(N,199,1024) -> (N,181,1024): adaptive average pool to each scale, concatenated."""
        x = feat.transpose(1, 2)
        return torch.cat([F.adaptive_avg_pool1d(x, pn).transpose(1, 2) for pn in self.pn], dim=1)

    # ---- A.3 style (app/modules/style_encoder.py:26-60, app/models.py:67-73) ----
    def style_cond(self, style_motion: Optional[torch.Tensor], batch: int) -> torch.Tensor:
        null = self.w("null_style_cond")
        if style_motion is None:
            return null.expand(batch, -1, -1).clone()
        c, p = self.cfg, "style_encoder."
        m = (style_motion.float() - self.w(p + "motion_mean")) / self.w(p + "motion_std")
        f = self.lin(m, p + "encoder.motion_proj")
        f = f + self.w(p + "PE.pe")[:, m.shape[1], :]           # quirk 1: one PE row for all tokens
        nh, hd = c.style_heads, c.style_dim // c.style_heads
        for l in range(c.style_layers):
            q = p + "encoder.transformer.layers.%d." % l
            B, S, _ = f.shape
            qkv = F.linear(f, self.w(q + "self_attn.in_proj_weight"), self.w(q + "self_attn.in_proj_bias"))
            qq, kk, vv = [t.view(B, S, nh, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
            att = torch.softmax(qq @ kk.transpose(-1, -2) * (hd ** -0.5), dim=-1) @ vv
            att = self.lin(att.transpose(1, 2).reshape(B, S, nh * hd), q + "self_attn.out_proj")
            f = _ln(f + att, self.w(q + "norm1.weight"), self.w(q + "norm1.bias"), 1e-5)   # post-norm
            ff = self.lin(F.gelu(self.lin(f, q + "linear1")), q + "linear2")
            f = _ln(f + ff, self.w(q + "norm2.weight"), self.w(q + "norm2.bias"), 1e-5)
        s = self.lin(f.mean(dim=1), "style_cond_embed")[:, None]
        return s * 1.1 - null * 0.1

    # ---- A.7 bit-plane operators (app/modules/bitwise_vae.py:264-305) ------------
    def _h(self, bits):
        return (bits.float() * 2 - 1.0) / (self.cfg.code_dim ** 0.5)

    def _up(self, h_BLC):          # linear upsample L -> T along time, align_corners=False
        return F.interpolate(h_BLC.transpose(1, 2), size=self.T, mode="linear").transpose(1, 2)

    def _area(self, f_BTC, n):     # area pool T -> n
        return F.interpolate(f_BTC.transpose(1, 2), size=n, mode="area").transpose(1, 2)

    def bits_to_ms_feat(self, bits: torch.Tensor, upto: Optional[int] = None) -> torch.Tensor:
        """bits (B, >=cum[upto], 32) -> next-scale features for scales 1..upto+1, cat along time.
        upto=None: all but the last scale (180 rows) = ``vqidx_to_feat(multi_scale=True)``;
        upto=pidx: ``vqidx_to_ar_vqfeat(pidx, .)`` (prefix of the former)."""
        h = self._h(bits)
        n_lv = len(self.pn) - 1 if upto is None else upto + 1
        f_hat = torch.zeros(bits.shape[0], self.T, self.cfg.code_dim)
        outs, start = [], 0
        for k in range(n_lv):
            f_hat = f_hat + self._up(h[:, start:start + self.pn[k]])
            start += self.pn[k]
            outs.append(self._area(f_hat, self.pn[k + 1]))
        return torch.cat(outs, dim=1)

    def bits_to_latent(self, bits: torch.Tensor) -> torch.Tensor:
        """``vqidx_to_feat(multi_scale=False)`` (bitwise_vae.py:280-288): (B,181,32) -> (B,100,32)."""
        h = self._h(bits)
        f_hat = torch.zeros(bits.shape[0], self.T, self.cfg.code_dim)
        start = 0
        for k in range(len(self.pn) - 1):
            f_hat = f_hat + self._up(h[:, start:start + self.pn[k]])
            start += self.pn[k]
        return f_hat + h[:, start:]

    def tokens_from_bits(self, style, bits, upto=None):
        """cat[style, vqfeat_embed(ms_feat)] (app/models.py:89,107,113)."""
        return torch.cat([style, self.lin(self.bits_to_ms_feat(bits, upto), "vqfeat_embed")], dim=1)

    # ---- A.8 VAE encoder / decoder (app/modules/bitwise_vae.py:128-215) ----------
    def _vae_stack(self, x, prefix, stack, mask):
        c = self.cfg
        nh, hd = c.vae_heads, c.vae_hidden // c.vae_heads
        x = F.leaky_relu(self.lin(x, prefix + "inp_mapping.0"), 0.2)
        for i in range(c.vae_depth):
            a, m = "%s%s.%d." % (prefix, stack, 2 * i), "%s%s.%d." % (prefix, stack, 2 * i + 1)
            B, S, _ = x.shape
            n = _ln(x, self.w(a + "norm.weight"), self.w(a + "norm.bias"), 1e-5)
            qkv = F.linear(n, self.w(a + "to_qkv.weight")).view(B, S, 3, nh, hd)      # quirk 3
            qq, kk, vv = [qkv[:, :, j].transpose(1, 2) for j in range(3)]
            sc = qq @ kk.transpose(-1, -2) * (c.vae_hidden ** -0.5)                   # quirk 2
            if mask is not None:
                sc = sc + mask
            att = (torch.softmax(sc, dim=-1) @ vv).transpose(1, 2).reshape(B, S, nh * hd)
            x = x + self.lin(att, a + "to_out")
            x = x + self.lin(F.gelu(self.lin(x, m + "0"), approximate="tanh"), m + "2")
        return x

    def vae_decode(self, prev_bits, bits) -> torch.Tensor:
        """``vqidx_to_motion`` (bitwise_vae.py:105-113): returns the *new* half (B,100,106)."""
        z = torch.cat([self.bits_to_latent(prev_bits), self.bits_to_latent(bits)], dim=1)
        x = self._vae_stack(z + self.w("basic_vae.dec_pos_embed"), "basic_vae.decoder.",
                            "decoder_transformer", self.w("basic_vae.attn_mask")[0, 0])
        out = self.lin(x, "basic_vae.decoder.out_mapping")
        out = out * self.w("basic_vae.motion_std") + self.w("basic_vae.motion_mean")
        return out[:, self.T:]

    def vae_encode(self, motion) -> torch.Tensor:
        """encoder half of ``quant_to_vqidx(prev, None)`` (bitwise_vae.py:88-90): (B,100,106)->(B,100,32)."""
        x = (motion - self.w("basic_vae.motion_mean")) / self.w("basic_vae.motion_std")
        x = x + self.w("basic_vae.enc_pos_embed")[:, :self.T]
        x = self._vae_stack(x, "basic_vae.encoder.", "encoder_transformer", None)
        return self.lin(x, "basic_vae.encoder.code_mapping")

    # ---- A.9 residual multi-scale BSQ (bitwise_vae.py:227-242,316-334) -----------
    def bsq_bits(self, enc_out: torch.Tensor) -> torch.Tensor:
        q_scale = 1.0 / (self.cfg.code_dim ** 0.5)
        r, out = enc_out, []
        for pt in self.pn:
            a = self._area(r, pt) if pt != self.T else r
            z = F.normalize(a, dim=-1)
            zhat = q_scale * torch.where(z > 0, torch.ones_like(z), -torch.ones_like(z))
            qz = z + (zhat - z)                                   # quirk 5
            out.append((qz > 0).int())
            r = r - (self._up(qz) if pt != self.T else qz)
        return torch.cat(out, dim=1)

    def motion_to_bits(self, motion):
        return self.bsq_bits(self.vae_encode(motion))

    # ---- A.5/A.6 AR transformer (app/transformer.py:30-79, app/models.py:97-107,145-148) ----
    def _ar_block(self, l, x, prev_tok, cond, bias):
        c, p = self.cfg, "attn_blocks.%d." % l
        B, Lx, C = x.shape
        nh, hd = c.ar_heads, C // c.ar_heads
        ada = self.lin(F.silu(cond), p + "ada_lin.1").view(B, Lx, 6, C)
        g1, g2, s1, s2, b1, b2 = ada.unbind(2)                    # quirk 9
        u = _ln(x, None, None, 1e-6) * (1 + s1) + b1
        z = torch.cat([prev_tok, u], dim=1)
        q = self.lin(u, p + "attn.query").view(B, Lx, nh, hd).transpose(1, 2)
        k = self.lin(z, p + "attn.key", bias=False).view(B, -1, nh, hd).transpose(1, 2)
        v = self.lin(z, p + "attn.value").view(B, -1, nh, hd).transpose(1, 2)
        smul = self.w(p + "attn.scale_mul_1H11").clamp_max(math.log(100)).exp()
        q = F.normalize(q, dim=-1) * smul
        k = F.normalize(k, dim=-1)
        att = torch.softmax(q @ k.transpose(-1, -2) + bias, dim=-1) @ v
        att = self.lin(att.transpose(1, 2).reshape(B, Lx, C), p + "attn.proj")
        x = x + att * g1
        wv = _ln(x, None, None, 1e-6) * (1 + s2) + b2
        ff = self.lin(F.gelu(self.lin(wv, p + "ffn.0"), approximate="tanh"), p + "ffn.2")
        return x + ff * g2

    def _head(self, x, cond):
        B, Lx, C = x.shape
        sc, sh = self.lin(F.silu(cond), "cond_logits_head.ada_lin.1").view(B, Lx, 2, C).unbind(2)
        return self.lin(_ln(x, None, None, 1e-6) * (1 + sc) + sh, "logits_head")

    def lvl_pos(self):
        lv = self.w("lvl_embed.weight")[self.w("lvl_idx")[0]]
        return lv[None] + self.w("pos_embed"), lv[None].repeat(1, self.cfg.prev_ratio, 1) + self.w("prev_pos_embed")

    def ar_chunk(self, cond, style, prev_tok, forced_bits: Optional[torch.Tensor] = None):
        """One 100-frame chunk of the scale loop (app/models.py:96-107).
        cond (B,181,1024), style (B,1,768), prev_tok (B,181,768) = cat[style, embed(ms feat of prev bits)].
        forced_bits (B,181,32): teacher forcing — next-step inputs come from these instead of argmax.
        Returns (logits (B,181,64) of the final step, bits (B,181,32), per-step logits list)."""
        lvl_pos, prev_lvl_pos = self.lvl_pos()
        mask = self.w("attn_bias_for_masking")[0, 0]
        P = self.L * self.cfg.prev_ratio
        prev_in = prev_tok + prev_lvl_pos
        nxt, steps, bits = style, [], None
        for pidx in range(len(self.pn)):
            n = self.cum[pidx]
            x = nxt + lvl_pos[:, :n]
            for l in range(self.cfg.ar_depth):
                x = self._ar_block(l, x, prev_in, cond[:, :n], mask[:n, :n + P])
            logits = self._head(x, cond[:, :n])
            steps.append(logits)
            B = logits.shape[0]
            bits = logits.view(B, n, -1, 2).argmax(dim=-1)        # tie -> bit 0 (app/models.py:104)
            src = forced_bits[:, :n] if forced_bits is not None else bits
            if pidx < len(self.pn) - 1:
                nxt = self.tokens_from_bits(style, src, upto=pidx)
        return steps[-1], bits.int(), steps

    # ---- app/models.py:62-121 ------------------------------------------------------
    def split_chunks(self, audio: torch.Tensor):
        c = self.cfg
        seq = c.frames_for_samples(audio.shape[-1])
        padded_frames = math.ceil(seq / self.T) * self.T
        padded = int(padded_frames / c.fps * c.sample_rate)
        a = torch.cat([audio, audio.new_zeros(audio.shape[0], padded - audio.shape[1])], dim=-1)
        return seq, list(a.split(c.chunk_samples, dim=-1))

    def inference(self, audio: torch.Tensor, style_motion: Optional[torch.Tensor] = None,
                  trace: Optional[dict] = None) -> torch.Tensor:
        """audio (B,S) [+ style (B,50,106)] -> motion (B, ceil(S/640), 106). Batched == per-clip loop."""
        B = audio.shape[0]
        seq, chunks = self.split_chunks(audio.float())
        style = self.style_cond(style_motion, B)
        prev_bits = self.motion_to_bits(torch.zeros(B, self.T, self.cfg.motion_dim))
        prev_tok = self.tokens_from_bits(style, prev_bits)
        outs: List[torch.Tensor] = []
        if trace is not None:
            trace.update(style=style, init_bits=prev_bits, cond=[], logits=[], bits=[], motion=[],
                         enc_out=[], prev_bits=[])
        for ch in chunks:
            cond = self.audio_cond(self.audio_encode(ch))
            logits, bits, _ = self.ar_chunk(cond, style, prev_tok)
            motion = self.vae_decode(prev_bits, bits)               # quirk 7: re-encoded prev bits
            outs.append(motion)
            enc = self.vae_encode(motion)
            prev_bits = self.bsq_bits(enc)
            prev_tok = self.tokens_from_bits(style, prev_bits)
            if trace is not None:
                trace["cond"].append(cond); trace["logits"].append(logits); trace["bits"].append(bits)
                trace["motion"].append(motion); trace["enc_out"].append(enc); trace["prev_bits"].append(prev_bits)
        return torch.cat(outs, dim=1)[:, :seq]

    # ---- inference.py:47-57,89-95 ----------------------------------------------------
    @staticmethod
    def smooth_savgol(motion: torch.Tensor) -> torch.Tensor:
        from scipy.signal import savgol_filter
        m = motion.detach().cpu().numpy()
        sm = savgol_filter(m, window_length=5, polyorder=2, axis=0)
        sm[..., 100:103] = savgol_filter(m[..., 100:103], window_length=9, polyorder=3, axis=0)
        return torch.tensor(sm).type_as(motion)

    def engine_inference(self, audio_1d, style_motion=None, clip_length=750, fix_pose=False):
        m = self.inference(audio_1d[None], style_motion)[0]
        m = self.smooth_savgol(m)[:clip_length]
        if fix_pose:
            m[..., 100:103] *= 0.0
        m[..., 104:] *= 0.0
        return m


# ---- A.10 FLAME (app/flame_model/FLAME.py:117-149, app/flame_model/lbs.py:142-383) ----
def rodrigues(r: torch.Tensor) -> torch.Tensor:
    """(N,3) axis-angle -> (N,3,3); angle = ||r + 1e-8|| (quirk 6, lbs.py:294)."""
    angle = torch.norm(r + 1e-8, dim=1, keepdim=True)
    d = r / angle
    c, s = torch.cos(angle)[:, :, None], torch.sin(angle)[:, :, None]
    z = torch.zeros_like(d[:, 0])
    K = torch.stack([z, -d[:, 2], d[:, 1], d[:, 2], z, -d[:, 0], -d[:, 1], d[:, 0], z], dim=1).view(-1, 3, 3)
    return torch.eye(3)[None] + s * K + (1 - c) * (K @ K)


def flame_vertices(asset: dict, shape, expr, pose6, n_shape=300, n_exp=100, scale=1.0) -> torch.Tensor:
    """shape (N,300), expr (N,100), pose6 (N,6)=[global rot, jaw] -> (N,5023,3)."""
    fm = asset["flame_model"]
    N = shape.shape[0]
    sdirs = torch.cat([fm["shapedirs"][:, :, :n_shape], fm["shapedirs"][:, :, 300:300 + n_exp]], dim=2)
    betas = torch.cat([shape, expr], dim=1)
    z3 = torch.zeros(N, 3)
    full_pose = torch.cat([pose6[:, :3], z3, pose6[:, 3:], z3, z3], dim=1)       # global, neck, jaw, eyes
    v_shaped = fm["v_template"][None] + torch.einsum("bl,mkl->bmk", betas, sdirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, fm["J_regressor"])
    R = rodrigues(full_pose.view(-1, 3)).view(N, 5, 3, 3)
    pose_feat = (R[:, 1:] - torch.eye(3)).reshape(N, -1)
    posedirs = fm["posedirs"].reshape(-1, fm["posedirs"].shape[-1]).T          # (36, 15069)
    v_posed = v_shaped + (pose_feat @ posedirs).view(N, -1, 3)
    parents = [-1, 0, 1, 1, 1]
    G = []
    for j in range(5):
        t = J[:, j] if j == 0 else J[:, j] - J[:, parents[j]]
        M = torch.zeros(N, 4, 4)
        M[:, :3, :3], M[:, :3, 3], M[:, 3, 3] = R[:, j], t, 1.0
        G.append(M if j == 0 else G[parents[j]] @ M)
    G = torch.stack(G, dim=1)                                                    # (N,5,4,4)
    Jh = torch.cat([J, torch.zeros(N, 5, 1)], dim=2)[..., None]
    A = G.clone()
    A[..., :, 3:] = G[..., :, 3:] - G @ Jh                                      # A = G - [0 | G (J,0)]
    Tm = (fm["weights"][None] @ A.view(N, 5, 16)).view(N, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(N, v_posed.shape[1], 1)], dim=2)[..., None]
    return (Tm @ vh)[:, :, :3, 0] * scale


def get_flame_verts(asset, shape_params, motion, with_global=False, scale=1.0):
    """``BITWISE_VAE.get_flame_verts`` (bitwise_vae.py:43-57) for 2-D shape params."""
    expr, pose = motion[..., :100], motion[..., 100:]
    if not with_global:
        pose = torch.cat([torch.zeros_like(pose[..., :3]), pose[..., 3:]], dim=-1)
    return flame_vertices(asset, shape_params, expr, pose, scale=scale)


def gaga_t_points(asset, shapecode, motion, forehead_indices, scale=5.0, keep=0.98):
    """app/GAGAvatar/models.py:112-126 run frame by frame like the reference's render loop (inference.py:78-84):
    FLAME(scale 5.0, avatar shape, pose [0,0,0,jaw], zero eye pose) then the forehead EMA. Pinned against the live
    ``GAGAvatar.build_forward_batch`` (rasteriser / pytorch3d imports stubbed, called frame by frame on a stand-in ``self``
    holding a synthetic tracked avatar) by oracle/make_golden.py::run_gaga -> tests/golden/gaga.npz."""
    idx = torch.as_tensor(list(forehead_indices), dtype=torch.long)
    out, upper = [], None
    for f in range(motion.shape[0]):
        mc = motion[f:f + 1]
        pose = torch.cat([mc.new_zeros(1, 3), mc[:, 103:]], dim=-1)
        pts = flame_vertices(asset, shapecode, mc[:, :100], pose, scale=scale).float()
        if upper is None:
            upper = pts[:, idx].clone()
        else:
            upper = keep * upper + (1.0 - keep) * pts[:, idx]
            pts[:, idx] = upper
        out.append(pts[0])
    return torch.stack(out, 0)


def vertex_normals(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """pytorch3d ``Meshes.verts_normals`` (what the reference's mesh renderer shades with, app/flame_model/renderer_utils.py
    builds Meshes(verts, faces); pytorch3d is not installed here, so this restates its published formula: PARITY UNPINNED):
    at every corner of every face accumulate the cross product of the two edges leaving that corner, normalise with eps 1e-6.
    verts (N,V,3), faces (F,3) -> (N,V,3)."""
    f = faces.long()
    v0, v1, v2 = verts[:, f[:, 0]], verts[:, f[:, 1]], verts[:, f[:, 2]]
    n = torch.zeros_like(verts)
    n.index_add_(1, f[:, 1], torch.cross(v2 - v1, v0 - v1, dim=-1))
    n.index_add_(1, f[:, 2], torch.cross(v0 - v2, v1 - v2, dim=-1))
    n.index_add_(1, f[:, 0], torch.cross(v1 - v0, v2 - v0, dim=-1))
    return torch.nn.functional.normalize(n, eps=1e-6, dim=-1)
