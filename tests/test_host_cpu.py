"""CPU: host-side logic — C-ABI library loads and exports every symbol the header declares, strict checkpoint
validation, operator tables against ATen / scipy, repacking algebra, loud failure without CUDA."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from artalk_b200 import config, synthetic, _lib
from artalk_b200 import weights as W
import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "artalk_b200.h")).read()
    declared = set(re.findall(r"\b(artalk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    l = _lib.lib()                                  # raises if the .so is missing (no fallback)
    for name in declared:
        assert hasattr(l, name), name
    assert l.artalk_abi_version() == 2


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from artalk_b200.model import BitwiseARModel
    m = BitwiseARModel(config.TINY, device="cuda")
    with pytest.raises(_lib.ArtalkError):
        m.load_state_dict({})
    with pytest.raises(_lib.ArtalkError):
        BitwiseARModel(config.TINY, device="cpu").load_state_dict({})


def test_config_roundtrip_and_lengths():
    j = config.FULL.to_reference_json()
    j["AR_CONFIG"]["AUDIO_ENCODER"] = "wav2vec"
    c = config.ModelConfig.from_reference_json(j)
    assert c == config.FULL
    assert c.seq_tokens == 181 and c.chunk_frames == 100 and c.chunk_samples == 64000 and c.audio_frames == 199
    assert c.wav2vec.conv_lengths(64000) == [12799, 6399, 3199, 1599, 799, 399, 199]
    assert c.frames_for_samples(217088) == 340 and c.chunks_for_samples(217088) == 4     # demo/eng1.wav
    assert c.frames_for_samples(64160) == 101
    j["AR_CONFIG"]["AUDIO_ENCODER"] = "mimi"
    with pytest.raises(ValueError):
        config.ModelConfig.from_reference_json(j)


def test_strict_validation_cpu():
    sd = dict(gu.state_dict("TINY"))
    W.validate_state_dict(sd, config.TINY)
    bad = dict(sd); bad.pop("lvl_embed.weight")
    with pytest.raises(W.CheckpointError, match="Missing"):
        W.validate_state_dict(bad, config.TINY)
    bad = dict(sd); bad["attn_bias_for_masking"] = torch.zeros_like(sd["attn_bias_for_masking"])
    with pytest.raises(W.CheckpointError, match="block-causal"):
        W.validate_state_dict(bad, config.TINY)
    with pytest.raises(W.CheckpointError):
        W.validate_state_dict(sd, config.FULL)        # depth mismatch -> missing keys


def test_interp_tables_match_aten():
    cfg = config.FULL
    i0, i1, w1, ps, pe = W.interp_tables(cfg)
    T = cfg.chunk_frames
    for k, p in enumerate(cfg.patch_nums):
        eye = torch.eye(p)[None]                                   # (1, C=p, L=p): channel j is the j-th unit impulse
        up = F.interpolate(eye, size=T, mode="linear")[0]          # (p, T)
        mine = torch.zeros(p, T)
        for t in range(T):
            mine[i0[k, t], t] += 1.0 - w1[k, t]
            mine[i1[k, t], t] += w1[k, t]
        assert torch.allclose(up, mine, atol=1e-7), k
        eyeT = torch.eye(T)[None]
        down = F.interpolate(eyeT, size=p, mode="area")[0]         # (T, p)
        mine = torch.zeros(T, p)
        for i in range(p):
            mine[ps[k, i]:pe[k, i], i] = 1.0 / (pe[k, i] - ps[k, i])
        assert torch.allclose(down, mine, atol=1e-7), k


def test_savgol_hat_matches_scipy():
    from scipy.signal import savgol_coeffs, savgol_filter
    assert np.allclose(W.savgol_hat(5, 2)[2], savgol_coeffs(5, 2)[::-1], atol=1e-6)
    assert np.allclose(W.savgol_hat(9, 3)[4], savgol_coeffs(9, 3)[::-1], atol=1e-6)
    x = np.random.RandomState(0).randn(9)
    assert np.allclose(W.savgol_hat(9, 3) @ x, savgol_filter(x, 9, 3), atol=1e-5)       # T == window: all edge rows
    x = np.random.RandomState(1).randn(30)
    y = savgol_filter(x, 5, 2)
    H = W.savgol_hat(5, 2)
    assert np.allclose(H[:2] @ x[:5], y[:2], atol=1e-5) and np.allclose(H[3:] @ x[-5:], y[-2:], atol=1e-5)


def test_repack_algebra_cpu():
    """Folded operands reproduce the un-folded reference arithmetic (checked with the oracle's formulas on CPU)."""
    cfg = config.TINY
    sd = gu.state_dict("TINY")
    t = W.repack(sd, cfg, "cpu", "fp32")
    # decoder out_mapping with motion stats folded
    x = torch.randn(7, 512)
    ref = F.linear(x, sd["basic_vae.decoder.out_mapping.weight"], sd["basic_vae.decoder.out_mapping.bias"]) \
        * sd["basic_vae.motion_std"] + sd["basic_vae.motion_mean"]
    assert torch.allclose(F.linear(x, t["vae.dec.out.w"], t["vae.dec.out.b"]), ref, atol=1e-5)
    # style mix folded into the embed
    s = torch.randn(3, 128)
    ref = F.linear(s, sd["style_cond_embed.weight"], sd["style_cond_embed.bias"]) * 1.1 - sd["null_style_cond"].reshape(-1) * 0.1
    assert torch.allclose(F.linear(s, t["style.embed.w"], t["style.embed.b"]), ref, atol=1e-5)
    # pos-conv weight norm folded + grouped tap layout == F.conv1d
    h = torch.randn(1, 40, 1024)
    g_, v_ = sd["audio_encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original0"], \
        sd["audio_encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    wn = torch._weight_norm(v_, g_, 2)
    ref = F.conv1d(h.transpose(1, 2), wn, None, padding=64, groups=16)[:, :, :-1].transpose(1, 2)
    hp = F.pad(h, (0, 0, 64, 64))
    win = hp.unfold(1, 128, 1)[:, :40]                                 # (1, 40, 1024, 128): [t][c][tap]
    win = win.view(1, 40, 16, 64, 128).permute(0, 1, 2, 4, 3).reshape(1, 40, 16, 128 * 64)
    mine = torch.einsum("btgk,gok->btgo", win, t["w2v.pos.w"]).reshape(1, 40, 1024)
    assert torch.allclose(mine, ref, atol=1e-4)
    # the same conv with four output frames per operand row (csrc/posconv_tc.cu): A'[t'][j'] = x[4t' + j' - 64] against the four
    # shifted filter copies of weights.posconv_shift4 gives out[4t' + s] in column block s
    w4 = W.posconv_shift4(wn, 16)                                       # (16, 256, 131 * 64)
    hp4 = F.pad(h, (0, 0, 64, 67 + 3))                                  # frames -64 .. 40 + 69
    a4 = torch.stack([hp4[0, 4 * tp:4 * tp + 131] for tp in range(10)])  # (10, 131, 1024): [t'][j'][c]
    a4 = a4.view(10, 131, 16, 64).permute(0, 2, 1, 3).reshape(10, 16, 131 * 64)
    o4 = torch.einsum("tgk,gnk->tgn", a4, w4).view(10, 16, 4, 64)       # [t'][g][s][co]
    mine4 = o4.permute(0, 2, 1, 3).reshape(1, 40, 1024)                 # frame 4t' + s, channel g * 64 + co
    assert torch.allclose(mine4, ref, atol=1e-4)
    # conv layer 0 with the LayerNorm folded through the conv (csrc/conv0_fold.cu): centred filters + quadratic-form variance
    cw0, b0 = sd["audio_encoder.feature_extractor.conv_layers.0.conv.weight"][:, 0, :], sd["audio_encoder.feature_extractor.conv_layers.0.conv.bias"]
    g0, be0 = sd["audio_encoder.feature_extractor.conv_layers.0.layer_norm.weight"], sd["audio_encoder.feature_extractor.conv_layers.0.layer_norm.bias"]
    wq, bq, qf = W.conv0_fold(cw0, b0, g0)
    xa = torch.randn(37, 10) + 0.3
    ref0 = F.layer_norm(xa @ cw0.t() + b0, (512,), g0, be0, 1e-5)
    z = torch.cat([xa, torch.ones(37, 1)], 1)
    var = ((z @ qf[:, :11].t()) ** 2).sum(1, keepdim=True)
    mine0 = (xa @ wq + bq) * torch.rsqrt(var + 1e-5) + be0
    assert torch.allclose(mine0, ref0, atol=2e-5)
    # conv layer 1 in channels-last implicit-GEMM form
    x = torch.randn(1, 512, 21)
    ref = F.conv1d(x, sd["audio_encoder.feature_extractor.conv_layers.1.conv.weight"], None, stride=2)
    xl = x.transpose(1, 2)                                               # (1, 21, 512)
    rows = torch.stack([xl[0, 2 * i:2 * i + 3].reshape(-1) for i in range(10)])
    assert torch.allclose(rows @ t["w2v.conv1.w"].t(), ref[0].t(), atol=1e-4)
    assert t["ar.ada.w"].shape == (2 * 4608 + 1536, 1024) and t["ar.prevkv.w"].shape == (2 * 1536, 768)
    assert float(t["ar.l0.head_scale"].max()) <= 100.0 + 1e-3 and abs(float(t["ar.l0.head_scale"][0]) - 100.0) < 1e-3


def test_word_packing_roundtrip():
    from artalk_b200.model import pack_words, unpack_words
    bits = torch.randint(0, 2, (4, 181, 32), dtype=torch.int32)
    bits[0, 0] = 1                                                     # all ones -> 0xFFFFFFFF -> int32 -1
    w = pack_words(bits)
    assert w.dtype == torch.int32 and int(w[0, 0]) == -1
    assert torch.equal(unpack_words(w), bits)
    assert torch.equal(gu.unpack_bits(w.numpy().view(np.uint32)), bits)


def test_l2_band_walk_is_a_permutation_of_the_tiles():
    """The CTA-pair GEMM walks N in L2-sized bands when the weights exceed L2 (artalk_b200/csrc/gemm_tc.cu, `decode` of
    gemm_tc2_kernel). Host restatement of that index map: every (M tile, N tile) is visited exactly once, bands are visited in
    order with all M tiles inside a band, and the last band may be narrower."""
    def decode(big, m_all, n_tiles_n, band_n):
        per_band = band_n * m_all
        n_bands = (n_tiles_n + band_n - 1) // band_n
        band = min(big // per_band, n_bands - 1)
        r = big - band * per_band
        width = n_tiles_n - band * band_n if band == n_bands - 1 else band_n
        rest = r // width
        return rest, band * band_n + (r - rest * width)          # (M tile, N tile)

    for m_all, n_tiles_n, band_n in [(46, 222, 64), (5, 8, 4), (74, 5, 2), (3, 7, 7), (3, 7, 100), (1, 1, 1), (10, 9, 4)]:
        seen = [decode(t, m_all, n_tiles_n, band_n) for t in range(m_all * n_tiles_n)]
        assert sorted(seen) == [(m, n) for m in range(m_all) for n in range(n_tiles_n)]
        bands = [n // band_n for _, n in seen]
        assert bands == sorted(bands)                              # one band after the other
        first_band = [mn for mn in seen if mn[1] < band_n]
        assert len({m for m, _ in first_band}) == m_all            # every M tile inside the band
