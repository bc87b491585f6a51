"""GPU: single-kernel parity through the C ABI (artalk_op_* entry points) against plain torch fp32 on the device."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from artalk_b200 import _lib  # noqa: E402


def dev():
    return torch.device("cuda:0")


def rowmap(rpb=0, bs=0, rs=0):
    return _lib.RowMap(rpb, bs, rs)


def run_gemm(precision, A, W, M, N, K, *, a_map=None, bias=None, act=0, gate=None, gate_map=None, resid=None,
             resid_map=None, out32=None, out_act=None, c_map=None, tap_w=0, tap_pad=0, groups=1, a_gs=0, w_gs=0, c_gs=0,
             bias_gs=0, ldw=None, tap_slots=0, exact=0, split_acc=0):
    g = _lib.Gemm()
    g.tap_slots, g.exact, g.split_acc = tap_slots, exact, split_acc
    g.A, g.W = A.data_ptr(), W.data_ptr()
    g.a_map = a_map or rowmap(0, 0, K)
    g.ldw = ldw if ldw is not None else K
    g.M, g.N, g.K = M, N, K
    g.tap_w, g.tap_pad, g.groups, g.a_gs, g.w_gs, g.c_gs, g.bias_gs = tap_w, tap_pad, groups, a_gs, w_gs, c_gs, bias_gs
    g.bias = _lib.ptr(bias)
    g.act = act
    g.gate = _lib.ptr(gate)
    g.gate_dt = _lib.F32 if gate is None or gate.dtype == torch.float32 else _lib.BF16
    g.gate_map = gate_map or rowmap(0, 0, N)
    g.resid = _lib.ptr(resid)
    g.resid_map = resid_map or rowmap(0, 0, N)
    g.out32 = _lib.ptr(out32)
    g.out_act = _lib.ptr(out_act)
    g.out_act_dt = _lib.F32 if out_act is None or out_act.dtype == torch.float32 else _lib.BF16
    g.c_map = c_map or rowmap(0, 0, N)
    _lib.check(_lib.lib().artalk_op_gemm(C.byref(g), precision, _lib.stream_ptr(dev())))
    torch.cuda.synchronize()


ACTS = {0: lambda x: x, 1: F.gelu, 2: lambda x: F.gelu(x, approximate="tanh"), 3: lambda x: F.leaky_relu(x, 0.2), 4: F.silu}


def precisions():
    return [("fp32", 0, torch.float32, 2e-4), ("bf16", 1, torch.bfloat16, 3e-2)]


@pytest.mark.parametrize("pname,prec,dt,tol", precisions())
@pytest.mark.parametrize("M,N,K,act", [(300, 768, 768, 0), (77, 2304, 768, 2), (1, 64, 768, 0), (200, 106, 512, 0),
                                       (513, 512, 32, 3), (129, 1024, 4096, 1), (256, 32, 512, 0), (100, 4608, 1024, 4)])
def test_gemm_bias_act(pname, prec, dt, tol, M, N, K, act):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).to(dev(), dt)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev(), dt)
    b = torch.randn(N, generator=g).to(dev())
    out = torch.full((M, N), float("nan"), device=dev())
    run_gemm(prec, A, W, M, N, K, bias=b, act=act, out32=out)
    ref = ACTS[act](A.double() @ W.double().t() + b.double()).float()     # fp64: torch fp32 matmul may use TF32
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K,act,resid,gated", [(9472, 2048, 256, 1, False, False), (9700, 2048, 512, 0, True, True),
                                                   (18944, 1280, 192, 2, True, True), (9728, 2048, 128, 0, False, False),
                                                   (9700, 2048, 512, 0, True, False), (18944, 1280, 192, 3, True, False)])
def test_gemm_cta_pair_kernel(M, N, K, act, resid, gated):
    """Large plain GEMMs take the cta_group::2 kernel (256x256 tiles over two SMs): same results as the fp64 reference,
    including a ragged last M tile, a ragged N tile (1280 = 5 x 256), the TMA-fed residual tile (residual without gate), a ragged last wave (304 tiles on 74 CTA pairs: the last 8
    tiles run as 32-column slices), gate + in-place residual and dual outputs."""
    dt, tol = torch.bfloat16, 3e-2
    g = torch.Generator(device="cpu").manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(dev(), dt)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev(), dt)
    b = torch.randn(N, generator=g).to(dev())
    if resid:
        x = torch.randn(M, N, generator=g).to(dev())
        gate = torch.randn(M, N, generator=g).to(dev(), dt) if gated else None
        x0 = x.clone()
        xa = torch.empty(M, N, device=dev(), dtype=dt)
        run_gemm(1, A, W, M, N, K, bias=b, act=act, gate=gate, resid=x, out32=x, out_act=xa)
        ref = ACTS[act](A.double() @ W.double().t() + b.double())
        ref = (x0.double() + (ref * gate.double() if gated else ref)).float()
        scale = max(1.0, ref.abs().max().item())
        assert (x - ref).abs().max().item() < tol * scale
        assert (xa.float() - ref).abs().max().item() < (tol + 8e-3) * scale
    else:
        out = torch.full((M, N), float("nan"), device=dev(), dtype=dt)
        run_gemm(1, A, W, M, N, K, bias=b, act=act, out_act=out)
        ref = ACTS[act](A.double() @ W.double().t() + b.double()).float()
        assert torch.isfinite(out.float()).all()
        assert (out.float() - ref).abs().max().item() < (tol + 8e-3) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K,resid", [(9700, 2048, 512, True), (9700, 2048, 512, False), (18944, 1280, 192, True),
                                          (38000, 1024, 1024, True)])
def test_gemm_cta_pair_kernel_tma_store(M, N, K, resid):
    """fp32-only outputs of the CTA-pair kernel leave as TMA stores of each warp's 32 x 32 scratch tile (option gemm_tma_out):
    in-place residual update and plain output, ragged last M tile (rows past M clipped by the tensor map), sliced ragged wave;
    identical to the st.global epilogue bit for bit."""
    dt, tol = torch.bfloat16, 3e-2
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N)
    A = torch.randn(M, K, generator=g).to(dev(), dt)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev(), dt)
    b = torch.randn(N, generator=g).to(dev())
    x0 = torch.randn(M + 7, N, generator=g).to(dev())             # 7 guard rows after the matrix
    ref = (A.double() @ W.double().t() + b.double()).float()
    outs = []
    bouts = []
    for on in (2, 0):
        _lib.check(_lib.lib().artalk_set_option(b"gemm_tma_out", on))
        try:
            x = x0.clone()
            if resid:
                run_gemm(1, A, W, M, N, K, bias=b, resid=x, out32=x)
            else:
                run_gemm(1, A, W, M, N, K, bias=b, out32=x)
            xb = torch.full((M + 7, N), 7.0, device=dev(), dtype=dt)   # bf16-only output (64-byte-swizzle tiles), GELU epilogue
            run_gemm(1, A, W, M, N, K, bias=b, act=1, out_act=xb)
        finally:
            _lib.check(_lib.lib().artalk_set_option(b"gemm_tma_out", 2))
        assert torch.equal(x[M:], x0[M:]) and bool((xb[M:] == 7.0).all())      # nothing written past row M
        outs.append(x[:M].clone()); bouts.append(xb[:M].clone())
    want = (x0[:M] + ref) if resid else ref
    assert (outs[0] - want).abs().max().item() < tol * max(1.0, want.abs().max().item())
    assert torch.equal(outs[0], outs[1])
    gref = F.gelu(ref.double()).float()
    assert (bouts[0].float() - gref).abs().max().item() < (tol + 8e-3) * max(1.0, gref.abs().max().item())
    assert torch.equal(bouts[0], bouts[1])


@pytest.mark.parametrize("M,N,K", [(9700, 2048, 512), (18944, 1280, 1024), (9472, 2304, 256)])
def test_gemm_cta_pair_kernel_l2_bands(M, N, K):
    """Weights larger than L2 (the hoisted AdaLN GEMM, 116 MB) are walked in N bands so that a band of W stays L2-resident.
    Forced here with a 1 MB band: even bands, a narrower last band (5 N tiles in bands of 2) and a single band, with the ragged
    last wave cut into column slices on top."""
    dt, tol = torch.bfloat16, 3e-2
    g = torch.Generator(device="cpu").manual_seed(M + K)
    A = torch.randn(M, K, generator=g).to(dev(), dt)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev(), dt)
    b = torch.randn(N, generator=g).to(dev())
    ref = (A.double() @ W.double().t() + b.double()).float()
    _lib.check(_lib.lib().artalk_set_option(b"gemm_band_mb", 1))
    try:
        out = torch.full((M, N), float("nan"), device=dev(), dtype=dt)
        run_gemm(1, A, W, M, N, K, bias=b, out_act=out)
    finally:
        _lib.check(_lib.lib().artalk_set_option(b"gemm_band_mb", 32))
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < (tol + 8e-3) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K,act,mode", [(64, 768, 3072, 0, "f32"), (64, 3072, 768, 2, "bf16"), (320, 768, 3072, 0, "f32"),
                                             (320, 3072, 768, 2, "bf16"), (1, 768, 768, 0, "f32"), (5, 64, 768, 0, "rows"),
                                             (100, 2304, 768, 1, "dual"), (200, 512, 512, 3, "resid"), (257, 1536, 512, 0, "bf16"),
                                             (512, 768, 768, 4, "resid"), (1280, 2304, 768, 0, "bf16")])
def test_gemm_skinny_latency_path(M, N, K, act, mode):
    """GEMMs with few rows (the 1- and 5-token scale steps, every batch-1 step) take the mma.sync / cp.async kernel of
    skinny.cu: same results as the fp64 reference and as the tcgen05 kernel (option skinny_max_m = 0) on ragged row slabs,
    every column-tile width (8..64), deep and shallow rings, fp32 / bf16 / dual outputs, residual and batched output rows."""
    dt, tol = torch.bfloat16, 3e-2
    g = torch.Generator(device="cpu").manual_seed(M * 11 + N)
    A = torch.randn(M, K, generator=g).to(dev(), dt)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev(), dt)
    b = torch.randn(N, generator=g).to(dev())
    ref = ACTS[act](A.double() @ W.double().t() + b.double()).float()
    outs = []
    for max_m in (2048, 0):
        _lib.check(_lib.lib().artalk_set_option(b"skinny_max_m", max_m))
        try:
            if mode == "resid":
                x = torch.randn(M, N, generator=torch.Generator(device="cpu").manual_seed(3)).to(dev())
                want = x + ref
                run_gemm(1, A, W, M, N, K, bias=b, act=act, resid=x, out32=x)
                got = x
            elif mode == "rows":          # logits rows scattered as (clip, token) like the AR head
                buf = torch.full((M, 7, N), float("nan"), device=dev())
                run_gemm(1, A, W, M, N, K, bias=b, act=act, out32=buf.view(-1)[2 * N:], c_map=rowmap(1, 7 * N, N))
                assert torch.isnan(buf[:, :2]).all() and torch.isnan(buf[:, 3:]).all()
                got, want = buf[:, 2], ref
            elif mode == "dual":
                got = torch.full((M, N), float("nan"), device=dev())
                xa = torch.empty(M, N, device=dev(), dtype=dt)
                run_gemm(1, A, W, M, N, K, bias=b, act=act, out32=got, out_act=xa)
                assert (xa.float() - got).abs().max().item() <= 8e-3 * max(1.0, got.abs().max().item())
                want = ref
            elif mode == "bf16":
                o = torch.full((M, N), float("nan"), device=dev(), dtype=dt)
                run_gemm(1, A, W, M, N, K, bias=b, act=act, out_act=o)
                got, want = o.float(), ref
            else:
                got = torch.full((M, N), float("nan"), device=dev())
                run_gemm(1, A, W, M, N, K, bias=b, act=act, out32=got)
                want = ref
        finally:
            _lib.check(_lib.lib().artalk_set_option(b"skinny_max_m", 512))
        assert torch.isfinite(got).all()
        extra = 8e-3 if mode == "bf16" else 0.0
        assert (got - want).abs().max().item() < (tol + extra) * max(1.0, want.abs().max().item())
        outs.append(got.clone())
    # the two kernels accumulate the same bf16 products in fp32: they agree far inside the bf16-vs-fp64 tolerance
    assert (outs[0] - outs[1]).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("pname,prec,dt,tol", precisions())
def test_gemm_gate_residual_rowmaps(pname, prec, dt, tol):
    """AR epilogue: x += (A W^T + b) * gamma with gamma rows gathered through (clip, token) maps; dual outputs."""
    B, n_new, L, C, K = 3, 25, 181, 768, 3072
    M = B * n_new
    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randn(M, K, generator=g).to(dev(), dt)
    W = (torch.randn(C, K, generator=g) / math.sqrt(K)).to(dev(), dt)
    b = torch.randn(C, generator=g).to(dev())
    n_ada = 6 * C
    ada = torch.randn(B, L, n_ada, generator=g).to(dev(), dt)
    off = 31
    x = torch.randn(M, C, generator=g).to(dev())
    x0 = x.clone()
    xa = torch.empty(M, C, device=dev(), dtype=dt)
    gate_view = ada[:, off:off + n_new, C:2 * C]                   # gamma2 slice
    gate_ptr_t = ada.view(-1)[off * n_ada + C:]
    run_gemm(prec, A, W, M, C, K, bias=b, gate=gate_ptr_t, gate_map=rowmap(n_new, L * n_ada, n_ada), resid=x, out32=x,
             out_act=xa)
    ref = (x0.double() + (A.double() @ W.double().t() + b.double()) * gate_view.reshape(M, C).double()).float()
    scale = max(1.0, ref.abs().max().item())
    assert (x - ref).abs().max().item() < tol * scale
    assert (xa.float() - ref).abs().max().item() < (tol + (0 if dt == torch.float32 else 8e-3)) * scale


@pytest.mark.parametrize("pname,prec,dt,tol", precisions())
@pytest.mark.parametrize("k,s", [(3, 2), (2, 2)])
def test_conv_as_overlapping_row_gemm(pname, prec, dt, tol, k, s):
    """wav2vec conv layers 1-6: channels-last implicit GEMM with overlapping A rows == F.conv1d."""
    n, L_in, Cc = 3, 401, 512
    L_out = (L_in - k) // s + 1
    g = torch.Generator(device="cpu").manual_seed(k)
    x = torch.randn(n, L_in, Cc, generator=g).to(dev(), dt)                      # channels last
    w = (torch.randn(Cc, Cc, k, generator=g) / math.sqrt(Cc * k)).to(dev(), dt)   # torch layout (out,in,k)
    b = torch.randn(Cc, generator=g).to(dev())
    wp = w.permute(0, 2, 1).reshape(Cc, k * Cc).contiguous()
    out = torch.empty(n * L_out, Cc, device=dev())
    run_gemm(prec, x, wp, n * L_out, Cc, k * Cc, a_map=rowmap(L_out, L_in * Cc, s * Cc), bias=b, out32=out)
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), stride=s).transpose(1, 2).reshape(n * L_out, Cc).float()
    assert (out - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("pname,prec,dt,tol", precisions())
def test_pos_conv_grouped_tap_gemm(pname, prec, dt, tol):
    """wav2vec positional conv: grouped conv1d k=128 pad 64 (last sample dropped) + GELU + residual."""
    n, Fr, H, G, K = 2, 199, 1024, 16, 128
    gw = H // G
    g = torch.Generator(device="cpu").manual_seed(11)
    h = torch.randn(n, Fr, H, generator=g).to(dev())
    w = (torch.randn(H, gw, K, generator=g) / math.sqrt(gw * K)).to(dev())
    b = torch.randn(H, generator=g).to(dev())
    A = h.to(dt)
    wp = w.view(G, gw, gw, K).permute(0, 1, 3, 2).reshape(G, gw, K * gw).contiguous().to(dt)
    out = torch.empty(n * Fr, H, device=dev())
    run_gemm(prec, A, wp, n * Fr, gw, K * gw, a_map=rowmap(Fr, Fr * H, H), tap_w=gw, tap_pad=K // 2, groups=G, a_gs=gw,
             w_gs=gw * K * gw, c_gs=gw, bias_gs=gw, bias=b, act=1, resid=h.view(-1, H), resid_map=rowmap(0, 0, H),
             out32=out, c_map=rowmap(0, 0, H), ldw=K * gw)
    pc = F.conv1d(A.double().transpose(1, 2), w.to(dt).double(), b.double(), padding=K // 2, groups=G)[:, :, :-1]
    ref = (h.double() + F.gelu(pc).transpose(1, 2)).reshape(n * Fr, H).float()
    assert (out - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("n,S", [(1, 64000), (3, 64000), (2, 3212)])
def test_conv0_layernorm_folded_through_the_conv(n, S):
    """conv0_fold.cu: conv layer 0 + LayerNorm + GELU with the LayerNorm statistics taken from the frame's 10 samples (quadratic
    form of the channel-centred filters, weights.conv0_fold) against (a) the fp64 reference and (b) the direct kernel, whose
    bf16 outputs it must match to within one bf16 step of the GELU output (the two differ in fp32 rounding only). Inputs with a
    DC offset and a common filter component exercise the centring."""
    from artalk_b200 import weights
    g = torch.Generator(device="cpu").manual_seed(5 + n)
    audio = (0.1 * torch.randn(n, S, generator=g) + 0.05).to(dev())
    cw = (torch.randn(512, 10, generator=g) * 0.3 + 0.1 * torch.randn(1, 10, generator=g)).to(dev())
    b = (0.2 * torch.randn(512, generator=g) + 0.1).to(dev())
    lg = (1.0 + 0.2 * torch.randn(512, generator=g)).to(dev())
    lb = (0.1 * torch.randn(512, generator=g)).to(dev())
    wq, bq, qf = [t.to(dev()).contiguous() for t in weights.conv0_fold(cw.cpu(), b.cpu(), lg.cpu())]
    L = (S - 10) // 5 + 1
    ws = torch.empty(2 * n, device=dev())
    w_kc = cw.t().contiguous()
    outs = []
    for fold in (False, True):
        out = torch.full((n, L, 512), float("nan"), device=dev(), dtype=torch.bfloat16)
        _lib.check(_lib.lib().artalk_op_conv0(audio.data_ptr(), n, S, w_kc.data_ptr(), b.data_ptr(), lg.data_ptr(), lb.data_ptr(),
                                              wq.data_ptr() if fold else None, bq.data_ptr() if fold else None,
                                              qf.data_ptr() if fold else None, ws.data_ptr(), out.data_ptr(), _lib.BF16, 1e-5,
                                              _lib.stream_ptr(dev())))
        torch.cuda.synchronize()
        assert bool(torch.isfinite(out.float()).all())
        outs.append(out.float())
    a64 = audio.double()
    xn = (a64 - a64.mean(1, keepdim=True)) / (a64.std(1, keepdim=True) + 1e-6)          # app/modules/wav2vec.py:23-27
    y = F.conv1d(xn[:, None, :], cw.double()[:, None, :], b.double(), stride=5).transpose(1, 2)
    ref = F.gelu(F.layer_norm(y, (512,), lg.double(), lb.double(), 1e-5)).float()
    for o in outs:
        assert (o - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    d = (outs[0] - outs[1]).abs()
    assert d.max().item() <= 2.0 ** -7 * max(1.0, ref.abs().max().item())      # at most one bf16 step of the largest output
    assert (d > 0).float().mean().item() < 0.05                                 # and only where the fp32 value sat on a rounding boundary


@pytest.mark.parametrize("n", [1, 2, 5, 11])
def test_pos_conv_four_frames_per_row(n):
    """posconv_tc.cu: the positional conv with four output frames per A row (N = 256 CTA-pair tiles, shifted filter copies from
    weights.posconv_shift4). Same products in the same tap order as the N = 64 tap-mode GEMM (the added ones are exact zeros):
    bit-identical to it, and within the bf16 tolerance of the fp64 reference; chunk counts that leave the last CTA / tile partly
    or wholly out of range (1, 2, 5) and more than one tile (11)."""
    from artalk_b200 import weights
    Fr, H, G, K = 199, 1024, 16, 128
    gw = H // G
    g = torch.Generator(device="cpu").manual_seed(17 + n)
    h = torch.randn(n, Fr, H, generator=g).to(dev())
    w = (torch.randn(H, gw, K, generator=g) / math.sqrt(gw * K)).to(dev())
    b = torch.randn(H, generator=g).to(dev())
    A = h.to(torch.bfloat16)
    w4 = weights.posconv_shift4(w, G).to(torch.bfloat16).contiguous()
    assert w4.shape == (G, 256, (K + 3) * gw)
    out = torch.full((n * Fr, H), float("nan"), device=dev())
    _lib.check(_lib.lib().artalk_op_posconv4(A.data_ptr(), w4.data_ptr(), b.data_ptr(), h.data_ptr(), out.data_ptr(), n, Fr, H, G, K,
                                             _lib.stream_ptr(dev())))
    torch.cuda.synchronize()
    pc = F.conv1d(A.double().transpose(1, 2), w.to(torch.bfloat16).double(), b.double(), padding=K // 2, groups=G)[:, :, :-1]
    ref = (h.double() + F.gelu(pc).transpose(1, 2)).reshape(n * Fr, H).float()
    assert bool(torch.isfinite(out).all())
    assert (out - ref).abs().max().item() < 3e-2 * max(1.0, ref.abs().max().item())
    wp = w.view(G, gw, gw, K).permute(0, 1, 3, 2).reshape(G, gw, K * gw).contiguous().to(torch.bfloat16)
    old = torch.empty(n * Fr, H, device=dev())
    run_gemm(1, A, wp, n * Fr, gw, K * gw, a_map=rowmap(Fr, Fr * H, H), tap_w=gw, tap_pad=K // 2, groups=G, a_gs=gw,
             w_gs=gw * K * gw, c_gs=gw, bias_gs=gw, bias=b, act=1, resid=h.view(-1, H), resid_map=rowmap(0, 0, H),
             out32=old, c_map=rowmap(0, 0, H), ldw=K * gw)
    assert torch.equal(out, old)


# ----------------------------------------------------------------------------- parity-grade tensor-core mode (split.cu)
def split_op(x, slots, is_w):
    """fp32 tensor -> bf16 piece blocks [numel / 64][slots][64] through artalk_op_split_bf16."""
    x = x.contiguous()
    out = torch.empty(x.numel() * slots, device=dev(), dtype=torch.bfloat16)
    _lib.check(_lib.lib().artalk_op_split_bf16(x.data_ptr(), out.data_ptr(), x.numel(), slots, int(is_w), _lib.stream_ptr(dev())))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("slots", [3, 6])
def test_split_pieces_reconstruct_operand(slots):
    """The piece blocks hold p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) in the slot order A: 0 1 0 | 1 2 0,
    W: 0 0 1 | 1 0 2, and the pieces sum back to x to 2^-17 (2 pieces) / 2^-24 (3 pieces) relative."""
    g = torch.Generator(device="cpu").manual_seed(slots)
    x = (torch.randn(3 * 64 * 5, generator=g) * torch.logspace(-3, 3, 960)).to(dev())
    p0 = x.to(torch.bfloat16).float()
    p1 = (x - p0).to(torch.bfloat16).float()
    p2 = (x - p0 - p1).to(torch.bfloat16).float()
    pieces = [p0, p1, p2]
    order = {(3, 0): [0, 1, 0], (3, 1): [0, 0, 1], (6, 0): [0, 1, 0, 1, 2, 0], (6, 1): [0, 0, 1, 1, 0, 2]}
    for is_w in (0, 1):
        got = split_op(x, slots, is_w).float().view(-1, slots, 64)
        for s_, pc in enumerate(order[(slots, is_w)]):
            assert torch.equal(got[:, s_], pieces[pc].view(-1, 64)), (slots, is_w, s_)
    n_p = 2 if slots == 3 else 3
    rel = ((sum(pieces[:n_p]) - x).abs() / x.abs().clamp_min(1e-30)).max().item()
    assert rel < (2.0 ** -16 if slots == 3 else 2.0 ** -23), rel


@pytest.mark.parametrize("slots,tol", [(3, 3e-5), (6, 1.5e-6)])
@pytest.mark.parametrize("M,N,K,act", [(300, 768, 768, 0), (129, 1024, 4096, 1), (1, 64, 768, 0), (9700, 2048, 512, 2),
                                       (200, 106, 512, 0), (38000, 1024, 1024, 0)])
def test_split_gemm_is_fp32_grade(slots, tol, M, N, K, act):
    """The UNCHANGED bf16 tcgen05 kernels (1-CTA and CTA-pair) on piece blocks with K' = slots * K reproduce the fp64 product
    of the fp32 operands: ~2^-17 relative per product with 2 pieces (3 passes), fp32 grade with 3 pieces (6 passes), measured
    against sum_k |a_k||w_k| (the scale rounding errors live on). exact = 1: libm-accurate GELU in the epilogue."""
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + slots)
    A = torch.randn(M, K, generator=g).to(dev())
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev())
    b = torch.randn(N, generator=g).to(dev())
    As, Ws = split_op(A, slots, 0), split_op(W, slots, 1)
    out = torch.full((M, N), float("nan"), device=dev())
    run_gemm(1, As, Ws, M, N, slots * K, bias=b, act=act, out32=out, exact=1, split_acc=slots)
    pre = A.double() @ W.double().t() + b.double()
    ref = ACTS[act](pre)
    mag = (A.double().abs() @ W.double().abs().t()).clamp_min(1.0)           # error scale of a length-K dot product
    err = ((out.double() - ref).abs() / mag).max().item()
    assert torch.isfinite(out).all()
    assert err < tol, err


@pytest.mark.parametrize("slots,tol", [(3, 3e-5), (6, 1.5e-6)])
def test_split_conv_window_and_tap_views(slots, tol):
    """Piece blocks keep the strided views working: overlapping-row conv windows (every stride x slots) and the grouped tap
    mode of the positional conv (tap_slots: k-block kb reads slot kb % slots of tap kb / slots)."""
    # conv k=3, s=2 over channels-last input
    n, L_in, Cc, k, s = 3, 401, 512, 3, 2
    L_out = (L_in - k) // s + 1
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(n, L_in, Cc, generator=g).to(dev())
    w = (torch.randn(Cc, Cc, k, generator=g) / math.sqrt(Cc * k)).to(dev())
    b = torch.randn(Cc, generator=g).to(dev())
    wp = w.permute(0, 2, 1).reshape(Cc, k * Cc).contiguous()
    out = torch.empty(n * L_out, Cc, device=dev())
    S = slots
    run_gemm(1, split_op(x, S, 0), split_op(wp, S, 1), n * L_out, Cc, S * k * Cc, a_map=rowmap(L_out, S * L_in * Cc, S * s * Cc),
             bias=b, out32=out, exact=1, split_acc=S)
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), stride=s).transpose(1, 2).reshape(n * L_out, Cc)
    assert ((out.double() - ref).abs().max().item()) < tol * 40
    # positional conv: grouped, k=128, pad 64, GELU + residual
    n, Fr, H, G, K = 2, 199, 1024, 16, 128
    gw = H // G
    h = torch.randn(n, Fr, H, generator=g).to(dev())
    w = (torch.randn(H, gw, K, generator=g) / math.sqrt(gw * K)).to(dev())
    b = torch.randn(H, generator=g).to(dev())
    wp = w.view(G, gw, gw, K).permute(0, 1, 3, 2).reshape(G, gw, K * gw).contiguous()
    out = torch.empty(n * Fr, H, device=dev())
    run_gemm(1, split_op(h, S, 0), split_op(wp, S, 1), n * Fr, gw, S * K * gw, a_map=rowmap(Fr, S * Fr * H, S * H), tap_w=gw,
             tap_pad=K // 2, groups=G, a_gs=gw, w_gs=S * gw * K * gw, c_gs=gw, bias_gs=gw, bias=b, act=1, resid=h.view(-1, H),
             resid_map=rowmap(0, 0, H), out32=out, c_map=rowmap(0, 0, H), ldw=S * K * gw, tap_slots=S, exact=1, split_acc=S)
    pc = F.conv1d(h.double().transpose(1, 2), w.double(), b.double(), padding=K // 2, groups=G)[:, :, :-1]
    ref = h.double() + F.gelu(pc).transpose(1, 2)
    assert (out.double() - ref.reshape(n * Fr, H)).abs().max().item() < tol * 40


def run_attn(q, k, v, out, n_seq, H, D, lq, lk, strides, scale, split=0, key_bound=None):
    a = _lib.Attn()
    a.key_bound = _lib.ptr(key_bound)
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.dt = _lib.F32 if q.dtype == torch.float32 else _lib.BF16
    a.n_seq, a.n_heads, a.head_dim, a.lq, a.lk = n_seq, H, D, lq, lk
    (a.q_ss, a.q_rs, a.k_ss, a.k_rs, a.v_ss, a.v_rs, a.o_ss, a.o_rs) = strides
    a.scale, a.split = scale, split
    _lib.check(_lib.lib().artalk_op_attention(C.byref(a), _lib.stream_ptr(dev())))
    torch.cuda.synchronize()


@pytest.mark.parametrize("dt,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("n_seq,H,D,lq,lk,split", [(3, 16, 64, 199, 199, 0), (2, 8, 64, 200, 200, 100), (4, 12, 64, 25, 212, 0),
                                                   (5, 12, 64, 1, 182, 0), (2, 4, 32, 50, 50, 0), (2, 12, 64, 100, 362, 0),
                                                   (7, 12, 64, 5, 187, 0), (3, 12, 64, 8, 256, 0), (2, 8, 64, 3, 17, 0)])
def test_attention(dt, tol, n_seq, H, D, lq, lk, split):
    g = torch.Generator(device="cpu").manual_seed(lq * 3 + lk)
    Cw = H * D
    q = torch.randn(n_seq, lq, Cw, generator=g).to(dev(), dt)
    k = torch.randn(n_seq, lk, Cw, generator=g).to(dev(), dt)
    v = torch.randn(n_seq, lk, Cw, generator=g).to(dev(), dt)
    out = torch.empty(n_seq, lq, Cw, device=dev(), dtype=dt)
    scale = 0.3
    out.fill_(float("nan"))
    run_attn(q, k, v, out, n_seq, H, D, lq, lk, (lq * Cw, Cw, lk * Cw, Cw, lk * Cw, Cw, lq * Cw, Cw), scale, split)
    _check_attn(out, q, k, v, n_seq, H, D, lq, lk, scale, split, tol)


@pytest.mark.parametrize("n_seq,H,lq,lk", [(4, 12, 25, 212), (5, 12, 1, 182), (2, 12, 100, 362), (3, 12, 50, 262), (7, 12, 5, 187)])
def test_attention_bounded_scores_single_pass(n_seq, H, lq, lk):
    """AR attention (app/transformer.py:72-77): q and k are L2-normalised per head and q is scaled by the head's
    exp(min(scale_mul, ln 100)), so |q.k| <= head_scale[h]. Given that bound the tcgen05 kernel skips its row-maximum pass
    (softmax is shift invariant): same result as the two-pass kernel and the fp32 reference, over ping-pong (<= 256 keys) and
    split-key (> 256 keys) items, from 1 to 100 query rows, with scales from 1 to 30; a head at the clamp (100 > 32) keeps its max pass."""
    D = 64
    g = torch.Generator(device="cpu").manual_seed(lq * 5 + lk)
    Cw = H * D
    hs = torch.linspace(1.0, 30.0, H)
    hs[H - 1] = 100.0                                      # the clamp value (ln 100): this head keeps the max pass
    q = F.normalize(torch.randn(n_seq, lq, H, D, generator=g), dim=-1) * hs.view(1, 1, H, 1)
    k = F.normalize(torch.randn(n_seq, lk, H, D, generator=g), dim=-1)
    q = q.reshape(n_seq, lq, Cw).to(dev(), torch.bfloat16)
    k = k.reshape(n_seq, lk, Cw).to(dev(), torch.bfloat16)
    v = torch.randn(n_seq, lk, Cw, generator=g).to(dev(), torch.bfloat16)
    strides = (lq * Cw, Cw, lk * Cw, Cw, lk * Cw, Cw, lq * Cw, Cw)
    outs = []
    for bound in (hs.to(dev()), None):
        out = torch.full((n_seq, lq, Cw), float("nan"), device=dev(), dtype=torch.bfloat16)
        run_attn(q, k, v, out, n_seq, H, D, lq, lk, strides, 1.0, 0, key_bound=bound)
        _check_attn(out, q, k, v, n_seq, H, D, lq, lk, 1.0, 0, 2e-2)
        outs.append(out.float())
    assert (outs[0] - outs[1]).abs().max().item() < 4e-2        # the two evaluations differ by bf16 ulps of P and of the output (|v| <= 4)


@pytest.mark.parametrize("n_seq,H,lq,lk", [(2, 12, 100, 362), (3, 12, 50, 262), (70, 12, 100, 362), (64, 12, 50, 262), (40, 12, 150, 384),
                                           (1, 1, 7, 257)])
def test_attention_block_wise_kernel(n_seq, H, lq, lk):
    """attn_blk_kernel: 257..384 keys in two key blocks, one item per warpgroup (two item chains per SM); heads with a usable
    bound take it as the softmax shift, the head at the clamp (100) runs an online softmax (O rescaled in TMEM). Same result as
    the split-key kernel within bf16 rounding, and within tolerance of the fp32 reference; from a single item to five items per
    CTA (stage / barrier phases wrap), ragged last block (262 = 144 + 118 keys), two query tiles (150 rows)."""
    D = 64
    g = torch.Generator(device="cpu").manual_seed(lq * 7 + lk + n_seq)
    Cw = H * D
    hs = torch.linspace(1.0, 30.0, H)
    hs[H - 1] = 100.0                                      # the clamp value: this head runs the two blocks as an online softmax
    q = F.normalize(torch.randn(n_seq, lq, H, D, generator=g), dim=-1) * hs.view(1, 1, H, 1)
    k = F.normalize(torch.randn(n_seq, lk, H, D, generator=g), dim=-1)
    q = q.reshape(n_seq, lq, Cw).to(dev(), torch.bfloat16)
    k = k.reshape(n_seq, lk, Cw).to(dev(), torch.bfloat16)
    v = torch.randn(n_seq, lk, Cw, generator=g).to(dev(), torch.bfloat16)
    strides = (lq * Cw, Cw, lk * Cw, Cw, lk * Cw, Cw, lq * Cw, Cw)
    bound = hs.to(dev())
    outs = []
    try:
        for blk in (0, 1):
            _lib.check(_lib.lib().artalk_set_option(b"attn_blk", blk))
            out = torch.full((n_seq, lq, Cw), float("nan"), device=dev(), dtype=torch.bfloat16)
            run_attn(q, k, v, out, n_seq, H, D, lq, lk, strides, 1.0, 0, key_bound=bound)
            _check_attn(out, q, k, v, n_seq, H, D, lq, lk, 1.0, 0, 2e-2)
            outs.append(out.float())
    finally:
        _lib.check(_lib.lib().artalk_set_option(b"attn_blk", 1))
    assert (outs[0] - outs[1]).abs().max().item() < 4e-2


@pytest.mark.parametrize("n_seq,H,lq,lk,split,bounded", [(3, 16, 199, 199, 0, False), (2, 8, 200, 200, 100, False), (2, 8, 100, 100, 0, False),
                                                         (4, 12, 100, 362, 0, True), (3, 12, 50, 262, 0, True), (5, 12, 25, 212, 0, True),
                                                         (40, 12, 1, 182, 0, True), (1, 1, 3, 17, 0, False)])
def test_attention_fp32_grade_on_tensor_cores(n_seq, H, lq, lk, split, bounded):
    """Parity-grade attention (precision bf16x3): fp32 q / k / v split into two bf16 pieces each, both contractions keep the
    three piece products, P split in place over S in TMEM. Against the fp32 reference at 2e-5 absolute (the fp32 SIMT kernel is
    held to the same bound in test_attention; 5e-4 where the scores reach 100); wav2vec / VAE (two-block mask) / AR shapes incl. score bounds
    with one head at the clamp, several items per CTA."""
    D = 64
    g = torch.Generator(device="cpu").manual_seed(lq * 3 + lk + n_seq)
    Cw = H * D
    if bounded:
        hs = torch.linspace(1.0, 30.0, H)
        hs[H - 1] = 100.0
        q = F.normalize(torch.randn(n_seq, lq, H, D, generator=g), dim=-1) * hs.view(1, 1, H, 1)
        k = F.normalize(torch.randn(n_seq, lk, H, D, generator=g), dim=-1)
        q, k = q.reshape(n_seq, lq, Cw), k.reshape(n_seq, lk, Cw)
        scale, bound = 1.0, hs.to(dev())
    else:
        q, k = torch.randn(n_seq, lq, Cw, generator=g), torch.randn(n_seq, lk, Cw, generator=g)
        scale, bound = 0.125, None
    v = torch.randn(n_seq, lk, Cw, generator=g)
    q, k, v = q.to(dev()).contiguous(), k.to(dev()).contiguous(), v.to(dev()).contiguous()
    out = torch.full((n_seq, lq, Cw), float("nan"), device=dev())
    a = _lib.Attn()
    a.key_bound = _lib.ptr(bound)
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.dt = _lib.F32
    a.n_seq, a.n_heads, a.head_dim, a.lq, a.lk = n_seq, H, D, lq, lk
    (a.q_ss, a.q_rs, a.k_ss, a.k_rs, a.v_ss, a.v_rs, a.o_ss, a.o_rs) = (lq * Cw, Cw, lk * Cw, Cw, lk * Cw, Cw, lq * Cw, Cw)
    a.scale, a.split = scale, split
    nbytes = 4 * n_seq * (lq + 2 * lk) * Cw + 4096
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    _lib.check(_lib.lib().artalk_op_attention_split(C.byref(a), scratch.data_ptr(), nbytes, _lib.stream_ptr(dev())))
    torch.cuda.synchronize()
    # two pieces keep 16 mantissa bits per operand: the dropped terms are ~2^-17 of |q||k|, i.e. ~1e-4 in a score of the head at
    # the clamp (|q| = 100) and as much, relative, in its probabilities (measured 2.6e-4 in the output)
    _check_attn(out, q, k, v, n_seq, H, D, lq, lk, scale, split, 5e-4 if bounded else 2e-5)


def _check_attn(out, q, k, v, n_seq, H, D, lq, lk, scale, split, tol):
    Cw = H * D
    qq = q.float().view(n_seq, lq, H, D).transpose(1, 2)
    kk = k.float().view(n_seq, lk, H, D).transpose(1, 2)
    vv = v.float().view(n_seq, lk, H, D).transpose(1, 2)
    s = qq @ kk.transpose(-1, -2) * scale
    if split:
        m = torch.zeros(lq, lk, device=dev())
        m[:split, split:] = -math.inf
        s = s + m
    ref = (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(n_seq, lq, Cw)
    assert (out.float() - ref).abs().max().item() < tol


@pytest.mark.parametrize("cols,act", [(512, 1), (1024, 0), (768, 0), (128, 0)])
def test_layernorm(cols, act):
    g = torch.Generator(device="cpu").manual_seed(cols)
    x = (3 * torch.randn(333, cols, generator=g) + 1.5).to(dev())
    gam, bet = torch.randn(cols, generator=g).to(dev()), torch.randn(cols, generator=g).to(dev())
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 8e-3)):
        out = torch.empty(333, cols, device=dev(), dtype=dt)
        _lib.check(_lib.lib().artalk_op_layernorm(x.data_ptr(), out.data_ptr(), _lib.F32 if dt == torch.float32 else _lib.BF16,
                                                  gam.data_ptr(), bet.data_ptr(), 333, cols, 1e-5, act, _lib.stream_ptr(dev())))
        torch.cuda.synchronize()
        ref = F.layer_norm(x, (cols,), gam, bet, 1e-5)
        if act:
            ref = F.gelu(ref)
        assert ((out.float() - ref).abs() / (1.0 + ref.abs())).max().item() < tol
