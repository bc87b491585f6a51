"""Developer tool: per-shape timing of the path's GEMM / attention launches through the C ABI op entry points.

Every launch of a shape uses its own copy of the operands (rotating pool larger than the 126 MB L2), the launches of one
shape are captured into a CUDA graph (no host launch overhead in the numbers) and timed with CUDA events.
  python tools_opbench.py [--only gemm|attn] [--clips 64] [--chunks 3] [--iters 24]
Shapes = what bench.py's default workload launches."""
import argparse, ctypes as C, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--only", default=""); ap.add_argument("--clips", type=int, default=64); ap.add_argument("--chunks", type=int, default=3)
ap.add_argument("--iters", type=int, default=24); ap.add_argument("--pool-mb", type=int, default=400)
ap.add_argument("--filter", default="", help="substring of the shape label")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = _lib.lib()
B, NCH = a.clips, a.chunks
bf = torch.bfloat16


def rm(rpb=0, bs=0, rs=0):
    return _lib.RowMap(rpb, bs, rs)


def time_graph(launch, copies, iters):
    """launch(i) enqueues one op on the current stream using operand copy i % copies."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(min(copies, 3)):
            launch(i)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(iters):
                launch(i % copies)
        g.replay(); s.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); g.replay(); e1.record(s); s.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    return best


def bench_gemm(label, M, N, K, act=0, resid=False, gate=False, out="bf16", dual=False):
    per = M * K * 2 + N * K * 2 + M * N * (4 if (out == "f32" or resid) else 2) + (M * N * 2 if gate else 0)
    copies = max(2, min(a.iters, int(math.ceil(a.pool_mb * 1e6 / per))))
    As = [torch.randn(M, K, device=dev).to(bf) for _ in range(copies)]
    Ws = [(torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf) for _ in range(copies)]
    bias = torch.randn(N, device=dev)
    X = [torch.randn(M, N, device=dev) for _ in range(copies)] if (resid or out == "f32") else None
    O = [torch.empty(M, N, device=dev, dtype=bf) for _ in range(copies)] if (out == "bf16" or dual) else None
    G = [torch.randn(M, N, device=dev).to(bf) for _ in range(copies)] if gate else None
    structs = []
    for i in range(copies):
        g = _lib.Gemm()
        g.A, g.W, g.a_map, g.ldw, g.M, g.N, g.K = As[i].data_ptr(), Ws[i].data_ptr(), rm(0, 0, K), K, M, N, K
        g.groups = 1
        g.bias, g.act = bias.data_ptr(), act
        g.gate = G[i].data_ptr() if gate else None
        g.gate_dt, g.gate_map = _lib.BF16, rm(0, 0, N)
        g.resid = X[i].data_ptr() if resid else None
        g.resid_map = rm(0, 0, N)
        g.out32 = X[i].data_ptr() if X is not None else None
        g.out_act = O[i].data_ptr() if O is not None else None
        g.out_act_dt = _lib.BF16
        g.c_map = rm(0, 0, N)
        structs.append(g)

    def launch(i):
        _lib.check(lib.artalk_op_gemm(C.byref(structs[i]), 1, torch.cuda.current_stream().cuda_stream))
    us = time_graph(launch, copies, a.iters)
    print("gemm %-34s M=%6d N=%6d K=%5d  %8.1f us  %7.0f TFLOP/s" % (label, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)
    return us


def bench_attn(label, n_seq, H, lq, lk, split=0, fused_qkv=False):
    Cw = H * 64
    per = n_seq * (lq + 2 * lk) * Cw * 2 + n_seq * lq * Cw * 2
    copies = max(2, min(a.iters, int(math.ceil(a.pool_mb * 1e6 / per))))
    if fused_qkv:      # wav2vec / VAE layout: q|k|v interleaved per row
        QKV = [torch.randn(n_seq, lq, 3 * Cw, device=dev).to(bf) for _ in range(copies)]
    else:
        Q = [torch.randn(n_seq, lq, Cw, device=dev).to(bf) for _ in range(copies)]
        Kc = [torch.randn(n_seq, 362, Cw, device=dev).to(bf) for _ in range(copies)]
        Vc = [torch.randn(n_seq, 362, Cw, device=dev).to(bf) for _ in range(copies)]
    O = [torch.empty(n_seq, lq, Cw, device=dev, dtype=bf) for _ in range(copies)]
    structs = []
    for i in range(copies):
        t = _lib.Attn()
        if fused_qkv:
            base = QKV[i].data_ptr()
            t.q, t.k, t.v = base, base + Cw * 2, base + 2 * Cw * 2
            t.q_ss = t.k_ss = t.v_ss = lq * 3 * Cw; t.q_rs = t.k_rs = t.v_rs = 3 * Cw
        else:
            t.q, t.k, t.v = Q[i].data_ptr(), Kc[i].data_ptr(), Vc[i].data_ptr()
            t.q_ss, t.q_rs, t.k_ss, t.k_rs, t.v_ss, t.v_rs = lq * Cw, Cw, 362 * Cw, Cw, 362 * Cw, Cw
        t.out, t.o_ss, t.o_rs = O[i].data_ptr(), lq * Cw, Cw
        t.dt, t.n_seq, t.n_heads, t.head_dim, t.lq, t.lk, t.scale, t.split = _lib.BF16, n_seq, H, 64, lq, lk, 0.125, split
        structs.append(t)

    def launch(i):
        _lib.check(lib.artalk_op_attention(C.byref(structs[i]), torch.cuda.current_stream().cuda_stream))
    us = time_graph(launch, copies, a.iters)
    fl = 4.0 * n_seq * H * 64 * lq * lk
    print("attn %-34s seq=%4d H=%2d lq=%3d lk=%3d  %8.1f us  %7.0f TFLOP/s" % (label, n_seq, H, lq, lk, us, fl / us / 1e6), flush=True)
    return us


tot = 0.0
want = lambda lab: (a.filter in lab)
if a.only in ("", "gemm"):
    Mw = B * NCH * 199
    w2v = 0.0
    for lab, args in [("w2v qkv", dict(M=Mw, N=3072, K=1024)), ("w2v out+resid", dict(M=Mw, N=1024, K=1024, resid=True, out="f32")),
                      ("w2v ff1 gelu", dict(M=Mw, N=4096, K=1024, act=1)), ("w2v ff2+resid", dict(M=Mw, N=1024, K=4096, resid=True, out="f32"))]:
        if want(lab):
            w2v += 24 * bench_gemm(lab, **args)
    print("  -> wav2vec layer GEMMs per step: %.2f ms" % (w2v / 1e3))
    if want("ada"):
        tot += NCH * bench_gemm("ar ada (hoisted AdaLN)", B * 181, 56832, 1024)
    ar = 0.0
    for n_new in (1, 5, 25, 50, 100):
        M = B * n_new
        for lab, args in [("ar qkv", dict(M=M, N=2304, K=768)), ("ar proj (fp32 y)", dict(M=M, N=768, K=768, out="f32")),
                          ("ar ff1 gelu", dict(M=M, N=3072, K=768, act=2)), ("ar ff2 (fp32 y)", dict(M=M, N=768, K=3072, out="f32"))]:
            if want(lab):
                ar += 12 * NCH * bench_gemm("%s n=%d" % (lab, n_new), **args)
    print("  -> AR block GEMMs per step: %.2f ms" % (ar / 1e3))
    vae = 0.0
    for rows in (200, 100):
        M = B * rows
        for lab, args in [("vae qkv", dict(M=M, N=1536, K=512)), ("vae out+resid", dict(M=M, N=512, K=512, resid=True, out="f32", dual=True)),
                          ("vae ff1", dict(M=M, N=768, K=512, act=2)), ("vae ff2+resid", dict(M=M, N=512, K=768, resid=True, out="f32", dual=True))]:
            if want(lab):
                vae += 8 * NCH * bench_gemm("%s rows=%d" % (lab, rows), **args)
    print("  -> VAE GEMMs per step: %.2f ms" % (vae / 1e3))
if a.only in ("", "attn"):
    t = 0.0
    if want("w2v"):
        t += 24 * bench_attn("w2v", B * NCH, 16, 199, 199, fused_qkv=True)
    for n_new, lk in ((1, 182), (5, 187), (25, 212), (50, 262), (100, 362)):
        if want("ar"):
            t += 12 * NCH * bench_attn("ar n=%d" % n_new, B, 12, n_new, lk)
    if want("vae"):
        t += 8 * NCH * bench_attn("vae dec", B, 8, 200, 200, split=100, fused_qkv=True)
        t += 8 * NCH * bench_attn("vae enc", B, 8, 100, 100, fused_qkv=True)
    print("  -> attention per step: %.2f ms" % (t / 1e3))
