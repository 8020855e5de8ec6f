"""One DiT block as the sampler launches it (rmsnorm -> qkvg -> attention -> wo -> rmsnorm -> w13 -> w2), 24 'layers'
with their own weights in one CUDA graph: is the chain slower than the sum of its kernels timed alone
(tools/bench_ops.py)? Kernel-switch costs (instruction cache, PDL hand-over between different kernels) show up here."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
L, D, I, H = 24, 2048, 5888, 16
only = os.environ.get("ONLY", "")


def graph_ms(fn, reps=3):
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for b in (3, 1):
    M = b * 640
    X = torch.randn(M, D, device=dev)
    sc, sh = 1 + 0.1 * torch.randn(1, D, device=dev), 0.1 * torch.randn(1, D, device=dev)
    gate = 0.1 * torch.randn(1, D, device=dev)
    wq = [torch.randn(4 * D, D, device=dev).bfloat16() * D ** -0.5 for _ in range(L)]
    wo = [torch.randn(D, D, device=dev).bfloat16() * D ** -0.5 for _ in range(L)]
    w13 = [torch.randn(2 * I, D, device=dev).bfloat16() * D ** -0.5 for _ in range(L)]
    w2 = [torch.randn(D, I, device=dev).bfloat16() * I ** -0.5 for _ in range(L)]
    outs = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    nw = torch.ones(D, device=dev)
    pos = torch.arange(4096, device=dev)[:, None] * (1e4 ** (-torch.arange(64, device=dev) / 64.0))[None]
    cos, sin = torch.cos(pos).contiguous(), torch.sin(pos).contiguous()
    hh = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
    ao = torch.empty(b, 640, D, device=dev, dtype=torch.bfloat16)
    kt = torch.randn(1, 768, H, 128, device=dev).bfloat16()
    ks = torch.randn(1, 53, H, 128, device=dev).bfloat16()
    eff = torch.tensor([36, 0, 36][:b], dtype=torch.int32, device=dev)
    effs = torch.tensor([53, 53, 0][:b], dtype=torch.int32, device=dev)

    def layer(i, parts):
        if "rms" in parts:
            xn = ops.rmsnorm_affine(X, sc, sh)
        else:
            xn = outs[0]
        if "qkvg" in parts:
            ops.gemm_qkv(xn, wq[i], outs, [nw, nw, None, None], [8, 8, 0, 0], [0, 0, 0, 1], D, cos, sin, 128, pos_period=640)
        if "attn" in parts:
            q, k, v, g = (t.view(b, 640, H, 128) for t in outs)
            segs = [dict(k=k, v=v), dict(k=kt, v=kt, eff_len=eff, batch_mod=1), dict(k=ks, v=ks, eff_len=effs, batch_mod=1)]
            ops.attention(q, segs, ao.view(b, 640, H, 128), gate=g)
        if "wo" in parts:
            ops.gemm(ao.view(M, D), wo[i], gate=gate, resid=X, out_f32=X)
        if "rms" in parts:
            xn = ops.rmsnorm_affine(X, sc, sh)
        if "w13" in parts:
            ops.gemm_swiglu(xn, w13[i], hh)
        if "w2" in parts:
            ops.gemm(hh, w2[i], gate=gate, resid=X, out_f32=X)

    allp = ("rms", "qkvg", "attn", "wo", "w13", "w2")
    total = graph_ms(lambda: [layer(i, allp) for i in range(L)]) / L * 1e3
    print(f"M={M}: full block chain {total:7.1f} us per layer", flush=True)
    s = 0.0
    for part in allp:
        t = graph_ms(lambda: [layer(i, (part,)) for i in range(L)]) / L * 1e3
        s += t
        print(f"   {part:5s} alone {t:7.1f} us" + (" (both norms)" if part == "rms" else ""), flush=True)
    print(f"   sum of parts {s:7.1f} us -> chain / sum = {total / s:.3f}", flush=True)
    for drop in allp:
        t = graph_ms(lambda: [layer(i, tuple(p for p in allp if p != drop)) for i in range(L)]) / L * 1e3
        print(f"   chain without {drop:5s} {t:7.1f} us  (marginal cost of {drop}: {total - t:6.1f} us)", flush=True)
