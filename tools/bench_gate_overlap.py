"""Feasibility check (not product code): at M = 640 the fused wq|wk|wv|gate GEMM is 160 tiles = 1.08 waves of the 148
SMs. Would [q|k|v] alone (120 tiles, one wave) followed by attention, with the gate projection (40 tiles) running NEXT TO
the b = 1 attention (80 CTAs) on a second stream, be faster than qkvg -> attention? Chain per layer, CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
L, D, H, M, b = 24, 2048, 16, 640, 1
x = torch.randn(M, D, device=dev).bfloat16()
wq = [torch.randn(4 * D, D, device=dev).bfloat16() * D ** -0.5 for _ in range(L)]
outs = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(4)]
nw = torch.ones(D, device=dev)
pos = torch.arange(4096, device=dev)[:, None] * (1e4 ** (-torch.arange(64, device=dev) / 64.0))[None]
cos, sin = torch.cos(pos).contiguous(), torch.sin(pos).contiguous()
ao = torch.empty(b, 640, D, device=dev, dtype=torch.bfloat16)
kt = torch.randn(1, 768, H, 128, device=dev).bfloat16()
ks = torch.randn(1, 53, H, 128, device=dev).bfloat16()
eff = torch.tensor([36], dtype=torch.int32, device=dev)
effs = torch.tensor([53], dtype=torch.int32, device=dev)
wo = [torch.randn(D, D, device=dev).bfloat16() * D ** -0.5 for _ in range(L)]
X = torch.randn(M, D, device=dev)
gate_mod = 0.1 * torch.randn(1, D, device=dev)
side = torch.cuda.Stream()


def attn(gated):
    q, k, v, g = (t.view(b, 640, H, 128) for t in outs)
    segs = [dict(k=k, v=v), dict(k=kt, v=kt, eff_len=eff, batch_mod=1), dict(k=ks, v=ks, eff_len=effs, batch_mod=1)]
    ops.attention(q, segs, ao.view(b, 640, H, 128), gate=g if gated else None)


def fused(i):
    ops.gemm_qkv(x, wq[i], outs, [nw, nw, None, None], [8, 8, 0, 0], [0, 0, 0, 1], D, cos, sin, 128, pos_period=640)
    attn(True)
    ops.gemm(ao.view(M, D), wo[i], gate=gate_mod, resid=X, out_f32=X)


def split(i, overlap, mul):
    cur = torch.cuda.current_stream()
    ops.gemm_qkv(x, wq[i][: 3 * D], outs[:3], [nw, nw, None], [8, 8, 0], [0, 0, 0], D, cos, sin, 128, pos_period=640)
    if overlap:
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ops.gemm_qkv(x, wq[i][3 * D:], outs[3:], [None], [0], [1], D, cos, sin, 128, pos_period=640)
        attn(False)
        cur.wait_stream(side)
    else:
        ops.gemm_qkv(x, wq[i][3 * D:], outs[3:], [None], [0], [1], D, cos, sin, 128, pos_period=640)
        attn(False)
    if mul:
        ao.mul_(outs[3].view(b, 640, D))  # stand-in for applying the gate somewhere downstream
    ops.gemm(ao.view(M, D), wo[i], gate=gate_mod, resid=X, out_f32=X)


def graph_us(fn):
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(L):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3 / L * 1e3


print(f"qkvg -> gated attention -> wo                                   {graph_us(fused):7.1f} us per layer")
print(f"qkv -> gate -> attention -> wo (serial, gate never applied)      {graph_us(lambda i: split(i, False, False)):7.1f}")
print(f"qkv -> (attention || gate) -> wo (gate never applied)            {graph_us(lambda i: split(i, True, False)):7.1f}")
print(f"qkv -> (attention || gate) -> elementwise gate -> wo             {graph_us(lambda i: split(i, True, True)):7.1f}")
