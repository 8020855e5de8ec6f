"""Three rounds of the four DiT GEMM kinds + the joint attention at M = 3*640 (env M=640 for the non-CFG shapes),
for `ncu --set full -k regex:"gemm_tc|attn_tc" -s 10 -c 5`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
D, I, M = 2048, 5888, int(os.environ.get("M", "1920"))
x = torch.randn(M, D, device=dev).bfloat16()
h = torch.randn(M, I, device=dev).bfloat16()
res = torch.zeros(M, D, device=dev)
gate = torch.randn(1, D, device=dev)
w_qkv = torch.randn(4 * D, D, device=dev).bfloat16() * D ** -0.5
w_13 = torch.randn(2 * I, D, device=dev).bfloat16() * D ** -0.5
w_o = torch.randn(D, D, device=dev).bfloat16() * D ** -0.5
w_2 = torch.randn(D, I, device=dev).bfloat16() * I ** -0.5
outs = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(4)]
nw = torch.ones(D, device=dev)
pos = torch.arange(4096, device=dev)[:, None] * (1e4 ** (-torch.arange(64, device=dev) / 64.0))[None]
cos, sin = torch.cos(pos).contiguous(), torch.sin(pos).contiguous()
hh = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
b, S, H, Dh = M // 640, 640, 16, 128
q = torch.randn(b, S, H, Dh, device=dev).bfloat16()
k = torch.randn(b, S, H, Dh, device=dev).bfloat16()
v = torch.randn(b, S, H, Dh, device=dev).bfloat16()
kt = torch.randn(1, 768, H, Dh, device=dev).bfloat16()
ks = torch.randn(1, 53, H, Dh, device=dev).bfloat16()
g = torch.rand(b, S, H * Dh, device=dev).bfloat16()
ao = torch.empty(b, S, H * Dh, device=dev, dtype=torch.bfloat16)
eff = torch.tensor([36, 0, 36][:b], dtype=torch.int32, device=dev)
effs = torch.tensor([53, 53, 0][:b], dtype=torch.int32, device=dev)
segs = [dict(k=k, v=v), dict(k=kt, v=kt, eff_len=eff, batch_mod=1), dict(k=ks, v=ks, eff_len=effs, batch_mod=1)]
for _ in range(3):
    ops.gemm_qkv(x, w_qkv, outs, [nw, nw, None, None], [8, 8, 0, 0], [0, 0, 0, 1], D, cos, sin, 128, pos_period=640)
    ops.gemm_swiglu(x, w_13, hh)
    ops.gemm(x, w_o, gate=gate, resid=res, out_f32=res)
    ops.gemm(h, w_2, gate=gate, resid=res, out_f32=res)
    ops.attention(q, segs, ao, gate=g)
torch.cuda.synchronize()
print("ok")
