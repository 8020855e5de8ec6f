"""Where the exposed epilogue of the QKV GEMM goes: per-CTA last-tile epilogue time (in-kernel clock64) with rope + norm +
sigmoid (the real DiT call), without the sigmoid, without the rope, norm only, plain stores."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
from echo_tts_b200 import ops
dev = "cuda"; D = 2048
for M in (640, 1920):
    x = torch.randn(M, D, device=dev).bfloat16()
    w = torch.randn(4 * D, D, device=dev).bfloat16() * D ** -0.5
    outs = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    nw = torch.ones(D, device=dev)
    cos, sin = torch.ones(4096, 64, device=dev), torch.zeros(4096, 64, device=dev)
    for label, norms, ropes, sig in (("rope+norm+sigmoid (real)", [nw, nw, None, None], [8, 8, 0, 0], [0, 0, 0, 1]),
                                     ("rope+norm, no sigmoid", [nw, nw, None, None], [8, 8, 0, 0], [0, 0, 0, 0]),
                                     ("norm+sigmoid, no rope", [nw, nw, None, None], [0, 0, 0, 0], [0, 0, 0, 1]),
                                     ("norm only", [nw, nw, None, None], [0, 0, 0, 0], [0, 0, 0, 0]),
                                     ("plain stores", [None] * 4, [0, 0, 0, 0], [0, 0, 0, 0])):
        trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
        fn = lambda: ops.gemm_qkv(x, w, outs, norms, ropes, sig, D, cos, sin, 128, pos_period=640, trace=trace)
        for _ in range(4): fn()
        torch.cuda.synchronize(); trace.zero_(); fn(); torch.cuda.synchronize()
        t = trace.view(148, 16).cpu(); t = t[t[:, 0] > 0]
        epi = (t[:, 6] - t[:, 5]).float() / 1.9e3
        tot = (t[:, 7] - t[:, 0]).float() / 1.9e3
        print(f"M={M} {label:26s}: last-tile epilogue per CTA: median {epi.median():.2f} max {epi.max():.2f} min {epi.min():.2f} us; CTA lifetime max {tot.max():.2f} us")
