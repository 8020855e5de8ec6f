"""Micro-benchmarks of the hot kernels at the DiT shapes (CUDA events, weights rotated across 24 'layers' so each
launch streams its weights from HBM like the real step does)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402


def time_ms(fn, iters):
    """Launches are captured in a CUDA graph so the measurement is GPU-bound (Python launch overhead excluded)."""
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):  # warm up on the capture stream: per-stream workspaces are allocated on first use
            fn(0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=48)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = "cuda"
    L = 24
    D, I = 2048, 5888
    for M in (1920, 640):
        x = torch.randn(M, D, device=dev).bfloat16()
        h = torch.randn(M, I, device=dev).bfloat16()
        res = torch.zeros(M, D, device=dev)
        gate = torch.randn(1, D, device=dev)
        shapes = {
            "qkvg": (4 * D, D), "wo": (D, D), "w13": (2 * I, D), "w2": (D, I),
        }
        for name, (N, K) in shapes.items():
            if a.only and a.only != name:
                continue
            ws = [torch.randn(N, K, device=dev).bfloat16() * K ** -0.5 for _ in range(L)]
            if name == "qkvg":
                outs = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(4)]
                nw = torch.ones(D, device=dev)
                cos = torch.ones(4096, 64, device=dev)
                sin = torch.zeros(4096, 64, device=dev)
                fn = lambda i: ops.gemm_qkv(x, ws[i % L], outs, [nw, nw, None, None], [8, 8, 0, 0], [0, 0, 0, 1], D, cos,
                                            sin, 128, pos_period=640)
            elif name == "w13":
                out = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
                fn = lambda i: ops.gemm_swiglu(x, ws[i % L], out)
            elif name == "wo":
                fn = lambda i: ops.gemm(x, ws[i % L], gate=gate, resid=res, out_f32=res)
            else:
                fn = lambda i: ops.gemm(h, ws[i % L], gate=gate, resid=res, out_f32=res)
            ms = time_ms(fn, a.iters)
            print(f"M={M:5d} {name:5s} N={N:6d} K={K:5d}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)
    # LowRankAdaLN modulate + RMSNorm: alone (X hot in L2 from the previous iteration) and right behind the wo GEMM
    for M in (1920, 640):
        if a.only and a.only != "rmsnorm":
            continue
        res = torch.randn(M, D, device=dev)
        sc, sh = 1 + 0.1 * torch.randn(1, D, device=dev), 0.1 * torch.randn(1, D, device=dev)
        x = torch.randn(M, D, device=dev).bfloat16()
        wo = torch.randn(D, D, device=dev).bfloat16() * D ** -0.5
        gate = torch.randn(1, D, device=dev)
        ms = time_ms(lambda i: ops.rmsnorm_affine(res, sc, sh), a.iters)
        print(f"M={M:5d} rmsnorm_affine alone: {ms*1e3:8.1f} us  {M*D*6/ms/1e6:8.1f} GB/s (algorithmic 6 B/elem)", flush=True)

        def pair(i):
            ops.gemm(x, wo, gate=gate, resid=res, out_f32=res)
            ops.rmsnorm_affine(res, sc, sh)
        ms2 = time_ms(pair, a.iters)
        ms1 = time_ms(lambda i: ops.gemm(x, wo, gate=gate, resid=res, out_f32=res), a.iters)
        print(f"M={M:5d} wo GEMM {ms1*1e3:6.1f} us; wo GEMM + rmsnorm_affine {ms2*1e3:6.1f} us -> rmsnorm in the chain {1e3*(ms2-ms1):6.1f} us", flush=True)
    # attention at the CFG-step shape
    for b, S in ((3, 640), (1, 640)):
        H, Dh = 16, 128
        q = torch.randn(b, S, H, Dh, device=dev).bfloat16()
        k = torch.randn(b, S, H, Dh, device=dev).bfloat16()
        v = torch.randn(b, S, H, Dh, device=dev).bfloat16()
        kt = torch.randn(1, 768, H, Dh, device=dev).bfloat16()
        ks = torch.randn(1, 53, H, Dh, device=dev).bfloat16()
        g = torch.rand(b, S, H * Dh, device=dev).bfloat16()
        out = torch.empty(b, S, H * Dh, device=dev, dtype=torch.bfloat16)
        eff = torch.tensor([36, 0, 36][:b], dtype=torch.int32, device=dev)
        effs = torch.tensor([53, 53, 0][:b], dtype=torch.int32, device=dev)
        segs = [dict(k=k, v=v), dict(k=kt, v=kt, eff_len=eff, batch_mod=1), dict(k=ks, v=ks, eff_len=effs, batch_mod=1)]
        ms = time_ms(lambda i: ops.attention(q, segs, out, gate=g), a.iters)
        keys = [S + 36 + 53, S + 53, S + 36][:b]
        fl = sum(4 * S * kk * Dh * H for kk in keys)
        print(f"attention b={b} S={S}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
