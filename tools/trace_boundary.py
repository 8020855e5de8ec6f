"""Kernel-boundary latency in a PDL chain, from %globaltimer stamps inside the GEMM kernel: when did the last CTA of
kernel N finish its work, and when did the dependency wait of kernel N+1 return? (wo -> w13 and w13 -> w2.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
D, I = 2048, 5888
for M in (640, 1920):
    x = torch.randn(M, D, device=dev).bfloat16()
    ao = torch.randn(M, D, device=dev).bfloat16()
    hh = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
    X = torch.zeros(M, D, device=dev)
    gate = torch.randn(1, D, device=dev)
    w_o = torch.randn(D, D, device=dev).bfloat16() * D ** -0.5
    w_13 = torch.randn(2 * I, D, device=dev).bfloat16() * D ** -0.5
    w_2 = torch.randn(D, I, device=dev).bfloat16() * I ** -0.5
    tr = [torch.zeros(148 * 16, dtype=torch.int64, device=dev) for _ in range(3)]

    def chain(traces):
        ops.gemm(ao, w_o, gate=gate, resid=X, out_f32=X, trace=traces[0])
        ops.gemm_swiglu(x, w_13, hh, trace=traces[1])
        ops.gemm(hh, w_2, gate=gate, resid=X, out_f32=X, trace=traces[2])

    for _ in range(5):
        chain(tr)
    torch.cuda.synchronize()
    for t in tr:
        t.zero_()
    torch.cuda.synchronize()
    chain(tr)
    torch.cuda.synchronize()
    T = [t.view(148, 16).cpu() for t in tr]
    T = [t[t[:, 0] > 0] for t in T]
    names = ["wo", "w13", "w2"]
    for a in range(2):
        done = T[a][:, 14]
        start = T[a + 1][:, 13]
        print(f"M={M} {names[a]:3s} -> {names[a+1]:3s}: last CTA of {names[a]} done -> first / median / last wait-return of {names[a+1]}: "
              f"{(start.min() - done.max()) / 1e3:6.2f} / {(start.median() - done.max()) / 1e3:6.2f} / {(start.max() - done.max()) / 1e3:6.2f} us; "
              f"{names[a]} CTAs finish over {(done.max() - done.min()) / 1e3:5.2f} us (median {(done.max() - done.median()) / 1e3:5.2f} us before the last)")
    for a in range(3):
        print(f"      {names[a]:3s}: wait-return -> CTA done, median {((T[a][:, 14] - T[a][:, 13]).float().median()) / 1e3:6.2f} us, "
              f"kernel span (first wait-return -> last done) {(T[a][:, 14].max() - T[a][:, 13].min()) / 1e3:6.2f} us")
