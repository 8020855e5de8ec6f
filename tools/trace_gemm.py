"""In-kernel timeline of the tcgen05 GEMM (clock64 stamps) for the small DiT shapes: where do the fixed costs go?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
names = ["entry", "setup done", "1st TMA issued", "1st full", "last MMA issued", "acc ready", "epilogue done", "exit"]
cases = [(1920, 2048, 2048, 256, 1, 0), (1920, 2048, 2048, 256, 2, 0), (640, 2048, 2048, 128, 1, 0), (640, 2048, 2048, 256, 1, 0),
         (1920, 2048, 5888, 256, 1, 0), (640, 2048, 5888, 128, 1, 0), (1920, 2048, 2048, 256, 1, 8), (1920, 2048, 2048, 128, 1, 0)]
for (M, N, K, bn, cg, dbg) in cases:
    a = torch.randn(M, K, device=dev).bfloat16()
    ws = [torch.randn(N, K, device=dev).bfloat16() * K ** -0.5 for _ in range(8)]
    res = torch.zeros(M, N, device=dev)
    gate = torch.randn(1, N, device=dev)
    trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    for i in range(5):
        ops.gemm(a, ws[i % 8], gate=gate, resid=res, out_f32=res, bn=bn, cg=cg, trace=trace, dbg=dbg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(a, ws[7], gate=gate, resid=res, out_f32=res, bn=bn, cg=cg, trace=trace, dbg=dbg)
    e1.record()
    torch.cuda.synchronize()
    t = trace.view(148, 16).cpu()
    used = t[:, 0] > 0
    t = t[used]
    # leader CTAs (even) have all stamps; report the median over CTAs of each delta from entry, in us at 1.9 GHz
    lead = t[::cg] if cg == 2 else t
    rel = (lead - lead[:, :1]).float() / 1.9e3
    med = rel.median(0).values[:8]
    print(f"M={M} N={N} K={K} bn={bn} cg={cg} dbg={dbg}: {int(used.sum())} CTAs, kernel {e0.elapsed_time(e1)*1e3:.1f} us (event)")
    print("   " + "  ".join(f"{n}={v:.2f}" for n, v in zip(names, med.tolist())))
    span = (t[:, 7].max() - t[:, 0].min()).item() / 1.9e3
    print(f"   first entry -> last exit over all CTAs (per-SM clocks, approx): {span:.2f} us")
