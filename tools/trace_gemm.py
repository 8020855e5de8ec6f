"""In-kernel timeline of the tcgen05 GEMM (clock64 stamps): where do the fixed costs go?
Slots: 0 entry, 1 setup done, 2 1st TMA issued, 3 1st full, 4 last MMA issued, 5 last acc ready, 6 last epilogue
done, 7 exit, 8+2i / 9+2i = accumulator ready / epilogue done of the CTA's i-th tile (i < 4)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
names = ["entry", "setup", "tma0", "full0", "mma_last", "acc_last", "epi_last", "exit"]
D, I = 2048, 5888


def run(label, M, fn_factory):
    trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    fn = fn_factory(trace)
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    t = trace.view(148, 16).cpu()
    t = t[t[:, 0] > 0]
    rel = (t - t[:, :1]).float() / 1.9e3
    rel[t == 0] = float("nan")
    med = rel.nanmedian(0).values
    print(f"{label} M={M}: {t.shape[0]} CTAs, event {e0.elapsed_time(e1)*1e3:.1f} us")
    print("   " + "  ".join(f"{n}={v:.2f}" for n, v in zip(names, med[:8].tolist())))
    print("   tiles (acc ready -> epilogue done): " + "  ".join(
        f"[{med[8 + 2 * i]:.2f} -> {med[9 + 2 * i]:.2f}]" for i in range(4) if med[9 + 2 * i] == med[9 + 2 * i]))
    # slot 12 is the producer's wait-return stamp, but a CTA's third tile overwrites it (tile stamps use slots 8 .. 15)
    if med[12] == med[12] and med[13] != med[13]:
        print(f"   producer: dependency wait returned {med[12]:.2f}, A tiles of the first fill issued {med[2]:.2f}, first stage full {med[3]:.2f}")
    print(f"   slowest CTA exit: {rel[:, 7].max():.2f} us", flush=True)


for M in (1920, 640):
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
    cos, sin = torch.ones(4096, 64, device=dev), torch.zeros(4096, 64, device=dev)
    hh = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
    run("qkvg", M, lambda tr: (lambda: ops.gemm_qkv(x, w_qkv, outs, [nw, nw, None, None], [8, 8, 0, 0], [0, 0, 0, 1], D,
                                                    cos, sin, 128, pos_period=640, trace=tr)))
    run("w13 ", M, lambda tr: (lambda: ops.gemm_swiglu(x, w_13, hh, trace=tr)))
    run("wo  ", M, lambda tr: (lambda: ops.gemm(x, w_o, gate=gate, resid=res, out_f32=res, trace=tr)))
    run("w2  ", M, lambda tr: (lambda: ops.gemm(h, w_2, gate=gate, resid=res, out_f32=res, trace=tr)))
