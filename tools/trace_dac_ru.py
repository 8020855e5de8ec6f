"""In-kernel timeline of the fused DAC ResidualUnit (clock64 stamps of epilogue warp 0, third tile of every CTA):
previous tile's end -> phase 1 begins -> conv7 accumulator ready -> phase 1 done (operand of the second MMA in shared
memory) -> second accumulator ready -> phase 2 done. ECHO_DAC_RU_WINDOW=0 traces the ring form."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402


def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to("cuda", dtype)


for C, T in ((96, 1310720), (192, 655360)):
    a = rnd((T, C), 1)
    w7, w1 = rnd((C, 7 * C), 2, (7 * C) ** -0.5), rnd((C, C), 3, C ** -0.5)
    b7, b1 = rnd((C,), 4, 0.1, torch.float32), rnd((C,), 5, 0.1, torch.float32)
    al2, alo = torch.exp(0.3 * rnd((C,), 6, 1, torch.float32)), torch.exp(0.3 * rnd((C,), 7, 1, torch.float32))
    x = rnd((T, C), 8, 1, torch.float32)
    nxt = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    for dil in (1, 9):
        for _ in range(3):
            ops.residual_unit(a, w7, b7, al2, w1, b1, x, alo, nxt, dil)
        torch.cuda.synchronize()
        trace.zero_()
        ops.residual_unit(a, w7, b7, al2, w1, b1, x, alo, nxt, dil, trace=trace)
        torch.cuda.synchronize()
        t = trace.view(148, 16).cpu().double()
        t = t[t[:, 0] > 0]
        ghz = 1.9
        seg = {"prev tile end -> phase 1 begins (barrier)": t[:, 8] - t[:, 14], "wait for the conv7 accumulator": t[:, 9] - t[:, 8],
               "phase 1 (Snake -> bf16 operand in smem)": t[:, 10] - t[:, 9], "wait for the second accumulator": t[:, 11] - t[:, 10],
               "phase 2 (+ x, stores, Snake)": t[:, 13] - t[:, 11], "whole tile": t[:, 13] - t[:, 14]}
        ntile = (T + 127) // 128 / 148
        print(f"C={C} dilation {dil}: CTA lifetime median {((t[:, 7] - t[:, 0]) / ghz / 1e3).median():.1f} us, {ntile:.1f} tiles per CTA")
        for k, v in seg.items():
            v = v / ghz / 1e3
            print(f"    {k:44s} median {v.median():6.2f} us   min {v.min():6.2f}   max {v.max():6.2f}")
