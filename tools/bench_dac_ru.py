"""DAC ResidualUnit at the decoder's real shapes (T = 640 latents): the fused kernel (conv7 -> Snake -> bf16 in shared memory
-> conv1 -> + x, one launch) against the two tap-GEMM launches it replaces. CUDA events over a loop; algorithmic DRAM bytes
per unit: fused = read a (2 B) + read / write the fp32 stream (8 B) + write the next input (2 B) = 12 B per element,
two launches = 16 B (the bf16 intermediate goes out and comes back)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402
from echo_tts_b200._lib import ACT_SNAKE  # noqa: E402


def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to("cuda", dtype)


for C, T in ((96, 1310720), (192, 655360)):
    a = rnd((T, C), 1)
    w7, w1 = rnd((C, 7 * C), 2, (7 * C) ** -0.5), rnd((C, C), 3, C ** -0.5)
    b7, b1 = rnd((C,), 4, 0.1, torch.float32), rnd((C,), 5, 0.1, torch.float32)
    al2, alo = torch.exp(0.3 * rnd((C,), 6, 1, torch.float32)), torch.exp(0.3 * rnd((C,), 7, 1, torch.float32))
    x = rnd((T, C), 8, 1, torch.float32)
    hb = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    nxt = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    for dil in (1, 3, 9):
        shifts = [-(6 - j) * dil for j in range(7)]

        def two():
            ops.gemm(a, w7, taps=7, tap_shift=shifts, bias=b7, out_bf16=hb, act=ACT_SNAKE, alpha=al2, col_mod=C)
            ops.gemm(hb, w1, bias=b1, resid=x, out_f32=x, out_bf16=nxt, act=ACT_SNAKE, alpha=alo, col_mod=C)

        def fused():
            ops.residual_unit(a, w7, b7, al2, w1, b1, x, alo, nxt, dil)

        res = []
        for name, fn in (("two launches", two), ("fused", fused)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / 10
            gb = T * C * (16 if name == "two launches" else 12) / 1e9
            res.append(f"{name} {us:7.1f} us ({gb / us * 1e6 / 1e3:.2f} TB/s algorithmic)")
        fl = 2.0 * T * C * C * 8
        print(f"C={C} T={T} dilation {dil}: " + "; ".join(res) + f"; {fl / 1e9:.0f} GFLOP", flush=True)
