"""Stress of the fused DAC ResidualUnit (window and ring forms) and the lean conv epilogue: random lengths / dilations,
bit equality against the two-launch form every time and run-to-run (a race on the shared-memory operand, the aliased
transpose patches or the barriers would show up as a flipped bit somewhere in a few hundred launches)."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402
from echo_tts_b200._lib import ACT_SNAKE  # noqa: E402


def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to("cuda", dtype)


random.seed(7)
bad = 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for it in range(n):
    C = random.choice((96, 96, 192))
    T = random.choice((random.randint(1, 300), random.randint(300, 40000), 128 * 148 * random.randint(1, 6) + random.randint(-130, 130)))
    dil = random.choice((1, 3, 9))
    a = rnd((T, C), 10 * it + 1)
    w7, w1 = rnd((C, 7 * C), 10 * it + 2, (7 * C) ** -0.5), rnd((C, C), 10 * it + 3, C ** -0.5)
    b7, b1 = rnd((C,), 10 * it + 4, 0.1, torch.float32), rnd((C,), 10 * it + 5, 0.1, torch.float32)
    al2, alo = torch.exp(0.3 * rnd((C,), 10 * it + 6, 1, torch.float32)), torch.exp(0.3 * rnd((C,), 10 * it + 7, 1, torch.float32))
    x = rnd((T, C), 10 * it + 8, 1, torch.float32)
    shifts = [-(6 - j) * dil for j in range(7)]
    hb = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    x2, n2 = x.clone(), torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w7, taps=7, tap_shift=shifts, bias=b7, out_bf16=hb, act=ACT_SNAKE, alpha=al2, col_mod=C)
    ops.gemm(hb, w1, bias=b1, resid=x2, out_f32=x2, out_bf16=n2, act=ACT_SNAKE, alpha=alo, col_mod=C)
    ok = True
    for rep in range(3):
        x1, n1 = x.clone(), torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
        ops.residual_unit(a, w7, b7, al2, w1, b1, x1, alo, n1, dil)
        ok = ok and torch.equal(x1, x2) and torch.equal(n1, n2)
    if not ok:
        bad += 1
        print(f"MISMATCH C={C} T={T} dilation {dil}", flush=True)
print(f"{n} random cases x 3 runs: {bad} mismatches")
sys.exit(1 if bad else 0)
