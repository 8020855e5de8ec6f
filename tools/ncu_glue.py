"""The bandwidth-bound glue of a sampler step at its real shapes, three rounds, for
`ncu --set full -k regex:"rmsnorm|cfg_euler" -s 8 -c 4`: LowRankAdaLN modulate + RMSNorm at M = 1920 / 640 (fp32 residual
stream in, bf16 GEMM operand out: 6 B per element algorithmic) and the CFG combine + Euler update of the latents."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
D = 2048
xs = {M: torch.randn(M, D, device=dev) for M in (1920, 640)}
sc, sh = 1 + 0.1 * torch.randn(1, D, device=dev), 0.1 * torch.randn(1, D, device=dev)
x = torch.randn(1, 640, 80, device=dev)
v3, v1 = torch.randn(3, 640, 80, device=dev), torch.randn(1, 640, 80, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()  # 256 MB > L2: the captures below see a cold cache, as ncu's replay does anyway
    ops.rmsnorm_affine(xs[1920], sc, sh)
    ops.rmsnorm_affine(xs[640], sc, sh)
    ops.cfg_euler_update(x, v3, True, 3.0, 8.0, -0.025)
    ops.cfg_euler_update(x, v1, False, 3.0, 8.0, -0.025)
torch.cuda.synchronize()
print("ok")
