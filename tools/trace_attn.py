"""In-kernel timeline of the tcgen05 attention kernel (clock64 stamps of the first softmax warp of every CTA)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402

dev = "cuda"
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 1  # split-KV shares (1 = off)
for b in (1, 3):
    S, H, Dh = 640, 16, 128
    if NS == -2 and b == 3:
        continue  # two softmax warpgroups: one CTA per SM, the b = 1 shape
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
    ncta = 5 * H * b * max(NS, 1)
    trace = torch.zeros(ncta * 64, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.attention(q, segs, out, gate=g, trace=trace, nsplit=NS)
    torch.cuda.synchronize()
    trace.zero_()
    ops.attention(q, segs, out, gate=g, trace=trace, nsplit=NS)
    torch.cuda.synchronize()
    t = trace.view(ncta, 64).cpu()
    rel = (t - t[:, :1]).float() / 1.9e3  # stamps are taken by thread 64 (first softmax warp)
    rel[t == 0] = float("nan")
    med = rel.nanmedian(0).values
    print(f"b={b}: {ncta} CTAs; us @1.9GHz: setup={med[1]:.2f} tiles={med[2]:.2f} q_landed={med[3]:.2f} O_done={med[30]:.2f} end={med[31]:.2f}")
    print(f"   tile 5 of the softmax warp: scores ready={med[14]:.3f} ld done={med[27]:.3f} max done={med[28]:.3f} rescale check done={med[29]:.3f} "
          f"exp+pack done={med[61]:.3f} st done={med[62]:.3f} P published={med[15]:.3f}; tile 6 scores ready={med[16]:.3f}")
    if NS > 1:
        print(f"   split-KV x{NS}: slab staged={med[24]:.2f} cluster barrier passed={med[25]:.2f} merged={med[26]:.2f}")
    print("   softmax [start -> end] per tile: " + "  ".join(f"[{med[4+2*j]:.2f}->{med[5+2*j]:.2f}]" for j in range(13) if med[4+2*j] == med[4+2*j]))
    print("   MMA thread S_j issued: " + "  ".join(f"{med[32+j]:.2f}" for j in range(13) if med[32+j] == med[32+j]))
    print("   MMA thread PV_j issued: " + "  ".join(f"{med[48+j]:.2f}" for j in range(13) if med[48+j] == med[48+j]))
    span = (t[:, 31].max() - t[:, 0].min()).item() / 1.9e3
    print(f"   first entry -> last end over all CTAs (different SM clocks, approximate): {span:.2f} us", flush=True)
