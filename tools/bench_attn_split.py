"""Times the tcgen05 attention kernel (op-level C ABI) with and without split-KV on the shapes of the hot path:
configs[1] plain step (b=1, S=640), configs[4] plain / CFG steps (S=160 against ~2700 keys). CUDA events over a loop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import ops  # noqa: E402


def rnd(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).to("cuda", torch.bfloat16)


def run(name, b, S, H, D, segs_len, eff=None, iters=200):
    q = rnd((b, S, H, D), 1)
    segs = [dict(k=rnd((b, S, H, D), 2), v=rnd((b, S, H, D), 3))]
    for i, L in enumerate(segs_len):
        sg = dict(k=rnd((1, L, H, D), 10 + i), v=rnd((1, L, H, D), 20 + i), batch_mod=1)
        if eff is not None and eff[i] is not None:
            sg["eff_len"] = torch.tensor(eff[i], dtype=torch.int32, device="cuda")
        segs.append(sg)
    gate = rnd((b, S, H * D), 4)
    out = torch.empty(b, S, H * D, device="cuda", dtype=torch.bfloat16)
    res = []
    for ns in (1, 2, 3, 4, 6, 8, 0):
        for _ in range(3):
            ops.attention(q, segs, out, gate=gate, nsplit=ns)
        torch.cuda.synchronize()
        # a CUDA graph of 20 back-to-back launches: the Python / ctypes call costs more than the kernel
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                ops.attention(q, segs, out, gate=gate, nsplit=ns)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters // 20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res.append(f"nsplit={ns if ns else 'auto'}: {1e3 * e0.elapsed_time(e1) / (iters // 20 * 20):.1f} us")
    print(f"{name}: " + "  ".join(res), flush=True)


run("configs[1] plain  b=1 S=640 text 768(eff 35) spk 53", 1, 640, 16, 128, [768, 53], eff=[[35], None])
run("configs[1] CFG    b=3 S=640 text 768(eff 35,0,35) spk 53", 3, 640, 16, 128, [768, 53], eff=[[35, 0, 35], [53, 53, 0]])
run("configs[4] plain  b=1 S=160 lat 160 text 768(eff 35) spk 1600", 1, 160, 16, 128, [160, 768, 1600], eff=[None, [35], None])
run("configs[4] CFG    b=3 S=160 lat 160 text 768 spk 1600", 3, 160, 16, 128, [160, 768, 1600], eff=[None, [35, 0, 35], [1600, 1600, 0]])
