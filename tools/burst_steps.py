"""Burst timing of a few sampler steps (short calls separated by idle gaps, so the part is not power-capped): what a
kernel-level change is worth before the sustained run's power cap takes its share. Usage: python tools/burst_steps.py"""
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model, dac, pca = bench.load_models(dev, 0, 1)
ids_h, mask_h = bench.tokens(bench.PROMPT)
spk = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1)).to(dev)
smask = torch.ones(1, 212, dtype=torch.bool, device=dev)
ids, mask = ids_h.to(dev), bench.mask_to(mask_h, dev)
noise = torch.randn(1, 640, 80, generator=torch.Generator().manual_seed(5)).to(dev)


def burst(steps, cfg_min_t, reps=15, gap=0.05):
    k = dict(bench.KNOBS)
    k.update(num_steps=steps, cfg_min_t=cfg_min_t)
    ts = []
    for i in range(reps + 3):
        time.sleep(gap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **k)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


for name, cmt in (("CFG", 0.0), ("plain", 2.0)):
    t1, t3 = burst(1, cmt), burst(3, cmt)
    print(f"{name:5s} step in a burst: {(t3 - t1) / 2:6.3f} ms ({(t3 - t1) / 2 / 24 * 1e3:6.1f} us per layer); 1-step call {t1:6.3f} ms", flush=True)
