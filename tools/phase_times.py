"""Where does one configs[1] request spend its time? CUDA-event timing of the phases through the public Python API:
KV caches, the sampler with all-CFG / no-CFG / default step mixes (=> per-step cost of each kind), DAC decode."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echo_tts_b200.autoencoder import ae_decode  # noqa: E402
from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model, dac, pca = bench.load_models(dev, 0, 1)
ids_h, mask_h = bench.tokens(bench.PROMPT)
spk = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1)).to(dev)
smask = torch.ones(1, 212, dtype=torch.bool, device=dev)
ids, mask = ids_h.to(dev), bench.mask_to(mask_h, dev)
noise = torch.randn(1, 640, 80, generator=torch.Generator().manual_seed(5)).to(dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def knobs(**kw):
    k = dict(bench.KNOBS)
    k.update(kw)
    return k


lat = sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **bench.KNOBS)
rows = [
    ("get_kv_cache_text (768 tokens)", lambda: model.get_kv_cache_text(ids, mask)),
    ("get_kv_cache_speaker (212 latents)", lambda: model.get_kv_cache_speaker(spk.bfloat16())),
    ("sampler, default (20 CFG + 20 plain steps)", lambda: sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **bench.KNOBS)),
    ("sampler, 40 CFG steps (cfg_min_t=0)", lambda: sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **knobs(cfg_min_t=0.0))),
    ("sampler, 40 plain steps (cfg_min_t=2)", lambda: sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **knobs(cfg_min_t=2.0))),
    ("sampler, 1 step", lambda: sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **knobs(num_steps=1))),
    ("ae_decode (640 latents)", lambda: ae_decode(dac, pca, lat)),
]
res = {}
for name, fn in rows:
    res[name] = timed(fn)
    print(f"{name:48s} {res[name]:8.3f} ms", flush=True)
cfg = res["sampler, 40 CFG steps (cfg_min_t=0)"]
plain = res["sampler, 40 plain steps (cfg_min_t=2)"]
one = res["sampler, 1 step"]
print(f"per CFG step   ~ {(cfg - one) / 39:6.3f} ms  (24 layers: {(cfg - one) / 39 / 24 * 1e3:6.1f} us per layer)")
print(f"per plain step ~ {(plain - one) / 39:6.3f} ms  (24 layers: {(plain - one) / 39 / 24 * 1e3:6.1f} us per layer)")
print(f"sampler fixed cost (KV caches, AdaLN tables, 1 step incl.) ~ {one:6.3f} ms")
