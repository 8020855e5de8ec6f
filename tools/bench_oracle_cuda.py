"""Same-box GPU CONTEXT number (not the target, not a product path): the oracle port -- plain eager PyTorch, the same
operator sequence as the reference's model.py / inference.py / autoencoder.py -- on the B200 in bf16 through
cuBLAS / SDPA, one full configs[1] request (KV caches + 40 Euler steps with CFG + DAC decode). SURVEY 2.1 / BASELINE.md
section 1 name "the reference's PyTorch code on the B200" as the same-box bar; the reference itself cannot travel to
the GPU box, so its restatement stands in. Usage: python tools/bench_oracle_cuda.py [--tiny] [--cpu] [--runs N]"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echo_tts_b200.config import DacConfig, DitConfig  # noqa: E402
from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state  # noqa: E402
from oracle import echo_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tiny", action="store_true")
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--runs", type=int, default=3)
a = ap.parse_args()
dev = "cpu" if a.cpu else "cuda"
dt = torch.bfloat16
torch.set_default_device(dev)  # the oracle creates its tables / schedules without a device argument


def linear(x, w, b=None):  # activations follow the weight dtype, as nn.Linear modules in a bf16 model do
    return F.linear(x.to(w.dtype), w, None if b is None else b.to(w.dtype))


def masked_attention(q, k, v, mask):  # the reference calls F.scaled_dot_product_attention (model.py:148-154, 255-261)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2).to(q.dtype), v.transpose(1, 2).to(q.dtype),
                                       attn_mask=mask)
    return o.transpose(1, 2)


O.linear = linear
O.masked_attention = masked_attention
cfg, dcfg = (DitConfig.tiny(), DacConfig.tiny()) if a.tiny else (DitConfig.base(), DacConfig.base())
S, Lt, Ls = (16, 48, 16) if a.tiny else (640, 768, 212)
with torch.device("cpu"):
    sd = make_dit_weights(cfg, 1234, include_latent=False)
    dsd = make_dac_weights(dcfg, 4321)
    comps, mean, scale = make_pca_state(dcfg)
    ids, mask = bench.tokens(bench.PROMPT, Lt)
    spk = torch.randn(1, Ls, 80, generator=torch.Generator().manual_seed(1))
    noise = torch.randn(1, S, 80, generator=torch.Generator().manual_seed(0))
sd = {k: v.to(dev, dt) for k, v in sd.items()}
dsd = {k: v.to(dev) for k, v in dsd.items()}  # DAC stays fp32 (TF32 tensor cores): the oracle's conv stack is not dtype-generic
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
comps, mean = comps.to(dev), mean.to(dev)
ids, mask, spk, noise = ids.to(dev), mask.to(dev), spk.to(dev, dt), noise.to(dev)
smask = torch.ones(1, Ls, dtype=torch.bool, device=dev)


def sync():
    if dev == "cuda":
        torch.cuda.synchronize()


times = []
with torch.inference_mode():
    for r in range(a.runs + 1):
        sync()
        t0 = time.perf_counter()
        lat = O.sample_euler_cfg_independent_guidances(sd, cfg, spk, smask, ids, mask, noise, t_dtype=dt, **bench.KNOBS)
        sync()
        t1 = time.perf_counter()
        audio = O.ae_decode(dsd, dcfg, comps, mean, scale, lat.float())
        sync()
        t2 = time.perf_counter()
        assert torch.isfinite(audio.float()).all()
        if r > 0:
            times.append((t1 - t0, t2 - t1))
        print(f"run {r}: sampler {1e3 * (t1 - t0):.1f} ms, DAC decode {1e3 * (t2 - t1):.1f} ms", flush=True)
smp = sorted(t[0] for t in times)[len(times) // 2]
dec = sorted(t[1] for t in times)[len(times) // 2]
secs = S * 2048 / 44100.0
print(f"eager PyTorch ({dt}, {dev}{', ' + torch.cuda.get_device_name(0) if dev == 'cuda' else ''}) oracle port, one request (DiT bf16, DAC fp32/TF32): sampler {1e3 * smp:.1f} ms + "
      f"decode {1e3 * dec:.1f} ms = {1e3 * (smp + dec):.1f} ms -> {secs / (smp + dec):.1f} audio-s/s")
