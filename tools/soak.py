"""Soak test on one B200: many requests with changing shapes through one model / DAC pair -- text lengths, speaker
reference lengths (none, 10 s, minutes), batch sizes, step counts, CFG windows, Euler and blockwise samplers, cached
voices -- checking that every result is finite, that a rerun in deterministic mode is bit-identical, and that the
library's workspaces / tensor-map cache survive the churn. Usage: python tools/soak.py [seconds]"""
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import echo_tts_b200  # noqa: E402
from echo_tts_b200 import pipeline as P  # noqa: E402
from echo_tts_b200.autoencoder import ae_decode  # noqa: E402
from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as blockwise  # noqa: E402
from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
from echo_tts_b200.autoencoder import B200DAC, PCAState  # noqa: E402
from echo_tts_b200.config import DacConfig, DitConfig  # noqa: E402
from echo_tts_b200.model import B200EchoDiT  # noqa: E402
from echo_tts_b200.weights import iter_dit_weights, make_dac_weights, make_pca_state  # noqa: E402
cfg, dcfg = DitConfig.base(), DacConfig.base()
model = B200EchoDiT(cfg, dev).load_state_dict(iter_dit_weights(cfg, 1234, include_latent=True))
dac = B200DAC.from_state_dict(make_dac_weights(dcfg, 4321), dcfg, dev)
comps, mean, scale = make_pca_state(dcfg)
pca = PCAState(comps.to(dev), mean.to(dev), scale)
voices = P.VoiceCache(model, dac, pca, max_bytes=1 << 30)
rnd = random.Random(0)
t0 = time.time()
n = 0
peak0 = torch.cuda.memory_allocated()
while time.time() - t0 < budget:
    B = rnd.choice([1, 1, 1, 2, 3, 4])
    S = rnd.choice([640, 640, 160, 96, 333, 37])
    steps = rnd.choice([2, 3, 5, 8])
    Ls = rnd.choice([4, 212, 212, 644, 1600, 6400]) if B == 1 else rnd.choice([4, 212])
    words = " ".join(rnd.choice(["hello", "world", "echo", "b200", "tts", "speaker", "voice", "latent"]) for _ in range(rnd.randint(1, 60)))
    ids, mask = P.get_text_input_ids_and_mask([f"[S1] {words} {i}." for i in range(B)], rnd.choice([64, 256, 768]), device=dev,
                                              pad_to_max=rnd.random() < 0.5)
    g = torch.Generator().manual_seed(n)
    spk = (torch.randn(B, Ls, 80, generator=g) if Ls > 4 else torch.zeros(B, Ls, 80)).to(dev)
    smask = torch.ones(B, Ls, dtype=torch.bool, device=dev) if Ls > 4 else torch.zeros(B, Ls, dtype=torch.bool, device=dev)
    knobs = dict(bench.KNOBS, num_steps=steps, cfg_min_t=rnd.choice([0.0, 0.5, 2.0]))
    if rnd.random() < 0.4:
        knobs.update(speaker_kv_scale=1.5, speaker_kv_min_t=0.9, speaker_kv_max_layers=rnd.choice([None, 12, 24]))
    if rnd.random() < 0.3:
        knobs.update(truncation_factor=0.8, rescale_k=1.2, rescale_sigma=3.0)
    kv = {}
    if B == 1 and Ls > 4 and rnd.random() < 0.5:
        v = voices.get(f"voice-{Ls}-{n % 3}", speaker_latent=spk, speaker_mask=smask)
        spk, smask, kv = v.speaker_latent, v.speaker_mask, dict(speaker_kv_cache=v.kv)
    echo_tts_b200.set_deterministic(True)
    if rnd.random() < 0.35 and S >= 96 and S % 4 == 0:  # the latent prefix is patchified by 4 (the reference reshapes, model.py:459)
        blocks = [S // 3, S // 3, S - 2 * (S // 3)]
        nb = [torch.randn((B, b, 80), generator=g) for b in blocks]
        run = lambda: blockwise(model, spk, smask, ids, mask, 0, blocks, noise_blocks=nb, **knobs, **kv)
        kind = f"blockwise{blocks}"
    else:
        noise = torch.randn(B, S, 80, generator=g)
        run = lambda: sample(model, spk, smask, ids, mask, 0, sequence_length=S, noise=noise, **knobs, **kv)
        kind = "euler"
    a = run()
    b = run()
    assert torch.isfinite(a).all(), (n, kind)
    assert torch.equal(a, b), (n, kind, "deterministic rerun differs")
    echo_tts_b200.set_deterministic(False)
    c = run()
    err = ((c - a).norm() / a.norm()).item()
    assert err < 2e-2, (n, kind, err)
    audio = ae_decode(dac, pca, a[:, : min(a.shape[1], 200)])
    assert torch.isfinite(audio).all() and audio.shape[-1] == min(a.shape[1], 200) * 2048
    n += 1
    if n % 10 == 0:
        torch.cuda.synchronize()
        print(f"{n:4d} requests, {time.time() - t0:5.1f} s, torch allocator {torch.cuda.memory_allocated() / 1e9:.2f} GB, "
              f"last: B={B} S={S} Ls={Ls} Lt={ids.shape[1]} steps={steps} {kind} atomic-vs-deterministic {err:.1e}", flush=True)
torch.cuda.synchronize()
free, total = torch.cuda.mem_get_info()
print(f"soak ok: {n} requests in {time.time() - t0:.1f} s; device memory in use {(total - free) / 1e9:.1f} GB; "
      f"voices cached {len(voices)} ({voices.nbytes / 1e6:.0f} MB, {voices.hits} hits / {voices.misses} misses)")
