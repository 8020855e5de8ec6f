"""BASELINE configs[4] on B200: sample_blockwise 4 x 160 with a 5-minute speaker reference (1600 speaker-KV patches),
speaker_kv_scale 1.5 until t < 0.9, streaming decode. Reports time-to-first-audio (block 0 decoded) and total."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echo_tts_b200 import pipeline as P  # noqa: E402
from echo_tts_b200.autoencoder import B200DAC, PCAState  # noqa: E402
from echo_tts_b200.config import DacConfig, DitConfig  # noqa: E402
from echo_tts_b200.model import B200EchoDiT  # noqa: E402
from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as blockwise  # noqa: E402
from echo_tts_b200.weights import iter_dit_weights, make_dac_weights, make_pca_state  # noqa: E402

dev = torch.device("cuda", 0)
cfg, dcfg = DitConfig.base(), DacConfig.base()
model = B200EchoDiT(cfg, dev).load_state_dict(iter_dit_weights(cfg, 1234, include_latent=True))
dac = B200DAC.from_state_dict(make_dac_weights(dcfg, 4321), dcfg, dev)
comps, mean, scale = make_pca_state(dcfg)
pca = PCAState(comps.to(dev), mean.to(dev), scale)
ids, mask = bench.tokens(bench.PROMPT)
ids, mask = ids.to(dev), mask.to(dev)
spk = torch.randn(1, 6400, 80, generator=torch.Generator().manual_seed(1)).to(dev)
smask = torch.ones(1, 6400, dtype=torch.bool, device=dev)
knobs = dict(bench.KNOBS, speaker_kv_scale=1.5, speaker_kv_min_t=0.9, speaker_kv_max_layers=24)
cache = P.VoiceCache(model, dac, pca)
voice = cache.get("five-minute-voice", speaker_latent=spk, speaker_mask=smask)  # per-voice persistence
print(f"cached voice: {voice.nbytes / 1e6:.0f} MB (latents + mask + 24-layer speaker KV)", flush=True)
for it in range(7):
    kv = dict(speaker_kv_cache=voice.kv) if it >= 4 else {}
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    lat, parts = P.stream_blockwise_audio(model, dac, pca, blockwise, spk, smask, ids, mask, it, [160] * 4, **knobs, **kv)
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    stamps = [ev0.elapsed_time(ev) for _, ev in parts]  # device timeline: when each block's audio was complete
    print(f"run {it}{' (cached speaker KV)' if kv else ''}: audio of block i complete after " + ", ".join(f"{s:.1f}" for s in stamps) +
          f" ms (device timeline; the first = time to first audio, incl. text + 1600-patch speaker KV); host enqueue "
          f"returned after {t_enq*1e3:.1f} ms; total {total*1e3:.1f} ms for {4*160*2048/44100:.1f} s of audio", flush=True)
