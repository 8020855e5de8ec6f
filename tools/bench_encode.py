"""Speaker-reference encoding on B200: get_speaker_latent_and_mask (reference inference.py:240-283) for a 10 s and a
5 min reference, full-size random-init Fish S1-DAC. Prints ms per call (CUDA events, after warm-up)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200 import pipeline as P  # noqa: E402
from echo_tts_b200.autoencoder import B200DAC, PCAState  # noqa: E402
from echo_tts_b200.config import DacConfig  # noqa: E402
from echo_tts_b200.weights import make_dac_weights, make_pca_state  # noqa: E402

cfg = DacConfig.base()
dac = B200DAC.from_state_dict(make_dac_weights(cfg, 4321, include_encoder=True), cfg, "cuda:0")
comps, mean, scale = make_pca_state(cfg)
pca = PCAState(comps.cuda(), mean.cuda(), scale)
for seconds in (10, 30, 300):
    wav = 0.3 * torch.randn(1, int(seconds * 44100), generator=torch.Generator().manual_seed(1)).cuda()
    for _ in range(2):
        lat, mask = P.get_speaker_latent_and_mask(dac, pca, wav)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lat, mask = P.get_speaker_latent_and_mask(dac, pca, wav)
    e1.record()
    torch.cuda.synchronize()
    print(f"{seconds:4d} s reference -> latents {tuple(lat.shape)}: {e0.elapsed_time(e1) / 3:8.2f} ms per call "
          f"({seconds / (e0.elapsed_time(e1) / 3e3):8.0f} audio-s/s); finite={bool(torch.isfinite(lat).all())}", flush=True)
print("peak memory GB", torch.cuda.max_memory_allocated() / 1e9)
