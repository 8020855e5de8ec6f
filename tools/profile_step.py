"""One full request (sampler + DAC decode) of BASELINE configs[1] between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echo_tts_b200.autoencoder import ae_decode  # noqa: E402
from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample  # noqa: E402

dev = torch.device("cuda", 0)
model, dac, pca = bench.load_models(dev, 0, 1)
ids, mask = bench.tokens(bench.PROMPT)
spk = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1))
smask = torch.ones(1, 212, dtype=torch.bool)
noise = torch.randn(2, 1, 640, 80, generator=torch.Generator().manual_seed(1000)).to(dev)
steps = int(os.environ.get("ECHO_PROFILE_STEPS", "40"))
knobs = dict(bench.KNOBS, num_steps=steps)


def request(i):
    lat = sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise[i], **knobs)
    return ae_decode(dac, pca, lat)


request(0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
audio = request(1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(audio.abs().mean()))
