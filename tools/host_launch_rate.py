"""Host-side cost of one kernel launch of the library (descriptor set-up + tensor-map lookup + cudaLaunchKernelEx with
PDL): the tiny model's kernels take ~2-3 us each, so a long tiny sampler run is bound by the host's enqueue rate.
Compare with the launch rates of the real requests: configs[1] 7 200 launches / 180 ms = 25 us per launch, configs[4]
29 000 launches / 377 ms = 13 us per launch."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echo_tts_b200.config import DitConfig  # noqa: E402
from echo_tts_b200.model import B200EchoDiT  # noqa: E402
from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample  # noqa: E402
from echo_tts_b200.weights import make_dit_weights  # noqa: E402

cfg = DitConfig.tiny()
model = B200EchoDiT.from_state_dict(make_dit_weights(cfg, 1234), cfg, "cuda:0")
ids = torch.zeros(1, 32, dtype=torch.int32)
mask = torch.zeros(1, 32, dtype=torch.bool)
mask[0, :10] = True
spk, smask = torch.randn(1, 16, 80), torch.ones(1, 16, dtype=torch.bool)
knobs = dict(num_steps=400, cfg_scale_text=3.0, cfg_scale_speaker=8.0, cfg_min_t=0.5, cfg_max_t=1.0, truncation_factor=None,
             rescale_k=None, rescale_sigma=None, speaker_kv_scale=None, speaker_kv_max_layers=None, speaker_kv_min_t=None)
noise = torch.randn(1, 16, 80)
for it in range(3):
    n0 = model.h.num_launches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sample(model, spk, smask, ids, mask, 0, sequence_length=16, noise=noise, **knobs)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = model.h.num_launches() - n0
    print(f"run {it}: {n} launches, host enqueue {1e3 * (t1 - t0):.1f} ms = {1e6 * (t1 - t0) / n:.2f} us per launch; "
          f"GPU done {1e3 * (t2 - t0):.1f} ms = {1e6 * (t2 - t0) / n:.2f} us per launch", flush=True)
