"""Per-launch CUDA-event profile (the library's own echo_profile_start / _stop, ECHO_PROFILE_DUMP=1 prints the table on
stderr) of one request: `cfg1` = BASELINE configs[1] (sampler + DAC decode), `cfg5` = configs[4] (blockwise 4 x 160,
1600-patch speaker KV, streaming decode). Events serialise the chain (no PDL overlap), so the SHARES matter, not the sum.
usage: ECHO_PROFILE_DUMP=1 python tools/profile_events.py cfg1|cfg5"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ECHO_PROFILE_DUMP", "1")
import bench  # noqa: E402
from echo_tts_b200 import _lib  # noqa: E402
from echo_tts_b200 import pipeline as P  # noqa: E402
from echo_tts_b200.autoencoder import ae_decode  # noqa: E402
from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as blockwise  # noqa: E402
from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
dev = torch.device("cuda", 0)
model, dac, pca = bench.load_models(dev, 0, 1)
ids, mask = bench.tokens(bench.PROMPT)
ids, mask = ids.to(dev), mask.to(dev)
if which == "cfg1":
    spk = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1)).to(dev)
    smask = torch.ones(1, 212, dtype=torch.bool, device=dev)
    noise = torch.randn(1, 640, 80, generator=torch.Generator().manual_seed(1000)).to(dev)

    def request():
        return ae_decode(dac, pca, sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **bench.KNOBS))
else:
    spk = torch.randn(1, 6400, 80, generator=torch.Generator().manual_seed(1)).to(dev)
    smask = torch.ones(1, 6400, dtype=torch.bool, device=dev)
    knobs = dict(bench.KNOBS, speaker_kv_scale=1.5, speaker_kv_min_t=0.9, speaker_kv_max_layers=24)

    def request():
        return P.stream_blockwise_audio(model, dac, pca, blockwise, spk, smask, ids, mask, 0, [160] * 4, **knobs)[0]

request()
torch.cuda.synchronize()
rep = _lib.ProfileReport()
_lib.check(model.lib.echo_profile_start(model.h.ptr), "echo_profile_start")
request()
_lib.check(model.lib.echo_profile_stop(model.h.ptr, C.byref(rep)), "echo_profile_stop")
print(f"{which}: gemm {rep.ms[0]:.2f} ms ({rep.launches[0]} launches), attention {rep.ms[1]:.2f} ms ({rep.launches[1]}), "
      f"glue {rep.ms[2]:.2f} ms ({rep.launches[2]})")
