"""Full-size (echo-tts-base dimensions) GPU parity against goldens produced by the REAL reference in fp32 on CPU
(oracle/pin_reference.py --full cfg1|cfg2): BASELINE.json configs[0] and configs[1].

North-star tolerances: per-layer outputs rel-L2 <= 1e-2, final latents rel-L2 <= 2e-2 vs the fp32 reference.
"""
import os

import pytest
import torch

from echo_tts_b200.config import DitConfig
from echo_tts_b200.weights import iter_dit_weights
from tests.util import GOLD, HANDLER_KNOBS, byte_tokens, gold, rel_l2

pytestmark = pytest.mark.gpu
PROMPT = "[S1] Hello from Echo-TTS on B200."


@pytest.fixture(scope="module")
def base_model():
    from echo_tts_b200.model import B200EchoDiT
    cfg = DitConfig.base()
    model = B200EchoDiT(cfg, "cuda:0").load_state_dict(iter_dit_weights(cfg, 1234, include_latent=False))
    model.round_t_to_model_dtype = False  # goldens are the fp32 reference (fp32 t)
    return model


def _inputs(which):
    ids, mask = byte_tokens([PROMPT], 768)  # sample_pipeline pads to 768 (reference inference.py:327)
    if which == "cfg1":
        return ids, mask, torch.zeros(1, 4, 80), torch.zeros(1, 4, dtype=torch.bool)
    spk = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1))
    return ids, mask, spk, torch.ones(1, 212, dtype=torch.bool)


@pytest.mark.parametrize("which", ["cfg1", "cfg2"])
def test_full_size_sampler_and_layers(base_model, which):
    path = os.path.join(GOLD, f"dit_full_{which}.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    g = gold(f"dit_full_{which}.pt")
    model = base_model
    ids, mask, spk, smask = _inputs(which)
    noise = torch.randn((1, 640, 80), generator=torch.Generator().manual_seed(0))  # the reference's CPU draw, seed 0

    # step 0 of the sampler = one forward with the CFG batch layout: per-layer outputs on three probe rows
    kt = model.get_kv_cache_text(ids, mask)
    ks = model.get_kv_cache_speaker(spk)
    rep3 = lambda c: [(k.repeat(3, 1, 1, 1), v.repeat(3, 1, 1, 1)) for k, v in c]
    mt = torch.cat([mask, torch.zeros_like(mask), mask])
    ms = torch.cat([smask, smask, torch.zeros_like(smask)])
    t0 = (torch.linspace(1., 0., 41) * 0.999)[0]
    layers = []
    v = model(x=noise.repeat(3, 1, 1), t=torch.ones(3) * t0, text_mask=mt, speaker_mask=ms, kv_cache_text=rep3(kt),
              kv_cache_speaker=rep3(ks), layer_outputs=layers)
    ref_rows = g["step0_layers_rows"]  # (24, 3, 3, 2048): rows 0, 319, 639 of each branch
    errs = [rel_l2(l[:, [0, 319, 639]], ref_rows[i]) for i, l in enumerate(layers)]
    print(which, "per-layer rel-L2 max", max(errs), "v", rel_l2(v, g["step0_v"]))
    assert max(errs) < 1e-2, errs
    assert rel_l2(v, g["step0_v"]) < 1e-2
    del layers

    out = sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **HANDLER_KNOBS)
    e = rel_l2(out, g["latent"])
    print(which, "final latent rel-L2", e)
    assert tuple(out.shape) == (1, 640, 80) and e < 2e-2, e


def test_full_size_determinism_switch(base_model):
    """echo_set_deterministic(1): the same request twice is bit-identical (no atomic split-K). Default mode: two runs
    may differ in the last bits of the residual stream, which decorrelates them to the bf16 rounding-noise floor --
    both stay within the north-star tolerance of the fp32 reference."""
    import echo_tts_b200
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    g = gold("dit_full_cfg2.pt")
    ids, mask, spk, smask = _inputs("cfg2")
    noise = torch.randn((1, 640, 80), generator=torch.Generator().manual_seed(0))
    knobs = dict(HANDLER_KNOBS, num_steps=40)
    try:
        echo_tts_b200.set_deterministic(True)
        a = sample(base_model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **knobs)
        b = sample(base_model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **knobs)
        assert torch.equal(a, b)
        assert rel_l2(a, g["latent"]) < 2e-2
    finally:
        echo_tts_b200.set_deterministic(False)
    c = sample(base_model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **knobs)
    d = sample(base_model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **knobs)
    print("fast mode: run-to-run rel-L2", rel_l2(c, d), "vs reference", rel_l2(c, g["latent"]), rel_l2(d, g["latent"]))
    assert rel_l2(c, g["latent"]) < 2e-2 and rel_l2(d, g["latent"]) < 2e-2
    assert rel_l2(c, d) < 2e-2


def test_full_size_blockwise_long_speaker_cfg5():
    """BASELINE configs[4]: sample_blockwise 4 x 160 with a 5-minute speaker reference (6400 latents -> 1600 speaker
    KV patches), speaker_kv_scale 1.5 on all layers until t < 0.9, against the fp32 reference golden
    (oracle/pin_reference.py --full cfg5). Needs the latent_* (blockwise) weights."""
    path = os.path.join(GOLD, "dit_full_cfg5.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    from echo_tts_b200.model import B200EchoDiT
    from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as sample_blockwise
    cfg = DitConfig.base()
    model = B200EchoDiT(cfg, "cuda:0").load_state_dict(iter_dit_weights(cfg, 1234, include_latent=True))
    model.round_t_to_model_dtype = False
    g = gold("dit_full_cfg5.pt")
    ids, mask = byte_tokens([PROMPT], 768)
    spk = torch.randn(1, 6400, 80, generator=torch.Generator().manual_seed(1))
    smask = torch.ones(1, 6400, dtype=torch.bool)
    rng = torch.Generator().manual_seed(0)  # the reference draws one block of noise per block from one CPU generator
    noise_blocks = [torch.randn((1, 160, 80), generator=rng) for _ in range(4)]
    knobs = dict(HANDLER_KNOBS, speaker_kv_scale=1.5, speaker_kv_min_t=0.9, speaker_kv_max_layers=24)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    out = sample_blockwise(model, spk, smask, ids, mask, 0, [160] * 4, noise_blocks=noise_blocks, **knobs)
    ev1.record()
    torch.cuda.synchronize()
    e = rel_l2(out, g["latent"])
    print(f"cfg5 blockwise final latent rel-L2 {e:.3e}; {ev0.elapsed_time(ev1):.1f} ms on B200 "
          f"(reference fp32 CPU: {float(g['seconds']):.0f} s on {int(g['threads'])} threads)")
    assert tuple(out.shape) == (1, 640, 80) and e < 2e-2, e
    del model
    torch.cuda.empty_cache()
