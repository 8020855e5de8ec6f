"""Full-size (echo-tts-base dimensions) GPU parity against goldens produced by the REAL reference in fp32 on CPU
(oracle/pin_reference.py --full cfg1|cfg2): BASELINE.json configs[0] and configs[1].

North-star tolerances: per-layer outputs rel-L2 <= 1e-2, final latents rel-L2 <= 2e-2 vs the fp32 reference.

With default-style random init every residual branch adds only ~1 % to the stream, so the per-layer bar alone would
not notice a branch that is wrong by 30 %. Two additions close that hole:
  * INCREMENT parity: for every block the attention increment (stream after the attention branch - stream before) and
    the MLP increment (model.py:388-389) are compared with the reference's own, rel-L2 <= INCREMENT_TOL, on the CFG
    step 0 (three branches) and on a plain step at t = 0.25;
  * a "trained-like" checkpoint (cfg2t: wo / w2 of the DiT blocks x 32, branches are O(0.3) of the stream) on which
    the same north-star bars are held.
"""
import os

import pytest
import torch

from echo_tts_b200.config import DitConfig
from echo_tts_b200.weights import iter_dit_weights
from tests.util import GOLD, HANDLER_KNOBS, byte_tokens, gold, rel_l2

pytestmark = pytest.mark.gpu
PROMPT = "[S1] Hello from Echo-TTS on B200."
INCREMENT_TOL = 3e-2  # rel-L2 of a block's attention / MLP increment vs the fp32 reference's increment
ROWS = [0, 319, 639]  # probe rows stored by oracle/pin_reference.py


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


def _increment_errors(layers, mids, x0_rows, mid_rows, out_rows):
    """rel-L2 of every block's two residual increments (CUDA vs reference) on the probe rows."""
    errs_a, errs_m, size_a, size_m = [], [], [], []
    prev_c, prev_r = None, x0_rows
    for i in range(len(layers)):
        out_c, mid_c = layers[i][:, ROWS].cpu(), mids[i][:, ROWS].cpu()
        in_c = prev_c if prev_c is not None else None
        if in_c is not None:  # block 0's input (in_proj output) is not probed on the CUDA side: its increment starts at block 1
            errs_a.append(rel_l2(mid_c - in_c, mid_rows[i] - prev_r))
            size_a.append(((mid_rows[i] - prev_r).norm() / prev_r.norm()).item())
        errs_m.append(rel_l2(out_c - mid_c, out_rows[i] - mid_rows[i]))
        size_m.append(((out_rows[i] - mid_rows[i]).norm() / mid_rows[i].norm()).item())
        prev_c, prev_r = out_c, out_rows[i]
    return errs_a, errs_m, size_a, size_m


def _probe_forward(model, x, t, mask, smask, kt, ks):
    layers, mids = [], []
    v = model(x=x, t=t, text_mask=mask, speaker_mask=smask, kv_cache_text=kt, kv_cache_speaker=ks, layer_outputs=layers,
              layer_mids=mids)
    return v, layers, mids


def _check_increments(model, g, which):
    ids, mask, spk, smask = _inputs("cfg2")
    noise = torch.randn((1, 640, 80), generator=torch.Generator().manual_seed(0))
    kt, ks = model.get_kv_cache_text(ids, mask), model.get_kv_cache_speaker(spk)
    rep3 = lambda c: [(k.repeat(3, 1, 1, 1), v.repeat(3, 1, 1, 1)) for k, v in c]
    mt = torch.cat([mask, torch.zeros_like(mask), mask])
    ms = torch.cat([smask, smask, torch.zeros_like(smask)])
    t0 = (torch.linspace(1., 0., 41) * 0.999)[0]
    # ---- CFG step 0: the three branches
    v, layers, mids = _probe_forward(model, noise.repeat(3, 1, 1), torch.ones(3) * t0, mt, ms, rep3(kt), rep3(ks))
    ea, em, sa, sm_ = _increment_errors(layers, mids, g["step0_x0_rows"], g["step0_mid_rows"], g["step0_layers_rows"])
    lay = [rel_l2(l[:, ROWS], g["step0_layers_rows"][i]) for i, l in enumerate(layers)]
    print(f"{which} step 0 (b=3): increments |attn|/|x| {min(sa):.3f}..{max(sa):.3f} |mlp|/|x| {min(sm_):.3f}..{max(sm_):.3f}; "
          f"increment rel-L2 attention max {max(ea):.3e} mlp max {max(em):.3e}; per-layer max {max(lay):.3e}; "
          f"v {rel_l2(v, g['step0_v']):.3e}")
    assert max(ea) < INCREMENT_TOL and max(em) < INCREMENT_TOL, (ea, em)
    assert max(lay) < 1e-2, lay
    assert rel_l2(v, g["step0_v"]) < 1e-2
    del layers, mids
    # ---- a plain step (b = 1, t < cfg_min_t) on the reference's own x_t of that step
    v, layers, mids = _probe_forward(model, g["plain_x"], g["plain_t"], mask, smask, kt, ks)
    ea, em, sa, sm_ = _increment_errors(layers, mids, g["plain_x0_rows"], g["plain_mid_rows"], g["plain_out_rows"])
    lay = [rel_l2(l[:, ROWS], g["plain_out_rows"][i]) for i, l in enumerate(layers)]
    print(f"{which} plain step (call {int(g['plain_call'])}, t = {float(g['plain_t'][0]):.4f}): increment rel-L2 attention max "
          f"{max(ea):.3e} mlp max {max(em):.3e}; per-layer max {max(lay):.3e}; v {rel_l2(v, g['plain_v']):.3e}")
    assert max(ea) < INCREMENT_TOL and max(em) < INCREMENT_TOL, (ea, em)
    assert max(lay) < 1e-2, lay
    assert rel_l2(v, g["plain_v"]) < 1e-2


def test_full_size_block_increments(base_model):
    """Per-block attention / MLP INCREMENTS against the reference (cfg2 golden, default-style init)."""
    g = gold("dit_full_cfg2.pt")
    if "step0_mid_rows" not in g:
        pytest.skip("dit_full_cfg2.pt predates the increment probes (oracle/pin_reference.py --full cfg2)")
    _check_increments(base_model, g, "cfg2")


def test_full_size_trained_like_weights():
    """cfg2t: the same request on a checkpoint whose residual branches are O(0.3) of the stream (wo / w2 x 32):
    increments, per-layer outputs (<= 1e-2), step velocities and the final latents (<= 2e-2) vs the fp32 reference."""
    path = os.path.join(GOLD, "dit_full_cfg2t.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    from echo_tts_b200.model import B200EchoDiT
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    g = gold("dit_full_cfg2t.pt")
    cfg = DitConfig.base()
    model = B200EchoDiT(cfg, "cuda:0").load_state_dict(
        iter_dit_weights(cfg, 1234, include_latent=False, branch_gain=float(g["branch_gain"])))
    model.round_t_to_model_dtype = False
    _check_increments(model, g, "cfg2t")
    ids, mask, spk, smask = _inputs("cfg2")
    noise = torch.randn((1, 640, 80), generator=torch.Generator().manual_seed(0))
    out = sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise, **HANDLER_KNOBS)
    e = rel_l2(out, g["latent"])
    print("cfg2t final latent rel-L2", e)
    assert e < 2e-2, e
    del model
    torch.cuda.empty_cache()
