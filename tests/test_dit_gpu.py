"""GPU parity of the CUDA DiT path (through the C ABI) against goldens from the REAL reference (fp32) and the oracle.

Tolerances (BASELINE.json north_star): per-layer outputs rel-L2 <= 1e-2, final latents rel-L2 <= 2e-2 vs the fp32
reference; KV caches must match the reference LAYOUT exactly ((B, L, H, 128) contiguous, list of num_layers (K, V))
and their values within 1.5e-2 (speaker/latent states cross a 14-layer bf16-operand encoder).
"""
import pytest
import torch

from echo_tts_b200.config import DitConfig
from echo_tts_b200.weights import make_dit_weights
from tests.util import EULER_KNOBS, PLAIN_KNOBS, gold, rel_l2

pytestmark = pytest.mark.gpu

LAYER_TOL = 1e-2
LATENT_TOL = 2e-2
KV_TOL = 1.5e-2


@pytest.fixture(scope="module")
def tiny():
    from echo_tts_b200.model import B200EchoDiT
    cfg = DitConfig.tiny()
    sd = make_dit_weights(cfg, seed=1234)
    model = B200EchoDiT.from_state_dict(sd, cfg, "cuda:0")
    return cfg, sd, model, gold("dit_tiny.pt")


def test_kv_cache_layout_and_values(tiny):
    cfg, sd, model, g = tiny
    caches = dict(kv_text=model.get_kv_cache_text(g["kv_ids"], g["kv_tmask"]),
                  kv_speaker=model.get_kv_cache_speaker(g["kv_spk"]),
                  kv_latent=model.get_kv_cache_latent(g["kv_pre"]))
    for name, cache in caches.items():
        ref = g[name]
        assert isinstance(cache, list) and len(cache) == cfg.num_layers
        for i, (k, v) in enumerate(cache):
            for t, r in ((k, ref[i, 0]), (v, ref[i, 1])):
                assert t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape) == tuple(r.shape)
                assert t.stride() == r.contiguous().stride()
            if name == "kv_text":
                # padded positions hold whatever the reference computes there too, but only valid keys are compared
                m = g["kv_tmask"]
                assert rel_l2(k[m.cuda()], ref[i, 0][m]) < KV_TOL and rel_l2(v[m.cuda()], ref[i, 1][m]) < KV_TOL
            else:
                assert rel_l2(k, ref[i, 0]) < KV_TOL, (name, i, rel_l2(k, ref[i, 0]))
                assert rel_l2(v, ref[i, 1]) < KV_TOL, (name, i, rel_l2(v, ref[i, 1]))


def test_forward_per_layer(tiny):
    """model(...) with the CFG batch layout (cond, no-text, no-speaker), a latent prefix and start_pos > 0."""
    cfg, sd, model, g = tiny
    ids, tm = g["kv_ids"][:1], g["kv_tmask"][:1]
    rep3 = lambda c: [(k.repeat(3, 1, 1, 1), v.repeat(3, 1, 1, 1)) for k, v in c]
    kt = rep3(model.get_kv_cache_text(ids, tm))
    ks = rep3(model.get_kv_cache_speaker(g["kv_spk"][:1]))
    kl = model.get_kv_cache_latent(g["kv_pre"][:1].repeat(3, 1, 1))
    sm = g["fw_smask"]
    mt = torch.cat([tm, torch.zeros_like(tm), tm])
    ms = torch.cat([sm, sm, torch.zeros_like(sm)])
    layers = []
    v = model(x=g["fw_x"], t=g["fw_t"], text_mask=mt, speaker_mask=ms, kv_cache_text=kt, kv_cache_speaker=ks,
              start_pos=int(g["fw_start"]), kv_cache_latent=kl, layer_outputs=layers)
    assert v.dtype == torch.float32 and tuple(v.shape) == tuple(g["fw_v"].shape)
    errs = [rel_l2(l, g["fw_layers"][i]) for i, l in enumerate(layers)]
    assert max(errs) < LAYER_TOL, errs
    assert rel_l2(v, g["fw_v"]) < LAYER_TOL, rel_l2(v, g["fw_v"])
    # the three branches must differ (masks took effect) and branch order must follow torch.cat([x, x, x])
    assert rel_l2(v[0], v[1]) > 1e-3 and rel_l2(v[0], v[2]) > 1e-3


def test_forward_matches_oracle_per_row_t(tiny):
    """Different t per batch row (the drop-in forward takes any t vector), no latent prefix."""
    from oracle import echo_oracle as O
    cfg, sd, model, g = tiny
    torch.manual_seed(3)
    b, S = 2, 70
    x = torch.randn(b, S, 80)
    t = torch.tensor([0.93, 0.21])
    ids, tm, spk = g["kv_ids"], g["kv_tmask"], g["kv_spk"]
    sm = g["eu_smask"]
    with torch.inference_mode():
        ref = O.dit_forward(sd, cfg, x, t, tm, sm, O.kv_cache_text(sd, cfg, ids, tm), O.kv_cache_speaker(sd, cfg, spk))
    v = model(x=x, t=t, text_mask=tm, speaker_mask=sm, kv_cache_text=model.get_kv_cache_text(ids, tm),
              kv_cache_speaker=model.get_kv_cache_speaker(spk))
    assert rel_l2(v, ref) < LAYER_TOL, rel_l2(v, ref)


@pytest.mark.parametrize("knobs,key", [(EULER_KNOBS, "eu_latent"), (PLAIN_KNOBS, "eu_latent_plain")])
def test_sampler_euler_vs_reference(tiny, knobs, key):
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    seed, S = int(g["eu_seed"]), int(g["eu_S"])
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(seed))  # the reference's CPU draw
    model.round_t_to_model_dtype = False  # golden = fp32 reference (fp32 t)
    try:
        out = sample(model, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], seed, sequence_length=S, noise=noise,
                     **knobs)
    finally:
        model.round_t_to_model_dtype = True
    assert out.dtype == torch.float32 and tuple(out.shape) == (2, S, 80)
    assert rel_l2(out, g[key]) < LATENT_TOL, rel_l2(out, g[key])


def test_sampler_euler_bf16_t_matches_oracle(tiny):
    """Default mode mirrors the reference's bf16 handler path: t rounded to bf16 before the embedding."""
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    from oracle import echo_oracle as O
    cfg, sd, model, g = tiny
    seed, S = 5, 24
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(seed))
    with torch.inference_mode():
        ref = O.sample_euler_cfg_independent_guidances(sd, cfg, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"],
                                                       noise, t_dtype=torch.bfloat16, **PLAIN_KNOBS)
    out = sample(model, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], seed, sequence_length=S, noise=noise,
                 **PLAIN_KNOBS)
    assert rel_l2(out, ref) < LATENT_TOL, rel_l2(out, ref)


def test_sampler_seed_draw_matches_torch(tiny):
    """Without injected noise the sampler draws torch.randn from a device generator exactly like the reference."""
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    S = 16
    rng = torch.Generator(device="cuda:0").manual_seed(123)
    noise = torch.randn((2, S, 80), device="cuda:0", dtype=torch.float32, generator=rng)
    a = sample(model, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], 123, sequence_length=S, **PLAIN_KNOBS)
    b = sample(model, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], 0, sequence_length=S, noise=noise,
               **PLAIN_KNOBS)
    assert torch.equal(a, b)


def test_sampler_blockwise_vs_reference(tiny):
    from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    seed = int(g["eu_seed"])
    blocks = [16, 8]
    rng = torch.Generator().manual_seed(seed)
    nb = [torch.randn((2, b, 80), generator=rng) for b in blocks]
    model.round_t_to_model_dtype = False
    try:
        out = sample(model, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], seed, blocks,
                     continuation_latent=g["bw_cont"], noise_blocks=nb, **dict(EULER_KNOBS, num_steps=6))
    finally:
        model.round_t_to_model_dtype = True
    assert tuple(out.shape) == (2, 32, 80)
    assert torch.equal(out[:, :8].cpu(), g["bw_cont"])
    assert rel_l2(out, g["bw_latent"]) < LATENT_TOL, rel_l2(out, g["bw_latent"])


def test_kv_scale_requires_min_t(tiny):
    """Reference behaviour: speaker_kv_scale without speaker_kv_min_t raises TypeError (inference.py:511)."""
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    with pytest.raises(TypeError):
        sample(model, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], 0, sequence_length=8,
               **dict(PLAIN_KNOBS, speaker_kv_scale=1.5))


def test_sample_euler_host_entry_is_bit_equal_to_the_device_entry(tiny):
    """`echo_sample_euler_host` (the entry a non-PyTorch host binds, INTEGRATION.md section 2: plain host pointers in,
    host pointer out, synchronous) must give exactly what `echo_sample_euler` gives on device tensors. Deterministic
    mode, so that bit equality is meaningful."""
    import ctypes as C

    import echo_tts_b200
    from echo_tts_b200 import _lib
    from echo_tts_b200.sampler import _args, sample_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    S = int(g["eu_S"])
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(int(g["eu_seed"])))
    spk, smask, ids, tmask = g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"]
    try:
        echo_tts_b200.set_deterministic(True)
        dev = sample(model, spk, smask, ids, tmask, 0, sequence_length=S, noise=noise, **EULER_KNOBS).cpu()
        a = _args(model, sequence_length=S, **EULER_KNOBS)
        # the host entry takes the fp32 speaker latents and rounds them to bf16 itself (inference.py:465)
        spk_h = spk.float().contiguous()
        sm_h = smask.to(torch.uint8).contiguous()
        ids_h = ids.to(torch.int32).contiguous()
        tm_h = tmask.to(torch.uint8).contiguous()
        nz_h = noise.float().contiguous()
        out_h = torch.empty(2, S, 80, dtype=torch.float32)
        with torch.cuda.device(model.device):
            _lib.check(model.lib.echo_sample_euler_host(model.h.ptr, C.byref(a), spk_h.data_ptr(), sm_h.data_ptr(),
                                                        sm_h.shape[1], ids_h.data_ptr(), tm_h.data_ptr(), tm_h.shape[1], 2,
                                                        nz_h.data_ptr(), out_h.data_ptr()), "echo_sample_euler_host")
    finally:
        echo_tts_b200.set_deterministic(False)
    assert torch.equal(out_h, dev)
    assert rel_l2(out_h, g["eu_latent"]) < LATENT_TOL


def test_sampler_skips_masked_text_rows_without_changing_the_result(tiny):
    """When the sampler knows the longest unmasked text prefix without a device round trip (mask still on the host, or
    built by pipeline.get_text_input_ids_and_mask), it encodes and projects only those text rows
    (echo_sampler_args::text_valid_len); with a device mask it runs all padded rows like the reference
    (model.py:606-613). The rows behind the prefix are masked out of every attention, so both must agree bit for bit."""
    import echo_tts_b200
    from echo_tts_b200.sampler import _text_valid_len, sample_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    S = int(g["eu_S"])
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(int(g["eu_seed"])))
    spk, smask, ids, tmask = g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"]
    # pad the text to twice its length so that there is a masked tail whatever the golden's mask looks like
    Lt = tmask.shape[1]
    ids2 = torch.cat([ids, torch.zeros_like(ids)], 1)
    tm2 = torch.cat([tmask, torch.zeros_like(tmask)], 1)
    n = _text_valid_len(tm2)
    assert 0 < n <= Lt and not bool(tm2[:, n:].any()) and bool(tm2[:, n - 1].any())
    assert _text_valid_len(tm2.cuda()) == 0  # a device mask: not known without a round trip
    try:
        echo_tts_b200.set_deterministic(True)
        short = sample(model, spk, smask, ids2, tm2, 0, sequence_length=S, noise=noise, **EULER_KNOBS)
        full = sample(model, spk, smask, ids2.cuda(), tm2.cuda(), 0, sequence_length=S, noise=noise, **EULER_KNOBS)
    finally:
        echo_tts_b200.set_deterministic(False)
    assert torch.equal(short, full)
    assert rel_l2(short.cpu(), g["eu_latent"]) < LATENT_TOL  # zero-padding the text changes nothing either


def test_speaker_kv_max_layers_zero_scales_no_layer(tiny):
    """reference `_multiply_kv_cache(cache, s, max_layers)` scales min(max_layers, L) layers: None = all, 0 = none
    (inference.py:408-414). With 0 layers the scaling knobs must change nothing."""
    import echo_tts_b200
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    S = 16
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(3))
    args = (model, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], 0)
    try:
        echo_tts_b200.set_deterministic(True)
        base = sample(*args, sequence_length=S, noise=noise, **dict(PLAIN_KNOBS, num_steps=4))
        zero = sample(*args, sequence_length=S, noise=noise,
                      **dict(PLAIN_KNOBS, num_steps=4, speaker_kv_scale=1.5, speaker_kv_min_t=0.6, speaker_kv_max_layers=0))
        allv = sample(*args, sequence_length=S, noise=noise,
                      **dict(PLAIN_KNOBS, num_steps=4, speaker_kv_scale=1.5, speaker_kv_min_t=0.6, speaker_kv_max_layers=None))
        huge = sample(*args, sequence_length=S, noise=noise,
                      **dict(PLAIN_KNOBS, num_steps=4, speaker_kv_scale=1.5, speaker_kv_min_t=0.6, speaker_kv_max_layers=99))
    finally:
        echo_tts_b200.set_deterministic(False)
    assert torch.equal(zero, base)
    assert not torch.equal(allv, base) and torch.equal(allv, huge)


def test_two_handles_two_streams_one_process(tiny):
    """Two handles in one process, each on its own CUDA stream (and on a second device when there is one): per-device
    kernel attributes and SM counts, and no shared scratch between handles."""
    from echo_tts_b200.model import B200EchoDiT
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    cfg, sd, model, g = tiny
    dev2 = "cuda:1" if torch.cuda.device_count() > 1 else "cuda:0"
    other = B200EchoDiT.from_state_dict(sd, cfg, dev2)
    S = 16
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(4))
    kn = dict(PLAIN_KNOBS, num_steps=4)
    args = (g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], 0)
    ref = sample(model, *args, sequence_length=S, noise=noise, **kn).cpu()
    s1, s2 = torch.cuda.Stream(model.device), torch.cuda.Stream(other.device)
    outs = []
    for _ in range(3):  # interleave the two handles; nothing synchronises between the calls
        with torch.cuda.stream(s1):
            a = sample(model, *args, sequence_length=S, noise=noise, **kn)
        with torch.cuda.device(other.device), torch.cuda.stream(s2):
            b = sample(other, *args, sequence_length=S, noise=noise, **kn)
        outs.append((a, b))
    torch.cuda.synchronize(model.device)
    torch.cuda.synchronize(other.device)
    for a, b in outs:
        assert rel_l2(a, ref) < 5e-3 and rel_l2(b, ref) < 5e-3  # atomics: run-to-run noise only
    # one handle, two streams back to back: the second call must wait for the first (shared workspace)
    with torch.cuda.stream(s1):
        a = sample(model, *args, sequence_length=S, noise=noise, **kn)
    with torch.cuda.stream(torch.cuda.Stream(model.device)):
        b = sample(model, *args, sequence_length=S, noise=noise, **kn)
    torch.cuda.synchronize(model.device)
    assert rel_l2(a, ref) < 5e-3 and rel_l2(b, ref) < 5e-3
