"""CPU: the oracle restatement against the golden vectors produced by the REAL reference (oracle/pin_reference.py).

The reference ships no tests or fixtures; these goldens are outputs of the reference's own
model.py / inference.py / inference_blockwise.py / autoencoder.py on the deterministic synthetic checkpoints.
"""
import pytest
import torch

from echo_tts_b200.config import DacConfig, DitConfig
from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state
from oracle import echo_oracle as O
from tests.util import EULER_KNOBS, PLAIN_KNOBS, byte_tokens, gold, rel_l2

TOL = 2e-5  # fp32 vs fp32, different op order


@pytest.fixture(scope="module")
def tiny():
    cfg = DitConfig.tiny()
    return cfg, make_dit_weights(cfg, seed=1234), gold("dit_tiny.pt")


def test_tokens_match_reference_tokenizer(tiny):
    _, _, g = tiny
    ids, mask = byte_tokens(["[S1] Hello there.", "[S2] A longer second prompt, with commas."], 48)
    # the reference normaliser rewrites ":" ";" to "," -- these prompts are already normal, so bytes must agree
    assert torch.equal(ids, g["kv_ids"]) and torch.equal(mask, g["kv_tmask"])


@torch.inference_mode()
def test_kv_caches(tiny):
    cfg, sd, g = tiny
    for name, fn, args in (("kv_text", O.kv_cache_text, (g["kv_ids"], g["kv_tmask"])),
                           ("kv_speaker", O.kv_cache_speaker, (g["kv_spk"],)),
                           ("kv_latent", O.kv_cache_latent, (g["kv_pre"],))):
        out = fn(sd, cfg, *args)
        ref = g[name]  # (layers, 2, B, L, H, 128)
        assert len(out) == cfg.num_layers
        for i, (k, v) in enumerate(out):
            assert k.shape == ref[i, 0].shape == v.shape
            assert rel_l2(k, ref[i, 0]) < TOL and rel_l2(v, ref[i, 1]) < TOL


@torch.inference_mode()
def test_forward_and_layers(tiny):
    cfg, sd, g = tiny
    ids, tm = g["kv_ids"][:1], g["kv_tmask"][:1]
    kt = O._batch3(O.kv_cache_text(sd, cfg, ids, tm))
    ks = O._batch3(O.kv_cache_speaker(sd, cfg, g["kv_spk"][:1]))
    kl = O.kv_cache_latent(sd, cfg, g["kv_pre"][:1].repeat(3, 1, 1))
    sm = g["fw_smask"]
    mt = torch.cat([tm, torch.zeros_like(tm), tm])
    ms = torch.cat([sm, sm, torch.zeros_like(sm)])
    layers = []
    v = O.dit_forward(sd, cfg, g["fw_x"], g["fw_t"], mt, ms, kt, ks, int(g["fw_start"]), kl, layers)
    assert rel_l2(v, g["fw_v"]) < TOL
    for i, l in enumerate(layers):
        assert rel_l2(l, g["fw_layers"][i]) < TOL


@torch.inference_mode()
def test_samplers(tiny):
    cfg, sd, g = tiny
    seed, S = int(g["eu_seed"]), int(g["eu_S"])
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(seed))
    args = (sd, cfg, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], noise)
    assert rel_l2(O.sample_euler_cfg_independent_guidances(*args, **EULER_KNOBS), g["eu_latent"]) < 1e-4
    assert rel_l2(O.sample_euler_cfg_independent_guidances(*args, **PLAIN_KNOBS), g["eu_latent_plain"]) < 1e-4
    rng = torch.Generator().manual_seed(seed)
    blocks = [16, 8]
    nb = [torch.randn((2, b, 80), generator=rng) for b in blocks]
    out = O.sample_blockwise_euler_cfg_independent_guidances(
        sd, cfg, g["kv_spk"], g["eu_smask"], g["kv_ids"], g["kv_tmask"], nb, block_sizes=blocks,
        continuation_latent=g["bw_cont"], **dict(EULER_KNOBS, num_steps=6))
    assert out.shape == (2, 8 + 24, 80)
    assert torch.equal(out[:, :8], g["bw_cont"])  # the continuation is returned untouched
    assert rel_l2(out, g["bw_latent"]) < 1e-4


@torch.inference_mode()
def test_dac_decode():
    cfg = DacConfig.tiny()
    sd = make_dac_weights(cfg, seed=4321)
    g = gold("dac_tiny.pt")
    comps, mean, scale = make_pca_state(cfg)
    audio = O.ae_decode(sd, cfg, comps, mean, scale, g["z"])
    assert audio.shape == (2, 1, 6 * cfg.hop) and cfg.hop == 2048
    assert rel_l2(audio, g["audio"]) < 1e-4
    zq = O.pca_unproject(comps, mean, scale, g["z"]).transpose(1, 2)
    assert rel_l2(O.dac_post_module(sd, cfg, zq), g["post"]) < TOL
    assert rel_l2(O.dac_upsample(sd, cfg, g["post"]), g["up"]) < TOL
    # decode is strictly causal: the first half of the latents decodes to the first half of the audio
    half = O.ae_decode(sd, cfg, comps, mean, scale, g["z"][:, :3])
    assert rel_l2(half, audio[..., : 3 * cfg.hop]) < 1e-4


def test_t_schedule_matches_torch():
    """The C++ fallback schedule (used when the caller passes none) is torch.linspace's fp32 closed form; torch's
    vectorised CPU kernel may differ from it by one ulp, which is why the Python wrapper passes torch's own values."""
    for n in (1, 2, 7, 8, 40, 41, 100):
        ref = torch.linspace(1.0, 0.0, n + 1) * 0.999
        steps, step = n + 1, torch.tensor(-1.0) / torch.tensor(float(n))
        mine = []
        for i in range(steps):
            v = (torch.tensor(1.0) + step * i) if i < steps // 2 else (torch.tensor(0.0) - step * (steps - 1 - i))
            mine.append(v * torch.tensor(0.999))
        assert torch.allclose(torch.stack(mine).float(), ref, rtol=0, atol=1.2e-7), n


@torch.inference_mode()
def test_dac_encode_oracle_matches_reference_golden():
    """Encode path: encoder, quantizer front, codes, z_q, PCA latents and get_speaker_latent_and_mask of
    the oracle against outputs of the reference's DAC.encode_zq / ae_encode / get_speaker_latent_and_mask (tiny)."""
    cfg = DacConfig.tiny()
    sd = make_dac_weights(cfg, seed=4321, include_encoder=True)
    comps, mean, scale = make_pca_state(cfg)
    g = gold("dac_encode_tiny.pt")
    zq, parts = O.dac_encode_zq(sd, cfg, g["audio"], return_parts=True)
    assert rel_l2(parts["z_enc"], g["z_enc"]) < TOL
    assert torch.equal(parts["codes"], g["codes"])
    assert rel_l2(zq, g["zq"]) < TOL
    assert rel_l2(O.ae_encode(sd, cfg, comps, mean, scale, g["audio"]), g["latent"]) < TOL
    lat, mask = O.get_speaker_latent_and_mask(sd, cfg, (comps, mean, scale), g["spk_wav"], max_speaker_latent_length=48,
                                              audio_chunk_size=8 * cfg.frame_length)
    assert torch.equal(mask, g["spk_mask"]) and rel_l2(lat, g["spk_latent"]) < TOL
    assert lat.shape[1] % 4 == 0 and lat.shape[1] == 40
