"""GPU parity of the DAC ENCODE path (Encoder -> quantizer.downsample -> pre_module -> semantic + residual VQ
-> z_q -> PCA latents) against goldens from the REAL reference (fp32 autoencoder.py / inference.py).

The quantizers take discrete decisions, so parity is checked stage by stage:
  * the quantizers' INPUT (continuous) against the reference within the latent tolerance;
  * the VQ kernel itself, exactly: the codes it picks must be the codes the oracle picks for the SAME fp32 input, and
    z_q must be the oracle's from_codes of those codes (fp32 tolerance);
  * end to end: share of codes equal to the reference's and the resulting latent error are reported and bounded.
"""
import pytest
import torch

from echo_tts_b200.config import DacConfig
from echo_tts_b200.weights import make_dac_weights, make_pca_state
from oracle import echo_oracle as O
from tests.util import gold, rel_l2

pytestmark = pytest.mark.gpu
Z_TOL = 2e-2


def _build(cfg):
    from echo_tts_b200.autoencoder import B200DAC, PCAState
    sd = make_dac_weights(cfg, seed=4321, include_encoder=True)
    dac = B200DAC.from_state_dict(sd, cfg, "cuda:0")
    comps, mean, scale = make_pca_state(cfg)
    return sd, dac, PCAState(comps, mean, scale)


@pytest.mark.parametrize("which", ["tiny", "full"])
def test_dac_encode_stages(which):
    cfg = DacConfig.tiny() if which == "tiny" else DacConfig.base()
    sd, dac, pca = _build(cfg)
    g = gold(f"dac_encode_{which}.pt")
    zq, codes, z_pre = dac.encode_zq(g["audio"], return_codes=True, return_z_pre=True)
    B, _, T = g["zq"].shape
    assert tuple(zq.shape) == (B, cfg.latent_dim, T) and tuple(codes.shape) == (B, 1 + cfg.n_codebooks, T)
    # (1) continuous part: encoder + downsample + pre_module
    e_pre = rel_l2(z_pre.transpose(1, 2), g["z_pre"])
    # (2) the VQ kernel on its own input: exact codes, fp32-accurate z_q
    with torch.inference_mode():
        codes_o = O.dac_encode_codes(sd, cfg, z_pre.transpose(1, 2).cpu())
        zq_o = O.dac_zq_from_codes(sd, cfg, codes.cpu())
    same_input = (codes.cpu() == codes_o).float().mean().item()
    # (3) end to end against the reference
    agree = (codes.cpu() == g["codes"]).float().mean().item()
    first = (codes.cpu()[:, 0] == g["codes"][:, 0]).float().mean().item()
    print(f"{which}: z_pre rel-L2 {e_pre:.3e}; VQ codes vs oracle on the same input {same_input:.4f}; "
          f"codes vs reference {agree:.3f} (semantic codebook {first:.3f}); z_q vs reference {rel_l2(zq, g['zq']):.3e}")
    assert e_pre < Z_TOL, e_pre
    assert same_input > 0.999, same_input      # only exact fp32 near-ties may differ
    assert rel_l2(zq, zq_o) < 1e-5
    # (2) proves the codes are exactly right for the kernel's own input, (1) bounds that input: every disagreement with
    # the reference's codes is therefore a near-tie flipped by the <= 1e-2 bf16 operand noise upstream. A flip at residual
    # level q changes the residual and so usually every later level of that frame (up to 9 of 10 codes per flip), which
    # makes the overall share noisy: measured on B200 at full size 0.998 (1 flipped frame of 48, round 1) and 0.979
    # (round 2 build), tiny 0.993; the semantic codebook, which carries most of the signal, 1.000 in every build.
    frames_flipped = (codes.cpu() != g["codes"]).any(dim=1).float().mean().item()
    print(f"{which}: frames with any flipped code {frames_flipped:.3f}")
    # (1) + (2) are the correctness criteria; the shares below are consequences of them (a flip needs a decision margin
    # smaller than the input perturbation allows) and are bounded loosely: observed 0.973-0.998 / 2-6 % of the frames
    assert first >= 0.99, first
    assert agree >= 0.90, agree
    assert frames_flipped <= 0.25, frames_flipped


def test_ae_encode_and_speaker_latents_tiny():
    """ae_encode == PCA projection of encode_zq (inference.py:219-224); get_speaker_latent_and_mask chunking / masks /
    trimming (inference.py:240-283) reproduce the reference's shapes and mask exactly."""
    from echo_tts_b200 import pipeline as P
    from echo_tts_b200.autoencoder import ae_encode
    cfg = DacConfig.tiny()
    sd, dac, pca = _build(cfg)
    g = gold("dac_encode_tiny.pt")
    zq = dac.encode_zq(g["audio"])
    lat = ae_encode(dac, pca, g["audio"])
    ref = ((zq.cpu().transpose(1, 2) - pca.pca_mean) @ pca.pca_components.T) * pca.latent_scale
    assert tuple(lat.shape) == tuple(g["latent"].shape) and rel_l2(lat, ref) < 1e-5
    sl, sm = P.get_speaker_latent_and_mask(dac, pca, g["spk_wav"].cuda(), max_speaker_latent_length=48,
                                           audio_chunk_size=8 * cfg.frame_length)
    assert tuple(sl.shape) == tuple(g["spk_latent"].shape) and torch.equal(sm.cpu(), g["spk_mask"])
    # chunk independence: the batched call equals chunk-by-chunk encoding
    fl = cfg.frame_length
    wav = g["spk_wav"][:, : 48 * fl]
    parts = [ae_encode(dac, pca, torch.nn.functional.pad(wav[:, i:i + 8 * fl], (0, max(0, 8 * fl - wav[:, i:i + 8 * fl].shape[1])))[None].cuda())
             for i in range(0, wav.shape[1], 8 * fl)]
    assert torch.equal(torch.cat(parts, 1)[:, : sl.shape[1]], sl)


def test_encode_requires_encoder_weights():
    from echo_tts_b200 import _lib
    from echo_tts_b200.autoencoder import B200DAC
    cfg = DacConfig.tiny()
    dac = B200DAC.from_state_dict(make_dac_weights(cfg, seed=4321), cfg, "cuda:0")  # decode path only
    with pytest.raises(_lib.EchoError):
        dac.encode_zq(torch.zeros(1, 1, cfg.frame_length))
