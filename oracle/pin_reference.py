"""Pins oracle/echo_oracle.py against the REAL reference and writes the golden fixtures under tests/golden/.

Runs only in the build container, where /root/reference exists (it does not exist on the GPU box). It imports the
reference's own model.py / inference.py / inference_blockwise.py / autoencoder.py (torchcodec, the one missing
dependency, is stubbed: only load_audio uses it), loads the deterministic synthetic checkpoints from
echo_tts_b200.weights into the reference modules, runs reference and oracle on identical inputs, asserts agreement,
and saves the REFERENCE outputs as fixtures.

  python oracle/pin_reference.py --tiny          # seconds: tiny-config goldens (DiT, samplers, DAC)
  python oracle/pin_reference.py --full cfg1     # minutes: full-size echo-tts-base goldens (BASELINE configs)
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import types
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("ECHO_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: goldens can only be regenerated where the reference is mounted")
    sys.path.insert(0, REF)
    for name in ("torchcodec", "torchcodec.decoders"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torchcodec.decoders"].AudioDecoder = object
    for name in ("huggingface_hub", "safetensors", "safetensors.torch", "torchaudio"):
        try:
            __import__(name)
        except Exception:  # pragma: no cover - only if the image lacks them
            m = types.ModuleType(name)
            m.hf_hub_download = None
            sys.modules[name] = m
    import autoencoder  # noqa
    import inference  # noqa
    import inference_blockwise  # noqa
    import model  # noqa
    return model, inference, inference_blockwise, autoencoder


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def build_ref_dit(refmodel, cfg, sd):
    with torch.device("meta"):
        m = refmodel.EchoDiT(**cfg.as_dict())
    m.load_state_dict(sd, strict=True, assign=True)
    return m.eval()


def build_ref_dac(refae, cfg, sd):
    """Reference DAC with the decode-path weights replaced; encoder etc. keep meta tensors (never touched)."""
    if cfg.latent_dim == 1024 and cfg.post_layers == 8:
        with torch.device("meta"):
            ae = refae.build_ae()
    else:
        q_config = refae.ModelArgs(block_size=cfg.post_block_size, n_layer=cfg.post_layers, n_head=cfg.post_heads,
                                   dim=cfg.latent_dim, intermediate_size=cfg.post_intermediate,
                                   head_dim=cfg.latent_dim // cfg.post_heads, norm_eps=cfg.post_norm_eps,
                                   dropout_rate=0.1, attn_dropout_rate=0.1, channels_first=True)
        with torch.device("meta"):
            quant = refae.DownsampleResidualVectorQuantize(
                input_dim=cfg.latent_dim, n_codebooks=1, codebook_size=8, codebook_dim=8, downsample_factor=(2, 2),
                semantic_codebook_size=8,
                post_module=refae.WindowLimitedTransformer(causal=True, window_size=cfg.post_window,
                                                           input_dim=cfg.latent_dim, config=q_config))
            ae = refae.DAC(encoder_dim=8, encoder_rates=[2, 2], latent_dim=cfg.latent_dim, decoder_dim=cfg.decoder_dim,
                           decoder_rates=list(cfg.rates), quantizer=quant, causal=True,
                           decoder_transformer_layers=[0] * len(cfg.rates))
    missing, unexpected = ae.load_state_dict(sd, strict=False, assign=True)
    assert not unexpected, unexpected
    # non-persistent / derived buffers of the post_module must be real tensors
    pm = ae.quantizer.post_module
    pm.freqs_cis = refae.precompute_freqs_cis(pm.config.block_size, pm.config.head_dim, pm.config.rope_base)
    pm.causal_mask = torch.tril(torch.ones(pm.config.block_size, pm.config.block_size, dtype=torch.bool))
    return ae.eval()


def build_ref_dac_with_encoder(refae, cfg, sd):
    """Reference DAC including the encode path (Encoder, downsample, pre_module, vector quantizers) for `cfg`."""
    def tcfg(block_size, layers, heads, dim, inter):
        return refae.ModelArgs(block_size=block_size, n_layer=layers, n_head=heads, dim=dim, intermediate_size=inter,
                               head_dim=64, norm_eps=1e-5, dropout_rate=0.1, attn_dropout_rate=0.1, channels_first=True)

    def general(**kw):  # build_ae's transformer_general_config (autoencoder.py:1169-1183)
        return tcfg(kw.get("block_size", 16384), kw.get("n_layer", 8), kw.get("n_head", 8), kw.get("dim", 512),
                    kw.get("intermediate_size", 1536))

    C = cfg.latent_dim
    nb = len(cfg.enc_rates)
    with torch.device("meta"):
        mk = lambda: refae.WindowLimitedTransformer(
            causal=True, window_size=cfg.post_window, input_dim=C,
            config=tcfg(cfg.post_block_size, cfg.post_layers, cfg.post_heads, C, cfg.post_intermediate))
        quant = refae.DownsampleResidualVectorQuantize(
            input_dim=C, n_codebooks=cfg.n_codebooks, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim,
            downsample_factor=(2,) * cfg.num_upsample, semantic_codebook_size=cfg.semantic_codebook_size,
            pre_module=mk(), post_module=mk())
        ae = refae.DAC(encoder_dim=cfg.enc_dim, encoder_rates=list(cfg.enc_rates), latent_dim=C,
                       decoder_dim=cfg.decoder_dim, decoder_rates=list(cfg.rates), quantizer=quant, causal=True,
                       encoder_transformer_layers=[0] * (nb - 1) + [cfg.enc_t_layers],
                       decoder_transformer_layers=[0] * len(cfg.rates), transformer_general_config=general)
    missing, unexpected = ae.load_state_dict(sd, strict=False, assign=True)
    assert not unexpected, unexpected
    assert all(k.endswith(("freqs_cis", "causal_mask")) for k in missing), missing
    mods = [ae.quantizer.post_module, ae.quantizer.pre_module]
    if cfg.enc_t_layers > 0:
        mods.append(ae.encoder.block[nb].block[5])
    for pm in mods:  # non-persistent / derived buffers must be real tensors
        pm.freqs_cis = refae.precompute_freqs_cis(pm.config.block_size, pm.config.head_dim, pm.config.rope_base)
        pm.causal_mask = torch.tril(torch.ones(pm.config.block_size, pm.config.block_size, dtype=torch.bool))
    return ae.eval()


def run_encode(which: str):
    """Pins the encode-path oracle (dac_encode_zq, ae_encode, get_speaker_latent_and_mask) against the reference's
    DAC.encode_zq / inference.ae_encode / inference.get_speaker_latent_and_mask and stores goldens."""
    from echo_tts_b200.config import DacConfig
    from echo_tts_b200.weights import make_dac_weights, make_pca_state
    from oracle import echo_oracle as O
    _, refinf, _, refae = import_reference()
    torch.set_num_threads(os.cpu_count())
    cfg = DacConfig.tiny() if which == "tiny" else DacConfig.base()
    sd = make_dac_weights(cfg, seed=4321, include_encoder=True)
    ae = build_ref_dac_with_encoder(refae, cfg, sd)
    comps, mean, scale = make_pca_state(cfg)
    pca = refinf.PCAState(pca_components=comps, pca_mean=mean, latent_scale=scale)
    g = torch.Generator().manual_seed(17)
    fl = cfg.frame_length
    n_frames = 37 if which == "tiny" else 48
    audio = 0.3 * torch.randn(2 if which == "tiny" else 1, 1, n_frames * fl - fl // 3, generator=g)  # ragged: not a frame multiple
    t0 = time.time()
    with torch.inference_mode():
        zq_ref = ae.encode_zq(audio)
        codes_ref, _ = ae.encode(audio)
        lat_ref = refinf.ae_encode(ae, pca, audio)
        z_enc_ref = ae.encoder(torch.nn.functional.pad(audio, (0, n_frames * fl - audio.shape[-1])))
        zq_o, parts = O.dac_encode_zq(sd, cfg, audio, return_parts=True)
        lat_o = O.ae_encode(sd, cfg, comps, mean, scale, audio)
    print(f"{which}: reference encode {time.time() - t0:.1f} s; frames {zq_ref.shape[-1]}")
    agree = (parts["codes"] == codes_ref).float().mean().item()
    print("encoder out oracle vs reference", rel(parts["z_enc"], z_enc_ref), "| code agreement", agree,
          "| z_q", rel(zq_o, zq_ref), "| latents", rel(lat_o, lat_ref))
    assert rel(parts["z_enc"], z_enc_ref) < 2e-5 and agree == 1.0 and rel(zq_o, zq_ref) < 2e-5 and rel(lat_o, lat_ref) < 2e-5
    out = dict(audio=audio, z_enc=z_enc_ref.clone(), z_pre=parts["z_pre"].clone(), codes=codes_ref.clone(),
               zq=zq_ref.clone(), latent=lat_ref.clone())
    if which == "tiny":
        # get_speaker_latent_and_mask with a small chunk size (the chunking logic is size independent)
        wav = 0.3 * torch.randn(1, 5 * 8 * fl + 3 * fl + 11, generator=g)
        kw = dict(max_speaker_latent_length=48, audio_chunk_size=8 * fl)
        orig = refinf.get_speaker_latent_and_mask.__wrapped__ if hasattr(refinf.get_speaker_latent_and_mask, "__wrapped__") else None
        src = open(os.path.join(REF, "inference.py")).read()
        assert "AE_DOWNSAMPLE_FACTOR = 2048" in src  # the reference hard-codes the base hop; tiny uses frame_length
        ns = dict(vars(refinf))
        import ast as _ast
        node = [n for n in _ast.parse(src).body if isinstance(n, _ast.FunctionDef) and n.name == "get_speaker_latent_and_mask"][0]
        node.decorator_list = []
        code = _ast.unparse(node).replace("AE_DOWNSAMPLE_FACTOR = 2048", f"AE_DOWNSAMPLE_FACTOR = {fl}")
        exec(code, ns)
        with torch.inference_mode():
            sl_ref, sm_ref = ns["get_speaker_latent_and_mask"](ae, pca, wav, **kw)
            sl_o, sm_o = O.get_speaker_latent_and_mask(sd, cfg, (comps, mean, scale), wav, **kw)
        assert sl_ref.shape == sl_o.shape and torch.equal(sm_ref, sm_o) and rel(sl_o, sl_ref) < 2e-5, (sl_ref.shape, sl_o.shape)
        print("get_speaker_latent_and_mask: oracle == reference", tuple(sl_ref.shape), rel(sl_o, sl_ref))
        out.update(spk_wav=wav, spk_latent=sl_ref.clone(), spk_mask=sm_ref.clone())
    torch.save(out, os.path.join(GOLD, f"dac_encode_{which}.pt"))
    print("wrote", os.path.join(GOLD, f"dac_encode_{which}.pt"))


def text_ids_mask(refinf, prompts, max_length):
    return refinf.get_text_input_ids_and_mask(prompts, max_length=max_length, device=None)


# ------------------------------------------------------------------------------------------------------ tiny
@torch.inference_mode()
def run_tiny():
    from echo_tts_b200.config import DacConfig, DitConfig
    from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state
    from oracle import echo_oracle as O

    refmodel, refinf, refblk, refae = import_reference()
    torch.manual_seed(0)
    cfg = DitConfig.tiny()
    sd = make_dit_weights(cfg, seed=1234)
    m = build_ref_dit(refmodel, cfg, sd)
    g = torch.Generator().manual_seed(7)
    out = {}

    # ---- KV caches (ragged text masks, B=2)
    ids, tmask = text_ids_mask(refinf, ["[S1] Hello there.", "[S2] A longer second prompt, with commas."], 48)
    spk = torch.randn(2, 16, 80, generator=g)
    pre = torch.randn(2, 32, 80, generator=g)
    kt_r, ks_r, kl_r = m.get_kv_cache_text(ids, tmask), m.get_kv_cache_speaker(spk), m.get_kv_cache_latent(pre)
    kt_o, ks_o, kl_o = O.kv_cache_text(sd, cfg, ids, tmask), O.kv_cache_speaker(sd, cfg, spk), O.kv_cache_latent(sd, cfg, pre)
    for name, r, o in (("text", kt_r, kt_o), ("speaker", ks_r, ks_o), ("latent", kl_r, kl_o)):
        for i in range(cfg.num_layers):
            assert r[i][0].shape == o[i][0].shape and r[i][0].is_contiguous()
            e = max(rel(o[i][0], r[i][0]), rel(o[i][1], r[i][1]))
            assert e < 2e-5, (name, i, e)
    out.update(kv_ids=ids, kv_tmask=tmask, kv_spk=spk, kv_pre=pre,
               kv_text=torch.stack([torch.stack(p) for p in kt_r]), kv_speaker=torch.stack([torch.stack(p) for p in ks_r]),
               kv_latent=torch.stack([torch.stack(p) for p in kl_r]))
    print("kv caches: oracle == reference")

    # ---- forward with the CFG batch layout + latent prefix
    S, start = 40, 16
    x = torch.randn(1, S, 80, generator=g).repeat(3, 1, 1)
    t = torch.full((3,), 0.7312)
    ids1, tm1 = ids[:1], tmask[:1]
    spk1 = spk[:1]
    smask1 = torch.ones(1, 16, dtype=torch.bool)
    smask1[0, 12:] = False
    kt = refinf._concat_kv_caches(*([m.get_kv_cache_text(ids1, tm1)] * 3))
    ks = refinf._concat_kv_caches(*([m.get_kv_cache_speaker(spk1)] * 3))
    kl = m.get_kv_cache_latent(pre[:1].repeat(3, 1, 1))
    mt = torch.cat([tm1, torch.zeros_like(tm1), tm1])
    ms = torch.cat([smask1, smask1, torch.zeros_like(smask1)])
    layers_r = []
    hooks = [blk.register_forward_hook(lambda _m, _i, o: layers_r.append(o.clone())) for blk in m.blocks]
    v_r = m(x=x, t=t, text_mask=mt, speaker_mask=ms, kv_cache_text=kt, kv_cache_speaker=ks, start_pos=start,
            kv_cache_latent=kl)
    for h in hooks:
        h.remove()
    layers_o = []
    v_o = O.dit_forward(sd, cfg, x, t, mt, ms, kt, ks, start, kl, layers_o)
    assert rel(v_o, v_r) < 2e-5, rel(v_o, v_r)
    for a, b in zip(layers_o, layers_r):
        assert rel(a, b) < 2e-5
    out.update(fw_x=x, fw_t=t, fw_smask=smask1, fw_start=torch.tensor(start), fw_v=v_r, fw_layers=torch.stack(layers_r))
    print("forward: oracle == reference", rel(v_o, v_r))

    # ---- Euler sampler, every knob on; noise drawn exactly as the reference does on CPU
    knobs = dict(num_steps=8, cfg_scale_text=3.0, cfg_scale_speaker=8.0, cfg_min_t=0.5, cfg_max_t=1.0,
                 truncation_factor=0.8, rescale_k=1.2, rescale_sigma=3.0, speaker_kv_scale=1.5, speaker_kv_max_layers=2,
                 speaker_kv_min_t=0.6)
    smask = torch.ones(2, 16, dtype=torch.bool)
    smask[1, 8:] = False
    seed, S = 11, 32
    lat_r = refinf.sample_euler_cfg_independent_guidances(m, spk, smask, ids, tmask, seed, sequence_length=S, **knobs)
    noise = torch.randn((2, S, 80), generator=torch.Generator().manual_seed(seed))
    lat_o = O.sample_euler_cfg_independent_guidances(sd, cfg, spk, smask, ids, tmask, noise, **knobs)
    assert rel(lat_o, lat_r) < 1e-4, rel(lat_o, lat_r)
    out.update(eu_smask=smask, eu_seed=torch.tensor(seed), eu_S=torch.tensor(S), eu_latent=lat_r)
    plain = dict(knobs, truncation_factor=None, rescale_k=None, rescale_sigma=None, speaker_kv_scale=None,
                 speaker_kv_max_layers=None, speaker_kv_min_t=None)
    lat_r2 = refinf.sample_euler_cfg_independent_guidances(m, spk, smask, ids, tmask, seed, sequence_length=S, **plain)
    lat_o2 = O.sample_euler_cfg_independent_guidances(sd, cfg, spk, smask, ids, tmask, noise, **plain)
    assert rel(lat_o2, lat_r2) < 1e-4
    out.update(eu_latent_plain=lat_r2)
    print("euler sampler: oracle == reference", rel(lat_o, lat_r), rel(lat_o2, lat_r2))

    # ---- blockwise sampler with a continuation latent
    bknobs = dict(knobs, num_steps=6)
    blocks = [16, 8]
    cont = torch.randn(2, 8, 80, generator=g)
    lat_rb = refblk.sample_blockwise_euler_cfg_independent_guidances(m, spk, smask, ids, tmask, seed, blocks,
                                                                     continuation_latent=cont, **bknobs)
    rng = torch.Generator().manual_seed(seed)
    nb = [torch.randn((2, b, 80), generator=rng) for b in blocks]
    lat_ob = O.sample_blockwise_euler_cfg_independent_guidances(sd, cfg, spk, smask, ids, tmask, nb, block_sizes=blocks,
                                                                continuation_latent=cont, **bknobs)
    assert rel(lat_ob, lat_rb) < 1e-4, rel(lat_ob, lat_rb)
    out.update(bw_cont=cont, bw_latent=lat_rb)
    print("blockwise sampler: oracle == reference", rel(lat_ob, lat_rb))
    torch.save({k: v.clone() for k, v in out.items()}, os.path.join(GOLD, "dit_tiny.pt"))

    # ---- DAC decode (tiny) + PCA
    dcfg = DacConfig.tiny()
    dsd = make_dac_weights(dcfg, seed=4321)
    ae = build_ref_dac(refae, dcfg, dsd)
    comps, mean, scale = make_pca_state(dcfg)
    z = torch.randn(2, 6, 80, generator=g)
    pca = refinf.PCAState(pca_components=comps, pca_mean=mean, latent_scale=scale)
    audio_r = refinf.ae_decode(ae, pca, z)
    audio_o = O.ae_decode(dsd, dcfg, comps, mean, scale, z)
    assert audio_r.shape == (2, 1, 6 * dcfg.hop)
    assert rel(audio_o, audio_r) < 1e-4, rel(audio_o, audio_r)
    zq = O.pca_unproject(comps, mean, scale, z).transpose(1, 2)
    post_r = ae.quantizer.post_module(zq)
    up_r = ae.quantizer.upsample(post_r)
    assert rel(O.dac_post_module(dsd, dcfg, zq), post_r) < 2e-5
    assert rel(O.dac_upsample(dsd, dcfg, post_r), up_r) < 2e-5
    torch.save(dict(z=z, audio=audio_r, post=post_r, up=up_r), os.path.join(GOLD, "dac_tiny.pt"))
    print("dac decode: oracle == reference", rel(audio_o, audio_r), "audio rms", audio_r.pow(2).mean().sqrt().item())


# ------------------------------------------------------------------------------------------------------ full size
BASE_PROMPT = "[S1] Hello from Echo-TTS on B200."
HANDLER_KNOBS = dict(num_steps=40, cfg_scale_text=3.0, cfg_scale_speaker=8.0, cfg_min_t=0.5, cfg_max_t=1.0,
                     truncation_factor=None, rescale_k=None, rescale_sigma=None, speaker_kv_scale=None,
                     speaker_kv_max_layers=None, speaker_kv_min_t=None)  # handler.py:431-442 defaults


@torch.inference_mode()
def run_full(which: str):
    from echo_tts_b200.config import DacConfig, DitConfig
    from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state

    refmodel, refinf, refblk, refae = import_reference()
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    if which == "dac":
        dcfg = DacConfig.base()
        dsd = make_dac_weights(dcfg, seed=4321)
        ae = build_ref_dac(refae, dcfg, dsd)
        comps, mean, scale = make_pca_state(dcfg)
        pca = refinf.PCAState(pca_components=comps, pca_mean=mean, latent_scale=scale)
        z = torch.randn(1, 64, 80, generator=torch.Generator().manual_seed(5))
        t1 = time.time()
        audio = refinf.ae_decode(ae, pca, z)
        print("dac T=64 decode", time.time() - t1, "s; rms", audio.pow(2).mean().sqrt().item())
        torch.save(dict(z=z, audio=audio.clone()), os.path.join(GOLD, "dac_full_T64.pt"))
        return
    if which == "dac640":
        # the full-length decode of BASELINE configs[1] (640 latents -> 1 310 720 samples), stored as fp16 (2.6 MB):
        # the waveform is tanh-bounded, fp16 keeps ~1e-3 relative precision per sample (rel-L2 ~3e-4 overall)
        dcfg = DacConfig.base()
        dsd = make_dac_weights(dcfg, seed=4321)
        ae = build_ref_dac(refae, dcfg, dsd)
        comps, mean, scale = make_pca_state(dcfg)
        pca = refinf.PCAState(pca_components=comps, pca_mean=mean, latent_scale=scale)
        z = torch.randn(1, 640, 80, generator=torch.Generator().manual_seed(6))
        t1 = time.time()
        audio = refinf.ae_decode(ae, pca, z)
        dt = time.time() - t1
        a16 = audio.to(torch.float16)
        print(f"dac T=640 decode {dt:.1f} s; rms {audio.pow(2).mean().sqrt().item():.4f}; fp16 storage rel-L2 "
              f"{rel(a16.float(), audio):.2e}")
        torch.save(dict(z=z, audio_f16=a16.clone(), seconds=torch.tensor(dt)), os.path.join(GOLD, "dac_full_T640.pt"))
        return
    cfg = DitConfig.base()
    gain = 32.0 if which == "cfg2t" else 1.0  # cfg2t: "trained-like" blocks, see weights.iter_dit_weights
    sd = make_dit_weights(cfg, seed=1234, branch_gain=gain)
    m = build_ref_dit(refmodel, cfg, sd)
    print("weights + reference model ready", time.time() - t0, "s")
    ids, tmask = text_ids_mask(refinf, [BASE_PROMPT], 768)  # sample_pipeline always pads to 768 (inference.py:327)
    seed = 0
    res = {}
    if which == "cfg1":
        spk = torch.zeros(1, 4, 80)
        smask = torch.zeros(1, 4, dtype=torch.bool)  # no speaker audio (inference.py:329-331)
    elif which in ("cfg2", "cfg2t"):
        spk = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1))
        smask = torch.ones(1, 212, dtype=torch.bool)
    elif which == "cfg5":
        # BASELINE configs[4]: blockwise 4 x 160, 5-minute speaker reference (6400 latents = 1600 patches),
        # speaker_kv_scale 1.5 on all layers until t < 0.9 (gradio defaults), other knobs as the handler
        spk = torch.randn(1, 6400, 80, generator=torch.Generator().manual_seed(1))
        smask = torch.ones(1, 6400, dtype=torch.bool)
        knobs = dict(HANDLER_KNOBS, speaker_kv_scale=1.5, speaker_kv_min_t=0.9, speaker_kv_max_layers=24)
        t1 = time.time()
        lat = refblk.sample_blockwise_euler_cfg_independent_guidances(m, spk, smask, ids, tmask, seed, [160] * 4, **knobs)
        dt = time.time() - t1
        print(f"cfg5: reference fp32 blockwise sampler {dt:.1f} s on {torch.get_num_threads()} threads")
        torch.save(dict(latent=lat.clone(), seconds=torch.tensor(dt), threads=torch.tensor(torch.get_num_threads())),
                   os.path.join(GOLD, "dit_full_cfg5.pt"))
        return
    else:
        raise SystemExit(which)
    # Per-block probes on rows 0 / 319 / 639 of every batch row, at two model calls of the sampler: call 0 (step 0, the
    # three CFG branches, t = 0.999) and call PLAIN_CALL (a plain b = 1 step, t < cfg_min_t). For every block the stream
    # BEFORE the block, AFTER the attention branch (= the input of mlp_adaln, model.py:388) and AFTER the MLP branch
    # (model.py:389) are kept, so that the tests can compare the two residual INCREMENTS of each block separately.
    PLAIN_CALL = 30
    ROWS = [0, 319, 639]
    call = {"i": -1}
    probes = {0: dict(mid=[], out=[]), PLAIN_CALL: dict(mid=[], out=[])}

    def want():
        return probes.get(call["i"])

    def pre_block0(_m, args, kwargs):
        if want() is not None:
            want()["x0"] = (args[0] if args else kwargs["x"])[:, ROWS].clone()

    def pre_mlp_adaln(_m, args):
        if want() is not None:
            want()["mid"].append(args[0][:, ROWS].clone())

    def post_block(_m, _i, o):
        if want() is not None:
            want()["out"].append(o[:, ROWS].clone())

    hooks = [m.blocks[0].register_forward_pre_hook(pre_block0, with_kwargs=True)]
    for blk in m.blocks:
        hooks.append(blk.mlp_adaln.register_forward_pre_hook(pre_mlp_adaln))
        hooks.append(blk.register_forward_hook(post_block))
    orig_forward = m.forward

    def fwd(*a, **k):
        call["i"] += 1
        if want() is not None:
            want()["x"], want()["t"] = k["x"].clone(), k["t"].clone()
        r = orig_forward(*a, **k)
        if want() is not None:
            want()["v"] = r.clone()
        return r

    m.forward = fwd
    t1 = time.time()
    lat = refinf.sample_euler_cfg_independent_guidances(m, spk, smask, ids, tmask, seed, sequence_length=640,
                                                        **HANDLER_KNOBS)
    dt = time.time() - t1
    for h in hooks:
        h.remove()
    print(f"{which}: reference fp32 sampler {dt:.1f} s on {torch.get_num_threads()} threads, {call['i'] + 1} model calls")
    p0, p1 = probes[0], probes[PLAIN_CALL]
    assert p1["x"].shape[0] == 1 and float(p1["t"][0]) < 0.5 and p0["x"].shape[0] == 3
    for name, p in (("step0", p0), ("plain", p1)):
        mid, out = torch.stack(p["mid"]), torch.stack(p["out"])
        xin = torch.cat([p["x0"][None], out[:-1]])
        ra = ((mid - xin).flatten(1).norm(dim=1) / xin.flatten(1).norm(dim=1))
        rm = ((out - mid).flatten(1).norm(dim=1) / mid.flatten(1).norm(dim=1))
        print(f"  {name}: |attention increment| / |stream| per block: min {ra.min():.3f} max {ra.max():.3f}; "
              f"|mlp increment| / |stream|: min {rm.min():.3f} max {rm.max():.3f}")
    res.update(latent=lat.clone(), step0_v=p0["v"], step0_layers_rows=torch.stack(p0["out"]),
               step0_x0_rows=p0["x0"], step0_mid_rows=torch.stack(p0["mid"]),
               plain_x=p1["x"], plain_t=p1["t"], plain_v=p1["v"], plain_x0_rows=p1["x0"],
               plain_mid_rows=torch.stack(p1["mid"]), plain_out_rows=torch.stack(p1["out"]),
               plain_call=torch.tensor(PLAIN_CALL), branch_gain=torch.tensor(gain),
               seconds=torch.tensor(dt), threads=torch.tensor(torch.get_num_threads()))
    torch.save(res, os.path.join(GOLD, f"dit_full_{which}.pt"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--full", choices=["cfg1", "cfg2", "cfg2t", "cfg5", "dac", "dac640"])
    ap.add_argument("--encode", choices=["tiny", "full"])
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if a.tiny:
        run_tiny()
    if a.full:
        run_full(a.full)
    if a.encode:
        run_encode(a.encode)
