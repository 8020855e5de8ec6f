"""Deterministic synthetic checkpoints with the reference's state-dict keys and shapes.

The pretrained weights (jordand/echo-tts-base, jordand/fish-s1-dac-min; reference inference.py:14,56) are gated and
unreachable offline, so tests and bench use random-init weights of the same architecture. Every tensor is drawn
from its own generator seeded by crc32(key) ^ seed, so any subset can be regenerated anywhere (CPU torch RNG is
machine independent) and the GPU box reproduces exactly the tensors the golden fixtures were made from.

Values are rounded to bf16 (the model dtype on the B200 path) and returned as fp32, so the fp32 oracle/reference
and the bf16 CUDA path see IDENTICAL weight values; only arithmetic precision differs.
Key names/shapes follow reference model.py:472-559 and autoencoder.py (decode path only), see SURVEY.md 8(b).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterator, Tuple

import torch

from .config import DacConfig, DitConfig


def _gen(key: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _uniform(key, seed, shape, bound):
    return _bf16_round((torch.rand(shape, generator=_gen(key, seed)) * 2 - 1) * bound)


def _normal(key, seed, shape, std, mean=0.0):
    return _bf16_round(torch.randn(shape, generator=_gen(key, seed)) * std + mean)


# ----------------------------------------------------------------------------------------------- EchoDiT
def dit_param_specs(cfg: DitConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yields (key, shape, kind) for every EchoDiT parameter. kind: linear | bias:<fan_in> | norm | embed."""
    D, H, I, r = cfg.model_size, cfg.num_heads, cfg.intermediate_size, cfg.adaln_rank

    def encoder(prefix, E, heads, inter, layers):
        for i in range(layers):
            p = f"{prefix}.blocks.{i}"
            for w in ("wq", "wk", "wv", "wo", "gate"):
                yield f"{p}.attention.{w}.weight", (E, E), "linear"
            yield f"{p}.attention.q_norm.weight", (heads, E // heads), "norm"
            yield f"{p}.attention.k_norm.weight", (heads, E // heads), "norm"
            yield f"{p}.mlp.w1.weight", (inter, E), "linear"
            yield f"{p}.mlp.w3.weight", (inter, E), "linear"
            yield f"{p}.mlp.w2.weight", (E, inter), "linear"
            yield f"{p}.attention_norm.weight", (E,), "norm"
            yield f"{p}.mlp_norm.weight", (E,), "norm"

    Et, Es = cfg.text_model_size, cfg.speaker_model_size
    yield "text_encoder.text_embedding.weight", (cfg.text_vocab_size, Et), "embed"
    yield from encoder("text_encoder", Et, cfg.text_num_heads, cfg.text_intermediate_size, cfg.text_num_layers)
    for enc in ("speaker_encoder", "latent_encoder"):
        fan = cfg.latent_size * cfg.speaker_patch_size
        yield f"{enc}.in_proj.weight", (Es, fan), "linear"
        yield f"{enc}.in_proj.bias", (Es,), f"bias:{fan}"
        yield from encoder(enc, Es, cfg.speaker_num_heads, cfg.speaker_intermediate_size, cfg.speaker_num_layers)
    yield "text_norm.weight", (Et,), "norm"
    yield "speaker_norm.weight", (Es,), "norm"
    yield "latent_norm.weight", (Es,), "norm"
    yield "cond_module.0.weight", (D, cfg.timestep_embed_size), "linear"
    yield "cond_module.2.weight", (D, D), "linear"
    yield "cond_module.4.weight", (3 * D, D), "linear"
    yield "in_proj.weight", (D, cfg.latent_size), "linear"
    yield "in_proj.bias", (D,), f"bias:{cfg.latent_size}"
    for i in range(cfg.num_layers):
        p = f"blocks.{i}"
        for w in ("wq", "wk", "wv", "gate", "wo"):
            yield f"{p}.attention.{w}.weight", (D, D), "linear"
        for w in ("wk_text", "wv_text"):
            yield f"{p}.attention.{w}.weight", (D, Et), "linear"
        for w in ("wk_speaker", "wv_speaker", "wk_latent", "wv_latent"):
            yield f"{p}.attention.{w}.weight", (D, Es), "linear"
        yield f"{p}.attention.q_norm.weight", (H, D // H), "norm"
        yield f"{p}.attention.k_norm.weight", (H, D // H), "norm"
        yield f"{p}.mlp.w1.weight", (I, D), "linear"
        yield f"{p}.mlp.w3.weight", (I, D), "linear"
        yield f"{p}.mlp.w2.weight", (D, I), "linear"
        for ad in ("attention_adaln", "mlp_adaln"):
            for part in ("shift", "scale", "gate"):
                yield f"{p}.{ad}.{part}_down.weight", (r, D), "linear"
                yield f"{p}.{ad}.{part}_up.weight", (D, r), "linear"
                yield f"{p}.{ad}.{part}_up.bias", (D,), f"bias:{r}"
    yield "out_norm.weight", (D,), "norm"
    yield "out_proj.weight", (cfg.latent_size, D), "linear"
    yield "out_proj.bias", (cfg.latent_size,), f"bias:{D}"


def _make(key, shape, kind, seed):
    if kind == "linear":
        return _uniform(key, seed, shape, 1.0 / math.sqrt(shape[-1]))
    if kind.startswith("bias:"):
        return _uniform(key, seed, shape, 1.0 / math.sqrt(int(kind[5:])))
    if kind == "norm":
        return _normal(key, seed, shape, 0.1, 1.0)
    if kind == "embed":
        return _normal(key, seed, shape, 1.0)
    if kind.startswith("normal:"):
        return _normal(key, seed, shape, float(kind[7:]))
    if kind.startswith("scale:"):  # layer-scale style: c * (1 + 0.1 N)
        return _normal(key, seed, shape, 0.1 * float(kind[6:]), float(kind[6:]))
    if kind == "alpha":
        return _bf16_round(torch.exp(0.3 * torch.randn(shape, generator=_gen(key, seed))))
    raise ValueError(kind)


def iter_dit_weights(cfg: DitConfig, seed: int = 1234, include_latent: bool = True, branch_gain: float = 1.0):
    """`branch_gain` (a power of two, so the bf16 rounding of the values is unchanged) multiplies the output
    projections `attention.wo` / `mlp.w2` of the DiT blocks. With the default-style init each residual branch adds only
    ~1 % to the stream, which hides branch-level errors behind the stream itself; gain 32 gives "trained-like" blocks
    whose branches are O(0.3) of the stream (tests/test_full_gpu.py: dit_full_cfg2t golden)."""
    for key, shape, kind in dit_param_specs(cfg):
        if not include_latent and (key.startswith("latent_encoder.") or key.startswith("latent_norm")
                                   or ".wk_latent" in key or ".wv_latent" in key):
            continue  # mirrors delete_blockwise_modules (reference inference.py:28-34)
        w = _make(key, shape, kind, seed)
        if branch_gain != 1.0 and key.startswith("blocks.") and key.endswith(("attention.wo.weight", "mlp.w2.weight")):
            w = w * branch_gain
        yield key, w


def make_dit_weights(cfg: DitConfig, seed: int = 1234, include_latent: bool = True,
                     branch_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    return dict(iter_dit_weights(cfg, seed, include_latent, branch_gain))


# ----------------------------------------------------------------------------------------------- DAC (decode path)
def dac_param_specs(cfg: DacConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    C = cfg.latent_dim
    hd = C // cfg.post_heads
    # quantizer.post_module: window-limited causal transformer (autoencoder.py:744-802)
    for i in range(cfg.post_layers):
        p = f"quantizer.post_module.layers.{i}"
        yield f"{p}.attention.wqkv.weight", (3 * C, C), "linear"
        yield f"{p}.attention.wo.weight", (C, C), "linear"
        yield f"{p}.feed_forward.w1.weight", (cfg.post_intermediate, C), "linear"
        yield f"{p}.feed_forward.w3.weight", (cfg.post_intermediate, C), "linear"
        yield f"{p}.feed_forward.w2.weight", (C, cfg.post_intermediate), "linear"
        yield f"{p}.ffn_norm.weight", (C,), "norm"
        yield f"{p}.attention_norm.weight", (C,), "norm"
        yield f"{p}.attention_layer_scale.gamma", (C,), "scale:0.2"
        yield f"{p}.ffn_layer_scale.gamma", (C,), "scale:0.2"
    yield "quantizer.post_module.norm.weight", (C,), "norm"
    # quantizer.upsample: [ConvTranspose k2 s2 ; ConvNeXt] x num_upsample (autoencoder.py:427-435)
    M = cfg.convnext_mlp_ratio * C
    for i in range(cfg.num_upsample):
        p = f"quantizer.upsample.{i}"
        yield f"{p}.0.conv.weight", (C, C, 2), f"normal:{1.0 / math.sqrt(C)}"
        yield f"{p}.0.conv.bias", (C,), "normal:0.02"
        yield f"{p}.1.gamma", (C,), "scale:0.2"
        yield f"{p}.1.dwconv.conv.weight", (C, 1, 7), f"normal:{1.0 / math.sqrt(7)}"
        yield f"{p}.1.dwconv.conv.bias", (C,), "normal:0.02"
        yield f"{p}.1.norm.weight", (C,), "norm"
        yield f"{p}.1.norm.bias", (C,), "normal:0.1"
        yield f"{p}.1.pwconv1.weight", (M, C), "linear"
        yield f"{p}.1.pwconv1.bias", (M,), f"bias:{C}"
        yield f"{p}.1.pwconv2.weight", (C, M), "linear"
        yield f"{p}.1.pwconv2.bias", (C,), f"bias:{M}"

    # decoder (autoencoder.py:971-998): weight-normed convs keep (g, v) = (original0, original1)
    ch = cfg.decoder_dim
    yield "decoder.model.0.conv.bias", (ch,), "normal:0.02"
    yield "decoder.model.0.conv.parametrizations.weight.original0", (ch, 1, 1), "wn_g"
    yield "decoder.model.0.conv.parametrizations.weight.original1", (ch, C, 7), f"normal:{1.0 / math.sqrt(7 * C)}"
    for bi, stride in enumerate(cfg.rates):
        cin, cout = ch // 2 ** bi, ch // 2 ** (bi + 1)
        p = f"decoder.model.{bi + 1}.block"
        yield f"{p}.0.alpha", (1, cin, 1), "alpha"
        yield f"{p}.1.conv.bias", (cout,), "normal:0.02"
        yield f"{p}.1.conv.parametrizations.weight.original0", (cin, 1, 1), "wn_g"
        yield f"{p}.1.conv.parametrizations.weight.original1", (cin, cout, 2 * stride), f"normal:{1.0 / math.sqrt(2 * cin)}"
        for ui in range(3):
            q = f"{p}.{ui + 2}.block"
            yield f"{q}.0.alpha", (1, cout, 1), "alpha"
            yield f"{q}.1.conv.bias", (cout,), "normal:0.02"
            yield f"{q}.1.conv.parametrizations.weight.original0", (cout, 1, 1), "wn_g"
            yield f"{q}.1.conv.parametrizations.weight.original1", (cout, cout, 7), f"normal:{1.0 / math.sqrt(7 * cout)}"
            yield f"{q}.2.alpha", (1, cout, 1), "alpha"
            yield f"{q}.3.conv.bias", (cout,), "normal:0.02"
            yield f"{q}.3.conv.parametrizations.weight.original0", (cout, 1, 1), "wn_g"
            yield f"{q}.3.conv.parametrizations.weight.original1", (cout, cout, 1), f"normal:{0.3 / math.sqrt(cout)}"
    n = len(cfg.rates)
    clast = ch // 2 ** n
    yield f"decoder.model.{n + 1}.alpha", (1, clast, 1), "alpha"
    yield f"decoder.model.{n + 2}.conv.bias", (1,), "normal:0.02"
    yield f"decoder.model.{n + 2}.conv.parametrizations.weight.original0", (1, 1, 1), "wn_g"
    yield f"decoder.model.{n + 2}.conv.parametrizations.weight.original1", (1, clast, 7), f"normal:{0.2 / math.sqrt(7 * clast)}"


def dac_encoder_param_specs(cfg: DacConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Encode-path parameters of build_ae(): Encoder (autoencoder.py:903-929), quantizer.downsample / pre_module and
    the semantic + residual vector quantizers (:117-157, 376-440)."""
    def wn_conv(p, cout, cin, k, std):
        yield f"{p}.conv.bias", (cout,), "normal:0.02"
        yield f"{p}.conv.parametrizations.weight.original0", (cout, 1, 1), "wn_g"
        yield f"{p}.conv.parametrizations.weight.original1", (cout, cin, k), f"normal:{std}"

    def transformer(p, Cd, layers, inter):
        for i in range(layers):
            q = f"{p}.layers.{i}"
            yield f"{q}.attention.wqkv.weight", (3 * Cd, Cd), "linear"
            yield f"{q}.attention.wo.weight", (Cd, Cd), "linear"
            yield f"{q}.feed_forward.w1.weight", (inter, Cd), "linear"
            yield f"{q}.feed_forward.w3.weight", (inter, Cd), "linear"
            yield f"{q}.feed_forward.w2.weight", (Cd, inter), "linear"
            yield f"{q}.ffn_norm.weight", (Cd,), "norm"
            yield f"{q}.attention_norm.weight", (Cd,), "norm"
            yield f"{q}.attention_layer_scale.gamma", (Cd,), "scale:0.2"
            yield f"{q}.ffn_layer_scale.gamma", (Cd,), "scale:0.2"
        yield f"{p}.norm.weight", (Cd,), "norm"

    d = cfg.enc_dim
    yield from wn_conv("encoder.block.0", d, 1, 7, 1.0 / math.sqrt(7))
    nb = len(cfg.enc_rates)
    for bi, stride in enumerate(cfg.enc_rates):
        cin, d = d, d * 2
        p = f"encoder.block.{bi + 1}.block"
        for ui in range(3):
            q = f"{p}.{ui}.block"
            yield f"{q}.0.alpha", (1, cin, 1), "alpha"
            yield from wn_conv(f"{q}.1", cin, cin, 7, 1.0 / math.sqrt(7 * cin))
            yield f"{q}.2.alpha", (1, cin, 1), "alpha"
            yield from wn_conv(f"{q}.3", cin, cin, 1, 0.3 / math.sqrt(cin))
        yield f"{p}.3.alpha", (1, cin, 1), "alpha"
        yield from wn_conv(f"{p}.4", d, cin, 2 * stride, 1.0 / math.sqrt(2 * stride * cin))
        if bi == nb - 1 and cfg.enc_t_layers > 0:
            yield from transformer(f"{p}.5", d, cfg.enc_t_layers, 3 * d)
    assert d == cfg.latent_dim, "latent_dim must equal enc_dim * 2 ** len(enc_rates)"
    C = cfg.latent_dim
    yield f"encoder.block.{nb + 1}.alpha", (1, C, 1), "alpha"
    yield from wn_conv(f"encoder.block.{nb + 2}", C, C, 3, 1.0 / math.sqrt(3 * C))
    M = cfg.convnext_mlp_ratio * C
    for i in range(cfg.num_upsample):
        p = f"quantizer.downsample.{i}"
        yield f"{p}.0.conv.weight", (C, C, 2), f"normal:{1.0 / math.sqrt(2 * C)}"
        yield f"{p}.0.conv.bias", (C,), "normal:0.02"
        yield f"{p}.1.gamma", (C,), "scale:0.2"
        yield f"{p}.1.dwconv.conv.weight", (C, 1, 7), f"normal:{1.0 / math.sqrt(7)}"
        yield f"{p}.1.dwconv.conv.bias", (C,), "normal:0.02"
        yield f"{p}.1.norm.weight", (C,), "norm"
        yield f"{p}.1.norm.bias", (C,), "normal:0.1"
        yield f"{p}.1.pwconv1.weight", (M, C), "linear"
        yield f"{p}.1.pwconv1.bias", (M,), f"bias:{C}"
        yield f"{p}.1.pwconv2.weight", (C, M), "linear"
        yield f"{p}.1.pwconv2.bias", (C,), f"bias:{M}"
    yield from transformer("quantizer.pre_module", C, cfg.post_layers, cfg.post_intermediate)
    for name, n, size in (("semantic_quantizer", 1, cfg.semantic_codebook_size), ("quantizer", cfg.n_codebooks, cfg.codebook_size)):
        for i in range(n):
            p = f"quantizer.{name}.quantizers.{i}"
            yield f"{p}.in_proj.bias", (cfg.codebook_dim,), "normal:0.02"
            yield f"{p}.in_proj.parametrizations.weight.original0", (cfg.codebook_dim, 1, 1), "wn_g"
            yield f"{p}.in_proj.parametrizations.weight.original1", (cfg.codebook_dim, C, 1), f"normal:{1.0 / math.sqrt(C)}"
            yield f"{p}.out_proj.bias", (C,), "normal:0.02"
            yield f"{p}.out_proj.parametrizations.weight.original0", (C, 1, 1), "wn_g"
            yield f"{p}.out_proj.parametrizations.weight.original1", (C, cfg.codebook_dim, 1), f"normal:{1.0 / math.sqrt(cfg.codebook_dim)}"
            yield f"{p}.codebook.weight", (size, cfg.codebook_dim), "embed"


def make_dac_weights(cfg: DacConfig, seed: int = 4321, include_encoder: bool = False) -> Dict[str, torch.Tensor]:
    """Decode-path parameters of build_ae() (post_module, upsample, decoder); with include_encoder also the encode
    path (Encoder, quantizer.downsample / pre_module, the vector quantizers)."""
    out: Dict[str, torch.Tensor] = {}
    specs = list(dac_param_specs(cfg)) + (list(dac_encoder_param_specs(cfg)) if include_encoder else [])
    for key, shape, kind in specs:
        if kind != "wn_g":
            out[key] = _make(key, shape, kind, seed)
    for key, shape, kind in specs:
        if kind == "wn_g":
            # weight_norm: w = g * v / ||v|| with the norm over all dims but 0 (autoencoder.py:90-94, 291-293).
            v = out[key.replace("original0", "original1")]
            nrm = v.flatten(1).norm(dim=1).view(shape)
            jitter = 1.0 + 0.1 * torch.randn(shape, generator=_gen(key, seed))
            out[key] = _bf16_round(nrm * jitter)
    return out


def make_pca_state(cfg: DacConfig, seed: int = 99):
    """Synthetic PCA state with the reference's fields (inference.py:86-99): orthonormal-ish components."""
    g = _gen("pca", seed)
    q, _ = torch.linalg.qr(torch.randn(cfg.latent_dim, cfg.pca_dim, generator=g))
    comps = q.T.contiguous()  # (80, 1024)
    mean = 0.1 * torch.randn(cfg.latent_dim, generator=g)
    return comps.float(), mean.float(), 0.35
