"""CPU oracle for the Echo-TTS sampling hot path -- TEST INFRASTRUCTURE ONLY.

A functional, fp32, plain-PyTorch restatement of the reference algorithm over a flat state dict (reference key
names). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (echo_tts_b200/) never does and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against the reference ITSELF: oracle/pin_reference.py imports /root/reference (model.py, inference.py,
inference_blockwise.py, autoencoder.py), runs both on identical weights/inputs, asserts agreement to fp32
round-off, and writes the fixtures under tests/golden/ that tests/test_oracle.py re-checks without the reference.

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

SD = Dict[str, torch.Tensor]
KV = List[Tuple[torch.Tensor, torch.Tensor]]


# ------------------------------------------------------------------------------------------------ primitives
def rope_table(head_dim: int, n_pos: int, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin of pos / theta^(2i/head_dim)  (model.py:9-14)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2] / head_dim))
    ang = torch.outer(torch.arange(n_pos), inv)
    return torch.cos(ang), torch.sin(ang)


def rope_rotate(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x (b, s, h, d); (x[2i], x[2i+1]) rotated as a complex number by angle pos*freq_i  (model.py:17-24)."""
    a, b = x.float()[..., 0::2], x.float()[..., 1::2]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    out = torch.stack((a * c - b * s, a * s + b * c), dim=-1)
    return out.flatten(-2).to(x.dtype)


def rms_norm(x: torch.Tensor, weight: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """x * rsqrt(mean(x^2) + eps) [* weight]  (model.py:99-104; LowRankAdaLN uses it without a weight, :78)."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    if weight is not None:
        y = y * weight
    return y.to(x.dtype)


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    y = x @ w.T
    return y if b is None else y + b


def silu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)


def masked_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """softmax(q k^T / sqrt(d) + mask) v with q,k,v (b, s, h, d) and a boolean mask broadcastable to (b, h, sq, sk);
    what F.scaled_dot_product_attention computes at model.py:148-154, 255-261 and autoencoder.py:698-702."""
    d = q.shape[-1]
    logits = torch.einsum("bqhd,bkhd->bhqk", q, k) / math.sqrt(d)
    if mask is not None:
        logits = logits.masked_fill(~mask, float("-inf"))
    p = torch.softmax(logits, dim=-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v)


def timestep_embedding(t: torch.Tensor, size: int) -> torch.Tensor:
    """[cos, sin](t * 1000 * exp(-ln(1e4) * i / half)), cast to t.dtype  (model.py:27-43)."""
    half = size // 2
    freqs = 1000 * torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[..., None] * freqs[None]
    return torch.cat((torch.cos(args), torch.sin(args)), dim=-1).to(t.dtype)


# ------------------------------------------------------------------------------------------------ EchoDiT
def low_rank_adaln(sd: SD, p: str, x: torch.Tensor, cond: torch.Tensor, eps: float):
    """model.py:64-83: three low-rank residual MLPs on the thirds of cond; weightless RMS norm of x; modulate; tanh gate."""
    parts = cond.chunk(3, dim=-1)
    mod = []
    for name, c in zip(("shift", "scale", "gate"), parts):
        h = linear(silu(c), sd[f"{p}.{name}_down.weight"])
        mod.append(linear(h, sd[f"{p}.{name}_up.weight"], sd[f"{p}.{name}_up.bias"]) + c)
    shift, scale, gate = mod
    xn = rms_norm(x, None, eps).float() * (scale + 1) + shift
    return xn.to(x.dtype), torch.tanh(gate)


def encoder_block(sd: SD, p: str, x: torch.Tensor, key_mask: Optional[torch.Tensor], causal: bool, heads: int,
                  cos: torch.Tensor, sin: torch.Tensor, eps: float) -> torch.Tensor:
    """Pre-norm block of the text / speaker / latent encoders (model.py:311-339 with SelfAttention :128-161)."""
    b, s, _ = x.shape
    h = rms_norm(x, sd[f"{p}.attention_norm.weight"], eps)
    a = f"{p}.attention"
    q = linear(h, sd[f"{a}.wq.weight"]).view(b, s, heads, -1)
    k = linear(h, sd[f"{a}.wk.weight"]).view(b, s, heads, -1)
    v = linear(h, sd[f"{a}.wv.weight"]).view(b, s, heads, -1)
    g = linear(h, sd[f"{a}.gate.weight"])
    q = rope_rotate(rms_norm(q, sd[f"{a}.q_norm.weight"], eps), cos[:s], sin[:s])  # RoPE on ALL heads here
    k = rope_rotate(rms_norm(k, sd[f"{a}.k_norm.weight"], eps), cos[:s], sin[:s])
    mask = None
    if key_mask is not None:
        mask = key_mask[:, None, None, :]
    if causal:
        tri = torch.ones(s, s, dtype=torch.bool).tril()[None, None]
        mask = tri if mask is None else (mask & tri)
    o = masked_attention(q, k, v, mask).reshape(b, s, -1) * torch.sigmoid(g)
    x = x + linear(o, sd[f"{a}.wo.weight"])
    h = rms_norm(x, sd[f"{p}.mlp_norm.weight"], eps)
    m = f"{p}.mlp"
    return x + linear(silu(linear(h, sd[f"{m}.w1.weight"])) * linear(h, sd[f"{m}.w3.weight"]), sd[f"{m}.w2.weight"])


def _num_blocks(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.{n}.attention.wq.weight" in sd:
        n += 1
    return n


def _k_norm_proj(sd: SD, i: int, state: torch.Tensor, which: str, heads: int, eps: float):
    """model.py:270-293: K = k_norm(wk_x state), V = wv_x state, both (b, L, heads, 128)."""
    b, L, _ = state.shape
    a = f"blocks.{i}.attention"
    k = linear(state, sd[f"{a}.wk_{which}.weight"]).view(b, L, heads, -1)
    v = linear(state, sd[f"{a}.wv_{which}.weight"]).view(b, L, heads, -1)
    return rms_norm(k, sd[f"{a}.k_norm.weight"], eps), v


def kv_cache_text(sd: SD, cfg, ids: torch.Tensor, mask: Optional[torch.Tensor]) -> KV:
    """model.py:606-613 + TextEncoder :419-427."""
    x = sd["text_encoder.text_embedding.weight"][ids.long()]
    hd = cfg.text_model_size // cfg.text_num_heads
    cos, sin = rope_table(hd, ids.shape[1])
    for i in range(cfg.text_num_layers):
        x = encoder_block(sd, f"text_encoder.blocks.{i}", x, mask, False, cfg.text_num_heads, cos, sin, cfg.norm_eps)
    x = rms_norm(x, sd["text_norm.weight"], cfg.norm_eps)
    return [_k_norm_proj(sd, i, x, "text", cfg.num_heads, cfg.norm_eps) for i in range(cfg.num_layers)]


def _patch_encoder(sd: SD, cfg, prefix: str, latent: torch.Tensor) -> torch.Tensor:
    """SpeakerEncoder.forward (model.py:458-469): patchify x4, in_proj, /6, causal blocks."""
    ps = cfg.speaker_patch_size
    b, L, C = latent.shape
    x = latent.reshape(b, L // ps, C * ps)
    x = linear(x, sd[f"{prefix}.in_proj.weight"], sd[f"{prefix}.in_proj.bias"]) / 6.0
    hd = cfg.speaker_model_size // cfg.speaker_num_heads
    cos, sin = rope_table(hd, x.shape[1])
    for i in range(cfg.speaker_num_layers):
        x = encoder_block(sd, f"{prefix}.blocks.{i}", x, None, True, cfg.speaker_num_heads, cos, sin, cfg.norm_eps)
    return x


def kv_cache_speaker(sd: SD, cfg, speaker_latent: torch.Tensor) -> KV:
    """model.py:615-621."""
    x = rms_norm(_patch_encoder(sd, cfg, "speaker_encoder", speaker_latent), sd["speaker_norm.weight"], cfg.norm_eps)
    return [_k_norm_proj(sd, i, x, "speaker", cfg.num_heads, cfg.norm_eps) for i in range(cfg.num_layers)]


def _rope_first_half_heads(y: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """JointAttention._apply_rotary_half (model.py:199-202): the split is over HEADS (dim -2), not head_dim."""
    h = y.shape[-2] // 2
    return torch.cat((rope_rotate(y[..., :h, :], cos, sin), y[..., h:, :]), dim=-2)


def kv_cache_latent(sd: SD, cfg, prefix_latent: torch.Tensor) -> KV:
    """model.py:623-636: as the speaker path with latent_* weights, plus RoPE (first half of heads) at positions 4j."""
    x = rms_norm(_patch_encoder(sd, cfg, "latent_encoder", prefix_latent), sd["latent_norm.weight"], cfg.norm_eps)
    n = x.shape[1]
    cos, sin = rope_table(cfg.model_size // cfg.num_heads, max(n * cfg.speaker_patch_size, 1))
    pos = torch.arange(n) * cfg.speaker_patch_size
    out = []
    for i in range(cfg.num_layers):
        k, v = _k_norm_proj(sd, i, x, "latent", cfg.num_heads, cfg.norm_eps)
        out.append((_rope_first_half_heads(k, cos[pos], sin[pos]), v))
    return out


def dit_forward(sd: SD, cfg, x: torch.Tensor, t: torch.Tensor, text_mask: torch.Tensor, speaker_mask: torch.Tensor,
                kv_text: KV, kv_speaker: KV, start_pos: Optional[int] = None, kv_latent: Optional[KV] = None,
                layer_outputs: Optional[list] = None) -> torch.Tensor:
    """EchoDiT.forward (model.py:563-604) with TransformerBlock (:371-390) and JointAttention (:204-268)."""
    start_pos = start_pos or 0
    b, S, _ = x.shape
    H, eps = cfg.num_heads, cfg.norm_eps
    cos, sin = rope_table(cfg.model_size // H, start_pos + S)
    cos, sin = cos[start_pos:start_pos + S], sin[start_pos:start_pos + S]
    spk_mask = speaker_mask[..., :: cfg.speaker_patch_size]
    temb = timestep_embedding(t, cfg.timestep_embed_size)
    c = linear(silu(linear(silu(linear(temb, sd["cond_module.0.weight"])), sd["cond_module.2.weight"])),
               sd["cond_module.4.weight"])[:, None]
    x = linear(x, sd["in_proj.weight"], sd["in_proj.bias"])
    for i in range(cfg.num_layers):
        p = f"blocks.{i}"
        a = f"{p}.attention"
        xn, gate_a = low_rank_adaln(sd, f"{p}.attention_adaln", x, c, eps)
        q = linear(xn, sd[f"{a}.wq.weight"]).view(b, S, H, -1)
        k = linear(xn, sd[f"{a}.wk.weight"]).view(b, S, H, -1)
        v = linear(xn, sd[f"{a}.wv.weight"]).view(b, S, H, -1)
        g = linear(xn, sd[f"{a}.gate.weight"])
        q = _rope_first_half_heads(rms_norm(q, sd[f"{a}.q_norm.weight"], eps), cos, sin)
        k = _rope_first_half_heads(rms_norm(k, sd[f"{a}.k_norm.weight"], eps), cos, sin)
        keys, vals = [k], [v]
        masks = [torch.ones(b, S, dtype=torch.bool)]
        if kv_latent is not None and kv_latent[i][0].shape[1] > 0:
            kl, vl = kv_latent[i]
            keys.append(kl); vals.append(vl)
            pos = torch.arange(kl.shape[1]) * cfg.speaker_patch_size
            masks.append((pos[None] < start_pos).expand(b, -1))
        keys += [kv_text[i][0], kv_speaker[i][0]]
        vals += [kv_text[i][1], kv_speaker[i][1]]
        masks += [text_mask, spk_mask]
        o = masked_attention(q, torch.cat(keys, 1), torch.cat(vals, 1), torch.cat(masks, 1)[:, None, None, :])
        o = o.reshape(b, S, -1) * torch.sigmoid(g)
        x = x + gate_a * linear(o, sd[f"{a}.wo.weight"])
        xn, gate_m = low_rank_adaln(sd, f"{p}.mlp_adaln", x, c, eps)
        m = f"{p}.mlp"
        x = x + gate_m * linear(silu(linear(xn, sd[f"{m}.w1.weight"])) * linear(xn, sd[f"{m}.w3.weight"]),
                                sd[f"{m}.w2.weight"])
        if layer_outputs is not None:
            layer_outputs.append(x.clone())
    x = rms_norm(x, sd["out_norm.weight"], eps)
    return linear(x, sd["out_proj.weight"], sd["out_proj.bias"]).float()


# ------------------------------------------------------------------------------------------------ samplers
def _batch3(cache: KV) -> KV:
    """_concat_kv_caches(c, c, c) (inference.py:398-406)."""
    return [(torch.cat((k, k, k), 0), torch.cat((v, v, v), 0)) for k, v in cache]


def _scale_kv(cache: KV, s: float, max_layers: Optional[int]) -> None:
    """_multiply_kv_cache (inference.py:408-414): in place, first max_layers layers."""
    n = len(cache) if max_layers is None else min(max_layers, len(cache))
    for i in range(n):
        cache[i][0].mul_(s)
        cache[i][1].mul_(s)


def temporal_score_rescale(v: torch.Tensor, x: torch.Tensor, t: float, k: float, sigma: float) -> torch.Tensor:
    """inference.py:416-424."""
    if t < 1:
        snr = (1 - t) ** 2 / (t ** 2)
        ratio = (snr * sigma ** 2 + 1) / (snr * sigma ** 2 / k + 1)
        return 1 / (1 - t) * (ratio * ((1 - t) * v + x) - x)
    return v


def _euler_loop(sd, cfg, x_t, t_sched, num_steps, text_mask, speaker_mask, kv_text, kv_speaker, kv_text3, kv_speaker3,
                mask_t3, mask_s3, cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t, rescale_k, rescale_sigma,
                speaker_kv_scale, speaker_kv_max_layers, speaker_kv_min_t, start_pos=None, kv_latent=None,
                kv_latent3=None, t_dtype=None):
    """The 40-step loop shared by inference.py:481-515 and inference_blockwise.py:80-118."""
    B = x_t.shape[0]
    for i in range(num_steps):
        t, t_next = t_sched[i], t_sched[i + 1]
        tt = t if t_dtype is None else t.to(t_dtype).float()
        if bool((t >= cfg_min_t) * (t <= cfg_max_t)):
            v = dit_forward(sd, cfg, torch.cat((x_t, x_t, x_t), 0), torch.ones(3 * B) * tt, mask_t3, mask_s3,
                            kv_text3, kv_speaker3, start_pos, kv_latent3)
            vc, vt, vs = v.chunk(3, 0)
            v = vc + cfg_scale_text * (vc - vt) + cfg_scale_speaker * (vc - vs)
        else:
            v = dit_forward(sd, cfg, x_t, torch.ones(B) * tt, text_mask, speaker_mask, kv_text, kv_speaker,
                            start_pos, kv_latent)
        if rescale_k is not None and rescale_sigma is not None:
            v = temporal_score_rescale(v, x_t, t, rescale_k, rescale_sigma)
        if speaker_kv_scale is not None and t_next < speaker_kv_min_t and t >= speaker_kv_min_t:
            _scale_kv(kv_speaker, 1.0 / speaker_kv_scale, speaker_kv_max_layers)
            kv_speaker3 = _batch3(kv_speaker)
        x_t = x_t + v * (t_next - t)
    return x_t, kv_speaker3


def sample_euler_cfg_independent_guidances(sd, cfg, speaker_latent, speaker_mask, text_ids, text_mask, noise, *,
                                           num_steps, cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t,
                                           truncation_factor=None, rescale_k=None, rescale_sigma=None,
                                           speaker_kv_scale=None, speaker_kv_max_layers=None, speaker_kv_min_t=None,
                                           t_dtype=None) -> torch.Tensor:
    """inference.py:427-517 with the initial noise INJECTED (the reference draws it from a device generator, :457,477).
    `noise` (B, S, 80) fp32 replaces torch.randn; everything else follows the reference step for step.
    t_dtype=torch.bfloat16 reproduces the reference's `.to(model.dtype)` rounding of t (:489) inside an fp32 run."""
    t_sched = torch.linspace(1.0, 0.0, num_steps + 1) * 0.999
    kv_t = kv_cache_text(sd, cfg, text_ids, text_mask)
    kv_s = kv_cache_speaker(sd, cfg, speaker_latent)
    if speaker_kv_scale is not None:
        _scale_kv(kv_s, speaker_kv_scale, speaker_kv_max_layers)
    mask_t3 = torch.cat((text_mask, torch.zeros_like(text_mask), text_mask), 0)
    mask_s3 = torch.cat((speaker_mask, speaker_mask, torch.zeros_like(speaker_mask)), 0)
    x_t = noise.clone().float()
    if truncation_factor is not None:
        x_t = x_t * truncation_factor
    x_t, _ = _euler_loop(sd, cfg, x_t, t_sched, num_steps, text_mask, speaker_mask, kv_t, kv_s, _batch3(kv_t),
                         _batch3(kv_s), mask_t3, mask_s3, cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t,
                         rescale_k, rescale_sigma, speaker_kv_scale, speaker_kv_max_layers, speaker_kv_min_t,
                         t_dtype=t_dtype)
    return x_t


def sample_blockwise_euler_cfg_independent_guidances(sd, cfg, speaker_latent, speaker_mask, text_ids, text_mask,
                                                     noise_blocks: Sequence[torch.Tensor], *, block_sizes, num_steps,
                                                     cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t,
                                                     truncation_factor=None, rescale_k=None, rescale_sigma=None,
                                                     speaker_kv_scale=None, speaker_kv_max_layers=None,
                                                     speaker_kv_min_t=None, continuation_latent=None,
                                                     t_dtype=None) -> torch.Tensor:
    """inference_blockwise.py:15-123, noise per block injected. Keeps the reference's quirk that the speaker-KV scale
    is re-applied at the start of every block (:68-70) and only undone when the min_t threshold is crossed."""
    B = text_ids.shape[0]
    t_sched = torch.linspace(1.0, 0.0, num_steps + 1) * 0.999
    kv_t = kv_cache_text(sd, cfg, text_ids, text_mask)
    kv_s = kv_cache_speaker(sd, cfg, speaker_latent)
    kv_t3, kv_s3 = _batch3(kv_t), _batch3(kv_s)
    mask_t3 = torch.cat((text_mask, torch.zeros_like(text_mask), text_mask), 0)
    mask_s3 = torch.cat((speaker_mask, speaker_mask, torch.zeros_like(speaker_mask)), 0)
    prefix = torch.zeros(B, sum(block_sizes), cfg.latent_size)
    start = 0
    if continuation_latent is not None:
        prefix = torch.cat((continuation_latent.float(), prefix), 1)
        start = continuation_latent.shape[1]
    for bs, noise in zip(block_sizes, noise_blocks):
        if speaker_kv_scale is not None:
            _scale_kv(kv_s, speaker_kv_scale, speaker_kv_max_layers)
            kv_s3 = _batch3(kv_s)
        kv_l3 = kv_cache_latent(sd, cfg, torch.cat((prefix, prefix, prefix), 0))
        kv_l = [(k[:B], v[:B]) for k, v in kv_l3]
        x_t = noise.clone().float()
        if truncation_factor is not None:
            x_t = x_t * truncation_factor
        x_t, kv_s3 = _euler_loop(sd, cfg, x_t, t_sched, num_steps, text_mask, speaker_mask, kv_t, kv_s, kv_t3, kv_s3,
                                 mask_t3, mask_s3, cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t, rescale_k,
                                 rescale_sigma, speaker_kv_scale, speaker_kv_max_layers, speaker_kv_min_t,
                                 start_pos=start, kv_latent=kv_l, kv_latent3=kv_l3, t_dtype=t_dtype)
        prefix[:, start:start + bs] = x_t
        start += bs
    return prefix


# ------------------------------------------------------------------------------------------------ Fish S1-DAC decode
def _wn(sd: SD, p: str) -> torch.Tensor:
    """weight_norm: g * v / ||v||, norm over every dim but 0 (torch parametrizations; autoencoder.py:90-94, 291-293)."""
    g, v = sd[f"{p}.parametrizations.weight.original0"], sd[f"{p}.parametrizations.weight.original1"]
    return v * (g / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1))))


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """x + sin^2(alpha x) / (alpha + 1e-9)  (autoencoder.py:96-102)."""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def causal_conv1d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, dilation: int = 1, groups: int = 1) -> torch.Tensor:
    """CausalConvNet.forward, stride 1: left pad (k-1)*dilation zeros (autoencoder.py:285-289)."""
    pad = (w.shape[-1] - 1) * dilation
    return torch.nn.functional.conv1d(torch.nn.functional.pad(x, (pad, 0)), w, b, dilation=dilation, groups=groups)


def causal_conv_transpose1d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int) -> torch.Tensor:
    """CausalTransConvNet.forward (autoencoder.py:310-316): full transposed conv, then trim (k - stride) on the right."""
    y = torch.nn.functional.conv_transpose1d(x, w, b, stride=stride)
    trim = w.shape[-1] - stride
    return y[..., : y.shape[-1] - trim] if trim > 0 else y


def dac_rope_table(n_pos: int, head_dim: int, base: float = 10000.0):
    """autoencoder.py:805-813: the cos/sin cache is stored in BFLOAT16; values are used in fp32 math (:815-826)."""
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(n_pos), inv)
    return torch.cos(ang).to(torch.bfloat16).float(), torch.sin(ang).to(torch.bfloat16).float()


def dac_window_transformer(sd: SD, prefix: str, z: torch.Tensor, layers: int, heads: int, window: int,
                           eps: float) -> torch.Tensor:
    """WindowLimitedTransformer.forward (autoencoder.py:786-802) over (B, C, T): causal window attention
    (:762-773), RoPE on all heads with the bf16 table (:805-826), LayerScale residuals (:621-626), final RMSNorm
    (:607). Used for quantizer.post_module / pre_module (window 128) and the last EncoderBlock (window 512)."""
    x = z.transpose(1, 2)
    B, T, C = x.shape
    hd = C // heads
    cos, sin = dac_rope_table(T, hd)
    i = torch.arange(T)
    mask = ((i[None, :] <= i[:, None]) & (i[None, :] >= (i[:, None] - window + 1).clamp(min=0)))[None, None]
    for li in range(layers):
        p = f"{prefix}.layers.{li}"
        h = rms_norm(x, None, eps) * sd[f"{p}.attention_norm.weight"]
        q, k, v = linear(h, sd[f"{p}.attention.wqkv.weight"]).split(C, dim=-1)
        q = rope_rotate(q.reshape(B, T, heads, hd), cos, sin)
        k = rope_rotate(k.reshape(B, T, heads, hd), cos, sin)
        o = masked_attention(q, k, v.reshape(B, T, heads, hd), mask).reshape(B, T, C)
        x = x + linear(o, sd[f"{p}.attention.wo.weight"]) * sd[f"{p}.attention_layer_scale.gamma"]
        h = rms_norm(x, None, eps) * sd[f"{p}.ffn_norm.weight"]
        f = p + ".feed_forward"
        y = linear(silu(linear(h, sd[f"{f}.w1.weight"])) * linear(h, sd[f"{f}.w3.weight"]), sd[f"{f}.w2.weight"])
        x = x + y * sd[f"{p}.ffn_layer_scale.gamma"]
    x = rms_norm(x, None, eps) * sd[f"{prefix}.norm.weight"]
    return x.transpose(1, 2)


def dac_post_module(sd: SD, cfg, z: torch.Tensor) -> torch.Tensor:
    return dac_window_transformer(sd, "quantizer.post_module", z, cfg.post_layers, cfg.post_heads, cfg.post_window,
                                  cfg.post_norm_eps)


def dac_upsample(sd: SD, cfg, z: torch.Tensor) -> torch.Tensor:
    """quantizer.upsample (autoencoder.py:427-435): per stage ConvTranspose(k=2,s=2) then ConvNeXtBlock (:360-373)."""
    for i in range(cfg.num_upsample):
        p = f"quantizer.upsample.{i}"
        z = causal_conv_transpose1d(z, sd[f"{p}.0.conv.weight"], sd[f"{p}.0.conv.bias"], 2)
        C = z.shape[1]
        y = causal_conv1d(z, sd[f"{p}.1.dwconv.conv.weight"], sd[f"{p}.1.dwconv.conv.bias"], groups=C).transpose(1, 2)
        y = torch.nn.functional.layer_norm(y, (C,), sd[f"{p}.1.norm.weight"], sd[f"{p}.1.norm.bias"], eps=1e-6)
        y = linear(torch.nn.functional.gelu(linear(y, sd[f"{p}.1.pwconv1.weight"], sd[f"{p}.1.pwconv1.bias"])),
                   sd[f"{p}.1.pwconv2.weight"], sd[f"{p}.1.pwconv2.bias"])
        z = z + (y * sd[f"{p}.1.gamma"]).transpose(1, 2)
    return z


def dac_decoder(sd: SD, cfg, z: torch.Tensor) -> torch.Tensor:
    """Decoder.model (autoencoder.py:984-998). DecoderBlock (:959-965) = Snake, WN-ConvTranspose(k=2s, stride s, right
    trim s), three ResidualUnits (:884-900) with dilations 1, 3, 9. The block's transformer_module is constructed but
    never added to the Sequential, so there is none here either."""
    x = causal_conv1d(z, _wn(sd, "decoder.model.0.conv"), sd["decoder.model.0.conv.bias"])
    for bi, stride in enumerate(cfg.rates):
        p = f"decoder.model.{bi + 1}.block"
        x = snake(x, sd[f"{p}.0.alpha"])
        x = causal_conv_transpose1d(x, _wn(sd, f"{p}.1.conv"), sd[f"{p}.1.conv.bias"], stride)
        for ui, dil in enumerate((1, 3, 9)):
            q = f"{p}.{ui + 2}.block"
            y = snake(x, sd[f"{q}.0.alpha"])
            y = causal_conv1d(y, _wn(sd, f"{q}.1.conv"), sd[f"{q}.1.conv.bias"], dilation=dil)
            y = snake(y, sd[f"{q}.2.alpha"])
            y = causal_conv1d(y, _wn(sd, f"{q}.3.conv"), sd[f"{q}.3.conv.bias"])
            x = x + y
    n = len(cfg.rates)
    x = snake(x, sd[f"decoder.model.{n + 1}.alpha"])
    x = causal_conv1d(x, _wn(sd, f"decoder.model.{n + 2}.conv"), sd[f"decoder.model.{n + 2}.conv.bias"])
    return torch.tanh(x)


def dac_decode_zq(sd: SD, cfg, zq: torch.Tensor) -> torch.Tensor:
    """DAC.decode_zq (autoencoder.py:1128-1132): (B, 1024, T) -> (B, 1, 2048 T)."""
    return dac_decoder(sd, cfg, dac_upsample(sd, cfg, dac_post_module(sd, cfg, zq)))


def ae_decode(sd: SD, cfg, pca_components: torch.Tensor, pca_mean: torch.Tensor, latent_scale: float,
              z: torch.Tensor) -> torch.Tensor:
    """inference.ae_decode (inference.py:226-229): PCA un-projection, then decode_zq."""
    zq = (z / latent_scale) @ pca_components + pca_mean
    return dac_decode_zq(sd, cfg, zq.transpose(1, 2)).float()


def pca_unproject(pca_components, pca_mean, latent_scale, z):
    return (z / latent_scale) @ pca_components + pca_mean


# ------------------------------------------------------------------------------------------------ Fish S1-DAC encode
def causal_conv1d_strided(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int) -> torch.Tensor:
    """CausalConvNet.forward with stride (autoencoder.py:269-289): left pad (k - stride) zeros, plus the right pad of
    get_extra_padding_for_conv1d (:48-55) so that the last window is complete."""
    import math
    k = w.shape[-1]
    pad = k - stride
    length = x.shape[-1]
    n_frames = (length - k + pad) / stride + 1
    extra = (math.ceil(n_frames) - 1) * stride + (k - pad) - length
    return torch.nn.functional.conv1d(torch.nn.functional.pad(x, (pad, extra)), w, b, stride=stride)


def _convnext(sd: SD, p: str, z: torch.Tensor) -> torch.Tensor:
    """ConvNeXtBlock.forward (autoencoder.py:360-373): causal depthwise conv7, LayerNorm(eps 1e-6), pw1, GELU, pw2,
    gamma, residual."""
    C = z.shape[1]
    y = causal_conv1d(z, sd[f"{p}.dwconv.conv.weight"], sd[f"{p}.dwconv.conv.bias"], groups=C).transpose(1, 2)
    y = torch.nn.functional.layer_norm(y, (C,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-6)
    y = linear(torch.nn.functional.gelu(linear(y, sd[f"{p}.pwconv1.weight"], sd[f"{p}.pwconv1.bias"])),
               sd[f"{p}.pwconv2.weight"], sd[f"{p}.pwconv2.bias"])
    return z + (y * sd[f"{p}.gamma"]).transpose(1, 2)


def dac_encoder(sd: SD, cfg, audio: torch.Tensor) -> torch.Tensor:
    """Encoder.forward (autoencoder.py:903-929), causal: WN-conv7 1 -> enc_dim; per rate an EncoderBlock (:839-876) =
    three ResidualUnits (dilations 1, 3, 9) at the input width, Snake, WN-conv(k = 2 stride, stride) doubling the
    width, and -- last block only -- a window-512 transformer; then Snake and WN-conv3. (B, 1, L) -> (B, C, L/hop)."""
    x = causal_conv1d(audio, _wn(sd, "encoder.block.0.conv"), sd["encoder.block.0.conv.bias"])
    nb = len(cfg.enc_rates)
    for bi, stride in enumerate(cfg.enc_rates):
        p = f"encoder.block.{bi + 1}.block"
        for ui, dil in enumerate((1, 3, 9)):
            q = f"{p}.{ui}.block"
            y = snake(x, sd[f"{q}.0.alpha"])
            y = causal_conv1d(y, _wn(sd, f"{q}.1.conv"), sd[f"{q}.1.conv.bias"], dilation=dil)
            y = snake(y, sd[f"{q}.2.alpha"])
            y = causal_conv1d(y, _wn(sd, f"{q}.3.conv"), sd[f"{q}.3.conv.bias"])
            x = x + y
        x = snake(x, sd[f"{p}.3.alpha"])
        x = causal_conv1d_strided(x, _wn(sd, f"{p}.4.conv"), sd[f"{p}.4.conv.bias"], stride)
        if bi == nb - 1 and cfg.enc_t_layers > 0:
            x = dac_window_transformer(sd, f"{p}.5", x, cfg.enc_t_layers, x.shape[1] // 64, cfg.enc_window, 1e-5)
    x = snake(x, sd[f"encoder.block.{nb + 1}.alpha"])
    return causal_conv1d(x, _wn(sd, f"encoder.block.{nb + 2}.conv"), sd[f"encoder.block.{nb + 2}.conv.bias"])


def dac_quantizer_front(sd: SD, cfg, z: torch.Tensor) -> torch.Tensor:
    """DownsampleResidualVectorQuantize.forward up to the quantizers (autoencoder.py:452-453): per factor a causal
    conv(k = 2, stride 2) + ConvNeXtBlock (:418-424), then pre_module (window-128 transformer)."""
    for i in range(cfg.num_upsample):
        p = f"quantizer.downsample.{i}"
        z = causal_conv1d_strided(z, sd[f"{p}.0.conv.weight"], sd[f"{p}.0.conv.bias"], 2)
        z = _convnext(sd, f"{p}.1", z)
    return dac_window_transformer(sd, "quantizer.pre_module", z, cfg.post_layers, cfg.post_heads, cfg.post_window,
                                  cfg.post_norm_eps)


def vq_nearest(sd: SD, p: str, residual: torch.Tensor):
    """VectorQuantize.forward in eval mode (autoencoder.py:132-157): z_e = in_proj(residual); nearest code by
    L2 distance between the L2-normalised z_e and the L2-normalised codebook; z_q = out_proj(codebook[idx])."""
    z_e = torch.nn.functional.conv1d(residual, _wn(sd, f"{p}.in_proj"), sd[f"{p}.in_proj.bias"])
    B, D, T = z_e.shape
    enc = torch.nn.functional.normalize(z_e.transpose(1, 2).reshape(B * T, D))
    cb = torch.nn.functional.normalize(sd[f"{p}.codebook.weight"])
    dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cb.t() + cb.pow(2).sum(1, keepdim=True).t()
    idx = (-dist).max(1)[1].reshape(B, T)
    z_q = torch.nn.functional.conv1d(sd[f"{p}.codebook.weight"][idx].transpose(1, 2), _wn(sd, f"{p}.out_proj"),
                                     sd[f"{p}.out_proj.bias"])
    return z_q, idx


def dac_encode_codes(sd: SD, cfg, z: torch.Tensor) -> torch.Tensor:
    """The code path of DownsampleResidualVectorQuantize.forward (autoencoder.py:455-463): one semantic codebook on
    z, then n_codebooks residual codebooks on z - z_semantic. Returns codes (B, 1 + n_codebooks, T)."""
    z_sem, idx = vq_nearest(sd, "quantizer.semantic_quantizer.quantizers.0", z)
    codes = [idx]
    residual = z - z_sem
    for i in range(cfg.n_codebooks):
        z_q_i, idx = vq_nearest(sd, f"quantizer.quantizer.quantizers.{i}", residual)
        residual = residual - z_q_i
        codes.append(idx)
    return torch.stack(codes, dim=1)


def dac_zq_from_codes(sd: SD, cfg, codes: torch.Tensor) -> torch.Tensor:
    """DAC.encode_zq after encode() (autoencoder.py:1116-1126): clamp, then from_codes of both quantizers (:215-225),
    z_q = z_q_semantic + sum_i out_proj_i(codebook_i[code_i])."""
    def from_codes(name, c, size):
        z_q = 0.0
        for i in range(c.shape[1]):
            p = f"quantizer.{name}.quantizers.{i}"
            e = sd[f"{p}.codebook.weight"][c[:, i].clamp(max=size - 1)].transpose(1, 2)
            z_q = z_q + torch.nn.functional.conv1d(e, _wn(sd, f"{p}.out_proj"), sd[f"{p}.out_proj.bias"])
        return z_q
    return from_codes("semantic_quantizer", codes[:, :1], cfg.semantic_codebook_size) + \
        from_codes("quantizer", codes[:, 1:], cfg.codebook_size)


def dac_encode_zq(sd: SD, cfg, audio: torch.Tensor, return_parts: bool = False):
    """DAC.encode_zq (autoencoder.py:1080-1126): right-pad to a multiple of frame_length, encoder, quantizer front,
    codes, z_q. (B, 1, L) -> (B, C, ceil(L / frame_length)). The post_module / upsample that DAC.encode also runs
    inside quantizer.forward (:464-465) do not influence the codes and are skipped."""
    import math
    if audio.dim() == 2:
        audio = audio.unsqueeze(1)
    L = audio.shape[-1]
    audio = torch.nn.functional.pad(audio, (0, math.ceil(L / cfg.frame_length) * cfg.frame_length - L))
    z_enc = dac_encoder(sd, cfg, audio)
    z = dac_quantizer_front(sd, cfg, z_enc)
    codes = dac_encode_codes(sd, cfg, z)
    zq = dac_zq_from_codes(sd, cfg, codes)
    return (zq, dict(z_enc=z_enc, z_pre=z, codes=codes)) if return_parts else zq


def ae_encode(sd: SD, cfg, pca_components: torch.Tensor, pca_mean: torch.Tensor, latent_scale: float,
              audio: torch.Tensor) -> torch.Tensor:
    """inference.ae_encode (inference.py:219-224): z_q -> PCA projection -> * latent_scale. (B, 1, L) -> (B, T, 80)."""
    zq = dac_encode_zq(sd, cfg, audio).float()
    return ((zq.transpose(1, 2) - pca_mean) @ pca_components.T) * latent_scale


def get_speaker_latent_and_mask(sd: SD, cfg, pca, audio: torch.Tensor, max_speaker_latent_length: int = 6400,
                                audio_chunk_size: int = 640 * 2048, divis_by_patch_size: int = 4):
    """inference.get_speaker_latent_and_mask (inference.py:240-283), pad_to_max=False: encode in 640-latent chunks
    (last one zero padded), keep the latents of complete frames, trim to a multiple of the patch size."""
    hop = cfg.frame_length
    audio = audio[:, : max_speaker_latent_length * hop]
    lat = []
    for i in range(0, audio.shape[1], audio_chunk_size):
        chunk = audio[:, i:i + audio_chunk_size]
        if chunk.shape[1] < audio_chunk_size:
            chunk = torch.nn.functional.pad(chunk, (0, audio_chunk_size - chunk.shape[1]))
        lat.append(ae_encode(sd, cfg, pca[0], pca[1], pca[2], chunk.unsqueeze(0)))
    lat = torch.cat(lat, dim=1)
    n = audio.shape[1] // hop
    mask = (torch.arange(lat.shape[1]) < n).unsqueeze(0)
    lat, mask = lat[:, :n], mask[:, :n]
    n4 = lat.shape[1] // divis_by_patch_size * divis_by_patch_size
    return lat[:, :n4], mask[:, :n4]
