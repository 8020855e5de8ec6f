"""Samplers with the reference's call signatures, executed by libecho_b200.so.

`sample_euler_cfg_independent_guidances` mirrors reference inference.py:427-517 and
`sample_blockwise_euler_cfg_independent_guidances` mirrors inference_blockwise.py:15-123: same positional and keyword
arguments, same fp32 (B, S, 80) result, so `functools.partial(...)` objects built by handler._build_sample_fn
(handler.py:426-443) work unchanged as the `sample_fn` of inference.sample_pipeline.

The initial noise is drawn exactly as the reference draws it (`torch.Generator(device).manual_seed(seed)` +
`torch.randn`, inference.py:457,477), so seeds stay compatible on the same device; the optional keyword-only `noise`
lets tests inject a tensor (the CPU and CUDA generators differ).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from .model import B200EchoDiT, _stream, _u8


_SCHEDULES: dict = {}


def _args(model: B200EchoDiT, num_steps, cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t, truncation_factor,
          rescale_k, rescale_sigma, speaker_kv_scale, speaker_kv_max_layers, speaker_kv_min_t, sequence_length):
    a = _lib.SamplerArgs()
    a.num_steps = int(num_steps)
    a.cfg_scale_text, a.cfg_scale_speaker = float(cfg_scale_text), float(cfg_scale_speaker)
    a.cfg_min_t, a.cfg_max_t = float(cfg_min_t), float(cfg_max_t)
    a.has_truncation = int(truncation_factor is not None)
    a.truncation_factor = float(truncation_factor or 0.0)
    a.has_rescale = int(rescale_k is not None and rescale_sigma is not None)
    a.rescale_k, a.rescale_sigma = float(rescale_k or 0.0), float(rescale_sigma or 0.0)
    a.has_kv_scale = int(speaker_kv_scale is not None)
    if speaker_kv_scale is not None and speaker_kv_min_t is None:
        # the reference evaluates `t_next < None` and raises (inference.py:511)
        raise TypeError("'<' not supported between instances of 'Tensor' and 'NoneType' (speaker_kv_min_t is None)")
    a.speaker_kv_scale = float(speaker_kv_scale or 0.0)
    # None -> every layer (C side: negative); an explicit n scales min(n, num_layers) layers, n <= 0 none (inference.py:408-414)
    a.speaker_kv_max_layers = max(int(speaker_kv_max_layers), 0) if speaker_kv_max_layers is not None else -1
    a.speaker_kv_min_t = float(speaker_kv_min_t or 0.0)
    a.sequence_length = int(sequence_length)
    a.round_t_to_bf16 = int(model.round_t_to_model_dtype)
    # the schedule is computed by torch on the model's device, exactly as the reference does (inference.py:459)
    key = (a.num_steps, str(model.device))
    sched = _SCHEDULES.get(key)
    if sched is None:  # one device round trip per (num_steps, device), not per request
        INIT_SCALE = 0.999
        sched = (torch.linspace(1., 0., a.num_steps + 1, device=model.device) * INIT_SCALE).cpu().contiguous()
        _SCHEDULES[key] = sched
    a._sched_keepalive = sched
    a.t_schedule = C.cast(sched.data_ptr(), C.POINTER(C.c_float))
    return a


def _attach_speaker_kv(a, model: B200EchoDiT, speaker_kv_cache, B: int, Ls: int) -> None:
    """Point the sampler at a speaker KV cache kept from an earlier `model.get_kv_cache_speaker(speaker_latent)` (per-voice
    persistence): list of num_layers (K, V), each (B, Ls/4, heads, 128) contiguous model-dtype on the
    model's device -- the layout the reference's `get_kv_cache_speaker` returns (model.py:615-621)."""
    L, P = model.cfg.num_layers, Ls // model.cfg.speaker_patch_size
    if len(speaker_kv_cache) != L:
        raise ValueError(f"speaker_kv_cache has {len(speaker_kv_cache)} layers, the model {L}")
    Ks, Vs = (C.c_void_p * L)(), (C.c_void_p * L)()
    for i, (k, v) in enumerate(speaker_kv_cache):
        for t in (k, v):
            if (t.device != model.device or t.dtype != torch.bfloat16 or not t.is_contiguous()
                    or t.shape[0] != B or t.shape[1] != P):
                raise ValueError(f"speaker_kv_cache[{i}]: expected contiguous bf16 ({B}, {P}, heads, head_dim) on "
                                 f"{model.device}, got {tuple(t.shape)} {t.dtype} on {t.device}")
        Ks[i], Vs[i] = k.data_ptr(), v.data_ptr()
    a._kv_keepalive = (Ks, Vs, speaker_kv_cache)
    a.speaker_K, a.speaker_V = C.cast(Ks, C.POINTER(C.c_void_p)), C.cast(Vs, C.POINTER(C.c_void_p))


def _text_valid_len(text_mask: torch.Tensor) -> int:
    """Length of the longest unmasked text prefix over the batch, when it is known WITHOUT a device round trip: the mask is
    still in host memory, or `pipeline.get_text_input_ids_and_mask` built it and left the number on the tensor. 0 = not
    known (the text encoder then runs over all padded rows, as the reference does). The samplers skip the text rows behind
    it: they are masked out of every attention (echo_sampler_args::text_valid_len)."""
    n = getattr(text_mask, "_echo_valid_len", None)
    if n is not None:
        return max(int(n), 1)
    if text_mask.device.type == "cpu" and text_mask.numel() > 0:
        cols = text_mask.reshape(-1, text_mask.shape[-1]).any(0).nonzero()
        return int(cols.max()) + 1 if cols.numel() else 1
    return 0


def _inputs(model, speaker_latent, speaker_mask, text_input_ids, text_mask):
    dev = model.device
    spk = speaker_latent.to(dev, torch.bfloat16).contiguous()  # reference: speaker_latent.to(dtype) (inference.py:465)
    sm = _u8(speaker_mask.to(dev))
    ids = text_input_ids.to(dev, torch.int32).contiguous()
    tm = _u8(text_mask.to(dev))
    return spk, sm, ids, tm


@torch.inference_mode()
def sample_euler_cfg_independent_guidances(
    model: B200EchoDiT,
    speaker_latent: torch.Tensor,
    speaker_mask: torch.Tensor,
    text_input_ids: torch.Tensor,
    text_mask: torch.Tensor,
    rng_seed: int,
    num_steps: int,
    cfg_scale_text: float,
    cfg_scale_speaker: float,
    cfg_min_t: float,
    cfg_max_t: float,
    truncation_factor: Optional[float],
    rescale_k: Optional[float],
    rescale_sigma: Optional[float],
    speaker_kv_scale: Optional[float],
    speaker_kv_max_layers: Optional[int],
    speaker_kv_min_t: Optional[float],
    sequence_length: Optional[int] = None,
    *,
    noise: Optional[torch.Tensor] = None,
    speaker_kv_cache=None,
) -> torch.Tensor:
    """`speaker_kv_cache` (optional, keyword-only extension): the result of an earlier
    `model.get_kv_cache_speaker(speaker_latent)` for the same voice; the speaker encoder is then skipped."""
    if sequence_length is None:
        sequence_length = 640  # max sequence length during training (inference.py:449-450)
    dev = model.device
    B = text_input_ids.shape[0]
    C_lat = model.cfg.latent_size
    if noise is None:
        rng = torch.Generator(device=dev).manual_seed(rng_seed)
        noise = torch.randn((B, sequence_length, C_lat), device=dev, dtype=torch.float32, generator=rng)
    noise = noise.to(dev, torch.float32).contiguous()
    assert noise.shape == (B, sequence_length, C_lat)
    a = _args(model, num_steps, cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t, truncation_factor, rescale_k,
              rescale_sigma, speaker_kv_scale, speaker_kv_max_layers, speaker_kv_min_t, sequence_length)
    a.text_valid_len = _text_valid_len(text_mask)
    spk, sm, ids, tm = _inputs(model, speaker_latent, speaker_mask, text_input_ids, text_mask)
    if speaker_kv_cache is not None:
        _attach_speaker_kv(a, model, speaker_kv_cache, B, sm.shape[1])
    out = torch.empty_like(noise)
    with torch.cuda.device(dev):
        _lib.check(model.lib.echo_sample_euler(model.h.ptr, C.byref(a), spk.data_ptr(), sm.data_ptr(), sm.shape[1],
                                               ids.data_ptr(), tm.data_ptr(), tm.shape[1], B, noise.data_ptr(),
                                               out.data_ptr(), _stream(dev)), "echo_sample_euler")
    return out


@torch.inference_mode()
def sample_blockwise_euler_cfg_independent_guidances(
    model: B200EchoDiT,
    speaker_latent: torch.Tensor,
    speaker_mask: torch.Tensor,
    text_input_ids: torch.Tensor,
    text_mask: torch.Tensor,
    rng_seed: int,
    block_sizes: List[int],
    num_steps: int,
    cfg_scale_text: float,
    cfg_scale_speaker: float,
    cfg_min_t: float,
    cfg_max_t: float,
    truncation_factor: Optional[float],
    rescale_k: Optional[float],
    rescale_sigma: Optional[float],
    speaker_kv_scale: Optional[float],
    speaker_kv_max_layers: Optional[int],
    speaker_kv_min_t: Optional[float],
    continuation_latent: Optional[torch.Tensor] = None,
    *,
    noise_blocks: Optional[List[torch.Tensor]] = None,
    on_block=None,
    speaker_kv_cache=None,
) -> torch.Tensor:
    """`on_block(index, start, length, prefix)` (optional, keyword-only extension): called on the host as soon as the
    kernels of a block have been enqueued; `prefix[:, :start + length]` is final in stream order, so the callback can
    enqueue the DAC decode of that prefix while the next block samples (streaming)."""
    dev = model.device
    B = text_input_ids.shape[0]
    C_lat = model.cfg.latent_size
    if noise_blocks is None:  # one generator, one draw per block, in block order (inference_blockwise.py:42,76)
        rng = torch.Generator(device=dev).manual_seed(rng_seed)
        noise_blocks = [torch.randn((B, bs, C_lat), device=dev, dtype=torch.float32, generator=rng) for bs in block_sizes]
    noise = torch.cat([n.to(dev, torch.float32).reshape(-1) for n in noise_blocks]).contiguous()
    Lc = 0
    cont = None
    if continuation_latent is not None:
        cont = continuation_latent.to(dev, torch.float32).contiguous()
        Lc = cont.shape[1]
    total = Lc + sum(block_sizes)
    a = _args(model, num_steps, cfg_scale_text, cfg_scale_speaker, cfg_min_t, cfg_max_t, truncation_factor, rescale_k,
              rescale_sigma, speaker_kv_scale, speaker_kv_max_layers, speaker_kv_min_t, max(block_sizes))
    a.text_valid_len = _text_valid_len(text_mask)
    spk, sm, ids, tm = _inputs(model, speaker_latent, speaker_mask, text_input_ids, text_mask)
    if speaker_kv_cache is not None:  # same extension as in sample_euler_cfg_independent_guidances
        _attach_speaker_kv(a, model, speaker_kv_cache, B, sm.shape[1])
    out = torch.empty(B, total, C_lat, device=dev, dtype=torch.float32)
    blocks = (C.c_int * len(block_sizes))(*[int(b) for b in block_sizes])
    errors = []

    def _cb(_user, index, start, length):
        try:
            on_block(index, start, length, out)
        except BaseException as e:  # never unwind through the C frame
            errors.append(e)

    cb = _lib.BLOCK_CB(_cb) if on_block is not None else C.cast(None, _lib.BLOCK_CB)
    with torch.cuda.device(dev):
        _lib.check(model.lib.echo_sample_blockwise_stream(
            model.h.ptr, C.byref(a), blocks, len(block_sizes), spk.data_ptr(), sm.data_ptr(), sm.shape[1],
            ids.data_ptr(), tm.data_ptr(), tm.shape[1], B, None if cont is None else cont.data_ptr(), Lc,
            noise.data_ptr(), out.data_ptr(), cb, None, _stream(dev)), "echo_sample_blockwise")
    if errors:
        raise errors[0]
    return out
