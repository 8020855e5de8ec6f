"""Host side of one synthesis job: text -> chunks -> (per chunk) tokens -> sampler -> DAC decode -> crop -> stitch.

Mirrors the callers either side of the CUDA hot path, with the reference's names, argument meaning and results:

  tokenizer_encode / get_text_input_ids_and_mask    reference inference.py:115-138, 192-215
  chunk_text / chunk_text_for_audio                 reference inference.py:140-190, handler.py:102-123
  find_flattening_point / crop_audio_to_...         reference inference.py:288-301   (vectorised: one pass, no
                                                    per-window device sync)
  get_speaker_latent_and_mask                       reference inference.py:240-283   (chunks batched into one call)
  sample_pipeline                                   reference inference.py:309-347
  crossfade_chunks / normalize_chunk_boundaries     reference handler.py:126-240     (the per-sample Python loop of
                                                    handler.py:214-218 becomes one reduction)
  synthesize                                        reference handler.py:736-768     (chunk loop + stitching)

Multi-GPU: the chunks of a long prompt -- or whole requests -- are independent units. `shard_units`
assigns unit i to rank i % world, every rank runs its units on its own GPU replica, and the finished audio (<= 5 MB
per chunk) is gathered once per job; the stitching runs on the host of rank 0. There is no collective on the
per-step path.
"""
from __future__ import annotations

import re
from typing import Callable, List, Optional, Sequence, Tuple

import torch

MAX_TEXT_LENGTH = 768             # inference.py:324
MAX_SPEAKER_LATENT_LENGTH = 6400  # inference.py:323
AE_DOWNSAMPLE_FACTOR = 2048
SAMPLE_RATE = 44100

_WS = re.compile(r"\s+")
_REPLACEMENTS = (("…", "..."), ("’", "'"), ("”", '"'), ("\n", " "), (":", ","), (";", ","),
                 ("—", ", "))


# ------------------------------------------------------------------------------------------------ text
def tokenizer_encode(text: str, append_bos: bool = True, normalize: bool = True,
                     return_normalized_text: bool = False):
    if normalize:
        for a, b in _REPLACEMENTS:
            text = text.replace(a, b)
        if not text.startswith(("[", "(")) and "S1" not in text and "S2" not in text:
            text = "[S1] " + text
    ids = ([0] if append_bos else []) + list(text.encode("utf-8"))
    t = torch.tensor(ids)
    return (t, text) if return_normalized_text else t


def get_text_input_ids_and_mask(text_arr: List[str], max_length: Optional[int], device=None, normalize: bool = True,
                                return_normalized_text: bool = False, pad_to_max: bool = True):
    enc = [tokenizer_encode(t, normalize=normalize, return_normalized_text=True) for t in text_arr]
    if max_length is None:
        max_length = max(len(e) for e, _ in enc)
    tokens = torch.zeros((len(text_arr), max_length), dtype=torch.int32)
    mask = torch.zeros((len(text_arr), max_length), dtype=torch.bool)
    for i, (e, _) in enumerate(enc):
        n = min(len(e), max_length)
        tokens[i, :n] = e[:n]
        mask[i, :n] = True
    valid = max((min(len(e), max_length) for e, _ in enc), default=0)
    if device is not None:
        tokens, mask = tokens.to(device), mask.to(device)
    # the samplers of this package skip the text rows behind the longest unmasked prefix when they know it without a
    # device round trip (sampler._text_valid_len); a copy / slice of the tensor loses the attribute, which is the safe side
    mask._echo_valid_len = valid
    if return_normalized_text:
        return tokens, mask, [t for _, t in enc]
    return tokens, mask


_SENTENCE, _CLAUSE = ".!?", ",;:"
_CLOSERS = "\"')]}”’"


def _split_point(window: str) -> int:
    """Last whitespace of the window that follows a sentence end, else a clause end, else any; 0 if none."""
    best = [0, 0, 0]  # sentence, clause, space
    for m in _WS_CHAR.finditer(window, 1):
        i = m.start()
        best[2] = i
        p1 = window[i - 1]
        p0 = window[i - 2] if i >= 2 else ""
        closing = p1 in _CLOSERS
        if p1 in _SENTENCE or (closing and p0 != "" and p0 in _SENTENCE):
            best[0] = i
        elif p1 in _CLAUSE or (closing and p0 != "" and p0 in _CLAUSE):
            best[1] = i
    return best[0] or best[1] or best[2]


_WS_CHAR = re.compile(r"\s")


def chunk_text(text: str, max_chars: int = 300) -> List[str]:
    """Split into <= max_chars chunks, preferring sentence, then clause, then word boundaries."""
    if max_chars <= 0:
        raise ValueError("max_chars must be > 0")
    rest = _WS.sub(" ", text or "").strip()
    out: List[str] = []
    while len(rest) > max_chars:
        cut = _split_point(rest[: max_chars + 1]) or max_chars
        head = rest[:cut].strip()
        if head:
            out.append(head)
        rest = rest[cut:].strip()
    if rest:
        out.append(rest)
    return out


def chunk_text_for_audio(text: str, max_chars: int = 300, target_duration_seconds: float = 10.0) -> List[str]:
    chunks = chunk_text(text, max_chars=min(max_chars, int(target_duration_seconds * 12)))  # ~12 chars / s of speech
    if len(chunks) > 1 and len(chunks[-1]) < 24:  # a < 2 s tail joins the previous chunk
        tail = chunks.pop()
        chunks[-1] += " " + tail
    return chunks


# ------------------------------------------------------------------------------------------------ audio -> latents
@torch.inference_mode()
def get_speaker_latent_and_mask(fish_ae, pca_state, audio: torch.Tensor,
                                max_speaker_latent_length: int = MAX_SPEAKER_LATENT_LENGTH,
                                audio_chunk_size: Optional[int] = None, pad_to_max: bool = False,
                                divis_by_patch_size: Optional[int] = 4) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference inference.get_speaker_latent_and_mask (inference.py:240-283): the speaker reference (1, L) is encoded
    in chunks of 640 latents (the last one zero padded), latents of incomplete frames are dropped, and the length is
    cut to a multiple of the speaker patch size. All chunks go through ONE batched encoder call here (the reference
    loops); results are identical because chunks are independent."""
    from .autoencoder import ae_encode
    hop = getattr(getattr(fish_ae, "cfg", None), "frame_length", AE_DOWNSAMPLE_FACTOR)
    if audio_chunk_size is None:
        audio_chunk_size = 640 * hop
    assert audio.ndim == 2 and audio.shape[0] == 1  # (1, length)
    audio = audio[:, : max_speaker_latent_length * hop]
    n_chunks = max(1, -(-audio.shape[1] // audio_chunk_size))
    padded = torch.nn.functional.pad(audio, (0, n_chunks * audio_chunk_size - audio.shape[1]))
    chunks = padded.reshape(n_chunks, 1, audio_chunk_size)
    latent = ae_encode(fish_ae, pca_state, chunks)                      # (n_chunks, frames, 80)
    speaker_latent = latent.reshape(1, -1, latent.shape[-1])
    actual = audio.shape[1] // hop
    speaker_mask = (torch.arange(speaker_latent.shape[1], device=speaker_latent.device) < actual).unsqueeze(0)
    if pad_to_max and speaker_latent.shape[1] < max_speaker_latent_length:
        extra = max_speaker_latent_length - speaker_latent.shape[1]
        speaker_latent = torch.nn.functional.pad(speaker_latent, (0, 0, 0, extra))
        speaker_mask = torch.nn.functional.pad(speaker_mask, (0, extra))
    elif not pad_to_max:
        speaker_latent, speaker_mask = speaker_latent[:, :actual], speaker_mask[:, :actual]
    if divis_by_patch_size is not None:
        n = speaker_latent.shape[1] // divis_by_patch_size * divis_by_patch_size
        speaker_latent, speaker_mask = speaker_latent[:, :n], speaker_mask[:, :n]
    return speaker_latent, speaker_mask


# ------------------------------------------------------------------------------------------------ per-voice cache
class Voice:
    """What a speaker reference costs once: PCA latents + mask (DAC encoder + RVQ) and the 24-layer speaker KV cache."""
    __slots__ = ("speaker_latent", "speaker_mask", "kv", "nbytes")

    def __init__(self, speaker_latent, speaker_mask, kv):
        self.speaker_latent, self.speaker_mask, self.kv = speaker_latent, speaker_mask, kv
        self.nbytes = speaker_latent.numel() * speaker_latent.element_size() + speaker_mask.numel() + sum(
            k.numel() * k.element_size() + v.numel() * v.element_size() for k, v in kv)


class VoiceCache:
    """Per-voice persistence. The reference re-encodes the speaker audio and re-runs the speaker encoder
    + 24 K/V projections on every request (inference.py:332-339, 465); a server that sees the same voices again keeps
    both here: `get(voice_id, audio)` encodes on the first request and afterwards returns the stored latents, mask and
    KV cache, which `sample_pipeline(..., voice=...)` / the samplers' `speaker_kv_cache=` consume without touching them
    (the samplers copy the cache when they have to scale it). Least-recently-used voices are dropped once the stored
    bytes exceed `max_bytes` (a 10 s reference is 10 MB, a 5-minute one 315 MB; a B200 has 180 GB).

    `encode(audio) -> (speaker_latent, speaker_mask)` and `build_kv(speaker_latent) -> [(K, V)] * layers` default to
    `get_speaker_latent_and_mask` and `model.get_kv_cache_speaker`."""

    def __init__(self, model, fish_ae=None, pca_state=None, max_bytes: int = 8 << 30, encode: Optional[Callable] = None,
                 build_kv: Optional[Callable] = None):
        from collections import OrderedDict
        self._voices: "OrderedDict[str, Voice]" = OrderedDict()
        self.max_bytes, self.nbytes = int(max_bytes), 0
        self.hits = self.misses = 0
        self._encode = encode or (lambda audio: get_speaker_latent_and_mask(fish_ae, pca_state, audio.to(model.device)))
        self._build_kv = build_kv or (lambda latent: model.get_kv_cache_speaker(latent.to(model.dtype)))

    def __len__(self) -> int:
        return len(self._voices)

    def __contains__(self, voice_id: str) -> bool:
        return voice_id in self._voices

    def get(self, voice_id: str, audio: Optional[torch.Tensor] = None, speaker_latent: Optional[torch.Tensor] = None,
            speaker_mask: Optional[torch.Tensor] = None) -> Voice:
        """The stored voice, or encode `audio` ((1, L) @ 44.1 kHz) / take already-encoded latents + mask, build its KV
        cache and store it. KeyError if the voice is unknown and nothing to build it from is given."""
        v = self._voices.get(voice_id)
        if v is not None:
            self._voices.move_to_end(voice_id)
            self.hits += 1
            return v
        if speaker_latent is None:
            if audio is None:
                raise KeyError(f"voice {voice_id!r} is not cached and no audio / latents were given")
            speaker_latent, speaker_mask = self._encode(audio)
        self.misses += 1
        v = Voice(speaker_latent, speaker_mask, self._build_kv(speaker_latent))
        self._voices[voice_id] = v
        self.nbytes += v.nbytes
        while self.nbytes > self.max_bytes and len(self._voices) > 1:
            _, old = self._voices.popitem(last=False)
            self.nbytes -= old.nbytes
        return v

    def drop(self, voice_id: str) -> None:
        v = self._voices.pop(voice_id, None)
        if v is not None:
            self.nbytes -= v.nbytes


# ------------------------------------------------------------------------------------------------ latents -> audio
def find_flattening_point(data: torch.Tensor, target_value: float = 0.0, window_size: int = 20,
                          std_threshold: float = 0.05) -> int:
    """First index whose `window_size`-latent window (zero padded past the end) is flat: std < std_threshold and
    |mean - target| < 0.1 (reference inference.py:288-296; the reference loops over the windows with two device syncs
    each). Latents on the GPU -- the sampler's output, the product path -- go through one warp-per-window kernel
    (`echo_op_flattening_point`) and the index leaves the device once. Host tensors (the stitching side of the
    pipeline works on the host) are evaluated in one vectorised pass."""
    n = data.shape[0]
    if n == 0:
        return 0
    if data.is_cuda:
        import ctypes as C

        from . import _lib
        x = data.reshape(n, -1).to(torch.float32).contiguous()
        out = torch.empty(1, dtype=torch.int32, device=data.device)
        with torch.cuda.device(data.device):
            _lib.check(_lib.load().echo_op_flattening_point(
                x.data_ptr(), n, x.shape[1], C.c_float(target_value), int(window_size), C.c_float(std_threshold),
                out.data_ptr(), torch.cuda.current_stream(data.device).cuda_stream), "echo_op_flattening_point")
        return int(out.item())
    padded = torch.cat([data, data.new_zeros((window_size,) + tuple(data.shape[1:]))])
    win = padded.reshape(padded.shape[0], -1).unfold(0, window_size, 1)[:n]  # (n, features, window)
    win = win.reshape(n, -1).float()
    flat = (win.std(dim=1) < std_threshold) & ((win.mean(dim=1) - target_value).abs() < 0.1)
    idx = torch.nonzero(flat)
    return int(idx[0]) if idx.numel() else n


def crop_audio_to_flattening_point(audio: torch.Tensor, latent: torch.Tensor) -> torch.Tensor:
    return audio[..., : find_flattening_point(latent) * AE_DOWNSAMPLE_FACTOR]


@torch.inference_mode()
def sample_pipeline(model, fish_ae, pca_state, sample_fn: Callable, text_prompt: str,
                    speaker_audio: Optional[torch.Tensor] = None, rng_seed: int = 0,
                    pad_to_max_speaker_latent_length: Optional[int] = None,
                    pad_to_max_text_length: Optional[int] = None, normalize_text: bool = True, *,
                    speaker_latent: Optional[torch.Tensor] = None, speaker_mask: Optional[torch.Tensor] = None,
                    voice: Optional[Voice] = None) -> Tuple[torch.Tensor, str]:
    """One chunk: reference inference.py:309-347, same positional order (model, fish_ae, pca_state, sample_fn,
    text_prompt, speaker_audio, rng_seed, pad_..., normalize_text). The speaker reference is raw audio (1, L) @ 44.1 kHz
    (`speaker_audio`, encoded here like the reference does) or, through the keyword-only extensions, already-encoded
    PCA latents + mask, or a `Voice` from a `VoiceCache` (latents, mask and the speaker KV cache: `sample_fn` must
    then accept `speaker_kv_cache=`, as the samplers of this package do). Returns (audio (1, 1, n) fp32, normalised
    text)."""
    from .autoencoder import ae_decode
    device = model.device
    ids, mask, norm = get_text_input_ids_and_mask(
        [text_prompt], max_length=min(pad_to_max_text_length or MAX_TEXT_LENGTH, MAX_TEXT_LENGTH), device=device,
        normalize=normalize_text, return_normalized_text=True, pad_to_max=(pad_to_max_text_length is not None))
    if voice is None and speaker_audio is not None and speaker_latent is None:  # inference.py:332-339
        speaker_latent, speaker_mask = get_speaker_latent_and_mask(
            fish_ae, pca_state, speaker_audio.to(device),
            max_speaker_latent_length=pad_to_max_speaker_latent_length or MAX_SPEAKER_LATENT_LENGTH,
            pad_to_max=(pad_to_max_speaker_latent_length is not None))
    if voice is None and speaker_latent is None:  # inference.py:329-331
        n = pad_to_max_speaker_latent_length or 4
        speaker_latent = torch.zeros((1, n, 80), device=device, dtype=model.dtype)
        speaker_mask = torch.zeros((1, n), device=device, dtype=torch.bool)
    if voice is not None:
        latent = sample_fn(model, voice.speaker_latent, voice.speaker_mask, ids, mask, rng_seed, speaker_kv_cache=voice.kv)
    else:
        latent = sample_fn(model, speaker_latent, speaker_mask, ids, mask, rng_seed)
    audio = ae_decode(fish_ae, pca_state, latent)
    return crop_audio_to_flattening_point(audio, latent[0]), norm[0]


@torch.inference_mode()
def sample_pipeline_batch(model, fish_ae, pca_state, sample_fn: Callable, text_prompts: Sequence[str],
                          rng_seeds: Sequence[int], speaker_audio: Optional[torch.Tensor] = None,
                          pad_to_max_speaker_latent_length: Optional[int] = None,
                          pad_to_max_text_length: Optional[int] = None, normalize_text: bool = True, *,
                          speaker_latent: Optional[torch.Tensor] = None, speaker_mask: Optional[torch.Tensor] = None,
                          voice: Optional[Voice] = None) -> List[Tuple[torch.Tensor, str]]:
    """`sample_pipeline` for several prompts that share ONE speaker reference -- the chunks of a long prompt
    (handler.py:746-759 calls sample_pipeline once per chunk) -- in ONE sampler call and ONE DAC decode: the GEMMs then
    run at M = B x 640 rows, where the tensor cores are 15-20 % better used than at batch 1. Item i draws its noise
    exactly as a single call with `rng_seeds[i]` would (its own generator), so a chunk's latents differ from the
    one-at-a-time result only by the bf16 noise floor (other tile shapes / summation orders), never by its inputs.
    `sample_fn` must accept the keyword `noise=` (the samplers of this package do). Returns [(audio (1, 1, n), text)]."""
    from .autoencoder import ae_decode
    device = model.device
    B = len(text_prompts)
    assert B == len(rng_seeds) and B > 0
    ids, mask, norm = get_text_input_ids_and_mask(
        list(text_prompts), max_length=min(pad_to_max_text_length or MAX_TEXT_LENGTH, MAX_TEXT_LENGTH), device=device,
        normalize=normalize_text, return_normalized_text=True, pad_to_max=(pad_to_max_text_length is not None))
    kw = {}
    if voice is not None:
        speaker_latent, speaker_mask = voice.speaker_latent, voice.speaker_mask
        kw["speaker_kv_cache"] = [(k.expand(B, *k.shape[1:]).contiguous(), v.expand(B, *v.shape[1:]).contiguous())
                                  for k, v in voice.kv]
    elif speaker_audio is not None and speaker_latent is None:
        speaker_latent, speaker_mask = get_speaker_latent_and_mask(
            fish_ae, pca_state, speaker_audio.to(device),
            max_speaker_latent_length=pad_to_max_speaker_latent_length or MAX_SPEAKER_LATENT_LENGTH,
            pad_to_max=(pad_to_max_speaker_latent_length is not None))
    if speaker_latent is None:
        n = pad_to_max_speaker_latent_length or 4
        speaker_latent = torch.zeros((1, n, 80), device=device, dtype=model.dtype)
        speaker_mask = torch.zeros((1, n), device=device, dtype=torch.bool)
    S = getattr(sample_fn, "keywords", {}).get("sequence_length") or 640
    noise = torch.cat([torch.randn((1, S, model.cfg.latent_size), device=device, dtype=torch.float32,
                                   generator=torch.Generator(device=device).manual_seed(int(sd))) for sd in rng_seeds])
    spk = speaker_latent.to(device).expand(B, *speaker_latent.shape[1:])
    smk = speaker_mask.to(device).expand(B, *speaker_mask.shape[1:])
    latent = sample_fn(model, spk, smk, ids, mask, int(rng_seeds[0]), noise=noise, **kw)
    audio = ae_decode(fish_ae, pca_state, latent)
    return [(crop_audio_to_flattening_point(audio[i:i + 1], latent[i]), norm[i]) for i in range(B)]


@torch.inference_mode()
def stream_blockwise_audio(model, fish_ae, pca_state, sample_blockwise_fn: Callable, speaker_latent, speaker_mask,
                           text_input_ids, text_mask, rng_seed: int, block_sizes: Sequence[int], on_audio=None,
                           **sampler_kwargs):
    """Streaming synthesis: audio of block i is decoded while block i + 1 is being sampled.
    `sample_blockwise_fn` is the blockwise sampler (or a functools.partial of it, as handler._build_sample_fn would
    build). Every finished block of latents goes through a STATEFUL streaming decode (`B200DAC.new_stream`: the decoder
    is exactly causal; the stream carries the window-128 attention keys / values and the conv halos), so a block costs
    what its own latents cost and the samples are bit-identical to decoding everything at the end.
    Returns (final_latents, [(audio_block (B, 1, n) fp32 on the device, ready_event), ...]); `ready_event` completes
    when that block's audio is final -- for the first block long before the sampler has finished. The call itself
    returns once everything is enqueued (the CUDA launch queue back-pressures the host), so a consumer that wants the
    early blocks polls the events from another thread or from `on_audio`.
    `on_audio(index, audio_block, ready_event)` (optional) is invoked on the host right after block `index`'s decode
    has been enqueued. A `continuation_latent` is decoded first (its audio is not returned: the caller already has it)
    so that the stream state matches the prefix."""
    blocks = []
    B = text_input_ids.shape[0]
    cont = sampler_kwargs.get("continuation_latent")
    total = sum(block_sizes) + (cont.shape[1] if cont is not None else 0)
    streams = [fish_ae.borrow_stream(total) for _ in range(B)]  # pooled per decoder: no allocation after the first request
    try:
        if cont is not None and cont.shape[1] > 0:
            for b, st in enumerate(streams):
                st.decode(pca_state, cont[b:b + 1])

        def on_block(index, start, length, prefix):
            # enqueued behind the block's kernels; one stream (= one sequence) per batch item
            audio = torch.cat([st.decode(pca_state, prefix[b:b + 1, start:start + length]) for b, st in enumerate(streams)])
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            blocks.append((audio, ev))
            if on_audio is not None:
                on_audio(index, audio, ev)

        latents = sample_blockwise_fn(model, speaker_latent, speaker_mask, text_input_ids, text_mask, rng_seed,
                                      list(block_sizes), on_block=on_block, **sampler_kwargs)
    finally:
        # back to the pool without synchronising: the next user resets the state through the same handle, which orders the
        # work on the device (HandleScope)
        for st in streams:
            fish_ae.return_stream(st)
    return latents, blocks


# ------------------------------------------------------------------------------------------------ stitching (host)
def crossfade_chunks(audio_chunks: Sequence[torch.Tensor], overlap_samples: int = 4410) -> torch.Tensor:
    """Linear 100 ms cross-fades; the overlap shrinks to a quarter of the shorter side."""
    if not audio_chunks:
        return torch.tensor([])
    pieces = [audio_chunks[0]]
    total = audio_chunks[0].shape[-1]
    for cur in audio_chunks[1:]:
        ov = min(overlap_samples, cur.shape[-1] // 4, total // 4)
        if ov > 0:
            ramp = torch.linspace(0, 1, ov, device=cur.device)
            down = torch.linspace(1, 0, ov, device=cur.device)
            if cur.dim() == 2:
                ramp, down = ramp.view(1, -1), down.view(1, -1)
            last = pieces.pop()
            # the fade-out may reach back over several short pieces: materialise only the tail that is needed
            while last.shape[-1] < ov:
                last = torch.cat([pieces.pop(), last], dim=-1)
            pieces += [last[..., :-ov], last[..., -ov:] * down + cur[..., :ov] * ramp, cur[..., ov:]]
            total += cur.shape[-1] - ov
        else:
            pieces.append(cur)
            total += cur.shape[-1]
    return torch.cat(pieces, dim=-1)


def _trailing_silence(chunk: torch.Tensor, threshold: float, max_tail: int) -> int:
    tail = chunk[..., -min(chunk.shape[-1], max_tail):].abs().flatten()
    loud = torch.nonzero(tail >= threshold)
    return tail.numel() if loud.numel() == 0 else tail.numel() - 1 - int(loud[-1])


def normalize_chunk_boundaries(audio_chunks: Sequence[torch.Tensor], sample_rate: int = SAMPLE_RATE,
                               silence_threshold: float = 0.01, min_silence_samples: int = 22050) -> torch.Tensor:
    """Every chunk but the last ends with exactly `min_silence_samples` of silence, then cross-fade."""
    if not audio_chunks:
        return torch.tensor([])
    if len(audio_chunks) == 1:
        return audio_chunks[0]
    fixed = []
    for i, chunk in enumerate(audio_chunks):
        if chunk.dim() == 1:
            chunk = chunk.unsqueeze(0)
        if i + 1 < len(audio_chunks):
            quiet = _trailing_silence(chunk, silence_threshold, 2 * min_silence_samples)
            if quiet > min_silence_samples:
                chunk = chunk[..., : chunk.shape[-1] - (quiet - min_silence_samples)]
            elif quiet < min_silence_samples:
                pad = chunk.new_zeros(tuple(chunk.shape[:-1]) + (min_silence_samples - quiet,)).float()
                chunk = torch.cat([chunk, pad.to(chunk.device)], dim=-1)
        fixed.append(chunk)
    return crossfade_chunks(fixed)


def stitch(audio_chunks: Sequence[torch.Tensor], normalize_boundaries: bool = True, enable_crossfade: bool = True):
    """handler.py:763-768."""
    if normalize_boundaries and len(audio_chunks) > 1:
        return normalize_chunk_boundaries(audio_chunks, sample_rate=SAMPLE_RATE)
    if enable_crossfade and len(audio_chunks) > 1:
        return crossfade_chunks(audio_chunks)
    return torch.cat(list(audio_chunks), dim=-1)


# ------------------------------------------------------------------------------------------------ sharding
def shard_units(n_units: int, rank: int, world: int) -> List[int]:
    """Unit i (a text chunk or a whole request) runs on rank i % world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_units, world))


def gather_audio(local: Sequence[Tuple[int, torch.Tensor]], n_units: int, group=None, dst: int = 0):
    """Collect (unit index, 2-D audio) pairs from every rank on `dst`, in unit order. One exchange per job:
    lengths first, then one padded buffer per rank. Works on gloo (CPU tensors) and NCCL (CUDA tensors)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        out = [None] * n_units
        for i, a in local:
            out[i] = a
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    per_rank = (n_units + world - 1) // world
    lens = torch.zeros(per_rank, dtype=torch.int64, device=dev)
    for slot, (_, a) in enumerate(local):
        lens[slot] = a.shape[-1]
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    width = max(int(torch.stack(all_lens).max()), 1)
    buf = torch.zeros(per_rank, width, dtype=torch.float32, device=dev)
    for slot, (_, a) in enumerate(local):
        buf[slot, : a.shape[-1]] = a.reshape(-1).to(dev, torch.float32)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    if rank != dst:
        return None
    out = [None] * n_units
    for r in range(world):
        for slot, i in enumerate(shard_units(n_units, r, world)):
            out[i] = bufs[r][slot, : int(all_lens[r][slot])].unsqueeze(0).cpu()
    return out


def synthesize(text: str, synth_chunk: Optional[Callable[[str, int], torch.Tensor]], seed: int = 0,
               max_chars_per_chunk: Optional[int] = 300, target_duration: float = 10.0,
               normalize_boundaries: bool = True, enable_crossfade: bool = True, group=None, *,
               synth_chunks: Optional[Callable[[List[str], List[int]], List[torch.Tensor]]] = None,
               chunks_per_call: int = 4):
    """The chunk loop of handler._synthesize (handler.py:736-768), sharded over the ranks of `group` when
    torch.distributed is initialised. `synth_chunk(chunk_text, chunk_seed)` returns the chunk's audio (1, n);
    chunk i uses seed + 1000 * i (handler.py:749). Rank 0 returns the stitched audio, other ranks None.
    `synth_chunks(texts, seeds) -> [audio]` (optional, keyword-only extension): when given, a rank's chunks are
    synthesised `chunks_per_call` at a time in one batched call (see `sample_pipeline_batch`) instead of one by one."""
    import torch.distributed as dist
    chunks = chunk_text_for_audio(text, max_chars_per_chunk, target_duration) if max_chars_per_chunk and \
        max_chars_per_chunk > 0 else [text]
    if not chunks:
        raise ValueError("Text is empty after normalization")
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    local = []
    mine = shard_units(len(chunks), rank, world)
    if synth_chunks is not None:
        for j in range(0, len(mine), max(1, chunks_per_call)):
            idx = mine[j:j + max(1, chunks_per_call)]
            for i, a in zip(idx, synth_chunks([chunks[i] for i in idx], [seed + i * 1000 for i in idx])):
                local.append((i, a.reshape(1, -1) if a.dim() != 2 else a))
    else:
        for i in mine:
            a = synth_chunk(chunks[i], seed + i * 1000)
            local.append((i, a.reshape(1, -1) if a.dim() != 2 else a))
    audio = gather_audio(local, len(chunks), group)
    if audio is None:
        return None
    return stitch([a.cpu() for a in audio], normalize_boundaries, enable_crossfade)
