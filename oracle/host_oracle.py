"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's host-side steps either side of the sampling path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this module; the product
(echo_tts_b200/pipeline.py) never does. Every function follows the reference line by line (pure-Python loops where
the reference loops) so that it can be compared with the reference's own outputs; oracle/pin_host_reference.py
executes the REAL reference functions in the build container and stores their outputs in
tests/golden/host_pipeline.pt, which pins this file.

  tokenizer_encode / text_ids_and_mask      reference inference.py:115-138, 192-215
  chunk_text                                reference inference.py:140-190 (== handler.py:49-99)
  chunk_text_for_audio                      reference handler.py:102-123
  find_flattening_point / crop              reference inference.py:288-301
  crossfade_chunks                          reference handler.py:126-171
  normalize_chunk_boundaries                reference handler.py:174-240
"""
from __future__ import annotations

import re
from typing import List, Tuple

import torch

_WS = re.compile(r"\s+")


def tokenizer_encode(text: str, append_bos: bool = True, normalize: bool = True) -> Tuple[List[int], str]:
    """inference.py:115-138: typographic normalisation, '[S1] ' prefix, UTF-8 bytes, BOS = 0."""
    if normalize:
        for a, b in (("…", "..."), ("’", "'"), ("”", '"'), ("\n", " "), (":", ","), (";", ","),
                     ("—", ", ")):
            text = text.replace(a, b)
        if not text.startswith("[") and not text.startswith("(") and "S1" not in text and "S2" not in text:
            text = "[S1] " + text
    b = list(text.encode("utf-8"))
    if append_bos:
        b.insert(0, 0)
    return b, text


def text_ids_and_mask(texts: List[str], max_length, normalize: bool = True):
    """inference.py:192-215 (pad_to_max irrelevant: both branches give (n, max_length))."""
    enc = [tokenizer_encode(t, normalize=normalize) for t in texts]
    if max_length is None:
        max_length = max(len(e) for e, _ in enc)
    ids = torch.zeros((len(texts), max_length), dtype=torch.int32)
    mask = torch.zeros((len(texts), max_length), dtype=torch.bool)
    for i, (e, _) in enumerate(enc):
        n = min(len(e), max_length)
        ids[i, :n] = torch.tensor(e[:n], dtype=torch.int32)
        mask[i, :n] = True
    return ids, mask, [t for _, t in enc]


def chunk_text(text: str, max_chars: int = 300) -> List[str]:
    """inference.py:140-190: greedy split at the LAST sentence end, else clause end, else space inside the window."""
    if max_chars <= 0:
        raise ValueError("max_chars must be > 0")
    text = _WS.sub(" ", (text or "")).strip()
    if not text:
        return []
    if len(text) <= max_chars:
        return [text]
    sentence_enders, clause_enders = {".", "!", "?"}, {",", ";", ":"}
    closers = {'"', "'", ")", "]", "}", "”", "’"}
    chunks, remaining = [], text
    while remaining:
        if len(remaining) <= max_chars:
            chunks.append(remaining)
            break
        window = remaining[: max_chars + 1]
        cand_sentence = cand_clause = cand_space = None
        for i in range(1, len(window)):
            if not window[i].isspace():
                continue
            cand_space = i
            prev = window[i - 1]
            prev2 = window[i - 2] if i >= 2 else ""
            if prev in sentence_enders or (prev in closers and prev2 in sentence_enders):
                cand_sentence = i
            elif prev in clause_enders or (prev in closers and prev2 in clause_enders):
                cand_clause = i
        split_at = cand_sentence or cand_clause or cand_space or max_chars
        chunk = remaining[:split_at].strip()
        if chunk:
            chunks.append(chunk)
        remaining = remaining[split_at:].strip()
    return chunks


def chunk_text_for_audio(text: str, max_chars: int = 300, target_duration_seconds: float = 10.0) -> List[str]:
    """handler.py:102-123: ~12 chars per second of speech; a last chunk shorter than 24 chars joins its neighbour."""
    target_chars = min(max_chars, int(target_duration_seconds * 12))
    chunks = chunk_text(text, max_chars=target_chars)
    if len(chunks) > 1 and len(chunks[-1]) < 24:
        chunks[-2] += " " + chunks[-1]
        chunks.pop()
    return chunks


def find_flattening_point(data: torch.Tensor, target_value=0.0, window_size=20, std_threshold=0.05) -> int:
    """inference.py:288-296: first i whose 20-latent window (zero padded past the end) has std < 0.05 (unbiased, over
    all window elements) and |mean - target| < 0.1."""
    padded = torch.cat([data, torch.zeros(window_size, *data.shape[1:], dtype=data.dtype)])
    for i in range(len(padded) - window_size):
        w = padded[i:i + window_size]
        if w.std() < std_threshold and abs(w.mean() - target_value) < 0.1:
            return i
    return len(data)


def crop_audio_to_flattening_point(audio: torch.Tensor, latent: torch.Tensor) -> torch.Tensor:
    """inference.py:298-301."""
    return audio[..., : find_flattening_point(latent) * 2048]


def crossfade_chunks(chunks: List[torch.Tensor], overlap_samples: int = 4410) -> torch.Tensor:
    """handler.py:126-171: linear 100 ms fades; overlap clipped to a quarter of either side."""
    if len(chunks) <= 1:
        return torch.cat(chunks, dim=-1) if chunks else torch.tensor([])
    result = chunks[0]
    for i in range(1, len(chunks)):
        ov = min(overlap_samples, chunks[i].shape[-1] // 4, result.shape[-1] // 4)
        if ov > 0:
            fade_out, fade_in = torch.linspace(1, 0, ov), torch.linspace(0, 1, ov)
            if chunks[i].dim() == 2:
                fade_out, fade_in = fade_out.view(1, -1), fade_in.view(1, -1)
            tail = result[..., -ov:] * fade_out
            result = result[..., :-ov]
            head = chunks[i][..., :ov] * fade_in
            result = torch.cat([result, tail + head, chunks[i][..., ov:]], dim=-1)
        else:
            result = torch.cat([result, chunks[i]], dim=-1)
    return result


def normalize_chunk_boundaries(chunks: List[torch.Tensor], sample_rate: int = 44100, silence_threshold: float = 0.01,
                               min_silence_samples: int = 22050) -> torch.Tensor:
    """handler.py:174-240: every chunk but the last ends with exactly min_silence_samples of "silence"
    (|x| < threshold counted backwards over at most the last 2*min_silence_samples), then crossfade."""
    if not chunks:
        return torch.tensor([])
    if len(chunks) == 1:
        return chunks[0]
    out = []
    for i, chunk in enumerate(chunks):
        if chunk.dim() == 1:
            chunk = chunk.unsqueeze(0)
        if i < len(chunks) - 1:
            tail_samples = min(chunk.shape[-1], min_silence_samples * 2)
            flat = torch.abs(chunk[..., -tail_samples:]).flatten()
            trailing = 0
            for j in range(len(flat) - 1, -1, -1):
                if flat[j] < silence_threshold:
                    trailing += 1
                else:
                    break
            if trailing > min_silence_samples:
                chunk = chunk[..., : -(trailing - min_silence_samples)]
            elif 0 < trailing < min_silence_samples:
                chunk = torch.cat([chunk, torch.zeros(*chunk.shape[:-1], min_silence_samples - trailing)], dim=-1)
            elif trailing == 0:
                chunk = torch.cat([chunk, torch.zeros(*chunk.shape[:-1], min_silence_samples)], dim=-1)
        out.append(chunk)
    return crossfade_chunks(out)
