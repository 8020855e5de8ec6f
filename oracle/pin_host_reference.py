"""Pins oracle/host_oracle.py against the REAL reference host functions and writes tests/golden/host_pipeline.pt.

Runs only in the build container (needs /root/reference). inference.py is imported as is (torchcodec stubbed);
handler.py cannot be imported (runpod / boto3 are absent), so its three pure functions -- chunk_text_for_audio,
crossfade_chunks, normalize_chunk_boundaries (handler.py:102-240) -- are lifted out of its source with `ast` and
executed unmodified in a namespace that holds `torch` and the reference's own chunk_text.

  python oracle/pin_host_reference.py
"""
from __future__ import annotations

import ast
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import host_oracle as H  # noqa: E402
from oracle.pin_reference import GOLD, REF, import_reference  # noqa: E402


def lift_handler_functions(names):
    src = open(os.path.join(REF, "handler.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "re": __import__("re")}
    exec("_WHITESPACE_RE = re.compile(r'\\s+')", ns)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "handler.py", "exec"), ns)
    return ns


TEXTS = [
    "Hello from Echo-TTS on B200.",
    "[S1] Already tagged: with a colon; a semicolon — and a dash… plus ‘quotes’ and ”these”.",
    "(laughs) Parenthesised start\nwith a newline.",
    "Speaker S2 mentioned inline, no prefix expected.",
    ("The quick brown fox jumps over the lazy dog. " * 70).strip(),
    "One very long sentence without any punctuation at all " * 30,
    "Short. Tiny tail! Ok?",
    "Clause one, clause two, clause three, " * 25 + "and the end.",
    "He said \"stop.\" Then 'go!' (really?) [sure.] {fine,} done " * 12,
    "   \n\t  ",
    "",
]


def main():
    _, inference, _, _ = import_reference()
    hz = lift_handler_functions({"chunk_text", "chunk_text_for_audio", "crossfade_chunks", "normalize_chunk_boundaries"})
    g = {"texts": TEXTS}

    # ---- tokenizer / padding
    ids, mask, norm = inference.get_text_input_ids_and_mask(TEXTS[:4], max_length=96, return_normalized_text=True)
    o_ids, o_mask, o_norm = H.text_ids_and_mask(TEXTS[:4], 96)
    assert torch.equal(ids, o_ids) and torch.equal(mask, o_mask) and norm == o_norm
    g["tok_ids"], g["tok_mask"], g["tok_norm"] = ids, mask, norm

    # ---- chunking (inference.chunk_text, handler.chunk_text, handler.chunk_text_for_audio)
    g["chunks"], g["chunks_audio"] = {}, {}
    for t in TEXTS:
        for mc in (300, 120, 40, 7):
            ref = inference.chunk_text(t, max_chars=mc)
            assert ref == hz["chunk_text"](t, max_chars=mc) == H.chunk_text(t, mc), (t[:30], mc)
            g["chunks"][(t, mc)] = ref
        for mc, dur in ((300, 10.0), (300, 30.0), (50, 10.0)):
            ref = hz["chunk_text_for_audio"](t, max_chars=mc, target_duration_seconds=dur)
            assert ref == H.chunk_text_for_audio(t, mc, dur)
            g["chunks_audio"][(t, mc, dur)] = ref

    # ---- flattening point: latents that go flat at a known index, never, immediately; flat but off target
    gen = torch.Generator().manual_seed(5)
    lat_cases = []
    for n, flat_from, level in ((64, 30, 0.0), (64, 64, 0.0), (40, 0, 0.0), (50, 20, 0.5), (25, 10, 0.02), (640, 411, 0.0)):
        x = torch.randn(n, 80, generator=gen)
        x[flat_from:] = level + 0.01 * torch.randn(n - flat_from, 80, generator=gen)
        lat_cases.append(x)
    g["flat_latents"] = lat_cases
    g["flat_points"] = [inference.find_flattening_point(x) for x in lat_cases]
    assert g["flat_points"] == [H.find_flattening_point(x) for x in lat_cases]
    audio = torch.randn(1, 1, 64 * 2048, generator=gen)
    ref = inference.crop_audio_to_flattening_point(audio, lat_cases[0])
    assert torch.equal(ref, H.crop_audio_to_flattening_point(audio, lat_cases[0]))
    g["crop_audio_len"] = ref.shape[-1]

    # ---- crossfade / boundary normalisation: 2-D (1, n) chunks as sample_pipeline returns audio_out[0]-style rows
    def mk(n, tail_silence, seed, amp=0.5):
        x = amp * torch.randn(1, n, generator=torch.Generator().manual_seed(seed))
        x = torch.where(x.abs() < 0.02, torch.full_like(x, 0.02), x)  # "speech": never below the silence threshold
        if tail_silence:
            x[..., n - tail_silence:] = 0.003 * torch.randn(1, tail_silence, generator=torch.Generator().manual_seed(seed + 1))
        return x

    sets = {
        "three_mixed": [mk(60000, 0, 1), mk(90000, 30000, 2), mk(50000, 5000, 3)],
        "short_chunks": [mk(3000, 100, 4), mk(800, 0, 5), mk(10, 0, 6), mk(20000, 20000, 7)],
        "single": [mk(5000, 0, 8)],
        "all_silent_tail": [mk(30000, 30000, 9), mk(40000, 0, 10)],
        "one_dim": [mk(9000, 0, 11)[0], mk(12000, 2000, 12)[0]],
    }
    g["stitch_inputs"] = sets
    g["crossfade"], g["normalized"] = {}, {}
    for k, chunks in sets.items():
        ref = hz["crossfade_chunks"]([c.clone() for c in chunks])
        assert torch.equal(ref, H.crossfade_chunks([c.clone() for c in chunks])), k
        g["crossfade"][k] = ref
        ref = hz["normalize_chunk_boundaries"]([c.clone() for c in chunks], sample_rate=44100)
        assert torch.equal(ref, H.normalize_chunk_boundaries([c.clone() for c in chunks])), k
        g["normalized"][k] = ref
    assert hz["crossfade_chunks"]([]).numel() == 0 and H.crossfade_chunks([]).numel() == 0

    path = os.path.join(GOLD, "host_pipeline.pt")
    torch.save(g, path)
    print(f"host oracle pinned against the reference; wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
