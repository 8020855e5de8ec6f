"""CPU: host side of a synthesis job (echo_tts_b200/pipeline.py) against the oracle restatement and against the
golden vectors produced by the REAL reference host functions (oracle/pin_host_reference.py -> host_pipeline.pt).
Everything here is index / byte / fp32-elementwise work: the bar is bit-exact."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from echo_tts_b200 import pipeline as P
from oracle import host_oracle as H
from tests.util import gold


@pytest.fixture(scope="module")
def g():
    return gold("host_pipeline.pt")


def test_oracle_matches_reference_goldens(g):
    """The oracle itself is pinned: same answers as the reference's own functions recorded in the fixture."""
    ids, mask, norm = H.text_ids_and_mask(g["texts"][:4], 96)
    assert torch.equal(ids, g["tok_ids"]) and torch.equal(mask, g["tok_mask"]) and norm == g["tok_norm"]
    for (t, mc), ref in g["chunks"].items():
        assert H.chunk_text(t, mc) == ref
    for (t, mc, dur), ref in g["chunks_audio"].items():
        assert H.chunk_text_for_audio(t, mc, dur) == ref
    assert [H.find_flattening_point(x) for x in g["flat_latents"]] == g["flat_points"]
    for k, chunks in g["stitch_inputs"].items():
        assert torch.equal(H.crossfade_chunks([c.clone() for c in chunks]), g["crossfade"][k])
        assert torch.equal(H.normalize_chunk_boundaries([c.clone() for c in chunks]), g["normalized"][k])


def test_tokenizer_and_padding(g):
    ids, mask, norm = P.get_text_input_ids_and_mask(g["texts"][:4], max_length=96, return_normalized_text=True)
    assert ids.dtype == torch.int32 and mask.dtype == torch.bool
    assert torch.equal(ids, g["tok_ids"]) and torch.equal(mask, g["tok_mask"]) and norm == g["tok_norm"]
    # truncation at max_length and max_length=None (longest prompt)
    ids2, mask2 = P.get_text_input_ids_and_mask(g["texts"][:4], max_length=10)
    assert torch.equal(ids2, g["tok_ids"][:, :10]) and bool(mask2.all())
    ids3, _ = P.get_text_input_ids_and_mask(["ab", "abcdef"], max_length=None, normalize=False)
    assert tuple(ids3.shape) == (2, 7) and ids3[0].tolist() == [0, 97, 98, 0, 0, 0, 0]
    assert P.tokenizer_encode("x", append_bos=False, normalize=False).tolist() == [120]


def test_chunking_bit_exact(g):
    for (t, mc), ref in g["chunks"].items():
        assert P.chunk_text(t, mc) == ref, (t[:40], mc)
        assert all(len(c) <= mc for c in ref)
    for (t, mc, dur), ref in g["chunks_audio"].items():
        assert P.chunk_text_for_audio(t, mc, dur) == ref, (t[:40], mc, dur)
    with pytest.raises(ValueError):
        P.chunk_text("abc", 0)
    assert P.chunk_text(None) == [] and P.chunk_text(" \n ") == []


def test_chunking_property_random_text():
    """Same chunks as the oracle on random punctuation soup; the concatenation keeps every non-space character."""
    gen = torch.Generator().manual_seed(3)
    alphabet = list("abcdefgh   ,.;:!?\"')]}\n\t") + ["”", "’", "…"]
    for trial in range(60):
        n = int(torch.randint(1, 900, (1,), generator=gen))
        text = "".join(alphabet[int(i)] for i in torch.randint(0, len(alphabet), (n,), generator=gen))
        for mc in (5, 33, 120):
            mine, ref = P.chunk_text(text, mc), H.chunk_text(text, mc)
            assert mine == ref
            assert "".join("".join(c.split()) for c in mine) == "".join(text.split())


def test_flattening_point(g):
    for x, ref in zip(g["flat_latents"], g["flat_points"]):
        assert P.find_flattening_point(x) == ref
    audio = torch.zeros(1, 1, 64 * 2048)
    assert P.crop_audio_to_flattening_point(audio, g["flat_latents"][0]).shape[-1] == g["crop_audio_len"]
    assert P.find_flattening_point(torch.zeros(0, 80)) == 0


def test_stitching_bit_exact(g):
    for k, chunks in g["stitch_inputs"].items():
        assert torch.equal(P.crossfade_chunks([c.clone() for c in chunks]), g["crossfade"][k]), k
        assert torch.equal(P.normalize_chunk_boundaries([c.clone() for c in chunks]), g["normalized"][k]), k
        assert torch.equal(P.stitch(chunks, True, True), g["normalized"][k] if len(chunks) > 1 else chunks[0])
        assert torch.equal(P.stitch(chunks, False, True), g["crossfade"][k])
        assert torch.equal(P.stitch(chunks, False, False), torch.cat(chunks, -1))
    assert P.crossfade_chunks([]).numel() == 0 and P.normalize_chunk_boundaries([]).numel() == 0


def test_shard_units():
    for n in (0, 1, 7, 11, 32):
        for world in (1, 2, 4, 8):
            parts = [P.shard_units(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        P.shard_units(4, 2, 2)


LONG_TEXT = ("The quick brown fox jumps over the lazy dog, twice. " * 40).strip()


def fake_synth(chunk: str, seed: int) -> torch.Tensor:
    """Deterministic stand-in for sampler + DAC decode: audio that depends on the chunk text and its seed only."""
    gen = torch.Generator().manual_seed(seed + len(chunk))
    n = 30000 + 50 * len(chunk)
    a = 0.3 * torch.randn(1, n, generator=gen)
    a[..., -(seed // 1000 % 3) * 9000 - 1:] *= 0.001  # different amounts of trailing silence per chunk
    return a


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def synth(chunk, seed):
        calls.append(seed)
        return fake_synth(chunk, seed)

    audio = P.synthesize(LONG_TEXT, synth, seed=42)
    torch.save({"audio": audio, "calls": calls}, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_synthesis_equals_single_process(tmp_path, world):
    """N > 1: chunks of one long prompt are sharded rank-round-robin over a gloo group; rank 0 stitches on the host.
    The result must equal the single-process job bit for bit, and every chunk must be synthesised exactly once with
    the reference's seed progression seed + 1000 * idx (handler.py:749)."""
    single = P.synthesize(LONG_TEXT, fake_synth, seed=42)
    chunks = P.chunk_text_for_audio(LONG_TEXT, 300, 10.0)
    assert len(chunks) >= 5 and torch.equal(single, H.normalize_chunk_boundaries(
        [fake_synth(c, 42 + 1000 * i) for i, c in enumerate(chunks)]))
    out = str(tmp_path / "shard")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    assert torch.equal(res[0]["audio"], single)
    assert all(r["audio"] is None for r in res[1:])
    seeds = sorted(s for r in res for s in r["calls"])
    assert seeds == [42 + 1000 * i for i in range(len(chunks))]
    for r in range(world):
        assert res[r]["calls"] == [42 + 1000 * i for i in P.shard_units(len(chunks), r, world)]


def test_voice_cache_lru_host_logic():
    """VoiceCache bookkeeping on the CPU with stub encoders: hit/miss counting, least-recently-used eviction by bytes,
    KeyError for an unknown voice with nothing to build it from (per-voice persistence)."""
    from echo_tts_b200 import pipeline as P
    calls = {"encode": 0, "kv": 0}

    def encode(audio):
        calls["encode"] += 1
        n = audio.shape[1] // 2048 // 4 * 4
        return torch.zeros(1, n, 80), torch.ones(1, n, dtype=torch.bool)

    def build_kv(latent):
        calls["kv"] += 1
        p = latent.shape[1] // 4
        return [(torch.zeros(1, p, 2, 128, dtype=torch.bfloat16), torch.zeros(1, p, 2, 128, dtype=torch.bfloat16))
                for _ in range(3)]

    one = P.Voice(*encode(torch.zeros(1, 2048 * 8)), build_kv(torch.zeros(1, 8, 80)))
    cache = P.VoiceCache(model=None, max_bytes=int(2.5 * one.nbytes), encode=encode, build_kv=build_kv)
    calls.update(encode=0, kv=0)
    a = cache.get("a", audio=torch.zeros(1, 2048 * 8))
    assert cache.get("a") is a and calls == {"encode": 1, "kv": 1} and (cache.hits, cache.misses) == (1, 1)
    cache.get("b", audio=torch.zeros(1, 2048 * 8))
    cache.get("a")                                        # a is now the most recently used
    cache.get("c", audio=torch.zeros(1, 2048 * 8))        # 3 voices > 2.5: the least recently used (b) goes
    assert "a" in cache and "c" in cache and "b" not in cache and len(cache) == 2
    assert cache.nbytes == 2 * one.nbytes
    with pytest.raises(KeyError):
        cache.get("b")
    cache.get("d", speaker_latent=one.speaker_latent, speaker_mask=one.speaker_mask)  # pre-encoded: no encode call
    assert calls["encode"] == 3 and calls["kv"] == 4
    cache.drop("d")
    assert "d" not in cache


def test_synthesize_batched_callback_equals_per_chunk_callback():
    """`synthesize(..., synth_chunks=...)` hands a rank's chunks to one batched callback `chunks_per_call` at a time: same
    chunks, same seeds (seed + 1000 * idx, handler.py:749), same stitched result as the per-chunk callback."""
    seen = []

    def batched(texts, seeds):
        seen.append(list(seeds))
        return [fake_synth(t, s) for t, s in zip(texts, seeds)]

    ref = P.synthesize(LONG_TEXT, fake_synth, seed=42)
    for per_call in (1, 3, 100):
        seen.clear()
        got = P.synthesize(LONG_TEXT, None, seed=42, synth_chunks=batched, chunks_per_call=per_call)
        assert torch.equal(got, ref)
        n = len(P.chunk_text_for_audio(LONG_TEXT, 300, 10.0))
        assert [s for call in seen for s in call] == [42 + 1000 * i for i in range(n)]
        assert max(len(c) for c in seen) <= per_call
