"""GPU: the host pipeline (echo_tts_b200/pipeline.py) driving the CUDA sampler + DAC decoder end to end, tiny config.
Mirrors what handler._synthesize does with the drop-in objects (reference handler.py:736-768, inference.py:309-347)."""
import functools

import pytest
import torch

import echo_tts_b200
from echo_tts_b200 import pipeline as P
from echo_tts_b200.config import DacConfig, DitConfig
from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state
from tests.util import PLAIN_KNOBS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack():
    from echo_tts_b200.autoencoder import B200DAC, PCAState
    from echo_tts_b200.model import B200EchoDiT
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    cfg, dcfg = DitConfig.tiny(), DacConfig.tiny()
    model = B200EchoDiT.from_state_dict(make_dit_weights(cfg, 1234), cfg, "cuda:0")
    dac = B200DAC.from_state_dict(make_dac_weights(dcfg, 4321), dcfg, "cuda:0")
    comps, mean, scale = make_pca_state(dcfg)
    pca = PCAState(comps.cuda(), mean.cuda(), scale)
    # the sample_fn exactly as handler._build_sample_fn builds it: functools.partial over keyword knobs
    sample_fn = functools.partial(sample, **dict(PLAIN_KNOBS, num_steps=4), sequence_length=24)
    echo_tts_b200.set_deterministic(True)
    yield model, dac, pca, sample_fn
    echo_tts_b200.set_deterministic(False)


def test_sample_pipeline_matches_manual_composition(stack):
    from echo_tts_b200.autoencoder import ae_decode
    model, dac, pca, sample_fn = stack
    audio, norm = P.sample_pipeline(model, dac, pca, sample_fn, "Hello there: friend", rng_seed=7, pad_to_max_text_length=64)
    assert norm == "[S1] Hello there, friend"
    ids, mask = P.get_text_input_ids_and_mask(["Hello there: friend"], 64, device="cuda")
    spk = torch.zeros(1, 4, 80, device="cuda", dtype=model.dtype)
    smask = torch.zeros(1, 4, dtype=torch.bool, device="cuda")
    lat = sample_fn(model, spk, smask, ids, mask, 7)
    ref = P.crop_audio_to_flattening_point(ae_decode(dac, pca, lat), lat[0])
    assert audio.dtype == torch.float32 and audio.dim() == 3 and audio.shape[:2] == (1, 1)
    assert audio.shape[-1] % 2048 == 0 and audio.shape[-1] <= 24 * 2048
    assert torch.equal(audio, ref)


def test_sample_pipeline_with_speaker_latents(stack):
    model, dac, pca, sample_fn = stack
    spk = torch.randn(1, 16, 80, generator=torch.Generator().manual_seed(3)).cuda()
    smask = torch.ones(1, 16, dtype=torch.bool, device="cuda")
    a, _ = P.sample_pipeline(model, dac, pca, sample_fn, "[S1] Same text.", rng_seed=1, pad_to_max_text_length=64,
                             speaker_latent=spk, speaker_mask=smask)
    b, _ = P.sample_pipeline(model, dac, pca, sample_fn, "[S1] Same text.", None, 1, pad_to_max_text_length=64)  # the reference's positional order
    assert torch.isfinite(a).all() and a.abs().max() <= 1.0
    n = min(a.shape[-1], b.shape[-1])
    assert n == 0 or not torch.equal(a[..., :n], b[..., :n])  # the speaker condition changes the result


def test_synthesize_long_prompt_chunks_and_stitches(stack):
    """Long prompt -> chunk_text_for_audio -> one sample_pipeline call per chunk with seed + 1000 * idx -> boundary
    normalisation + crossfade on the host, as handler._synthesize (handler.py:736-768)."""
    model, dac, pca, sample_fn = stack
    text = ("This is the first sentence of a long prompt. " * 8).strip()

    def synth_chunk(chunk, seed):
        audio, _ = P.sample_pipeline(model, dac, pca, sample_fn, chunk, rng_seed=seed, pad_to_max_text_length=160)
        return audio[0]  # (1, n) like handler's audio chunks

    out = P.synthesize(text, synth_chunk, seed=11, max_chars_per_chunk=300, target_duration=10.0)
    chunks = P.chunk_text_for_audio(text, 300, 10.0)
    assert len(chunks) >= 3
    manual = P.normalize_chunk_boundaries([synth_chunk(c, 11 + 1000 * i).cpu() for i, c in enumerate(chunks)])
    assert out.dim() == 2 and torch.equal(out, manual)
    plain = P.synthesize(text, synth_chunk, seed=11, normalize_boundaries=False, enable_crossfade=False)
    assert plain.shape[-1] == sum(synth_chunk(c, 11 + 1000 * i).shape[-1] for i, c in enumerate(chunks))


def test_sample_pipeline_with_speaker_audio(stack):
    """Raw speaker audio -> get_speaker_latent_and_mask (DAC encoder on the GPU) -> sampler -> decode, the complete
    reference sample_pipeline (inference.py:309-347). Equals the two-step call with pre-encoded latents."""
    from echo_tts_b200.autoencoder import B200DAC
    model, _, pca, sample_fn = stack
    dcfg = DacConfig.tiny()
    dac = B200DAC.from_state_dict(make_dac_weights(dcfg, 4321, include_encoder=True), dcfg, "cuda:0")
    wav = 0.3 * torch.randn(1, 70 * dcfg.frame_length + 5, generator=torch.Generator().manual_seed(2))
    spk, smask = P.get_speaker_latent_and_mask(dac, pca, wav.cuda())
    assert spk.shape[1] == 68 and bool(smask.all())  # 70 complete frames -> trimmed to a multiple of 4
    a, _ = P.sample_pipeline(model, dac, pca, sample_fn, "[S1] With a voice.", rng_seed=3, pad_to_max_text_length=64,
                             speaker_audio=wav)
    b, _ = P.sample_pipeline(model, dac, pca, sample_fn, "[S1] With a voice.", rng_seed=3, pad_to_max_text_length=64,
                             speaker_latent=spk, speaker_mask=smask)
    assert torch.equal(a, b) and torch.isfinite(a).all()


def test_streaming_blockwise_equals_offline(stack):
    """audio decoded block by block while the next block samples == decoding the final latents once
    (the DAC is causal), and the latents equal the non-streaming sampler's."""
    from echo_tts_b200.autoencoder import ae_decode
    from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as blockwise
    model, dac, pca, _ = stack
    ids, mask = P.get_text_input_ids_and_mask(["[S1] Streaming test."], 64, device="cuda")
    spk = torch.randn(1, 16, 80, generator=torch.Generator().manual_seed(5)).cuda()
    smask = torch.ones(1, 16, dtype=torch.bool, device="cuda")
    knobs = dict(PLAIN_KNOBS, num_steps=4)
    blocks = [8, 8, 4]
    rng = torch.Generator().manual_seed(9)
    nb = [torch.randn((1, b, 80), generator=rng) for b in blocks]
    offline = blockwise(model, spk, smask, ids, mask, 0, blocks, noise_blocks=nb, **knobs)
    lat, parts = P.stream_blockwise_audio(model, dac, pca, blockwise, spk, smask, ids, mask, 0, blocks, noise_blocks=nb, **knobs)
    assert torch.equal(lat, offline) and len(parts) == 3
    for _, ev in parts:
        ev.synchronize()
    streamed = torch.cat([a for a, _ in parts], dim=-1)
    full = ae_decode(dac, pca, lat)
    assert streamed.shape == full.shape == (1, 1, 20 * 2048)
    assert [a.shape[-1] for a, _ in parts] == [8 * 2048, 8 * 2048, 4 * 2048]
    assert torch.equal(streamed, full)  # exact: causal decoder, deterministic kernels


def test_voice_cache_skips_the_speaker_encoder_and_changes_nothing(stack):
    """SURVEY 8 f4, per-voice persistence: a cached speaker KV gives bit-identical latents (deterministic mode), for
    the plain sampler, with speaker_kv_scale (the sampler must scale a COPY: the stored cache stays untouched) and for
    the blockwise sampler; and it saves the speaker encoder's kernel launches."""
    from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as blockwise
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    from tests.util import EULER_KNOBS
    model, dac, pca, _ = stack
    ids, mask = P.get_text_input_ids_and_mask(["[S1] Same voice again."], 64, device="cuda")
    spk = torch.randn(1, 16, 80, generator=torch.Generator().manual_seed(11)).cuda()
    smask = torch.ones(1, 16, dtype=torch.bool, device="cuda")
    cache = P.VoiceCache(model, dac, pca)
    voice = cache.get("alice", speaker_latent=spk, speaker_mask=smask)
    assert cache.get("alice") is voice and (cache.hits, cache.misses) == (1, 1)
    assert len(voice.kv) == model.cfg.num_layers and voice.kv[0][0].shape[:2] == (1, 4)
    pristine = [(k.clone(), v.clone()) for k, v in voice.kv]
    noise = torch.randn(1, 24, 80, generator=torch.Generator().manual_seed(12))
    for knobs in (dict(PLAIN_KNOBS, num_steps=4), dict(EULER_KNOBS, num_steps=4)):  # the latter scales the speaker KV
        n0 = model.h.num_launches()
        ref = sample(model, spk, smask, ids, mask, 0, sequence_length=24, noise=noise, **knobs)
        n1 = model.h.num_launches()
        got = sample(model, spk, smask, ids, mask, 0, sequence_length=24, noise=noise, speaker_kv_cache=voice.kv, **knobs)
        n2 = model.h.num_launches()
        assert torch.equal(got, ref)
        assert n2 - n1 < n1 - n0  # no speaker encoder, no K/V projections
        for (k, v), (k0, v0) in zip(voice.kv, pristine):
            assert torch.equal(k, k0) and torch.equal(v, v0)
    nb = [torch.randn((1, b, 80), generator=torch.Generator().manual_seed(13 + b)) for b in (8, 8)]
    kn = dict(EULER_KNOBS, num_steps=4)
    ref = blockwise(model, spk, smask, ids, mask, 0, [8, 8], noise_blocks=nb, **kn)
    got = blockwise(model, spk, smask, ids, mask, 0, [8, 8], noise_blocks=nb, speaker_kv_cache=voice.kv, **kn)
    assert torch.equal(got, ref)
    # through sample_pipeline, as a server would use it
    sample_fn = functools.partial(sample, **dict(PLAIN_KNOBS, num_steps=4), sequence_length=24)
    a, _ = P.sample_pipeline(model, dac, pca, sample_fn, "[S1] Same voice again.", rng_seed=3, pad_to_max_text_length=64,
                             voice=voice)
    b, _ = P.sample_pipeline(model, dac, pca, sample_fn, "[S1] Same voice again.", rng_seed=3,
                             pad_to_max_text_length=64, speaker_latent=spk, speaker_mask=smask)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        sample(model, spk, smask, ids, mask, 0, sequence_length=24, noise=noise, speaker_kv_cache=voice.kv[:-1],
               **dict(PLAIN_KNOBS, num_steps=4))


# ------------------------------------------------------------------------------------------------ N > 1 over NCCL
NCCL_TEXT = ("This is the first sentence of a long prompt. " * 8).strip()


def _tiny_synth_chunk(device):
    """The real tiny sampler + DAC decode on `device`, as a `synth_chunk(chunk_text, seed)` for pipeline.synthesize."""
    from echo_tts_b200.autoencoder import B200DAC, PCAState
    from echo_tts_b200.model import B200EchoDiT
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample
    cfg, dcfg = DitConfig.tiny(), DacConfig.tiny()
    model = B200EchoDiT.from_state_dict(make_dit_weights(cfg, 1234), cfg, device)
    dac = B200DAC.from_state_dict(make_dac_weights(dcfg, 4321), dcfg, device)
    comps, mean, scale = make_pca_state(dcfg)
    pca = PCAState(comps.to(device), mean.to(device), scale)
    sample_fn = functools.partial(sample, **dict(PLAIN_KNOBS, num_steps=4), sequence_length=24)

    def synth_chunk(chunk, seed):
        audio, _ = P.sample_pipeline(model, dac, pca, sample_fn, chunk, rng_seed=seed, pad_to_max_text_length=160)
        return audio[0]

    return synth_chunk


def _nccl_worker(rank, world, port, out_path):
    import os
    import sys

    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    echo_tts_b200.set_deterministic(True)
    calls = []
    synth = _tiny_synth_chunk(f"cuda:{rank}")

    def synth_chunk(chunk, seed):
        calls.append(seed)
        return synth(chunk, seed)

    audio = P.synthesize(NCCL_TEXT, synth_chunk, seed=11)  # chunks i % world, NCCL all_gather of the audio, rank 0 stitches
    torch.save({"audio": audio, "calls": calls}, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_synthesis_over_nccl_equals_single_process(tmp_path):
    """BASELINE configs[3] in small: the chunks of one long prompt sharded over 2 GPUs (one process per GPU, NCCL
    gather of the finished audio, host stitching on rank 0) must equal the single-process job bit for bit
    (deterministic mode), with the reference's seed progression seed + 1000 * idx (handler.py:746-768)."""
    import socket

    import torch.multiprocessing as mp
    echo_tts_b200.set_deterministic(True)
    try:
        single = P.synthesize(NCCL_TEXT, _tiny_synth_chunk("cuda:0"), seed=11)
    finally:
        echo_tts_b200.set_deterministic(False)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "nccl")
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(2)]
    chunks = P.chunk_text_for_audio(NCCL_TEXT, 300, 10.0)
    assert len(chunks) >= 3
    assert res[1]["audio"] is None and torch.equal(res[0]["audio"], single)
    for r in range(2):
        assert res[r]["calls"] == [11 + 1000 * i for i in P.shard_units(len(chunks), r, 2)]


def test_flattening_point_kernel_bit_exact_vs_reference_goldens():
    """SURVEY 8 f2 on the device: `find_flattening_point` of GPU latents (one warp-per-window kernel through the C
    ABI) must return exactly the reference's index on the fixtures lifted from inference.py:288-301
    (tests/golden/host_pipeline.pt), agree with the host evaluation on random cases, and handle the edges."""
    from oracle import host_oracle as H
    from tests.util import gold
    g = gold("host_pipeline.pt")
    for x, ref in zip(g["flat_latents"], g["flat_points"]):
        assert P.find_flattening_point(x.cuda()) == ref
    audio = torch.zeros(1, 1, 64 * 2048, device="cuda")
    assert P.crop_audio_to_flattening_point(audio, g["flat_latents"][0].cuda()).shape[-1] == g["crop_audio_len"]
    gen = torch.Generator().manual_seed(5)
    for T, cut in ((640, 400), (640, 0), (640, 640), (37, 20), (19, 3), (1, 0), (1, 1), (160, 159)):
        x = torch.randn(T, 80, generator=gen)
        x[cut:] = 0.02 * torch.randn(T - cut, 80, generator=gen)  # flat tail: std 0.02 < 0.05, mean ~ 0
        assert P.find_flattening_point(x.cuda()) == H.find_flattening_point(x) == P.find_flattening_point(x), (T, cut)
    # near the thresholds: windows whose std straddles 0.05
    x = 0.05 * torch.randn(300, 80, generator=gen) * torch.linspace(1.2, 0.8, 300)[:, None]
    assert P.find_flattening_point(x.cuda()) == H.find_flattening_point(x)
    assert P.find_flattening_point(torch.zeros(0, 80, device="cuda")) == 0


def test_sample_pipeline_batch_matches_one_at_a_time(stack):
    """The chunks of one prompt batched into ONE sampler call + ONE decode (`sample_pipeline_batch`, used by
    `synthesize(..., synth_chunks=...)`): every item draws its noise from its own seed exactly as a single call does,
    so the audio equals the one-at-a-time result up to the bf16 noise floor, with identical crop lengths."""
    model, dac, pca, sample_fn = stack
    spk = torch.randn(1, 16, 80, generator=torch.Generator().manual_seed(3)).cuda()
    smask = torch.ones(1, 16, dtype=torch.bool, device="cuda")
    texts = ["[S1] First chunk of a prompt.", "[S1] The second one, a little longer than the first.", "[S1] Third."]
    seeds = [11, 1011, 2011]
    batch = P.sample_pipeline_batch(model, dac, pca, sample_fn, texts, seeds, pad_to_max_text_length=64,
                                    speaker_latent=spk, speaker_mask=smask)
    assert len(batch) == 3
    for (a, norm), t, sd in zip(batch, texts, seeds):
        one, norm1 = P.sample_pipeline(model, dac, pca, sample_fn, t, rng_seed=sd, pad_to_max_text_length=64,
                                       speaker_latent=spk, speaker_mask=smask)
        assert norm == norm1 and a.shape == one.shape
        assert ((a - one).norm() / one.norm().clamp_min(1e-9)).item() < 1e-2
    # through synthesize: the batched callback gives the same job as the per-chunk callback (same stitching)
    text = ("This is the first sentence of a long prompt. " * 8).strip()

    def synth_chunk(chunk, seed):
        return P.sample_pipeline(model, dac, pca, sample_fn, chunk, rng_seed=seed, pad_to_max_text_length=160)[0][0]

    def synth_chunks(chunks, seeds_):
        return [a[0] for a, _ in P.sample_pipeline_batch(model, dac, pca, sample_fn, chunks, seeds_, pad_to_max_text_length=160)]

    ref = P.synthesize(text, synth_chunk, seed=11)
    got = P.synthesize(text, None, seed=11, synth_chunks=synth_chunks, chunks_per_call=2)
    assert got.shape == ref.shape and ((got - ref).norm() / ref.norm()).item() < 1e-2
