#!/usr/bin/env python
"""bench.py -- Echo-TTS sampling hot path on B200: audio-seconds per second (RTF^-1).

One "step" = one full request of BASELINE.json configs[1]: echo-tts-base architecture (random-init), one short [S1]
prompt padded to 768 tokens (as inference.sample_pipeline does), a 10 s synthetic speaker reference (212 latents),
cfg_scale_text 3.0 / cfg_scale_speaker 8.0 / cfg_min_t 0.5, 40 Euler steps at sequence_length 640, batch 1, followed
by the Fish S1-DAC decode of the 640 latents to 1 310 720 samples @ 44.1 kHz = 29.72 s of audio.

  value  : audio-s/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the public drop-in call `pipeline.sample_pipeline` (reference inference.py:309-347: host
           text -> tokens -> H2D, pinned host speaker latents -> H2D, sampler, DAC decode, flattening-point crop on the
           device) with the cropped audio copied back to pinned host memory (D2H), every step
  N > 1  : every rank is a full replica processing its own requests (no data-path collective): weak scaling
  --impl reference : the reference algorithm (oracle port, fp32 PyTorch on the host cores) on a bounded sample
  config.extras    : the other BASELINE configs measured in the same run (not part of `value`):
           N = 1: cfg5 = configs[4] blockwise streaming latency (first audio / total), batch4 = 4 requests per call
           any N: cfg3 = configs[2], 32 requests sharded 32/N per GPU (4 per sampler call), aggregate audio-s/s
                  cfg4 = configs[3], one ~3000-char prompt -> ~300-char chunks i % N -> NCCL gather -> host stitch,
                         job wall-clock latency

Weights are generated on rank 0 and broadcast over NCCL once at start-up (the only collective; none per step).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AUDIO_SECONDS = 640 * 2048 / 44100.0
PROMPT = "[S1] Hello from Echo-TTS on B200."
KNOBS = dict(num_steps=40, cfg_scale_text=3.0, cfg_scale_speaker=8.0, cfg_min_t=0.5, cfg_max_t=1.0,
             truncation_factor=None, rescale_k=None, rescale_sigma=None, speaker_kv_scale=None,
             speaker_kv_max_layers=None, speaker_kv_min_t=None)  # handler.py:431-442 defaults
WORKLOAD = ("configs[1]: echo-tts-base random-init, 1 short [S1] prompt padded to 768 tokens, 10 s synthetic speaker "
            "(212 latents), cfg 3.0/8.0, cfg_min_t 0.5, 40 Euler steps, sequence_length 640, batch 1, + DAC decode "
            "to 1310720 samples")


def tokens(prompt, max_length=768):
    ids = torch.zeros(1, max_length, dtype=torch.int32)
    mask = torch.zeros(1, max_length, dtype=torch.bool)
    b = [0] + list(prompt.encode("utf-8"))
    ids[0, :len(b)] = torch.tensor(b, dtype=torch.int32)
    mask[0, :len(b)] = True
    return ids, mask


def mask_to(mask_h, device):
    """Text mask -> device, with the length of its longest unmasked prefix (host metadata of the prompt) left on the tensor:
    the samplers then skip the text rows behind it (echo_sampler_args::text_valid_len), which are masked out of every
    attention. A mask built by pipeline.get_text_input_ids_and_mask carries the same number."""
    m = mask_h.to(device)
    cols = mask_h.reshape(-1, mask_h.shape[-1]).any(0).nonzero()
    m._echo_valid_len = int(cols.max()) + 1 if cols.numel() else 1
    return m


def request_flops(n_text_valid: int, n_spk_patches: int):
    """Algorithmic FLOPs of one request: 2MNK for the DiT linears, 4*S*keys*D per layer for attention
    over UNMASKED keys, DAC decode 4.34 TFLOP measured by the survey."""
    S, D, L = 640, 2048, 24
    lin_tok = 2 * (5 * D * D + 3 * D * 5888)
    linear = lin_tok * L * S * (20 * 3 + 20 * 1)
    att = 0
    for branch_keys in (S + n_text_valid + n_spk_patches, S + n_spk_patches, S + n_text_valid):
        att += 4 * S * branch_keys * D * L * 20
    att += 4 * S * (S + n_text_valid + n_spk_patches) * D * L * 20
    return dict(linear=float(linear), attention=float(att), dac=4.34e12, total=float(linear + att) + 4.34e12)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def ncu_traffic():
    """DRAM bytes per launch of the main kernels from the committed `ncu --set full` capture (profiles/), or {}."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        return json.load(open(p))["kernels"]
    except Exception:
        return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1379.5), d.get("hbm_gbs", 6542.7), "measured"
    return 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_sample_seconds(sd, dsd, pca, threads):
    """Bounded sample of the workload on the host cores with the oracle (fp32 PyTorch port of the reference):
    one b=1 DiT forward at S=640 (Lt=768 padded, 53 speaker patches) + one DAC decode of T=16 latents."""
    from echo_tts_b200.config import DacConfig, DitConfig
    from oracle import echo_oracle as O
    cfg, dcfg = DitConfig.base(), DacConfig.base()
    torch.set_num_threads(threads)
    ids, mask = tokens(PROMPT)
    spk = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1))
    smask = torch.ones(1, 212, dtype=torch.bool)
    state = {}

    def prepare():
        with torch.inference_mode():
            t0 = time.time()
            state["kt"] = O.kv_cache_text(sd, cfg, ids, mask)
            state["ks"] = O.kv_cache_speaker(sd, cfg, spk)
            state["kv_s"] = time.time() - t0
            state["kt3"], state["ks3"] = O._batch3(state["kt"]), O._batch3(state["ks"])
            state["mask3"] = torch.cat([mask, torch.zeros_like(mask), mask])       # inference.py:474
            state["smask3"] = torch.cat([smask, smask, torch.zeros_like(smask)])   # inference.py:475

    def step(seed):
        """One CFG forward (3 branches stacked, as inference.py:487-494) + one plain forward (inference.py:497-504)
        + a 16-latent DAC decode; a request is 20 of each forward kind, the KV caches and 40 x 16 latents of decode."""
        with torch.inference_mode():
            x = torch.randn(1, 640, 80, generator=torch.Generator().manual_seed(seed))
            t0 = time.time()
            O.dit_forward(sd, cfg, x.repeat(3, 1, 1), torch.full((3,), 0.75), state["mask3"], state["smask3"], state["kt3"],
                          state["ks3"])
            t_cfg = time.time() - t0
            t0 = time.time()
            O.dit_forward(sd, cfg, x, torch.full((1,), 0.25), mask, smask, state["kt"], state["ks"])
            t_plain = time.time() - t0
            z = torch.randn(1, 16, 80, generator=torch.Generator().manual_seed(seed))
            t0 = time.time()
            O.ae_decode(dsd, dcfg, pca[0], pca[1], pca[2], z)
            t_dac = time.time() - t0
        return 20 * t_cfg + 20 * t_plain + state["kv_s"] + 40 * t_dac, t_cfg + t_plain + t_dac

    return prepare, step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from echo_tts_b200.config import DacConfig, DitConfig
    from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = make_dit_weights(DitConfig.base(), seed=1234, include_latent=False)
    dsd = make_dac_weights(DacConfig.base(), seed=4321)
    pca = make_pca_state(DacConfig.base())
    prepare, step = oracle_sample_seconds(sd, dsd, pca, threads)
    prepare()
    for i in range(args.warmup):
        step(i)
    est, wall = [], 0.0
    for i in range(args.steps):
        e, w = step(100 + i)
        est.append(e)
        wall += w
    req_s = sum(est) / len(est)
    value = AUDIO_SECONDS / req_s
    sample = ("per step: one fp32 CFG forward (3 x 640 rows) + one plain forward (640 rows) at Lt=768 padded / 53 speaker "
              "patches + one DAC decode of 16 latents; a request = 20 x each forward + KV caches + 40 x decode(16)")
    line = {"impl": "reference", "metric": "audio-sec/sec (RTF^-1) per GPU at seq 640/40 steps", "value": value,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample,
                             "extrapolated": True},
            "extrapolated": True,
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "estimated_request_seconds": req_s}
    emit(line)


# ------------------------------------------------------------------------------------------------ B200 arm
def load_models(device, rank, world):
    """Rank 0 generates the deterministic checkpoints; other ranks receive them over NCCL (one-time broadcast)."""
    import torch.distributed as dist
    from echo_tts_b200.autoencoder import B200DAC, PCAState
    from echo_tts_b200.config import DacConfig, DitConfig
    from echo_tts_b200.model import B200EchoDiT
    from echo_tts_b200.weights import dac_param_specs, dit_param_specs, iter_dit_weights, make_dac_weights, make_pca_state
    cfg, dcfg = DitConfig.base(), DacConfig.base()
    model, dac = B200EchoDiT(cfg, device), B200DAC(dcfg, device)

    # the latent_* (blockwise) tensors are loaded too: configs[4] (extras.cfg5) needs them, configs[1] never touches them
    if world == 1:
        model.load_state_dict(iter_dit_weights(cfg, 1234, include_latent=True))
        dac.load_state_dict(make_dac_weights(dcfg, 4321))
    else:
        specs = [(k, s) for k, s, _ in dit_param_specs(cfg)]
        gen = iter_dit_weights(cfg, 1234, include_latent=True) if rank == 0 else None

        def dit_items():
            for k, shape in specs:
                if rank == 0:
                    kk, t = next(gen)
                    assert kk == k
                    t = t.to(device=device, dtype=torch.bfloat16)
                else:
                    t = torch.empty(shape, device=device, dtype=torch.bfloat16)
                dist.broadcast(t, 0)
                yield k, t

        model.load_state_dict(dit_items())
        dsd = make_dac_weights(dcfg, 4321) if rank == 0 else None

        def dac_items():
            for k, shape, _ in dac_param_specs(dcfg):
                t = dsd[k].to(device) if rank == 0 else torch.empty(shape, device=device, dtype=torch.float32)
                dist.broadcast(t, 0)
                yield k, t

        dac.load_state_dict(dac_items())
    comps, mean, scale = make_pca_state(dcfg)
    return model, dac, PCAState(comps.to(device), mean.to(device), scale)


LONG_PROMPT_SENTENCES = (
    "The lighthouse keeper counted the ships that passed before dawn, and wrote each name in a ledger bound in green cloth.",
    "When the fog came in from the north, nobody in the village could tell the harbour wall from the open sea.",
    "She had promised her brother that the lamp would never go out, not for storms, not for sickness, not for grief.",
    "Every winter the gulls grew bolder, stealing bread from the windowsill and shrieking at the cat until it fled.",
    "A letter arrived in spring, stamped in a city she had never seen, asking whether the light was still for sale.",
)


def long_prompt(n_chars=3000):
    """Deterministic ~3000-character prompt for BASELINE configs[3]."""
    out, i = [], 0
    while sum(len(x) + 1 for x in out) < n_chars:
        out.append(LONG_PROMPT_SENTENCES[i % len(LONG_PROMPT_SENTENCES)])
        i += 1
    return " ".join(out)


def run_extras(args, model, dac, pca, device, rank, world, barrier):
    """config.extras: BASELINE configs[2], [3] (any N) and [4], batch 4 (N = 1), each with its own warm-up and timed on
    the device (CUDA events, max over ranks) or -- for the chunked job latency -- by the wall clock of rank 0 between
    two barriers. Every rank returns the same dict (rank 0 prints it)."""
    import functools

    import torch.distributed as dist
    from echo_tts_b200 import pipeline as P
    from echo_tts_b200.autoencoder import ae_decode
    from echo_tts_b200.sampler import sample_blockwise_euler_cfg_independent_guidances as blockwise
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    out = {}
    ids_h, mask_h = tokens(PROMPT)
    spk1 = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1)).to(device)
    smask1 = torch.ones(1, 212, dtype=torch.bool, device=device)

    # ---- configs[2]: 32 independent requests, request-sharded 32 / N per GPU, 4 requests per sampler call
    n_req = 32
    per_call = 8 if n_req // world >= 8 else 4  # 8 per call: +2 % over 4 (profiles/r01_bench_v22_batch8.json)
    mine = list(range(rank, n_req, world))
    calls = [mine[i:i + per_call] for i in range(0, len(mine), per_call)]

    def run_calls():
        for c in calls:
            b = len(c)
            nz = torch.randn(b, 640, 80, device=device, generator=torch.Generator(device).manual_seed(c[0]))
            lat = sample(model, spk1.repeat(b, 1, 1), smask1.repeat(b, 1), ids_h.repeat(b, 1).to(device),
                         mask_to(mask_h.repeat(b, 1), device), 0, sequence_length=640, noise=nz, **KNOBS)
            ae_decode(dac, pca, lat)

    run_calls()  # warm-up (workspaces grow to the batch-4 sizes)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_calls()
    e1.record()
    barrier()
    ms3 = max_over_ranks(e0.elapsed_time(e1))
    out["cfg3"] = {"workload": f"configs[2]: 32 independent requests (each = configs[1]), sharded 32/N per GPU, {per_call} per sampler call",
                   "audio_s_per_s": n_req * AUDIO_SECONDS / (ms3 / 1e3), "job_ms": ms3, "requests": n_req,
                   "requests_per_gpu": len(mine), "requests_per_call": per_call}

    # ---- configs[3]: one ~3000-char prompt -> ~300-char chunks -> chunk i on rank i % N -> gather -> host stitch
    text = long_prompt()
    sample_fn = functools.partial(sample, **KNOBS, sequence_length=640)

    def synth_chunk(chunk, seed):
        audio, _ = P.sample_pipeline(model, dac, pca, sample_fn, chunk, None, seed, speaker_latent=spk1, speaker_mask=smask1)
        return audio[0]

    # target_duration 25 s: chunk_text_for_audio caps chunks at 12 chars/s * target (handler.py:114), so ~300-char
    # chunks (BASELINE configs[3]) need 25 s; the handler's 10 s default would give ~120-char chunks
    def synth_chunks(chunk_list, seeds):  # a rank's chunks batched into one sampler call + one decode (chunks of one
        res = P.sample_pipeline_batch(model, dac, pca, sample_fn, chunk_list, seeds, speaker_latent=spk1,  # prompt share the voice)
                                      speaker_mask=smask1)
        return [a[0] for a, _ in res]

    job = functools.partial(P.synthesize, text, synth_chunk, seed=0, max_chars_per_chunk=300, target_duration=25.0,
                            synth_chunks=synth_chunks, chunks_per_call=4)
    n_chunks = len(P.chunk_text_for_audio(text, 300, 25.0))
    job()  # warm-up
    lat_s = []
    stitched = None
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        stitched = job()
        dt = time.perf_counter() - t0  # rank 0: includes the gather and the host stitching
        barrier()
        lat_s.append(dt)
    job_s = min(lat_s) if rank == 0 else 0.0
    if world > 1:
        t = torch.tensor([job_s], device=device, dtype=torch.float64)
        dist.broadcast(t, 0)
        job_s = t.item()
    audio_s = (stitched.shape[-1] / 44100.0) if stitched is not None else 0.0
    if world > 1:
        t = torch.tensor([audio_s], device=device, dtype=torch.float64)
        dist.broadcast(t, 0)
        audio_s = t.item()
    out["cfg4"] = {"workload": f"configs[3]: one {len(text)}-char prompt, chunk_text_for_audio(300, target 25 s) -> {n_chunks} chunks, "
                               f"chunk i on GPU i % N (a GPU's chunks batched 4 per sampler call), NCCL all_gather of the audio, boundary "
                               f"normalisation + crossfade on rank 0's host",
                   "job_latency_ms": job_s * 1e3, "chunks": n_chunks, "critical_path_chunks": -(-n_chunks // world),
                   "stitched_audio_seconds": audio_s, "audio_s_per_s": audio_s / job_s if job_s > 0 else None,
                   "timing": "wall clock on rank 0 between two barriers (min of 2 jobs after 1 warm-up job)"}

    if world == 1:
        # ---- 4 requests per call (one GPU)
        def batch4():
            nz = torch.randn(4, 640, 80, device=device, generator=torch.Generator(device).manual_seed(7))
            lat = sample(model, spk1.repeat(4, 1, 1), smask1.repeat(4, 1), ids_h.repeat(4, 1).to(device),
                         mask_to(mask_h.repeat(4, 1), device), 0, sequence_length=640, noise=nz, **KNOBS)
            return ae_decode(dac, pca, lat)

        batch4()
        torch.cuda.synchronize(device)
        e0.record()
        for _ in range(3):
            batch4()
        e1.record()
        torch.cuda.synchronize(device)
        out["batch4_audio_s_per_s"] = 3 * 4 * AUDIO_SECONDS / (e0.elapsed_time(e1) / 1e3)

        # ---- configs[4]: blockwise 4 x 160, 5-minute speaker reference (1600 KV patches), speaker_kv_scale 1.5 until
        # t < 0.9, streaming DAC decode; device timeline of when each block's audio is complete
        spk5 = torch.randn(1, 6400, 80, generator=torch.Generator().manual_seed(1)).to(device)
        smask5 = torch.ones(1, 6400, dtype=torch.bool, device=device)
        knobs5 = dict(KNOBS, speaker_kv_scale=1.5, speaker_kv_min_t=0.9, speaker_kv_max_layers=24)
        ids, mask = ids_h.to(device), mask_to(mask_h, device)
        runs = []
        for it in range(4):
            torch.cuda.synchronize(device)
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            _, parts = P.stream_blockwise_audio(model, dac, pca, blockwise, spk5, smask5, ids, mask, it, [160] * 4, **knobs5)
            torch.cuda.synchronize(device)
            runs.append([ev0.elapsed_time(ev) for _, ev in parts])
        runs = runs[1:]  # the first run sizes the workspaces
        first = sorted(r[0] for r in runs)[len(runs) // 2]
        total = sorted(r[-1] for r in runs)[len(runs) // 2]
        out["cfg5"] = {"workload": "configs[4]: sample_blockwise 4 x 160 latents, 40 steps per block, 6400-latent (5 min) speaker "
                                   "reference = 1600 speaker-KV patches, speaker_kv_scale 1.5 until t < 0.9, streaming DAC decode",
                       "first_audio_ms": first, "total_ms": total, "block_ready_ms": runs[-1],
                       "audio_s_per_s": AUDIO_SECONDS / (total / 1e3),
                       "timing": "CUDA events on the stream (median of 3 runs after 1 warm-up); first audio includes the "
                                 "text + 1600-patch speaker KV caches"}
    return out


def run_b200(args):
    import torch.distributed as dist
    from echo_tts_b200 import _lib
    from echo_tts_b200.autoencoder import ae_decode
    from echo_tts_b200.sampler import sample_euler_cfg_independent_guidances as sample

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    t_setup = time.time()
    model, dac, pca = load_models(device, rank, world)
    setup_s = time.time() - t_setup

    Bq = max(1, args.batch)
    ids_h, mask_h = tokens(PROMPT)
    ids_h, mask_h = ids_h.repeat(Bq, 1), mask_h.repeat(Bq, 1)
    spk_h = torch.randn(1, 212, 80, generator=torch.Generator().manual_seed(1)).repeat(Bq, 1, 1)
    smask_h = torch.ones(Bq, 212, dtype=torch.bool)
    n_req = args.warmup + args.steps + 1
    noise_h = torch.randn(n_req, Bq, 640, 80, generator=torch.Generator().manual_seed(1000 + rank))
    ids, spk, smask, noise = (t.to(device) for t in (ids_h, spk_h, smask_h, noise_h))
    mask = mask_to(mask_h, device)

    def request(i):
        lat = sample(model, spk, smask, ids, mask, 0, sequence_length=640, noise=noise[i], **KNOBS)
        return ae_decode(dac, pca, lat)

    # ---- value: inputs resident in HBM
    for i in range(args.warmup):
        request(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = model.h.num_launches()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        audio = request(args.warmup + i)
        evs[i + 1].record()
    barrier()
    sampler.stop_flag = True
    ms = evs[0].elapsed_time(evs[-1])
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    p50_ms = per_step[len(per_step) // 2]
    launches = model.h.num_launches() - l0
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * args.steps * Bq * AUDIO_SECONDS / (ms / 1e3)
    assert torch.isfinite(audio).all() and tuple(audio.shape) == (Bq, 1, 640 * 2048)

    # ---- e2e: the public drop-in call (pipeline.sample_pipeline == reference inference.sample_pipeline) with HOST inputs:
    # prompt string -> tokens on the host -> H2D, pinned host speaker latents + mask -> H2D, noise drawn on the device
    # from the seed exactly as the reference draws it (inference.py:457,477), sampler, DAC decode, flattening-point
    # crop on the device (one int32 D2H), cropped audio -> pinned host memory (D2H). Every step.
    import functools
    from echo_tts_b200 import pipeline as P
    sample_fn = functools.partial(sample, **KNOBS, sequence_length=640)  # as handler._build_sample_fn (handler.py:426-443)
    h_spk, h_smask = spk_h[:1].pin_memory(), smask_h[:1].pin_memory()
    h_audio = torch.empty(1, 1, 640 * 2048, dtype=torch.float32).pin_memory()
    h2d = h_spk.numel() * 4 + h_smask.numel() + 768 * 4 + 768  # speaker latents fp32 + mask, token ids int32 + mask
    d2h_box = [0]

    def request_e2e(i):
        audio, _ = P.sample_pipeline(model, dac, pca, sample_fn, PROMPT, None, 1000 * rank + i,
                                     speaker_latent=h_spk, speaker_mask=h_smask)
        n = audio.shape[-1]
        h_audio[..., :n].copy_(audio, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        d2h_box[0] = n * 4 + 4  # cropped audio + the flattening index
        return n

    e2e = e2e_s = None
    d2h = 0
    if Bq == 1:
        request_e2e(0)
        barrier()
        t0 = time.perf_counter()
        samples = 0
        for i in range(args.steps):
            samples += request_e2e(args.warmup + i)
        barrier()
        e2e_s = time.perf_counter() - t0
        d2h = d2h_box[0]
        if world > 1:
            t = torch.tensor([e2e_s], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = t.item()
        # audio-seconds = what the call returned AFTER the crop (random-init latents never flatten, so normally all
        # 1 310 720 samples of every request: then this is the same unit as `value`)
        e2e = world * (samples / 44100.0) / e2e_s

    # ---- extras: the other BASELINE configs, measured in the same run on the same model (never part of `value`)
    extras = {}
    if not args.no_extras and Bq == 1:
        extras = run_extras(args, model, dac, pca, device, rank, world, barrier)

    # ---- roofline: one more request with per-launch CUDA events (same kernels, same stream)
    rep = _lib.ProfileReport()
    import ctypes as C
    _lib.check(model.lib.echo_profile_start(model.h.ptr), "echo_profile_start")
    request(n_req - 1)
    _lib.check(model.lib.echo_profile_stop(model.h.ptr, C.byref(rep)), "echo_profile_stop")
    peak_tf, peak_hbm, peak_kind = measured_peaks()
    gemm_tf = rep.flops[0] / (rep.ms[0] * 1e-3) / 1e12 if rep.ms[0] > 0 else 0.0
    prof_total = sum(rep.ms)
    fl = request_flops(n_text_valid=len(PROMPT.encode()) + 1, n_spk_patches=53)
    traffic = ncu_traffic()
    w13 = traffic.get("gemm_w13_M1920")  # largest single share of a request: SwiGLU up-projection on CFG steps
    traffic_note = None
    if w13:
        traffic_note = {"kernel": "gemm_tc_kernel<256,64,1,EPI_SWIGLU,pair> M=1920 N=11776 K=2048 (29 % of the GEMM time)",
                        "dram_bytes_per_launch": w13["dram_bytes_read"] + w13["dram_bytes_write"],
                        "algorithmic_bytes_per_launch": 2 * (1920 * 2048 + 11776 * 2048 + 1920 * 5888),
                        "tensor_pipe_active_pct_ncu": w13["tensor_pipe_active_pct"],
                        "source": "profiles/r02_ncu_traffic.json (ncu --set full, cold cache)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from echo_tts_b200.config import DacConfig, DitConfig
        from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state
        threads = os.cpu_count() or 1
        sd = make_dit_weights(DitConfig.base(), seed=1234, include_latent=False)
        dsd = make_dac_weights(DacConfig.base(), seed=4321)
        prepare, step = oracle_sample_seconds(sd, dsd, make_pca_state(DacConfig.base()), threads)
        prepare()
        step(6)  # warm-up (thread pool, allocator)
        est, _ = step(7)
        cpu = {"value": AUDIO_SECONDS / est, "unit": "audio-s/s", "cores": threads, "kind": "port", "extrapolated": True,
               "sample": "one fp32 CFG forward (3 x 640 rows) + one plain forward + KV caches + DAC decode of 16 latents "
                         "with the oracle, extrapolated to a request (20 x each forward + 40 x decode(16))"}

    if rank == 0:
        line = {
            "metric": "audio-sec/sec (RTF^-1) per GPU at seq 640/40 steps", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"replicas x{world} (requests sharded, no collective per step)",
                       "l2": "per step the kernels stream the 2.7 GB of bf16 weights of the 24 DiT blocks (>> 126 MB L2): inputs larger than L2",
                       "p50_latency_ms": p50_ms, "requests_per_call": Bq, "extras": extras,
                       "reproducibility": "default mode uses atomic split-K in the M=640 residual GEMMs (run-to-run "
                                          "differences at the bf16 noise floor); echo_set_deterministic(1) is bit-exact and ~1 % slower"},
            "e2e": ({"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "api": "echo_tts_b200.pipeline.sample_pipeline (== reference inference.sample_pipeline): host prompt + "
                            "pinned host speaker latents in, sampler + DAC decode + flattening-point crop on the GPU, "
                            "cropped audio to pinned host memory"} if e2e is not None else None),
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": gemm_tf / peak_tf,
                         "traffic": (traffic_note["dram_bytes_per_launch"] if traffic_note else None),
                         "traffic_detail": traffic_note, "peak_source": f"{peak_kind} bf16_tflops_sustained",
                         "kernel": "gemm_tc_kernel (tcgen05 GEMM, all DiT/encoder/DAC contractions of one request)",
                         "gemm_launches": int(rep.launches[0]), "gemm_ms": rep.ms[0], "gemm_flops": rep.flops[0],
                         "attention_ms": rep.ms[1], "glue_ms": rep.ms[2],
                         "gemm_share_of_step": rep.ms[0] / prof_total if prof_total else None,
                         "note": "achieved = algorithmic FLOPs / summed per-launch CUDA-event time of every GEMM launch of one "
                                 "request (events serialise the chain: no PDL overlap, ~2 us per launch included)",
                         "request_algorithmic_tflop": Bq * fl["total"] / 1e12,
                         "request_frac_of_peak": Bq * fl["total"] / (ms / args.steps * 1e-3) / 1e12 / peak_tf},
            "cpu_baseline": cpu,
            "setup_seconds": setup_s,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = 1


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


def main():
    # Libraries print to stdout behind our back (NCCL: "NCCL version 2.28.9+cuda12.9" on the 8-GPU box). Keep the
    # original stdout for the JSON line only and send everything else to stderr.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip config.extras (cfg3 / cfg4 / cfg5 / batch4)")
    ap.add_argument("--batch", type=int, default=1,
                    help="independent requests per sampler call on every GPU (1 = BASELINE configs[1]; > 1 = configs[2] style)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
