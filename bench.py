#!/usr/bin/env python
"""bench.py -- Echo-TTS sampling hot path on B200: audio-seconds per second (RTF^-1).

One "step" = one full request of BASELINE.json configs[1]: echo-tts-base architecture (random-init), one short [S1]
prompt padded to 768 tokens (as inference.sample_pipeline does), a 10 s synthetic speaker reference (212 latents),
cfg_scale_text 3.0 / cfg_scale_speaker 8.0 / cfg_min_t 0.5, 40 Euler steps at sequence_length 640, batch 1, followed
by the Fish S1-DAC decode of the 640 latents to 1 310 720 samples @ 44.1 kHz = 29.72 s of audio.

  value  : audio-s/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the public Python API with pinned HOST inputs (H2D) and the audio copied back (D2H)
  N > 1  : every rank is a full replica processing its own requests (no data-path collective): weak scaling
  --impl reference : the reference algorithm (oracle port, fp32 PyTorch on the host cores) on a bounded sample

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
    p = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
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
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
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

    def skip(k):
        return k.startswith("latent_encoder.") or k.startswith("latent_norm") or ".wk_latent" in k or ".wv_latent" in k

    if world == 1:
        model.load_state_dict(iter_dit_weights(cfg, 1234, include_latent=False))
        dac.load_state_dict(make_dac_weights(dcfg, 4321))
    else:
        specs = [(k, s) for k, s, _ in dit_param_specs(cfg) if not skip(k)]
        gen = iter_dit_weights(cfg, 1234, include_latent=False) if rank == 0 else None

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
    ids, mask, spk, smask, noise = (t.to(device) for t in (ids_h, mask_h, spk_h, smask_h, noise_h))

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

    # ---- e2e: pinned host inputs -> H2D, public API, D2H of the audio, every step
    pin = lambda t: t.pin_memory()
    h_in = [pin(ids_h), pin(mask_h), pin(spk_h), pin(smask_h)]
    h_noise = pin(noise_h)
    h_audio = torch.empty(Bq, 1, 640 * 2048, dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in h_in) + noise_h[0].numel() * 4
    d2h = h_audio.numel() * 4

    def request_e2e(i):
        d = [t.to(device, non_blocking=True) for t in h_in]
        nz = h_noise[i].to(device, non_blocking=True)
        lat = sample(model, d[2], d[3], d[0], d[1], 0, sequence_length=640, noise=nz, **KNOBS)
        h_audio.copy_(ae_decode(dac, pca, lat), non_blocking=True)
        torch.cuda.current_stream(device).synchronize()

    request_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        request_e2e(args.warmup + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e = world * args.steps * Bq * AUDIO_SECONDS / e2e_s

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
                        "source": "profiles/r01_ncu_traffic.json (ncu --set full, cold cache)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from echo_tts_b200.config import DacConfig, DitConfig
        from echo_tts_b200.weights import make_dac_weights, make_dit_weights, make_pca_state
        threads = os.cpu_count() or 1
        sd = make_dit_weights(DitConfig.base(), seed=1234, include_latent=False)
        dsd = make_dac_weights(DacConfig.base(), seed=4321)
        prepare, step = oracle_sample_seconds(sd, dsd, make_pca_state(DacConfig.base()), threads)
        prepare()
        est, _ = step(7)
        cpu = {"value": AUDIO_SECONDS / est, "unit": "audio-s/s", "cores": threads, "kind": "port",
               "sample": "one fp32 CFG forward (3 x 640 rows) + one plain forward + KV caches + DAC decode of 16 latents "
                         "with the oracle, extrapolated to a request (20 x each forward + 40 x decode(16))"}

    if rank == 0:
        line = {
            "metric": "audio-sec/sec (RTF^-1) per GPU at seq 640/40 steps", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"replicas x{world} (requests sharded, no collective per step)",
                       "l2": "per step the kernels stream the 2.7 GB of bf16 weights of the 24 DiT blocks (>> 126 MB L2): inputs larger than L2",
                       "p50_latency_ms": p50_ms, "requests_per_call": Bq,
                       "reproducibility": "default mode uses atomic split-K in the M=640 residual GEMMs (run-to-run "
                                          "differences at the bf16 noise floor); echo_set_deterministic(1) is bit-exact and ~1 % slower"},
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
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
                         # the per-launch events serialise the chain (no PDL overlap) and add ~2 us per launch; the
                         # same FLOPs over the kernel's SHARE of the un-instrumented step time:
                         "achieved_in_uninstrumented_step": (rep.flops[0] / (ms / args.steps * 1e-3 * rep.ms[0] / prof_total) / 1e12
                                                             if prof_total and rep.ms[0] > 0 else None),
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
