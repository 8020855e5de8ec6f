"""`B200DAC`: drop-in for the reference Fish S1-DAC object, plus `PCAState` / `ae_decode` / `ae_encode`.

Mirrors what the reference callers touch (reference inference.py:86-99, 219-229; autoencoder.py:1080-1138):
`fish_ae.decode_zq(z (B, 1024, T)) -> (B, 1, 2048 T)`, `fish_ae.encode_zq(audio (B, 1, L)) -> (B, 1024, ceil(L / 2048))`,
`.dtype`, `.device`, `ae_decode(fish_ae, pca_state, z_q)` and `ae_encode(fish_ae, pca_state, audio)`.
The encode path needs the encoder / quantizer tensors of the checkpoint; without them `encode_zq` raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Tuple

import torch

from . import _lib
from .config import DacConfig
from .model import Handle, _stream


@dataclass
class PCAState:  # same fields as reference inference.py:86-90
    pca_components: torch.Tensor
    pca_mean: torch.Tensor
    latent_scale: float


class B200DAC:
    def __init__(self, cfg: DacConfig = DacConfig.base(), device="cuda"):
        self.cfg = cfg
        self.h = Handle(torch.device(device))
        self.lib = self.h.lib
        c = _lib.DacConfig()
        c.latent_dim, c.pca_dim, c.post_layers, c.post_heads = cfg.latent_dim, cfg.pca_dim, cfg.post_layers, cfg.post_heads
        c.post_intermediate, c.post_window, c.post_norm_eps = cfg.post_intermediate, cfg.post_window, cfg.post_norm_eps
        c.num_upsample, c.decoder_dim, c.num_rates = cfg.num_upsample, cfg.decoder_dim, len(cfg.rates)
        for i, r in enumerate(cfg.rates):
            c.rates[i] = r
        c.enc_dim, c.num_enc_rates, c.enc_t_layers, c.enc_window = cfg.enc_dim, len(cfg.enc_rates), cfg.enc_t_layers, cfg.enc_window
        for i, r in enumerate(cfg.enc_rates):
            c.enc_rates[i] = r
        c.n_codebooks, c.codebook_size = cfg.n_codebooks, cfg.codebook_size
        c.semantic_codebook_size, c.codebook_dim = cfg.semantic_codebook_size, cfg.codebook_dim
        self.has_encoder = False
        _lib.check(self.lib.echo_dac_configure(self.h.ptr, C.byref(c)), "echo_dac_configure")

    def load_state_dict(self, state: Iterable[Tuple[str, torch.Tensor]] | dict, strict: bool = False, assign: bool = False):
        """Accepts a full reference DAC state dict. The decode path is mandatory; the encode path (encoder.*,
        quantizer.downsample / pre_module / *quantizer.quantizers) is used when present."""
        items = state.items() if isinstance(state, dict) else state
        for k, v in items:
            if k.endswith(("freqs_cis", "causal_mask")):
                continue  # derived buffers
            if k.startswith(("decoder.", "quantizer.upsample.", "quantizer.post_module.")):
                self.h.set_weight("dac." + k, v)
            elif k.startswith(("encoder.", "quantizer.downsample.", "quantizer.pre_module.",
                               "quantizer.semantic_quantizer.", "quantizer.quantizer.")):
                self.h.set_weight("dac." + k, v)
                self.has_encoder = True
        with torch.cuda.device(self.device):
            _lib.check(self.lib.echo_dac_finalize(self.h.ptr, _stream(self.device)), "echo_dac_finalize")
        return self

    @classmethod
    def from_state_dict(cls, state, cfg: DacConfig = DacConfig.base(), device="cuda") -> "B200DAC":
        return cls(cfg, device).load_state_dict(state)

    @property
    def device(self) -> torch.device:
        return self.h.device

    @property
    def dtype(self) -> torch.dtype:
        # interface dtype: ae_decode casts the PCA un-projection to fish_ae.dtype (inference.py:229); keeping it
        # float32 hands the kernels the un-rounded latent. Internally GEMM operands are bf16 with fp32 accumulation.
        return torch.float32

    def eval(self):
        return self

    @torch.inference_mode()
    def decode_zq(self, z_q: torch.Tensor) -> torch.Tensor:
        z = z_q.to(self.device, torch.float32).contiguous()
        B, Cc, T = z.shape
        assert Cc == self.cfg.latent_dim
        audio = torch.empty(B, 1, T * self.cfg.hop, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.echo_dac_decode_zq(self.h.ptr, z.data_ptr(), B, T, audio.data_ptr(), _stream(self.device)),
                       "echo_dac_decode_zq")
        return audio

    def _padded_audio(self, audio_data: torch.Tensor) -> torch.Tensor:
        """(B, L) or (B, 1, L) -> (B, 1, L') fp32 on the device, right-padded with zeros to a multiple of the frame
        length exactly as DAC.encode does (autoencoder.py:1090-1095)."""
        a = audio_data
        if a.dim() == 2:
            a = a.unsqueeze(1)
        assert a.dim() == 3 and a.shape[1] == 1, "audio must be (B, 1, L)"
        a = a.to(self.device, torch.float32)
        fl = self.cfg.frame_length
        pad = (-a.shape[-1]) % fl
        if pad:
            a = torch.nn.functional.pad(a, (0, pad))
        return a.contiguous()

    @torch.inference_mode()
    def encode_zq(self, audio_data: torch.Tensor, return_codes: bool = False, return_z_pre: bool = False):
        """reference DAC.encode_zq (autoencoder.py:1116-1126): (B, 1, L) -> z_q (B, latent_dim, ceil(L / 2048)) fp32.
        With return_codes also the (B, 1 + n_codebooks, T) code indices of DAC.encode."""
        if not self.has_encoder:
            raise _lib.EchoError("B200DAC.encode_zq: the checkpoint had no encoder / quantizer tensors")
        a = self._padded_audio(audio_data)
        B, _, L = a.shape
        T = L // self.cfg.frame_length
        zq = torch.empty(B, self.cfg.latent_dim, T, device=self.device, dtype=torch.float32)
        codes = torch.empty(B, 1 + self.cfg.n_codebooks, T, device=self.device, dtype=torch.int32) if return_codes else None
        z_pre = torch.empty(B, T, self.cfg.latent_dim, device=self.device, dtype=torch.float32) if return_z_pre else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.echo_dac_encode_zq(self.h.ptr, a.data_ptr(), B, L, zq.data_ptr(),
                                                   None if codes is None else codes.data_ptr(),
                                                   None if z_pre is None else z_pre.data_ptr(), _stream(self.device)),
                       "echo_dac_encode_zq")
        if not (return_codes or return_z_pre):
            return zq
        return (zq,) + ((codes.long(),) if return_codes else ()) + ((z_pre,) if return_z_pre else ())

    @torch.inference_mode()
    def encode_latent(self, pca_state: PCAState, audio: torch.Tensor) -> torch.Tensor:
        """Fused encode + PCA projection: audio (B, 1, L) -> (B, ceil(L / 2048), 80) fp32 (reference ae_encode)."""
        if not self.has_encoder:
            raise _lib.EchoError("B200DAC.encode_latent: the checkpoint had no encoder / quantizer tensors")
        dev = self.device
        a = self._padded_audio(audio)
        B, _, L = a.shape
        T = L // self.cfg.frame_length
        comps = pca_state.pca_components.to(dev, torch.float32).contiguous()
        mean = pca_state.pca_mean.to(dev, torch.float32).contiguous()
        K = comps.shape[0]
        assert comps.shape == (K, self.cfg.latent_dim) and K == self.cfg.pca_dim
        out = torch.empty(B, T, K, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(self.lib.echo_dac_encode(self.h.ptr, a.data_ptr(), comps.data_ptr(), mean.data_ptr(),
                                                float(pca_state.latent_scale), B, L, out.data_ptr(), _stream(dev)),
                       "echo_dac_encode")
        return out

    def new_stream(self, max_latents: int = 640) -> "DacStream":
        """A streaming-decode state for one sequence of up to `max_latents` latents (see `DacStream`)."""
        return DacStream(self, max_latents)

    def borrow_stream(self, max_latents: int = 640) -> "DacStream":
        """A reset stream from this decoder's pool (or a new one): a server that streams request after request does not
        pay the ~26 device allocations of a stream state per request. Give it back with `return_stream`."""
        pool = self.__dict__.setdefault("_stream_pool", [])
        for i, st in enumerate(pool):
            if st.max_latents >= max_latents:
                pool.pop(i)
                st.reset()
                return st
        return DacStream(self, max_latents)

    def return_stream(self, st: "DacStream") -> None:
        pool = self.__dict__.setdefault("_stream_pool", [])
        if len(pool) < 16:
            pool.append(st)
        else:
            st.close()

    @torch.inference_mode()
    def decode_latent(self, pca_state: PCAState, z: torch.Tensor) -> torch.Tensor:
        """Fused PCA un-projection + decode: z (B, T, 80) fp32 -> (B, 1, 2048 T) fp32."""
        dev = self.device
        zf = z.to(dev, torch.float32).contiguous()
        B, T, K = zf.shape
        comps = pca_state.pca_components.to(dev, torch.float32).contiguous()
        mean = pca_state.pca_mean.to(dev, torch.float32).contiguous()
        assert comps.shape == (K, self.cfg.latent_dim)
        audio = torch.empty(B, 1, T * self.cfg.hop, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(self.lib.echo_dac_decode(self.h.ptr, zf.data_ptr(), comps.data_ptr(), mean.data_ptr(),
                                                float(pca_state.latent_scale), B, T, audio.data_ptr(), _stream(dev)),
                       "echo_dac_decode")
        return audio


class DacStream:
    """Stateful streaming decode: feed the latents of ONE sequence block by block, get its audio block by
    block -- each block costs what its own latents cost, and the samples are bit-identical to an offline
    `ae_decode` of the whole sequence. The state (the window-128 keys / values of the post_module, the input halos of
    every causal conv) lives in the library (`echo_dac_stream_*`). One sequence per stream; `reset()` starts a new one."""

    def __init__(self, dac: "B200DAC", max_latents: int = 640):
        self.dac, self.max_latents = dac, int(max_latents)
        self.ptr = C.c_void_p()
        with torch.cuda.device(dac.device):
            _lib.check(dac.lib.echo_dac_stream_create(dac.h.ptr, self.max_latents, C.byref(self.ptr), _stream(dac.device)),
                       "echo_dac_stream_create")

    @property
    def position(self) -> int:
        n = C.c_int()
        _lib.check(self.dac.lib.echo_dac_stream_position(self.dac.h.ptr, self.ptr, C.byref(n)), "echo_dac_stream_position")
        return n.value

    def reset(self) -> None:
        with torch.cuda.device(self.dac.device):
            _lib.check(self.dac.lib.echo_dac_stream_reset(self.dac.h.ptr, self.ptr, _stream(self.dac.device)),
                       "echo_dac_stream_reset")

    @torch.inference_mode()
    def decode(self, pca_state: PCAState, z: torch.Tensor) -> torch.Tensor:
        """z (1, T, 80) or (T, 80): the NEXT T latents of the sequence -> (1, 1, 2048 T) fp32, the next samples."""
        dac, dev = self.dac, self.dac.device
        zf = z.to(dev, torch.float32).contiguous()
        if zf.dim() == 2:
            zf = zf.unsqueeze(0)
        assert zf.dim() == 3 and zf.shape[0] == 1, "one sequence per stream"
        T, K = zf.shape[1], zf.shape[2]
        comps = pca_state.pca_components.to(dev, torch.float32).contiguous()
        mean = pca_state.pca_mean.to(dev, torch.float32).contiguous()
        assert comps.shape == (K, dac.cfg.latent_dim)
        audio = torch.empty(1, 1, T * dac.cfg.hop, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(dac.lib.echo_dac_stream_decode(dac.h.ptr, self.ptr, zf.data_ptr(), comps.data_ptr(), mean.data_ptr(),
                                                      float(pca_state.latent_scale), T, audio.data_ptr(), _stream(dev)),
                       "echo_dac_stream_decode")
        return audio

    def close(self) -> None:
        if getattr(self, "ptr", None) and self.ptr.value and getattr(self.dac.h, "ptr", None) and self.dac.h.ptr.value:
            self.dac.lib.echo_dac_stream_destroy(self.dac.h.ptr, self.ptr)
        self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@torch.inference_mode()
def ae_decode(fish_ae, pca_state: PCAState, z_q: torch.Tensor) -> torch.Tensor:
    """reference inference.ae_decode (inference.py:226-229); the PCA un-projection is fused into the decode kernels.
    `fish_ae` must be a `B200DAC`: there is no second backend behind this API (no eager / CPU path)."""
    if not isinstance(fish_ae, B200DAC):
        raise TypeError(f"ae_decode: fish_ae must be a B200DAC (got {type(fish_ae).__name__}); echo_tts_b200 has no "
                        "eager fallback -- build one with B200DAC.from_state_dict(reference_dac.state_dict())")
    return fish_ae.decode_latent(pca_state, z_q)


@torch.inference_mode()
def ae_encode(fish_ae, pca_state: PCAState, audio: torch.Tensor) -> torch.Tensor:
    """reference inference.ae_encode (inference.py:219-224): (B, 1, L) -> (B, T, 80), PCA projection fused.
    `fish_ae` must be a `B200DAC` (no eager fallback)."""
    assert audio.ndim == 3 and audio.shape[1] == 1  # (b, 1, length)
    if not isinstance(fish_ae, B200DAC):
        raise TypeError(f"ae_encode: fish_ae must be a B200DAC (got {type(fish_ae).__name__}); echo_tts_b200 has no "
                        "eager fallback")
    return fish_ae.encode_latent(pca_state, audio)
