"""`B200DAC`: drop-in for the reference Fish S1-DAC object on the DECODE path, plus `PCAState` / `ae_decode`.

Mirrors what the reference callers touch (reference inference.py:86-99, 226-229; autoencoder.py:1128-1138):
`fish_ae.decode_zq(z (B, 1024, T)) -> (B, 1, 2048 T)`, `.dtype`, `.device`, and `ae_decode(fish_ae, pca_state, z_q)`.
`encode_zq` (speaker-reference encoding) is out of scope for this round and raises.
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
        _lib.check(self.lib.echo_dac_configure(self.h.ptr, C.byref(c)), "echo_dac_configure")

    def load_state_dict(self, state: Iterable[Tuple[str, torch.Tensor]] | dict, strict: bool = False, assign: bool = False):
        """Accepts a full reference DAC state dict; only decode-path tensors are consumed."""
        items = state.items() if isinstance(state, dict) else state
        for k, v in items:
            if k.startswith("decoder.") or k.startswith("quantizer.upsample.") or (
                    k.startswith("quantizer.post_module.") and not k.endswith(("freqs_cis", "causal_mask"))):
                self.h.set_weight("dac." + k, v)
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

    def encode_zq(self, audio_data: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("DAC encode / RVQ is outside this round's scope (SURVEY.md 8(f)); "
                                  "run the reference encoder to obtain speaker latents")

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


@torch.inference_mode()
def ae_decode(fish_ae, pca_state: PCAState, z_q: torch.Tensor) -> torch.Tensor:
    """reference inference.ae_decode (inference.py:226-229). With a B200DAC the PCA un-projection is fused."""
    if isinstance(fish_ae, B200DAC):
        return fish_ae.decode_latent(pca_state, z_q)
    z = (z_q / pca_state.latent_scale) @ pca_state.pca_components + pca_state.pca_mean
    return fish_ae.decode_zq(z.transpose(1, 2).to(fish_ae.dtype)).float()
