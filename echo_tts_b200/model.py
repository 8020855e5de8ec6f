"""`B200EchoDiT`: drop-in for the reference `EchoDiT` object on the sampling path.

Mirrors the members the reference callers touch (reference model.py:563-642; inference.py:325,454,464-465,487-504;
inference_blockwise.py:49-50,73,86-107): `model(x=, t=, text_mask=, speaker_mask=, kv_cache_text=,
kv_cache_speaker=[, start_pos=, kv_cache_latent=])`, `get_kv_cache_text/speaker/latent`, `.device`, `.dtype`,
`.eval()`. All arithmetic happens in libecho_b200.so (hand-written sm_100a kernels); PyTorch only owns the tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .config import DitConfig

KVCache = List[Tuple[torch.Tensor, torch.Tensor]]


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _u8(mask: torch.Tensor) -> torch.Tensor:
    m = mask.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class Handle:
    """Owns one echo_handle (one CUDA device)."""

    def __init__(self, device: torch.device):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.EchoError("echo_tts_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.ptr = C.c_void_p()
        _lib.check(self.lib.echo_create(C.byref(self.ptr), idx), "echo_create")

    def set_weight(self, key: str, tensor: torch.Tensor) -> None:
        t = tensor.detach()
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        t = t.to(self.device, non_blocking=False).contiguous()
        shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
        dt = _lib.DTYPE_BF16 if t.dtype == torch.bfloat16 else _lib.DTYPE_F32
        with torch.cuda.device(self.device):
            _lib.check(self.lib.echo_set_weight(self.ptr, key.encode(), t.data_ptr(), shape, max(t.dim(), 1), dt,
                                                _stream(self.device)), f"echo_set_weight({key})")
            torch.cuda.current_stream(self.device).synchronize()  # `t` may be freed right after

    def num_launches(self) -> int:
        n = C.c_int64()
        _lib.check(self.lib.echo_num_launches(self.ptr, C.byref(n)), "echo_num_launches")
        return n.value

    def close(self):
        if getattr(self, "ptr", None) and self.ptr.value:
            self.lib.echo_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200EchoDiT:
    def __init__(self, cfg: DitConfig = DitConfig.base(), device="cuda"):
        self.cfg = cfg
        self.h = Handle(torch.device(device))
        self.lib = self.h.lib
        c = _lib.DitConfig(**cfg.as_dict())
        _lib.check(self.lib.echo_dit_configure(self.h.ptr, C.byref(c)), "echo_dit_configure")
        self._ready = False
        self.has_latent = False
        # the reference rounds t to model.dtype (bf16) before the timestep embedding (inference.py:489)
        self.round_t_to_model_dtype = True

    # ---- weights ---------------------------------------------------------------------------------------------
    def load_state_dict(self, state: Iterable[Tuple[str, torch.Tensor]] | dict, strict: bool = False, assign: bool = False):
        items = state.items() if isinstance(state, dict) else state
        n_lat = 0
        for k, v in items:
            self.h.set_weight(k, v)
            n_lat += int("latent" in k)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.echo_dit_finalize(self.h.ptr, _stream(self.device)), "echo_dit_finalize")
        self._ready = True
        self.has_latent = n_lat > 0
        return self

    @classmethod
    def from_state_dict(cls, state, cfg: DitConfig = DitConfig.base(), device="cuda") -> "B200EchoDiT":
        return cls(cfg, device).load_state_dict(state)

    @classmethod
    def from_reference(cls, ref_model, cfg: DitConfig = DitConfig.base(), device="cuda") -> "B200EchoDiT":
        """Build from a loaded reference EchoDiT (any dtype/device): same state-dict keys."""
        return cls(cfg, device).load_state_dict(ref_model.state_dict())

    # ---- reference surface -----------------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.h.device

    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    def eval(self):
        return self

    def _new_cache(self, B: int, L: int) -> KVCache:
        c = self.cfg
        return [(torch.empty(B, L, c.num_heads, c.head_dim, device=self.device, dtype=torch.bfloat16),
                 torch.empty(B, L, c.num_heads, c.head_dim, device=self.device, dtype=torch.bfloat16))
                for _ in range(c.num_layers)]

    @torch.inference_mode()
    def get_kv_cache_text(self, text_input_ids: torch.Tensor, text_mask: Optional[torch.Tensor]) -> KVCache:
        ids = text_input_ids.to(self.device, torch.int32).contiguous()
        B, Lt = ids.shape
        m = None if text_mask is None else _u8(text_mask.to(self.device))
        out = self._new_cache(B, Lt)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.echo_kv_text(self.h.ptr, ids.data_ptr(), None if m is None else m.data_ptr(), B, Lt,
                                             _ptr_array([k for k, _ in out]), _ptr_array([v for _, v in out]),
                                             _stream(self.device)), "echo_kv_text")
        return out

    def _kv_patch(self, fn, latent: torch.Tensor, what: str) -> KVCache:
        x = latent.to(self.device, torch.bfloat16).contiguous()
        B, L, _ = x.shape
        out = self._new_cache(B, L // self.cfg.speaker_patch_size)
        with torch.cuda.device(self.device):
            _lib.check(fn(self.h.ptr, x.data_ptr(), B, L, _ptr_array([k for k, _ in out]),
                          _ptr_array([v for _, v in out]), _stream(self.device)), what)
        return out

    @torch.inference_mode()
    def get_kv_cache_speaker(self, speaker_latent: torch.Tensor) -> KVCache:
        return self._kv_patch(self.lib.echo_kv_speaker, speaker_latent, "echo_kv_speaker")

    @torch.inference_mode()
    def get_kv_cache_latent(self, prefix_latent: torch.Tensor) -> KVCache:
        return self._kv_patch(self.lib.echo_kv_latent, prefix_latent, "echo_kv_latent")

    @torch.inference_mode()
    def forward(self, x: torch.Tensor, t: torch.Tensor, text_mask: torch.Tensor, speaker_mask: torch.Tensor,
                kv_cache_text: KVCache, kv_cache_speaker: KVCache, start_pos: Optional[int] = None,
                kv_cache_latent: Optional[KVCache] = None, layer_outputs: Optional[list] = None,
                layer_mids: Optional[list] = None) -> torch.Tensor:
        """`layer_outputs` / `layer_mids` (optional lists, parity probes): receive the fp32 stream after every block /
        after every block's attention branch (model.py:388-389)."""
        dev = self.device
        xf = x.to(dev, torch.float32).contiguous()
        b, S, _ = xf.shape
        tf = t.to(dev, torch.float32).contiguous()
        tm, sm = _u8(text_mask.to(dev)), _u8(speaker_mask.to(dev))
        Lt, Ls = tm.shape[1], sm.shape[1]
        for cache, L in ((kv_cache_text, Lt), (kv_cache_speaker, Ls // self.cfg.speaker_patch_size)):
            k0 = cache[0][0]
            if k0.shape[0] != b or k0.shape[1] != L or k0.dtype != torch.bfloat16 or not k0.is_contiguous():
                raise ValueError(f"kv cache must be contiguous bf16 (b={b}, L={L}, H, 128); got {tuple(k0.shape)} {k0.dtype}")
        Pl = 0
        kl = vl = None
        if kv_cache_latent is not None and kv_cache_latent[0][0].shape[1] > 0:
            Pl = kv_cache_latent[0][0].shape[1]
            kl, vl = _ptr_array([k for k, _ in kv_cache_latent]), _ptr_array([v for _, v in kv_cache_latent])
        out = torch.empty(b, S, self.cfg.latent_size, device=dev, dtype=torch.float32)
        lo = None
        if layer_outputs is not None:
            bufs = [torch.empty(b, S, self.cfg.model_size, device=dev, dtype=torch.float32)
                    for _ in range(self.cfg.num_layers)]
            layer_outputs.extend(bufs)
            lo = _ptr_array(bufs)
        lm = None
        if layer_mids is not None:
            mids = [torch.empty(b, S, self.cfg.model_size, device=dev, dtype=torch.float32)
                    for _ in range(self.cfg.num_layers)]
            layer_mids.extend(mids)
            lm = _ptr_array(mids)
        with torch.cuda.device(dev):
            _lib.check(self.lib.echo_dit_forward_probe(
                self.h.ptr, xf.data_ptr(), tf.data_ptr(), tm.data_ptr(), sm.data_ptr(),
                _ptr_array([k for k, _ in kv_cache_text]), _ptr_array([v for _, v in kv_cache_text]), Lt,
                _ptr_array([k for k, _ in kv_cache_speaker]), _ptr_array([v for _, v in kv_cache_speaker]), Ls,
                kl, vl, Pl, int(start_pos or 0), b, S, out.data_ptr(), lo, lm, _stream(dev)), "echo_dit_forward")
        return out

    __call__ = forward
