"""Op-level wrappers over the C ABI (echo_op_gemm / echo_op_attention) taking torch CUDA tensors.

Used by the parity tests and micro-benchmarks; the model-level calls launch the same kernels from C++.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_NONE, EPI_GENERIC, EPI_QKV, EPI_SWIGLU, AttnDesc, GemmDesc


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def gemm(a: torch.Tensor, w: torch.Tensor, *, taps: int = 1, tap_shift: Sequence[int] = (0,), bias=None,
         scale: float = 1.0, gate=None, rows_per_gate: int = 0, resid=None, out_f32=None, out_bf16=None,
         act: int = ACT_NONE, alpha=None, col_mod: int = 0, bn: int = 0, cg: int = 0, trace=None,
         split_k: int = 0) -> None:
    """out = epilogue(sum_taps a[rows + shift] @ w[:, tap*Kc:(tap+1)*Kc].T).  a: (batches, M, Kc) or (M, Kc) bf16."""
    lib = _lib.load(strict=False)
    if a.dim() == 2:
        a = a.unsqueeze(0)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(-1) == 1 and w.stride(-1) == 1
    batches, M, Kc = a.shape
    N = w.shape[0]
    d = GemmDesc()
    d.A, d.lda, d.a_batch_stride = a.data_ptr(), a.stride(1), a.stride(0)
    d.B, d.ldb, d.b_rows = w.data_ptr(), w.stride(0), N
    d.M, d.N, d.Kc, d.batches, d.taps = M, N, Kc, batches, taps
    for i, s in enumerate(tap_shift):
        d.tap_shift[i] = int(s)
    d.epi = EPI_GENERIC
    d.bias, d.scale = _ptr(bias), scale
    d.gate, d.rows_per_gate, d.gate_ld = _ptr(gate), rows_per_gate, (gate.stride(0) if gate is not None and gate.dim() > 1 else 0)
    d.resid, d.out_f32 = _ptr(resid), _ptr(out_f32)
    d.ld_f32 = (out_f32 if out_f32 is not None else resid).stride(-2) if (out_f32 is not None or resid is not None) else 0
    d.out_bf16, d.ld_bf16 = _ptr(out_bf16), (out_bf16.stride(-2) if out_bf16 is not None else 0)
    d.act, d.alpha, d.col_mod, d.bn, d.cg = act, _ptr(alpha), col_mod, bn, cg
    d.trace = _ptr(trace)
    d.split_k = split_k
    _lib.check(lib.echo_op_gemm(C.byref(d), _stream()), "echo_op_gemm")


def residual_unit(a: torch.Tensor, w7: torch.Tensor, bias7, alpha2, w1: torch.Tensor, bias1, stream_f32: torch.Tensor,
                  alpha_out, out_bf16: torch.Tensor, dilation: int, trace=None) -> None:
    """Fused DAC ResidualUnit (autoencoder.py:884-900) on time-major activations: a (T, C) bf16 = Snake(alpha1)(x) (the
    conv7 input), w7 (C, 7 C) bf16 [cout][tap][cin], w1 (C, C) bf16; stream_f32 (T, C) fp32 is x and is UPDATED in place
    (x + conv1(Snake(alpha2)(conv7(a)))); out_bf16 (T, C) = Snake(alpha_out)(new x), the next unit's conv input."""
    from ._lib import EPI_RU
    lib = _lib.load(strict=False)
    T, Cc = a.shape
    d = GemmDesc()
    d.A, d.lda, d.a_batch_stride = a.data_ptr(), a.stride(0), 0
    d.B, d.ldb, d.b_rows = w7.data_ptr(), w7.stride(0), Cc
    d.M, d.N, d.Kc, d.batches, d.taps = T, Cc, Cc, 1, 7
    for j in range(7):
        d.tap_shift[j] = -(6 - j) * dilation
    d.epi = EPI_RU
    d.bias, d.scale, d.alpha = _ptr(bias7), 1.0, _ptr(alpha2)
    d.B1, d.ldb1, d.ru_bias1, d.ru_alpha_out = w1.data_ptr(), w1.stride(0), _ptr(bias1), _ptr(alpha_out)
    d.resid, d.out_f32, d.ld_f32 = stream_f32.data_ptr(), stream_f32.data_ptr(), stream_f32.stride(0)
    d.out_bf16, d.ld_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    d.trace = _ptr(trace)
    _lib.check(lib.echo_op_gemm(C.byref(d), _stream()), "echo_op_gemm(residual unit)")


def gemm_swiglu(a: torch.Tensor, w13: torch.Tensor, out_bf16: torch.Tensor, cg: int = 0, trace=None) -> None:
    """w13: (2*I, K) packed so that each 256-row tile is [128 rows of w1 | the matching 128 rows of w3]."""
    lib = _lib.load(strict=False)
    M, K = a.shape
    d = GemmDesc()
    d.A, d.lda, d.a_batch_stride = a.data_ptr(), a.stride(0), 0
    d.B, d.ldb, d.b_rows = w13.data_ptr(), w13.stride(0), w13.shape[0]
    d.M, d.N, d.Kc, d.batches, d.taps = M, w13.shape[0], K, 1, 1
    d.epi = EPI_SWIGLU
    d.cg = cg
    d.trace = _ptr(trace)
    d.out_bf16, d.ld_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    _lib.check(lib.echo_op_gemm(C.byref(d), _stream()), "echo_op_gemm(swiglu)")


def gemm_qkv(a: torch.Tensor, w: torch.Tensor, outs, norm_ws, rope_heads, sigmoids, sec_width: int, rope_cos=None,
             rope_sin=None, head_dim: int = 128, pos_period: int = 1, pos_offset: int = 0, pos_mult: int = 1,
             eps: float = 1e-5, cg: int = 0, trace=None, bn: int = 0) -> None:
    lib = _lib.load(strict=False)
    M, K = a.shape
    d = GemmDesc()
    d.A, d.lda, d.a_batch_stride = a.data_ptr(), a.stride(0), 0
    d.B, d.ldb, d.b_rows = w.data_ptr(), w.stride(0), w.shape[0]
    d.M, d.N, d.Kc, d.batches, d.taps = M, w.shape[0], K, 1, 1
    d.epi = EPI_QKV
    d.cg = cg
    d.bn = bn  # 0 = auto; 384 = the single-wave pair tile (two MMAs per K step)
    d.trace = _ptr(trace)
    for i, o in enumerate(outs):
        d.sec_out[i] = o.data_ptr()
        d.sec_norm_w[i] = _ptr(norm_ws[i])
        d.sec_rope_heads[i] = rope_heads[i]
        d.sec_sigmoid[i] = sigmoids[i]
    d.sec_width, d.rope_cos, d.rope_sin, d.head_dim = sec_width, _ptr(rope_cos), _ptr(rope_sin), head_dim
    d.pos_period, d.pos_offset, d.pos_mult, d.eps = pos_period, pos_offset, pos_mult, eps
    _lib.check(lib.echo_op_gemm(C.byref(d), _stream()), "echo_op_gemm(qkv)")


def attention(q: torch.Tensor, segments, out: torch.Tensor, gate: Optional[torch.Tensor] = None,
              scale: Optional[float] = None, trace: Optional[torch.Tensor] = None, nsplit: int = 0) -> None:
    """q: (b, S, H, D) bf16. segments: list of dicts with keys k, v ((b, L, H, D) bf16) and optional mask (b, L) bool,
    eff_len (b,) int32, pos_limit_mult, pos_limit, causal, window, q_offset. nsplit: split-KV over a cluster of CTAs
    (0 = auto, 1 = off, 2..8 = force)."""
    lib = _lib.load(strict=False)
    b, S, H, D = q.shape
    d = AttnDesc()
    d.Q, d.q_batch_stride, d.q_row_stride = q.data_ptr(), q.stride(0), q.stride(1)
    d.gate, d.out = _ptr(gate), out.data_ptr()
    d.b, d.S, d.H, d.D = b, S, H, D
    d.scale = scale if scale is not None else D ** -0.5
    d.nseg = len(segments)
    d.trace = _ptr(trace)
    d.nsplit = int(nsplit)
    keep = []
    for i, sg in enumerate(segments):
        k, v = sg["k"], sg["v"]
        assert k.stride() == v.stride() and k.stride(3) == 1 and k.stride(2) == D
        s = d.seg[i]
        s.K, s.V = k.data_ptr(), v.data_ptr()
        s.batch_stride, s.row_stride, s.len = (k.stride(0) if k.shape[0] > 1 else 0), k.stride(1), k.shape[1]
        m = sg.get("mask")
        if m is not None:
            m8 = m.view(torch.uint8) if m.dtype == torch.bool else m
            keep.append(m8)
            s.mask, s.mask_ld, s.mask_stride = m8.data_ptr(), m8.stride(0), m8.stride(1)
        e = sg.get("eff_len")
        if e is not None:
            s.eff_len = e.data_ptr()
        s.pos_limit_mult, s.pos_limit = sg.get("pos_limit_mult", 0), sg.get("pos_limit", 0)
        s.causal, s.window = int(sg.get("causal", 0)), int(sg.get("window", 0))
        s.q_offset = int(sg.get("q_offset", 0))
        s.kv_scale = float(sg.get("kv_scale", 0.0))
        s.batch_mod = int(sg.get("batch_mod", 0))
    _lib.check(lib.echo_op_attention(C.byref(d), _stream()), "echo_op_attention")


def rmsnorm_affine(x: torch.Tensor, a: torch.Tensor, c0: Optional[torch.Tensor] = None, rows_per_group: int = 0,
                   eps: float = 1e-5) -> torch.Tensor:
    """x (rows, W) fp32; a / c0 (groups, W) fp32 -> bf16 (rows, W)."""
    lib = _lib.load(strict=False)
    rows, W = x.shape
    out = torch.empty(rows, W, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.echo_op_rmsnorm_affine(x.data_ptr(), out.data_ptr(), a.data_ptr(), _ptr(c0), rows, W, rows_per_group,
                                          a.stride(0) if a.dim() == 2 else 0, eps, _stream()), "echo_op_rmsnorm_affine")
    return out


def cfg_euler_update(x: torch.Tensor, v: torch.Tensor, has_cfg: bool, s_text: float, s_spk: float, dt: float,
                     rescale: Optional[tuple] = None) -> None:
    """In place: x += dt * combine(v). v is (3, ...) when has_cfg else (1, ...); rescale = (one_minus_t, ratio)."""
    lib = _lib.load(strict=False)
    omt, ratio = rescale if rescale is not None else (0.0, 1.0)
    _lib.check(lib.echo_op_cfg_euler_update(x.data_ptr(), v.data_ptr(), x.numel(), int(has_cfg), s_text, s_spk,
                                            int(rescale is not None), omt, ratio, dt, _stream()),
               "echo_op_cfg_euler_update")
