"""ctypes binding of libecho_b200.so (see include/echo_b200.h).

The shared library is the product; this module only loads it and mirrors its C structs. There is no CPU
fallback: if the library is missing or was not built, importing a compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libecho_b200.so"

c_void_pp = C.POINTER(C.c_void_p)


class DitConfig(C.Structure):
    _fields_ = [
        ("latent_size", C.c_int), ("model_size", C.c_int), ("num_layers", C.c_int), ("num_heads", C.c_int),
        ("intermediate_size", C.c_int), ("norm_eps", C.c_float),
        ("text_vocab_size", C.c_int), ("text_model_size", C.c_int), ("text_num_layers", C.c_int),
        ("text_num_heads", C.c_int), ("text_intermediate_size", C.c_int),
        ("speaker_patch_size", C.c_int), ("speaker_model_size", C.c_int), ("speaker_num_layers", C.c_int),
        ("speaker_num_heads", C.c_int), ("speaker_intermediate_size", C.c_int),
        ("timestep_embed_size", C.c_int), ("adaln_rank", C.c_int),
    ]


class DacConfig(C.Structure):
    _fields_ = [
        ("latent_dim", C.c_int), ("pca_dim", C.c_int), ("post_layers", C.c_int), ("post_heads", C.c_int),
        ("post_intermediate", C.c_int), ("post_window", C.c_int), ("post_norm_eps", C.c_float),
        ("num_upsample", C.c_int), ("decoder_dim", C.c_int), ("num_rates", C.c_int), ("rates", C.c_int * 8),
        ("enc_dim", C.c_int), ("num_enc_rates", C.c_int), ("enc_rates", C.c_int * 8), ("enc_t_layers", C.c_int),
        ("enc_window", C.c_int), ("n_codebooks", C.c_int), ("codebook_size", C.c_int),
        ("semantic_codebook_size", C.c_int), ("codebook_dim", C.c_int),
    ]


class SamplerArgs(C.Structure):
    _fields_ = [
        ("num_steps", C.c_int), ("cfg_scale_text", C.c_float), ("cfg_scale_speaker", C.c_float),
        ("cfg_min_t", C.c_float), ("cfg_max_t", C.c_float),
        ("has_truncation", C.c_int), ("truncation_factor", C.c_float),
        ("has_rescale", C.c_int), ("rescale_k", C.c_float), ("rescale_sigma", C.c_float),
        ("has_kv_scale", C.c_int), ("speaker_kv_scale", C.c_float), ("speaker_kv_max_layers", C.c_int),
        ("speaker_kv_min_t", C.c_float),
        ("sequence_length", C.c_int), ("round_t_to_bf16", C.c_int),
        ("t_schedule", C.POINTER(C.c_float)),
        ("speaker_K", C.POINTER(C.c_void_p)), ("speaker_V", C.POINTER(C.c_void_p)),
        ("text_valid_len", C.c_int),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_batch_stride", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_rows", C.c_int64),
        ("M", C.c_int), ("N", C.c_int), ("Kc", C.c_int), ("batches", C.c_int), ("taps", C.c_int),
        ("tap_shift", C.c_int * 8),
        ("epi", C.c_int),
        ("bias", C.c_void_p), ("scale", C.c_float),
        ("gate", C.c_void_p), ("rows_per_gate", C.c_int), ("gate_ld", C.c_int),
        ("resid", C.c_void_p), ("out_f32", C.c_void_p), ("ld_f32", C.c_int),
        ("out_bf16", C.c_void_p), ("ld_bf16", C.c_int),
        ("act", C.c_int), ("alpha", C.c_void_p), ("col_mod", C.c_int),
        ("sec_out", C.c_void_p * 4), ("sec_norm_w", C.c_void_p * 4), ("sec_rope_heads", C.c_int * 4),
        ("sec_sigmoid", C.c_int * 4),
        ("sec_width", C.c_int), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("head_dim", C.c_int),
        ("pos_period", C.c_int), ("pos_offset", C.c_int), ("pos_mult", C.c_int), ("eps", C.c_float),
        ("bn", C.c_int), ("cg", C.c_int), ("reserved0", C.c_int), ("trace", C.c_void_p), ("split_k", C.c_int),
        ("B1", C.c_void_p), ("ldb1", C.c_int64), ("ru_bias1", C.c_void_p), ("ru_alpha_out", C.c_void_p),
    ]


class AttnSegment(C.Structure):
    _fields_ = [
        ("K", C.c_void_p), ("V", C.c_void_p), ("batch_stride", C.c_int64), ("row_stride", C.c_int64),
        ("batch_mod", C.c_int),
        ("len", C.c_int), ("eff_len", C.c_void_p),
        ("mask", C.c_void_p), ("mask_ld", C.c_int), ("mask_stride", C.c_int),
        ("pos_limit_mult", C.c_int), ("pos_limit", C.c_int), ("causal", C.c_int), ("window", C.c_int),
        ("q_offset", C.c_int), ("kv_scale", C.c_float),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p), ("q_batch_stride", C.c_int64), ("q_row_stride", C.c_int64),
        ("gate", C.c_void_p), ("out", C.c_void_p),
        ("b", C.c_int), ("S", C.c_int), ("H", C.c_int), ("D", C.c_int), ("scale", C.c_float),
        ("nseg", C.c_int), ("seg", AttnSegment * 4), ("trace", C.c_void_p),
        ("nsplit", C.c_int),
    ]


class ProfileReport(C.Structure):
    _fields_ = [("launches", C.c_int64 * 3), ("ms", C.c_double * 3), ("flops", C.c_double * 3),
                ("bytes", C.c_double * 3)]


EPI_GENERIC, EPI_SWIGLU, EPI_QKV, EPI_RU = 0, 1, 2, 4
ACT_NONE, ACT_GELU, ACT_SNAKE, ACT_TANH, ACT_SIGMOID, ACT_SILU = 0, 1, 2, 3, 4, 5
DTYPE_F32, DTYPE_BF16 = 0, 1

# every symbol include/echo_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
BLOCK_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)  # echo_block_cb(user, index, start, len)
SYMBOLS = {
    "echo_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "echo_destroy": (C.c_int, [_P]),
    "echo_last_error": (C.c_char_p, []),
    "echo_num_launches": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "echo_dit_configure": (C.c_int, [_P, C.POINTER(DitConfig)]),
    "echo_dac_configure": (C.c_int, [_P, C.POINTER(DacConfig)]),
    "echo_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, C.c_int, _P]),
    "echo_dit_finalize": (C.c_int, [_P, _P]),
    "echo_dac_finalize": (C.c_int, [_P, _P]),
    "echo_kv_text": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, c_void_pp, c_void_pp, _P]),
    "echo_kv_speaker": (C.c_int, [_P, _P, C.c_int, C.c_int, c_void_pp, c_void_pp, _P]),
    "echo_kv_latent": (C.c_int, [_P, _P, C.c_int, C.c_int, c_void_pp, c_void_pp, _P]),
    "echo_dit_forward": (C.c_int, [_P, _P, _P, _P, _P, c_void_pp, c_void_pp, C.c_int, c_void_pp, c_void_pp, C.c_int,
                                   c_void_pp, c_void_pp, C.c_int, C.c_int, C.c_int, C.c_int, _P, c_void_pp, _P]),
    "echo_dit_forward_probe": (C.c_int, [_P, _P, _P, _P, _P, c_void_pp, c_void_pp, C.c_int, c_void_pp, c_void_pp, C.c_int,
                                         c_void_pp, c_void_pp, C.c_int, C.c_int, C.c_int, C.c_int, _P, c_void_pp, c_void_pp,
                                         _P]),
    "echo_sample_euler": (C.c_int, [_P, C.POINTER(SamplerArgs), _P, _P, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "echo_sample_blockwise": (C.c_int, [_P, C.POINTER(SamplerArgs), C.POINTER(C.c_int), C.c_int, _P, _P, C.c_int, _P,
                                        _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P]),
    "echo_sample_blockwise_stream": (C.c_int, [_P, C.POINTER(SamplerArgs), C.POINTER(C.c_int), C.c_int, _P, _P, C.c_int,
                                               _P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, BLOCK_CB, _P, _P]),
    "echo_dac_decode": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int, C.c_int, _P, _P]),
    "echo_dac_stream_create": (C.c_int, [_P, C.c_int, C.POINTER(_P), _P]),
    "echo_dac_stream_reset": (C.c_int, [_P, _P, _P]),
    "echo_dac_stream_decode": (C.c_int, [_P, _P, _P, _P, _P, C.c_float, C.c_int, _P, _P]),
    "echo_dac_stream_position": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "echo_dac_stream_destroy": (C.c_int, [_P, _P]),
    "echo_dac_encode_zq": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "echo_dac_encode": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int, C.c_int, _P, _P]),
    "echo_dac_decode_zq": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "echo_sample_euler_host": (C.c_int, [_P, C.POINTER(SamplerArgs), _P, _P, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P]),
    "echo_profile_start": (C.c_int, [_P]),
    "echo_profile_stop": (C.c_int, [_P, C.POINTER(ProfileReport)]),
    "echo_set_deterministic": (C.c_int, [C.c_int]),
    "echo_op_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "echo_op_attention": (C.c_int, [C.POINTER(AttnDesc), _P]),
    "echo_op_rmsnorm_affine": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_float, _P]),
    "echo_op_flattening_point": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, _P, _P]),
    "echo_op_cfg_euler_update": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                          C.c_float, C.c_float, _P]),
}

_lib = None


class EchoError(RuntimeError):
    pass


def load(strict: bool = True) -> C.CDLL:
    """Load libecho_b200.so and bind every declared symbol. Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ECHO_B200_LIB", str(LIB_PATH))
    if not os.path.exists(path):
        raise EchoError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(echo_tts_b200/csrc/build.sh). There is no CPU fallback.")
    lib = C.CDLL(path)
    missing = []
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing and strict:
        raise EchoError(f"libecho_b200.so lacks symbols: {missing}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().echo_last_error()
        raise EchoError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
