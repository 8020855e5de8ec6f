"""echo-b200: B200-native (sm_100a) implementation of the Echo-TTS sampling hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + C ABI), `_lib` (ctypes binding) and thin Python
mirrors of the reference's call signatures (model / sample_fn / fish_ae / pca_state).
"""
__all__ = ["_lib", "ops", "set_deterministic"]


def set_deterministic(on: bool = True) -> None:
    """Process-wide switch: True = bit-reproducible results (the split-K slices of the residual-accumulate GEMMs are
    summed in a fixed order instead of with fp32 atomics, about 1 % slower at batch 1); False (default) = fastest.
    Both meet the same tolerances against the reference."""
    from . import _lib
    _lib.check(_lib.load().echo_set_deterministic(int(bool(on))), "echo_set_deterministic")
