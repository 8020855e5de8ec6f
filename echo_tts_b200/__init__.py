"""echo-b200: B200-native (sm_100a) implementation of the Echo-TTS sampling hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + C ABI), `_lib` (ctypes binding) and thin Python
mirrors of the reference's call signatures (model / sample_fn / fish_ae / pca_state).
"""
__all__ = ["_lib", "ops"]
