"""Architecture descriptions of the two networks on the hot path (plain dataclasses, no torch).

Values for `DitConfig.base()` are the echo-tts-base constructor arguments at reference inference.py:16-24;
`DacConfig.base()` follows build_ae() at reference autoencoder.py:1144-1192.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import List


@dataclass(frozen=True)
class DitConfig:
    latent_size: int = 80
    model_size: int = 2048
    num_layers: int = 24
    num_heads: int = 16
    intermediate_size: int = 5888
    norm_eps: float = 1e-5
    text_vocab_size: int = 256
    text_model_size: int = 1280
    text_num_layers: int = 14
    text_num_heads: int = 10
    text_intermediate_size: int = 3328
    speaker_patch_size: int = 4
    speaker_model_size: int = 1280
    speaker_num_layers: int = 14
    speaker_num_heads: int = 10
    speaker_intermediate_size: int = 3328
    timestep_embed_size: int = 512
    adaln_rank: int = 256

    @staticmethod
    def base() -> "DitConfig":
        return DitConfig()

    @staticmethod
    def tiny() -> "DitConfig":
        """Small but structurally complete (head_dim stays 128, which the kernels require)."""
        return DitConfig(model_size=256, num_layers=3, num_heads=2, intermediate_size=512,
                         text_model_size=256, text_num_layers=2, text_num_heads=2, text_intermediate_size=384,
                         speaker_model_size=256, speaker_num_layers=2, speaker_num_heads=2,
                         speaker_intermediate_size=384, timestep_embed_size=128, adaln_rank=64)

    @property
    def head_dim(self) -> int:
        return self.model_size // self.num_heads

    def as_dict(self):
        return asdict(self)


@dataclass(frozen=True)
class DacConfig:
    latent_dim: int = 1024
    pca_dim: int = 80
    post_layers: int = 8
    post_heads: int = 16
    post_intermediate: int = 3072
    post_window: int = 128
    post_norm_eps: float = 1e-5
    post_block_size: int = 4096
    num_upsample: int = 2
    convnext_mlp_ratio: int = 4
    decoder_dim: int = 1536
    rates: List[int] = field(default_factory=lambda: [8, 8, 4, 2])
    # ---- encode path (reference build_ae, autoencoder.py:1144-1195): Encoder + downsample + pre_module + RVQ
    enc_dim: int = 64
    enc_rates: List[int] = field(default_factory=lambda: [2, 4, 8, 8])
    enc_t_layers: int = 4          # transformer layers of the LAST encoder block (encoder_transformer_layers=[0,0,0,4])
    enc_window: int = 512          # EncoderBlock's WindowLimitedTransformer window (autoencoder.py:855)
    n_codebooks: int = 9
    codebook_size: int = 1024
    semantic_codebook_size: int = 4096
    codebook_dim: int = 8

    @staticmethod
    def base() -> "DacConfig":
        return DacConfig()

    @staticmethod
    def tiny() -> "DacConfig":
        """Same topology at 1/4 width, 2 transformer layers; last stage still hits the 96-channel GEMM path."""
        return DacConfig(latent_dim=256, post_layers=2, post_heads=4, post_intermediate=768, post_window=16,
                         decoder_dim=1536, rates=[8, 8, 4, 2], enc_dim=64, enc_rates=[2, 4], enc_t_layers=1,
                         enc_window=512, n_codebooks=3, codebook_size=32, semantic_codebook_size=64)

    @property
    def hop(self) -> int:
        h = 2 ** self.num_upsample
        for r in self.rates:
            h *= r
        return h

    @property
    def enc_hop(self) -> int:
        """Encoder stride product (DAC.hop_length); one z_q frame = enc_hop * 2**num_upsample samples (frame_length)."""
        h = 1
        for r in self.enc_rates:
            h *= r
        return h

    @property
    def frame_length(self) -> int:
        return self.enc_hop * 2 ** self.num_upsample

    def as_dict(self):
        return asdict(self)
