/*
 * echo_b200.h -- C ABI of libecho_b200.so: the B200-native (sm_100a) implementation of the Echo-TTS sampling
 * hot path (Euler sampler + independent text/speaker CFG + 24-layer latent DiT + Fish S1-DAC decoder).
 *
 * The reference (sruckh/echo-tts) has no FFI: its boundary is four Python objects (model, sample_fn, fish_ae,
 * pca_state) threaded through inference.sample_pipeline (reference inference.py:309-347). Each entry point below
 * names the reference function it replaces; the echo_tts_b200 Python package mirrors the Python call signatures on top of
 * these symbols, INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success and a negative ECHO_ERR_* code on failure; echo_last_error() returns a
 *     human-readable message for the calling thread. No exception crosses the ABI.
 *   - all data pointers are CUDA DEVICE pointers owned by the caller unless a parameter says "host".
 *   - `stream` is a cudaStream_t passed as void*; the library never synchronises the device behind the caller's
 *     back except in the *_host entry points (documented there).
 *   - a handle is bound to one device. Its compute entry points are serialised by a per-handle mutex, and a call on
 *     a different stream than the previous call first waits (on the device) for the work the handle enqueued on the old
 *     stream: the activation / KV workspaces are per handle. For concurrency use one handle per worker. A stream
 *     handed to the library must stay alive until the next call on the same handle has been issued.
 *   - one process may drive several devices (one handle each); per-device kernel attributes are set on first use.
 *   - there is NO CPU fallback: without an sm_100 device echo_create fails with ECHO_ERR_DEVICE.
 */
#ifndef ECHO_B200_H
#define ECHO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECHO_OK 0
#define ECHO_ERR_ARG (-1)     /* bad argument / unsupported shape */
#define ECHO_ERR_CUDA (-2)    /* CUDA runtime / driver error */
#define ECHO_ERR_DEVICE (-3)  /* no sm_100 device */
#define ECHO_ERR_STATE (-4)   /* call order (e.g. weights missing) */

#define ECHO_DTYPE_F32 0
#define ECHO_DTYPE_BF16 1

typedef struct echo_handle echo_handle;

/* Architecture of EchoDiT (reference model.py:472-559; echo-tts-base values at inference.py:16-24). */
typedef struct echo_dit_config {
  int latent_size;       /* 80   */
  int model_size;        /* 2048 */
  int num_layers;        /* 24   */
  int num_heads;         /* 16   (head_dim must be 128) */
  int intermediate_size; /* 5888 */
  float norm_eps;        /* 1e-5 */
  int text_vocab_size;   /* 256  */
  int text_model_size;   /* 1280 */
  int text_num_layers;   /* 14   */
  int text_num_heads;    /* 10   */
  int text_intermediate_size; /* 3328 */
  int speaker_patch_size;     /* 4    */
  int speaker_model_size;     /* 1280 */
  int speaker_num_layers;     /* 14   */
  int speaker_num_heads;      /* 10   */
  int speaker_intermediate_size; /* 3328 */
  int timestep_embed_size;       /* 512  */
  int adaln_rank;                /* 256  */
} echo_dit_config;

/* Architecture of the Fish S1-DAC decode path (reference autoencoder.py:1144-1192 build_ae). */
typedef struct echo_dac_config {
  int latent_dim;      /* 1024 */
  int pca_dim;         /* 80   */
  int post_layers;     /* 8    quantizer.post_module (window-limited causal transformer) */
  int post_heads;      /* 16   (head_dim 64) */
  int post_intermediate; /* 3072 */
  int post_window;     /* 128  */
  float post_norm_eps; /* 1e-5 */
  int num_upsample;    /* 2    quantizer.upsample stages (ConvTranspose k2 s2 + ConvNeXt) */
  int decoder_dim;     /* 1536 */
  int num_rates;       /* 4    */
  int rates[8];        /* 8, 8, 4, 2 */
  /* encode path (Encoder + quantizer.downsample / pre_module + RVQ); only used when the encoder weights are loaded */
  int enc_dim;         /* 64   */
  int num_enc_rates;   /* 4    */
  int enc_rates[8];    /* 2, 4, 8, 8 */
  int enc_t_layers;    /* 4    transformer layers of the last EncoderBlock (window enc_window, head_dim 64) */
  int enc_window;      /* 512  */
  int n_codebooks;     /* 9    residual codebooks (plus one semantic codebook) */
  int codebook_size;   /* 1024 */
  int semantic_codebook_size; /* 4096 */
  int codebook_dim;    /* 8    */
} echo_dac_config;

/* Knobs of sample_euler_cfg_independent_guidances (reference inference.py:427-447). */
typedef struct echo_sampler_args {
  int num_steps;
  float cfg_scale_text;
  float cfg_scale_speaker;
  float cfg_min_t;
  float cfg_max_t;
  int has_truncation; float truncation_factor;
  int has_rescale;    float rescale_k; float rescale_sigma;
  int has_kv_scale;   float speaker_kv_scale; int speaker_kv_max_layers; /* < 0: all layers (None); else min(n, num_layers), 0 = none */ float speaker_kv_min_t;
  int sequence_length;   /* latents to generate (<= 640 in the reference) */
  int round_t_to_bf16;   /* 1: t is rounded to bf16 before the timestep embedding, as the reference does when
                            model.dtype is bfloat16 (inference.py:489); 0: fp32 t */
  const float* t_schedule; /* optional HOST array of num_steps+1 floats (torch.linspace(1,0,n+1)*0.999 as the caller's
                              torch computes it, inference.py:459); NULL: the library's own fp32 closed form */
  /* Optional per-voice persistence: HOST arrays of num_layers device pointers to a speaker KV cache built
     earlier by echo_kv_speaker for the SAME speaker_latent (each (B, Ls/4, 16, 128) bf16 contiguous). When both are
     non-NULL the samplers skip the speaker encoder + 24 K/V projections (the reference recomputes them on every call,
     inference.py:465). The cache is never written: speaker_kv_scale is applied inside the attention kernel
     (the reference scales its per-call cache in place, inference.py:467-468, 511-513). */
  void* const* speaker_K;
  void* const* speaker_V;
  /* Optional hint, 0 = unknown: every text_mask entry at index >= text_valid_len is 0 in every batch row (the masks
     get_text_input_ids_and_mask builds are such prefixes, inference.py:192-215). The samplers then run the text encoder
     and the 24 K/V projections over the first text_valid_len rows only -- the rows behind it are masked out of every
     attention, in the reference too (model.py:606-613 computes them and never uses them). The host entry derives it
     from the mask it is given. */
  int text_valid_len;
} echo_sampler_args;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int echo_create(echo_handle** out, int device);
int echo_destroy(echo_handle* h);
const char* echo_last_error(void);
int echo_num_launches(echo_handle* h, int64_t* out); /* kernels launched by this handle so far */
/* Process-wide: 1 = bit-reproducible results (the split-K slices of the residual-accumulate GEMMs are summed in a
 * fixed order instead of with fp32 atomics, ~1 % slower at batch 1); 0 (default) = fastest. ECHO_DETERMINISTIC=1 in the
 * environment sets it at start-up. Both settings meet the same tolerances against the reference. */
int echo_set_deterministic(int on);

/* ---- weights: replaces load_state_dict (reference inference.py:14-47, 56-76) ---------------------------- */
int echo_dit_configure(echo_handle* h, const echo_dit_config* cfg);
int echo_dac_configure(echo_handle* h, const echo_dac_config* cfg);
/* `key` is the reference state-dict key (e.g. "blocks.3.attention.wq.weight", "dac.decoder.model.1.block.0.alpha"
 * -- DAC keys carry a "dac." prefix). The tensor is copied/packed into library-owned bf16/fp32 buffers (QKV|gate and
 * W1|W3 fused, weight-norm folded), so the caller may free `data` afterwards. dtype: ECHO_DTYPE_*. */
int echo_set_weight(echo_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype,
                    void* stream);
int echo_dit_finalize(echo_handle* h, void* stream); /* checks that every required key arrived, builds tables */
int echo_dac_finalize(echo_handle* h, void* stream);

/* ---- EchoDiT: replaces model.py:563-636 ------------------------------------------------------------------ */
/* K[i], V[i]: 24 (num_layers) output device pointers, each (B, L, num_heads, 128) bf16 contiguous. */
int echo_kv_text(echo_handle* h, const int32_t* ids, const uint8_t* mask /* (B,Lt) bool or NULL */, int B, int Lt,
                 void* const* K, void* const* V, void* stream);
int echo_kv_speaker(echo_handle* h, const void* latent_bf16 /* (B,Ls,80) */, int B, int Ls, void* const* K,
                    void* const* V, void* stream);
int echo_kv_latent(echo_handle* h, const void* prefix_bf16 /* (B,Lp,80) */, int B, int Lp, void* const* K,
                   void* const* V, void* stream);
/* == EchoDiT.forward. x (b,S,80) fp32; t (b,) fp32; masks uint8 (bool) (b,Lt) / (b,Ls_unstrided);
 * caches as produced above with batch b; Kl/Vl may be NULL (Pl = 0). out (b,S,80) fp32.
 * layer_out: optional (num_layers) device pointers receiving each block's output (b,S,D) fp32 (parity tests). */
int echo_dit_forward(echo_handle* h, const float* x, const float* t, const uint8_t* text_mask,
                     const uint8_t* speaker_mask, void* const* Kt, void* const* Vt, int Lt, void* const* Ks,
                     void* const* Vs, int Ls_unstrided, void* const* Kl, void* const* Vl, int Pl, int start_pos,
                     int b, int S, float* out, void* const* layer_out, void* stream);

/* Same call with a second probe array: layer_mid[i] (optional, (b,S,D) fp32) receives the stream after block i's
 * attention branch, i.e. the input of mlp_adaln (model.py:388), so that parity tests can compare the two residual
 * increments of every block (attention and MLP, model.py:384-389) with the reference's separately. */
int echo_dit_forward_probe(echo_handle* h, const float* x, const float* t, const uint8_t* text_mask,
                           const uint8_t* speaker_mask, void* const* Kt, void* const* Vt, int Lt, void* const* Ks,
                           void* const* Vs, int Ls_unstrided, void* const* Kl, void* const* Vl, int Pl, int start_pos,
                           int b, int S, float* out, void* const* layer_out, void* const* layer_mid, void* stream);

/* ---- samplers: replace inference.py:427-517 and inference_blockwise.py:15-123 -------------------------- */
/* noise: (B, sequence_length, 80) fp32 drawn by the caller (torch.randn with the reference's generator, so seeds
 * stay compatible). x_out: (B, sequence_length, 80) fp32. */
int echo_sample_euler(echo_handle* h, const echo_sampler_args* a, const void* speaker_latent_bf16,
                      const uint8_t* speaker_mask, int Ls, const int32_t* text_ids, const uint8_t* text_mask,
                      int Lt, int B, const float* noise, float* x_out, void* stream);
/* noise: concatenation over blocks of (B, block, 80); continuation (B,Lc,80) fp32 or NULL;
 * prefix_out (B, Lc + sum(blocks), 80) fp32. */
int echo_sample_blockwise(echo_handle* h, const echo_sampler_args* a, const int* block_sizes, int nblocks,
                          const void* speaker_latent_bf16, const uint8_t* speaker_mask, int Ls,
                          const int32_t* text_ids, const uint8_t* text_mask, int Lt, int B,
                          const float* continuation, int Lc, const float* noise, float* prefix_out, void* stream);
/* Streaming form: identical arithmetic; after the kernels of block `index` have been ENQUEUED, `cb` is
 * called on the host with the position of the block inside prefix_out (rows [start, start + len) of every batch
 * item are final in stream order). The callback may enqueue work on the same stream -- e.g. echo_dac_decode of
 * prefix_out[:, :start + len] -- so audio of block i is produced while block i + 1 samples. cb may be NULL. */
typedef void (*echo_block_cb)(void* user, int index, int start, int len);
int echo_sample_blockwise_stream(echo_handle* h, const echo_sampler_args* a, const int* block_sizes, int nblocks,
                                 const void* speaker_latent_bf16, const uint8_t* speaker_mask, int Ls,
                                 const int32_t* text_ids, const uint8_t* text_mask, int Lt, int B,
                                 const float* continuation, int Lc, const float* noise, float* prefix_out,
                                 echo_block_cb cb, void* user, void* stream);

/* ---- DAC decode: replaces ae_decode (inference.py:226-229) + DAC.decode_zq (autoencoder.py:1128-1132) ---- */
/* z (B,T,80) fp32 PCA latents; pca_components (80,1024) fp32; pca_mean (1024) fp32; audio (B,1,2048*T) fp32. */
int echo_dac_decode(echo_handle* h, const float* z, const float* pca_components, const float* pca_mean,
                    float latent_scale, int B, int T, float* audio, void* stream);
/* == DAC.decode_zq: zq (B,1024,T) fp32 channels-first. */
int echo_dac_decode_zq(echo_handle* h, const float* zq, int B, int T, float* audio, void* stream);

/* ---- streaming DAC decode (reference gradio_app.py:43 "decode chunk-wise", the blockwise sampler writes
 * its blocks back one by one, inference_blockwise.py:120-123). A stream carries what the causal decoder remembers
 * between blocks -- the window-128 keys / values of the 8 post_module layers (autoencoder.py:762-773) and the (k-1)*d
 * input rows in front of every causal conv (autoencoder.py:285-289) -- so decoding the latents block by block costs the
 * same as decoding them once and produces BIT-IDENTICAL samples. One batch item per stream; max_latents <= 4096.
 *   echo_dac_stream_decode : z (1, T, 80) fp32 = the next T latents; audio (1, 1, 2048*T) fp32 = the next samples. */
typedef struct echo_dac_stream echo_dac_stream;
int echo_dac_stream_create(echo_handle* h, int max_latents, echo_dac_stream** out, void* stream);
int echo_dac_stream_reset(echo_handle* h, echo_dac_stream* st, void* stream);  /* start a new sequence */
int echo_dac_stream_decode(echo_handle* h, echo_dac_stream* st, const float* z, const float* pca_components,
                           const float* pca_mean, float latent_scale, int T, float* audio, void* stream);
int echo_dac_stream_position(echo_handle* h, echo_dac_stream* st, int* latents_decoded);
int echo_dac_stream_destroy(echo_handle* h, echo_dac_stream* st);

/* ---- Fish S1-DAC encode: replaces inference.ae_encode (inference.py:219-224) / DAC.encode_zq (autoencoder.py:1080-1126)
 * audio: (B, 1, L) fp32 on the device, L a multiple of the frame length (enc hop * 2^num_upsample = 2048; the caller
 * right-pads with zeros as DAC.encode does). T = L / frame length.
 *   echo_dac_encode_zq : z_q (B, latent_dim, T) fp32; codes (B, 1 + n_codebooks, T) int32 (optional, may be NULL);
 *                        z_pre (B, T, latent_dim) fp32 = the quantizers' input, for parity tests (optional, may be NULL)
 *   echo_dac_encode    : PCA-projected, scaled latents (B, T, pca_dim) fp32 = ((z_q^T - mean) @ components^T) * scale
 * Needs the encoder / quantizer weights ("dac.encoder.*", "dac.quantizer.downsample|pre_module|*quantizer.*"). */
int echo_dac_encode_zq(echo_handle* h, const float* audio, int B, int L, float* zq, int32_t* codes, float* z_pre,
                       void* stream);
int echo_dac_encode(echo_handle* h, const float* audio, const float* pca_components, const float* pca_mean,
                    float latent_scale, int B, int L, float* latent, void* stream);

/* ---- host-buffer convenience (what a non-PyTorch host binds): copies in, runs, copies out, synchronises. -- */
int echo_sample_euler_host(echo_handle* h, const echo_sampler_args* a, const float* speaker_latent_f32_host,
                           const uint8_t* speaker_mask_host, int Ls, const int32_t* text_ids_host,
                           const uint8_t* text_mask_host, int Lt, int B, const float* noise_host,
                           float* x_out_host);

/* ---- per-launch CUDA-event timing for the roofline report (bench.py); off unless started -------------------- */
typedef struct echo_profile_report {   /* index 0: tcgen05 GEMM launches, 1: attention, 2: bandwidth-bound glue */
  int64_t launches[3];
  double ms[3];      /* summed CUDA-event durations on the launching stream */
  double flops[3];   /* algorithmic FLOPs (2*M*N*K per GEMM) */
  double bytes[3];   /* algorithmic bytes (glue kernels) */
} echo_profile_report;
int echo_profile_start(echo_handle* h);
int echo_profile_stop(echo_handle* h, echo_profile_report* out); /* synchronises the device */

/* ---- op-level entry points (used by the parity tests; same kernels the calls above launch) -------------- */
typedef struct echo_gemm_desc {
  const void* A; int64_t lda; int64_t a_batch_stride; /* bf16 [batches][M][lda] */
  const void* B; int64_t ldb; int64_t b_rows;          /* bf16 [b_rows][ldb]; K extent = taps*Kc */
  int M, N, Kc, batches, taps;
  int tap_shift[8];
  int epi; /* 0 generic, 1 swiglu, 2 qkv, 4 fused DAC ResidualUnit (see B1 below) */
  const float* bias; float scale;
  const float* gate; int rows_per_gate; int gate_ld;
  const float* resid; float* out_f32; int ld_f32;
  void* out_bf16; int ld_bf16;
  int act; const float* alpha; int col_mod;
  /* qkv */
  void* sec_out[4]; const float* sec_norm_w[4]; int sec_rope_heads[4]; int sec_sigmoid[4];
  int sec_width; const float* rope_cos; const float* rope_sin; int head_dim; int pos_period; int pos_offset;
  int pos_mult; float eps;
  int bn; /* 0 = auto */
  int cg; /* 0 = auto, 1 = one CTA per tile, 2 = CTA pair per 256-row tile (tcgen05 cta_group::2) */
  int reserved0;    /* must be 0 */
  long long* trace; /* optional device buffer, 16 clock64 stamps per CTA (kernel timeline for tuning); NULL normally */
  int split_k;      /* 0 = auto, 1 = off, n > 1 = n K-splits per tile (only when out_f32 == resid: atomic accumulate) */
  /* epi == 4, fused ResidualUnit (autoencoder.py:884-900), N == Kc == C in {96, 192}: A / B / taps / tap_shift / bias / alpha
     describe the dilated conv7 and the Snake behind it; B1 [C][ldb1] bf16 + ru_bias1 the 1 x 1 conv; resid / out_f32 the
     fp32 stream (may alias); out_bf16 = Snake(ru_alpha_out) of the new stream (must not alias A). */
  const void* B1; int64_t ldb1; const float* ru_bias1; const float* ru_alpha_out;
} echo_gemm_desc;
int echo_op_gemm(const echo_gemm_desc* d, void* stream);

typedef struct echo_attn_segment {
  const void* K; const void* V;   /* bf16, key j of batch b at K + (b*batch_stride + j*row_stride) elements */
  int64_t batch_stride; int64_t row_stride;
  int batch_mod;                  /* >0: K/V/mask are indexed by (b % batch_mod): CFG branches share one cache */
  int len;                        /* keys in this segment */
  const int32_t* eff_len;         /* optional device (b): keys >= eff_len[batch] are all invalid (tile skipping) */
  const uint8_t* mask;            /* (batch, mask_ld) bool per key, or NULL = all valid */
  int mask_ld; int mask_stride;   /* key j reads mask[b*mask_ld + j*mask_stride] */
  int pos_limit_mult;             /* >0: key j valid iff j*pos_limit_mult < pos_limit (latent prefix, model.py:243-244) */
  int pos_limit;
  int causal;                     /* 1: key j valid iff j <= q (+ window), q = query row + q_offset */
  int window;                     /* >0 with causal: also j > q - window */
  int q_offset;                   /* causal only: index of query row 0 on this segment's key axis (streaming decode: the
                                     keys are a growing cache, the queries its newest rows); 0 = plain self attention */
  float kv_scale;                 /* 0 or 1: none. Otherwise the segment behaves as if its K and V were multiplied by this
                                     factor (speaker_kv_scale, inference.py:408-414, 467-468): scores x factor, P V x factor --
                                     applied in the kernel, the cache itself is never rewritten */
} echo_attn_segment;
typedef struct echo_attn_desc {
  const void* Q; int64_t q_batch_stride; int64_t q_row_stride; /* bf16 (b, S, H, D) */
  const void* gate;  /* optional bf16 (b,S,H*D): output multiplied by it (already sigmoid-ed) */
  void* out;         /* bf16 (b, S, H*D) */
  int b, S, H, D;    /* D = 128 or 64 */
  float scale;       /* softmax scale, 1/sqrt(D) */
  int nseg; echo_attn_segment seg[4];
  long long* trace;  /* optional device buffer, 64 clock64 stamps per CTA (tcgen05 kernel timeline); NULL normally */
  /* Split-KV: when one CTA per (128 queries, head, batch row) leaves most SMs idle and every CTA walks a long key list
     (blockwise S = 160 against ~2700 keys: 32 CTAs x 43 tiles), the key tiles are divided over a thread-block cluster
     of nsplit CTAs that merge their partial (O, max, sum) through distributed shared memory in a fixed order
     (deterministic; no workspace, no atomics). 0 = auto, 1 = off, 2..8 = force. */
  int nsplit;
} echo_attn_desc;
int echo_op_attention(const echo_attn_desc* d, void* stream);

/* out[r,c] = bf16( x[r,c] * rsqrt(mean_c x[r,:]^2 + eps) * a[g(r),c] + c0[g(r),c] ),  g(r) = rows_per_group > 0 ?
 * (r / rows_per_group) * group_ld : 0; c0 may be NULL. Covers RMSNorm (model.py:86-104: a = weight) and the
 * LowRankAdaLN modulate (model.py:76-79: a = 1 + scale, c0 = shift). x fp32 (rows, W), a / c0 fp32. */
int echo_op_rmsnorm_affine(const float* x, void* out_bf16, const float* a, const float* c0, int rows, int W,
                           int rows_per_group, int64_t group_ld, float eps, void* stream);
/* x += dt * v', v' = CFG combine of the 3 branches in v (inference.py:495) when has_cfg, else v; optional temporal
 * score rescale (inference.py:416-424) with one_minus_t and ratio precomputed by the caller. x, v fp32. */
int echo_op_cfg_euler_update(float* x, const float* v, int64_t n_per_branch, int has_cfg, float cfg_scale_text,
                             float cfg_scale_speaker, int has_rescale, float one_minus_t, float ratio, float dt,
                             void* stream);

/* find_flattening_point (reference inference.py:288-296; the step after ae_decode in sample_pipeline, :345): *out_index
 * (device int32) = first i in [0, T) whose window of window_size rows of latent (T, C) fp32 -- zero padded past the
 * end -- has unbiased std < std_threshold and |mean - target_value| < 0.1, else T. One warp per window, no host sync
 * (the reference loops over up to 640 windows with two device syncs each). */
int echo_op_flattening_point(const float* latent, int T, int C, float target_value, int window_size,
                             float std_threshold, int32_t* out_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECHO_B200_H */
