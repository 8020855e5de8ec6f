// Bandwidth-bound glue kernels of the DiT step and the samplers (declarations). All take a stream, none syncs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace echo {

typedef __nv_bfloat16 bf16;

// X[r, :] = table[ids[r], :]                                   (TextEncoder embedding, model.py:420)
void embed_rows(const int32_t* ids, const bf16* table, float* X, int rows, int E, int vocab, cudaStream_t s);

// out[r, c] = x[r, c] * rsqrt(mean_c x^2 + eps) * a[g(r), c] + c0[g(r), c]     (model.py:76-79, 99-104)
// a / c0 are fp32; g(r) = rows_per_group > 0 ? (r / rows_per_group) * group_ld : 0. c0 may be null.
// Optional split-K planes (W == 2048 only): x[r, :] += sum_{k < nparts} parts[k * part_stride + r * W + :] is folded in
// first and written back to X (the residual-stream update the preceding wo / w2 GEMM left in planes, see GemmCall::part_ws).
void rmsnorm_affine(const float* X, bf16* out, const float* a, const float* c0, int rows, int W, int rows_per_group,
                    int64_t group_ld, float eps, cudaStream_t s, const float* parts = nullptr, int nparts = 0,
                    int64_t part_stride = 0);

// X[c * rows + r, n] = sum_k x[r, k] W[n, k] + bias[n]   for c < copies          (EchoDiT.in_proj, model.py:586)
void in_proj(const float* x, const bf16* W, const float* bias, float* X, int rows, int K, int D, int copies,
             cudaStream_t s);

// v[r, n] = sum_k (X[r,k] * rstd_r * wn[k]) Wout[n, k] + bias[n]                  (out_norm + out_proj, model.py:601-604)
void out_norm_proj(const float* X, const float* wn, const bf16* Wout, const float* bias, float* v, int rows, int D,
                   int Nout, float eps, cudaStream_t s);

// emb[j, :] = [cos(t_j f_i), sin(t_j f_i)] as bf16, optionally rounding t to bf16 first   (model.py:27-43)
void timestep_embed(const float* t, const float* freqs, bf16* emb, int n, int half, int round_t_bf16, cudaStream_t s);

// scond[p][j][c] = silu(cond[j][p*D + c])  (bf16), p in {shift, scale, gate}       (model.py:70-74)
void adaln_prep(const float* cond, bf16* scond, int n, int D, cudaStream_t s);
// mod[p][q][j][c] = f_p(up[p][q][j][c] + cond[j][p*D + c]);  f = id, +1, tanh          (model.py:72-81)
void adaln_finish(const float* up, const float* cond, float* mod, int n, int D, int Q, cudaStream_t s);

// Euler update with independent CFG and optional temporal score rescale              (inference.py:495, 416-424, 515)
// v holds 3 branches (cond, no-text, no-speaker) of (B*S*C) when has_cfg, else one.
void cfg_euler_update(float* x, const float* v, int64_t n_per_branch, int has_cfg, float s_text, float s_spk,
                      int has_rescale, float one_minus_t, float ratio, float dt, cudaStream_t s);

void scale_copy_f32(const float* src, float* dst, int64_t n, float sc, cudaStream_t s);
void cast_f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t s);

// eff[j] = 1 + last index i with mask[j*ld + i*stride] != 0 (0 if none), i < len
void mask_eff_len(const uint8_t* mask, int32_t* eff, int n, int len, int ld, int stride, cudaStream_t s);

// weight packing: dst[(r / blk) * blk_stride + blk_off + r % blk][c] = src[r][c]  (src fp32 or bf16, dst bf16 or fp32)
void pack_rows(const void* src, int src_is_bf16, void* dst, int dst_is_bf16, int64_t rows, int64_t cols, int64_t dst_ld,
               int64_t blk, int64_t blk_stride, int64_t blk_off, cudaStream_t s);

// *out = first i in [0, T) whose window of `window` rows (zero padded past T) of latent (T, C) fp32 is flat, else T
// (find_flattening_point, inference.py:288-296)
void flattening_point(const float* latent, int T, int C, int window, float target, float std_threshold, int32_t* out,
                      cudaStream_t s);

int64_t glue_launch_count();

}  // namespace echo
