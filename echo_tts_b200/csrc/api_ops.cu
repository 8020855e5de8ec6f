// Op-level C-ABI entry points (include/echo_b200.h): thin translations from the C descriptors to the internal
// launch structs. The parity tests drive the very same kernels the model-level calls use through these.
#include <cstdio>
#include <cstring>

#include "attention.h"
#include "echo_b200.h"
#include "errors.h"
#include "gemm.h"
#include "glue.h"

using namespace echo;

namespace echo {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace echo

extern "C" const char* echo_last_error(void) { return echo::g_err; }

extern "C" int echo_op_gemm(const echo_gemm_desc* d, void* stream) {
  if (!d) { set_error("echo_op_gemm: null descriptor"); return ECHO_ERR_ARG; }
  GemmCall c;
  std::memset(&c, 0, sizeof(c));
  c.A = static_cast<const bf16*>(d->A);
  c.lda = d->lda;
  c.a_batch_stride = d->a_batch_stride;
  c.B = static_cast<const bf16*>(d->B);
  c.ldb = d->ldb;
  c.b_rows = d->b_rows;
  c.bn = d->bn;
  c.cg = d->cg;
  c.split_k = d->split_k;
  c.B1 = static_cast<const bf16*>(d->B1);
  c.ldb1 = d->ldb1;
  GemmParams& p = c.p;
  p.M = d->M; p.N = d->N; p.Kc = d->Kc; p.batches = d->batches; p.taps = d->taps;
  for (int i = 0; i < 8; ++i) p.tap_shift[i] = d->tap_shift[i];
  p.a_batch_div = 1;
  p.epi = d->epi;
  p.bias = d->bias; p.scale = d->scale;
  p.gate = d->gate; p.rows_per_gate = d->rows_per_gate; p.gate_ld = d->gate_ld;
  p.resid = d->resid; p.out_f32 = d->out_f32; p.ld_f32 = d->ld_f32;
  p.out_bf16 = static_cast<bf16*>(d->out_bf16); p.ld_bf16 = d->ld_bf16;
  p.act = d->act; p.alpha = d->alpha; p.col_mod = d->col_mod;
  for (int i = 0; i < 4; ++i) {
    p.sec[i].out = static_cast<bf16*>(d->sec_out[i]);
    p.sec[i].norm_w = d->sec_norm_w[i];
    p.sec[i].rope_heads = d->sec_rope_heads[i];
    p.sec[i].sigmoid = d->sec_sigmoid[i];
  }
  p.ru_bias1 = d->ru_bias1; p.ru_alpha_out = d->ru_alpha_out; p.ru_alpha_out_inv = nullptr;
  p.sec_width = d->sec_width; p.rope_cos = d->rope_cos; p.rope_sin = d->rope_sin;
  p.head_dim = d->head_dim ? d->head_dim : 128;
  p.pos_period = d->pos_period; p.pos_offset = d->pos_offset; p.pos_mult = d->pos_mult ? d->pos_mult : 1;
  p.eps = d->eps;
  p.trace = d->trace;
  cudaError_t e = gemm_launch(c, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    set_error("echo_op_gemm: %s (M=%d N=%d Kc=%d taps=%d epi=%d)", cudaGetErrorString(e), d->M, d->N, d->Kc, d->taps,
              d->epi);
    return e == cudaErrorInvalidValue ? ECHO_ERR_ARG : ECHO_ERR_CUDA;
  }
  return ECHO_OK;
}

extern "C" int echo_op_attention(const echo_attn_desc* d, void* stream) {
  if (!d) { set_error("echo_op_attention: null descriptor"); return ECHO_ERR_ARG; }
  cudaError_t e = attention_launch(*d, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    set_error("echo_op_attention: %s (b=%d S=%d H=%d D=%d nseg=%d)", cudaGetErrorString(e), d->b, d->S, d->H, d->D,
              d->nseg);
    return e == cudaErrorInvalidValue ? ECHO_ERR_ARG : ECHO_ERR_CUDA;
  }
  return ECHO_OK;
}

extern "C" int echo_op_rmsnorm_affine(const float* x, void* out_bf16, const float* a, const float* c0, int rows, int W,
                                      int rows_per_group, int64_t group_ld, float eps, void* stream) {
  if (!x || !out_bf16 || !a || rows <= 0 || W <= 0 || W % 4 != 0 || W > 4096) {
    set_error("echo_op_rmsnorm_affine: bad argument (rows=%d W=%d)", rows, W);
    return ECHO_ERR_ARG;
  }
  rmsnorm_affine(x, static_cast<bf16*>(out_bf16), a, c0, rows, W, rows_per_group, group_ld, eps,
                 static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("echo_op_rmsnorm_affine: %s", cudaGetErrorString(e)); return ECHO_ERR_CUDA; }
  return ECHO_OK;
}

extern "C" int echo_op_cfg_euler_update(float* x, const float* v, int64_t n_per_branch, int has_cfg, float cfg_scale_text,
                                        float cfg_scale_speaker, int has_rescale, float one_minus_t, float ratio, float dt,
                                        void* stream) {
  if (!x || !v || n_per_branch <= 0) { set_error("echo_op_cfg_euler_update: bad argument"); return ECHO_ERR_ARG; }
  cfg_euler_update(x, v, n_per_branch, has_cfg, cfg_scale_text, cfg_scale_speaker, has_rescale, one_minus_t, ratio, dt,
                   static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("echo_op_cfg_euler_update: %s", cudaGetErrorString(e)); return ECHO_ERR_CUDA; }
  return ECHO_OK;
}

extern "C" int echo_op_flattening_point(const float* latent, int T, int C, float target_value, int window_size,
                                        float std_threshold, int32_t* out_index, void* stream) {
  if (!latent || !out_index || T < 0 || C <= 0 || window_size <= 0) { set_error("echo_op_flattening_point: bad argument"); return ECHO_ERR_ARG; }
  flattening_point(latent, T, C, window_size, target_value, std_threshold, out_index, static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("echo_op_flattening_point: %s", cudaGetErrorString(e)); return ECHO_ERR_CUDA; }
  return ECHO_OK;
}

extern "C" int echo_set_deterministic(int on) {
  gemm_set_deterministic(on);
  return ECHO_OK;
}
