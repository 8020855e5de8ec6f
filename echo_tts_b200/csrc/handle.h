// Internal state behind the opaque echo_handle: packed weights, tables and a growable device workspace.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "echo_b200.h"
#include "errors.h"

namespace echo {

typedef __nv_bfloat16 bf16;

#define ECHO_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return ECHO_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define ECHO_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != ECHO_OK) return _rc; \
  } while (0)

struct EncoderLayerW {
  bf16 *wqkvg, *wo, *w13, *w2;
  float *q_norm, *k_norm, *attn_norm, *mlp_norm;
};
struct EncoderW {
  int E = 0, heads = 0, inter = 0, layers = 0;
  bf16* embed = nullptr;      // text only (vocab, E)
  bf16* in_proj_w = nullptr;  // speaker / latent (E, patch*latent)
  float* in_proj_b = nullptr;
  float* final_norm = nullptr;
  std::vector<EncoderLayerW> L;
};
struct BlockW {
  bf16 *wqkvg, *wo, *w13, *w2, *wkv_text, *wkv_speaker, *wkv_latent;
  float *q_norm, *k_norm;
};

struct RawTensor {  // DAC tensors are stashed as fp32 until finalize (weight-norm folding needs g and v together)
  float* p = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

struct DacConvW {   // one (transposed) conv lowered to a tap GEMM
  bf16* w = nullptr;  // [N][taps*Cin]
  float* bias = nullptr;
  int cin = 0, n = 0, taps = 0;
};
struct DacPostLayerW {
  bf16 *wqkv, *wo, *w13, *w2;
  float *attn_norm, *ffn_norm, *attn_gamma, *ffn_gamma;
};
struct DacUpW {
  DacConvW convt;  // ConvTranspose k2 s2 as a 1-tap GEMM with N = 2*C
  float *dw_w, *dw_b, *ln_w, *ln_b, *b1, *b2, *gamma;
  bf16 *w1, *w2;
};
struct DacResUnitW {
  float *alpha1, *alpha2;
  DacConvW conv7, conv1;
};
struct DacStageW {
  float* alpha_in;
  DacConvW convt;
  int stride = 0, cin = 0, cout = 0;
  DacResUnitW ru[3];
};

struct DacEncBlockW {  // EncoderBlock (autoencoder.py:839-876): 3 ResidualUnits at cin, Snake, strided conv cin -> cout
  DacResUnitW ru[3];
  float* alpha_out = nullptr;
  DacConvW down;  // k = 2 * stride, lowered to a 2-tap GEMM over the (T / stride, stride * cin) view
  int stride = 0, cin = 0, cout = 0;
};
struct DacVqW {  // VectorQuantize (autoencoder.py:117-157), everything fp32
  float* in_w = nullptr;     // [codebook_dim][C]   weight-norm folded in_proj
  float* in_b = nullptr;     // [codebook_dim]
  float* cb_norm = nullptr;  // [size][codebook_dim] L2-normalised codebook
  float* cb_sq = nullptr;    // [size]               |normalised code|^2 (the reference keeps this term)
  float* out_table = nullptr;  // [size][C]          out_proj(codebook[i]) + bias
  int size = 0;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace echo

struct echo_handle {
  int device = 0;
  int num_sms = 148;
  std::vector<void*> owned;  // every cudaMalloc'ed weight block
  std::map<std::string, echo::DevBuf> ws;

  // ---- EchoDiT
  bool dit_configured = false, dit_ready = false, has_latent = false;
  echo_dit_config cfg{};
  echo::EncoderW enc[3];  // text, speaker, latent
  std::vector<echo::BlockW> blk;
  echo::bf16 *ada_down = nullptr, *ada_up = nullptr;
  float* ada_up_bias = nullptr;
  echo::bf16 *cond_w0 = nullptr, *cond_w2 = nullptr, *cond_w4 = nullptr;
  echo::bf16 *in_proj_w = nullptr, *out_proj_w = nullptr;
  float *in_proj_b = nullptr, *out_norm = nullptr, *out_proj_b = nullptr;
  float *rope_cos = nullptr, *rope_sin = nullptr, *temb_freqs = nullptr;
  int rope_positions = 0;
  std::set<std::string> got;

  // ---- DAC
  bool dac_configured = false, dac_ready = false;
  echo_dac_config dcfg{};
  std::map<std::string, echo::RawTensor> dac_raw;
  std::vector<echo::DacPostLayerW> post;
  float* post_final_norm = nullptr;
  float *dac_rope_cos = nullptr, *dac_rope_sin = nullptr;
  std::vector<echo::DacUpW> up;
  echo::DacConvW dec_conv0;
  std::vector<echo::DacStageW> stage;
  float* final_alpha = nullptr;
  float* final_w = nullptr;  // [7][C] fp32
  float final_b = 0.f;
  std::map<const float*, float*> dac_alpha_inv;  // snake alpha -> 1 / (alpha + 1e-9)
  // ---- DAC encode path (optional)
  bool dac_enc_ready = false;
  float *enc_conv0_w = nullptr, *enc_conv0_b = nullptr;  // first conv, Cin = 1: [7][enc_dim] fp32
  std::vector<echo::DacEncBlockW> enc_blk;
  std::vector<echo::DacPostLayerW> enc_tf;
  float* enc_tf_norm = nullptr;
  float* enc_alpha_out = nullptr;
  echo::DacConvW enc_conv_out;  // k = 3
  std::vector<echo::DacUpW> down;  // quantizer.downsample: `convt` holds the k2 s2 conv as a 1-tap GEMM
  std::vector<echo::DacPostLayerW> pre;
  float* pre_final_norm = nullptr;
  std::vector<echo::DacVqW> vq;  // [0] semantic, [1 ..] residual
  std::set<echo_dac_stream*> dac_streams;  // live streaming-decode states (freed by echo_destroy if the caller did not)

  // ---- serialisation of the public entry points (see echo::HandleScope)
  std::recursive_mutex mu;
  int depth = 0;
  bool used = false;
  cudaStream_t last_stream = nullptr;
  cudaEvent_t order_ev = nullptr;
  // side stream of the samplers: the speaker KV cache is built next to the text KV cache (both are a hundred small,
  // latency-bound launches), forked from and joined back into the caller's stream with events
  cudaStream_t side_stream = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;

  void* wsget(const char* name, size_t bytes, cudaStream_t s);
  void* dalloc(size_t bytes);
};

namespace echo {
// Every compute entry point of the C ABI opens one of these. The named workspace buffers (activations, KV scratch,
// AdaLN tables) are shared by all calls on a handle, so (1) host threads are serialised by a per-handle mutex and
// (2) when a call arrives on a different stream than the previous one, the new stream first waits for everything the
// handle enqueued on the old stream (one event, no host synchronisation). Calls on ONE stream -- the normal case --
// pay a mutex lock and a pointer compare. Contract for callers: a stream handed to the library must stay alive until
// the next call on the same handle has been issued.
struct HandleScope {
  echo_handle* h;
  HandleScope(echo_handle* hh, cudaStream_t s) : h(hh) {
    h->mu.lock();
    if (h->depth++ == 0) {
      if (h->used && h->last_stream != s) {
        if (!h->order_ev) cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming);
        if (h->order_ev && cudaEventRecord(h->order_ev, h->last_stream) == cudaSuccess) cudaStreamWaitEvent(s, h->order_ev, 0);
        else cudaGetLastError();  // the old stream is gone: everything on it has completed or was abandoned
      }
      h->last_stream = s;
      h->used = true;
    }
  }
  ~HandleScope() {
    --h->depth;
    h->mu.unlock();
  }
  HandleScope(const HandleScope&) = delete;
  HandleScope& operator=(const HandleScope&) = delete;
};

// dac.cu
int dac_set_weight(echo_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype,
                   cudaStream_t s);
}  // namespace echo
