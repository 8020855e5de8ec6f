// Fish S1-DAC decode path on B200: ae_decode (reference inference.py:226-229) -> DAC.decode_zq
// (autoencoder.py:1128-1132) = quantizer.post_module -> quantizer.upsample -> Decoder.
//
// Activations are kept TIME-MAJOR (B, T, C). Then
//   * every Linear of the post_module transformer and of the ConvNeXt blocks,
//   * every causal (dilated) Conv1d           : y[t] = sum_j W_j x[t - (k-1-j) d]      (autoencoder.py:285-289)
//   * every causal ConvTranspose1d (k = 2s)   : y[q s + r] = W_r x[q] + W_{r+s} x[q-1]  (autoencoder.py:310-316)
// is one launch of the tcgen05 tap-GEMM (gemm_tc.cuh): the taps are row-shifted TMA loads of the same activation
// matrix, TMA's zero fill is the causal padding, the polyphase transposed conv is a 2-tap GEMM with N = s*Cout whose
// row-major output IS the upsampled (T*s, Cout) signal. Bias, Snake, GELU, LayerScale, residual adds and the bf16
// re-quantisation for the next layer all happen in the GEMM epilogue. Weight-norm is folded at load time.
#include <cmath>
#include <cstring>

#include "attention.h"
#include "counters.h"
#include "launch.h"
#include "gemm.h"
#include "glue.h"
#include "handle.h"

using namespace echo;

namespace {

// ------------------------------------------------------------------------------------------------ kernels
__global__ void slice_norm_kernel(const float* __restrict__ v, float* __restrict__ scale, const float* __restrict__ g,
                                  int64_t slice) {
  // scale[i] = g[i] / ||v[i, ...]||   (weight_norm over every dim but 0; autoencoder.py:90-94)
  __shared__ float sh[32];
  const int i = blockIdx.x;
  float ss = 0.f;
  for (int64_t k = threadIdx.x; k < slice; k += blockDim.x) {
    const float t = v[(size_t)i * slice + k];
    ss += t * t;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) scale[i] = g[i] / sqrtf(t);
  }
}

// Conv1d weight (Cout, Cin, k) -> [Cout][k][Cin] bf16 (tap j multiplies x[t - (k-1-j) d])
__global__ void pack_conv_kernel(const float* __restrict__ v, const float* __restrict__ scale, bf16* __restrict__ out,
                                 int Cout, int Cin, int k) {
  const int64_t total = (int64_t)Cout * Cin * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int j = (int)((i / Cin) % k);
    const int co = (int)(i / ((int64_t)Cin * k));
    const float w = v[((size_t)co * Cin + ci) * k + j] * (scale ? scale[co] : 1.f);
    out[i] = __float2bfloat16_rn(w);
  }
}

// ConvTranspose1d weight (Cin, Cout, k = taps*s) -> [r*Cout + co][tap*Cin + ci] = v[ci, co, r + tap*s] * scale[ci]
__global__ void pack_convt_kernel(const float* __restrict__ v, const float* __restrict__ scale, bf16* __restrict__ out,
                                  int Cin, int Cout, int s, int taps) {
  const int k = taps * s;
  const int64_t total = (int64_t)s * Cout * taps * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int tap = (int)((i / Cin) % taps);
    const int n = (int)(i / ((int64_t)Cin * taps));
    const int co = n % Cout, r = n / Cout;
    const float w = v[((size_t)ci * Cout + co) * k + r + tap * s] * (scale ? scale[ci] : 1.f);
    out[i] = __float2bfloat16_rn(w);
  }
}

__global__ void alpha_inv_kernel(const float* __restrict__ a, float* __restrict__ inv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[i] = 1.f / (a[i] + 1e-9f);
}

// final conv (1, C, 7) -> [7][C] fp32
__global__ void pack_final_kernel(const float* __restrict__ v, const float* __restrict__ scale, float* __restrict__ out,
                                  int C, int k) {
  for (int i = threadIdx.x; i < C * k; i += blockDim.x) {
    const int c = i % C, j = i / C;
    out[i] = v[(size_t)c * k + j] * scale[0];
  }
}

// zq[r, n] = sum_k (z[r, k] / scale) * comps[k, n] + mean[n]     (inference.py:228), fp32
__global__ void pca_unproject_kernel(const float* __restrict__ z, const float* __restrict__ comps,
                                     const float* __restrict__ mean, float inv_is_div_scale, float* __restrict__ X,
                                     int K, int N) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sz[];
  const int r = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) sz[k] = z[(size_t)r * K + k] / inv_is_div_scale;
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(sz[k], comps[(size_t)k * N + n], acc);
    X[(size_t)r * N + n] = acc + mean[n];
  }
}

// (B, C, T) fp32 -> (B, T, C) fp32
__global__ void transpose_ct_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? in[((size_t)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (c < C && t < T) out[((size_t)b * T + t) * C + c] = tile[threadIdx.x][i];
  }
}

// ConvNeXt front: depthwise causal conv k=7 + LayerNorm(eps 1e-6) -> bf16      (autoencoder.py:362-364)
// one block per (b, t) row, C <= 4096, 256 threads
__global__ void __launch_bounds__(256) dwconv_ln_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                        const float* __restrict__ wb, const float* __restrict__ lnw,
                                                        const float* __restrict__ lnb, bf16* __restrict__ out, int T,
                                                        int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[32];
  __shared__ float stat[2];
  const int row = blockIdx.x;
  const int t = row % T;
  float y[16];
  float s1 = 0.f;
  int cnt = 0;
  for (int c = threadIdx.x; c < C; c += 256, ++cnt) {
    float acc = wb[c];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int tt = t - 6 + j;
      if (tt >= 0) acc = fmaf(w[c * 7 + j], z[((size_t)row - 6 + j) * C + c], acc);
    }
    y[cnt] = acc;
    s1 += acc;
  }
  // mean
  float v = s1;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float tt = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, o);
    if (threadIdx.x == 0) stat[0] = tt / (float)C;
  }
  __syncthreads();
  const float mean = stat[0];
  float s2 = 0.f;
  cnt = 0;
  for (int c = threadIdx.x; c < C; c += 256, ++cnt) { const float d = y[cnt] - mean; s2 += d * d; }
  v = s2;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float tt = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, o);
    if (threadIdx.x == 0) stat[1] = rsqrtf(tt / (float)C + 1e-6f);
  }
  __syncthreads();
  const float rstd = stat[1];
  cnt = 0;
  for (int c = threadIdx.x; c < C; c += 256, ++cnt)
    out[(size_t)row * C + c] = __float2bfloat16_rn((y[cnt] - mean) * rstd * lnw[c] + lnb[c]);
}

// final: audio[b, t] = tanh(bias + sum_j sum_c w[j][c] * sx[b, t - 6 + j, c])     (autoencoder.py:994)
__global__ void __launch_bounds__(256) final_conv_tanh_kernel(const bf16* __restrict__ sx, const float* __restrict__ w,
                                                              float bias, float* __restrict__ audio, int T, int C) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sw[];  // [7][C]
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const bf16* base = sx + (size_t)b * T * C;
  float acc = bias;
  for (int j = 0; j < 7; ++j) {
    const int tt = t - 6 + j;
    if (tt < 0) continue;
    const uint4* rp = reinterpret_cast<const uint4*>(base + (size_t)tt * C);
    const float* wj = sw + j * C;
    for (int c8 = 0; c8 < C / 8; ++c8) {
      const uint4 u = __ldg(rp + c8);
      const uint32_t* uu = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uu[q]));
        acc = fmaf(f.x, wj[c8 * 8 + 2 * q], acc);
        acc = fmaf(f.y, wj[c8 * 8 + 2 * q + 1], acc);
      }
    }
  }
  audio[(size_t)b * T + t] = tanhf(acc);
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}

bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

}  // namespace

// ------------------------------------------------------------------------------------------------ weights
namespace echo {

int dac_set_weight(echo_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype,
                   cudaStream_t s) {
  if (!h->dac_configured) { set_error("echo_set_weight(dac.*): call echo_dac_configure first"); return ECHO_ERR_STATE; }
  RawTensor& r = h->dac_raw[key];
  if (r.p) { cudaFree(r.p); r.p = nullptr; }
  r.shape.assign(shape, shape + ndim);
  r.numel = 1;
  for (int i = 0; i < ndim; ++i) r.numel *= shape[i];
  ECHO_CUDA(cudaMalloc(&r.p, (size_t)r.numel * 4));
  pack_rows(data, dtype == ECHO_DTYPE_BF16, r.p, 0, 1, r.numel, r.numel, 1, 0, 0, s);
  ECHO_CUDA(cudaGetLastError());
  return ECHO_OK;
}

}  // namespace echo

extern "C" int echo_dac_configure(echo_handle* h, const echo_dac_config* c) {
  if (!h || !c) { set_error("echo_dac_configure: null argument"); return ECHO_ERR_ARG; }
  if (h->dac_configured) { set_error("echo_dac_configure: already configured"); return ECHO_ERR_STATE; }
  const int C = c->latent_dim;
  if (C % 256 || c->post_heads <= 0 || C / c->post_heads != 64 || c->post_intermediate % 128 || C > 4096 ||
      c->num_rates < 1 || c->num_rates > 8 || c->pca_dim > 512) {
    set_error("echo_dac_configure: unsupported dims (latent_dim %% 256, head_dim 64, intermediate %% 128)");
    return ECHO_ERR_ARG;
  }
  int ch = c->decoder_dim;
  for (int i = 0; i < c->num_rates; ++i) {
    if (ch % 2 || (ch / 2) % 32) { set_error("decoder channels must stay multiples of 32"); return ECHO_ERR_ARG; }
    ch /= 2;
  }
  h->dcfg = *c;
  h->dac_configured = true;
  return ECHO_OK;
}

namespace {

struct Packer {
  echo_handle* h;
  cudaStream_t s;
  int rc = ECHO_OK;

  const RawTensor* get(const std::string& k) {
    auto it = h->dac_raw.find(k);
    if (it == h->dac_raw.end()) {
      if (rc == ECHO_OK) { set_error("echo_dac_finalize: missing weight 'dac.%s'", k.c_str()); rc = ECHO_ERR_STATE; }
      return nullptr;
    }
    return &it->second;
  }
  float* vecf(const std::string& k, int64_t n) {  // fp32 copy
    const RawTensor* r = get(k);
    if (!r) return nullptr;
    if (r->numel != n) { if (rc == ECHO_OK) { set_error("dac.%s: %lld elements, expected %lld", k.c_str(), (long long)r->numel, (long long)n); rc = ECHO_ERR_ARG; } return nullptr; }
    float* d = (float*)h->dalloc((size_t)n * 4);
    cudaMemcpyAsync(d, r->p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
    return d;
  }
  float* alpha(const std::string& k, int64_t n) {  // snake alpha + its reciprocal
    float* a = vecf(k, n);
    if (!a) return nullptr;
    float* inv = (float*)h->dalloc((size_t)n * 4);
    alpha_inv_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(a, inv, (int)n);
    h->dac_alpha_inv[a] = inv;
    return a;
  }
  bf16* matb(const std::string& k, int64_t rows, int64_t cols, bf16* dst = nullptr, int64_t blk = 0, int64_t blk_stride = 0,
             int64_t blk_off = 0) {
    const RawTensor* r = get(k);
    if (!r) return nullptr;
    if (r->numel != rows * cols) { if (rc == ECHO_OK) { set_error("dac.%s: bad size", k.c_str()); rc = ECHO_ERR_ARG; } return nullptr; }
    if (!dst) dst = (bf16*)h->dalloc((size_t)rows * cols * 2);
    pack_rows(r->p, 0, dst, 1, rows, cols, cols, blk ? blk : rows, blk_stride, blk_off, s);
    return dst;
  }
  // weight-normed (or plain when wn=false) conv -> tap GEMM operand
  DacConvW conv(const std::string& p, int cout, int cin, int k, bool wn) {
    DacConvW w;
    w.cin = cin; w.n = cout; w.taps = k;
    const RawTensor* v = get(p + (wn ? ".conv.parametrizations.weight.original1" : ".conv.weight"));
    if (!v) return w;
    if (v->numel != (int64_t)cout * cin * k) { if (rc == ECHO_OK) { set_error("dac.%s: bad conv size", p.c_str()); rc = ECHO_ERR_ARG; } return w; }
    float* scale = nullptr;
    if (wn) {
      const RawTensor* g = get(p + ".conv.parametrizations.weight.original0");
      if (!g) return w;
      scale = (float*)h->dalloc((size_t)cout * 4);
      slice_norm_kernel<<<cout, 256, 0, s>>>(v->p, scale, g->p, (int64_t)cin * k);
    }
    w.w = (bf16*)h->dalloc((size_t)cout * cin * k * 2);
    pack_conv_kernel<<<grid_for((int64_t)cout * cin * k), 256, 0, s>>>(v->p, scale, w.w, cout, cin, k);
    w.bias = vecf(p + ".conv.bias", cout);
    return w;
  }
  DacConvW convt(const std::string& p, int cin, int cout, int stride, int taps, bool wn) {
    DacConvW w;
    w.cin = cin; w.n = stride * cout; w.taps = taps;
    const int k = taps * stride;
    const RawTensor* v = get(p + (wn ? ".conv.parametrizations.weight.original1" : ".conv.weight"));
    if (!v) return w;
    if (v->numel != (int64_t)cin * cout * k) { if (rc == ECHO_OK) { set_error("dac.%s: bad convT size", p.c_str()); rc = ECHO_ERR_ARG; } return w; }
    float* scale = nullptr;
    if (wn) {
      const RawTensor* g = get(p + ".conv.parametrizations.weight.original0");
      if (!g) return w;
      scale = (float*)h->dalloc((size_t)cin * 4);
      slice_norm_kernel<<<cin, 256, 0, s>>>(v->p, scale, g->p, (int64_t)cout * k);
    }
    w.w = (bf16*)h->dalloc((size_t)stride * cout * taps * cin * 2);
    pack_convt_kernel<<<grid_for((int64_t)stride * cout * taps * cin), 256, 0, s>>>(v->p, scale, w.w, cin, cout, stride, taps);
    w.bias = vecf(p + ".conv.bias", cout);
    return w;
  }
};

}  // namespace

extern "C" int echo_dac_finalize(echo_handle* h, void* stream) {
  if (!h || !h->dac_configured) { set_error("echo_dac_finalize: not configured"); return ECHO_ERR_STATE; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ECHO_CUDA(cudaSetDevice(h->device));
  const echo_dac_config& c = h->dcfg;
  const int C = c.latent_dim, I = c.post_intermediate;
  Packer pk{h, s};
  h->post.resize(c.post_layers);
  for (int i = 0; i < c.post_layers; ++i) {
    const std::string p = "quantizer.post_module.layers." + std::to_string(i);
    DacPostLayerW& l = h->post[i];
    l.wqkv = pk.matb(p + ".attention.wqkv.weight", 3 * C, C);
    l.wo = pk.matb(p + ".attention.wo.weight", C, C);
    l.w13 = (bf16*)h->dalloc((size_t)2 * I * C * 2);
    pk.matb(p + ".feed_forward.w1.weight", I, C, l.w13, 128, 256, 0);
    pk.matb(p + ".feed_forward.w3.weight", I, C, l.w13, 128, 256, 128);
    l.w2 = pk.matb(p + ".feed_forward.w2.weight", C, I);
    l.attn_norm = pk.vecf(p + ".attention_norm.weight", C);
    l.ffn_norm = pk.vecf(p + ".ffn_norm.weight", C);
    l.attn_gamma = pk.vecf(p + ".attention_layer_scale.gamma", C);
    l.ffn_gamma = pk.vecf(p + ".ffn_layer_scale.gamma", C);
  }
  h->post_final_norm = pk.vecf("quantizer.post_module.norm.weight", C);
  h->up.resize(c.num_upsample);
  for (int i = 0; i < c.num_upsample; ++i) {
    const std::string p = "quantizer.upsample." + std::to_string(i);
    DacUpW& u = h->up[i];
    u.convt = pk.convt(p + ".0", C, C, 2, 1, false);
    u.dw_w = pk.vecf(p + ".1.dwconv.conv.weight", (int64_t)C * 7);
    u.dw_b = pk.vecf(p + ".1.dwconv.conv.bias", C);
    u.ln_w = pk.vecf(p + ".1.norm.weight", C);
    u.ln_b = pk.vecf(p + ".1.norm.bias", C);
    u.w1 = pk.matb(p + ".1.pwconv1.weight", 4 * C, C);
    u.b1 = pk.vecf(p + ".1.pwconv1.bias", 4 * C);
    u.w2 = pk.matb(p + ".1.pwconv2.weight", C, 4 * C);
    u.b2 = pk.vecf(p + ".1.pwconv2.bias", C);
    u.gamma = pk.vecf(p + ".1.gamma", C);
  }
  int ch = c.decoder_dim;
  h->dec_conv0 = pk.conv("decoder.model.0", ch, C, 7, true);
  h->stage.resize(c.num_rates);
  for (int b = 0; b < c.num_rates; ++b) {
    const std::string p = "decoder.model." + std::to_string(b + 1) + ".block";
    DacStageW& st = h->stage[b];
    st.stride = c.rates[b]; st.cin = ch; st.cout = ch / 2;
    st.alpha_in = pk.alpha(p + ".0.alpha", st.cin);
    st.convt = pk.convt(p + ".1", st.cin, st.cout, st.stride, 2, true);
    for (int u = 0; u < 3; ++u) {
      const std::string q = p + "." + std::to_string(u + 2) + ".block";
      st.ru[u].alpha1 = pk.alpha(q + ".0.alpha", st.cout);
      st.ru[u].conv7 = pk.conv(q + ".1", st.cout, st.cout, 7, true);
      st.ru[u].alpha2 = pk.alpha(q + ".2.alpha", st.cout);
      st.ru[u].conv1 = pk.conv(q + ".3", st.cout, st.cout, 1, true);
    }
    ch /= 2;
  }
  const int n = c.num_rates;
  h->final_alpha = pk.alpha("decoder.model." + std::to_string(n + 1) + ".alpha", ch);
  {
    const std::string p = "decoder.model." + std::to_string(n + 2);
    const RawTensor* v = pk.get(p + ".conv.parametrizations.weight.original1");
    const RawTensor* g = pk.get(p + ".conv.parametrizations.weight.original0");
    const RawTensor* b = pk.get(p + ".conv.bias");
    if (v && g && b) {
      float* scale = (float*)h->dalloc(4);
      slice_norm_kernel<<<1, 256, 0, s>>>(v->p, scale, g->p, (int64_t)ch * 7);
      h->final_w = (float*)h->dalloc((size_t)7 * ch * 4);
      pack_final_kernel<<<1, 256, 0, s>>>(v->p, scale, h->final_w, ch, 7);
      ECHO_CUDA(cudaMemcpyAsync(&h->final_b, b->p, 4, cudaMemcpyDeviceToHost, s));
    }
  }
  if (pk.rc != ECHO_OK) return pk.rc;
  // RoPE cache of the post_module: stored in bfloat16 by the reference (autoencoder.py:805-813), used in fp32 math
  const int P = 4096, half = 32;
  std::vector<float> cs((size_t)P * half), sn((size_t)P * half);
  for (int i = 0; i < half; ++i) {
    const float inv = 1.0f / powf(10000.0f, (float)(2 * i) / 64.0f);
    for (int p = 0; p < P; ++p) {
      const float ang = (float)p * inv;
      cs[(size_t)p * half + i] = __bfloat162float(__float2bfloat16_rn(cosf(ang)));
      sn[(size_t)p * half + i] = __bfloat162float(__float2bfloat16_rn(sinf(ang)));
    }
  }
  h->dac_rope_cos = (float*)h->dalloc(cs.size() * 4);
  h->dac_rope_sin = (float*)h->dalloc(sn.size() * 4);
  ECHO_CUDA(cudaMemcpyAsync(h->dac_rope_cos, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice, s));
  ECHO_CUDA(cudaMemcpyAsync(h->dac_rope_sin, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice, s));
  ECHO_CUDA(cudaStreamSynchronize(s));
  ECHO_CUDA(cudaGetLastError());
  for (auto& kv : h->dac_raw) if (kv.second.p) { cudaFree(kv.second.p); kv.second.p = nullptr; }
  h->dac_raw.clear();
  h->dac_ready = true;
  return ECHO_OK;
}

// ------------------------------------------------------------------------------------------------ decode
namespace {

#define DAC_GEMM(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = gemm_launch((call), s);                                                   \
    if (_e != cudaSuccess) {                                                                   \
      set_error("%s:%d dac gemm: %s", __FILE__, __LINE__, cudaGetErrorString(_e));             \
      return ECHO_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

GemmCall base_gemm(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, int batches, int M, int N, int Kc, int taps) {
  GemmCall c;
  std::memset(&c, 0, sizeof(c));
  c.A = A; c.lda = lda; c.a_batch_stride = (int64_t)M * lda; c.B = W; c.ldb = ldw;
  c.p.M = M; c.p.N = N; c.p.Kc = Kc; c.p.batches = batches; c.p.taps = taps; c.p.a_batch_div = 1;
  c.p.epi = EPI_GENERIC; c.p.scale = 1.f; c.p.pos_period = 1; c.p.pos_mult = 1; c.p.head_dim = 128;
  return c;
}

// causal conv (k taps, dilation d) over (B, T, Cin) -> epilogue outputs
GemmCall conv_gemm(const DacConvW& w, const bf16* A, int B, int T, int dil) {
  GemmCall c = base_gemm(A, w.cin, w.w, (int64_t)w.taps * w.cin, B, T, w.n, w.cin, w.taps);
  for (int j = 0; j < w.taps; ++j) c.p.tap_shift[j] = -(w.taps - 1 - j) * dil;
  c.p.bias = w.bias; c.p.col_mod = w.n;
  return c;
}

// polyphase causal transposed conv: taps = 1 (k == s) or 2 (k == 2s)
GemmCall convt_gemm(const DacConvW& w, const bf16* A, int B, int Tin, int cout) {
  GemmCall c = base_gemm(A, w.cin, w.w, (int64_t)w.taps * w.cin, B, Tin, w.n, w.cin, w.taps);
  c.p.tap_shift[0] = 0;
  c.p.tap_shift[1] = -1;
  c.p.bias = w.bias; c.p.col_mod = cout;
  return c;
}

int dac_run(echo_handle* h, float* X /* (B*T, C) fp32 latent, time-major */, int B, int T, float* audio, cudaStream_t s) {
  const echo_dac_config& c = h->dcfg;
  const int C = c.latent_dim, I = c.post_intermediate, H = c.post_heads, rows = B * T;
  if (T > 4096) { set_error("dac: T=%d exceeds the post_module block size 4096", T); return ECHO_ERR_ARG; }
  // sizes of the largest activations
  int64_t max_elems = (int64_t)rows * C * 4;  // ConvNeXt hidden
  {
    int64_t t = (int64_t)T;
    for (int i = 0; i < c.num_upsample; ++i) { t *= 2; max_elems = std::max<int64_t>(max_elems, (int64_t)B * t * C * 4); }
    int ch = c.decoder_dim;
    max_elems = std::max<int64_t>(max_elems, (int64_t)B * t * ch);
    for (int i = 0; i < c.num_rates; ++i) { t *= c.rates[i]; ch /= 2; max_elems = std::max<int64_t>(max_elems, (int64_t)B * t * ch); }
  }
  bf16* XN = (bf16*)h->wsget("dac.XN", (size_t)rows * C * 2, s);
  bf16* Q = (bf16*)h->wsget("dac.Q", (size_t)rows * C * 2, s);
  bf16* K = (bf16*)h->wsget("dac.K", (size_t)rows * C * 2, s);
  bf16* V = (bf16*)h->wsget("dac.V", (size_t)rows * C * 2, s);
  bf16* AO = (bf16*)h->wsget("dac.AO", (size_t)rows * C * 2, s);
  bf16* Hh = (bf16*)h->wsget("dac.Hh", (size_t)rows * I * 2, s);
  float* xa = (float*)h->wsget("dac.xa", (size_t)max_elems * 4, s);
  bf16* sa = (bf16*)h->wsget("dac.sa", (size_t)max_elems * 2, s);
  bf16* sb = (bf16*)h->wsget("dac.sb", (size_t)max_elems * 2, s);
  bf16* hb = (bf16*)h->wsget("dac.hb", (size_t)max_elems * 2, s);
  if (!XN || !Q || !K || !V || !AO || !Hh || !xa || !sa || !sb || !hb) { set_error("dac: workspace allocation failed"); return ECHO_ERR_CUDA; }
  const float eps = c.post_norm_eps;

  // ---- quantizer.post_module (autoencoder.py:786-802, 621-626)
  for (int i = 0; i < c.post_layers; ++i) {
    const DacPostLayerW& w = h->post[i];
    rmsnorm_affine(X, XN, w.attn_norm, nullptr, rows, C, 0, 0, eps, s);
    {
      GemmCall g = base_gemm(XN, C, w.wqkv, C, 1, rows, 3 * C, C, 1);
      g.p.epi = EPI_QKV;
      g.p.sec[0] = {Q, nullptr, 1 << 20, 0};
      g.p.sec[1] = {K, nullptr, 1 << 20, 0};
      g.p.sec[2] = {V, nullptr, 0, 0};
      g.p.sec_width = C; g.p.rope_cos = h->dac_rope_cos; g.p.rope_sin = h->dac_rope_sin; g.p.head_dim = 64;
      g.p.pos_period = T; g.p.eps = eps;
      DAC_GEMM(g);
    }
    {
      echo_attn_desc a;
      std::memset(&a, 0, sizeof(a));
      a.Q = Q; a.q_batch_stride = (int64_t)T * C; a.q_row_stride = C; a.out = AO;
      a.b = B; a.S = T; a.H = H; a.D = 64; a.scale = 0.125f; a.nseg = 1;
      a.seg[0].K = K; a.seg[0].V = V; a.seg[0].batch_stride = (int64_t)T * C; a.seg[0].row_stride = C;
      a.seg[0].len = T; a.seg[0].causal = 1; a.seg[0].window = c.post_window; a.seg[0].mask_stride = 1;
      cudaError_t er = attention_launch(a, s);
      if (er != cudaSuccess) { set_error("dac attention: %s", cudaGetErrorString(er)); return ECHO_ERR_CUDA; }
    }
    {
      GemmCall g = base_gemm(AO, C, w.wo, C, 1, rows, C, C, 1);
      g.p.gate = w.attn_gamma; g.p.resid = X; g.p.out_f32 = X; g.p.ld_f32 = C;
      g.split_k = 1;  // no atomic split-K here: < 0.1 ms to gain, and the decode stays bit-reproducible
      DAC_GEMM(g);
    }
    rmsnorm_affine(X, XN, w.ffn_norm, nullptr, rows, C, 0, 0, eps, s);
    {
      GemmCall g = base_gemm(XN, C, w.w13, C, 1, rows, 2 * I, C, 1);
      g.p.epi = EPI_SWIGLU; g.p.out_bf16 = Hh; g.p.ld_bf16 = I;
      DAC_GEMM(g);
    }
    {
      GemmCall g = base_gemm(Hh, I, w.w2, I, 1, rows, C, I, 1);
      g.p.gate = w.ffn_gamma; g.p.resid = X; g.p.out_f32 = X; g.p.ld_f32 = C;
      g.split_k = 1;
      DAC_GEMM(g);
    }
  }
  rmsnorm_affine(X, sa, h->post_final_norm, nullptr, rows, C, 0, 0, eps, s);  // sa = post_module output, bf16

  // ---- quantizer.upsample (autoencoder.py:427-435): [ConvTranspose k2 s2 ; ConvNeXt] per stage
  int Tc = T;
  bf16* cur = sa;
  bf16* nxt = sb;
  for (int i = 0; i < c.num_upsample; ++i) {
    const DacUpW& u = h->up[i];
    {
      GemmCall g = convt_gemm(u.convt, cur, B, Tc, C);
      g.p.out_f32 = xa; g.p.ld_f32 = 2 * C;
      DAC_GEMM(g);
    }
    Tc *= 2;
    const int r2 = B * Tc;
    launch_k(dwconv_ln_kernel, dim3(r2), dim3(256), 0, s, 1, xa, u.dw_w, u.dw_b, u.ln_w, u.ln_b, nxt, Tc, C);
    count_launch();
    {
      GemmCall g = base_gemm(nxt, C, u.w1, C, 1, r2, 4 * C, C, 1);
      g.p.bias = u.b1; g.p.out_bf16 = hb; g.p.ld_bf16 = 4 * C; g.p.act = ACT_GELU;
      DAC_GEMM(g);
    }
    {
      GemmCall g = base_gemm(hb, 4 * C, u.w2, 4 * C, 1, r2, C, 4 * C, 1);
      g.p.bias = u.b2; g.p.gate = u.gamma; g.p.resid = xa; g.p.out_f32 = xa; g.p.ld_f32 = C;
      g.p.out_bf16 = cur; g.p.ld_bf16 = C;  // bf16 copy feeds the next (transposed) conv
      DAC_GEMM(g);
    }
    // `cur` now holds this stage's output in bf16 (the convT that read it has completed in stream order)
  }

  // ---- decoder (autoencoder.py:984-998)
  {
    GemmCall g = conv_gemm(h->dec_conv0, cur, B, Tc, 1);
    g.p.out_bf16 = nxt; g.p.ld_bf16 = h->dec_conv0.n; g.p.act = ACT_SNAKE; g.p.alpha = h->stage[0].alpha_in; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
    DAC_GEMM(g);
  }
  std::swap(cur, nxt);  // cur = snake(conv0 out), channels decoder_dim
  for (int b = 0; b < c.num_rates; ++b) {
    const DacStageW& st = h->stage[b];
    {
      GemmCall g = convt_gemm(st.convt, cur, B, Tc, st.cout);
      g.p.out_f32 = xa; g.p.ld_f32 = st.stride * st.cout;
      g.p.out_bf16 = nxt; g.p.ld_bf16 = st.stride * st.cout; g.p.act = ACT_SNAKE; g.p.alpha = st.ru[0].alpha1; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
      DAC_GEMM(g);
    }
    Tc *= st.stride;
    std::swap(cur, nxt);  // cur = snake1(x) for residual unit 0
    static const int dil[3] = {1, 3, 9};
    for (int u = 0; u < 3; ++u) {
      const DacResUnitW& ru = st.ru[u];
      {
        GemmCall g = conv_gemm(ru.conv7, cur, B, Tc, dil[u]);
        g.p.out_bf16 = hb; g.p.ld_bf16 = st.cout; g.p.act = ACT_SNAKE; g.p.alpha = ru.alpha2; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
        DAC_GEMM(g);
      }
      {
        const float* next_alpha = (u < 2) ? st.ru[u + 1].alpha1
                                  : (b + 1 < c.num_rates ? h->stage[b + 1].alpha_in : h->final_alpha);
        GemmCall g = conv_gemm(ru.conv1, hb, B, Tc, 1);
        g.p.resid = xa; g.p.out_f32 = xa; g.p.ld_f32 = st.cout;
        g.p.out_bf16 = cur; g.p.ld_bf16 = st.cout; g.p.act = ACT_SNAKE; g.p.alpha = next_alpha; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
        DAC_GEMM(g);
      }
    }
  }
  const int Cl = h->stage.back().cout;
  dim3 grid((Tc + 255) / 256, B);
  launch_k(final_conv_tanh_kernel, dim3(grid), dim3(256), 7 * Cl * sizeof(float), s, 1, cur, h->final_w, h->final_b, audio, Tc, Cl);
  count_launch();
  ECHO_CUDA(cudaGetLastError());
  return ECHO_OK;
}

int dac_check(echo_handle* h, const char* who) {
  if (!h) { set_error("%s: null handle", who); return ECHO_ERR_ARG; }
  if (!h->dac_ready) { set_error("%s: DAC weights not finalized (echo_dac_finalize)", who); return ECHO_ERR_STATE; }
  cudaSetDevice(h->device);
  return ECHO_OK;
}

}  // namespace

extern "C" int echo_dac_decode(echo_handle* h, const float* z, const float* pca_components, const float* pca_mean,
                               float latent_scale, int B, int T, float* audio, void* stream) {
  ECHO_TRY(dac_check(h, "echo_dac_decode"));
  if (!z || !pca_components || !pca_mean || !audio || B <= 0 || T <= 0) { set_error("echo_dac_decode: bad argument"); return ECHO_ERR_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = h->dcfg.latent_dim, Kp = h->dcfg.pca_dim;
  float* X = (float*)h->wsget("dac.X", (size_t)B * T * C * 4, s);
  if (!X) { set_error("dac: out of memory"); return ECHO_ERR_CUDA; }
  launch_k(pca_unproject_kernel, dim3(B * T), dim3(256), Kp * sizeof(float), s, 1, z, pca_components, pca_mean, latent_scale, X, Kp, C);
  count_launch();
  return dac_run(h, X, B, T, audio, s);
}

extern "C" int echo_dac_decode_zq(echo_handle* h, const float* zq, int B, int T, float* audio, void* stream) {
  ECHO_TRY(dac_check(h, "echo_dac_decode_zq"));
  if (!zq || !audio || B <= 0 || T <= 0) { set_error("echo_dac_decode_zq: bad argument"); return ECHO_ERR_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = h->dcfg.latent_dim;
  float* X = (float*)h->wsget("dac.X", (size_t)B * T * C * 4, s);
  if (!X) { set_error("dac: out of memory"); return ECHO_ERR_CUDA; }
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_k(transpose_ct_kernel, dim3(grid), dim3(block), 0, s, 1, zq, X, C, T);
  count_launch();
  return dac_run(h, X, B, T, audio, s);
}
