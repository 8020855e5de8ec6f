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
#include <cstdlib>
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
                                                        int C, int halo) {
  // halo > 0 (streaming decode, one batch item): `halo` rows of the previous block sit in front of z's row 0
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
      if (tt >= -halo) acc = fmaf(w[c * 7 + j], z[((int64_t)row - 6 + j) * C + c], acc);
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
                                                              float bias, float* __restrict__ audio, int T, int C, int halo) {
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
    if (tt < -halo) continue;  // halo rows of the previous block (streaming) sit in front of row 0
    const uint4* rp = reinterpret_cast<const uint4*>(base + (int64_t)tt * C);
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

// Same op, restructured for C = 8 * CS channels (96 in the real model). The kernel above reads each weight from shared
// memory once per FMA (672 broadcast LDS per output sample) and took 362 us for 1.3 M samples, 9x its HBM floor.
// Here 8 threads share a row: each keeps its 7 x CS weights in registers, loads its CS channels of the row once
// (the 8 threads cover the 192-byte row with three 64-bit loads each), forms the 7 per-tap partial dot products and
// the octet reduces them by shuffles into d[tap][row] in shared memory; the output sample is then the sum of 7
// diagonal entries. Every activation byte is read once: the kernel is bound by the 252 MB it has to read.
template <int CS>
__global__ void __launch_bounds__(256) final_conv_tanh_rows_kernel(const bf16* __restrict__ sx, const float* __restrict__ w,
                                                                   float bias, float* __restrict__ audio, int T, int halo) {
  pdl_wait();
  pdl_trigger();
  constexpr int C = 8 * CS, TB = 256, ROWS = TB + 6;
  static_assert(CS % 4 == 0, "64-bit loads");
  __shared__ float d[7][ROWS + 2];
  const int b = blockIdx.y, t0 = blockIdx.x * TB;
  const int slice = threadIdx.x & 7, grp = threadIdx.x >> 3;  // 32 rows in flight per pass
  float wr[7][CS];
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int c = 0; c < CS; ++c) wr[j][c] = __ldg(w + j * C + slice * CS + c);
  const bf16* base = sx + (size_t)b * T * C;
  for (int r0 = 0; r0 < ROWS; r0 += 32) {
    const int r = r0 + grp;          // row of this block's window
    const int tt = t0 - 6 + r;       // time index (negative: causal zero padding)
    float x[CS];
    const bool live = r < ROWS && tt >= -halo && tt < T;
    if (live) {
      const uint2* rp = reinterpret_cast<const uint2*>(base + (int64_t)tt * C + slice * CS);
#pragma unroll
      for (int q = 0; q < CS / 4; ++q) {
        const uint2 u = __ldg(rp + q);
        const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        x[4 * q] = f0.x; x[4 * q + 1] = f0.y; x[4 * q + 2] = f1.x; x[4 * q + 3] = f1.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < CS; ++c) x[c] = 0.f;
    }
    float p[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) a = fmaf(x[c], wr[j][c], a);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      p[j] = a;
    }
    if (slice == 0 && r < ROWS) {
#pragma unroll
      for (int j = 0; j < 7; ++j) d[j][r] = p[j];
    }
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t < T) {
    float acc = bias;
#pragma unroll
    for (int j = 0; j < 7; ++j) acc += d[j][threadIdx.x + j];  // row (t - 6 + j) sits at window index threadIdx.x + j
    audio[(size_t)b * T + t] = tanhf(acc);
  }
}

// ---- encode path kernels -------------------------------------------------------------------------------------
// First encoder conv (Cin = 1, k = 7, causal; autoencoder.py:915): x[b, t, c] = b[c] + sum_j w[j][c] a[b, t-6+j];
// writes the fp32 residual stream and snake(x, alpha) in bf16 (the A operand of the first ResidualUnit's conv7).
__global__ void __launch_bounds__(256) enc_conv0_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ alpha,
                                                        float* __restrict__ xa, bf16* __restrict__ sx, int T, int C) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sw[];  // [7][C] weights, [C] bias, [C] alpha
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sw[7 * C + i] = bias[i]; sw[8 * C + i] = alpha[i]; }
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float a[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) a[j] = (t - 6 + j >= 0) ? audio[(size_t)b * T + t - 6 + j] : 0.f;
  float* xo = xa + ((size_t)b * T + t) * C;
  bf16* so = sx + ((size_t)b * T + t) * C;
  for (int c = 0; c < C; c += 2) {
    float v0 = sw[7 * C + c], v1 = sw[7 * C + c + 1];
#pragma unroll
    for (int j = 0; j < 7; ++j) { v0 = fmaf(sw[j * C + c], a[j], v0); v1 = fmaf(sw[j * C + c + 1], a[j], v1); }
    *reinterpret_cast<float2*>(xo + c) = make_float2(v0, v1);
    const float al0 = sw[8 * C + c], al1 = sw[8 * C + c + 1];
    const float s0 = sinf(al0 * v0), s1 = sinf(al1 * v1);
    *reinterpret_cast<__nv_bfloat162*>(so + c) = __floats2bfloat162_rn(v0 + s0 * s0 / (al0 + 1e-9f), v1 + s1 * s1 / (al1 + 1e-9f));
  }
}

// first conv weight (C, 1, 7) weight-normed -> [7][C] fp32
__global__ void pack_enc_conv0_kernel(const float* __restrict__ v, const float* __restrict__ scale, float* __restrict__ out,
                                      int C) {
  for (int i = threadIdx.x; i < C * 7; i += blockDim.x) {
    const int c = i % C, j = i / C;
    out[i] = v[(size_t)c * 7 + j] * scale[c];
  }
}

// RMSNorm with weight over fp32 rows (Transformer.norm, autoencoder.py:607), one warp per row, C = 128 * NV:
//   out_f32 = x_hat * w (optional);  out_bf16 = bf16(alpha ? snake(x_hat * w, alpha) : x_hat * w) (optional)
template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_out_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                                          const float* __restrict__ alpha, float* __restrict__ out_f32,
                                                          bf16* __restrict__ out_bf16, int rows, float eps) {
  pdl_wait();
  pdl_trigger();
  constexpr int W = 128 * NV;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(X + (size_t)r * W);
  float4 v[NV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)W + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + c);
    float4 o = make_float4(v[i].x * rstd * wv.x, v[i].y * rstd * wv.y, v[i].z * rstd * wv.z, v[i].w * rstd * wv.w);
    if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)r * W)[c] = o;
    if (out_bf16) {
      if (alpha) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(alpha) + c);
        const float s0 = sinf(a.x * o.x), s1 = sinf(a.y * o.y), s2 = sinf(a.z * o.z), s3 = sinf(a.w * o.w);
        o.x += s0 * s0 / (a.x + 1e-9f); o.y += s1 * s1 / (a.y + 1e-9f);
        o.z += s2 * s2 / (a.z + 1e-9f); o.w += s3 * s3 / (a.w + 1e-9f);
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
      reinterpret_cast<uint2*>(out_bf16 + (size_t)r * W)[c] =
          make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
    }
  }
}

// in_proj (C -> D, weight-normed 1x1 conv) as fp32 [D][C]; D <= 16
__global__ void pack_vq_in_kernel(const float* __restrict__ v, const float* __restrict__ scale, float* __restrict__ out,
                                  int D, int C) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D * C; i += gridDim.x * blockDim.x) out[i] = v[i] * scale[i / C];
}
// normalised codebook + squared norms (F.normalize: x / max(||x||, 1e-12); autoencoder.py:149)
__global__ void pack_vq_codebook_kernel(const float* __restrict__ cb, float* __restrict__ cbn, float* __restrict__ sq,
                                        int size, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= size) return;
  float ss = 0.f;
  for (int d = 0; d < D; ++d) ss += cb[i * D + d] * cb[i * D + d];
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  float s2 = 0.f;
  for (int d = 0; d < D; ++d) { const float t = cb[i * D + d] * inv; cbn[i * D + d] = t; s2 += t * t; }
  sq[i] = s2;
}
// out_table[i][c] = bias[c] + sum_d (v[c][d] * scale[c]) * codebook[i][d]     (out_proj of every code, fp32)
__global__ void pack_vq_out_kernel(const float* __restrict__ v, const float* __restrict__ scale, const float* __restrict__ bias,
                                   const float* __restrict__ cb, float* __restrict__ table, int size, int D, int C) {
  const int64_t total = (int64_t)size * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int code = (int)(i / C);
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(v[(size_t)c * D + d] * scale[c], cb[code * D + d], acc);
    table[i] = acc + bias[c];
  }
}

struct VqDev {  // device-side view of one DacVqW
  const float *in_w, *in_b, *cb_norm, *cb_sq, *out_table;
  int size;
};
struct VqAll { VqDev q[16]; int n; };

// Semantic + residual vector quantisation of one row (autoencoder.py:132-157, 455-463, 1116-1126), all fp32:
// per codebook z_e = in_proj(residual); nearest code between the L2-normalised z_e and codebook (first index on
// ties, like torch.max); residual -= out_proj(code). One block per row keeps the row in registers across all
// codebooks. z_q = semantic + (sum of the residual codebooks), summed in the reference's order.
template <int NPT>  // C = 256 * NPT
__global__ void __launch_bounds__(256) vq_encode_kernel(const float* __restrict__ Z, const __grid_constant__ VqAll vq,
                                                        float* __restrict__ zq, int32_t* __restrict__ codes, int T,
                                                        int D) {
  pdl_wait();
  pdl_trigger();
  constexpr int C = 256 * NPT;
  __shared__ float s_red[8][16];
  __shared__ float s_e[16];
  __shared__ float s_best[8];
  __shared__ int s_idx[8];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float res[NPT], acc_sem[NPT], acc_res[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) { res[i] = Z[(size_t)r * C + tid + 256 * i]; acc_sem[i] = 0.f; acc_res[i] = 0.f; }
  for (int qi = 0; qi < vq.n; ++qi) {
    const VqDev& q = vq.q[qi];
    // ---- z_e = in_proj(residual): D dot products over C
    for (int d = 0; d < D; ++d) {
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < NPT; ++i) p = fmaf(res[i], q.in_w[(size_t)d * C + tid + 256 * i], p);
      for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (lane == 0) s_red[wid][d] = p;
    }
    __syncthreads();
    if (tid < D) {
      float t = q.in_b[tid];
      for (int w = 0; w < 8; ++w) t += s_red[w][tid];
      s_e[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {  // F.normalize(z_e)
      float ss = 0.f;
      for (int d = 0; d < D; ++d) ss += s_e[d] * s_e[d];
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      float s2 = 0.f;
      for (int d = 0; d < D; ++d) { s_e[d] *= inv; s2 += s_e[d] * s_e[d]; }
      s_e[15] = s2;  // |e_n|^2 (D <= 15)
    }
    __syncthreads();
    // ---- nearest code: maximise -(|e|^2 - 2 e.c + |c|^2)
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int code = tid; code < q.size; code += 256) {
      float dot = 0.f;
      for (int d = 0; d < D; ++d) dot = fmaf(s_e[d], q.cb_norm[(size_t)code * D + d], dot);
      const float score = -(s_e[15] - 2.f * dot + q.cb_sq[code]);
      if (score > best) { best = score; bidx = code; }  // ascending codes per thread: first maximum is kept
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { s_best[wid] = best; s_idx[wid] = bidx; }
    __syncthreads();
    if (tid == 0) {
      float bb = s_best[0];
      int bi = s_idx[0];
      for (int w = 1; w < 8; ++w)
        if (s_best[w] > bb || (s_best[w] == bb && s_idx[w] < bi)) { bb = s_best[w]; bi = s_idx[w]; }
      s_idx[0] = bi;
      if (codes) codes[((size_t)(r / T) * vq.n + qi) * T + (r % T)] = bi;
    }
    __syncthreads();
    const int idx = s_idx[0];
    const float* row = q.out_table + (size_t)idx * C;
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const float t = row[tid + 256 * i];
      res[i] -= t;
      if (qi == 0) acc_sem[i] = t;
      else acc_res[i] += t;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NPT; ++i) zq[(size_t)r * C + tid + 256 * i] = acc_sem[i] + acc_res[i];
}

// latent[r, k] = scale * sum_c (zq[r, c] - mean[c]) * comps[k, c]        (inference.py:222-223), fp32; one block per row
__global__ void __launch_bounds__(256) pca_project_kernel(const float* __restrict__ zq, const float* __restrict__ comps,
                                                          const float* __restrict__ mean, float scale,
                                                          float* __restrict__ out, int C, int K) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sz[];
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sz[c] = zq[(size_t)r * C + c] - mean[c];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = wid; k < K; k += 8) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(sz[c], comps[(size_t)k * C + c], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(size_t)r * K + k] = acc * scale;
  }
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}


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
  if (c->num_enc_rates > 0) {  // encode path described
    int d = c->enc_dim;
    if (d % 64 || c->num_enc_rates > 8 || c->codebook_dim < 1 || c->codebook_dim > 15 || c->n_codebooks < 0 || c->n_codebooks > 15) {
      set_error("echo_dac_configure: unsupported encoder dims (enc_dim %% 64, codebook_dim <= 15, n_codebooks <= 15)");
      return ECHO_ERR_ARG;
    }
    for (int i = 0; i < c->num_enc_rates; ++i) d *= 2;
    if (d != C) { set_error("echo_dac_configure: enc_dim * 2^num_enc_rates must equal latent_dim"); return ECHO_ERR_ARG; }
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
  // ---- encode path (optional: present iff the encoder weights were loaded)
  h->dac_enc_ready = false;
  const bool has_enc = h->dac_raw.count("encoder.block.0.conv.bias") > 0;
  if (has_enc && pk.rc == ECHO_OK) {
    const int D0 = c.enc_dim, nb = c.num_enc_rates, cd = c.codebook_dim;
    {  // first conv, Cin = 1
      const RawTensor* v = pk.get("encoder.block.0.conv.parametrizations.weight.original1");
      const RawTensor* g = pk.get("encoder.block.0.conv.parametrizations.weight.original0");
      if (v && g && v->numel == (int64_t)D0 * 7) {
        float* scale = (float*)h->dalloc((size_t)D0 * 4);
        slice_norm_kernel<<<D0, 256, 0, s>>>(v->p, scale, g->p, 7);
        h->enc_conv0_w = (float*)h->dalloc((size_t)7 * D0 * 4);
        pack_enc_conv0_kernel<<<1, 256, 0, s>>>(v->p, scale, h->enc_conv0_w, D0);
      } else if (pk.rc == ECHO_OK) { set_error("dac.encoder.block.0: bad first conv"); pk.rc = ECHO_ERR_ARG; }
      h->enc_conv0_b = pk.vecf("encoder.block.0.conv.bias", D0);
    }
    h->enc_blk.resize(nb);
    int d = D0;
    for (int b = 0; b < nb; ++b) {
      const std::string p = "encoder.block." + std::to_string(b + 1) + ".block";
      DacEncBlockW& eb = h->enc_blk[b];
      eb.stride = c.enc_rates[b]; eb.cin = d; eb.cout = 2 * d;
      for (int u = 0; u < 3; ++u) {
        const std::string q = p + "." + std::to_string(u) + ".block";
        eb.ru[u].alpha1 = pk.alpha(q + ".0.alpha", d);
        eb.ru[u].conv7 = pk.conv(q + ".1", d, d, 7, true);
        eb.ru[u].alpha2 = pk.alpha(q + ".2.alpha", d);
        eb.ru[u].conv1 = pk.conv(q + ".3", d, d, 1, true);
      }
      eb.alpha_out = pk.alpha(p + ".3.alpha", d);
      eb.down = pk.conv(p + ".4", 2 * d, d, 2 * eb.stride, true);  // [Cout][k][Cin] == 2 taps of (stride * Cin)
      d *= 2;
      if (b == nb - 1 && c.enc_t_layers > 0) {
        const int It = 3 * d;
        h->enc_tf.resize(c.enc_t_layers);
        for (int i = 0; i < c.enc_t_layers; ++i) {
          const std::string q = p + ".5.layers." + std::to_string(i);
          DacPostLayerW& l = h->enc_tf[i];
          l.wqkv = pk.matb(q + ".attention.wqkv.weight", 3 * d, d);
          l.wo = pk.matb(q + ".attention.wo.weight", d, d);
          l.w13 = (bf16*)h->dalloc((size_t)2 * It * d * 2);
          pk.matb(q + ".feed_forward.w1.weight", It, d, l.w13, 128, 256, 0);
          pk.matb(q + ".feed_forward.w3.weight", It, d, l.w13, 128, 256, 128);
          l.w2 = pk.matb(q + ".feed_forward.w2.weight", d, It);
          l.attn_norm = pk.vecf(q + ".attention_norm.weight", d);
          l.ffn_norm = pk.vecf(q + ".ffn_norm.weight", d);
          l.attn_gamma = pk.vecf(q + ".attention_layer_scale.gamma", d);
          l.ffn_gamma = pk.vecf(q + ".ffn_layer_scale.gamma", d);
        }
        h->enc_tf_norm = pk.vecf(p + ".5.norm.weight", d);
      }
    }
    if (d != C && pk.rc == ECHO_OK) { set_error("encoder width %d != latent_dim %d", d, C); pk.rc = ECHO_ERR_ARG; }
    h->enc_alpha_out = pk.alpha("encoder.block." + std::to_string(nb + 1) + ".alpha", C);
    h->enc_conv_out = pk.conv("encoder.block." + std::to_string(nb + 2), C, C, 3, true);
    h->down.resize(c.num_upsample);
    for (int i = 0; i < c.num_upsample; ++i) {
      const std::string p = "quantizer.downsample." + std::to_string(i);
      DacUpW& u = h->down[i];
      u.convt = pk.conv(p + ".0", C, C, 2, false);  // k = s = 2: one tap over the (T / 2, 2 C) view
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
    h->pre.resize(c.post_layers);
    for (int i = 0; i < c.post_layers; ++i) {
      const std::string p = "quantizer.pre_module.layers." + std::to_string(i);
      DacPostLayerW& l = h->pre[i];
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
    h->pre_final_norm = pk.vecf("quantizer.pre_module.norm.weight", C);
    h->vq.resize(1 + c.n_codebooks);
    for (int qi = 0; qi <= c.n_codebooks; ++qi) {
      const std::string p = qi == 0 ? std::string("quantizer.semantic_quantizer.quantizers.0")
                                    : "quantizer.quantizer.quantizers." + std::to_string(qi - 1);
      DacVqW& q = h->vq[qi];
      q.size = qi == 0 ? c.semantic_codebook_size : c.codebook_size;
      const RawTensor* iv = pk.get(p + ".in_proj.parametrizations.weight.original1");
      const RawTensor* ig = pk.get(p + ".in_proj.parametrizations.weight.original0");
      const RawTensor* ov = pk.get(p + ".out_proj.parametrizations.weight.original1");
      const RawTensor* og = pk.get(p + ".out_proj.parametrizations.weight.original0");
      const RawTensor* ob = pk.get(p + ".out_proj.bias");
      const RawTensor* cb = pk.get(p + ".codebook.weight");
      q.in_b = pk.vecf(p + ".in_proj.bias", cd);
      if (!iv || !ig || !ov || !og || !ob || !cb) break;
      if (iv->numel != (int64_t)cd * C || ov->numel != (int64_t)C * cd || cb->numel != (int64_t)q.size * cd) {
        if (pk.rc == ECHO_OK) { set_error("dac.%s: bad quantizer tensor sizes", p.c_str()); pk.rc = ECHO_ERR_ARG; }
        break;
      }
      float* si = (float*)h->dalloc((size_t)cd * 4);
      slice_norm_kernel<<<cd, 256, 0, s>>>(iv->p, si, ig->p, C);
      q.in_w = (float*)h->dalloc((size_t)cd * C * 4);
      pack_vq_in_kernel<<<grid_for((int64_t)cd * C), 256, 0, s>>>(iv->p, si, q.in_w, cd, C);
      float* so = (float*)h->dalloc((size_t)C * 4);
      slice_norm_kernel<<<C, 256, 0, s>>>(ov->p, so, og->p, cd);
      q.cb_norm = (float*)h->dalloc((size_t)q.size * cd * 4);
      q.cb_sq = (float*)h->dalloc((size_t)q.size * 4);
      pack_vq_codebook_kernel<<<(q.size + 255) / 256, 256, 0, s>>>(cb->p, q.cb_norm, q.cb_sq, q.size, cd);
      q.out_table = (float*)h->dalloc((size_t)q.size * C * 4);
      pack_vq_out_kernel<<<grid_for((int64_t)q.size * C), 256, 0, s>>>(ov->p, so, ob->p, cb->p, q.out_table, q.size, cd, C);
    }
    h->dac_enc_ready = (pk.rc == ECHO_OK);
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

struct TfBuffers { bf16 *XN, *Q, *K, *V, *AO, *Hh; };

// The layers of a WindowLimitedTransformer (autoencoder.py:786-802, 621-626) over the fp32 stream X (B*T rows, in
// place): pre-RMSNorm, fused QKV with RoPE on all heads, causal window attention, LayerScale residuals, SwiGLU.
// Shared by quantizer.post_module / pre_module (window 128) and the last EncoderBlock (window 512). The final
// RMSNorm (:607) is left to the caller (its output format differs per use).
// Streaming (kcache != nullptr, B == 1): the T rows are positions pos .. pos + T - 1 of a longer sequence; layer i's
// keys / values are appended to kcache[i] / vcache[i] ((max_T, C) bf16, rows < pos written by earlier blocks) and the
// attention runs the T newest rows against the whole cache (causal + window with q_offset = pos).
int run_window_transformer(echo_handle* h, const std::vector<DacPostLayerW>& layers, float* X, int B, int T, int C, int I,
                           int H, int window, float eps, const TfBuffers& tb, cudaStream_t s,
                           bf16* const* kcache = nullptr, bf16* const* vcache = nullptr, int pos = 0) {
  const int rows = B * T;
  bf16 *XN = tb.XN, *Q = tb.Q, *AO = tb.AO, *Hh = tb.Hh;
  for (size_t i = 0; i < layers.size(); ++i) {
    const DacPostLayerW& w = layers[i];
    bf16* K = kcache ? kcache[i] + (size_t)pos * C : tb.K;
    bf16* V = vcache ? vcache[i] + (size_t)pos * C : tb.V;
    rmsnorm_affine(X, XN, w.attn_norm, nullptr, rows, C, 0, 0, eps, s);
    {
      GemmCall g = base_gemm(XN, C, w.wqkv, C, 1, rows, 3 * C, C, 1);
      g.p.epi = EPI_QKV;
      g.p.sec[0] = {Q, nullptr, 1 << 20, 0};
      g.p.sec[1] = {K, nullptr, 1 << 20, 0};
      g.p.sec[2] = {V, nullptr, 0, 0};
      g.p.sec_width = C; g.p.rope_cos = h->dac_rope_cos; g.p.rope_sin = h->dac_rope_sin; g.p.head_dim = 64;
      g.p.pos_period = T; g.p.pos_offset = pos; g.p.eps = eps;
      DAC_GEMM(g);
    }
    {
      echo_attn_desc a;
      std::memset(&a, 0, sizeof(a));
      a.Q = Q; a.q_batch_stride = (int64_t)T * C; a.q_row_stride = C; a.out = AO;
      a.b = B; a.S = T; a.H = H; a.D = 64; a.scale = 0.125f; a.nseg = 1;
      a.seg[0].K = kcache ? kcache[i] : K; a.seg[0].V = vcache ? vcache[i] : V;
      a.seg[0].batch_stride = (int64_t)(pos + T) * C; a.seg[0].row_stride = C;
      a.seg[0].len = pos + T; a.seg[0].q_offset = pos;
      a.seg[0].causal = 1; a.seg[0].window = window; a.seg[0].mask_stride = 1;
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
  return ECHO_OK;
}

}  // namespace

// State of one streaming decode: what the causal decoder has to remember between blocks of latents.
//   * quantizer.post_module: the keys / values of every layer for all latents so far (window-128 causal attention,
//     autoencoder.py:762-773) -- (max_T, C) bf16 per layer, the attention reads the last 127 rows + the new block;
//   * every causal conv with a receptive field into the past (autoencoder.py:285-289, left pad (k-1)*d): its last
//     (k-1)*d input rows ("halo"): the ConvNeXt depthwise convs (6 rows fp32), decoder conv0 (6), each stage's transposed
//     conv (1), the dilated conv7 of every ResidualUnit (6, 18, 54) and the final conv7 (6).
// All halos start as zeros == the causal zero padding at the start of the sequence.
struct echo_dac_stream {
  int max_T = 0, pos = 0;
  std::vector<bf16*> K, V;            // [post_layers]
  std::vector<float*> dw_halo;        // [num_upsample]  6 x C fp32
  bf16* conv0_halo = nullptr;         // 6 x C
  std::vector<bf16*> convt_halo;      // [num_rates]     1 x cin
  std::vector<bf16*> ru_halo;         // [num_rates * 3] 6 * dil x cout
  bf16* final_halo = nullptr;         // 6 x C_last
  std::vector<std::pair<void*, size_t>> bufs;  // every allocation (for reset / destroy)
};

namespace {

constexpr int64_t DAC_HALO_ELEMS = 64 * 1536;  // headroom in front of every bf16 activation buffer: >= 54 rows x 768 ch

// One decode. Offline (st == nullptr): B x T latents -> B x hop*T samples. Streaming (st != nullptr, B == 1): the next
// T latents of the stream -> the next hop*T samples, bit-identical to the same samples of an offline decode of the
// whole sequence: every kernel computes an output row from the same inputs in the same order, the rows a conv needs
// from before the block come from the halos instead of the same buffer.
int dac_run(echo_handle* h, float* X /* (B*T, C) fp32 latent, time-major */, int B, int T, float* audio, cudaStream_t s,
            echo_dac_stream* st = nullptr) {
  const echo_dac_config& c = h->dcfg;
  const int C = c.latent_dim, I = c.post_intermediate, H = c.post_heads, rows = B * T;
  if (T > 4096) { set_error("dac: T=%d exceeds the post_module block size 4096", T); return ECHO_ERR_ARG; }
  if (st && B != 1) { set_error("dac stream: one batch item per stream"); return ECHO_ERR_ARG; }
  if (st && st->pos + T > st->max_T) { set_error("dac stream: %d + %d latents exceed the stream's capacity %d", st->pos, T, st->max_T); return ECHO_ERR_ARG; }
  // sizes of the largest activations
  int64_t max_elems = (int64_t)rows * C * 4;  // ConvNeXt hidden
  {
    int64_t t = (int64_t)T;
    for (int i = 0; i < c.num_upsample; ++i) { t *= 2; max_elems = std::max<int64_t>(max_elems, (int64_t)B * t * C * 4); }
    int ch = c.decoder_dim;
    max_elems = std::max<int64_t>(max_elems, (int64_t)B * t * ch);
    for (int i = 0; i < c.num_rates; ++i) { t *= c.rates[i]; ch /= 2; max_elems = std::max<int64_t>(max_elems, (int64_t)B * t * ch); }
  }
  const int64_t HR = DAC_HALO_ELEMS;  // halo headroom (elements) in front of row 0 of xa / sa / sb
  bf16* XN = (bf16*)h->wsget("dac.XN", (size_t)rows * C * 2, s);
  bf16* Q = (bf16*)h->wsget("dac.Q", (size_t)rows * C * 2, s);
  bf16* K = (bf16*)h->wsget("dac.K", (size_t)rows * C * 2, s);
  bf16* V = (bf16*)h->wsget("dac.V", (size_t)rows * C * 2, s);
  bf16* AO = (bf16*)h->wsget("dac.AO", (size_t)rows * C * 2, s);
  bf16* Hh = (bf16*)h->wsget("dac.Hh", (size_t)rows * I * 2, s);
  float* xa = (float*)h->wsget("dac.xa", (size_t)(max_elems + HR) * 4, s);
  bf16* sa = (bf16*)h->wsget("dac.sa", (size_t)(max_elems + HR) * 2, s);
  bf16* sb = (bf16*)h->wsget("dac.sb", (size_t)(max_elems + HR) * 2, s);
  bf16* hb = (bf16*)h->wsget("dac.hb", (size_t)max_elems * 2, s);
  if (!XN || !Q || !K || !V || !AO || !Hh || !xa || !sa || !sb || !hb) { set_error("dac: workspace allocation failed"); return ECHO_ERR_CUDA; }
  xa += HR; sa += HR; sb += HR;
  const float eps = c.post_norm_eps;

  // halo plumbing (streaming only): `rows0` is row 0 of the block in a buffer of row width W; the h rows in front of it
  // are loaded from the stream state before the consumer runs and the last h rows of [halo | block] saved after it
#define DAC_HALO_IN(rows0, store, hrows, W, esz)                                                                         \
  do {                                                                                                                   \
    if (st) ECHO_CUDA(cudaMemcpyAsync((char*)(rows0) - (size_t)(hrows) * (W) * (esz), (store), (size_t)(hrows) * (W) * (esz),   \
                                      cudaMemcpyDeviceToDevice, s));                                                     \
  } while (0)
#define DAC_HALO_OUT(rows0, store, hrows, W, esz, nrows)                                                                 \
  do {                                                                                                                   \
    if (st) ECHO_CUDA(cudaMemcpyAsync((store), (char*)(rows0) + ((int64_t)(nrows) - (hrows)) * (W) * (esz),              \
                                      (size_t)(hrows) * (W) * (esz), cudaMemcpyDeviceToDevice, s));                      \
  } while (0)
  // a conv / transposed-conv GEMM whose A rows start `hrows` rows before the block (tap shifts become >= 0)
  auto with_halo = [&](GemmCall g, const bf16* rows0, int hrows, int n) {
    if (st && hrows > 0) {
      g.A = rows0 - (size_t)hrows * g.lda;
      for (int j = 0; j < g.p.taps; ++j) g.p.tap_shift[j] += hrows;
      g.a_rows = hrows + n;
      g.a_batch_stride = (int64_t)(hrows + n) * g.lda;
    }
    return g;
  };

  // ---- quantizer.post_module (autoencoder.py:786-802, 621-626)
  TfBuffers tb{XN, Q, K, V, AO, Hh};
  ECHO_TRY(run_window_transformer(h, h->post, X, B, T, C, I, H, c.post_window, eps, tb, s, st ? st->K.data() : nullptr,
                                  st ? st->V.data() : nullptr, st ? st->pos : 0));
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
    if (st) DAC_HALO_IN(xa, st->dw_halo[i], 6, C, 4);
    launch_k(dwconv_ln_kernel, dim3(r2), dim3(256), 0, s, 1, xa, u.dw_w, u.dw_b, u.ln_w, u.ln_b, nxt, Tc, C, st ? 6 : 0);
    count_launch();
    if (st) DAC_HALO_OUT(xa, st->dw_halo[i], 6, C, 4, Tc);  // before the ConvNeXt output overwrites xa below
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
    if (st) DAC_HALO_IN(cur, st->conv0_halo, 6, C, 2);
    GemmCall g = with_halo(conv_gemm(h->dec_conv0, cur, B, Tc, 1), cur, 6, Tc);
    g.p.out_bf16 = nxt; g.p.ld_bf16 = h->dec_conv0.n; g.p.act = ACT_SNAKE; g.p.alpha = h->stage[0].alpha_in; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
    DAC_GEMM(g);
    if (st) DAC_HALO_OUT(cur, st->conv0_halo, 6, C, 2, Tc);
  }
  std::swap(cur, nxt);  // cur = snake(conv0 out), channels decoder_dim
  for (int b = 0; b < c.num_rates; ++b) {
    const DacStageW& stg = h->stage[b];
    {
      if (st) DAC_HALO_IN(cur, st->convt_halo[b], 1, stg.cin, 2);
      GemmCall g = with_halo(convt_gemm(stg.convt, cur, B, Tc, stg.cout), cur, 1, Tc);
      g.p.out_f32 = xa; g.p.ld_f32 = stg.stride * stg.cout;
      g.p.out_bf16 = nxt; g.p.ld_bf16 = stg.stride * stg.cout; g.p.act = ACT_SNAKE; g.p.alpha = stg.ru[0].alpha1; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
      DAC_GEMM(g);
      if (st) DAC_HALO_OUT(cur, st->convt_halo[b], 1, stg.cin, 2, Tc);
    }
    Tc *= stg.stride;
    std::swap(cur, nxt);  // cur = snake1(x) for residual unit 0
    static const int dil[3] = {1, 3, 9};
    // ResidualUnit = Snake -> conv7 (dilated) -> Snake -> conv1 -> + x (autoencoder.py:884-900). For the two widest-in-time
    // stages (C = 192 and 96 channels, where the activations -- not the weights -- are the traffic) it is ONE kernel:
    // the conv7 accumulator goes Snake -> bf16 -> shared memory -> second MMA (EPI_RU, gemm_tc.cuh); the bf16
    // intermediate never reaches HBM and the stream is read and written once per unit: 831 -> 655 us (C = 96) and
    // 665 -> 528 us (C = 192) per unit, bit-identical to the two launches (tests/test_ops_gpu.py::test_fused_residual_unit).
    static const int env_fuse = [] { const char* e = std::getenv("ECHO_DAC_FUSED_RU"); return e ? atoi(e) : 1; }();
    const bool fuse = env_fuse != 0 && (stg.cout == 96 || stg.cout == 192);
    for (int u = 0; u < 3; ++u) {
      const DacResUnitW& ru = stg.ru[u];
      const int hr = 6 * dil[u];
      const float* next_alpha = (u < 2) ? stg.ru[u + 1].alpha1
                                : (b + 1 < c.num_rates ? h->stage[b + 1].alpha_in : h->final_alpha);
      if (fuse) {
        if (st) DAC_HALO_IN(cur, st->ru_halo[3 * b + u], hr, stg.cout, 2);
        GemmCall g = with_halo(conv_gemm(ru.conv7, cur, B, Tc, dil[u]), cur, hr, Tc);
        g.p.epi = EPI_RU;
        g.p.alpha = ru.alpha2; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
        g.B1 = ru.conv1.w; g.ldb1 = ru.conv1.cin; g.p.ru_bias1 = ru.conv1.bias;
        g.p.ru_alpha_out = next_alpha; g.p.ru_alpha_out_inv = h->dac_alpha_inv[next_alpha];
        g.p.resid = xa; g.p.out_f32 = xa; g.p.ld_f32 = stg.cout;
        g.p.out_bf16 = nxt; g.p.ld_bf16 = stg.cout;  // not `cur`: later tiles still read its rows (the conv's past)
        DAC_GEMM(g);
        if (st) DAC_HALO_OUT(cur, st->ru_halo[3 * b + u], hr, stg.cout, 2, Tc);
        std::swap(cur, nxt);
        continue;
      }
      {
        if (st) DAC_HALO_IN(cur, st->ru_halo[3 * b + u], hr, stg.cout, 2);
        GemmCall g = with_halo(conv_gemm(ru.conv7, cur, B, Tc, dil[u]), cur, hr, Tc);
        g.p.out_bf16 = hb; g.p.ld_bf16 = stg.cout; g.p.act = ACT_SNAKE; g.p.alpha = ru.alpha2; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
        DAC_GEMM(g);
        if (st) DAC_HALO_OUT(cur, st->ru_halo[3 * b + u], hr, stg.cout, 2, Tc);
      }
      {
        GemmCall g = conv_gemm(ru.conv1, hb, B, Tc, 1);
        g.p.resid = xa; g.p.out_f32 = xa; g.p.ld_f32 = stg.cout;
        g.p.out_bf16 = cur; g.p.ld_bf16 = stg.cout; g.p.act = ACT_SNAKE; g.p.alpha = next_alpha; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
        DAC_GEMM(g);
      }
    }
  }
  const int Cl = h->stage.back().cout;
  dim3 grid((Tc + 255) / 256, B);
  if (st) DAC_HALO_IN(cur, st->final_halo, 6, Cl, 2);
  if (Cl == 96)
    launch_k(final_conv_tanh_rows_kernel<12>, dim3(grid), dim3(256), 0, s, 1, cur, h->final_w, h->final_b, audio, Tc, st ? 6 : 0);
  else
    launch_k(final_conv_tanh_kernel, dim3(grid), dim3(256), 7 * Cl * sizeof(float), s, 1, cur, h->final_w, h->final_b, audio, Tc, Cl, st ? 6 : 0);
  count_launch();
  if (st) DAC_HALO_OUT(cur, st->final_halo, 6, Cl, 2, Tc);
  ECHO_CUDA(cudaGetLastError());
  if (st) st->pos += T;
#undef DAC_HALO_IN
#undef DAC_HALO_OUT
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
  HandleScope scope(h, s);
  const int C = h->dcfg.latent_dim, Kp = h->dcfg.pca_dim;
  float* X = (float*)h->wsget("dac.X", (size_t)B * T * C * 4, s);
  if (!X) { set_error("dac: out of memory"); return ECHO_ERR_CUDA; }
  launch_k(pca_unproject_kernel, dim3(B * T), dim3(256), Kp * sizeof(float), s, 1, z, pca_components, pca_mean, latent_scale, X, Kp, C);
  count_launch();
  return dac_run(h, X, B, T, audio, s);
}

// ------------------------------------------------------------------------------------------------ streaming decode
extern "C" int echo_dac_stream_reset(echo_handle* h, echo_dac_stream* st, void* stream) {
  ECHO_TRY(dac_check(h, "echo_dac_stream_reset"));
  if (!st) { set_error("echo_dac_stream_reset: null stream"); return ECHO_ERR_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HandleScope scope(h, s);
  // halos = zeros (the causal left padding at the start of a sequence); the KV caches need no clearing: rows >= pos are
  // never read
  for (size_t i = 2 * st->K.size(); i < st->bufs.size(); ++i) ECHO_CUDA(cudaMemsetAsync(st->bufs[i].first, 0, st->bufs[i].second, s));
  st->pos = 0;
  return ECHO_OK;
}

extern "C" int echo_dac_stream_create(echo_handle* h, int max_latents, echo_dac_stream** out, void* stream) {
  ECHO_TRY(dac_check(h, "echo_dac_stream_create"));
  if (!out || max_latents <= 0 || max_latents > 4096) { set_error("echo_dac_stream_create: bad argument (max_latents 1..4096)"); return ECHO_ERR_ARG; }
  const echo_dac_config& c = h->dcfg;
  const int C = c.latent_dim;
  echo_dac_stream* st = new echo_dac_stream();
  st->max_T = max_latents;
  bool ok = true;
  auto alloc = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { ok = false; return nullptr; }
    st->bufs.push_back({p, bytes});
    return p;
  };
  for (int i = 0; i < c.post_layers; ++i) st->K.push_back((bf16*)alloc((size_t)max_latents * C * 2));
  for (int i = 0; i < c.post_layers; ++i) st->V.push_back((bf16*)alloc((size_t)max_latents * C * 2));
  // (the first 2 * post_layers entries of bufs are the KV caches -- echo_dac_stream_reset skips them)
  for (int i = 0; i < c.num_upsample; ++i) st->dw_halo.push_back((float*)alloc((size_t)6 * C * 4));
  st->conv0_halo = (bf16*)alloc((size_t)6 * C * 2);
  static const int dil[3] = {1, 3, 9};
  for (int b = 0; b < c.num_rates; ++b) {
    st->convt_halo.push_back((bf16*)alloc((size_t)h->stage[b].cin * 2));
    for (int u = 0; u < 3; ++u) st->ru_halo.push_back((bf16*)alloc((size_t)6 * dil[u] * h->stage[b].cout * 2));
  }
  st->final_halo = (bf16*)alloc((size_t)6 * h->stage.back().cout * 2);
  if (!ok) {
    for (auto& b : st->bufs) cudaFree(b.first);
    delete st;
    set_error("echo_dac_stream_create: out of device memory");
    return ECHO_ERR_CUDA;
  }
  h->dac_streams.insert(st);
  int rc = echo_dac_stream_reset(h, st, stream);
  if (rc != ECHO_OK) return rc;
  *out = st;
  return ECHO_OK;
}

extern "C" int echo_dac_stream_destroy(echo_handle* h, echo_dac_stream* st) {
  if (!h || !st) return ECHO_OK;
  std::lock_guard<std::recursive_mutex> g(h->mu);
  if (!h->dac_streams.erase(st)) { set_error("echo_dac_stream_destroy: unknown stream"); return ECHO_ERR_ARG; }
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& b : st->bufs) cudaFree(b.first);
  delete st;
  return ECHO_OK;
}

extern "C" int echo_dac_stream_decode(echo_handle* h, echo_dac_stream* st, const float* z, const float* pca_components,
                                      const float* pca_mean, float latent_scale, int T, float* audio, void* stream) {
  ECHO_TRY(dac_check(h, "echo_dac_stream_decode"));
  if (!st || !z || !pca_components || !pca_mean || !audio || T <= 0) { set_error("echo_dac_stream_decode: bad argument"); return ECHO_ERR_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HandleScope scope(h, s);
  if (!h->dac_streams.count(st)) { set_error("echo_dac_stream_decode: unknown stream"); return ECHO_ERR_ARG; }
  const int C = h->dcfg.latent_dim, Kp = h->dcfg.pca_dim;
  float* X = (float*)h->wsget("dac.X", (size_t)T * C * 4, s);
  if (!X) { set_error("dac: out of memory"); return ECHO_ERR_CUDA; }
  launch_k(pca_unproject_kernel, dim3(T), dim3(256), Kp * sizeof(float), s, 1, z, pca_components, pca_mean, latent_scale, X, Kp, C);
  count_launch();
  return dac_run(h, X, 1, T, audio, s, st);
}

extern "C" int echo_dac_stream_position(echo_handle* h, echo_dac_stream* st, int* latents_decoded) {
  if (!h || !st || !latents_decoded) { set_error("echo_dac_stream_position: bad argument"); return ECHO_ERR_ARG; }
  *latents_decoded = st->pos;
  return ECHO_OK;
}

extern "C" int echo_dac_decode_zq(echo_handle* h, const float* zq, int B, int T, float* audio, void* stream) {
  ECHO_TRY(dac_check(h, "echo_dac_decode_zq"));
  if (!zq || !audio || B <= 0 || T <= 0) { set_error("echo_dac_decode_zq: bad argument"); return ECHO_ERR_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HandleScope scope(h, s);
  const int C = h->dcfg.latent_dim;
  float* X = (float*)h->wsget("dac.X", (size_t)B * T * C * 4, s);
  if (!X) { set_error("dac: out of memory"); return ECHO_ERR_CUDA; }
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_k(transpose_ct_kernel, dim3(grid), dim3(block), 0, s, 1, zq, X, C, T);
  count_launch();
  return dac_run(h, X, B, T, audio, s);
}

// ------------------------------------------------------------------------------------------------ encode
namespace {

template <int NV>
void launch_rmsnorm_out(const float* X, const float* w, const float* alpha, float* of, bf16* ob, int rows, float eps,
                        cudaStream_t s) {
  launch_k(rmsnorm_out_kernel<NV>, dim3((rows + 7) / 8), dim3(256), 0, s, 1, X, w, alpha, of, ob, rows, eps);
  count_launch();
}
int rmsnorm_out(const float* X, const float* w, const float* alpha, float* of, bf16* ob, int rows, int C, float eps,
                cudaStream_t s) {
  switch (C) {
    case 256: launch_rmsnorm_out<2>(X, w, alpha, of, ob, rows, eps, s); break;
    case 512: launch_rmsnorm_out<4>(X, w, alpha, of, ob, rows, eps, s); break;
    case 1024: launch_rmsnorm_out<8>(X, w, alpha, of, ob, rows, eps, s); break;
    case 2048: launch_rmsnorm_out<16>(X, w, alpha, of, ob, rows, eps, s); break;
    default: set_error("dac encode: unsupported transformer width %d", C); return ECHO_ERR_ARG;
  }
  return ECHO_OK;
}

// Encoder + quantizer front + RVQ (autoencoder.py:903-929, 452-463, 1116-1126). audio (B, 1, L) fp32; writes
// zq_rows (B*T, C) fp32 time-major and optionally codes (B, 1 + n_codebooks, T).
int dac_encode_run(echo_handle* h, const float* audio, int B, int L, float* zq_rows, int32_t* codes, float* z_pre_out,
                   cudaStream_t s) {
  const echo_dac_config& c = h->dcfg;
  const int C = c.latent_dim, I = c.post_intermediate, H = c.post_heads, nb = c.num_enc_rates;
  int hop = 1;
  for (int i = 0; i < nb; ++i) hop *= c.enc_rates[i];
  const int frame = hop << c.num_upsample;
  if (L <= 0 || L % frame) { set_error("dac encode: L=%d must be a positive multiple of the frame length %d", L, frame); return ECHO_ERR_ARG; }
  const int Tenc = L / hop, T = L / frame;
  if (Tenc > 4096) { set_error("dac encode: %d encoder frames exceed the RoPE table (4096); encode in chunks", Tenc); return ECHO_ERR_ARG; }
  // largest activation: stage 0 holds L x enc_dim, every later stage half of it or less; transformer buffers separately
  int64_t max_elems = (int64_t)B * L * c.enc_dim;
  max_elems = std::max<int64_t>(max_elems, (int64_t)B * Tenc * C * 4);  // ConvNeXt hidden / SwiGLU hidden
  float* xa = (float*)h->wsget("dac.xa", (size_t)max_elems * 4, s);
  bf16* sa = (bf16*)h->wsget("dac.sa", (size_t)max_elems * 2, s);
  bf16* sb = (bf16*)h->wsget("dac.sb", (size_t)max_elems * 2, s);
  bf16* hb = (bf16*)h->wsget("dac.hb", (size_t)max_elems * 2, s);
  const int64_t trow = (int64_t)B * Tenc;
  bf16* XN = (bf16*)h->wsget("dac.XN", (size_t)trow * C * 2, s);
  bf16* Q = (bf16*)h->wsget("dac.Q", (size_t)trow * C * 2, s);
  bf16* K = (bf16*)h->wsget("dac.K", (size_t)trow * C * 2, s);
  bf16* V = (bf16*)h->wsget("dac.V", (size_t)trow * C * 2, s);
  bf16* AO = (bf16*)h->wsget("dac.AO", (size_t)trow * C * 2, s);
  bf16* Hh = (bf16*)h->wsget("dac.Hh", (size_t)trow * 3 * C * 2, s);
  float* xq = (float*)h->wsget("dac.xq", (size_t)trow * C * 4, s);  // fp32 stream of the transformers / quantizer input
  if (!xa || !sa || !sb || !hb || !XN || !Q || !K || !V || !AO || !Hh || !xq) { set_error("dac encode: workspace allocation failed"); return ECHO_ERR_CUDA; }
  TfBuffers tb{XN, Q, K, V, AO, Hh};

  // ---- Encoder: conv7 (Cin = 1) fused with the first Snake
  int Tc = L, d = c.enc_dim;
  bf16* cur = sa;
  bf16* nxt = sb;
  {
    dim3 grid((Tc + 255) / 256, B);
    launch_k(enc_conv0_kernel, grid, dim3(256), (size_t)9 * d * sizeof(float), s, 1, audio, h->enc_conv0_w, h->enc_conv0_b,
             h->enc_blk[0].ru[0].alpha1, xa, cur, Tc, d);
    count_launch();
  }
  static const int dil[3] = {1, 3, 9};
  for (int b = 0; b < nb; ++b) {
    const DacEncBlockW& eb = h->enc_blk[b];
    for (int u = 0; u < 3; ++u) {  // ResidualUnit (autoencoder.py:879-900): cur = snake1(x) on entry
      const DacResUnitW& ru = eb.ru[u];
      {
        GemmCall g = conv_gemm(ru.conv7, cur, B, Tc, dil[u]);
        g.p.out_bf16 = hb; g.p.ld_bf16 = d; g.p.act = ACT_SNAKE; g.p.alpha = ru.alpha2; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
        DAC_GEMM(g);
      }
      {
        const float* next_alpha = (u < 2) ? eb.ru[u + 1].alpha1 : eb.alpha_out;
        GemmCall g = conv_gemm(ru.conv1, hb, B, Tc, 1);
        g.p.resid = xa; g.p.out_f32 = xa; g.p.ld_f32 = d;
        g.p.out_bf16 = cur; g.p.ld_bf16 = d; g.p.act = ACT_SNAKE; g.p.alpha = next_alpha; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
        DAC_GEMM(g);
      }
    }
    // strided conv k = 2s (left pad s): out[t] = W[:, :, 0:s] x[s t - s ... s t - 1] + W[:, :, s:2s] x[s t ... s t + s - 1]
    // == 2-tap GEMM over the (Tc / s, s * d) view of cur with row shifts {-1, 0}
    const int st = eb.stride, To = Tc / st;
    const bool last = (b == nb - 1);
    {
      DacConvW w2 = eb.down;
      w2.cin = st * d; w2.taps = 2;
      GemmCall g = base_gemm(cur, (int64_t)st * d, w2.w, (int64_t)2 * st * d, B, To, eb.cout, st * d, 2);
      g.p.tap_shift[0] = -1; g.p.tap_shift[1] = 0;
      g.p.bias = w2.bias; g.p.col_mod = eb.cout;
      if (last && !h->enc_tf.empty()) {
        g.p.out_f32 = xq; g.p.ld_f32 = eb.cout;  // fp32 stream of the block's transformer
      } else {
        const float* next_alpha = last ? h->enc_alpha_out : h->enc_blk[b + 1].ru[0].alpha1;
        g.p.out_f32 = xa; g.p.ld_f32 = eb.cout;
        g.p.out_bf16 = nxt; g.p.ld_bf16 = eb.cout; g.p.act = ACT_SNAKE; g.p.alpha = next_alpha; g.p.alpha_inv = h->dac_alpha_inv[g.p.alpha];
      }
      DAC_GEMM(g);
    }
    Tc = To; d = eb.cout;
    std::swap(cur, nxt);
  }
  // ---- last block's window-512 transformer, final norm fused with the encoder's closing Snake
  if (!h->enc_tf.empty()) {
    ECHO_TRY(run_window_transformer(h, h->enc_tf, xq, B, Tc, C, 3 * C, C / 64, c.enc_window, 1e-5f, tb, s));
    ECHO_TRY(rmsnorm_out(xq, h->enc_tf_norm, h->enc_alpha_out, nullptr, cur, B * Tc, C, 1e-5f, s));
  }
  {  // closing conv3 (causal) -> z_enc: bf16 for the downsample GEMM
    GemmCall g = conv_gemm(h->enc_conv_out, cur, B, Tc, 1);
    g.p.out_bf16 = nxt; g.p.ld_bf16 = C;
    DAC_GEMM(g);
  }
  std::swap(cur, nxt);
  // ---- quantizer.downsample: [conv k2 s2 ; ConvNeXt] per factor (autoencoder.py:418-424, 360-373)
  for (int i = 0; i < c.num_upsample; ++i) {
    const DacUpW& u = h->down[i];
    const int To = Tc / 2, r2 = B * To;
    {
      GemmCall g = base_gemm(cur, (int64_t)2 * C, u.convt.w, (int64_t)2 * C, B, To, C, 2 * C, 1);
      g.p.bias = u.convt.bias; g.p.col_mod = C;
      g.p.out_f32 = xa; g.p.ld_f32 = C;
      DAC_GEMM(g);
    }
    Tc = To;
    launch_k(dwconv_ln_kernel, dim3(r2), dim3(256), 0, s, 1, xa, u.dw_w, u.dw_b, u.ln_w, u.ln_b, nxt, Tc, C, 0);
    count_launch();
    {
      GemmCall g = base_gemm(nxt, C, u.w1, C, 1, r2, 4 * C, C, 1);
      g.p.bias = u.b1; g.p.out_bf16 = hb; g.p.ld_bf16 = 4 * C; g.p.act = ACT_GELU;
      DAC_GEMM(g);
    }
    {
      const bool lastd = (i == c.num_upsample - 1);
      GemmCall g = base_gemm(hb, 4 * C, u.w2, 4 * C, 1, r2, C, 4 * C, 1);
      g.p.bias = u.b2; g.p.gate = u.gamma; g.p.resid = xa; g.p.out_f32 = lastd ? xq : xa; g.p.ld_f32 = C;
      if (!lastd) { g.p.out_bf16 = cur; g.p.ld_bf16 = C; }  // bf16 copy feeds the next strided conv
      g.split_k = 1;
      DAC_GEMM(g);
    }
  }
  if (Tc != T) { set_error("dac encode: internal length mismatch %d vs %d", Tc, T); return ECHO_ERR_ARG; }
  // ---- quantizer.pre_module + final norm in fp32 (the quantizers work on the un-rounded stream)
  ECHO_TRY(run_window_transformer(h, h->pre, xq, B, T, C, I, H, c.post_window, c.post_norm_eps, tb, s));
  ECHO_TRY(rmsnorm_out(xq, h->pre_final_norm, nullptr, xa, nullptr, B * T, C, c.post_norm_eps, s));
  if (z_pre_out) ECHO_CUDA(cudaMemcpyAsync(z_pre_out, xa, (size_t)B * T * C * 4, cudaMemcpyDeviceToDevice, s));
  // ---- semantic + residual VQ
  VqAll va;
  va.n = (int)h->vq.size();
  for (int i = 0; i < va.n; ++i) {
    const DacVqW& q = h->vq[i];
    va.q[i] = VqDev{q.in_w, q.in_b, q.cb_norm, q.cb_sq, q.out_table, q.size};
  }
  switch (C) {
    case 256: launch_k(vq_encode_kernel<1>, dim3(B * T), dim3(256), 0, s, 1, xa, va, zq_rows, codes, T, c.codebook_dim); break;
    case 512: launch_k(vq_encode_kernel<2>, dim3(B * T), dim3(256), 0, s, 1, xa, va, zq_rows, codes, T, c.codebook_dim); break;
    case 1024: launch_k(vq_encode_kernel<4>, dim3(B * T), dim3(256), 0, s, 1, xa, va, zq_rows, codes, T, c.codebook_dim); break;
    default: set_error("dac encode: unsupported latent_dim %d", C); return ECHO_ERR_ARG;
  }
  count_launch();
  ECHO_CUDA(cudaGetLastError());
  return ECHO_OK;
}

int dac_enc_check(echo_handle* h, const char* who) {
  ECHO_TRY(dac_check(h, who));
  if (!h->dac_enc_ready) { set_error("%s: the DAC encoder / quantizer weights were not loaded", who); return ECHO_ERR_STATE; }
  return ECHO_OK;
}

}  // namespace

extern "C" int echo_dac_encode_zq(echo_handle* h, const float* audio, int B, int L, float* zq, int32_t* codes, float* z_pre,
                                  void* stream) {
  ECHO_TRY(dac_enc_check(h, "echo_dac_encode_zq"));
  if (!audio || !zq || B <= 0) { set_error("echo_dac_encode_zq: bad argument"); return ECHO_ERR_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HandleScope scope(h, s);
  const int C = h->dcfg.latent_dim;
  int hop = 1;
  for (int i = 0; i < h->dcfg.num_enc_rates; ++i) hop *= h->dcfg.enc_rates[i];
  const int frame = hop << h->dcfg.num_upsample;
  if (L <= 0 || L % frame) { set_error("echo_dac_encode_zq: L=%d is not a positive multiple of %d", L, frame); return ECHO_ERR_ARG; }
  const int T = L / frame;
  float* rows = (float*)h->wsget("dac.zq_rows", (size_t)B * T * C * 4, s);
  if (!rows) { set_error("out of memory"); return ECHO_ERR_CUDA; }
  ECHO_TRY(dac_encode_run(h, audio, B, L, rows, codes, z_pre, s));
  // (B, T, C) -> (B, C, T): the same tiled transpose with the roles of the two inner dims swapped
  dim3 block(32, 8), grid((C + 31) / 32, (T + 31) / 32, B);
  launch_k(transpose_ct_kernel, grid, block, 0, s, 1, (const float*)rows, zq, T, C);
  count_launch();
  ECHO_CUDA(cudaGetLastError());
  return ECHO_OK;
}

extern "C" int echo_dac_encode(echo_handle* h, const float* audio, const float* pca_components, const float* pca_mean,
                               float latent_scale, int B, int L, float* latent, void* stream) {
  ECHO_TRY(dac_enc_check(h, "echo_dac_encode"));
  if (!audio || !pca_components || !pca_mean || !latent || B <= 0) { set_error("echo_dac_encode: bad argument"); return ECHO_ERR_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HandleScope scope(h, s);
  const int C = h->dcfg.latent_dim, Kp = h->dcfg.pca_dim;
  int hop = 1;
  for (int i = 0; i < h->dcfg.num_enc_rates; ++i) hop *= h->dcfg.enc_rates[i];
  const int frame = hop << h->dcfg.num_upsample;
  if (L <= 0 || L % frame) { set_error("echo_dac_encode: L=%d is not a positive multiple of %d", L, frame); return ECHO_ERR_ARG; }
  const int T = L / frame;
  float* rows = (float*)h->wsget("dac.zq_rows", (size_t)B * T * C * 4, s);
  if (!rows) { set_error("out of memory"); return ECHO_ERR_CUDA; }
  ECHO_TRY(dac_encode_run(h, audio, B, L, rows, nullptr, nullptr, s));
  launch_k(pca_project_kernel, dim3(B * T), dim3(256), (size_t)C * sizeof(float), s, 1, (const float*)rows, pca_components,
           pca_mean, latent_scale, latent, C, Kp);
  count_launch();
  ECHO_CUDA(cudaGetLastError());
  return ECHO_OK;
}
