// Bandwidth-bound glue of the DiT step: every kernel touches each element once, uses 128-bit accesses and
// warp-shuffle reductions. See glue.h for the reference lines each one replaces.
#include "glue.h"

#include "counters.h"
#include "launch.h"
#include "profiler.h"

namespace echo {

std::atomic<int64_t> g_launches{0};
int64_t glue_launch_count() { return g_launches.load(); }

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for <= 1024 threads; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum_f(v);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum_f(t);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void embed_kernel(const int32_t* __restrict__ ids, const bf16* __restrict__ table, float* __restrict__ X,
                             int E, int vocab) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x;
  int id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const bf16* src = table + (size_t)id * E;
  float* dst = X + (size_t)r * E;
  for (int c = threadIdx.x * 2; c < E; c += blockDim.x * 2) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + c));
    *reinterpret_cast<float2*>(dst + c) = v;
  }
}

// one block (256 threads) per row, W <= 4096, W % 4 == 0
__global__ void __launch_bounds__(256) rmsnorm_affine_kernel(const float* __restrict__ X, bf16* __restrict__ out,
                                                             const float* __restrict__ a, const float* __restrict__ c0,
                                                             int W, int rows_per_group, int64_t group_ld, float eps) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[32];
  const int r = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(X + (size_t)r * W);
  const int n4 = W >> 2;
  float4 v[4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < n4) {
      v[i] = xr[c];
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = block_sum(ss, sh);
  const float rstd = rsqrtf(ss / (float)W + eps);
  const size_t g = rows_per_group > 0 ? (size_t)(r / rows_per_group) * group_ld : 0;
  const float4* ap = reinterpret_cast<const float4*>(a + g);
  const float4* cp = c0 ? reinterpret_cast<const float4*>(c0 + g) : nullptr;
  uint2* op = reinterpret_cast<uint2*>(out + (size_t)r * W);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < n4) {
      const float4 av = __ldg(ap + c);
      float4 o;
      o.x = v[i].x * rstd * av.x; o.y = v[i].y * rstd * av.y; o.z = v[i].z * rstd * av.z; o.w = v[i].w * rstd * av.w;
      if (cp) {
        const float4 cv = __ldg(cp + c);
        o.x += cv.x; o.y += cv.y; o.z += cv.z; o.w += cv.w;
      }
      op[c] = make_uint2(pack2(o.x, o.y), pack2(o.z, o.w));
    }
  }
}

// One WARP per row, NV float4 per lane (W = 128 * NV): the whole row lives in registers between the two passes,
// 16 independent 16-byte loads are in flight per lane, and the reduction is five shuffles -- no block barrier.
// (The block-per-row version above moved 2.5 TB/s at M = 1920: two __syncthreads and 2 loads in flight per thread.)
template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_affine_warp_kernel(const float* __restrict__ X, bf16* __restrict__ out,
                                                                  const float* __restrict__ a,
                                                                  const float* __restrict__ c0, int rows,
                                                                  int rows_per_group, int64_t group_ld, float eps,
                                                                  const float* __restrict__ parts, int nparts,
                                                                  int64_t part_stride) {
  constexpr int W = 128 * NV;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const bool live = r < rows;
  // The modulation rows (a = 1 + scale, c0 = shift: AdaLN tables built once per request, or static norm weights) do
  // not depend on the preceding kernel: they are in registers before the dependency wait returns, so that behind it
  // there is one L2 round trip (the row of X) instead of two. Contract: a / c0 are not written by an immediate
  // predecessor that triggers its dependents early.
  float4 av[NV], cv[NV];
  if (live) {
    const size_t g = rows_per_group > 0 ? (size_t)(r / rows_per_group) * group_ld : 0;
    const float4* ap = reinterpret_cast<const float4*>(a + g);
#pragma unroll
    for (int i = 0; i < NV; ++i) av[i] = __ldg(ap + lane + 32 * i);
    if (c0) {
      const float4* cp = reinterpret_cast<const float4*>(c0 + g);
#pragma unroll
      for (int i = 0; i < NV; ++i) cv[i] = __ldg(cp + lane + 32 * i);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) cv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  pdl_wait();
  pdl_trigger();
  if (!live) return;
  const float4* xr = reinterpret_cast<const float4*>(X + (size_t)r * W);
  float4 v[NV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  if (nparts > 0) {
    // the split-K planes of the wo / w2 GEMM right before this kernel: added in plane order (bit-reproducible), and the
    // updated residual row goes back to X
    for (int k = 0; k < nparts; ++k) {
      const float4* pr = reinterpret_cast<const float4*>(parts + (size_t)k * part_stride + (size_t)r * W);
#pragma unroll
      for (int i = 0; i < NV; ++i) {  // (the compiler batches these loads as far as the 255 registers allow)
        const float4 q = pr[lane + 32 * i];
        v[i].x += q.x; v[i].y += q.y; v[i].z += q.z; v[i].w += q.w;
      }
    }
    float4* xw = reinterpret_cast<float4*>(const_cast<float*>(X) + (size_t)r * W);
#pragma unroll
    for (int i = 0; i < NV; ++i) xw[lane + 32 * i] = v[i];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  ss = warp_sum_f(ss);
  const float rstd = rsqrtf(ss / (float)W + eps);
  uint2* op = reinterpret_cast<uint2*>(out + (size_t)r * W);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 o;
    o.x = fmaf(v[i].x * rstd, av[i].x, cv[i].x); o.y = fmaf(v[i].y * rstd, av[i].y, cv[i].y);
    o.z = fmaf(v[i].z * rstd, av[i].z, cv[i].z); o.w = fmaf(v[i].w * rstd, av[i].w, cv[i].w);
    op[lane + 32 * i] = make_uint2(pack2(o.x, o.y), pack2(o.z, o.w));
  }
}

// Variant for folding split-K planes (see GemmCall::part_ws): TWO warps per row, so that a lane holds only half as many
// values per tensor and can have the row of X AND all (<= 4) planes in flight at once -- one L2 round trip behind the
// dependency wait, like the plain kernel (with one warp per row the planes came in 4-7 dependent batches and the
// fold cost more than the atomics it replaced). 4 rows per block of 256 threads; W = 256 * NH.
template <int NH>
__global__ void __launch_bounds__(256) rmsnorm_affine_planes_kernel(float* __restrict__ X, bf16* __restrict__ out,
                                                                    const float* __restrict__ a,
                                                                    const float* __restrict__ c0, int rows,
                                                                    int rows_per_group, int64_t group_ld, float eps,
                                                                    const float* __restrict__ parts, int nparts,
                                                                    int64_t part_stride) {
  constexpr int W = 256 * NH;
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 4 + (warp >> 1);
  const int hf = warp & 1;  // which half of the row
  const bool live = r < rows;
  float4 av[NH], cv[NH];
  if (live) {
    const size_t g = (rows_per_group > 0 ? (size_t)(r / rows_per_group) * group_ld : 0) + (size_t)hf * (W / 2);
    const float4* ap = reinterpret_cast<const float4*>(a + g);
#pragma unroll
    for (int i = 0; i < NH; ++i) av[i] = __ldg(ap + lane + 32 * i);
    if (c0) {
      const float4* cp = reinterpret_cast<const float4*>(c0 + g);
#pragma unroll
      for (int i = 0; i < NH; ++i) cv[i] = __ldg(cp + lane + 32 * i);
    } else {
#pragma unroll
      for (int i = 0; i < NH; ++i) cv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  pdl_wait();
  pdl_trigger();
  float4 v[NH];
  float ss = 0.f;
  if (live) {
    const size_t off = (size_t)r * W + (size_t)hf * (W / 2);
    const float4* xr = reinterpret_cast<const float4*>(X + off);
    float4 q[4][NH];
#pragma unroll
    for (int i = 0; i < NH; ++i) v[i] = xr[lane + 32 * i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nparts) {
        const float4* pr = reinterpret_cast<const float4*>(parts + (size_t)k * part_stride + off);
#pragma unroll
        for (int i = 0; i < NH; ++i) q[k][i] = pr[lane + 32 * i];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nparts) {
#pragma unroll
        for (int i = 0; i < NH; ++i) { v[i].x += q[k][i].x; v[i].y += q[k][i].y; v[i].z += q[k][i].z; v[i].w += q[k][i].w; }
      }
    }
    float4* xw = reinterpret_cast<float4*>(X + off);
#pragma unroll
    for (int i = 0; i < NH; ++i) {
      xw[lane + 32 * i] = v[i];
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    ss = warp_sum_f(ss);
  }
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (!live) return;
  const float rstd = rsqrtf((red[warp & ~1] + red[warp | 1]) / (float)W + eps);  // fixed order: same value in both warps
  uint2* op = reinterpret_cast<uint2*>(out + (size_t)r * W + (size_t)hf * (W / 2));
#pragma unroll
  for (int i = 0; i < NH; ++i) {
    float4 o;
    o.x = fmaf(v[i].x * rstd, av[i].x, cv[i].x); o.y = fmaf(v[i].y * rstd, av[i].y, cv[i].y);
    o.z = fmaf(v[i].z * rstd, av[i].z, cv[i].z); o.w = fmaf(v[i].w * rstd, av[i].w, cv[i].w);
    op[lane + 32 * i] = make_uint2(pack2(o.x, o.y), pack2(o.z, o.w));
  }
}

// 8 rows per block, 256 threads, K <= 128 (K % 8 == 0)
__global__ void __launch_bounds__(256) in_proj_kernel(const float* __restrict__ x, const bf16* __restrict__ W,
                                                      const float* __restrict__ bias, float* __restrict__ X, int rows,
                                                      int K, int D, int copies) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sx[8][128];
  const int r0 = blockIdx.x * 8;
  for (int i = threadIdx.x; i < 8 * K; i += 256) {
    const int rr = i / K, k = i % K;
    sx[rr][k] = (r0 + rr < rows) ? x[(size_t)(r0 + rr) * K + k] : 0.f;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < D; n += 256) {
    float acc[8];
    const float bv = bias ? bias[n] : 0.f;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) acc[rr] = bv;
    const uint4* wr = reinterpret_cast<const uint4*>(W + (size_t)n * K);
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint4 w4 = __ldg(wr + k8);
      const uint32_t* wu = reinterpret_cast<const uint32_t*>(&w4);
      float wf[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wu[j]));
        wf[2 * j] = f.x; wf[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[rr] = fmaf(sx[rr][k8 * 8 + j], wf[j], acc[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      if (r0 + rr < rows) {
        for (int c = 0; c < copies; ++c) X[((size_t)c * rows + r0 + rr) * D + n] = acc[rr];
      }
    }
  }
}

// 4 rows per block (256 threads = 8 warps); normalised rows staged in smem; D <= 2048*2
template <int RB>
__global__ void __launch_bounds__(256) out_norm_proj_kernel(const float* __restrict__ X, const float* __restrict__ wn,
                                                            const bf16* __restrict__ Wout, const float* __restrict__ bias,
                                                            float* __restrict__ v, int rows, int D, int Nout, float eps) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sxn[];  // [RB][D]
  __shared__ float sh[32];
  const int r0 = blockIdx.x * RB;
  for (int rr = 0; rr < RB; ++rr) {
    const int r = r0 + rr;
    float ss = 0.f;
    if (r < rows) {
      for (int c = threadIdx.x; c < D; c += 256) {
        const float t = X[(size_t)r * D + c];
        sxn[rr * D + c] = t;
        ss += t * t;
      }
    }
    ss = block_sum(ss, sh);
    const float rstd = rsqrtf(ss / (float)D + eps);
    if (r < rows)
      for (int c = threadIdx.x; c < D; c += 256) sxn[rr * D + c] *= rstd * wn[c];
    else
      for (int c = threadIdx.x; c < D; c += 256) sxn[rr * D + c] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < Nout; n += 8) {
    float acc[RB];
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) acc[rr] = 0.f;
    const uint4* wr = reinterpret_cast<const uint4*>(Wout + (size_t)n * D);
    for (int k8 = lane; k8 < D / 8; k8 += 32) {
      const uint4 w4 = __ldg(wr + k8);
      const uint32_t* wu = reinterpret_cast<const uint32_t*>(&w4);
      float wf[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wu[j]));
        wf[2 * j] = f.x; wf[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sxn[rr * D + k8 * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sxn[rr * D + k8 * 8 + 4]);
        acc[rr] += a0.x * wf[0] + a0.y * wf[1] + a0.z * wf[2] + a0.w * wf[3] + a1.x * wf[4] + a1.y * wf[5] +
                   a1.z * wf[6] + a1.w * wf[7];
      }
    }
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      const float t = warp_sum_f(acc[rr]);
      if (lane == 0 && r0 + rr < rows) v[(size_t)(r0 + rr) * Nout + n] = t + (bias ? bias[n] : 0.f);
    }
  }
}

__global__ void timestep_embed_kernel(const float* __restrict__ t, const float* __restrict__ freqs, bf16* __restrict__ emb,
                                      int half, int round_t) {
  pdl_wait();
  pdl_trigger();
  const int j = blockIdx.x;
  float tv = t[j];
  if (round_t) tv = __bfloat162float(__float2bfloat16_rn(tv));
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = tv * freqs[i];
    emb[(size_t)j * 2 * half + i] = __float2bfloat16_rn(cosf(a));
    emb[(size_t)j * 2 * half + half + i] = __float2bfloat16_rn(sinf(a));
  }
}

__global__ void adaln_prep_kernel(const float* __restrict__ cond, bf16* __restrict__ scond, int n, int D) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = (int64_t)3 * n * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const int j = (int)((i / D) % n);
    const int p = (int)(i / ((int64_t)D * n));
    const float x = cond[(size_t)j * 3 * D + (size_t)p * D + c];
    scond[i] = __float2bfloat16_rn(x / (1.f + __expf(-x)));
  }
}

__global__ void adaln_finish_kernel(const float* __restrict__ up, const float* __restrict__ cond, float* __restrict__ mod,
                                    int n, int D, int Q) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = (int64_t)3 * Q * n * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const int j = (int)((i / D) % n);
    const int p = (int)(i / ((int64_t)D * n * Q));
    float v = up[i] + cond[(size_t)j * 3 * D + (size_t)p * D + c];
    if (p == 1) v += 1.f;
    else if (p == 2) v = tanhf(v);
    mod[i] = v;
  }
}

__global__ void cfg_euler_kernel(float* __restrict__ x, const float* __restrict__ v, int64_t n, int has_cfg, float s_text,
                                 float s_spk, int has_rescale, float omt, float ratio, float dt) {
  pdl_wait();
  pdl_trigger();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float vp = v[i];
    if (has_cfg) {
      const float vt = v[n + i], vs = v[2 * n + i];
      vp = vp + s_text * (vp - vt) + s_spk * (vp - vs);
    }
    const float xv = x[i];
    if (has_rescale) vp = 1.f / omt * (ratio * (omt * vp + xv) - xv);
    x[i] = xv + vp * dt;
  }
}


__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, float sc) {
  pdl_wait();
  pdl_trigger();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * sc;
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
  pdl_wait();
  pdl_trigger();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void mask_eff_len_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ eff, int len, int ld,
                                    int stride) {
  pdl_wait();
  pdl_trigger();
  __shared__ int best;
  if (threadIdx.x == 0) best = 0;
  __syncthreads();
  const int j = blockIdx.x;
  int loc = 0;
  for (int i = threadIdx.x; i < len; i += blockDim.x)
    if (mask[(size_t)j * ld + (size_t)i * stride]) loc = i + 1;
  atomicMax(&best, loc);
  __syncthreads();
  if (threadIdx.x == 0) eff[j] = best;
}

__global__ void pack_rows_kernel(const void* __restrict__ src, int src_bf16, void* __restrict__ dst, int dst_bf16,
                                 int64_t rows, int64_t cols, int64_t dst_ld, int64_t blk, int64_t blk_stride,
                                 int64_t blk_off) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    const float v = src_bf16 ? __bfloat162float(static_cast<const bf16*>(src)[i]) : static_cast<const float*>(src)[i];
    const int64_t dr = (r / blk) * blk_stride + blk_off + (r % blk);
    if (dst_bf16) static_cast<bf16*>(dst)[dr * dst_ld + c] = __float2bfloat16_rn(v);
    else static_cast<float*>(dst)[dr * dst_ld + c] = v;
  }
}

inline int grid_for(int64_t n, int threads = 256) {
  int64_t g = (n + threads - 1) / threads;
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

void embed_rows(const int32_t* ids, const bf16* table, float* X, int rows, int E, int vocab, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(embed_kernel, dim3(rows), dim3(128), 0, s, 1, ids, table, X, E, vocab);
  count_launch();
}
void rmsnorm_affine(const float* X, bf16* out, const float* a, const float* c0, int rows, int W, int rows_per_group,
                    int64_t group_ld, float eps, cudaStream_t s, const float* parts, int nparts, int64_t part_stride) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  if (nparts > 0 && W == 2048 && nparts <= 4) {
    launch_k(rmsnorm_affine_planes_kernel<8>, dim3((rows + 3) / 4), dim3(256), 0, s, 1, const_cast<float*>(X), out, a, c0,
             rows, rows_per_group, group_ld, eps, parts, nparts, part_stride);
    count_launch();
    return;
  }
  const dim3 grid((rows + 7) / 8), block(256);
  switch (W) {
#define ECHO_RMS_CASE(NV)                                                                                       \
  case 128 * NV:                                                                                                \
    launch_k(rmsnorm_affine_warp_kernel<NV>, grid, block, 0, s, 1, X, out, a, c0, rows, rows_per_group, group_ld, eps, \
             parts, nparts, part_stride);                                                                           \
    break;
    ECHO_RMS_CASE(2) ECHO_RMS_CASE(4) ECHO_RMS_CASE(8) ECHO_RMS_CASE(10) ECHO_RMS_CASE(16)
#undef ECHO_RMS_CASE
    default:
      launch_k(rmsnorm_affine_kernel, dim3(rows), dim3(256), 0, s, 1, X, out, a, c0, W, rows_per_group, group_ld, eps);
  }
  count_launch();
}
void in_proj(const float* x, const bf16* W, const float* bias, float* X, int rows, int K, int D, int copies,
             cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(in_proj_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, 1, x, W, bias, X, rows, K, D, copies);
  count_launch();
}
void out_norm_proj(const float* X, const float* wn, const bf16* Wout, const float* bias, float* v, int rows, int D,
                   int Nout, float eps, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  constexpr int RB = 4;
  launch_k(out_norm_proj_kernel<RB>, dim3((rows + RB - 1) / RB), dim3(256), RB * D * sizeof(float), s, 1, X, wn, Wout, bias, v, rows, D, Nout,
                                                                                    eps);
  count_launch();
}
void timestep_embed(const float* t, const float* freqs, bf16* emb, int n, int half, int round_t_bf16, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(timestep_embed_kernel, dim3(n), dim3(128), 0, s, 1, t, freqs, emb, half, round_t_bf16);
  count_launch();
}
void adaln_prep(const float* cond, bf16* scond, int n, int D, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(adaln_prep_kernel, dim3(grid_for((int64_t)3 * n * D)), dim3(256), 0, s, 1, cond, scond, n, D);
  count_launch();
}
void adaln_finish(const float* up, const float* cond, float* mod, int n, int D, int Q, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(adaln_finish_kernel, dim3(grid_for((int64_t)3 * Q * n * D)), dim3(256), 0, s, 1, up, cond, mod, n, D, Q);
  count_launch();
}
void cfg_euler_update(float* x, const float* v, int64_t n, int has_cfg, float s_text, float s_spk, int has_rescale,
                      float one_minus_t, float ratio, float dt, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(cfg_euler_kernel, dim3(grid_for(n)), dim3(256), 0, s, 1, x, v, n, has_cfg, s_text, s_spk, has_rescale, one_minus_t, ratio, dt);
  count_launch();
}
// ---- find_flattening_point (inference.py:288-296): first i whose `window` rows i .. i+window-1 of the zero-padded
// (T + window, C) latent are flat: unbiased std over all window*C elements < thr and |mean - target| < 0.1.
// One warp per window; sums in fp64 (the decision is a threshold test: it must not depend on summation order), then
// rounded to fp32 before the comparisons exactly as torch compares its fp32 0-dim results with the Python scalars.
__global__ void flat_init_kernel(int32_t* out, int T) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) *out = T;
}
__global__ void flat_scan_kernel(const float* __restrict__ x, int T, int C, int window, float target, float thr,
                                 int32_t* out) {
  pdl_wait();
  pdl_trigger();
  const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= T) return;
  const int r_hi = w + window < T ? w + window : T;  // rows >= T are the zero padding
  const int64_t n_valid = (int64_t)(r_hi - w) * C;
  const float* p = x + (int64_t)w * C;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = lane; i < n_valid; i += 32) {
    const double v = (double)p[i];
    s1 += v;
    s2 += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    const double n = (double)window * (double)C;
    const double mean = s1 / n;
    double var = n > 1.0 ? (s2 - s1 * mean) / (n - 1.0) : 0.0;
    if (var < 0.0) var = 0.0;
    const float sd = (float)sqrt(var), m = (float)mean;
    if (sd < thr && fabsf(m - target) < 0.1f) atomicMin(out, w);
  }
}
void flattening_point(const float* latent, int T, int C, int window, float target, float std_threshold, int32_t* out,
                      cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(flat_init_kernel, dim3(1), dim3(32), 0, s, 1, out, T);
  count_launch();
  if (T > 0) {
    launch_k(flat_scan_kernel, dim3((T + 7) / 8), dim3(256), 0, s, 1, latent, T, C, window, target, std_threshold, out);
    count_launch();
  }
}

void scale_copy_f32(const float* src, float* dst, int64_t n, float sc, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(scale_copy_kernel, dim3(grid_for(n)), dim3(256), 0, s, 1, src, dst, n, sc);
  count_launch();
}
void cast_f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(cast_bf16_kernel, dim3(grid_for(n)), dim3(256), 0, s, 1, src, dst, n);
  count_launch();
}
void mask_eff_len(const uint8_t* mask, int32_t* eff, int n, int len, int ld, int stride, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(mask_eff_len_kernel, dim3(n), dim3(256), 0, s, 1, mask, eff, len, ld, stride);
  count_launch();
}
void pack_rows(const void* src, int src_is_bf16, void* dst, int dst_is_bf16, int64_t rows, int64_t cols, int64_t dst_ld,
               int64_t blk, int64_t blk_stride, int64_t blk_off, cudaStream_t s) {
  ProfScope ps(PROF_GLUE, 0.0, 0.0, s, __func__);
  launch_k(pack_rows_kernel, dim3(grid_for(rows * cols)), dim3(256), 0, s, 1, src, src_is_bf16, dst, dst_is_bf16, rows, cols, dst_ld, blk,
                                                         blk_stride, blk_off);
  count_launch();
}

}  // namespace echo
