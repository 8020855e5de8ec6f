// Attention over segmented keys, flash-style (online softmax), bf16 in / fp32 accumulate.
//
// Replaces, for one head of one batch row, the reference's  cat([self, latent, text, speaker]) + bool key mask +
// F.scaled_dot_product_attention  (model.py:246-261), the encoder self-attention (model.py:141-154, key mask or
// causal) and the DAC post_module window-limited causal attention (autoencoder.py:698-702, 762-773) -- without
// materialising the concatenated K/V or the 3x CFG copies of the text/speaker caches: every CFG branch points at
// the same cache and differs only in its mask / eff_len.
//
// v1 data path: cp.async double-buffered K/V tiles in XOR-swizzled smem, ldmatrix, mma.sync.m16n8k16 (bf16).
// 64 queries x 64 keys per step, 4 warps per CTA, grid = (ceil(S/64), H, b).
#include "attention.h"

#include <cuda_bf16.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#include "counters.h"
#include "launch.h"
#include "profiler.h"

namespace echo {

typedef __nv_bfloat16 bf16;

namespace {

constexpr int ATT_BM = 64;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 128;
constexpr int ATT_MAX_TILES = 96;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// element offset of (row, 16-byte chunk) inside a [rows][D] bf16 tile with XOR swizzle on the chunk index
template <int D>
__device__ __forceinline__ int swz(int row, int chunk) { return row * D + ((chunk ^ (row & 7)) << 3); }

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attn_kernel(const echo_attn_desc d) {
  constexpr int CH = D / 8;    // 16-byte chunks per row
  constexpr int KS = D / 16;   // k-steps over the head dim
  constexpr int ONT = D / 8;   // output n-tiles
  extern __shared__ __align__(128) uint8_t smem[];
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sK = sQ + ATT_BM * D;
  bf16* sV = sK + 2 * ATT_BN * D;
  uint8_t* sValid = reinterpret_cast<uint8_t*>(sV + 2 * ATT_BN * D);
  int* sTiles = reinterpret_cast<int*>(sValid + 2 * ATT_BN);
  __shared__ int s_ntiles;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;

  pdl_wait();
  pdl_trigger();
  // ---- tile list over all segments (few dozen entries)
  if (tid == 0) {
    int n = 0;
    for (int s = 0; s < d.nseg; ++s) {
      const echo_attn_segment& sg = d.seg[s];
      int hi = sg.len;
      if (sg.eff_len) { int e = sg.eff_len[b]; hi = e < hi ? e : hi; }
      int lo = 0;
      if (sg.pos_limit_mult > 0) {
        int lim = (sg.pos_limit + sg.pos_limit_mult - 1) / sg.pos_limit_mult;  // keys j with j*mult < pos_limit
        hi = lim < hi ? lim : hi;
      }
      if (sg.causal) {
        int qe = q0 + ATT_BM; if (qe > d.S) qe = d.S;
        hi = qe < hi ? qe : hi;
        if (sg.window > 0) { lo = q0 - sg.window + 1; if (lo < 0) lo = 0; lo &= ~(ATT_BN - 1); }
      }
      for (int n0 = lo; n0 < hi && n < ATT_MAX_TILES; n0 += ATT_BN) sTiles[n++] = (s << 24) | n0;
    }
    s_ntiles = n;
  }

  // ---- Q tile
  {
    const bf16* qb = static_cast<const bf16*>(d.Q) + (size_t)b * d.q_batch_stride + (size_t)h * D;
    for (int c = tid; c < ATT_BM * CH; c += ATT_THREADS) {
      const int row = c / CH, ch = c % CH;
      const bool ok = (q0 + row) < d.S;
      const bf16* src = qb + (size_t)(ok ? (q0 + row) : 0) * d.q_row_stride + ch * 8;
      cp_async16(sQ + swz<D>(row, ch), src, ok);
    }
    cp_async_commit();
  }
  __syncthreads();
  const int ntiles = s_ntiles;

  auto load_tile = [&](int ti, int stage) {
    const int e = sTiles[ti];
    const echo_attn_segment& sg = d.seg[e >> 24];
    const int n0 = e & 0xFFFFFF;
    int len = sg.len;
    if (sg.eff_len) { int el = sg.eff_len[b]; len = el < len ? el : len; }
    const int cb = sg.batch_mod > 0 ? b % sg.batch_mod : b;
    const bf16* kb = static_cast<const bf16*>(sg.K) + (size_t)cb * sg.batch_stride + (size_t)h * D;
    const bf16* vb = static_cast<const bf16*>(sg.V) + (size_t)cb * sg.batch_stride + (size_t)h * D;
    bf16* dk = sK + stage * ATT_BN * D;
    bf16* dv = sV + stage * ATT_BN * D;
    for (int c = tid; c < ATT_BN * CH; c += ATT_THREADS) {
      const int row = c / CH, ch = c % CH;
      const int j = n0 + row;
      const bool ok = j < len;
      const size_t off = (size_t)(ok ? j : 0) * sg.row_stride + ch * 8;
      cp_async16(dk + swz<D>(row, ch), kb + off, ok);
      cp_async16(dv + swz<D>(row, ch), vb + off, ok);
    }
    if (tid < ATT_BN) {
      const int j = n0 + tid;
      bool ok = j < len;
      if (ok && sg.mask) ok = sg.mask[(size_t)cb * sg.mask_ld + (size_t)j * sg.mask_stride] != 0;
      if (ok && sg.pos_limit_mult > 0) ok = (j * sg.pos_limit_mult) < sg.pos_limit;
      sValid[stage * ATT_BN + tid] = ok ? 1 : 0;
    }
  };

  if (ntiles > 0) load_tile(0, 0);
  cp_async_commit();

  // Q fragments (wait for the Q group only: at most the first K/V group may still be in flight)
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[KS][4];
  {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int ch = 2 * ks + (lane >> 4);
      ldsm_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], sQ + swz<D>(row, ch));
    }
  }

  float o[ONT][4];
#pragma unroll
  for (int i = 0; i < ONT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sl2 = d.scale * 1.4426950408889634f;
  const int qrow0 = q0 + warp * 16 + g;  // this thread's rows: qrow0, qrow0 + 8

  for (int ti = 0; ti < ntiles; ++ti) {
    const int stage = ti & 1;
    if (ti + 1 < ntiles) {
      load_tile(ti + 1, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const int e = sTiles[ti];
    const echo_attn_segment& sg = d.seg[e >> 24];
    const int n0 = e & 0xFFFFFF;
    const bf16* tk = sK + stage * ATT_BN * D;
    const bf16* tv = sV + stage * ATT_BN * D;
    const uint8_t* valid = sValid + stage * ATT_BN;

    // ---- S = Q K^T
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int ch = 2 * ks + ((lane >> 3) & 1);
        uint32_t r0, r1, r2, r3;
        ldsm_x4(r0, r1, r2, r3, tk + swz<D>(key, ch));
        mma_bf16(s[2 * np], qf[ks], r0, r1);
        mma_bf16(s[2 * np + 1], qf[ks], r2, r3);
      }
    }

    // ---- scale, mask, online softmax
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        const int col = nt * 8 + 2 * t + (e2 & 1);
        const int r = e2 >> 1;
        bool ok = valid[col] != 0;
        if (sg.causal) {
          const int j = n0 + col, q = qrow0 + r * 8;
          ok = ok && (j <= q) && (sg.window <= 0 || j > q - sg.window);
        }
        const float v = ok ? s[nt][e2] * sl2 : -INFINITY;
        s[nt][e2] = v;
        mx[r] = fmaxf(mx[r], v);
      }
    }
    float corr[2], muse[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(m_run[r], mx[r]);
      muse[r] = (mnew == -INFINITY) ? 0.f : mnew;
      corr[r] = exp2f(m_run[r] - muse[r]);  // m_run = -inf -> 0
      m_run[r] = mnew;
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        const int r = e2 >> 1;
        const float pv = exp2f(s[nt][e2] - muse[r]);
        s[nt][e2] = pv;
        l_run[r] += pv;
      }
    }
#pragma unroll
    for (int i = 0; i < ONT; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }

    // ---- O += P V
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      uint32_t pa[4];
      pa[0] = pack2(s[2 * k2][0], s[2 * k2][1]);
      pa[1] = pack2(s[2 * k2][2], s[2 * k2][3]);
      pa[2] = pack2(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
      pa[3] = pack2(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
      for (int dp = 0; dp < ONT / 2; ++dp) {
        const int key = k2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = 2 * dp + (lane >> 4);
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(r0, r1, r2, r3, tv + swz<D>(key, ch));
        mma_bf16(o[2 * dp], pa, r0, r1);
        mma_bf16(o[2 * dp + 1], pa, r2, r3);
      }
    }
    __syncthreads();  // all warps done with this stage before it is refilled
  }

  // ---- finalise: divide by the row sum, stage through smem, apply the output gate, coalesced store
  float inv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    inv[r] = l > 0.f ? 1.f / l : 0.f;
  }
  // sQ is private to this warp's 16 rows from here on (Q fragments live in registers)
#pragma unroll
  for (int nt = 0; nt < ONT; ++nt) {
    const int row = warp * 16 + g;
    // chunk nt holds columns [8nt, 8nt+8); this thread owns columns 2t, 2t+1 of it
    uint32_t* p0 = reinterpret_cast<uint32_t*>(sQ + swz<D>(row, nt)) + t;
    uint32_t* p1 = reinterpret_cast<uint32_t*>(sQ + swz<D>(row + 8, nt)) + t;
    *p0 = pack2(o[nt][0] * inv[0], o[nt][1] * inv[0]);
    *p1 = pack2(o[nt][2] * inv[1], o[nt][3] * inv[1]);
  }
  __syncwarp();
  {
    const size_t HD = (size_t)d.H * D;
    for (int c = lane; c < 16 * CH; c += 32) {
      const int rl = c / CH, ch = c % CH;
      const int row = warp * 16 + rl;
      const int q = q0 + row;
      if (q < d.S) {
        uint4 v = *reinterpret_cast<const uint4*>(sQ + swz<D>(row, ch));
        const size_t off = ((size_t)b * d.S + q) * HD + (size_t)h * D + ch * 8;
        if (d.gate) {
          const uint4 gv = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(d.gate) + off);
          const uint32_t* vi = reinterpret_cast<const uint32_t*>(&v);
          const uint32_t* gi = reinterpret_cast<const uint32_t*>(&gv);
          uint32_t r[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vi[i]));
            const float2 gg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gi[i]));
            r[i] = pack2(a.x * gg.x, a.y * gg.y);
          }
          v = make_uint4(r[0], r[1], r[2], r[3]);
        }
        *reinterpret_cast<uint4*>(static_cast<bf16*>(d.out) + off) = v;
      }
    }
  }
}

template <int D>
cudaError_t launch(const echo_attn_desc& d, cudaStream_t s) {
  static bool configured = false;
  const int smem = (ATT_BM * D + 4 * ATT_BN * D) * 2 + 2 * ATT_BN + ATT_MAX_TILES * 4;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((d.S + ATT_BM - 1) / ATT_BM, d.H, d.b);
  cudaError_t err;
  {
    char tag[64];
    snprintf(tag, sizeof(tag), "attn D=%d b=%d S=%d H=%d nseg=%d", D, d.b, d.S, d.H, d.nseg);
    ProfScope ps(PROF_ATTN, 0.0, 0.0, s, tag);
    err = launch_k(attn_kernel<D>, grid, dim3(ATT_THREADS), (size_t)smem, s, 1, d);
  }
  count_launch();
  return err;
}

}  // namespace

cudaError_t attention_launch(const echo_attn_desc& d, cudaStream_t s) {
  if (d.b <= 0 || d.S <= 0 || d.H <= 0 || d.nseg < 1 || d.nseg > 4) return cudaErrorInvalidValue;
  for (int i = 0; i < d.nseg; ++i)
    if (d.seg[i].len > (1 << 24) - 1) return cudaErrorInvalidValue;
  static const bool legacy = [] { const char* e = std::getenv("ECHO_ATTN_LEGACY"); return e && e[0] == '1'; }();
  if (!legacy && attention_tc_supported(d)) return attention_tc_launch(d, s);
  if (d.D == 128) return launch<128>(d, s);
  if (d.D == 64) return launch<64>(d, s);
  return cudaErrorInvalidValue;
}

}  // namespace echo
