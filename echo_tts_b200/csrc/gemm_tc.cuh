// tcgen05 GEMM for sm_100a:  C[M,N] = sum_taps A[rows + shift(tap), Kc] * B[N, tap*Kc + k]^T  (+ fused epilogues).
//
// One kernel serves every dense contraction on the Echo-TTS hot path:
//   * DiT / encoder linears (reference model.py:217-224, 263-266, 308): taps = 1, A = activations [M,K] bf16,
//     B = nn.Linear weight [N,K] bf16 (already K-major, no transpose needed).
//   * DAC causal convs (reference autoencoder.py:285-289, 310-316, 884-900): activations are kept time-major
//     (B, T, C), so a dilated causal conv is a GEMM whose K loop walks (tap, channel-block) and whose A tile for
//     tap j is the SAME matrix loaded at row offset -(k-1-j)*dilation. TMA zero-fills negative rows, which is
//     exactly the causal left padding; a 3-D tensor map (C, T, batch) keeps batches from bleeding.
//
// Structure (persistent, warp-specialised, 384 threads, 1 CTA / SM):
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one elected thread, tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue       (tcgen05.ld 32x32b -> registers -> fused math -> global)
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "common.cuh"
#include "gemm.h"
#include "launch.h"

namespace echo {

#ifndef ECHO_EPI_UNROLL
// 0: keep the per-warp chunk loops of the epilogues ROLLED. Unrolled, the generic kernel was 8 300 instructions
// (133 KB) of straight-line code that every warp executes once per tile, and ncu's source view showed its epilogue
// dominated by stall_no_inst (instruction-cache misses), profiles/r01_ncu_gemm_v6_stalls.txt.
#define ECHO_EPI_UNROLL 0
#endif
#ifndef ECHO_MAX_STAGES
#define ECHO_MAX_STAGES 8
#endif
#if ECHO_EPI_UNROLL
#define ECHO_CHUNK_UNROLL _Pragma("unroll")
#else
#define ECHO_CHUNK_UNROLL _Pragma("unroll 1")
#endif
constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARP0 = 4;
constexpr int GEMM_EPI_WARPS = 8;

__host__ __device__ constexpr int gemm_acc_stride(int BN) { return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512; }
// TMEM has 512 columns: two accumulator stages (epilogue of tile i overlaps the MMAs of tile i + 1) up to BN = 256, one
// for the 384-wide single-wave tiles
__host__ __device__ constexpr int gemm_acc_stages(int BN) { return 2 * gemm_acc_stride(BN) <= 512 ? 2 : 1; }
// per-CTA bytes of one pipeline stage; with a CTA pair (CG = 2) each CTA holds its 128 A rows and HALF of the B rows
__host__ __device__ constexpr int gemm_stage_bytes(int BN, int BK, int ATOMS, int CG) {
  return (GEMM_BM + BN / CG) * BK * 2 * ATOMS;
}
// every epilogue transposes 32 x 32 fp32 chunks through 4 KB of smem per epilogue warp; EPI_RU adds the bf16 A operand of
// its second MMA (128 x C) and 8 KB of per-channel tables (bias, alpha, 1 / alpha of both convs / Snakes)
__host__ __device__ constexpr int gemm_ru_a2_bytes(int BN) { return GEMM_BM * BN * 2; }
// EPI_RUW keeps that operand in the (then idle) transpose patches instead: 128 x 96 x 2 = 24 KB of the 32 KB
__host__ __device__ constexpr int gemm_epi_bytes(int EPI, int BN) {
  return GEMM_EPI_WARPS * 4096 + (EPI == EPI_RU ? gemm_ru_a2_bytes(BN) + 8192 : EPI == EPI_RUW ? 8192 : 0);
}
// EPI_RUW has no operand ring: [7 taps of conv7 weights | 1 x 1 weights | one A window of 128 + 64 rows], all K-major atoms
constexpr int RUW_TAPS = 7;
constexpr int RUW_WIN_ROWS = 192;
__host__ __device__ constexpr int gemm_ruw_bytes(int BN, int BK, int ATOMS) {
  return (RUW_TAPS + 1) * ATOMS * BN * BK * 2 + ATOMS * RUW_WIN_ROWS * BK * 2;
}
__host__ __device__ constexpr int gemm_stages(int BN, int BK, int ATOMS, int CG, int EPI) {
  if (EPI == EPI_RUW) return 2;  // two barrier pairs: [0] weights resident, [1] window full / free
  int s = (227 * 1024 - 1024 - 256 - gemm_epi_bytes(EPI, BN)) / gemm_stage_bytes(BN, BK, ATOMS, CG);
  return s > ECHO_MAX_STAGES ? ECHO_MAX_STAGES : s;
}
__host__ __device__ constexpr int gemm_ring_bytes(int BN, int BK, int ATOMS, int CG, int EPI) {
  return EPI == EPI_RUW ? gemm_ruw_bytes(BN, BK, ATOMS) : gemm_stages(BN, BK, ATOMS, CG, EPI) * gemm_stage_bytes(BN, BK, ATOMS, CG);
}
__host__ __device__ constexpr int gemm_smem_bytes(int BN, int BK, int ATOMS, int CG, int EPI) {
  return gemm_ring_bytes(BN, BK, ATOMS, CG, EPI) + gemm_epi_bytes(EPI, BN) + 1024 /*align slack*/ + 256 /*barriers*/;
}

// Rare activations (cond_module SiLU, ConvNeXt GELU, ...). Deliberately NOT inlined: inlining erff/tanhf 32x per
// chunk blew the generic kernel up to 1 MB of SASS and the epilogue became instruction-fetch bound.
__device__ __noinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    case ACT_SNAKE: {
      float s = __sinf(alpha * v);
      return v + s * s / (alpha + 1e-9f);
    }
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return sigmoid_f(v);
    case ACT_SILU: return silu_f(v);
    default: return v;
  }
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile --
// each CTA stages its own 128 A rows and half of the B rows, the even CTA issues M=256 MMAs that read both CTAs'
// shared memory and write both CTAs' TMEM, so the per-SM operand ingest drops from (128+BN) to (128+BN/2) rows per K.
template <int BN, int BK, int ATOMS, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmB1 /* EPI_RU: the 1 x 1 conv's weights; else a copy of tmB */, const GemmParams p) {
  constexpr int STAGES = gemm_stages(BN, BK, ATOMS, CG, EPI);
  constexpr int A_ATOM = GEMM_BM * BK * 2;
  constexpr int BNH = BN / CG;  // B rows staged by this CTA
  constexpr int B_ATOM = BNH * BK * 2;
  constexpr int STAGE_BYTES = (A_ATOM + B_ATOM) * ATOMS;
  constexpr int ACC_STRIDE = gemm_acc_stride(BN);
  constexpr int ROW_BYTES = BK * 2;
  static_assert(BK == 64 || BK == 32, "BK must match a swizzle span");
  static_assert(BN % 32 == 0 && BN >= 32 && (BN <= 256 || (BN == 384 && CG == 2 && EPI == EPI_QKV && ATOMS == 1)), "BN");
  // BN = 384 (CTA pairs, QKV epilogue): a 256 x 384 tile is two MMAs per K step (N = 256 into TMEM columns [0, 256) and
  // N = 128 into [256, 384)); each CTA stages 128 + 64 B rows. It exists to make the QKV GEMM of a 640-row plain step ONE
  // wave (3 pair rows x 22 column tiles = 66 <= 74 CTA pairs) instead of two (96 tiles of 256 x 256).
  constexpr int ACC_STAGES = gemm_acc_stages(BN);
  static_assert(A_ATOM % 1024 == 0 && B_ATOM % 1024 == 0, "atoms must keep 1024B alignment");
  constexpr bool RUW = EPI == EPI_RUW;       // fused ResidualUnit, weights resident + A window (see the producer branch)
  constexpr bool RU = EPI == EPI_RU || RUW;  // fused ResidualUnit: second MMA + two-phase epilogue
  constexpr bool CONV = EPI == EPI_CONV;     // lean conv epilogue: bias (+ stream) -> fp32, Snake -> bf16 (the RU's phase 2)
  constexpr int RING_BYTES = gemm_ring_bytes(BN, BK, ATOMS, CG, EPI);
  constexpr int RU_W = (BN / 2) % 32 == 0 ? 32 : 16;  // fused ResidualUnit: columns per epilogue piece, pieces per warp
  constexpr int RU_NPC = BN / 2 / RU_W;
  constexpr int RUW_W7_BYTES = RUW_TAPS * ATOMS * B_ATOM, RUW_W1_BYTES = ATOMS * B_ATOM;
  constexpr int WIN_ATOM = RUW_WIN_ROWS * BK * 2, RUW_WIN_BYTES = ATOMS * WIN_ATOM;
  static_assert(!RUW || (CG == 1 && gemm_ru_a2_bytes(BN) <= GEMM_EPI_WARPS * 4096 && BN == BK * ATOMS), "EPI_RUW shape");
  static_assert(gemm_smem_bytes(BN, BK, ATOMS, CG, EPI) <= 227 * 1024, "shared memory");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + RING_BYTES);  // [8 warps][32 x 32] (generic epilogue)
  // EPI_RU: [atoms][128 rows][BK] bf16, swizzled like A; EPI_RUW: the same, aliasing the transpose patches
  uint8_t* ru_a2 = smem + RING_BYTES + (RUW ? 0 : GEMM_EPI_WARPS * 4096);
  // EPI_RU / EPI_RUW: [6][BN] bias, alpha, 1 / alpha of conv7 + first Snake, then of the 1 x 1 conv + the next unit's Snake
  float* ru_tab = reinterpret_cast<float*>(smem + RING_BYTES + GEMM_EPI_WARPS * 4096 + (RUW ? 0 : gemm_ru_a2_bytes(BN)));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES + gemm_epi_bytes(EPI, BN));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* a2_ready_bar = bars + 2 * STAGES + 5;  // EPI_RU: the epilogue warps have written the second MMA's A operand
  uint64_t* t2full_bar = bars + 2 * STAGES + 7;    // EPI_RU: the second MMA's accumulator is ready
  static_assert((2 * ECHO_MAX_STAGES + 9) * 8 <= 256, "barrier block");
  // EPI_RU: K blocks of the 1 x 1 conv (its K = BN channels) and where they sit in the ring: after the first RU_J conv7
  // blocks of the NEXT tile, so that the second MMA of tile i runs in the middle of tile i + 1's first MMA (its operand
  // is written by the epilogue warps ~1.5 us after tile i's first accumulator is complete) and the second epilogue
  // overlaps the rest of it
  constexpr int RU_KB1 = (BN + BK * ATOMS - 1) / (BK * ATOMS);
  constexpr int RU_J = BN >= 192 ? 4 : 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* trace = p.trace ? p.trace + (size_t)blockIdx.x * 16 : nullptr;
  long long t_entry = 0;
  if (trace && threadIdx.x == 0) t_entry = clock64();

  uint32_t cta_rank = 0;
  if constexpr (CG == 2) cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;
  const int tiles_m = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int tiles_n = (BN == 384 && p.qkv_tiles > 0) ? p.qkv_tiles : (p.N + BN - 1) / BN;
  const int kb_per_tap = (p.Kc + BK * ATOMS - 1) / (BK * ATOMS);
  const int num_kb = kb_per_tap * p.taps;
  // split-K (EPI_GENERIC accumulate mode only): `splits` consecutive work units share one output tile and each
  // reduces a contiguous slice of the K blocks; partial sums meet in out_f32 through fp32 vector atomics
  const int splits = p.split_k > 1 ? p.split_k : 1;
  const int num_tiles = tiles_m * p.batches * tiles_n * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], GEMM_EPI_WARPS * CG);
      if constexpr (RU) {
        mbar_init(&a2_ready_bar[s], GEMM_EPI_WARPS);
        mbar_init(&t2full_bar[s], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_pair<ACC_STAGES * ACC_STRIDE>(tmem_slot);
    else tmem_alloc<ACC_STAGES * ACC_STRIDE>(tmem_slot);
  }
  if constexpr (RU) {
    // per-channel constants of the first Snake (weights: never written by a kernel of the chain, safe before the wait)
    if (warp >= GEMM_EPI_WARP0) {
      for (int c = threadIdx.x - GEMM_EPI_WARP0 * 32; c < BN; c += GEMM_EPI_WARPS * 32) {
        const float al = p.alpha ? p.alpha[c] : 1.f;
        ru_tab[c] = p.bias ? p.bias[c] : 0.f;
        ru_tab[BN + c] = al;
        ru_tab[2 * BN + c] = p.alpha_inv ? p.alpha_inv[c] : 1.f / (al + 1e-9f);
        const float ao = p.ru_alpha_out ? p.ru_alpha_out[c] : 1.f;
        ru_tab[3 * BN + c] = p.ru_bias1 ? p.ru_bias1[c] : 0.f;
        ru_tab[4 * BN + c] = ao;
        ru_tab[5 * BN + c] = p.ru_alpha_out_inv ? p.ru_alpha_out_inv[c] : 1.f / (ao + 1e-9f);
      }
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above touched only shared memory / TMEM / the kernel parameters: it overlaps the predecessor's tail.
  // So does the first pipeline fill of the B operand: B is always a weight matrix (written at load time, never by a
  // kernel of the chain), so its TMA loads may start before the predecessor has finished; only A waits.
  // Weights that are streamed once per launch (a DiT linear is 8-48 MB, the whole block 117 MB against 126 MB of L2) are
  // loaded with evict-first priority so they do not push the activations the next kernels re-read out of L2; the
  // activations (re-read by every column tile, written by the preceding kernel) with evict-last. Measured on whole
  // requests, same box: 192.0 -> 189.8 ms for the weights, another 0.3 % for the activations.
  const uint64_t b_policy = p.b_stream ? kL2EvictFirst : kL2EvictNormal;
  const uint64_t a_policy = p.b_stream ? kL2EvictLast : kL2EvictNormal;
  // The producer thread decodes its first work unit and issues the B (weight) tiles of the first pipeline fill BEFORE the
  // dependency wait; the matching A tiles follow right after it, back to back, from values already in registers.
  // (In-kernel timeline: with the decode -- nine integer divisions and a dozen parameter loads -- behind the wait, the
  // first A load left 0.47 us after the predecessor had finished, on every one of the ~7 000 launches of a request.)
  // B tile of one pipeline stage: rows [nrow, nrow + BNH) of this CTA -- except for BN = 384, where the CTA's 192 rows
  // are the 128 rows its half of the N = 256 MMA needs and the 64 rows of the N = 128 MMA (three 64-row boxes)
  auto load_b = [&](uint8_t* sb, uint64_t* bar, int kcoord, int bt_, int nt_) {
    const int ntile0 = bt_ * p.b_batch_rows + nt_ * BN;
    if constexpr (BN == 384) {
      // the tile's three 128-row groups come from the map (any three heads)
      const int boff = bt_ * p.b_batch_rows;
      const int r0 = boff + (int)p.tile_groups[3 * nt_ + (int)cta_rank] * 128;
      const int r1 = boff + (int)p.tile_groups[3 * nt_ + 2] * 128 + (int)cta_rank * 64;
      tma_load_2d_pair_hint(sb, &tmB, bar, kcoord, r0, b_policy);
      tma_load_2d_pair_hint(sb + 64 * ROW_BYTES, &tmB, bar, kcoord, r0 + 64, b_policy);
      tma_load_2d_pair_hint(sb + 128 * ROW_BYTES, &tmB, bar, kcoord, r1, b_policy);
    } else if constexpr (CG == 2) {
      tma_load_2d_pair_hint(sb, &tmB, bar, kcoord, ntile0 + (int)cta_rank * BNH, b_policy);
    } else {
      tma_load_2d_hint(sb, &tmB, bar, kcoord, ntile0, b_policy);
    }
  };
  int pre = 0;                           // k-blocks of the first unit whose B tiles are in flight before the wait
  int f_m0 = 0, f_ab = 0, f_kb_lo = 0;   // first unit: A row origin, A batch coordinate, first K block
  if constexpr (RUW) {
    // the whole conv7 + 1 x 1 weight set (7 x C x C + C x C bf16 = 144 KB at C = 96) becomes resident: one barrier
    if (warp == 0 && lane == 0 && tile0 < num_tiles) {
      mbar_expect_tx(&full_bar[0], RUW_W7_BYTES + RUW_W1_BYTES);
#pragma unroll 1
      for (int j = 0; j < RUW_TAPS; ++j)
#pragma unroll
        for (int a = 0; a < ATOMS; ++a)
          tma_load_2d_hint(smem + (j * ATOMS + a) * B_ATOM, &tmB, &full_bar[0], j * p.Kc + a * BK, 0, kL2EvictNormal);
#pragma unroll
      for (int a = 0; a < ATOMS; ++a)
        tma_load_2d_hint(smem + RUW_W7_BYTES + a * B_ATOM, &tmB1, &full_bar[0], a * BK, 0, kL2EvictNormal);
    }
  } else
  if (warp == 0 && lane == 0 && tile0 < num_tiles) {
    const int sk = tile0 % splits, tile = tile0 / splits;
    const int mt = tile % tiles_m;
    const int rest = tile / tiles_m;
    const int bt = rest % p.batches;
    const int nt = rest / p.batches;
    const int kb_lo = num_kb * sk / splits, kb_hi = num_kb * (sk + 1) / splits;
    f_m0 = (mt * CG + (int)cta_rank) * GEMM_BM;
    f_ab = bt / p.a_batch_div;
    f_kb_lo = kb_lo;
    if (!p.no_b_prefetch) {
      pre = (kb_hi - kb_lo) < STAGES ? (kb_hi - kb_lo) : STAGES;
      for (int i = 0; i < pre; ++i) {
        const int kb = kb_lo + i;
        const int tap = kb / kb_per_tap;
        const int kc0 = (kb - tap * kb_per_tap) * (BK * ATOMS);
        if (leader) mbar_expect_tx(&full_bar[i], STAGE_BYTES * CG);  // A's bytes are counted too; they are issued after the wait
        uint8_t* sb = smem + i * STAGE_BYTES + A_ATOM * ATOMS;
#pragma unroll
        for (int a = 0; a < ATOMS; ++a) load_b(sb + a * B_ATOM, &full_bar[i], tap * p.Kc + kc0 + a * BK, bt, nt);
      }
    }
  }
  pdl_wait();
  if (warp == 0 && lane == 0) {
    if (trace) trace[12] = clock64();
    for (int i = 0; i < pre; ++i) {
      const int kb = f_kb_lo + i;
      const int tap = p.taps == 1 ? 0 : kb / kb_per_tap;
      const int kc0 = (kb - tap * kb_per_tap) * (BK * ATOMS);
      uint8_t* sa = smem + i * STAGE_BYTES;
#pragma unroll
      for (int a = 0; a < ATOMS; ++a) {
        if constexpr (CG == 2) tma_load_3d_pair_hint(sa + a * A_ATOM, &tmA, &full_bar[i], kc0 + a * BK, f_m0 + p.tap_shift[tap], f_ab, a_policy);
        else tma_load_3d_hint(sa + a * A_ATOM, &tmA, &full_bar[i], kc0 + a * BK, f_m0 + p.tap_shift[tap], f_ab, a_policy);
      }
    }
    if (trace) trace[2] = clock64();
  }
  pdl_trigger();
  if (trace && threadIdx.x == 0) { trace[0] = t_entry; trace[1] = clock64(); }

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if constexpr (RUW) {
      // One window per tile: the A rows [m0 + shift(tap 0), + 128 + 6 d) that the seven taps read -- each tap's operand is
      // the same shared-memory tile addressed (tap * d) rows further down (row-shifted descriptors in the MMA thread),
      // instead of seven separate 128-row loads of overlapping rows. The window is single-buffered: it is reloaded as soon
      // as the tile's conv7 MMAs have retired, under the two epilogue phases of the tile.
      if (lane == 0) {
        int it = 0;
        for (int unit = tile0; unit < num_tiles; unit += tile_step, ++it) {
          const int mt = unit % tiles_m, bt = unit / tiles_m;
          if (it > 0) mbar_wait(&empty_bar[1], (uint32_t)(it - 1) & 1);
          mbar_expect_tx(&full_bar[1], RUW_WIN_BYTES);
#pragma unroll
          for (int a = 0; a < ATOMS; ++a)
            tma_load_3d_hint(smem + RUW_W7_BYTES + RUW_W1_BYTES + a * WIN_ATOM, &tmA, &full_bar[1], a * BK,
                             mt * GEMM_BM + p.tap_shift[0], bt / p.a_batch_div, a_policy);
        }
      }
    } else
    if (lane == 0) {
      int stage = pre == STAGES ? 0 : pre;  // the first `pre` slots are already being filled
      uint32_t phase = pre == STAGES ? 1 : 0;
      for (int unit = tile0; unit < num_tiles; unit += tile_step) {
        const int sk = unit % splits, tile = unit / splits;
        const int mt = tile % tiles_m;
        const int rest = tile / tiles_m;
        const int bt = rest % p.batches;
        const int nt = rest / p.batches;
        const int m0 = (mt * CG + (int)cta_rank) * GEMM_BM;
        const int kb_lo = num_kb * sk / splits, kb_hi = num_kb * (sk + 1) / splits;
        // EPI_RU: the 1 x 1 conv's weight blocks (B only) for the PREVIOUS tile's second MMA, RU_J blocks into this tile
        auto load_w1 = [&]() {
#pragma unroll 1
          for (int b1 = 0; b1 < RU_KB1; ++b1) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], B_ATOM * ATOMS);
            uint8_t* sb = smem + stage * STAGE_BYTES + A_ATOM * ATOMS;
#pragma unroll
            for (int a = 0; a < ATOMS; ++a)
              tma_load_2d_hint(sb + a * B_ATOM, &tmB1, &full_bar[stage], (b1 * ATOMS + a) * BK, 0, kL2EvictNormal);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        };
        const int ru_at = kb_lo + (kb_hi - kb_lo < RU_J ? kb_hi - kb_lo : RU_J);  // == kb_hi: after the tile's last block
        for (int kb = (unit == tile0 ? kb_lo + pre : kb_lo); kb < kb_hi; ++kb) {
          if constexpr (RU) { if (unit != tile0 && kb == ru_at) load_w1(); }
          const int tap = kb / kb_per_tap;
          const int kc0 = (kb - tap * kb_per_tap) * (BK * ATOMS);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], STAGE_BYTES * CG);  // counts both CTAs' bytes
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_ATOM * ATOMS;
#pragma unroll
          for (int a = 0; a < ATOMS; ++a) {
            if constexpr (CG == 2) tma_load_3d_pair_hint(sa + a * A_ATOM, &tmA, &full_bar[stage], kc0 + a * BK, m0 + p.tap_shift[tap], bt / p.a_batch_div, a_policy);
            else tma_load_3d_hint(sa + a * A_ATOM, &tmA, &full_bar[stage], kc0 + a * BK, m0 + p.tap_shift[tap], bt / p.a_batch_div, a_policy);
            load_b(sb + a * B_ATOM, &full_bar[stage], tap * p.Kc + kc0 + a * BK, bt, nt);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (RU) {
          if (unit != tile0 && ru_at == kb_hi) load_w1();                // a tile shorter than RU_J blocks
          if (unit + tile_step >= num_tiles) load_w1();                   // the CTA's last tile: its own second MMA
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if constexpr (RUW) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
        const uint32_t w7 = smem_u32(smem), w1 = w7 + RUW_W7_BYTES, win = w1 + RUW_W1_BYTES, a2 = smem_u32(ru_a2);
        // second MMA of tile j (the 1 x 1 conv): A = the bf16 tile the epilogue warps wrote, B = the resident weights
        auto issue_mma2 = [&](int j) {
          const int pas = j % ACC_STAGES;
          mbar_wait(&a2_ready_bar[pas], (j / ACC_STAGES) & 1);
          tc_fence_after();
          const uint32_t d2 = tmem_base + pas * ACC_STRIDE;
#pragma unroll
          for (int a = 0; a < ATOMS; ++a) {
            const uint64_t adesc = make_smem_desc<ROW_BYTES>(a2 + a * A_ATOM);
            const uint64_t bdesc = make_smem_desc<ROW_BYTES>(w1 + a * B_ATOM);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) tc_mma_f16(d2, adesc + 2 * k, bdesc + 2 * k, idesc, (a | k) != 0 ? 1u : 0u);
          }
          tc_commit(&t2full_bar[pas]);
        };
        if (tile0 < num_tiles) {
          mbar_wait(&full_bar[0], 0);  // weights resident
          tc_fence_after();
        }
        int it = 0;
        for (int unit = tile0; unit < num_tiles; unit += tile_step, ++it) {
          const int as = it % ACC_STAGES;
          mbar_wait(&tempty_bar[as], ((it / ACC_STAGES) & 1) ^ 1);
          mbar_wait(&full_bar[1], (uint32_t)it & 1);
          tc_fence_after();
          if (trace && unit == tile0) trace[3] = clock64();
          const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
#pragma unroll 1
          for (int j = 0; j < RUW_TAPS; ++j) {
            // tap j reads window rows [shift(j) - shift(0), + 128): the descriptor's start address moves down by whole
            // 64-byte rows. The swizzle is a function of the absolute shared-memory address bits, so the shifted start needs
            // nothing else -- setting the descriptor's base-offset field (bits 49-51) to the start's phase inside the 512-byte
            // pattern gives WRONG results (both forms were run against the seven separate loads on the hardware).
            const uint32_t roff = (uint32_t)(p.tap_shift[j] - p.tap_shift[0]) * ROW_BYTES;
#pragma unroll
            for (int a = 0; a < ATOMS; ++a) {
              const uint32_t sa = win + a * WIN_ATOM + roff;
              const uint64_t adesc = make_smem_desc<ROW_BYTES>(sa);
              const uint64_t bdesc = make_smem_desc<ROW_BYTES>(w7 + (j * ATOMS + a) * B_ATOM);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (j | a | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&empty_bar[1]);    // window reusable once these MMAs retire
          tc_commit(&tfull_bar[as]);   // conv7 accumulator ready for epilogue phase 1
          if (trace) trace[4] = clock64();
          // the tile's own 1 x 1 conv next, BEFORE the next tile's conv7: that one waits for its window (requested when
          // this tile's conv7 retired, an HBM round trip away), and queued behind it the second MMA left the epilogue
          // warps idle for 1.65 us of every 5.5 us tile
          issue_mma2(it);
        }
      }
    } else
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM * CG, BN == 384 ? 256 : BN);
      constexpr uint32_t idesc2 = make_idesc_bf16(GEMM_BM * CG, 128);  // BN = 384: the second MMA of a K step
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      // EPI_RU: second MMA of tile `j` (the 1 x 1 conv): A = the bf16 tile the epilogue warps wrote to shared memory,
      // B = the RU_KB1 weight blocks next in the ring, D = the tile's own accumulator stage (its first accumulator has been
      // read by then)
      auto issue_mma2 = [&](int j) {
        const int pas = j % ACC_STAGES;
        mbar_wait(&a2_ready_bar[pas], (j / ACC_STAGES) & 1);
        tc_fence_after();
        const uint32_t d2 = tmem_base + pas * ACC_STRIDE;
#pragma unroll 1
        for (int b1 = 0; b1 < RU_KB1; ++b1) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sb = smem_u32(smem + stage * STAGE_BYTES) + A_ATOM * ATOMS;
#pragma unroll
          for (int a = 0; a < ATOMS; ++a) {
            const uint64_t adesc = make_smem_desc<ROW_BYTES>(smem_u32(ru_a2) + (b1 * ATOMS + a) * A_ATOM);
            const uint64_t bdesc = make_smem_desc<ROW_BYTES>(sb + a * B_ATOM);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) tc_mma_f16(d2, adesc + 2 * k, bdesc + 2 * k, idesc, (b1 | a | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&t2full_bar[pas]);
      };
      for (int unit = tile0; unit < num_tiles; unit += tile_step, ++it) {
        const int as = it % ACC_STAGES;
        mbar_wait(&tempty_bar[as], ((it / ACC_STAGES) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        const int sk = unit % splits;
        const int kb_lo = num_kb * sk / splits, kb_hi = num_kb * (sk + 1) / splits;
        const int ru_at = kb_lo + (kb_hi - kb_lo < RU_J ? kb_hi - kb_lo : RU_J);
        bool third = true;  // BN = 384: does this tile hold a third group (else only the N = 256 MMA runs)
        if constexpr (BN == 384) third = (int)p.tile_groups[3 * ((unit / splits) / tiles_m / p.batches) + 2] * 128 < p.N;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          if constexpr (RU) { if (it > 0 && kb == ru_at) issue_mma2(it - 1); }
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (trace && kb == kb_lo && unit == tile0) trace[3] = clock64();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_ATOM * ATOMS;
#pragma unroll
          for (int a = 0; a < ATOMS; ++a) {
            const uint64_t adesc = make_smem_desc<ROW_BYTES>(sa + a * A_ATOM);
            const uint64_t bdesc = make_smem_desc<ROW_BYTES>(sb + a * B_ATOM);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 (32 bytes) along K inside the swizzle span: +2 in the 16-byte address field
              if constexpr (CG == 2) tc_mma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb_lo) | a | k) != 0 ? 1u : 0u);
              else tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb_lo) | a | k) != 0 ? 1u : 0u);
              if (BN == 384 && third) {  // columns [256, 384): this CTA's 64 B rows sit behind its 128 rows of the first MMA
                const uint64_t bdesc2 = make_smem_desc<ROW_BYTES>(sb + a * B_ATOM + 128 * ROW_BYTES);
                tc_mma_f16_pair(d_tmem + 256, adesc + 2 * k, bdesc2 + 2 * k, idesc2, ((kb - kb_lo) | a | k) != 0 ? 1u : 0u);
              }
            }
          }
          // smem slot reusable (in both CTAs of a pair) once these MMAs retire
          if constexpr (CG == 2) tc_commit_pair(&empty_bar[stage]);
          else tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator ready for the epilogue warps (of both CTAs)
        if constexpr (CG == 2) tc_commit_pair(&tfull_bar[as]);
        else tc_commit(&tfull_bar[as]);
        if (trace) trace[4] = clock64();
        if constexpr (RU) {
          if (it > 0 && ru_at == kb_hi) issue_mma2(it - 1);
          if (unit + tile_step >= num_tiles) issue_mma2(it);  // the CTA's last tile
        }
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - GEMM_EPI_WARP0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const int half = ew >> 2;      // which half of the column chunks
    int it = 0;
    for (int unit = tile0; unit < num_tiles; unit += tile_step, ++it) {
      // fused ResidualUnit: one column tile, no split-K -- one division instead of four
      const int sk = RU ? 0 : unit % splits, tile = RU ? unit : unit / splits;
      const int rest = tile / tiles_m;
      const int mt = tile - rest * tiles_m;
      const int bt = RU ? rest : rest % p.batches;
      const int nt = RU ? 0 : rest / p.batches;
      const int mbase = (mt * CG + (int)cta_rank) * GEMM_BM + quarter * 32;  // first row of this warp's 32-row slab
      const int n0 = nt * BN;
      const int as = it % ACC_STAGES;
      const uint32_t acc_phase = (it / ACC_STAGES) & 1;
      const uint32_t tbase = tmem_base + as * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
      if constexpr (EPI != EPI_GENERIC && EPI != EPI_ACCUM && !RU && !CONV) {
        mbar_wait(&tfull_bar[as], acc_phase);
        tc_fence_after();
        if (trace && warp == GEMM_EPI_WARP0 && lane == 0) {
          trace[5] = clock64();
          if (it < 4) trace[8 + 2 * it] = trace[5];
        }
      }

      if constexpr (RU || CONV) {
        // ---- phase 1 of the fused ResidualUnit: conv7 accumulator -> + bias -> Snake -> bf16 -> the A operand of the
        // second MMA, written straight into shared memory in the K-major swizzled layout a TMA load of the same tile would
        // have produced (atoms of [128 rows][BK]; 16-byte chunk index XOR row bits: row & 7 for 128-byte rows, (row >> 1) & 3
        // for 64-byte rows). The single A2 buffer is free: this warp only gets here after its second epilogue of the
        // previous tile, i.e. after that tile's second MMA has completed.
        // Set-up of phase 2 first: its pointers, the L2 prefetch of the NEXT tile's rows of the stream and the loads of this
        // tile's first piece of the stream -- they are in flight during phase 1 instead of being waited for at its end.
        constexpr int CPR = RU_W / 4, RPI = 32 / CPR, NIT = 32 / RPI;  // lanes per row, rows per pass, passes
        const uint32_t stg = smem_u32(epi_stage + ew * 1024);
        const uint32_t tab = smem_u32(ru_tab);
        const int sub = lane / CPR, c4 = lane % CPR;
        const size_t row0 = (size_t)bt * p.M + mbase;
        const int rows_left = p.M - mbase;
        const bool all_rows = rows_left >= 32;  // warp-uniform
        const int colw = half * (BN / 2) + 4 * c4;  // this lane's first column of piece 0 (inside the tile)
        // EPI_CONV: the stream input and / or the fp32 output may be absent (warp-uniform)
        const bool has_resid = RU || p.resid != nullptr, has_f32 = RU || p.out_f32 != nullptr;
        const float* const rp = p.resid + (row0 + sub) * p.ld_f32 + n0 + colw;
        float* const xp = p.out_f32 + (row0 + sub) * p.ld_f32 + n0 + colw;
        bf16* const np = p.out_bf16 + (row0 + sub) * p.ld_bf16 + n0 + colw;
        const size_t step32 = (size_t)RPI * p.ld_f32, step16 = (size_t)RPI * p.ld_bf16;
        // rows sub + RPI i of the patch. 128-byte rows (RU_W = 32): chunk index XOR (row & 7) = sub | sub + 4;
        // 64-byte rows (RU_W = 16): XOR ((row >> 1) & 3), the same for every pass
        const uint32_t rd_even = RU_W == 32 ? stg + sub * 128 + ((c4 ^ sub) << 4) : stg + sub * 64 + ((c4 ^ ((sub >> 1) & 3)) << 4);
        const uint32_t rd_odd = stg + (sub + 4) * 128 + ((c4 ^ (sub + 4)) << 4);  // RU_W = 32 only
        float4 rcur[NIT], rnext[NIT];
        auto load_resid = [&](int pc, float4* rr) {
          if (pc < RU_NPC && has_resid) {
            const float* q = rp + pc * RU_W;
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
              if (all_rows || sub + RPI * i < rows_left) rr[i] = *reinterpret_cast<const float4*>(q);
              q += step32;
            }
          }
        };
        // the rows of this CTA's NEXT tile go to L2 now, a whole tile period ahead of their use
        {
          const int u2 = unit + tile_step;
          int mt2 = mt + tile_step, bt2 = bt;
          if (mt2 >= tiles_m) {  // next batch item(s): rare, so the division is off the common path
            const int q = mt2 / tiles_m;
            mt2 -= q * tiles_m;
            bt2 += q;
          }
          const int rown = mt2 * GEMM_BM + quarter * 32 + lane;
          if (RU && u2 < num_tiles && p.ld_f32 == BN && (mt2 + 1) * GEMM_BM <= p.M) {
            // contiguous rows: ONE bulk prefetch of the warp's 32 x BN slab (a per-lane prefetch.global.L2 is a wavefront of
            // the L1 data pipe per lane and line -- 768 per tile, on the pipe that bounds this epilogue)
            if (half == 0 && lane == 0) {
              const float* q = p.resid + ((size_t)bt2 * p.M + mt2 * GEMM_BM + quarter * 32) * BN;
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "n"(32 * BN * 4) : "memory");
            }
          } else if (RU && u2 < num_tiles && rown < p.M) {
            const char* q = reinterpret_cast<const char*>(p.resid + ((size_t)bt2 * p.M + rown) * p.ld_f32 + half * (BN / 2));
#pragma unroll
            for (int o = 0; o < BN * 2; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q + BN * 2 - 4));
          }
        }
        load_resid(0, rcur);
        const bool tr = trace && warp == GEMM_EPI_WARP0 && lane == 0 && it == 2;  // steady-state tile of the timeline
        if constexpr (RU) {
        // EPI_RUW: the operand lives in the transpose patches -- every warp must be through the previous tile's phase 2
        if constexpr (RUW) asm volatile("bar.sync 1, %0;" ::"n"(GEMM_EPI_WARPS * 32) : "memory");
        if (tr) trace[8] = clock64();
        mbar_wait(&tfull_bar[as], acc_phase);
        tc_fence_after();
        if (tr) trace[9] = clock64();
        const int r = quarter * 32 + lane;  // row of the tile == TMEM lane
        const int sw = ROW_BYTES == 128 ? (r & 7) : ROW_BYTES == 64 ? ((r >> 1) & 3) : ((r >> 2) & 1);
        const uint32_t a2row = smem_u32(ru_a2) + r * ROW_BYTES;
        // the two warps of a lane quarter split the columns in halves (BN / 2 each, RU_W columns at a time); per-channel
        // constants come from the shared-memory table four at a time, the arithmetic runs on fp32 pairs
#pragma unroll 1
        for (int pc = 0; pc < RU_NPC; ++pc) {
          const int c0 = half * (BN / 2) + pc * RU_W;
          float v[RU_W];
          if constexpr (RU_W == 32) tc_ld_32x32(tbase + c0, v);
          else tc_ld_32x16(tbase + c0, v);
          tc_wait_ld();
#pragma unroll
          for (int g = 0; g < RU_W / 8; ++g) {  // 16-byte chunks of the operand row = 8 channels each
            const int k0 = c0 + 8 * g;          // first channel of the chunk
            uint32_t w[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t ta = tab + (k0 + 4 * h) * 4;
              const float4 b = lds_f4(ta), al = lds_f4(ta + BN * 4), iv = lds_f4(ta + 2 * BN * 4);
              const float* vv = v + 8 * g + 4 * h;
              const float2 y0 = snake2(f2add(make_float2(vv[0], vv[1]), make_float2(b.x, b.y)), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
              const float2 y1 = snake2(f2add(make_float2(vv[2], vv[3]), make_float2(b.z, b.w)), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
              w[2 * h] = pack_bf16(y0.x, y0.y);
              w[2 * h + 1] = pack_bf16(y1.x, y1.y);
            }
            const int atom = k0 / BK, c16 = (k0 % BK) / 8;
            sts_u4(a2row + atom * A_ATOM + ((c16 ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
          }
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_ready_bar[as]);
        if (tr) trace[10] = clock64();
        }  // RU: phase 1
        // ---- phase 2: second accumulator + bias + x -> the fp32 stream, Snake of the next unit -> bf16. Same arithmetic, in
        // the same order, as the generic epilogue (the two-launch form is bit-identical), but nothing else: the generic
        // code re-tests its options per row and was ~1 000 instructions per 32 x 32 chunk and warp -- with two epilogue warps
        // per scheduler (one instruction per ~4.5 cycles) phase 2 alone took 5.5 us of an 8.2 us tile
        // (profiles/r02_dac_ru_timeline.txt). Chunks of RU_W columns go through the warp's transpose patch as before.
        mbar_wait(RU ? &t2full_bar[as] : &tfull_bar[as], acc_phase);
        tc_fence_after();
        if (trace && warp == GEMM_EPI_WARP0 && lane == 0) {
          trace[5] = clock64();
          if (it == 2) trace[11] = trace[5];
        }
        const int cmod = p.col_mod > 0 ? p.col_mod : p.N;  // EPI_CONV: multiple of 32, so a piece never wraps
#pragma unroll 1
        for (int pc = 0; pc < RU_NPC; ++pc) {
          float v[RU_W];
          if constexpr (RU_W == 32) tc_ld_32x32(tbase + half * (BN / 2) + pc * RU_W, v);
          else tc_ld_32x16(tbase + half * (BN / 2) + pc * RU_W, v);
          load_resid(pc + 1, rnext);
          float4 b4, a4, i4;
          if constexpr (RU) {
            const uint32_t ta = tab + (3 * BN + colw + pc * RU_W) * 4;
            b4 = lds_f4(ta), a4 = lds_f4(ta + BN * 4), i4 = lds_f4(ta + 2 * BN * 4);
          } else {
            const int cb = (n0 + half * (BN / 2) + pc * RU_W) % cmod + 4 * c4;  // one modulo per piece
            b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + cb)) : make_float4(0.f, 0.f, 0.f, 0.f);
            a4 = __ldg(reinterpret_cast<const float4*>(p.alpha + cb));
            if (p.alpha_inv) i4 = __ldg(reinterpret_cast<const float4*>(p.alpha_inv + cb));
            else i4 = make_float4(1.f / (a4.x + 1e-9f), 1.f / (a4.y + 1e-9f), 1.f / (a4.z + 1e-9f), 1.f / (a4.w + 1e-9f));
          }
          tc_wait_ld();
          if constexpr (RU_W == 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts_v4(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          __syncwarp();
          float* xo = xp + pc * RU_W;
          bf16* no = np + pc * RU_W;
          float4 tt[NIT];  // all rows of the piece first: the loads then overlap each other and the math of earlier rows
#pragma unroll
          for (int i = 0; i < NIT; ++i)
            tt[i] = lds_v4(RU_W == 32 ? ((i & 1) ? rd_odd : rd_even) + (i >> 1) * 1024 : rd_even + i * 512);
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            const float4 t = tt[i];
            if (all_rows || sub + RPI * i < rows_left) {
              float2 t0 = f2add(make_float2(t.x, t.y), make_float2(b4.x, b4.y));
              float2 t1 = f2add(make_float2(t.z, t.w), make_float2(b4.z, b4.w));
              if (has_resid) {
                t0 = f2add(t0, make_float2(rcur[i].x, rcur[i].y));
                t1 = f2add(t1, make_float2(rcur[i].z, rcur[i].w));
              }
              if (has_f32) *reinterpret_cast<float4*>(xo) = make_float4(t0.x, t0.y, t1.x, t1.y);
              t0 = snake2(t0, make_float2(a4.x, a4.y), make_float2(i4.x, i4.y));
              t1 = snake2(t1, make_float2(a4.z, a4.w), make_float2(i4.z, i4.w));
              *reinterpret_cast<uint2*>(no) = make_uint2(pack_bf16(t0.x, t0.y), pack_bf16(t1.x, t1.y));
            }
            xo += step32;
            no += step16;
          }
          __syncwarp();  // the patch is rewritten by the next chunk
#pragma unroll
          for (int i = 0; i < NIT; ++i) rcur[i] = rnext[i];
        }
      } else if constexpr (EPI == EPI_GENERIC) {
        const float* const e_bias = p.bias;
        const float* const e_alpha = p.alpha;
        const float* const e_alpha_inv = p.alpha_inv;
        uint64_t* const acc_bar = &tfull_bar[as];
        // tcgen05.ld hands every thread one ROW of the chunk (32 consecutive columns). Touching global memory in
        // that shape costs 32 L1 wavefronts per 128-bit access (32 lines, 16 B each) and made this epilogue -- an
        // fp32 read-modify-write of the residual stream -- longer than the K = 2048 mainloop (measured 14.8 us vs
        // 10.4 us per 128 x 256 tile, profiles/r01_gemm_inkernel_timeline_v3.txt). Each warp therefore transposes
        // its 32 x 32 chunk through a private XOR-swizzled 4 KB smem patch: afterwards a lane owns 4 consecutive
        // columns of 8 rows, every 128-bit access covers 4 full 128 B lines, and the per-column operands
        // (bias, gate, snake alpha) are one float4 per lane instead of 8 broadcast loads.
        // Software-pipelined: the residual of chunk c+1 is requested while chunk c is processed, and the first
        // chunk's residual before the accumulator barrier.
        constexpr int NCH = (BN / 32 + 1) / 2;  // chunks per warp (this warp takes ch = half, half + 2, ...)
        const uint32_t stg = smem_u32(epi_stage + ew * 1024);  // shared-space byte address of this warp's patch
        const int sub = lane >> 3;  // row inside a group of 4
        const int c4 = lane & 7;    // float4 column of the chunk owned after the transpose
        const int cmod = p.col_mod > 0 ? p.col_mod : p.N;  // multiple of 32, so a 32-column chunk never wraps
        const size_t row0 = (size_t)bt * p.M + mbase;      // global row of this warp's slab
        const int rows_left = p.M - mbase;                 // rows of the slab inside the matrix (may be <= 0)
        const bool all_rows = rows_left >= 32;             // warp-uniform
        // split-K: out_f32 already holds the residual; every split adds its gated partial sum atomically
        const bool atomic_out = splits > 1 || p.atomic_out != 0;
        const float* resid = atomic_out ? nullptr : p.resid;
        // Everything that does not depend on the row is hoisted out of the row loop, which then is: one LDS, one FMA per
        // element, residual add, stores, pointer steps. (The first version recomputed 64-bit row addresses and
        // re-tested every option per row: ~2 800 instructions per warp and tile, and with two epilogue warps per
        // scheduler the epilogue of a DAC conv tile -- K = 7 x 96 .. 7 x 192 -- took twice as long as its mainloop.)
        // gemm_launch guarantees rows_per_gate % 32 == 0: a warp's 32-row slab never straddles two gate rows.
        const float* gate = p.gate;
        if (gate != nullptr && p.rows_per_gate > 0 && rows_left > 0)
          gate += (size_t)((uint32_t)row0 / (uint32_t)p.rows_per_gate) * (size_t)p.gate_ld;
        const float* bias = (e_bias != nullptr && sk == 0) ? e_bias + (size_t)bt * p.bias_bstride : nullptr;
        const int nstore = p.n_valid > 0 ? p.n_valid : p.N;
        const uint32_t rd_even = stg + sub * 128 + ((c4 ^ sub) << 4);  // rows sub + 4i: (sub + 4i) & 7 = sub | sub + 4
        const uint32_t rd_odd = stg + (sub + 4) * 128 + ((c4 ^ (sub + 4)) << 4);
        const size_t step32 = (size_t)4 * p.ld_f32, step16 = (size_t)4 * p.ld_bf16;
        const bool snake = p.act == ACT_SNAKE;  // the hot activation (every DAC conv); the others go through apply_act
        float4 rcur[8], rnext[8];
        auto load_resid = [&](int ch, float4* r) {
          const int c0 = n0 + ch * 32;
          if (resid != nullptr && ch < BN / 32 && c0 < p.N) {
            const float* rp = resid + (row0 + sub) * p.ld_f32 + c0 + 4 * c4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (all_rows || sub + 4 * i < rows_left) r[i] = *reinterpret_cast<const float4*>(rp);
              rp += step32;
            }
          }
        };
        // Pull this tile's residual rows into L2 while the MMAs still run: the read-modify-write of the fp32
        // residual stream is a chip-wide burst (every CTA reaches its epilogue at the same time) and with only one
        // chunk of loads in flight per thread it was latency-bound at ~3 TB/s (10.7 us per 128 x 256 tile).
        if (resid != nullptr && lane < rows_left) {
          const char* rp = reinterpret_cast<const char*>(resid + (row0 + lane) * p.ld_f32 + n0);
          const int bytes = (p.N - n0 < BN ? p.N - n0 : BN) * 4;
          for (int o = half * 128; o < bytes; o += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + o));
        }
        load_resid(half, rcur);
        mbar_wait(acc_bar, acc_phase);
        tc_fence_after();
        if (trace && warp == GEMM_EPI_WARP0 && lane == 0) {
          trace[5] = clock64();
          if (it < 4) trace[8 + 2 * it] = trace[5];
        }
ECHO_CHUNK_UNROLL
        for (int ci = 0; ci < NCH; ++ci) {
          const int ch = half + 2 * ci;
          if (ch < BN / 32) {
            float v[32];
            tc_ld_32x32(tbase + ch * 32, v);
            load_resid(ch + 2, rnext);
            const int c0 = n0 + ch * 32;
            const bool col_ok = c0 < nstore;                   // warp-uniform
            const bool lane_ok = c0 + 4 * c4 < nstore;         // n_valid may cut a chunk
            const int cb = (col_ok ? c0 % cmod : 0) + 4 * c4;  // one modulo per chunk
            // t = (acc + bias) * scale * gate  ==  acc * gs + bs
            float4 gs = make_float4(p.scale, p.scale, p.scale, p.scale), bs = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 a4 = make_float4(1.f, 1.f, 1.f, 1.f), i4 = a4;
            if (col_ok) {
              if (gate != nullptr) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gate + cb));
                gs.x *= g.x; gs.y *= g.y; gs.z *= g.z; gs.w *= g.w;
              }
              if (bias != nullptr) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + cb));
                bs = make_float4(b.x * gs.x, b.y * gs.y, b.z * gs.z, b.w * gs.w);
              }
              if (snake && p.out_bf16) {
                a4 = __ldg(reinterpret_cast<const float4*>(e_alpha + cb));
                if (e_alpha_inv) i4 = __ldg(reinterpret_cast<const float4*>(e_alpha_inv + cb));
                else i4 = make_float4(1.f / (a4.x + 1e-9f), 1.f / (a4.y + 1e-9f), 1.f / (a4.z + 1e-9f), 1.f / (a4.w + 1e-9f));
              }
            }
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            if (col_ok) {
              float* o32 = p.out_f32 ? p.out_f32 + (row0 + sub) * p.ld_f32 + c0 + 4 * c4 : nullptr;
              bf16* o16 = p.out_bf16 ? p.out_bf16 + (row0 + sub) * p.ld_bf16 + c0 + 4 * c4 : nullptr;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 t = lds_v4(((i & 1) ? rd_odd : rd_even) + (i >> 1) * 1024);
                if (lane_ok && (all_rows || sub + 4 * i < rows_left)) {
                  t.x = fmaf(t.x, gs.x, bs.x); t.y = fmaf(t.y, gs.y, bs.y);
                  t.z = fmaf(t.z, gs.z, bs.z); t.w = fmaf(t.w, gs.w, bs.w);
                  if (resid) { t.x += rcur[i].x; t.y += rcur[i].y; t.z += rcur[i].z; t.w += rcur[i].w; }
                  if (o32) {
                    if (atomic_out) atomicAdd(reinterpret_cast<float4*>(o32), t);  // RED.ADD.F32x4
                    else *reinterpret_cast<float4*>(o32) = t;
                  }
                  if (o16) {
                    if (snake) {
                      // snake(x) = x + sin^2(alpha x) / (alpha + 1e-9)   (autoencoder.py:96-102); the result is
                      // rounded to bf16, so the SFU sine (abs err ~|x| 2^-22) is far below the output quantum
                      const float s0 = __sinf(a4.x * t.x), s1 = __sinf(a4.y * t.y);
                      const float s2 = __sinf(a4.z * t.z), s3 = __sinf(a4.w * t.w);
                      t.x = fmaf(s0 * s0, i4.x, t.x); t.y = fmaf(s1 * s1, i4.y, t.y);
                      t.z = fmaf(s2 * s2, i4.z, t.z); t.w = fmaf(s3 * s3, i4.w, t.w);
                    } else if (p.act != ACT_NONE) {
                      t.x = apply_act(t.x, p.act, 1.f); t.y = apply_act(t.y, p.act, 1.f);
                      t.z = apply_act(t.z, p.act, 1.f); t.w = apply_act(t.w, p.act, 1.f);
                    }
                    *reinterpret_cast<uint2*>(o16) = make_uint2(pack_bf16(t.x, t.y), pack_bf16(t.z, t.w));
                  }
                }
                if (o32) o32 += step32;
                if (o16) o16 += step16;
              }
            }
            __syncwarp();  // the patch is rewritten by the next chunk
#pragma unroll
            for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
          }
        }
      } else if constexpr (EPI == EPI_SWIGLU) {
        // h = silu(a) * b in the row-per-thread layout, then the same smem transpose as the generic epilogue so
        // the bf16 rows leave as 64-byte runs (a lane owns 4 consecutive columns of 8 rows) instead of 32 scattered
        // 16-byte pieces per store instruction.
        constexpr int HALF_CH = BN / 64;  // chunks in the w1 half
        const uint32_t stg = smem_u32(epi_stage + ew * 1024);
        const int sub = lane >> 3, c4 = lane & 7;
        const size_t row0 = (size_t)bt * p.M + mbase;
        const int rows_left = p.M - mbase;
ECHO_CHUNK_UNROLL
        for (int ch = half; ch < HALF_CH; ch += 2) {
          float a[32], b[32];
          tc_ld_32x32(tbase + ch * 32, a);
          tc_ld_32x32(tbase + (ch + HALF_CH) * 32, b);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), silu_fast(a[4 * j]) * b[4 * j], silu_fast(a[4 * j + 1]) * b[4 * j + 1],
                   silu_fast(a[4 * j + 2]) * b[4 * j + 2], silu_fast(a[4 * j + 3]) * b[4 * j + 3]);
          __syncwarp();
          bf16* op = p.out_bf16 + row0 * p.ld_bf16 + n0 / 2 + ch * 32 + 4 * c4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = sub + 4 * i;
            const float4 t = lds_v4(stg + rr * 128 + ((c4 ^ (rr & 7)) << 4));
            if (rr < rows_left)
              *reinterpret_cast<uint2*>(op + (size_t)rr * p.ld_bf16) = make_uint2(pack_bf16(t.x, t.y), pack_bf16(t.z, t.w));
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_ACCUM) {
        // out_f32[r, c] += gate[c] * scale * (acc[r, c] + bias[c]) with fire-and-forget fp32 vector reductions: the
        // residual-stream update of wo / w2 (model.py:388-389) and of every encoder / DAC transformer block. The
        // generic epilogue executed ~2 800 instructions per warp and tile for this (ncu: epilogue 58 % of the wo
        // kernel at M = 1920, stalls spread thin over `wait` / `selected` -- plain instruction count with only two
        // epilogue warps per scheduler); this instantiation keeps only what the accumulate needs (~350).
        constexpr int NCH = (BN / 32 + 1) / 2;
        const uint32_t stg = smem_u32(epi_stage + ew * 1024);
        const int sub = lane >> 3, c4 = lane & 7;
        const size_t row0 = (size_t)bt * p.M + mbase;
        const int rows_left = p.M - mbase;
        const float* gate = p.gate;
        // gemm_launch guarantees rows_per_gate % 32 == 0, so a warp's 32-row slab never straddles two gate rows
        if (gate != nullptr && p.rows_per_gate > 0 && rows_left > 0) gate += (row0 / p.rows_per_gate) * (size_t)p.gate_ld;
        const bool add_bias = p.bias != nullptr && sk == 0;
        // rows sub + 4i of the transposed patch: (sub + 4i) & 7 is sub for even i and sub + 4 for odd i
        const uint32_t rd_even = stg + sub * 128 + ((c4 ^ sub) << 4);
        const uint32_t rd_odd = stg + (sub + 4) * 128 + ((c4 ^ (sub + 4)) << 4);
        // gate / bias of all this warp's chunks are fetched while the MMAs still run (they come from the AdaLN tables /
        // the weights, never from the preceding kernel): behind the accumulator barrier there is no L2 round trip left
        float4 gq[NCH], bq[NCH];
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c0 = n0 + (half + 2 * ci) * 32;
          gq[ci] = make_float4(1.f, 1.f, 1.f, 1.f);
          bq[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (half + 2 * ci < BN / 32 && c0 < p.N) {
            if (gate != nullptr) gq[ci] = __ldg(reinterpret_cast<const float4*>(gate + c0 + 4 * c4));
            if (add_bias) bq[ci] = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4 * c4));
          }
        }
        mbar_wait(&tfull_bar[as], acc_phase);
        tc_fence_after();
        if (trace && warp == GEMM_EPI_WARP0 && lane == 0) {
          trace[5] = clock64();
          if (it < 4) trace[8 + 2 * it] = trace[5];
        }
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int ch = half + 2 * ci;
          const int c0 = n0 + ch * 32;
          if (ch < BN / 32 && c0 < p.N) {
            float v[32];
            tc_ld_32x32(tbase + ch * 32, v);
            const float4 g4 = make_float4(gq[ci].x * p.scale, gq[ci].y * p.scale, gq[ci].z * p.scale, gq[ci].w * p.scale);
            const float4 b4 = make_float4(bq[ci].x * g4.x, bq[ci].y * g4.y, bq[ci].z * g4.z, bq[ci].w * g4.w);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const size_t step = (size_t)4 * p.ld_f32;
            const bool all_rows = rows_left >= 32;  // warp-uniform
            if (p.part_out != nullptr) {
              // split-K without atomics: slice sk parks its gated partial sum in its own plane of a workspace and the
              // LowRankAdaLN / RMSNorm kernel that follows adds the planes to the residual stream (glue.cu). Fixed
              // summation order: bit-reproducible. Not faster than the reductions: with all SMs pushing 128 x 256 tiles,
              // red.global.add.v4.f32 drains 8 % slower than plain stores, three CTAs per tile or one
              // (tools/red_drain.cu, profiles/r02_tma_reduce_experiment.txt), and the planes cost the next kernel reads.
              float* dst = p.part_out + (size_t)sk * p.part_stride + (row0 + sub) * (size_t)p.ld_f32 + c0 + 4 * c4;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 t = lds_v4(((i & 1) ? rd_odd : rd_even) + (i >> 1) * 1024);
                t.x = fmaf(t.x, g4.x, b4.x); t.y = fmaf(t.y, g4.y, b4.y);
                t.z = fmaf(t.z, g4.z, b4.z); t.w = fmaf(t.w, g4.w, b4.w);
                if (all_rows || sub + 4 * i < rows_left) *reinterpret_cast<float4*>(dst) = t;
                dst += step;
              }
            } else {
              float* dst = p.out_f32 + (row0 + sub) * (size_t)p.ld_f32 + c0 + 4 * c4;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 t = lds_v4(((i & 1) ? rd_odd : rd_even) + (i >> 1) * 1024);
                t.x = fmaf(t.x, g4.x, b4.x); t.y = fmaf(t.y, g4.y, b4.y);
                t.z = fmaf(t.z, g4.z, b4.z); t.w = fmaf(t.w, g4.w, b4.w);
                if (all_rows || sub + 4 * i < rows_left) atomicAdd(reinterpret_cast<float4*>(dst), t);  // RED.ADD.F32x4
                dst += step;
              }
            }
            __syncwarp();  // the patch is rewritten by the next chunk
          }
        }
      } else {  // EPI_QKV: a warp owns 32 rows x one 128-column group per pass
        static_assert(EPI != EPI_QKV || BN == 256 || BN == 384, "QKV epilogue needs BN == 256 or 384");
        // BN = 256: warp half h takes group h. BN = 384: a second pass over group 2, whose four 32-column chunks are
        // shared by the two halves (both read the whole group once more for the row statistics).
        constexpr int PASSES = BN == 384 ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
        const int grp_t = pass == 0 ? half : 2;                    // 128-column group inside the tile
        const int ch_lo = pass == 0 ? 0 : 2 * half, ch_hi = pass == 0 ? 4 : 2 * half + 2;  // chunks this warp writes
        // first global column of the group: consecutive for 256-wide tiles, from the tile -> groups map for 384-wide ones
        const int g0 = BN == 384 ? (int)p.tile_groups[3 * nt + grp_t] * 128 : n0 + grp_t * 128;
        if (g0 < p.N) {
          const int si = g0 / p.sec_width;
          const int cs = g0 - si * p.sec_width;  // column inside the section
          const QkvSection sec = p.sec[si];
          const int grp = cs >> 7;
          const uint32_t stg = smem_u32(epi_stage + ew * 1024);
          const int sub = lane >> 3, c4 = lane & 7;
          const size_t row0 = (size_t)bt * p.M + mbase;
          const int rows_left = p.M - mbase;
          float rstd = 1.f;
          if (sec.norm_w) {
            // RMSNorm over head_dim columns (reference model.py:99-104); head_dim == 128 whenever norm_w is set.
            // Row statistics are taken in the row-per-thread layout tcgen05.ld delivers (no shuffles).
            float ss = 0.f;
ECHO_CHUNK_UNROLL
            for (int ch = 0; ch < 4; ++ch) {
              float v[32];
              tc_ld_32x32(tbase + (grp_t * 4 + ch) * 32, v);
              tc_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);
            }
            rstd = rsqrtf(ss * (1.f / 128.f) + p.eps);
          }
          const bool do_rope = grp < sec.rope_heads;
          const int hd2 = p.head_dim >> 1;
          // after the transpose this lane owns columns 4*c4 .. 4*c4+3 of rows sub + 4i: RoPE pairs stay inside a lane.
          // Table offsets of the 8 rows in 32-bit arithmetic with ONE division (the first version took eight 64-bit
          // modulos per tile, ~800 of the 2 400 instructions a warp spent on a tile); only the rotated groups pay it.
          int roff[8];
          if (do_rope) {
            const uint32_t period = (uint32_t)p.pos_period;
            uint32_t r = (uint32_t)(row0 + sub) % period;  // rows < 2^32
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              roff[i] = (p.pos_offset + p.pos_mult * (int)r) * hd2;
              r += 4;
              while (r >= period) r -= period;  // one subtraction unless the period is below 32 rows
            }
          }
          const uint32_t rd_even = stg + sub * 128 + ((c4 ^ sub) << 4);  // rows sub + 4i: (sub + 4i) & 7 = sub | sub + 4
          const uint32_t rd_odd = stg + (sub + 4) * 128 + ((c4 ^ (sub + 4)) << 4);
          const bool all_rows = rows_left >= 32;  // warp-uniform
          const size_t ostep = (size_t)4 * p.sec_width;
ECHO_CHUNK_UNROLL
          for (int ch = ch_lo; ch < ch_hi; ++ch) {
            float v[32];
            tc_ld_32x32(tbase + (grp_t * 4 + ch) * 32, v);
            const int cc = cs + ch * 32;
            float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f);
            if (sec.norm_w) w4 = __ldg(reinterpret_cast<const float4*>(sec.norm_w + cc + 4 * c4));
            const int pi = ((cc & (p.head_dim - 1)) >> 1) + 2 * c4;  // first of this lane's two rotation pairs (head_dim is 64 / 128)
            // cos/sin of the 8 rows, requested up front and unconditionally (positions are always in range): with
            // 227 KB of smem there is no L1 left, every table read is an L2 round trip, and loads issued row by row
            // behind a predicate were serialised (8 exposed L2 latencies per chunk).
            float2 rc[8], rs[8];
            if (do_rope) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                rc[i] = __ldg(reinterpret_cast<const float2*>(p.rope_cos + roff[i] + pi));
                rs[i] = __ldg(reinterpret_cast<const float2*>(p.rope_sin + roff[i] + pi));
              }
            }
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j] * rstd, v[4 * j + 1] * rstd, v[4 * j + 2] * rstd,
                     v[4 * j + 3] * rstd);
            __syncwarp();
            bf16* op = sec.out + (row0 + sub) * p.sec_width + cc + 4 * c4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 t = lds_v4(((i & 1) ? rd_odd : rd_even) + (i >> 1) * 1024);
              t.x *= w4.x; t.y *= w4.y; t.z *= w4.z; t.w *= w4.w;
              if (do_rope) {
                const float x0 = t.x, y0 = t.y, x1 = t.z, y1 = t.w;
                t.x = x0 * rc[i].x - y0 * rs[i].x; t.y = x0 * rs[i].x + y0 * rc[i].x;
                t.z = x1 * rc[i].y - y1 * rs[i].y; t.w = x1 * rs[i].y + y1 * rc[i].y;
              }
              if (sec.sigmoid) { t.x = sigmoid_fast(t.x); t.y = sigmoid_fast(t.y); t.z = sigmoid_fast(t.z); t.w = sigmoid_fast(t.w); }
              if (all_rows || sub + 4 * i < rows_left)
                *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16(t.x, t.y), pack_bf16(t.z, t.w));
              op += ostep;
            }
            __syncwarp();
          }
        }
        }  // pass
      }
      if (trace && warp == GEMM_EPI_WARP0 && lane == 0) {
        trace[6] = clock64();
        if (!RU && it < 4) trace[9 + 2 * it] = trace[6];
        if (RU && it == 2) trace[13] = trace[6];
        if (RU && it == 1) trace[14] = trace[6];  // end of the previous tile
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_pair_leader(&tempty_bar[as]);  // the MMA issuer lives in the even CTA
        else mbar_arrive(&tempty_bar[as]);
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA may free TMEM / exit while the peer still uses it
  else __syncthreads();
  if (trace && threadIdx.x == 0) trace[7] = clock64();
  if (warp == 2) {
    if constexpr (CG == 2) tmem_dealloc_pair<ACC_STAGES * ACC_STRIDE>(tmem_base);
    else tmem_dealloc<ACC_STAGES * ACC_STRIDE>(tmem_base);
  }
}

}  // namespace echo
