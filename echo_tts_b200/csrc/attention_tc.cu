// Every attention of the hot path on tcgen05 tensor cores, one kernel template over the head dim (128 or 64):
//   * joint attention of the DiT blocks: cat([self, latent, text, speaker]) + bool key mask +
//     F.scaled_dot_product_attention + "* sigmoid(gate)" (model.py:246-266), without materialising the concatenated
//     K/V or the 3x CFG copies of the text/speaker caches;
//   * encoder self-attention (model.py:141-154): key mask (text) or causal (speaker / latent encoders);
//   * the window-limited causal attention of the DAC transformers (autoencoder.py:698-702, 762-773; head_dim 64,
//     window 128 in post_module / pre_module, 512 in the encoder): the tile list of a CTA only holds the key tiles its
//     128 query rows can see, the diagonal / window edge tiles are masked per row.
// One (128 query rows, head, batch row) per CTA.
//
// Data path per 64-key tile j (stage = j & 1):
//   TMA        K_j, V_j  -> smem  (128B-swizzled boxes; rows beyond the tensor are zero-filled)
//   tcgen05    S_j = Q K_j^T       (M128 N64 K128, both operands K-major, fp32 in TMEM columns [64*stage, +64))
//   softmax    thread r owns row r: tcgen05.ld S_j -> mask/scale/max/exp2 -> bf16 P_j stored back into TMEM over the
//              first half of S_j's columns (tcgen05.st), lazy rescale of O only when the row max grows by > 2^8
//   tcgen05    O += P_j V_j        (M128 N128 K64, A = P from TMEM, B = V MN-major, fp32 in TMEM columns [128, 256))
// K_j's smem slot is released as soon as S_j has completed and V_j's when PV_j has, so the producer runs two tiles
// ahead of the tensor pipe (with P staged in smem over K_j the ring was one tile deep and every tile paid a TMA
// round trip: 21 us for 12 tiles at b = 1).
// The MMA thread issues S_{j+1} before it waits for P_j, so the tensor pipe computes the next scores while the
// softmax warps work; two CTAs are resident per SM (96 KB smem, 256 TMEM columns each), so one CTA's MMAs also
// overlap the other's softmax.
//
// Head dim 64: Q / K / V tiles are one 64-column atom instead of two, S = QK^T takes 4 MMAs (K = 64) and O is 64 TMEM columns.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = softmax,
// O correction and epilogue (TMEM lane quarter = warp & 3). The WG = 2 form (b = 1 joint attention, one CTA per SM, see
// the kernel) has 128-key tiles and a second softmax warpgroup in warps 6-9.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "attention.h"
#include "common.cuh"
#include "counters.h"
#include "gemm.h"
#include "launch.h"
#include "profiler.h"

namespace echo {

namespace {

constexpr int TQ = 128;  // query rows per CTA
constexpr int TK = 64;   // keys per tile
constexpr int AT_THREADS = 192;      // WG = 1: producer warp, MMA warp, one softmax warpgroup
constexpr int AT_THREADS_WG2 = 320;  // WG = 2: two softmax warpgroups, 64 keys of every 128-key tile each (one CTA per SM)
constexpr int AT_MAX_TILES = 120;
// per head dim D: Q = D/64 atoms [128 rows][64 d]; a K / V tile = D/64 atoms (boxes) [64 keys][64 d]
__host__ __device__ constexpr int at_q_bytes(int D) { return TQ * D * 2; }
__host__ __device__ constexpr int at_tile_keys(int WG) { return TK * WG; }  // keys per tile of the list: 64, or 128 with two warpgroups
__host__ __device__ constexpr int at_slot(int D, int WG = 1) { return at_tile_keys(WG) * D * 2; }
__host__ __device__ constexpr int at_tile_bytes(int D, int WG = 1) { return at_q_bytes(D) + 4 * at_slot(D, WG); }  // 96 KB (D = 128) / 48 KB; WG = 2: 160 KB
__host__ __device__ constexpr int at_smem(int D, int WG = 1) {
  // WG = 2 (160 KB of tiles): a second (max, sum) table; two CTAs never share an SM (each takes all of TMEM)
  return at_tile_bytes(D, WG) + 1024 /*align slack*/ + 1024 /*barriers + tile list*/ + 1024 * WG /*(max, sum) per row: split-KV / warpgroups*/;
}
constexpr float RESCALE_LOG2 = 8.f;  // O is rescaled only when a row max grows by more than 2^8

struct AttnMaps {
  CUtensorMap q;
  CUtensorMap k[4];
  CUtensorMap v[4];
};

struct SmemCtl {
  uint64_t q_full, k_full[2], v_full[2], s_full[2], p_ready[2], pv_done[2];
  uint32_t tmem_slot;
  int ntiles;   // tiles of THIS CTA (its share of the list when the keys are split over a cluster)
  int t_lo;     // first list entry of this CTA
  int seg_hi[4];
  int tiles[AT_MAX_TILES];
};
static_assert(sizeof(SmemCtl) <= 1024, "control block");

// WG = 2 (few CTAs, one per SM: the b = 1 steps of a request): key tiles of 128 and TWO softmax warpgroups. S = Q K^T is
// one M128 N128 MMA chain per tile (a K = 16 MMA costs the tensor pipe ~76 cycles whether N is 64 or 128 -- with 64-key
// tiles the pipe, not the softmax, was the period: profiles/r02_attn_two_warpgroups.txt); warpgroup g takes keys
// [64 g, 64 g + 64) of every tile, keeps its own running (max, sum) and its own O accumulator in TMEM (columns [256, 384)
// and [384, 512)) fed by its own P_g V_g MMAs, and the two partial results are merged in the epilogue like two split-KV
// shares, but inside the CTA (no cluster barrier, no DSMEM).
template <int D, int WG>
__global__ void __launch_bounds__(WG == 2 ? AT_THREADS_WG2 : AT_THREADS, WG == 2 ? 1 : 2)
attn_tc_kernel(const __grid_constant__ AttnMaps maps, const echo_attn_desc d) {
  static_assert(D == 128 || D == 64, "head dim");
  static_assert(WG == 1 || (WG == 2 && D == 128), "two warpgroups: head dim 128 only");
  constexpr int TMEM_COLS = WG == 2 ? 512 : 256;
  constexpr int NA = D / 64;  // 64-column atoms per row
  constexpr int TKT = at_tile_keys(WG);  // keys per tile of the list
  constexpr int Q_BYTES = at_q_bytes(D), KSLOT = at_slot(D, WG), VSLOT = at_slot(D, WG), STAGE = KSLOT + VSLOT;
  constexpr int TILE_BYTES = at_tile_bytes(D, WG);
  constexpr uint32_t O_BASE = 2 * TKT;  // TMEM: two score stages of TKT columns, then one O accumulator of D columns per warpgroup
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + TILE_BYTES);
  auto sK = [&](int st) { return smem + Q_BYTES + st * STAGE; };
  auto sV = [&](int st) { return smem + Q_BYTES + st * STAGE + KSLOT; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split-KV: a cluster of nsplit CTAs (consecutive blockIdx.x) shares one (query tile, head, batch row)
  const int nsplit = d.nsplit > 1 ? d.nsplit : 1;
  const int q0 = ((int)blockIdx.x / nsplit) * TQ, h = blockIdx.y, b = blockIdx.z, sp = (int)blockIdx.x % nsplit;
  // timeline (tuning): 0 entry, 1 setup done, 2 tile list done, 3 Q landed (MMA thread), 4+2j / 5+2j = softmax of
  // tile j starts (scores ready) / ends (P published), 30 O complete, 31 epilogue done   -- stamps of warp 2 lane 0
  long long* trace = d.trace ? d.trace + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64 : nullptr;
  if (trace && threadIdx.x == 64) trace[0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    for (int s = 0; s < d.nseg; ++s) { tma_prefetch_desc(&maps.k[s]); tma_prefetch_desc(&maps.v[s]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(&ctl->q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&ctl->k_full[s], 1);
        mbar_init(&ctl->v_full[s], 1);
        mbar_init(&ctl->s_full[s], 1);
        mbar_init(&ctl->p_ready[s], 4 * WG);
        mbar_init(&ctl->pv_done[s], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(&ctl->tmem_slot);
  }
  // ---- tile list, built BEFORE the dependency wait: it reads only the descriptor and eff_len[], which the kernel
  // chain computes once per request (mask_eff_len in sampler_prepare), never in the kernel right before this one.
  // (Contract: eff_len must not be written by an immediate predecessor that triggers its dependents early.)
  // (segment << 24 | first key) for every 64-key tile that can hold a valid key
  if (warp == 2) {
    int hi = 0, lo = 0;
    if (lane < d.nseg) {
      const echo_attn_segment& sg = d.seg[lane];
      hi = sg.len;
      if (sg.eff_len) { const int e = sg.eff_len[b]; hi = e < hi ? e : hi; }
      if (sg.pos_limit_mult > 0) {
        const int lim = (sg.pos_limit + sg.pos_limit_mult - 1) / sg.pos_limit_mult;  // keys j with j*mult < pos_limit
        hi = lim < hi ? lim : hi;
      }
      if (sg.causal) {  // keys this CTA's rows can see: j <= q (and j > q - window), q = q_offset + [q0, min(q0 + 128, S))
        const int qe = sg.q_offset + (q0 + TQ < d.S ? q0 + TQ : d.S);
        hi = qe < hi ? qe : hi;
        if (sg.window > 0) {
          lo = sg.q_offset + q0 - sg.window + 1;
          lo = lo < 0 ? 0 : (lo & ~(TK - 1));  // tile boundaries stay multiples of 64 on the KEY axis for every CTA
        }
      }
      if (hi < 0) hi = 0;
      if (lo > hi) lo = hi;
      ctl->seg_hi[lane] = hi;
    }
    const int nt = (hi - lo + TKT - 1) / TKT;
    int off = nt;  // inclusive prefix sum over the (<= 4) segment lanes
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, off, o);
      if (lane >= o) off += t;
    }
    const int total = __shfl_sync(0xffffffffu, off, 3);
    off -= nt;
    for (int i = 0; i < nt && off + i < AT_MAX_TILES; ++i) ctl->tiles[off + i] = (lane << 24) | (lo + i * TKT);
    if (lane == 0) {
      const int all = total < AT_MAX_TILES ? total : AT_MAX_TILES;
      const int t_lo = all * sp / nsplit, t_hi = all * (sp + 1) / nsplit;  // this CTA's share of the key tiles
      ctl->t_lo = t_lo;
      ctl->ntiles = t_hi - t_lo;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // only shared memory / TMEM / kernel parameters / eff_len were touched so far: overlaps the predecessor's tail
  pdl_wait();
  pdl_trigger();
  if (trace && threadIdx.x == 64) trace[1] = clock64();
  const int ntiles = lds_i32(&ctl->ntiles);
  if (warp == 0 && lane == 0 && ntiles > 0) {  // Q first; the K / V tiles follow from the producer loop below without another barrier
    mbar_expect_tx(&ctl->q_full, Q_BYTES);       // (a share without tiles loads nothing: no TMA may be in flight when it exits)
#pragma unroll
    for (int a = 0; a < NA; ++a) tma_load_3d(sQ + a * (Q_BYTES / NA), &maps.q, &ctl->q_full, h * D + a * 64, q0, b);
  }
  const int t_lo = lds_i32(&ctl->t_lo);
  const uint32_t tmem_base = (uint32_t)lds_i32(&ctl->tmem_slot);
  if (trace && threadIdx.x == 64) trace[2] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      auto load_kv = [&](int j, bool is_v) {
        const int st = j & 1, u = j >> 1;
        const int e = lds_i32(&ctl->tiles[t_lo + j]);
        const int si = e >> 24, n0 = e & 0xFFFFFF;
        const int bm = d.seg[si].batch_mod;
        const int cb = d.seg[si].batch_stride == 0 ? 0 : (bm > 0 ? b % bm : b);  // stride 0: one cache for all rows
        // a tile is WG boxes of 64 keys per atom: keys [64 g, 64 g + 64) sit 8 KB further down in each atom
        if (!is_v) {
          if (u > 0) mbar_wait(&ctl->s_full[st], (u - 1) & 1);  // S_{j-2} has consumed the slot
          mbar_expect_tx(&ctl->k_full[st], KSLOT);
#pragma unroll
          for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int g = 0; g < WG; ++g)
              tma_load_3d(sK(st) + a * (KSLOT / NA) + g * (TK * 128), &maps.k[si], &ctl->k_full[st], h * D + a * 64, n0 + g * TK, cb);
        } else {
          if (u > 0) mbar_wait(&ctl->pv_done[st], (u - 1) & 1);  // PV_{j-2} has consumed the slot
          mbar_expect_tx(&ctl->v_full[st], VSLOT);
#pragma unroll
          for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int g = 0; g < WG; ++g)
              tma_load_3d(sV(st) + a * (VSLOT / NA) + g * (TK * 128), &maps.v[si], &ctl->v_full[st], h * D + a * 64, n0 + g * TK, cb);
        }
      };
      // issue order K0 K1 V0 K2 V1 K3 ...: K_{j+2} only waits for S_j, which completes two tiles before S_{j+2} is issued
      if (ntiles > 0) load_kv(0, false);
      if (ntiles > 1) load_kv(1, false);
      for (int j = 0; j < ntiles; ++j) {
        load_kv(j, true);
        if (j + 2 < ntiles) load_kv(j + 2, false);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the (warp-uniform) control flow and one elected lane issues each tcgen05 instruction: the
    // descriptors then live in uniform registers. Per 64-key tile: 12 MMAs, 2 commits and 3 barrier waits (0.9 us per
    // tile before the issue path was trimmed, profiles/r01_attn_tc_timeline.txt). Round 2 measured what is left: two
    // back-to-back issues of 8 MMAs are 0.33 us apart whatever they wait for -- a K = 16 MMA occupies the pipe ~76 cycles
    // at N = 64 (nominal 32) -- so with two CTAs per SM the pipe's ~0.9 us per tile pair IS the period at b = 3
    // (profiles/r02_attn_two_warpgroups.txt); the WG = 2 form halves the S MMAs per key with N = 128.
    if (ntiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(TQ, TKT);
      constexpr uint32_t idesc_o = make_idesc_bf16(TQ, D) | kIdescBMajorMN;
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t t_o = tmem_base + O_BASE;
      const uint64_t qd0 = make_smem_desc<128>(q_addr), qd1 = make_smem_desc<128>(q_addr + Q_BYTES / NA * (NA - 1));
      mbar_wait(&ctl->q_full, 0);
      if (trace && lane == 0) trace[3] = clock64();
      auto issue_s = [&](int j) {
        const int st = j & 1, u = j >> 1;
        mbar_wait(&ctl->k_full[st], u & 1);
        tc_fence_after();
        if (trace && lane == 0 && j < 13) trace[32 + j] = clock64();
        const uint32_t k_addr = smem_u32(sK(st));
        const uint64_t kd0 = make_smem_desc<128>(k_addr), kd1 = make_smem_desc<128>(k_addr + KSLOT / NA * (NA - 1));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base + st * TKT, qd0 + 2 * k, kd0 + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          if constexpr (NA == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base + st * TKT, qd1 + 2 * k, kd1 + 2 * k, idesc_s, 1u);
          }
          tc_commit(&ctl->s_full[st]);  // scores ready; also releases K_j's smem slot to the producer
        }
        __syncwarp();
      };
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);  // scores of the next tile run under this tile's softmax
        const int st = j & 1, u = j >> 1;
        mbar_wait(&ctl->p_ready[st], u & 1);  // every softmax warp (of both warpgroups) has published its P
        mbar_wait(&ctl->v_full[st], u & 1);
        tc_fence_after();
        if (trace && lane == 0 && j < 13) trace[48 + j] = clock64();
        if (elect_one()) {
#pragma unroll
          for (int g = 0; g < WG; ++g) {
            // warpgroup g: bf16 P of keys [64 g, 64 g + 64), packed 2 keys per column over the first 32 columns of its half of
            // S_j, times the V rows of those keys (8 KB further down in each atom), into its own accumulator
            const uint32_t p_tmem = tmem_base + st * TKT + g * TK;
            const uint64_t vd = make_smem_desc_mn(smem_u32(sV(st)) + g * (TK * 128), VSLOT / NA, 1024);  // LBO = next 64-wide d group (D = 128 only)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // 16 keys per MMA: A advances 8 TMEM columns, B advances two 8-key atoms (2 KB = 128 x 16 B)
              tc_mma_f16_ts(t_o + g * 128, p_tmem + 8 * k, vd + 128 * k, idesc_o, (j | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&ctl->pv_done[st]);  // the accumulators hold tiles 0..j; also releases V_j's smem slot
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int quarter = warp & 3;
    const int wg = WG == 2 ? (warp - 2) >> 2 : 0;  // softmax warpgroup: warps 2-5 / 6-9
    const int row = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t o_col = O_BASE + (uint32_t)wg * 128u;  // this warpgroup's O accumulator
    const uint32_t s_col = (uint32_t)wg * TK;             // its 64 keys inside a score stage
    const float sl2 = d.scale * 1.4426950408889634f;
    float m_used = -INFINITY;  // running max in the log2 domain the exponentials are taken against
    float l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int st = j & 1, u = j >> 1;
      // ---- key validity of this warpgroup's 64 keys of the tile as a 64-bit mask (before the scores are ready)
      const int e = lds_i32(&ctl->tiles[t_lo + j]);
      const int si = e >> 24, n0 = (e & 0xFFFFFF) + wg * TK;
      const echo_attn_segment& sg = d.seg[si];
      const int hi = lds_i32(&ctl->seg_hi[si]);
      uint32_t vm_lo, vm_hi;
      {
        bool ok0 = (n0 + lane) < hi, ok1 = (n0 + 32 + lane) < hi;
        if (sg.mask) {
          const int cb = sg.batch_mod > 0 ? b % sg.batch_mod : b;
          const uint8_t* mp = sg.mask + (size_t)cb * sg.mask_ld;
          if (ok0) ok0 = mp[(size_t)(n0 + lane) * sg.mask_stride] != 0;
          if (ok1) ok1 = mp[(size_t)(n0 + 32 + lane) * sg.mask_stride] != 0;
        }
        vm_lo = __ballot_sync(0xffffffffu, ok0);
        vm_hi = __ballot_sync(0xffffffffu, ok1);
      }
      mbar_wait(&ctl->s_full[st], u & 1);
      tc_fence_after();
      if (trace && threadIdx.x == 64 && j < 13) trace[4 + 2 * j] = clock64();
      float v[64];
      tc_ld_32x32(tmem_base + lane_base + st * TKT + s_col, v);
      tc_ld_32x32(tmem_base + lane_base + st * TKT + s_col + 32, v + 32);
      tc_wait_ld();
      if (trace && threadIdx.x == 64 && j == 5) trace[27] = clock64();
      // causal / window segments: rows of this warp's slab see keys j with q - window < j <= q. Tiles entirely below the
      // diagonal and inside the window of all 32 rows take the row-uniform path; edge tiles are masked per row.
      int r_lo = 0, r_hi = TK;  // this row's valid key range inside the tile: [r_lo, r_hi)
      bool per_row = false;     // warp-uniform
      if (sg.causal) {
        const int qw0 = sg.q_offset + q0 + quarter * 32;  // first query row of the slab, on the key axis
        per_row = (n0 + TK - 1 > qw0) || (sg.window > 0 && n0 < qw0 + 31 - sg.window + 1);
        const int q = sg.q_offset + q0 + row;
        r_hi = q - n0 + 1;
        if (sg.window > 0) r_lo = q - sg.window + 1 - n0;
      }
      const bool full = (vm_lo & vm_hi) == 0xffffffffu && !per_row;  // warp-uniform
      float tmax = -INFINITY;
      if (full) {
#pragma unroll
        for (int i = 0; i < 64; ++i) tmax = fmaxf(tmax, v[i]);
      } else if (!per_row) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const bool ok = ((i < 32 ? vm_lo >> i : vm_hi >> (i - 32)) & 1u) != 0;
          v[i] = ok ? v[i] : -INFINITY;
          tmax = fmaxf(tmax, v[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const bool ok = ((i < 32 ? vm_lo >> i : vm_hi >> (i - 32)) & 1u) != 0 && i >= r_lo && i < r_hi;
          v[i] = ok ? v[i] : -INFINITY;
          tmax = fmaxf(tmax, v[i]);
        }
      }
      if (trace && threadIdx.x == 64 && j == 5) trace[28] = clock64();
      // speaker_kv_scale: a segment whose K and V count as multiplied by kv_scale -- the scores scale with it (folded into
      // the softmax scale of this tile), and so does its share of P V (folded into P below; the row sum stays unscaled)
      const float kvs = (sg.kv_scale != 0.f) ? sg.kv_scale : 1.f;
      const float sl2t = sl2 * kvs;
      const float m_new = fmaxf(m_used, tmax * sl2t);
      if (j == 0) {
        m_used = m_new;
      } else {
        const bool grow = m_new > m_used + RESCALE_LOG2;  // also true for -inf -> finite
        if (__any_sync(0xffffffffu, grow)) {
          float corr = 1.f;
          if (grow) {
            corr = (m_used == -INFINITY) ? 0.f : fast_exp2(m_used - m_new);
            m_used = m_new;
            l_run *= corr;
          }
          // O holds tiles 0..j-1 once PV_{j-1} has completed; PV_j is only issued after this warp's p_ready arrive
          mbar_wait(&ctl->pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            float o[32];
            tc_ld_32x32(tmem_base + lane_base + o_col + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= corr;
            tc_st_32x32(tmem_base + lane_base + o_col + c * 32, o);
          }
          tc_wait_st();
        }
      }
      if (trace && threadIdx.x == 64 && j == 5) trace[29] = clock64();
      const float muse = (m_used == -INFINITY) ? 0.f : m_used;
      // ---- P = exp2(s * scale*log2e - m) as bf16 pairs, stored over the first 32 columns of this row's S_j
      // The softmax warps run one per scheduler, so this loop is bound by its instruction count: the scale / subtract and
      // the row sum run on fp32 PAIRS (FFMA2 / FADD2: 32 + 32 instead of 64 + 64 issue slots per tile), and the kv_scale
      // multiply only exists for a scaled segment (warp-uniform).
      float pk[32];
      {
        const float2 s2 = make_float2(sl2t, sl2t), nm2 = make_float2(-muse, -muse);
        float2 ls2 = make_float2(0.f, 0.f);
        if (kvs == 1.f) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float2 a = f2fma(make_float2(v[2 * c], v[2 * c + 1]), s2, nm2);
            const float2 pp = make_float2(fast_exp2(a.x), fast_exp2(a.y));  // -inf -> 0
            ls2 = f2add(ls2, pp);
            pk[c] = __uint_as_float(pack_bf16(pp.x, pp.y));
          }
        } else {
          const float2 k2 = make_float2(kvs, kvs);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float2 a = f2fma(make_float2(v[2 * c], v[2 * c + 1]), s2, nm2);
            float2 pp = make_float2(fast_exp2(a.x), fast_exp2(a.y));
            ls2 = f2add(ls2, pp);
            pp = f2mul(pp, k2);
            pk[c] = __uint_as_float(pack_bf16(pp.x, pp.y));
          }
        }
        l_run += ls2.x + ls2.y;
      }
      if (trace && threadIdx.x == 64 && j == 5) trace[61] = clock64();
      tc_st_32x32(tmem_base + lane_base + st * TKT + s_col, pk);
      tc_wait_st();
      if (trace && threadIdx.x == 64 && j == 5) trace[62] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->p_ready[st]);
      if (trace && threadIdx.x == 64 && j < 13) trace[5 + 2 * j] = clock64();
    }

    if constexpr (WG == 2) {
      // ---- epilogue of the two-warpgroup form (never split over a cluster): merge the two partial results inside the CTA.
      // Warpgroup g fed accumulator g from its 64 keys of every tile; thread `row` of BOTH warpgroups owns that query row.
      constexpr int CPR = D / 8, RPI = 32 / CPR, NIT = 32 / RPI, NH = NIT / 2, ROWB = D * 2;
      const size_t HD = (size_t)d.H * D;
      const int rsel = lane / CPR, ch = lane % CPR;
      const int rbase = wg * (32 / 2);  // rows of the quarter's 32 this warp stores: [16 wg, 16 wg + 16)
      const size_t off0 = ((size_t)b * d.S + q0 + quarter * 32 + rbase + rsel) * HD + (size_t)h * D + ch * 8;
      const int rows_ok = d.S - (q0 + quarter * 32);
      uint4 gv[NH];
      if (d.gate) {
#pragma unroll
        for (int i = 0; i < NH; ++i)
          if (rbase + RPI * i + rsel < rows_ok)
            gv[i] = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(d.gate) + off0 + (size_t)(RPI * i) * HD);
      }
      const bool have_g = ntiles > 0;  // warp-uniform
      if (have_g) {
        mbar_wait(&ctl->pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
        tc_fence_after();
      }
      if (trace && threadIdx.x == 64) trace[30] = clock64();
      const uint32_t ml_mine = smem_u32(smem + TILE_BYTES + 1024 + wg * 1024), ml_other = smem_u32(smem + TILE_BYTES + 1024 + (1 - wg) * 1024);
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ml_mine + row * 8), "f"(have_g ? m_used : -INFINITY), "f"(have_g ? l_run : 0.f) : "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");  // both warpgroups are through their last PV: K / V stages are dead too
      float m_ot, l_ot;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_ot), "=f"(l_ot) : "r"(ml_other + row * 8) : "memory");
      const float l_me = have_g ? l_run : 0.f;
      float mx = -INFINITY;
      if (l_me > 0.f) mx = m_used;
      if (l_ot > 0.f) mx = fmaxf(mx, m_ot);
      const float w_me = l_me > 0.f ? fast_exp2(m_used - mx) : 0.f, w_ot = l_ot > 0.f ? fast_exp2(m_ot - mx) : 0.f;
      const float l_tot = fmaf(l_me, w_me, l_ot * w_ot);
      const float iv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      const float w0 = (wg == 0 ? w_me : w_ot) * iv, w1 = (wg == 0 ? w_ot : w_me) * iv;  // weights of accumulators 0 / 1
      const bool has0 = ntiles > 0, has1 = ntiles > 0;  // accumulators no tile ever wrote hold stale TMEM: never read
      const uint32_t stg = smem_u32(smem + Q_BYTES + quarter * (32 * ROWB));  // the quarter's 32 staged rows, shared by its two warps
      const uint32_t srow = stg + lane * ROWB;
#pragma unroll 1
      for (int cc = 0; cc < D / 64; ++cc) {
        const int c = wg * (D / 64) + cc;  // this warpgroup's half of the head dim, 32 columns at a time
        float oa[32], ob[32];
        if (has0) tc_ld_32x32(tmem_base + lane_base + O_BASE + c * 32, oa);
        if (has1) tc_ld_32x32(tmem_base + lane_base + O_BASE + 128 + c * 32, ob);
        if (has0 || has1) tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) oa[i] = (has0 ? oa[i] * w0 : 0.f) + (has1 ? ob[i] * w1 : 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int chk = c * 4 + k;  // 16-byte chunk of the row
          sts_u4(srow + ((chk ^ (lane & 7)) << 4),
                 make_uint4(pack_bf16(oa[8 * k], oa[8 * k + 1]), pack_bf16(oa[8 * k + 2], oa[8 * k + 3]),
                            pack_bf16(oa[8 * k + 4], oa[8 * k + 5]), pack_bf16(oa[8 * k + 6], oa[8 * k + 7])));
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // both halves of every row are staged
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        const int r = rbase + RPI * i + rsel;
        if (r < rows_ok) {
          uint4 val = lds_u4(stg + r * ROWB + ((ch ^ (r & 7)) << 4));
          if (d.gate) {
            const uint32_t* vi = reinterpret_cast<const uint32_t*>(&val);
            const uint32_t* gi = reinterpret_cast<const uint32_t*>(&gv[i]);
            uint32_t rr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 a = unpack_bf16(vi[k]), g = unpack_bf16(gi[k]);
              rr[k] = pack_bf16(a.x * g.x, a.y * g.y);
            }
            val = make_uint4(rr[0], rr[1], rr[2], rr[3]);
          }
          *reinterpret_cast<uint4*>(static_cast<bf16*>(d.out) + off0 + (size_t)(RPI * i) * HD) = val;
        }
      }
    } else {
    // ---- epilogue: O / l -> bf16 -> (* gate) -> global, staged through smem so rows are written as 256 B runs.
    // The 16 gate vectors this lane needs are requested BEFORE waiting for the last PV (4.3 us of exposed L2 round
    // trips otherwise: the loop below was load -> multiply -> store, one row pair at a time).
    const bool have = ntiles > 0;
    constexpr int CPR = D / 8;      // 16-byte chunks per output row (one head)
    constexpr int RPI = 32 / CPR;   // rows written per warp instruction (2 for D = 128, 4 for D = 64)
    constexpr int NIT = 32 / RPI;
    constexpr int ROWB = D * 2;     // bytes of one staged row
    const size_t HD = (size_t)d.H * D;
    const int rsel = lane / CPR, ch = lane % CPR;
    const size_t off0 = ((size_t)b * d.S + q0 + quarter * 32 + rsel) * HD + (size_t)h * D + ch * 8;
    const int rows_ok = d.S - (q0 + quarter * 32);  // rows of this warp's slab inside the sequence
    uint4 gv[NIT];
    if (d.gate && nsplit == 1) {
#pragma unroll
      for (int i = 0; i < NIT; ++i)
        if (RPI * i + rsel < rows_ok) gv[i] = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(d.gate) + off0 + (size_t)(RPI * i) * HD);
    }
    if (have) {
      mbar_wait(&ctl->pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
      tc_fence_after();
    }
    if (trace && threadIdx.x == 64) trace[30] = clock64();
    if (nsplit > 1) {
      // ---- split-KV, part 1: this share's un-normalised O slab goes into its own shared memory (over the dead K / V
      // stages) as bf16 -- the merge reads it through distributed shared memory, which moves only ~20 B / clk / SM, so
      // the slab is kept small (fp32 slabs made the merge of a 3-way split take 5 us) -- and the fp32 (max, sum) per row
      // behind the control block. bf16 has fp32's exponent range: the lazily rescaled O (up to 2^8 x #keys) cannot overflow.
      constexpr int ROWS = D * 2;  // bytes of one staged row
      const uint32_t slab = smem_u32(smem + Q_BYTES);
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(smem_u32(smem + TILE_BYTES + 1024) + row * 8), "f"(have ? m_used : -INFINITY),
                   "f"(have ? l_run : 0.f) : "memory");
      if (have) {
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          float o[32];
          tc_ld_32x32(tmem_base + lane_base + 128 + c * 32, o);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts_u4(slab + row * ROWS + (((c * 4 + j) ^ (row & 7)) << 4),
                   make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                              pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7])));
        }
      }
      if (trace && threadIdx.x == 64) trace[24] = clock64();
    } else {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const uint32_t stg = smem_u32(smem + Q_BYTES + (warp - 2) * (32 * ROWB));  // 32 staged rows, private to this warp; all tiles are dead
    const uint32_t srow = stg + lane * ROWB;
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      float o[32];
      if (have) {
        tc_ld_32x32(tmem_base + lane_base + 128 + c * 32, o);
        tc_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int chk = c * 4 + cc;  // 16-byte chunk of the row
        sts_u4(srow + ((chk ^ (lane & 7)) << 4),
               make_uint4(pack_bf16(o[8 * cc] * inv, o[8 * cc + 1] * inv), pack_bf16(o[8 * cc + 2] * inv, o[8 * cc + 3] * inv),
                          pack_bf16(o[8 * cc + 4] * inv, o[8 * cc + 5] * inv), pack_bf16(o[8 * cc + 6] * inv, o[8 * cc + 7] * inv)));
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int r = RPI * i + rsel;
      if (r < rows_ok) {
        uint4 val = lds_u4(stg + r * ROWB + ((ch ^ (r & 7)) << 4));
        if (d.gate) {
          const uint32_t* vi = reinterpret_cast<const uint32_t*>(&val);
          const uint32_t* gi = reinterpret_cast<const uint32_t*>(&gv[i]);
          uint32_t rr[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 a = unpack_bf16(vi[k]), g = unpack_bf16(gi[k]);
            rr[k] = pack_bf16(a.x * g.x, a.y * g.y);
          }
          val = make_uint4(rr[0], rr[1], rr[2], rr[3]);
        }
        *reinterpret_cast<uint4*>(static_cast<bf16*>(d.out) + off0 + (size_t)(RPI * i) * HD) = val;
      }
    }
    }  // nsplit == 1
    }  // WG == 1
  }

  if (nsplit > 1) {
    // ---- split-KV, part 2. Every thread of the cluster arrives here once its share is staged. CTA `sp` then merges
    // rows [128 sp / nsplit, 128 (sp + 1) / nsplit) of the query tile from the nsplit slabs (its own and its peers',
    // read through DSMEM) in share order -- no global workspace, no atomics, bit-reproducible -- and writes them.
    constexpr int ROWS = D * 2;    // bytes of one staged bf16 row
    constexpr int LPR = D / 8;     // lanes per row, 8 columns (16 bytes) each: 16 (D = 128) / 8 (D = 64)
    constexpr int RPW = 32 / LPR;  // rows per warp instruction
    constexpr int MAXIT = (TQ / 2) / (4 * RPW);  // rows a lane handles at most (nsplit = 2: 64 rows over 4 warps)
    const int r_lo = TQ * sp / nsplit, r_hi = TQ * (sp + 1) / nsplit;
    const int rsub = lane / LPR, c16 = lane % LPR;
    const size_t HDm = (size_t)d.H * D;
    // the gate of this CTA's output rows does not depend on the peers: requested before the barrier (one exposed L2 round
    // trip per row otherwise -- the first version of this merge took 10 us for 43 rows, profiles/r02_attn_split_kv.txt)
    uint4 gq[MAXIT];
    if (warp >= 2 && d.gate) {
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        const int r = r_lo + (warp - 2) * RPW + rsub + it * 4 * RPW;
        if (r < r_hi && q0 + r < d.S)
          gq[it] = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(d.gate) + ((size_t)b * d.S + q0 + r) * HDm + (size_t)h * D + c16 * 8);
      }
    }
    cluster_sync_all();
    if (trace && threadIdx.x == 64) trace[25] = clock64();
    if (warp >= 2) {
      const uint32_t ml_local = smem_u32(smem + TILE_BYTES + 1024), slab_local = smem_u32(smem + Q_BYTES);
      const uint32_t wtab = smem_u32(sQ);  // [row of the slice][8] normalised merge weights (Q is dead; empty shares never loaded it)
      uint32_t slab_peer[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) slab_peer[i] = dsmem_addr(slab_local, i < nsplit ? i : 0);
      // ---- phase A: one thread per row of the slice turns the shares' (max, sum) into normalised weights
      {
        const int rr = (warp - 2) * 32 + lane;
        if (rr < r_hi - r_lo) {
          float2 ml[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nsplit) ml[i] = ld_dsmem_v2(dsmem_addr(ml_local, i) + (r_lo + rr) * 8);
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nsplit && ml[i].y > 0.f) mx = fmaxf(mx, ml[i].x);
          float w[8], l_tot = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            w[i] = (i < nsplit && ml[i].y > 0.f) ? fast_exp2(ml[i].x - mx) : 0.f;  // a share without a valid key for this row: weight 0
            l_tot = fmaf(w[i], i < nsplit ? ml[i].y : 0.f, l_tot);
          }
          const float iv = l_tot > 0.f ? 1.f / l_tot : 0.f;
          sts_v4(wtab + rr * 32, w[0] * iv, w[1] * iv, w[2] * iv, w[3] * iv);
          sts_v4(wtab + rr * 32 + 16, w[4] * iv, w[5] * iv, w[6] * iv, w[7] * iv);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four softmax warps
      // ---- phase B: warp w takes rows r_lo + (w - 2) * RPW + rsub + it * 4 * RPW; two rows (x up to 8 shares) in flight
#pragma unroll
      for (int it0 = 0; it0 < MAXIT; it0 += 2) {
        uint4 t[2][8];
        int rws[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          rws[u] = r_lo + (warp - 2) * RPW + rsub + (it0 + u) * 4 * RPW;
          if (rws[u] < r_hi) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < nsplit) {
                const float4 f = ld_dsmem_v4(slab_peer[i] + rws[u] * ROWS + ((c16 ^ (rws[u] & 7)) << 4));
                t[u][i] = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = rws[u];
          if (r < r_hi) {
            const float4 w0 = lds_v4(wtab + (r - r_lo) * 32), w1 = lds_v4(wtab + (r - r_lo) * 32 + 16);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < nsplit && w[i] > 0.f) {  // a share that staged nothing usable for this row holds stale bytes: skip, never multiply
                const uint32_t* tv = reinterpret_cast<const uint32_t*>(&t[u][i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = unpack_bf16(tv[k]);
                  acc[2 * k] = fmaf(w[i], f.x, acc[2 * k]);
                  acc[2 * k + 1] = fmaf(w[i], f.y, acc[2 * k + 1]);
                }
              }
            }
            if (q0 + r < d.S) {
              const size_t ooff = ((size_t)b * d.S + q0 + r) * HDm + (size_t)h * D + c16 * 8;
              // two roundings like the unsplit path: O / l -> bf16, then (* gate) -> bf16
              uint32_t val[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) val[k] = pack_bf16(acc[2 * k], acc[2 * k + 1]);
              if (d.gate) {
                const uint32_t* gi = reinterpret_cast<const uint32_t*>(&gq[it0 + u]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 a = unpack_bf16(val[k]), g = unpack_bf16(gi[k]);
                  val[k] = pack_bf16(a.x * g.x, a.y * g.y);
                }
              }
              *reinterpret_cast<uint4*>(static_cast<bf16*>(d.out) + ooff) = make_uint4(val[0], val[1], val[2], val[3]);
            }
          }
        }
      }
    }
    if (trace && threadIdx.x == 64) trace[26] = clock64();
    cluster_sync_all();  // no CTA may exit (and free its shared memory) while a peer still reads it
  }
  if (trace && threadIdx.x == 64) trace[31] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

// Worst-case number of 64-key tiles one CTA walks, or -1 when the descriptor cannot run on this kernel.
static int attention_tile_bound(const echo_attn_desc& d) {
  if ((d.D != 128 && d.D != 64) || d.nseg < 1 || d.nseg > 4) return -1;
  if ((reinterpret_cast<uintptr_t>(d.Q) & 15) || (d.q_row_stride % 8) || (d.q_batch_stride % 8)) return -1;
  int tiles = 0;
  for (int i = 0; i < d.nseg; ++i) {
    const echo_attn_segment& g = d.seg[i];
    if (g.len <= 0 || g.len > (1 << 24) - 1) return -1;
    if ((reinterpret_cast<uintptr_t>(g.K) & 15) || (reinterpret_cast<uintptr_t>(g.V) & 15) || (g.row_stride % 8) ||
        (g.batch_stride % 8))
      return -1;
    int keys = g.len;
    if (g.causal) {
      keys = keys < d.S ? keys : d.S;
      if (g.window > 0 && g.window + TQ + TK < keys) keys = g.window + TQ + TK;  // window + the CTA's rows + alignment
    }
    tiles += (keys + TK - 1) / TK;
  }
  return tiles;
}

template <int D, int WG = 1>
static cudaError_t attention_tc_launch_d(const echo_attn_desc& d_in, int nsplit, cudaStream_t s) {
  echo_attn_desc d = d_in;
  d.nsplit = nsplit;
  static std::atomic<uint64_t> configured{0};  // per device, see ensure_dyn_smem
  {
    cudaError_t e = ensure_dyn_smem(configured, attn_tc_kernel<D, WG>, at_smem(D, WG));
    if (e != cudaSuccess) return e;
  }
  AttnMaps maps;
  std::memset(&maps, 0, sizeof(maps));
  const uint64_t W = (uint64_t)d.H * D;
  const uint64_t qbs = d.b > 1 ? (uint64_t)d.q_batch_stride : (uint64_t)d.S * d.q_row_stride;
  if (!tma_map_bf16(&maps.q, d.Q, 3, W, (uint64_t)d.S, (uint64_t)d.b, (uint64_t)d.q_row_stride * 2, qbs * 2, 64, TQ, 128))
    return cudaErrorInvalidValue;
  for (int i = 0; i < d.nseg; ++i) {
    const echo_attn_segment& g = d.seg[i];
    const uint64_t nb = g.batch_stride == 0 ? 1 : (g.batch_mod > 0 ? (uint64_t)g.batch_mod : (uint64_t)d.b);
    const uint64_t bs = nb > 1 ? (uint64_t)g.batch_stride : (uint64_t)g.len * g.row_stride;
    if (!tma_map_bf16(&maps.k[i], g.K, 3, W, (uint64_t)g.len, nb, (uint64_t)g.row_stride * 2, bs * 2, 64, TK, 128) ||
        !tma_map_bf16(&maps.v[i], g.V, 3, W, (uint64_t)g.len, nb, (uint64_t)g.row_stride * 2, bs * 2, 64, TK, 128))
      return cudaErrorInvalidValue;
  }
  for (int i = d.nseg; i < 4; ++i) { maps.k[i] = maps.k[0]; maps.v[i] = maps.v[0]; }
  dim3 grid(((d.S + TQ - 1) / TQ) * nsplit, d.H, d.b);  // the nsplit shares of a query tile are one cluster along x
  cudaError_t err;
  {
    char tag[64];
    if (prof_enabled()) snprintf(tag, sizeof(tag), "attn_tc D=%d b=%d S=%d H=%d nseg=%d split=%d wg=%d", D, d.b, d.S, d.H, d.nseg, nsplit, WG);
    else tag[0] = 0;
    double keys = 0;  // keys a query can see (upper bound for masked segments)
    for (int i = 0; i < d.nseg; ++i) {
      const echo_attn_segment& g = d.seg[i];
      keys += g.causal ? (g.window > 0 && g.window < d.S ? (double)g.window : 0.5 * d.S) : (double)g.len;
    }
    ProfScope ps(PROF_ATTN, 4.0 * d.b * d.H * (double)d.S * keys * D, 0.0, s, tag);
    err = launch_k(attn_tc_kernel<D, WG>, grid, dim3(WG == 2 ? AT_THREADS_WG2 : AT_THREADS), (size_t)at_smem(D, WG), s, nsplit, maps, d);
  }
  count_launch();
  return err;
}

// The one attention entry of the library. A descriptor whose worst-case tile list does not fit the kernel's control
// block (more than AT_MAX_TILES * 64 = 7680 visible keys per query tile) is rejected, never truncated.
cudaError_t attention_launch(const echo_attn_desc& d, cudaStream_t s) {
  if (d.b <= 0 || d.S <= 0 || d.H <= 0) return cudaErrorInvalidValue;
  const int tiles = attention_tile_bound(d);
  if (tiles < 0 || tiles > AT_MAX_TILES) return cudaErrorInvalidValue;
  // Split-KV: with few CTAs walking long key lists the kernel is bound by the per-tile latency chain of each CTA
  // (~0.7 us per 64 keys) while most SMs idle. The key tiles of every (query tile, head, batch row) are then divided
  // over a thread-block cluster of nsplit CTAs (<= 8), which merge their partial (O, max, sum) through distributed
  // shared memory. Auto: fill the 2 CTA / SM slots, keep >= 4 tiles per share; the merge costs ~2-3 us.
  int nsplit = 1;
  const size_t groups = (size_t)((d.S + TQ - 1) / TQ) * d.H * d.b;
  static const int env_split = [] { const char* e = std::getenv("ECHO_ATTN_SPLIT"); return e ? atoi(e) : 0; }();  // tuning: 1 = off
  if (d.nsplit > 1) {
    nsplit = d.nsplit;
  } else if (d.nsplit == 0 && env_split != 1) {
    if (env_split > 1) {
      nsplit = env_split;
    } else {
      const size_t slots = 2 * (size_t)gemm_num_sms();
      if (groups * 2 <= slots && tiles >= 8) {
        nsplit = (int)(slots / groups);
        if (nsplit > tiles / 4) nsplit = tiles / 4;
      }
    }
  }
  if (nsplit > 8) nsplit = 8;
  if (nsplit > tiles) nsplit = tiles;
  if (nsplit < 1) nsplit = 1;
  // 128-key tiles and two softmax warpgroups per CTA (one CTA per SM) instead of a split over a cluster when at most one
  // CTA per SM is there to begin with and the key list is short enough that the in-CTA form wins (the b = 1 steps of a
  // 640-latent request: 80 CTAs x ~12 tiles): no cluster barrier, no DSMEM merge. Joint attention only -- the encoders
  // keep one form whatever their length, so that their rows do not depend on how many padded rows are computed next to
  // them. nsplit = -2 in the descriptor forces it (tests), ECHO_ATTN_WG2=0 switches it off.
  static const int env_wg2 = [] { const char* e = std::getenv("ECHO_ATTN_WG2"); return e ? atoi(e) : 1; }();
  bool wg2_ok = d.D == 128;
  for (int i = 0; i < d.nseg; ++i) wg2_ok = wg2_ok && !d.seg[i].causal;
  const bool wg2 = wg2_ok && (d.nsplit == -2 || (d.nsplit == 0 && env_wg2 != 0 && d.nseg >= 2 &&
                                                 groups <= (size_t)gemm_num_sms() && tiles <= 28));
  if (wg2) return attention_tc_launch_d<128, 2>(d, 1, s);
  return d.D == 128 ? attention_tc_launch_d<128>(d, nsplit, s) : attention_tc_launch_d<64>(d, nsplit, s);
}

}  // namespace echo
