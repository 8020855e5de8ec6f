// Host side of the tcgen05 GEMM: TMA tensor-map construction (cached), tile-shape selection, launch.
#include "gemm.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "counters.h"
#include "profiler.h"
#include "gemm_tc.cuh"

namespace echo {

static inline bool b_batched_rows_check(const GemmParams& p) { return p.N % 128 != 0 || p.N / 128 > 250 || p.sec_width % 128 != 0; }

// ------------------------------------------------------------------ driver entry point for tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t ptr, d0, d1, d2, s1, s2;
  uint32_t b0, b1, rank, swz;
  bool operator==(const MapKey& o) const { return std::memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

// bf16 tensor, dim0 contiguous. rank 2: (d0, d1) ; rank 3: (d0, d1, d2). Strides in BYTES for dims >= 1.
static bool get_tensor_map(CUtensorMap* out, const void* ptr, int rank, uint64_t d0, uint64_t d1, uint64_t d2,
                           uint64_t s1, uint64_t s2, uint32_t box0, uint32_t box1, int row_bytes) {
  MapKey key;
  std::memset(&key, 0, sizeof(key));
  key.ptr = (uint64_t)ptr; key.d0 = d0; key.d1 = d1; key.d2 = d2; key.s1 = s1; key.s2 = s2;
  key.b0 = box0; key.b1 = box1; key.rank = rank; key.swz = row_bytes;
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return true; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                           : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps[key] = m;
  }
  *out = m;
  return true;
}

bool tma_map_bf16(void* out, const void* ptr, int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                  uint32_t box0, uint32_t box1, int row_bytes) {
  return get_tensor_map(static_cast<CUtensorMap*>(out), ptr, rank, d0, d1, d2, s1, s2, box0, box1, row_bytes);
}

// ECHO_DETERMINISTIC=1 in the environment is the same as calling echo_set_deterministic(1) at start-up
static std::atomic<int> g_deterministic{[] { const char* e = std::getenv("ECHO_DETERMINISTIC"); return (e && atoi(e) != 0) ? 1 : 0; }()};
void gemm_set_deterministic(int on) { g_deterministic.store(on ? 1 : 0); }
int gemm_get_deterministic() { return g_deterministic.load(); }

// SM count of the CURRENT device, cached per device (one process may drive several GPUs)
int gemm_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::atomic<int>& slot = cache[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

template <int BN, int BK, int ATOMS, int EPI, int CG>
static cudaError_t launch_inst(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t s,
                               const CUtensorMap* mb1 = nullptr) {
  static std::atomic<uint64_t> configured{0};  // per device, see ensure_dyn_smem
  constexpr int SMEM = gemm_smem_bytes(BN, BK, ATOMS, CG, EPI);
  auto kern = gemm_tc_kernel<BN, BK, ATOMS, EPI, CG>;
  {
    cudaError_t e = ensure_dyn_smem(configured, kern, SMEM);
    if (e != cudaSuccess) return e;
  }
  const int tiles_n = (BN == 384 && p.qkv_tiles > 0) ? p.qkv_tiles : (p.N + BN - 1) / BN;
  const int tiles = ((p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG)) * p.batches * tiles_n * (p.split_k > 1 ? p.split_k : 1);
  const int slots = gemm_num_sms() / CG;
  const int grid = (tiles < slots ? tiles : slots) * CG;
  cudaError_t err;
  {
    char tag[96];
    tag[0] = 0;
    if (prof_enabled())
      snprintf(tag, sizeof(tag), "gemm epi=%d M=%d x%d N=%d K=%d taps=%d bn=%d cg=%d sk=%d", EPI, p.M, p.batches, p.N,
               p.Kc, p.taps, BN, CG, p.split_k > 1 ? p.split_k : 1);
    ProfScope ps(PROF_GEMM, 2.0 * p.M * p.batches * (double)p.N * (double)p.Kc * p.taps, 0.0, s, tag);
    err = launch_k(kern, dim3(grid), dim3(GEMM_THREADS), SMEM, s, CG, ma, mb, mb1 ? *mb1 : mb, p);
  }
  count_launch();
  return err;
}

// Tile-shape choice. Cost model per K=16 slice, in SM cycles: the MMA needs BN/2 cycles (128 x BN per CTA), operand
// ingest (~64 B/clk/SM) needs (128 + BN/CG)/2, whichever is larger; times the number of waves over the 148 SMs.
struct TileCfg { int bn, cg; };
static TileCfg pick_cfg(const GemmCall& c) {
  const GemmParams& p = c.p;
  const int sms = gemm_num_sms();
  // CTA pairs (cta_group::2, 256-row tiles, B split across the pair) cut the per-SM operand ingest from 48 KB to
  // 32 KB per K block. Measured on B200 (profiles/r01_gemm_cta_pair_vs_single.txt): +6-7 % for the multi-tile QKV and
  // SwiGLU GEMMs, neutral-to-negative for the single-wave generic GEMMs (M = 640 / 1920 wastes half a pair tile
  // per column), so those stay opt-in.
  static const int env_cg_generic = [] { const char* e = std::getenv("ECHO_GEMM_CG_GENERIC"); return e ? atoi(e) : 0; }();
  // generic GEMMs: pairs pay off once there are enough rows to fill pair tiles (M = 1920: w2 48.0 -> 42.6 us,
  // wo 24.3 -> 23.1 us); at M = 640 they lose ~2-4 % (2.5 pair rows), so auto keeps single CTAs below 1024 rows
  const bool allow_pair = (c.cg == 2 || (c.cg == 0 && env_cg_generic != 1 && (env_cg_generic == 2 || p.M >= 1024))) &&
                          (p.N % 128 == 0);
  auto cost = [&](int bn, int cg) -> double {
    const long units = (long)((p.M + 128 * cg - 1) / (128 * cg)) * p.batches * ((p.N + bn - 1) / bn);
    const long slots = sms / cg;
    const long waves = (units + slots - 1) / slots;
    const double per = std::max(bn / 2.0, (128.0 + bn / (double)cg) / 2.0);
    return waves * per + 12.0;  // small constant: prefer fewer, larger tiles on ties
  };
  if (p.epi == EPI_RU) return {p.N, 1};
  if (p.epi != EPI_GENERIC && p.epi != EPI_ACCUM) {
    static const int env_cg = [] { const char* e = std::getenv("ECHO_GEMM_CG"); return e ? atoi(e) : 0; }();  // tuning only
    if (c.cg == 1 || env_cg == 1 || p.N % 256 != 0 || p.M <= 128) return {256, 1};
    if (p.epi == EPI_QKV) {
      // 384-column pair tiles (two MMAs per K step, one TMEM stage) when they save whole waves: the 640-row plain step
      // is 96 tiles of 256 x 256 = two waves over the 74 CTA pairs, but 66 tiles of 256 x 384 = one (-25 % MMA time).
      // ECHO_GEMM_BN384=0 switches it off, =1 forces it.
      static const int env_384 = [] { const char* e = std::getenv("ECHO_GEMM_BN384"); return e ? atoi(e) : -1; }();
      const long pm = (p.M + 255) / 256, slots = sms / 2;
      const long w256 = (pm * p.batches * ((p.N + 255) / 256) + slots - 1) / slots;
      const long w384 = (pm * p.batches * ((p.N + 383) / 384) + slots - 1) / slots;
      if (c.bn == 384 || env_384 == 1 || (c.bn == 0 && env_384 != 0 && w384 * 384 < w256 * 256)) return {384, 2};
    }
    return {256, 2};
  }
  if (c.bn) return {c.bn, (c.cg == 2 && (c.bn == 256 || c.bn == 128)) ? 2 : 1};
  if (p.N % 64 != 0) {
    if (p.N % 96 == 0 && p.N <= 192) return {p.N, 1};  // 96 or 192
    return {0, 1};
  }
  if (p.N == 192 || p.N == 384) return {192, 1};
  TileCfg best = {0, 1};
  double bc = 1e30;
  const int cand[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    if (p.N % cand[i] != 0) continue;
    if (p.epi == EPI_ACCUM && cand[i] == 64) continue;  // the accumulate kernel is instantiated for 256 / 128 only
    for (int cg = 1; cg <= 2; ++cg) {
      if (cg == 2 && (!allow_pair || cand[i] == 64)) continue;
      const double cs = cost(cand[i], cg);
      if (cs < bc - 1e-9) { bc = cs; best = {cand[i], cg}; }
    }
  }
  return best;
}

cudaError_t gemm_launch(const GemmCall& c, cudaStream_t s) {
  GemmParams p = c.p;
  if (p.batches < 1) p.batches = 1;
  if (p.taps < 1) p.taps = 1;
  if (p.a_batch_div < 1) p.a_batch_div = 1;
  if (p.scale == 0.f) p.scale = 1.f;
  if (p.pos_period < 1) p.pos_period = 1;
  if (p.M <= 0 || p.N <= 0 || p.Kc <= 0 || p.taps > 8) return cudaErrorInvalidValue;
  if (p.epi == EPI_ACCUM) return cudaErrorInvalidValue;  // internal mode, chosen below
  if (p.epi == EPI_RU) {
    // fused ResidualUnit: square channel counts the kernel is instantiated for, shared weights, bf16 output to a buffer
    // other than A (tile t reads A rows that tile t - 1 has already replaced otherwise)
    if ((p.N != 96 && p.N != 192) || p.Kc != p.N || p.b_batch_rows != 0 || c.B1 == nullptr || (c.ldb1 % 8) != 0 ||
        (reinterpret_cast<uintptr_t>(c.B1) & 15) || p.out_bf16 == nullptr || (const void*)p.out_bf16 == (const void*)c.A ||
        p.gate != nullptr || p.n_valid != 0)
      return cudaErrorInvalidValue;
    p.act = ACT_SNAKE;
    p.col_mod = p.N;
  }
  // one gate row per rows_per_gate rows: a warp's 32-row slab must not straddle two of them (callers with other
  // group sizes launch once per group, see run_dit_layers)
  if (p.gate != nullptr && p.rows_per_gate > 0 &&
      (p.rows_per_gate % 32 != 0 || (p.batches > 1 && p.M % 32 != 0)))
    return cudaErrorInvalidValue;
  if (p.N % 32 != 0 || (c.lda % 8) != 0 || (c.ldb % 8) != 0) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(c.A) & 15) || (reinterpret_cast<uintptr_t>(c.B) & 15)) return cudaErrorInvalidValue;

  // Split-K: a single-wave GEMM whose tiles cover well under the 148 SMs (wo / w2 at M = 640: 40 tiles of 128 x 256)
  // is bound by the per-SM operand ingest of its few CTAs. When the epilogue is a pure accumulate into the fp32
  // residual stream (out = resid + gate * acc, resid aliasing out) the K blocks are split over several CTAs per
  // tile and the partial sums meet in the residual through fp32 vector atomics. The sum order then varies from run
  // to run in the last fp32 bit; after a few bf16 re-quantisations that decorrelates two runs down to the bf16
  // rounding-noise floor (same size as the error against the fp32 reference).
  // echo_set_deterministic(1) switches it off (bit-reproducible, ~4 % slower at batch 1).
  // L2 eviction hints: B is "streamed" when it is a large weight matrix read once per launch (every DiT / encoder
  // linear); the small, tile-after-tile reused conv weights of the DAC keep the default policy. ECHO_B_STREAM=0: off.
  {
    static const int env_bs = [] { const char* e = std::getenv("ECHO_B_STREAM"); return e ? atoi(e) : 1; }();
    const size_t b_bytes = (size_t)p.N * p.taps * p.Kc * 2 * (p.b_batch_rows ? p.batches : 1);
    p.b_stream = (env_bs != 0 && b_bytes >= ((size_t)4 << 20)) ? 1 : 0;
  }
  GemmCall cc = c;
  p.split_k = 1;
  {
    static const int env_sk = [] { const char* e = std::getenv("ECHO_SPLIT_K"); return e ? atoi(e) : 0; }();
    const bool eligible = p.epi == EPI_GENERIC && p.resid != nullptr && p.resid == p.out_f32 && p.out_bf16 == nullptr &&
                          p.taps == 1 && p.batches == 1 && p.N % 256 == 0;
    // partial planes: no atomics, fixed order -- used in deterministic mode only (in the default mode the atomics are
    // faster: 179.7 vs 181.5 ms per request)
    static const int env_planes = [] { const char* e = std::getenv("ECHO_SPLITK_PLANES"); return e ? atoi(e) : -1; }();
    const bool parts_ok = c.part_ws != nullptr && c.parts_used != nullptr &&
                          (env_planes >= 0 ? env_planes != 0 : g_deterministic.load() != 0);
    int want = c.split_k ? c.split_k : ((g_deterministic.load() && !parts_ok) ? 1 : env_sk);
    if (eligible && want == 0) {
      const int tiles256 = ((p.M + 127) / 128) * (p.N / 256);
      const int kb = (p.Kc + 63) / 64;
      if (tiles256 * 2 <= gemm_num_sms() && kb >= 16) {
        want = gemm_num_sms() / tiles256;
        if (want > 4) want = 4;
        while (want > 1 && kb / want < 8) --want;
      }
    }
    if (eligible && want > 1) {
      p.split_k = want;
      cc.bn = 256;
      cc.cg = 1;
    }
    // Even without split-K the accumulate epilogue adds gate * acc into the residual with fire-and-forget fp32 vector
    // reductions instead of load-add-store: no exposed read latency (wo 24.3 -> 22.7 us, w2 45.1 -> 42.8 us at
    // M = 1920). One contribution per element, so this stays bit-reproducible. ECHO_RED_EPILOGUE=0 restores the RMW.
    static const int env_red = [] { const char* e = std::getenv("ECHO_RED_EPILOGUE"); return e ? atoi(e) : 1; }();
    p.atomic_out = (eligible && env_red != 0) ? 1 : 0;
    // ... and runs in the lean EPI_ACCUM instantiation when nothing but gate / bias / scale is asked for
    static const int env_accum = [] { const char* e = std::getenv("ECHO_ACCUM_EPILOGUE"); return e ? atoi(e) : 1; }();
    if (p.atomic_out && env_accum != 0 && p.act == ACT_NONE && p.n_valid == 0 && (p.col_mod == 0 || p.col_mod == p.N) &&
        (c.bn == 0 || c.bn == 256 || c.bn == 128))
      p.epi = EPI_ACCUM;
    p.part_out = nullptr;
    if (c.parts_used) *c.parts_used = 0;
    if (p.epi == EPI_ACCUM && p.split_k > 1 && parts_ok && p.split_k <= 4) {
      p.part_out = c.part_ws;
      p.part_stride = c.part_stride;
      *c.parts_used = p.split_k;
    }
  }
  cc.p = p;
  const TileCfg tc = pick_cfg(cc);
  const int bn = tc.bn, cg = tc.cg;
  if (bn == 0) return cudaErrorInvalidValue;
  const bool small_k = (p.Kc % 64 != 0) && (p.Kc % 96 == 0) && bn == 96;  // DAC last stage: 96 channels / tap
  const int bk = small_k ? 32 : 64;
  const int a_batches = (p.batches + p.a_batch_div - 1) / p.a_batch_div;
  const int64_t a_bstride = c.a_batch_stride ? c.a_batch_stride : (int64_t)p.M * c.lda;

  // Fused ResidualUnit with 96 channels: the window form (EPI_RUW, gemm_tc.cuh) when the seven taps are evenly spaced and
  // their rows fit one 192-row window (dilations 1, 3, 9). ECHO_DAC_RU_WINDOW=0 keeps the ring form.
  bool ru_window = false;
  if (p.epi == EPI_RU && bn == 96 && p.taps == RUW_TAPS) {
    static const int env_w = [] { const char* e = std::getenv("ECHO_DAC_RU_WINDOW"); return e ? atoi(e) : 1; }();
    const int d = p.tap_shift[1] - p.tap_shift[0];
    ru_window = env_w != 0 && d >= 0 && GEMM_BM + (RUW_TAPS - 1) * d <= RUW_WIN_ROWS;
    for (int j = 1; j < RUW_TAPS; ++j) ru_window = ru_window && (p.tap_shift[j] - p.tap_shift[0] == j * d);
  }
  CUtensorMap ma, mb;
  if (!get_tensor_map(&ma, c.A, 3, (uint64_t)p.Kc, (uint64_t)(c.a_rows > 0 ? c.a_rows : p.M), (uint64_t)a_batches, (uint64_t)c.lda * 2,
                      (uint64_t)a_bstride * 2, bk, ru_window ? RUW_WIN_ROWS : GEMM_BM, bk * 2))
    return cudaErrorInvalidValue;
  const uint64_t b_rows = c.b_rows ? (uint64_t)c.b_rows : (uint64_t)p.N * (p.b_batch_rows ? p.batches : 1);
  // B box rows = this CTA's share of the tile's rows; the 384-column tile stages its 192 rows as three 64-row boxes
  if (!get_tensor_map(&mb, c.B, 2, (uint64_t)p.taps * p.Kc, b_rows, 1, (uint64_t)c.ldb * 2, 0, bk, bn == 384 ? 64 : bn / cg, bk * 2))
    return cudaErrorInvalidValue;

  if (p.epi == EPI_QKV && bn == 384) {
    // tile -> groups map (see GemmParams::tile_groups). Cost classes of a 128-column group in the epilogue: RoPE + norm (3),
    // norm (2), sigmoid (1), plain (0). The most expensive groups go one per tile into the slot both warp halves share; the
    // rest fill the per-half slots in pairs of equal class, so the two halves of a tile finish together.
    const int ngroups = p.N / 128;
    int ntile = (p.N + 383) / 384;
    if (b_batched_rows_check(p)) return cudaErrorInvalidValue;
    // single wave with CTA pairs to spare: more, narrower tiles -- x tiles of two groups (one MMA per K step, 1/3 less MMA
    // time) take the most expensive groups, so that MMA + epilogue time is even across the tiles
    {
      const long pm = (long)((p.M + 255) / 256) * p.batches, slots = gemm_num_sms() / 2;
      const int fit = (int)(slots / pm);
      if (fit > ntile && pm * ntile <= slots) ntile = std::min(fit, (ngroups + 1) / 2);
    }
    if (3 * ntile > (int)sizeof(p.tile_groups)) return cudaErrorInvalidValue;
    const int two = std::max(0, std::min(ntile, 3 * ntile - ngroups));  // tiles that hold two groups only
    std::vector<std::pair<int, int>> order;  // (-class, group)
    for (int g = 0; g < ngroups; ++g) {
      const int col = g * 128, si = col / p.sec_width;
      const QkvSection& sc = p.sec[si];
      const int grp = (col - si * p.sec_width) >> 7;
      const int cls = grp < sc.rope_heads ? 3 : sc.norm_w ? 2 : sc.sigmoid ? 1 : 0;
      order.push_back({-cls, g});
    }
    std::stable_sort(order.begin(), order.end());
    for (int i = 0; i < (int)sizeof(p.tile_groups); ++i) p.tile_groups[i] = 255;  // empty slot: beyond N
    p.qkv_tiles = ntile;
    int next = 0, last = ngroups - 1;
    // two-group tiles first: the heaviest groups, one per warp half
    for (int t = 0; t < two && next <= last; ++t) {
      p.tile_groups[3 * t] = (uint8_t)order[next++].second;
      if (next <= last) p.tile_groups[3 * t + 1] = (uint8_t)order[next++].second;
    }
    // three-group tiles: shared slot = the heaviest remaining groups, per-half slots = the rest, cheapest pairs first
    for (int t = two; t < ntile && next <= last; ++t) p.tile_groups[3 * t + 2] = (uint8_t)order[next++].second;
    for (int t = two; t < ntile && next <= last; ++t) {
      p.tile_groups[3 * t] = (uint8_t)order[last--].second;
      if (next <= last) p.tile_groups[3 * t + 1] = (uint8_t)order[last--].second;
    }
  }
  // DAC convs (bias, optional fp32 stream in / out, Snake -> bf16, nothing else): the lean epilogue -- same arithmetic in the
  // same order as the generic one (bit-identical), a third of its instructions. ECHO_CONV_EPILOGUE=0 keeps the generic path.
  if (p.epi == EPI_GENERIC) {
    static const int env_conv = [] { const char* e = std::getenv("ECHO_CONV_EPILOGUE"); return e ? atoi(e) : 1; }();
    const int cmod = p.col_mod > 0 ? p.col_mod : p.N;
    const bool lean = env_conv != 0 && p.act == ACT_SNAKE && p.out_bf16 != nullptr && p.alpha != nullptr && p.gate == nullptr &&
                      p.scale == 1.f && p.n_valid == 0 && p.split_k <= 1 && p.atomic_out == 0 && p.bias_bstride == 0 &&
                      cmod % 32 == 0 && p.N % bn == 0 && (p.resid == nullptr || p.out_f32 != nullptr);
    if (lean) {
      if (small_k) return launch_inst<96, 32, 3, EPI_CONV, 1>(ma, mb, p, s);
      if (bn == 192 && cg == 1) return launch_inst<192, 64, 1, EPI_CONV, 1>(ma, mb, p, s);
      if (bn == 256 && cg == 2) return launch_inst<256, 64, 1, EPI_CONV, 2>(ma, mb, p, s);
    }
  }
  if (p.epi == EPI_RU) {
    CUtensorMap mb1;
    if (!get_tensor_map(&mb1, c.B1, 2, (uint64_t)p.Kc, (uint64_t)p.N, 1, (uint64_t)c.ldb1 * 2, 0, bk, bn, bk * 2))
      return cudaErrorInvalidValue;
    if (ru_window) return launch_inst<96, 32, 3, EPI_RUW, 1>(ma, mb, p, s, &mb1);
    return bn == 96 ? launch_inst<96, 32, 3, EPI_RU, 1>(ma, mb, p, s, &mb1) : launch_inst<192, 64, 1, EPI_RU, 1>(ma, mb, p, s, &mb1);
  }
  switch (p.epi) {
    case EPI_ACCUM:
      if (bn == 256) return cg == 2 ? launch_inst<256, 64, 1, EPI_ACCUM, 2>(ma, mb, p, s) : launch_inst<256, 64, 1, EPI_ACCUM, 1>(ma, mb, p, s);
      if (bn == 128) return cg == 2 ? launch_inst<128, 64, 1, EPI_ACCUM, 2>(ma, mb, p, s) : launch_inst<128, 64, 1, EPI_ACCUM, 1>(ma, mb, p, s);
      return cudaErrorInvalidValue;
    case EPI_SWIGLU:
      if (p.N % 256 != 0) return cudaErrorInvalidValue;
      return cg == 2 ? launch_inst<256, 64, 1, EPI_SWIGLU, 2>(ma, mb, p, s) : launch_inst<256, 64, 1, EPI_SWIGLU, 1>(ma, mb, p, s);
    case EPI_QKV:
      if (p.sec_width % 128 != 0 || p.N % 128 != 0) return cudaErrorInvalidValue;
      if (bn == 384) return launch_inst<384, 64, 1, EPI_QKV, 2>(ma, mb, p, s);
      return cg == 2 ? launch_inst<256, 64, 1, EPI_QKV, 2>(ma, mb, p, s) : launch_inst<256, 64, 1, EPI_QKV, 1>(ma, mb, p, s);
    case EPI_GENERIC:
      if (small_k) return launch_inst<96, 32, 3, EPI_GENERIC, 1>(ma, mb, p, s);
      switch (bn) {
        case 256: return cg == 2 ? launch_inst<256, 64, 1, EPI_GENERIC, 2>(ma, mb, p, s) : launch_inst<256, 64, 1, EPI_GENERIC, 1>(ma, mb, p, s);
        case 192: return launch_inst<192, 64, 1, EPI_GENERIC, 1>(ma, mb, p, s);
        case 128: return cg == 2 ? launch_inst<128, 64, 1, EPI_GENERIC, 2>(ma, mb, p, s) : launch_inst<128, 64, 1, EPI_GENERIC, 1>(ma, mb, p, s);
        case 64: return launch_inst<64, 64, 1, EPI_GENERIC, 1>(ma, mb, p, s);
        default: return cudaErrorInvalidValue;
      }
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace echo
