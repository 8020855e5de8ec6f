// Does TMA multicast lift the chip-wide L2 -> SM throughput cap (~6300 B/clk, the bound of the M = 640 GEMMs)?
// Every CTA pulls `iters` rounds of 4 x 32 KB tiles (256 rows x 64 bf16, 128B swizzle) from an L2-resident matrix into a
// 4-slot shared-memory ring, 16 tiles (512 KB) in flight per CTA and barrier phase (nobody reads the tiles, so slots are
// simply overwritten; this measures the delivery rate, not a consumer protocol). Modes: 0 = every CTA reads its own tiles; 1 = the CTAs of a cluster read the SAME tiles,
// unicast; 2 = the same tiles, each CTA loads 1 / cluster of the rows and multicasts them to the whole cluster.
// Reported: bytes DELIVERED into shared memory per second, chip-wide, and per SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mc_bandwidth tools/mc_bandwidth.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int ROWS = 256, COLS = 64, TILE_BYTES = ROWS * COLS * 2, SLOTS = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
  asm volatile(
      "{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(smem_u32(b)), "r"(phase)
      : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.aligned; barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

template <int CS>
__global__ void __launch_bounds__(128, 1) pull(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_part,
                                               int mode, int iters, int tiles_total, long long* cycles, int share) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SLOTS * TILE_BYTES);
  const uint32_t rank = CS > 1 ? cluster_rank() : 0;
  const int cluster_id = blockIdx.x / CS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CS > 1) cluster_sync();
  const long long t0 = clock64();
  // Nobody reads the tiles, so slots may be overwritten freely; only the barrier accounting has to be right: one barrier
  // phase per BATCH of 16 tiles (512 KB in flight per CTA, issued back to back), one cluster barrier per batch.
  constexpr int BATCH = 16;
  for (int it = 0; it < iters * SLOTS / BATCH; ++it) {
    if (threadIdx.x == 0) {
      mbar_expect(&full[0], BATCH * TILE_BYTES);
      for (int k = 0; k < BATCH; ++k) {
        const int n = it * BATCH + k, s = k % SLOTS;
        if (mode == 2) {
          // this CTA's 1 / CS of the rows of the cluster's tile, delivered to every CTA of the cluster
          const int tile = (cluster_id * 977 + n) % tiles_total;
          constexpr int PR = ROWS / (CS == 6 ? 8 : CS);  // (mode 2 is not run for 6)
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
              ::"r"(smem_u32(smem + s * TILE_BYTES + rank * PR * COLS * 2)), "l"(&tm_part), "r"(smem_u32(&full[0])), "r"(0),
                "r"(tile * ROWS + (int)rank * PR), "h"((uint16_t)((1u << CS) - 1))
              : "memory");
        } else {
          const int tile = mode == 1 ? ((blockIdx.x / share) * 977 + n) % tiles_total : (blockIdx.x * 977 + n) % tiles_total;
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(smem_u32(smem + s * TILE_BYTES)), "l"(&tm_full), "r"(smem_u32(&full[0])), "r"(0), "r"(tile * ROWS)
              : "memory");
        }
      }
      mbar_wait(&full[0], it & 1);
    }
    __syncthreads();
    if (CS > 1) cluster_sync();  // a CTA's next batch must not signal a peer's barrier before the peer has re-armed it
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void* ptr, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {COLS, rows};
  cuuint64_t strides[1] = {COLS * 2};
  cuuint32_t box[2] = {COLS, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int CS>
static void run(EncodeFn enc, void* buf, int tiles_total, int mode, int iters, long long* d_cycles, int sms, int share = 0) {
  if (share == 0) share = CS;
  const CUtensorMap full = make_map(enc, buf, (uint64_t)tiles_total * ROWS, ROWS);
  const CUtensorMap part = make_map(enc, buf, (uint64_t)tiles_total * ROWS, ROWS / (CS == 6 ? 8 : CS));
  const int grid = sms / CS * CS;
  const int smem = SLOTS * TILE_BYTES + 1024 + 128;
  CK(cudaFuncSetAttribute(pull<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (CS > 8) CK(cudaFuncSetAttribute(pull<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    cudaError_t le = cudaLaunchKernelEx(&cfg, pull<CS>, full, part, mode, iters, tiles_total, d_cycles, share);
    if (le != cudaSuccess) { printf("cluster %d mode %d: launch failed: %s\n", CS, mode, cudaGetErrorString(le)); cudaGetLastError(); return; }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (long long c : cyc) mean += (double)c / grid;
  const double per_cta = (double)iters * SLOTS * TILE_BYTES;
  printf("cluster %d share %d mode %d (%s): %3d CTAs, %.1f us, delivered %.2f TB/s chip-wide, %.1f B/clk/SM, %.0f B/clk chip (SM clocks)\n", CS, share, mode,
         mode == 0 ? "distinct tiles" : mode == 1 ? "same tiles, unicast" : "same tiles, multicast", grid, ms * 1e3,
         per_cta * grid / (ms * 1e-3) / 1e12, per_cta / mean, per_cta / mean * grid);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
  const int tiles_total = 1536;  // 48 MB: L2-resident
  void* buf;
  CK(cudaMalloc(&buf, (size_t)tiles_total * TILE_BYTES));
  CK(cudaMemset(buf, 1, (size_t)tiles_total * TILE_BYTES));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, 1024 * sizeof(long long)));
  const int iters = 200;
  for (int mode = 0; mode < 3; ++mode) {
    if (mode < 2) run<1>(enc, buf, tiles_total, mode, iters, d_cycles, sms);
    run<2>(enc, buf, tiles_total, mode, iters, d_cycles, sms);
    run<4>(enc, buf, tiles_total, mode, iters, d_cycles, sms);
    run<8>(enc, buf, tiles_total, mode, iters, d_cycles, sms);
  }
  // the same tiles read (unicast) by `share` consecutive CTAs that are NOT all in one cluster: does the merging of identical
  // requests need the cluster?
  run<1>(enc, buf, tiles_total, 1, iters, d_cycles, sms, 2);
  run<1>(enc, buf, tiles_total, 1, iters, d_cycles, sms, 4);
  run<1>(enc, buf, tiles_total, 1, iters, d_cycles, sms, 6);
  run<2>(enc, buf, tiles_total, 1, iters, d_cycles, sms, 4);
  run<2>(enc, buf, tiles_total, 1, iters, d_cycles, sms, 6);
  run<6>(enc, buf, tiles_total, 1, iters, d_cycles, sms, 6);
  run<6>(enc, buf, tiles_total, 0, iters, d_cycles, sms, 6);
  return 0;
}
