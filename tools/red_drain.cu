// How long does the residual-accumulate epilogue's output take to DRAIN? 120 CTAs x 4 warps each push one 128 x 256
// fp32 tile of the residual stream X [M, 2048] from a swizzled shared-memory patch (the layout of the GEMM's accumulate
// epilogue: one 32-row x 32-column patch per warp and chunk) to global memory, and the launch-to-launch period of a
// back-to-back stream of such kernels is timed (it contains the drain: the next launch starts when the writes of the
// previous one are visible). Modes:
//   0  st.global.v4.f32                       (plain stores: the floor)
//   1  red.global.add.v4.f32                  (what the epilogue does today)
//   2  cp.reduce.async.bulk.tensor.2d ...add  (TMA reduce-add of the same patch, one elected lane per warp and chunk)
//   3  cp.async.bulk.tensor.2d store          (TMA store: floor of the async path)
//   4  ld.global.v4 + add + st.global.v4      (read-modify-write by the owning CTA; only valid for split = 1)
// split = 3: three CTAs add into the same tile (M = 640 with 3 K-slices) instead of 120 distinct tiles (M = 1920).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/red_drain tools/red_drain.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int N = 2048, BM = 128, BN = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) push(const __grid_constant__ CUtensorMap tm, float* X, int M, int split, int mode, float val, int rounds) {
  extern __shared__ uint8_t raw[];
  if (mode == 5) return;  // empty kernel: the launch period itself
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM;
  const int tile = blockIdx.x / split;
  const int mt = tile % tiles_m, nt = tile / tiles_m;
  const int row0 = mt * BM + warp * 32;
  uint8_t* stg0 = smem + warp * 8192;  // two 4 KB patches per warp
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = val;
  const int sub = lane >> 3, c4 = lane & 7;
#pragma unroll 1
  for (int cc = 0; cc < rounds * (BN / 32); ++cc) {
    const int ch = cc % (BN / 32);
    uint8_t* stg = stg0 + (ch & 1) * 4096;
    const int c0 = nt * BN + ch * 32;
    if (mode == 2 || mode == 3) {
      if (cc >= 2) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t a = smem_u32(stg + lane * 128 + ((j ^ (lane & 7)) << 4));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
    }
    if (mode == 2 || mode == 3) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (mode == 2)
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tm), "r"(c0), "r"(row0), "r"(smem_u32(stg)) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tm), "r"(c0), "r"(row0), "r"(smem_u32(stg)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      __syncwarp();
      float* dst = X + (size_t)(row0 + sub) * N + c0 + 4 * c4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = sub + 4 * i;
        const uint32_t a = smem_u32(stg + r * 128 + ((c4 ^ (r & 7)) << 4));
        float4 t;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(a));
        if (row0 + r < M) {
          if (mode == 0) {
            *reinterpret_cast<float4*>(dst) = t;
          } else if (mode == 1) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
          } else {
            float4 o = *reinterpret_cast<const float4*>(dst);
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            *reinterpret_cast<float4*>(dst) = o;
          }
        }
        dst += (size_t)4 * N;
      }
      __syncwarp();
    }
  }
  if (mode == 2 || mode == 3) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

int main() {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int smem_bytes = 4 * 8192 + 1024;
  CK(cudaFuncSetAttribute(push, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int M = cfg == 0 ? 1920 : 640, split = cfg == 0 ? 1 : 3;
    float* X;
    CK(cudaMalloc(&X, (size_t)M * N * 4));
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int rounds = 1; rounds <= 16; rounds *= 4)
    for (int mode = 0; mode < 6; ++mode) {
      if (mode == 4 && split > 1) continue;
      CK(cudaMemset(X, 0, (size_t)M * N * 4));
      const int reps = 200;
      for (int i = 0; i < 20; ++i) push<<<120, 128, smem_bytes>>>(tm, X, M, split, mode, 0.f, rounds);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int i = 0; i < reps; ++i) push<<<120, 128, smem_bytes>>>(tm, X, M, split, mode, 1.f, rounds);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<float> h((size_t)M * N);
      CK(cudaMemcpy(h.data(), X, h.size() * 4, cudaMemcpyDeviceToHost));
      const float want = mode == 5 ? 0.f : (mode == 0 || mode == 3) ? 1.f : (float)(reps * split * rounds);
      size_t bad = 0;
      for (float f : h) bad += f != want;
      const char* names[6] = {"st.global.v4", "red.global.add.v4.f32", "TMA reduce-add (32x32 boxes)", "TMA store (32x32 boxes)", "ld + add + st", "empty kernel"};
      printf("M=%4d split=%d rounds=%2d  %-30s %7.2f us per launch   (%zu wrong of %zu)\n", M, split, rounds, names[mode], ms * 1e3 / reps, bad, h.size());
    }
    CK(cudaFree(X));
  }
  return 0;
}
