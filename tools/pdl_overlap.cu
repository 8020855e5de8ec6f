// Do the CTAs of a PDL-launched secondary kernel start on SMs the primary has already left, while other CTAs of the
// primary are still running? Primary: 148 full-SM CTAs (384 threads, 200 KB smem); all trigger their dependents at once,
// CTAs >= `busy` exit, the first `busy` CTAs spin for `spin_us`. Secondary: 240 CTAs (light, or 99 KB smem = 2 per SM),
// no griddepcontrol.wait; every CTA stamps %globaltimer at entry. Reported: when the secondary's CTAs started relative to
// the END of the primary's last CTA (negative = overlapped).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pdl_overlap tools/pdl_overlap.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int COLS>
__device__ __forceinline__ uint32_t tmem_alloc_warp(uint32_t* slot) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  __syncwarp();
  return *reinterpret_cast<volatile uint32_t*>(slot);
}
template <int COLS>
__device__ __forceinline__ void tmem_free_warp(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(COLS) : "memory");
}

__global__ void __launch_bounds__(384, 1) primary(unsigned long long* end, int busy, int spin_ns, int trigger_early, int tmem) {
  __shared__ uint32_t slot;
  uint32_t taddr = 0;
  if (tmem && threadIdx.x < 32) taddr = tmem_alloc_warp<512>(&slot);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    // the first `busy` CTAs run spin_ns, the others half of it: SMs become free in the MIDDLE of the primary's run
    const unsigned long long t0 = gtime(), dur = (int)blockIdx.x < busy ? (unsigned long long)spin_ns : (unsigned long long)spin_ns / 2;
    while (gtime() - t0 < dur) { }
  }
  if (threadIdx.x == 0) end[blockIdx.x] = gtime();
  if (tmem && threadIdx.x < 32) tmem_free_warp<512>(taddr);
}

__global__ void secondary(unsigned long long* start, int wait, int tmem) {
  __shared__ uint32_t slot;
  uint32_t taddr = 0;
  if (tmem && threadIdx.x < 32) taddr = tmem_alloc_warp<256>(&slot);
  if (wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) start[blockIdx.x] = gtime();
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < 3000ull) { }  // 3 us of "work"
  if (tmem && threadIdx.x < 32) tmem_free_warp<256>(taddr);
}

int main() {
  unsigned long long *end, *start;
  CK(cudaMalloc(&end, 148 * 8));
  CK(cudaMalloc(&start, 240 * 8));
  CK(cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(secondary, cudaFuncAttributeMaxDynamicSharedMemorySize, 99 * 1024));
  for (int variant = 0; variant < 10; ++variant) {
    const int sec_smem = (variant & 1) ? 99 * 1024 : 0;
    const int wait = (variant == 4 || variant == 5) ? 1 : 0;
    const int cluster = (variant == 6 || variant == 7 || variant == 9) ? 2 : 1;
    const int tmem = variant >= 8 ? 1 : 0;
    const int trigger_early = (variant == 2 || variant == 3) ? 0 : 1;
    std::vector<double> first, median, last;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaMemset(end, 0, 148 * 8));
      CK(cudaMemset(start, 0, 240 * 8));
      CK(cudaDeviceSynchronize());
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      at[1].id = cudaLaunchAttributeClusterDimension;
      at[1].val.clusterDim.x = cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
      cudaLaunchConfig_t c1 = {};
      c1.gridDim = dim3(148); c1.blockDim = dim3(384); c1.dynamicSmemBytes = 200 * 1024; c1.attrs = at; c1.numAttrs = cluster > 1 ? 2 : 1;
      cudaLaunchConfig_t c2 = {};
      c2.gridDim = dim3(240); c2.blockDim = dim3(192); c2.dynamicSmemBytes = sec_smem; c2.attrs = at; c2.numAttrs = 1;
      CK(cudaLaunchKernelEx(&c1, primary, end, 68, 20000, trigger_early, tmem));
      CK(cudaLaunchKernelEx(&c2, secondary, start, wait, tmem));
      CK(cudaDeviceSynchronize());
      unsigned long long he[148], hs[240];
      CK(cudaMemcpy(he, end, sizeof(he), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hs, start, sizeof(hs), cudaMemcpyDeviceToHost));
      const unsigned long long pend = *std::max_element(he, he + 148);
      std::sort(hs, hs + 240);
      first.push_back(((double)hs[0] - (double)pend) / 1e3);
      median.push_back(((double)hs[120] - (double)pend) / 1e3);
      last.push_back(((double)hs[239] - (double)pend) / 1e3);
    }
    std::sort(first.begin(), first.end()); std::sort(median.begin(), median.end()); std::sort(last.begin(), last.end());
    printf("[cluster %d, tmem %d] secondary %s, %s, primary triggers %s: first / median / last CTA start %+7.2f / %+7.2f / %+7.2f us after the primary's last CTA ended\n",
           cluster, tmem, sec_smem ? "99 KB smem (2 per SM)" : "no smem             ", wait ? "griddepcontrol.wait  " : "no wait              ",
           trigger_early ? "at its start" : "never (implicit at exit)", first[2], median[2], last[2]);
  }
  return 0;
}
