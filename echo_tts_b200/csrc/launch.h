// Kernel launch helper: every kernel of the per-step chain is launched with programmatic dependent launch (PDL) so
// that the next kernel's CTAs are scheduled -- and run their prologue (barrier init, TMEM alloc, tensor-map
// prefetch) -- while the previous kernel drains. The ~3.7 us launch gap measured between dependent kernels
// (profiles/r01_gemm_inkernel_timeline_v3.txt: 32.8 us per launch in a graph vs 29.1 us inside the kernel) is paid
// 7 200 times per request otherwise.
//
// Contract: a kernel launched through launch_k() MUST execute pdl_wait() before its first global-memory access
// (reads of the predecessor's outputs AND writes the predecessor might still read). pdl_wait() blocks until the
// whole predecessor grid has completed and its memory is visible, so completion stays transitive along the stream.
// pdl_trigger() only allows the dependent grid to be scheduled early; it has no memory semantics.
// ECHO_NO_PDL=1 launches everything fully serialised (A/B measurements).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <utility>

namespace echo {

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ECHO_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE (context) property of a kernel: a process that drives
// several GPUs (a threaded server instead of one process per GPU) must set it once on every device it launches on.
// `done` is a bit mask indexed by the current device; racing threads may both set the attribute, which is harmless.
template <typename K>
inline cudaError_t ensure_dyn_smem(std::atomic<uint64_t>& done, K kern, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  done.fetch_or(bit, std::memory_order_release);
  return cudaSuccess;
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

}  // namespace echo
