// What does one DEPENDENCY HOP cost on a B200? The per-step chain of the DiT is seven dependent kernels per block; each
// kernel's first useful load waits for the whole predecessor grid (griddepcontrol.wait). This measures the floor of such
// a hop with kernels that do nothing else -- every CTA reads one word its predecessor wrote and writes one word -- in
// three forms:
//   A  programmatic dependent launch, light CTAs (128 threads, no dynamic shared memory): the successor's CTAs are
//      resident and parked at the wait while the predecessor runs
//   B  programmatic dependent launch, full-SM CTAs (384 threads, 200 KB of dynamic shared memory, like the tcgen05 GEMM):
//      a successor CTA can only be placed on an SM once the predecessor's CTA there has exited
//   C  ONE persistent kernel (148 full-SM CTAs), the phases separated by a grid-wide barrier in global memory
//      (red.release.gpu + ld.acquire.gpu polling): what tile-level dependencies inside a persistent block kernel would pay
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hop_latency tools/hop_latency.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void hop(const int* in, int* out) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) out[blockIdx.x] = in[(blockIdx.x + 1) % gridDim.x] + 1;
}

__global__ void __launch_bounds__(384, 1) persistent(int* buf0, int* buf1, unsigned* counter, int phases) {
  const unsigned n = gridDim.x;
  for (int ph = 0; ph < phases; ++ph) {
    const int* in = (ph & 1) ? buf1 : buf0;
    int* out = (ph & 1) ? buf0 : buf1;
    if (threadIdx.x == 0) {
      out[blockIdx.x] = __ldcg(in + (blockIdx.x + 1) % n) + 1;
      // grid barrier: arrive (release), then poll until all CTAs of this phase have arrived
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      const unsigned want = n * (unsigned)(ph + 1);
      unsigned seen, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        if (++spins > (1u << 22)) __trap();  // not all CTAs co-resident: fail instead of hanging
      } while (seen < want);
    }
    __syncthreads();
  }
}

static float run_chain(int threads, size_t smem, int n, int* a, int* b, bool pdl) {
  CK(cudaFuncSetAttribute(hop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(a, 0, 148 * 4));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < n; ++i) CK(cudaLaunchKernelEx(&cfg, hop, (const int*)((i & 1) ? b : a), (i & 1) ? a : b));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  int h[148];
  CK(cudaMemcpy(h, (n & 1) ? b : a, sizeof(h), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 148; ++i)
    if (h[i] != n) { printf("chain result wrong: %d != %d\n", h[i], n); break; }
  return best * 1e3f / n;
}

int main() {
  int *a, *b;
  unsigned* counter;
  CK(cudaMalloc(&a, 148 * 4));
  CK(cudaMalloc(&b, 148 * 4));
  CK(cudaMalloc(&counter, 4));
  const int n = 2000;
  printf("A  PDL chain, light CTAs (128 threads, no smem):              %.2f us per hop\n", run_chain(128, 0, n, a, b, true));
  printf("B  PDL chain, full-SM CTAs (384 threads, 200 KB smem):        %.2f us per hop\n", run_chain(384, 200 * 1024, n, a, b, true));
  printf("   same without PDL (plain stream order), light CTAs:         %.2f us per hop\n", run_chain(128, 0, n, a, b, false));
  printf("   same without PDL (plain stream order), full-SM CTAs:       %.2f us per hop\n", run_chain(384, 200 * 1024, n, a, b, false));
  CK(cudaFuncSetAttribute(persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int phases : {2000, 20000}) {
    CK(cudaMemset(a, 0, 148 * 4));
    CK(cudaMemset(counter, 0, 4));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    persistent<<<148, 384, 200 * 1024>>>(a, b, counter, phases);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    int h[148];
    CK(cudaMemcpy(h, (phases & 1) ? b : a, sizeof(h), cudaMemcpyDeviceToHost));
    bool ok = true;
    for (int i = 0; i < 148; ++i) ok = ok && h[i] == phases;
    printf("C  persistent kernel, grid barrier per phase (%5d phases):    %.2f us per hop  (%s)\n", phases, ms * 1e3f / phases, ok ? "exact" : "WRONG");
  }
  return 0;
}
