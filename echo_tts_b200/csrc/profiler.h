// Optional per-launch CUDA-event timing (bench.py's roofline leg). Off by default: zero events are recorded on the
// normal path. When enabled, every launch wrapper brackets its kernel with two events on the launching stream.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace echo {
enum ProfClass : int { PROF_GEMM = 0, PROF_ATTN = 1, PROF_GLUE = 2, PROF_NCLASS = 3 };
bool prof_enabled();
void prof_begin(int cls, double flops, double bytes, cudaStream_t s, const char* tag);
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s;
  bool on;
  ProfScope(int cls, double flops, double bytes, cudaStream_t st, const char* tag = "") : s(st), on(prof_enabled()) {
    if (on) prof_begin(cls, flops, bytes, s, tag);
  }
  ~ProfScope() {
    if (on) prof_end(s);
  }
};
}  // namespace echo
