#include "profiler.h"

#include <vector>

#include "echo_b200.h"
#include "errors.h"

namespace echo {
namespace {
struct Rec { cudaEvent_t a, b; int cls; double flops, bytes; };
bool g_on = false;
std::vector<Rec> g_recs;
Rec g_cur;
}  // namespace

bool prof_enabled() { return g_on; }
void prof_begin(int cls, double flops, double bytes, cudaStream_t s) {
  g_cur.cls = cls; g_cur.flops = flops; g_cur.bytes = bytes;
  cudaEventCreate(&g_cur.a);
  cudaEventCreate(&g_cur.b);
  cudaEventRecord(g_cur.a, s);
}
void prof_end(cudaStream_t s) {
  cudaEventRecord(g_cur.b, s);
  g_recs.push_back(g_cur);
}
}  // namespace echo

using namespace echo;

extern "C" int echo_profile_start(echo_handle*) {
  for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_recs.clear();
  g_on = true;
  return ECHO_OK;
}

extern "C" int echo_profile_stop(echo_handle*, echo_profile_report* out) {
  g_on = false;
  if (!out) { set_error("echo_profile_stop: null report"); return ECHO_ERR_ARG; }
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("echo_profile_stop: sync failed"); return ECHO_ERR_CUDA; }
  for (int c = 0; c < 3; ++c) { out->launches[c] = 0; out->ms[c] = 0; out->flops[c] = 0; out->bytes[c] = 0; }
  for (auto& r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    out->launches[r.cls] += 1;
    out->ms[r.cls] += ms;
    out->flops[r.cls] += r.flops;
    out->bytes[r.cls] += r.bytes;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  return ECHO_OK;
}
