#include "profiler.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "echo_b200.h"
#include "errors.h"

namespace echo {
namespace {
struct Rec { cudaEvent_t a, b; int cls; double flops, bytes; std::string tag; };
bool g_on = false;
std::vector<Rec> g_recs;
Rec g_cur;
}  // namespace

bool prof_enabled() { return g_on; }
void prof_begin(int cls, double flops, double bytes, cudaStream_t s, const char* tag) {
  g_cur.cls = cls; g_cur.flops = flops; g_cur.bytes = bytes; g_cur.tag = tag ? tag : "";
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
  struct Agg { long n = 0; double ms = 0, flops = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Agg& ag = agg[r.tag];
    ag.n += 1; ag.ms += ms; ag.flops += r.flops;
    out->launches[r.cls] += 1;
    out->ms[r.cls] += ms;
    out->flops[r.cls] += r.flops;
    out->bytes[r.cls] += r.bytes;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  if (getenv("ECHO_PROFILE_DUMP")) {
    std::vector<std::pair<std::string, Agg>> v(agg.begin(), agg.end());
    std::sort(v.begin(), v.end(), [](const std::pair<std::string, Agg>& x, const std::pair<std::string, Agg>& y) { return x.second.ms > y.second.ms; });
    double tot = 0;
    for (auto& kv : v) tot += kv.second.ms;
    fprintf(stderr, "[echo profile] total %.3f ms\n", tot);
    for (auto& kv : v)
      fprintf(stderr, "[echo profile] %9.3f ms %5.1f%% n=%5ld avg=%8.1f us %8.1f TFLOP/s  %s\n", kv.second.ms,
              100.0 * kv.second.ms / tot, kv.second.n, 1e3 * kv.second.ms / kv.second.n,
              kv.second.ms > 0 ? kv.second.flops / kv.second.ms / 1e9 : 0.0, kv.first.c_str());
  }
  return ECHO_OK;
}
