// Process-wide count of kernels launched by this library (reported through echo_num_launches / bench.py).
#pragma once
#include <atomic>
#include <cstdint>
namespace echo {
extern std::atomic<int64_t> g_launches;
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace echo
