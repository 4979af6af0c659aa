// api.cu -- library-level entry points: version, last error, launch counter.
#include "common.cuh"
#include <atomic>

namespace y2 {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace y2

extern "C" {
int y2_version(void) { return 100; }
const char* y2_last_error(void) { return y2::g_err; }
unsigned long long y2_launch_count(void) { return y2::g_launches.load(std::memory_order_relaxed); }
}
