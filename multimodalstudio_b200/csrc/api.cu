// Error reporting, version and launch accounting for the C ABI (include/mms_b200.h).
#include "common.cuh"

#include <atomic>

namespace mmsb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mmsb

extern "C" const char* mmsb_version(void) { return "mms_b200 0.1.0 (sm_100a)"; }
extern "C" const char* mmsb_last_error(void) { return mmsb::g_err; }
extern "C" int64_t mmsb_launch_count(void) { return mmsb::g_launches.load(std::memory_order_relaxed); }
