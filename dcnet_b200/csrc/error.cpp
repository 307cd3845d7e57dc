#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/dcnet_b200.h"
#include "host_error.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

extern "C" int dcnet_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" void dcnet_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" const char* dcnet_last_error(void) { return g_err; }
extern "C" int dcnet_abi_version(void) { return DCNET_ABI_VERSION; }
extern "C" long long dcnet_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
