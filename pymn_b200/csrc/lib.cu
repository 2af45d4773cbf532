// Library-level plumbing of the C ABI: error string, version, launch counter.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace mn {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mn

extern "C" {

const char* mn_last_error(void) { return mn::g_err; }

int mn_abi_version(void) { return 3; }

int64_t mn_launch_count(void) { return (int64_t)mn::g_launches.load(std::memory_order_relaxed); }

}
