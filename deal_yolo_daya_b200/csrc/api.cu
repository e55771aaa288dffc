// Version + thread-local error reporting of libdyd.so.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace dyd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Kernel launches issued by this library since it was loaded (a statistic, read by bench.py for `gpu_launches`).
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

}  // namespace dyd

extern "C" int dyd_version(void) { return DYD_VERSION; }

extern "C" uint64_t dyd_launch_count(void) { return dyd::g_launches.load(std::memory_order_relaxed); }

extern "C" size_t dyd_last_error(char* buf, size_t cap) {
    size_t n = strlen(dyd::g_err);
    if (buf && cap) {
        size_t m = n < cap - 1 ? n : cap - 1;
        memcpy(buf, dyd::g_err, m);
        buf[m] = 0;
    }
    return n;
}
