// Library-wide state of the C ABI: version, thread-local error text, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

namespace ssl {
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int g_predict_tiled = 1;
int g_kmeans_rows_per_round = 4;
}  // namespace ssl

extern "C" int ssl_version(void) { return 100; }
extern "C" const char *ssl_last_error(void) { return g_err; }
extern "C" int64_t ssl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
