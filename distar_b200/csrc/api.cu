// C-ABI plumbing: error string, version, launch counter.
#include <stdarg.h>
#include "common.cuh"

namespace dsb {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dsb

extern "C" const char* dsb_last_error(void) { return dsb::g_err; }
extern "C" int dsb_version(void) { return 1; }
extern "C" int64_t dsb_launch_count(void) { return dsb::g_launches.load(); }
