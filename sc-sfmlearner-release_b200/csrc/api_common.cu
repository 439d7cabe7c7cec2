// Error reporting and version of libscsfm (C ABI: include/scsfm.h).
#include <stdarg.h>

#include "common.cuh"

namespace scsfm {
static thread_local char g_err[512] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace scsfm

extern "C" const char* scsfm_last_error(void) { return scsfm::g_err; }
extern "C" int scsfm_version(void) { return 100; }
extern "C" long long scsfm_launch_count(void) { return __atomic_load_n(&scsfm::g_launch_count, __ATOMIC_RELAXED); }
