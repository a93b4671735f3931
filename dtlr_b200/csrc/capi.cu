// dtlr_b200 -- library identification and per-thread error string of the C ABI (include/dtlr_b200.h)
#include "common.cuh"

namespace dtlr {
static thread_local char g_err[512] = "";
int g_debug_flags = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dtlr

extern "C" int dtlr_version(void) { return (0 << 16) | (1 << 8) | 0; }
extern "C" int dtlr_built_for_sm(void) { return 100; }
extern "C" const char* dtlr_last_error(void) { return dtlr::g_err; }
extern "C" int dtlr_debug_flags(int flags) { const int old = dtlr::g_debug_flags; dtlr::g_debug_flags = flags; return old; }
