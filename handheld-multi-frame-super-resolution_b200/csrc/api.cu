// Version and error reporting of libhhsr.so (no device state is kept anywhere in the library).
#include "common.cuh"

namespace hhsr {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace hhsr

extern "C" int hhsr_version(void) { return HHSR_VERSION; }
extern "C" const char *hhsr_last_error_string(void) { return hhsr::g_err; }
