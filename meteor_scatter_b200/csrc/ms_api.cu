// Error reporting and version entry points of the C-ABI (include/ms_b200.h).
#include <stdarg.h>
#include <string.h>

#include "ms_common.cuh"

namespace ms {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ms

extern "C" {
int ms_abi_version(void) { return MS_ABI_VERSION; }
const char* ms_last_error(void) { return ms::g_err; }
}
