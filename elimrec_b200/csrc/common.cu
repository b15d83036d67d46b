#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void elimrec_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

ELIMREC_API const char* elimrec_last_error(void) { return g_err; }
ELIMREC_API int elimrec_abi_version(void) { return 2; }
