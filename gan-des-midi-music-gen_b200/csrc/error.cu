#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void mmg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* mmg_last_error(void) { return g_err; }
extern "C" int mmg_abi_version(void) { return 1; }

unsigned long long g_mmg_launches = 0;
// number of CUDA kernels this library has launched in this process (bench.py's gpu_launches)
extern "C" uint64_t mmg_launch_count(void) { return g_mmg_launches; }
