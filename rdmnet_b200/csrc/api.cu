// Error reporting + version for librdm_sm100.so.
#include <stdarg.h>
#include "common.cuh"
#include "../../include/rdm_sm100.h"

static thread_local char g_err[1024] = "";

void rdm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* rdm_last_error(void) { return g_err; }
extern "C" int rdm_version(void) { return 101; }

unsigned long long g_rdm_launches = 0;
extern "C" unsigned long long rdm_launch_count(void) { return __atomic_load_n(&g_rdm_launches, __ATOMIC_RELAXED); }
