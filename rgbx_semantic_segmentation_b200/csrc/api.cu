// Library-wide plumbing: error string, version, launch counter.
#include "common.cuh"
#include "../../include/cmx_b200.h"
#include <atomic>

std::atomic<long long> g_cmx_launches{0};
static thread_local char g_err[512] = "";

void cmx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

CMX_API const char* cmx_last_error(void) { return g_err; }
CMX_API int cmx_version(void) { return 100; }
CMX_API long long cmx_launch_count(void) { return g_cmx_launches.load(); }
