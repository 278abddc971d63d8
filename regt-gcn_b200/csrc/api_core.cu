// error string, version, launch counter (thread-local; the library has no other globals)
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace regt {
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
}  // namespace regt

extern "C" int regt_version(void) { return REGT_VERSION; }
extern "C" const char* regt_last_error(void) { return regt::g_err; }
extern "C" int64_t regt_launch_count(int reset) {
  long long v = regt::g_launches;
  if (reset) regt::g_launches = 0;
  return v;
}
