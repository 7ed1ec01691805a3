// Error plumbing and version entry points of the C-ABI (include/tdeed_b200.h).
#include "common.cuh"
#include <cstdarg>

namespace tdeed {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tdeed

extern "C" int tdeed_abi_version(void) { return 1; }
extern "C" const char* tdeed_last_error(void) { return tdeed::g_err; }
