// Library-level entry points: error string, ABI version.
#include <stdarg.h>

#include "common.cuh"

namespace ddf {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
}  // namespace ddf

extern "C" {
const char* ddf_last_error(void) { return ddf::get_error(); }
int ddf_abi_version(void) { return 1; }
int ddf_compiled_arch(void) { return 100; }
}
