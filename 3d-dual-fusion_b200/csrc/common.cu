// Library-level entry points: error string, ABI version.
#include <stdlib.h>
#include <stdarg.h>

#include <atomic>

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

static std::atomic<long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int pdl_enabled() {
  static const int on = [] {
    const char* e = getenv("DDF_PDL");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on;
}
}  // namespace ddf

extern "C" {
const char* ddf_last_error(void) { return ddf::get_error(); }
int ddf_abi_version(void) { return 3; }
int ddf_compiled_arch(void) { return 100; }
int64_t ddf_launch_count(int reset) {
  long long v = ddf::g_launches.load();
  if (reset) ddf::g_launches.store(0);
  return v;
}
}
