// api.cu — error string, version, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace sml {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace sml

extern "C" const char* sml_last_error(void) { return sml::g_err; }
extern "C" int sml_version(void) { return 100; }
extern "C" int64_t sml_launch_count(void) { return (int64_t)sml::g_launches.load(); }
