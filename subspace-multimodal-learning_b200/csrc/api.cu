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

// Dropout salt.  Every dropout mask in the library is a counter-based hash of (seed, offset / pair index) passed BY VALUE, which
// a captured CUDA graph would freeze.  When a salt pointer is installed, each mask kernel xors its seed with the 64-bit value at
// that device address at run time; sml_seed_salt_advance (a one-thread kernel, captured as the first node of a step graph) moves
// it to the next value of a splitmix64 sequence, so every replay draws fresh masks while the forward and backward kernels of one
// step still agree.  NULL (the default) leaves the seeds as passed.
namespace sml {
static const unsigned long long* g_salt = nullptr;
const unsigned long long* seed_salt() { return g_salt; }
__global__ void salt_advance_kernel(unsigned long long* p) {
  unsigned long long z = (*p += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  p[1] = z ^ (z >> 31);
}
}  // namespace sml
/* salt: device array of two uint64 {counter, current salt}; kernels read salt[1].  NULL uninstalls. */
extern "C" int sml_set_seed_salt(const void* salt2) {
  sml::g_salt = salt2 ? reinterpret_cast<const unsigned long long*>(salt2) + 1 : nullptr;
  return 0;
}
extern "C" int sml_seed_salt_advance(void* salt2, void* stream) {
  sml::salt_advance_kernel<<<1, 1, 0, sml::as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(salt2));
  SML_LAUNCH_OK("sml_seed_salt_advance");
  return 0;
}

extern "C" const char* sml_last_error(void) { return sml::g_err; }
extern "C" int sml_version(void) { return 100; }
extern "C" int64_t sml_launch_count(void) { return (int64_t)sml::g_launches.load(); }
