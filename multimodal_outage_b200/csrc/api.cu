// Error reporting, version and device gate of libgwn.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace gwn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GWN_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
long long* trace_ptr(const char* env_name) {
#ifdef GWN_TRACE
  const char* e = getenv(env_name);
  return e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
#else
  (void)env_name;
  return nullptr;
#endif
}
}  // namespace gwn

extern "C" long long gwn_launch_count(void) { return gwn::g_launches.load(); }
extern "C" const char* gwn_last_error(void) { return gwn::g_err; }
extern "C" int gwn_version(void) { return 100; }

extern "C" int gwn_check_device(void) {
  int dev = 0;
  GWN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  GWN_CUDA(cudaGetDeviceProperties(&prop, dev));
  GWN_REQUIRE(prop.major == 10 && prop.minor == 0,
              "libgwn is built for sm_100a (B200) only; device %d is sm_%d%d (%s) - no fallback path exists", dev,
              prop.major, prop.minor, prop.name);
  return 0;
}
