// ABI housekeeping: version, thread-local error text, device check.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace stinet {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace stinet

extern "C" {
long long stinet_launch_count(void) { return stinet::launches(); }
int stinet_abi_version(void) { return STINET_ABI_VERSION; }
const char* stinet_last_error(void) { return stinet::g_err; }
int stinet_device_ok(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    stinet::set_error("no CUDA device");
    return 0;
  }
  if (p.major != 10) {
    stinet::set_error("device sm_%d%d is not a Blackwell sm_100 part", p.major, p.minor);
    return 0;
  }
  return 1;
}
}
