// Library-level entry points: version, thread-local error text, device properties cache.
#include <stdarg.h>

#include "ar_common.cuh"

namespace ar {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};                  // per device: a process may drive a device other than 0
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cached[dev];
}

}  // namespace ar

extern "C" int ar_version(void) { return AR_B200_VERSION; }
extern "C" const char* ar_last_error(void) { return ar::g_err; }
