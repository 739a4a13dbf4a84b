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
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

}  // namespace ar

extern "C" int ar_version(void) { return AR_B200_VERSION; }
extern "C" const char* ar_last_error(void) { return ar::g_err; }
