// Library plumbing: version, thread-local error text, launch counter.
#include "common.cuh"

namespace dcb {

unsigned long long g_launches = 0;

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
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

}  // namespace dcb

extern "C" int dcb_version(void) { return 100; }
extern "C" const char* dcb_last_error(void) { return dcb::last_error_buf(); }
extern "C" unsigned long long dcb_launch_count(void) { return dcb::g_launches; }
