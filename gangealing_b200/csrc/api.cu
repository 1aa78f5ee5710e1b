// api.cu -- library-level entry points of libgg_b200 (version, error string, device query).
#include "common.cuh"

namespace gg {
int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}
}  // namespace gg

extern "C" {
int gg_version(void) { return 1; }
const char* gg_last_error(void) { return gg::err_buf(); }
int gg_sm_count(void) { return gg::sm_count(); }
}
