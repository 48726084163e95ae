// Error state, version and device queries of the C ABI (include/evoworld_b200.h).
#include "common.h"
#include <string.h>

namespace evw {
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
      return 148;
  }
  return cached;
}
}  // namespace evw

extern "C" const char* evw_last_error(void) { return evw::g_err; }
extern "C" int evw_abi_version(void) { return 2; }
extern "C" int evw_device_info(int* sm, int64_t* l2, int* cc) {
  int dev = 0, v = 0, maj = 0, min = 0;
  EVW_CUDA(cudaGetDevice(&dev));
  EVW_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  if (sm) *sm = v;
  EVW_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
  if (l2) *l2 = v;
  EVW_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  EVW_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (cc) *cc = maj * 10 + min;
  return EVW_OK;
}
