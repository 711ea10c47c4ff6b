// Error reporting, ABI version and device check of the C ABI (include/eonerf_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace eonerf {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace eonerf

extern "C" int eonerf_abi_version(void) { return EONERF_ABI_VERSION; }

extern "C" const char* eonerf_last_error(void) { return eonerf::g_err; }

extern "C" int eonerf_check_device(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    eonerf::set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return EONERF_EDEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    eonerf::set_error("device %d has compute capability %d.%d; this library is built for sm_100a only", dev, major, minor);
    return EONERF_EDEVICE;
  }
  return EONERF_OK;
}
