// libtod.so: error plumbing and identification entry points (include/tod.h).
#include <cstdarg>
#include <cstdio>

#include "tod_common.cuh"

namespace tod {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TOD_OK;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return TOD_ERR_CUDA;
}

}  // namespace tod

extern "C" int tod_version(void) { return 100; }

extern "C" const char* tod_last_error(void) { return tod::g_err; }

extern "C" int tod_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return tod::check_cuda(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return tod::check_cuda(e, "cudaDeviceGetAttribute");
  return major == 10 ? 1 : 0;
}
