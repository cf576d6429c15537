// libtod.so: error plumbing and identification entry points (include/tod.h).
#include <cstdarg>
#include <cstdio>

#include <cstdlib>

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

// ---- timeline (tools only)
static unsigned long long* g_timeline = nullptr;
static int g_timeline_seq = 0;
static char g_timeline_names[4096][48];

TimelineTag timeline_tag(const char* name) {
  TimelineTag t{g_timeline, 0};
  if (g_timeline != nullptr) {
    t.id = g_timeline_seq < 4096 ? g_timeline_seq : 4095;
    snprintf(g_timeline_names[t.id], sizeof(g_timeline_names[0]), "%s", name);
    ++g_timeline_seq;
  }
  return t;
}

// ---- SM budget of the persistent kernels' grids (tma_host.cuh: num_sms)
static int g_sm_budget = -1;
int sm_budget_value() {
  if (g_sm_budget < 0) {
    const char* e = getenv("TOD_SM_BUDGET");
    g_sm_budget = (e != nullptr) ? atoi(e) : 0;
    if (g_sm_budget < 0) g_sm_budget = 0;
  }
  return g_sm_budget;
}

}  // namespace tod

extern "C" int tod_set_sm_budget(int32_t sms) {
  TOD_CHECK_ARG(sms >= 0, "tod_set_sm_budget: negative budget");
  tod::g_sm_budget = sms;
  return TOD_OK;
}
extern "C" int tod_get_sm_budget(void) { return tod::sm_budget_value(); }

extern "C" int tod_debug_set_timeline(void* d_buf) {
  tod::g_timeline = reinterpret_cast<unsigned long long*>(d_buf);
  if (d_buf == nullptr) tod::g_timeline_seq = 0;
  return TOD_OK;
}
extern "C" int tod_debug_timeline_launches(void) { return tod::g_timeline_seq; }
extern "C" const char* tod_debug_timeline_name(int id) {
  return (id >= 0 && id < tod::g_timeline_seq && id < 4096) ? tod::g_timeline_names[id] : "";
}

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
