// Error string, device queries and version of the lwpose_b200 C-ABI library.
#include "common.cuh"

#include <string.h>

namespace lwp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int sms[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  int &s = sms[dev % kMaxDevices];
  if (s == 0) {
    if (cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || s <= 0) s = 148;
  }
  return s;
}

// SMs the NETWORK kernels size their persistent grids for: all of them, minus LWP_NET_SM_RESERVE (experiment: SMs left
// free so that the post-processing of the previous batch, which runs on its own stream, does not have to wait for --
// and then delay -- a persistent network CTA).
int net_sms() {
  static int reserve = -1;
  if (reserve < 0) {
    const char *e = getenv("LWP_NET_SM_RESERVE");
    reserve = e ? atoi(e) : 0;
    if (reserve < 0 || reserve > 64) reserve = 0;
    reserve &= ~1;   // CTA pairs: keep the count even
  }
  return num_sms() - reserve;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) on = (getenv("LWP_NO_PDL") != nullptr && atoi(getenv("LWP_NO_PDL")) != 0) ? 0 : 1;
  return on != 0;
}

}  // namespace lwp

extern "C" int lwp_version(void) { return 200; }

extern "C" int lwp_timing_experiments(void) {
#ifdef LWP_TIMING_EXPERIMENTS
  return 1;
#else
  return 0;
#endif
}

extern "C" const char *lwp_last_error(void) { return lwp::g_err; }

extern "C" int lwp_check_device(int dev) {
  int major = 0;
  LWP_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    lwp::set_error("device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
    return LWP_EARCH;
  }
  return LWP_OK;
}
