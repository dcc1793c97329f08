// Shared host/device helpers for the lwpose_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <utility>

#include "../../include/lwpose_b200.h"

namespace lwp {

void set_error(const char *fmt, ...);

#define LWP_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      lwp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return LWP_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

#define LWP_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      lwp::set_error(__VA_ARGS__);    \
      return LWP_EINVAL;              \
    }                                 \
  } while (0)

// launch error check that never synchronises
#define LWP_LAUNCH_CHECK() LWP_CUDA_CHECK(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int num_sms();   // SM count of the CURRENT device (cached per device ordinal)
int net_sms();   // SMs the network kernels' persistent grids are sized for (num_sms() - LWP_NET_SM_RESERVE)

// One-time set-up that is per DEVICE (cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is): a process that drives
// several GPUs must repeat it on each of them.
constexpr int kMaxDevices = 64;
struct DeviceOnce {
  bool done[kMaxDevices] = {};
  // true while the current device still needs the set-up; *slot = index to mark in `done` afterwards
  bool pending(int *slot) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
    *slot = d % kMaxDevices;
    return !done[*slot];
  }
};

// Timing-experiment switches (LWP_DEBUG_GEMM / _DW / _HEADS / _DWPW skip loads, MMAs or epilogues: results are wrong
// by design).  They only exist in a library built with -DLWP_TIMING_EXPERIMENTS; the release build compiles the
// branches out (LWP_DBG(x) is the constant 0) and lwp_timing_experiments() reports 0.
#ifdef LWP_TIMING_EXPERIMENTS
#define LWP_DBG(x) (x)
static inline int debug_env(const char *name) { const char *e = getenv(name); return e ? atoi(e) : 0; }
#else
#define LWP_DBG(x) 0
static inline int debug_env(const char *) { return 0; }
#endif

// Programmatic dependent launch: a kernel launched through launch_pdl may begin (prologue: barrier init, TMEM
// allocation, tensor-map prefetch, constant loads) while the previous kernel of the stream is still draining; it must
// execute pdl_wait() before it reads anything an earlier kernel wrote or writes anything an earlier kernel may still
// read.  pdl_trigger() lets the NEXT kernel of the stream be scheduled as soon as SM resources free up.
// LWP_NO_PDL=1 launches everything as ordinary stream-ordered kernels (then both device calls are no-ops).
bool pdl_enabled();
#ifdef __CUDACC__
// ELU(alpha = 1) = x > 0 ? x : exp(x) - 1 with ONE MUFU: ex2.approx.ftz of x * log2(e) (abs error ~1e-7; very negative x
// flushes to 0 -> -1).  __expf() wraps the same instruction in a denormal-input rescue (FSETP / PLOP3 / 2 x FMUL) that
// doubled the instruction count of the ELU epilogues -- the Cpm trunk kernels are instruction-issue-bound.
__device__ __forceinline__ float lwp_elu(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
  return v > 0.f ? v : e - 1.f;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int cluster,
                              Args &&...args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

}  // namespace lwp
