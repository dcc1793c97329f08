// Shared host/device helpers for the lwpose_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

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

int num_sms();

}  // namespace lwp
