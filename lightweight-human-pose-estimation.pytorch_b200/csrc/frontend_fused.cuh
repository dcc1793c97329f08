// Fused full-resolution front end (stem -> depthwise -> pointwise -> stride-2 depthwise), frontend_fused.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace lwp {

struct FrontendArgs {
  const void *x = nullptr;           // float32 NCHW [n][3][H][W] or uint8 [n][H][W][3]
  bool x_is_u8 = false;
  double mean[3] = {0, 0, 0}, img_scale = 1.0;
  const float *stem_w = nullptr, *stem_scale = nullptr, *stem_shift = nullptr;   // [32][27], [32], [32]
  const float *dw1_w = nullptr, *dw1_scale = nullptr, *dw1_shift = nullptr;      // [9][32]
  const void *pw_w = nullptr;                                                   // bf16 [64][32]
  const float *pw_scale = nullptr, *pw_shift = nullptr;
  const float *dw2_w = nullptr, *dw2_scale = nullptr, *dw2_shift = nullptr;      // [9][64]
  void *out = nullptr;                                                          // bf16 NHWC [n][H/4][W/4][64]
  int n = 0, H = 0, W = 0;
};

int frontend_fused_launch(const FrontendArgs &a, int *err_flag, cudaStream_t st);
size_t frontend_fused_smem_bytes();

}  // namespace lwp
