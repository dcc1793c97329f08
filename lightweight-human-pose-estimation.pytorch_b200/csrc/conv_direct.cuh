// Launchers of the CUDA-core layer kernels (conv_direct.cu).
#pragma once
#include <cuda_runtime.h>

namespace lwp {
int stem_launch(bool f32, const float *x, const float *w, const float *scale, const float *shift, void *out, int n,
                int H, int W, cudaStream_t st);
int depthwise_launch(bool f32, const void *in, void *out, const float *w9c, const float *scale, const float *shift,
                     int n, int H, int W, int C, int stride, int dil, int act, cudaStream_t st);
int nhwc_to_nchw_launch(bool in_f32, const void *in, int ld, int c0, int c, float *out, int n, int HW,
                        cudaStream_t st);
}  // namespace lwp
