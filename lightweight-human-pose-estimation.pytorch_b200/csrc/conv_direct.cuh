// Launchers of the CUDA-core layer kernels (conv_direct.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lwp {
int stem_launch(bool f32, const void *x, bool x_is_u8, const double *mean3, double img_scale, const float *w,
                const float *scale, const float *shift, void *out, int n, int H, int W, cudaStream_t st);
// the same layer as an im2col GEMM on tcgen05 (stem_gemm.cu); err_flag: the plan's pipeline-timeout flag
int stem_gemm_launch(bool f32, const void *x, bool x_is_u8, const double *mean3, double img_scale, const float *w,
                     const float *scale, const float *shift, void *out, int n, int H, int W, int *err_flag,
                     cudaStream_t st);
int depthwise_launch(bool f32, const void *in, void *out, const float *w9c, const float *scale, const float *shift,
                     int n, int H, int W, int C, int stride, int dil, int act, cudaStream_t st);
struct DwTileGeom {
  int tw, th, iw, ih, cb, cv, Ho, Wo, tiles_x, tiles_y, cblocks, num_tiles;
  uint32_t stage_bytes;
};
int depthwise_tma_geometry(bool f32, int n, int H, int W, int C, int stride, int dil, DwTileGeom *g);
int depthwise_tma_init();
int depthwise_tma_launch(bool f32, const CUtensorMap &tm, const CUtensorMap *tm_out /* NULL: register stores */, void *out, const float *w9c, const float *scale,
                         const float *shift, int n, int H, int W, int C, int stride, int dil, int act,
                         const DwTileGeom &g, cudaStream_t st);
int nhwc_to_nchw_launch(bool in_f32, const void *in, int ld, int c0, int c, float *out, int n, int HW,
                        cudaStream_t st);
}  // namespace lwp
