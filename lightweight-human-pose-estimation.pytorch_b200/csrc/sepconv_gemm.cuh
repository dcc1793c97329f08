// Parameter block of the weight-resident fused depthwise + pointwise kernel (sepconv_gemm.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace lwp {

struct SepParams {
  int H, W, NIMG;                 // input size
  int Ho, Wo;                     // output size: (H - 1) / stride + 1
  int stride;                     // 1 or 2 (3x3, pad 1, dilation 1)
  int tile_w, tile_h;             // OUTPUT tile, tile_w * tile_h == 128, tile_w multiple of 4, tile_h multiple of 2
  int tiles_x, tiles_y, m_tiles;
  int iw, ih;                     // input halo box of one tile: (tile - 1) * stride + 3
  int cin, kblocks, kb_ch;        // depthwise channels == GEMM K; K blocks of 128 bytes (64 bf16 / 32 tf32 channels)
  int cout_pad, n_store;          // GEMM N (multiple of 64, <= 256) / columns written (multiple of 64)
  int slice_bytes;                // bytes of one staged output row of one warp's column quarter: n_store / 4 elements (32, 64 or 128)
  uint32_t idesc;
  int acc_stages;                 // TMEM accumulator ring: 512 / cout_pad (2..4)
  int in_stages, a_stages;        // halo-box ring / A-tile ring depths (one entry = one K block of one tile)
  uint32_t in_stage_bytes;        // iw * ih * 128
  int dw_act, act;
  const float *dw_consts;         // [kblocks][9 taps | scale | shift][kb_ch] fp32 (folded BN), loaded once per CTA
  const float *scale, *shift;     // pointwise epilogue
  const void *residual;
  int res_ld;
  int *err_flag;
};

size_t sepconv_smem_bytes(const SepParams &p);
int sepconv_init();
int sepconv_launch(bool tf32, const CUtensorMap &tmIn, const CUtensorMap &tmB, const CUtensorMap &tmC, const SepParams &p,
                   int grid, cudaStream_t st);

}  // namespace lwp
