// Parameter block of the fused depthwise + pointwise kernel (dwpw_gemm.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace lwp {

struct DwpwParams {
  int H, W, NIMG;                 // the block keeps the spatial size (depthwise stride 1)
  int tile_w, tile_h, tiles_x, tiles_y, m_tiles;
  int iw, ih;                     // input halo box of one tile
  int dil;
  int cin;                        // depthwise channels == GEMM K
  int kblocks, kb_ch;             // K blocks of 128 bytes: 64 bf16 / 32 tf32 channels
  int cout_pad, n_store;          // GEMM N (multiple of 64, <= 512) / columns written
  int n_mma, n_per_mma;           // N is issued as n_mma instructions of n_per_mma (<= 256) columns
  uint32_t idesc;
  int acc_stages;                 // 512 / cout_pad TMEM accumulator stages
  int in_stages, a_stages, b_stages, staging_bufs;
  uint32_t in_stage_bytes, b_stage_bytes;   // in_stage_bytes = halo tile + the K block's depthwise constants
  uint32_t in_tile_bytes, dw_const_bytes;   // dw constants of one K block: [9 taps | scale | shift][kb_ch] fp32
  int dw_act, act;
  const float *dw_consts;                      // depthwise: [kblocks][9 taps | scale | shift][kb_ch] fp32 (folded BN)
  const float *scale, *shift;                  // pointwise epilogue
  const void *residual;
  int res_ld;
  int *err_flag;
  int debug;                      // bit 0: skip depthwise math, bit 1: skip epilogue, bit 2: skip MMA (timing experiments only)
};

size_t dwpw_smem_bytes(const DwpwParams &p);
int dwpw_init();
int dwpw_launch(bool tf32, const CUtensorMap &tmIn, const CUtensorMap &tmB, const CUtensorMap &tmC, const DwpwParams &p,
                int grid, cudaStream_t st);

}  // namespace lwp
