// Parameter block shared by the tcgen05 implicit-GEMM convolution kernel and the plan that launches it.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "gemm_epilogue.cuh"

namespace lwp {

constexpr int kEpiWarps = 8;         // two epilogue warps per TMEM lane quarter (alternate 128-byte chunks)
constexpr int kGemmThreads = 64 + 32 * kEpiWarps + 32;  // warp 0: TMA producer (activations), warp 1: MMA issuer + TMEM owner, 8 epilogue warps, last warp: TMA producer (weights)
constexpr int kBProducerWarp = 2 + kEpiWarps;
#ifndef LWP_GEMM_BOUND_THREADS
#define LWP_GEMM_BOUND_THREADS kGemmThreads
#endif
constexpr int kGemmBoundThreads = LWP_GEMM_BOUND_THREADS;   // register cap of the GEMM kernels = 65536 / this (experiments: 512 -> 128 registers)
constexpr int kBlockM = 128;        // pixels per tile == UMMA M == TMEM lanes
constexpr int kKBlockBytes = 128;   // one SWIZZLE_128B row of K per pipeline stage
constexpr int kATileBytes = kBlockM * kKBlockBytes;
constexpr int kMaxStages = 8;
constexpr int kMaxAccStages = 8;    // TMEM accumulator stages: 512 columns / block_n
constexpr int kMaxCout = 1024;

struct GemmParams {
  int H, W, NIMG;          // geometry the A tensor map was built with (flattened 1x1: H = NIMG = 1, W = pixels)
  int tile_w, tile_h;      // tile_w * tile_h == 128
  int tiles_x, tiles_y;    // tiles per image
  int m_tiles, n_tiles;    // total M tiles, Cout_pad / block_n
  int block_n;             // UMMA N (multiple of 32, <= 256)
  int taps, dil;           // 1 or 9; tap offset = (k - 1) * dil
  int kblocks_per_tap;     // ceil(Cin * elem_size / 128)
  int kb_elems;            // elements per K block: 64 (bf16) or 32 (tf32); half of that for a 64-byte K block
  int kb_bytes;            // bytes of one K-block row in shared memory: 128 (SWIZZLE_128B), or 64 (SWIZZLE_64B) for a
                           // single-K-block layer whose Cin is exactly 64 bytes (a 128-byte box would be half out of
                           // bounds, which the TMA unit fills very slowly)
  int cin;                 // B's K coordinate of tap t, block b = t * cin + b * kb_elems
  int cout_pad, n_store;   // columns computed / columns written (multiple of 8)
  uint32_t idesc;          // UMMA instruction descriptor (M = 128, N = block_n)
  int act;                 // LWP_ACT_*
  int num_stages;
  uint32_t tmem_cols;      // power of two >= acc_stages * block_n
  int acc_stages;          // TMEM accumulator ring depth (2..8)
  const float *scale, *shift;
  const void *residual;    // plan dtype, pixel stride res_ld, or nullptr
  int res_ld;
  void *out;               // plan dtype, pixel stride out_ld, or nullptr
  int out_ld;
  float *out_f32;          // optional float32 copy, pixel stride out_f32_ld
  int out_f32_ld;
  int *err_flag;           // set non-zero if a pipeline wait timed out
  int debug;               // timing experiments only (LWP_DEBUG_GEMM): bit 0 skip A loads, 1 skip B loads, 2 skip MMAs, 3 skip epilogue,
                           // 4 skip the epilogue's tensor stores, 5 skip the epilogue's math
  // epilogue through shared memory + TMA store (plain single-output layers): each epilogue warp stages its
  // 32 pixel rows x 128 bytes of output and one lane issues a 4-D tensor store of that box
  int tma_store;           // 0: direct register -> global stores
  int store_bw, store_bh;  // pixel box of one warp's 32 rows (store_bw * store_bh == 32)
  int kbps;                // K blocks per pipeline stage (1 or 2): one barrier round trip then covers 4 or 8 MMAs;
                           // a stage is kbps consecutive [A tile | B tile] pairs (conv_gemm_kernel only)
  int c3_a_stages, c3_b_stages;   // conv3x3_pair_kernel: activation-strip / weight-tile ring depths
  // conv3x3_pair_kernel with a fused trailing 1x1 conv (bf16, Cout = 128 -> 128): out = act2((act(conv3x3 + ...) + residual) W2^T * scale2 + shift2)
  int fuse_pw;             // 0 / 1
  const float *scale2, *shift2;
  int act2;
  int staging_bufs;        // 1 or 2 staging buffers per epilogue warp (2: the next chunk is converted while the
                           // tensor store of the previous one still reads its buffer); conv_gemm_kernel only
  int wres_sub;            // conv_gemm_wres_kernel: 128-pixel tiles per pipeline stage (0 = another kernel runs the layer)
};

constexpr int kStagingBytes = kEpiWarps * kStageOutBytes;  // one staging buffer per epilogue warp

size_t conv_gemm_smem_bytes(const GemmParams &p);
int conv_gemm_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                     const GemmParams &p, int grid, cudaStream_t st);
int conv_gemm_init();
// CTA-pair (cta_group::2) variant, conv_gemm2.cu
// heads_gemm.cu: both 1x1 layers of a stage's heads as one back-to-back GEMM (bf16 plans)
int heads_fused_launch(const CUtensorMap &tmX, const CUtensorMap &tmW1, const CUtensorMap &tmW2, int n_px, int c_in,
                       int c_mid, const float *scale1, const float *shift1, const float *scale2, const float *shift2,
                       float *out_f32, int out_f32_ld, void *out_bf16, int out_ld, int *err_flag, cudaStream_t st);
size_t heads_fused_smem_bytes(int c_in, int c_mid);
int heads_fused_chunk_cols(int c_mid);   // rows of one W1 TMA box
size_t conv_gemm3_smem_bytes(const GemmParams &p);
int conv_gemm3_init();
int conv_gemm3_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                      const GemmParams &p, int grid, cudaStream_t st);
int conv_gemm3_pw_launch(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const CUtensorMap &tmW2,
                         const GemmParams &p, int grid, cudaStream_t st);
// weight-resident thin 1x1 layers, several M tiles per stage (conv_gemm_wres.cu)
size_t conv_gemm_wres_smem_bytes(const GemmParams &p);
int conv_gemm_wres_init();
int conv_gemm_wres_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                          const GemmParams &p, int grid, cudaStream_t st);
size_t conv_gemm2_smem_bytes(const GemmParams &p);
int conv_gemm2_init();
int conv_gemm2_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                      const GemmParams &p, int grid, cudaStream_t st);

}  // namespace lwp
