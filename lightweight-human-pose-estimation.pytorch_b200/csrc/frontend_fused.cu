// The full-resolution front end of the backbone as ONE kernel (bf16 plans):
//
//   reference: model[0] conv(3, 32, stride=2) -> model[1] conv_dw(32, 64) -> the depthwise half of model[2]
//   conv_dw(64, 128, stride=2) of models/with_mobilenet.py:92-95 (modules/conv.py:4-22: Conv2d + BatchNorm2d + ReLU each),
//   i.e. stem 3x3/s2 3->32, depthwise 3x3 32, pointwise 32->64, depthwise 3x3/s2 64.  With uint8 frames also val.normalize
//   (val.py:30-33).
//
// As separate kernels these four layers move 2.1 GB per 64-frame step through HBM (32- and 64-channel maps at 184 x 328,
// each written once and read once or twice) for 0.5 ms -- 13 % of the step -- although only the frame (46 / 185 MB) and
// the 64-channel map at 92 x 164 (124 MB) have to cross the chip boundary.  Here a CTA owns an 8 x 16 tile of the
// stride-4 output and recomputes the halos: 17 x 33 pointwise / depthwise-1 pixels (1.10 x redundant), 19 x 35 stem pixels
// (1.30 x), a 39 x 71 x 3 patch of the frame.  Per tile:
//
//   0  patch (prefetched into registers as aligned vectors during the previous tile) -> bf16 -> smem
//   1  im2col: one K row (27 taps + 5 zeros, 64 bytes, SWIZZLE_64B) per stem pixel          -> 6 A tiles of 128 rows
//   2  tcgen05.mma M128 N32 K32 x 6 (one thread)                                             -> TMEM columns 0..191
//   3  TMEM -> BN + ReLU -> bf16, zero outside the image (= the depthwise padding)           -> S tile [19][40][32] smem
//   4  depthwise 3x3 (4 channels x a column of 6 pixels per thread, taps in raster order, packed FFMA2), BN + ReLU ->
//      bf16, written as the K-major SWIZZLE_64B A operand of the pointwise GEMM             -> 5 A tiles of 128 rows
//   5  tcgen05.mma M128 N64 K32 x 5                                                          -> TMEM columns 192..511
//   6  TMEM -> BN + ReLU -> bf16, zero outside the image                                     -> P tile [17][40][64] smem
//   7  depthwise 3x3 stride 2 (4 channels x a column of 4 pixels per thread), BN + ReLU -> bf16 -> global [n][H/4][W/4][64]
//
// Every rounding point and operation order is that of the separate kernels (stem_gemm.cu, conv_direct.cu, conv_gemm.cu +
// gemm_epilogue.cuh), so the result is BIT-IDENTICAL to the unfused path (tests/test_net_gpu.py checks exactly that).
// One 512-thread CTA per SM (all 512 TMEM columns, ~190 KB of shared memory).
//
// Measured (64 x 368x656, bf16 plan): DRAM traffic 276 MB instead of 2.1 GB, but 546 us (float frames) / 587 us (uint8)
// against 471 / 503 us for the four separate HBM-bound kernels: the kernel is bound by instruction ISSUE on the CUDA
// cores -- 31 k warp instructions per tile (unpacking bf16 pairs, FFMA2 at 1.7 / clk / SM, F2FP conversions at 1 / clk /
// SM, swizzled addressing) at 46 % issue utilisation with the 16 warps an SM can hold at 128 registers, phases separated
// by block barriers (profiles/r02k_frontend_fused_ncu_full_summary.csv, scripts/exp_frontend_phases.py: depthwise-1 28 %,
// pointwise epilogue 15 %, depthwise-2 15 %, im2col 10 %, stem epilogue 9 %, the two MMA waits 10 %).  It is therefore
// OPT-IN (LWP_FRONTEND_FUSION=1); the next step would be the depthwise convs as tensor-core GEMMs over shifted
// no-swizzle operand planes (the tensor pipe is 3 % busy here), which gives up bit-identity with the fp32-weight kernels.
#include "common.cuh"
#include "conv_gemm.cuh"
#include "frontend_fused.cuh"
#include "tcgen05.cuh"

namespace lwp {

constexpr int kFeThreads = 512;
constexpr int kFeQH = 8, kFeQW = 16;                      // output tile (stride-4 pixels)
constexpr int kFePH = 2 * kFeQH + 1, kFePW = 2 * kFeQW + 1;   // 17 x 33 pointwise / depthwise-1 pixels
constexpr int kFeSH = kFePH + 2, kFeSW = kFePW + 2;           // 19 x 35 stem pixels
constexpr int kFeIH = 2 * kFeSH + 1, kFeIW = 2 * kFeSW + 1;   // 39 x 71 input pixels
constexpr int kFeIWp = 72;                                    // column pitch of the bf16 patch (float frames: planes [ci][row][72])
constexpr int kFeIBp = 216;                                   // uint8 frames: interleaved like the frame, [row][col * 3 + ci], 216 halfwords per row
constexpr int kFeSPix = kFeSH * kFeSW, kFePPix = kFePH * kFePW;   // 665, 561
constexpr int kFeSTiles = (kFeSPix + 127) / 128, kFePTiles = (kFePPix + 127) / 128;   // 6, 5
// The patch is fetched in aligned 16-byte (float frames) / 4-byte (uint8 frames) vectors: the input origin of a tile is
// 64 tx - 5, so float rows start 3 elements and uint8 rows (3 bytes per pixel) 1 byte after a vector boundary; an aligned
// vector is either completely inside the frame or completely outside (W is a multiple of 8).
constexpr int kFeF32Vec = 20;                                 // float4 per (channel, row) line: columns ix0 - 3 .. ix0 + 76
constexpr int kFeF32Lines = 3 * kFeIH;                        // 117
constexpr int kFeF32Pre = (kFeF32Lines * kFeF32Vec + kFeThreads - 1) / kFeThreads;   // 5 float4 per thread
constexpr int kFeU8Vec = 54;                                  // 32-bit words per row: bytes -1 .. 214 of the 213-byte row
constexpr int kFeU8Pre = (kFeIH * kFeU8Vec + kFeThreads - 1) / kFeThreads;            // 5 words per thread
constexpr int kFeTileBytes = 128 * 64;                        // one SWIZZLE_64B A tile: 128 rows x 64 bytes

// shared memory map (bytes).  The P tile re-uses the patch + stem A tiles (dead once the stem MMAs have completed).
constexpr int kFeOffPatch = 0;
constexpr int kFePatchBytes = ((3 * kFeIH * kFeIWp * 2 + 1023) / 1024) * 1024;            // 16 848 -> 17 408
constexpr int kFeOffStemA = kFeOffPatch + kFePatchBytes;
constexpr int kFePitch = 40;   // pixel pitch of the S and P tiles: a multiple of 8, so the XOR swizzle of a pixel's 16-byte chunks
                               // depends on its column only and the depthwise threads address whole columns with immediates
constexpr int kFePTileBytes = kFePH * kFePitch * 128;                                       // 87 040
constexpr int kFeRegionX = ((((kFeOffStemA + kFeSTiles * kFeTileBytes) > kFePTileBytes ? (kFeOffStemA + kFeSTiles * kFeTileBytes) : kFePTileBytes) + 1023) / 1024) * 1024;
constexpr int kFeOffS = kFeRegionX;                                                         // S tile: 19 x 40 pixels x 64 B
constexpr int kFeSBytes = ((kFeSH * kFePitch * 64 + 1023) / 1024) * 1024;
constexpr int kFeOffD = kFeOffS + kFeSBytes;                                                // 5 A tiles
constexpr int kFeOffW0 = kFeOffD + kFePTiles * kFeTileBytes;                                // stem B operand: 32 rows x 64 B
constexpr int kFeOffW1 = kFeOffW0 + 2048;                                                   // pointwise B operand: 64 rows x 64 B
constexpr int kFeOffConst = kFeOffW1 + 4096;   // floats: stem scale|shift (64), dw1 w[9][32] scale shift (352), pw scale|shift (128), dw2 w[9][64] scale shift (704)
constexpr int kFeConstFloats = 64 + 352 + 128 + 704;
constexpr int kFeOffLut = kFeOffConst + kFeConstFloats * 4;                                 // uint8 -> bf16 LUT [3][256]
constexpr int kFeOffBars = kFeOffLut + 3 * 256 * 2;
constexpr int kFeSmemBytes = kFeOffBars + 64;

struct FrontendParams {
  const void *x;            // NCHW float32 [n][3][H][W]  or  uint8 [n][H][W][3]
  const float *stem_w, *stem_scale, *stem_shift;      // [32][27]
  const float *dw1_w, *dw1_scale, *dw1_shift;         // [9][32]
  const void *pw_w;                                   // bf16 [64][32]
  const float *pw_scale, *pw_shift;
  const float *dw2_w, *dw2_scale, *dw2_shift;         // [9][64]
  void *out;                // bf16 NHWC [n][H/4][W/4][64]
  int n, H, W, Ho, Wo, H2, W2;
  int tiles_x, tiles_y, tiles;
  uint32_t idesc0, idesc1;
  double mean[3], img_scale;
  int *err_flag;
};

__device__ __forceinline__ float2 fe_bf16x2_to_f32x2(uint32_t x) {
  return make_float2(__uint_as_float(__byte_perm(x, 0u, 0x1044)), __uint_as_float(x & 0xffff0000u));
}
__device__ __forceinline__ uint32_t fe_pack_relu(float2 a) {   // round, then ReLU on the rounded pair
  const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
  const __nv_bfloat162 h = __hmax2(__float22bfloat162_rn(a), zero2);
  return *reinterpret_cast<const uint32_t *>(&h);
}

#ifdef LWP_TIMING_EXPERIMENTS
__device__ long long g_fe_prof[16];   // CTA 0, thread 0: cycles per phase summed over its tiles
#define FE_T(k) do { if (blockIdx.x == 0 && tid == 0) { const long long c_ = clock64(); g_fe_prof[k] += c_ - fe_ts; fe_ts = c_; } } while (0)
#else
#define FE_T(k) do {} while (0)
#endif

template <bool kU8>
__global__ void __launch_bounds__(kFeThreads, 1)
frontend_fused_kernel(const FrontendParams p) {
  extern __shared__ uint8_t fe_smem_raw[];
  uint8_t *smem = fe_smem_raw + ((1024u - (ptx::smem_u32(fe_smem_raw) & 1023u)) & 1023u);
  uint16_t *patch = reinterpret_cast<uint16_t *>(smem + kFeOffPatch);
  uint8_t *stemA = smem + kFeOffStemA, *ptile = smem, *stile = smem + kFeOffS, *dtile = smem + kFeOffD;
  uint8_t *w0 = smem + kFeOffW0, *w1 = smem + kFeOffW1;
  float *cst = reinterpret_cast<float *>(smem + kFeOffConst);
  float *c_s0 = cst, *c_dw1 = cst + 64, *c_pw = cst + 416, *c_dw2 = cst + 544;
  uint16_t *lut = reinterpret_cast<uint16_t *>(smem + kFeOffLut);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kFeOffBars);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time set-up: barriers, TMEM, weights and constants (constants only: legal before griddepcontrol.wait) ----
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 512);
  if (tid < 32) {   // stem weights row `tid`: 27 taps + zero padding, K-major SWIZZLE_64B
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = 2 * j < 27 ? p.stem_w[tid * 27 + 2 * j] : 0.f, b = 2 * j + 1 < 27 ? p.stem_w[tid * 27 + 2 * j + 1] : 0.f;
      const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      pk[j] = *reinterpret_cast<const uint32_t *>(&h);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4 *>(w0 + tid * 64 + ((j ^ ((tid >> 1) & 3)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  } else if (tid < 96) {   // pointwise weights row r: 32 bf16 = 64 bytes
    const int r = tid - 32;
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(p.pw_w) + r * 64);
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4 *>(w1 + r * 64 + ((j ^ ((r >> 1) & 3)) << 4)) = src[j];
  }
  for (int i = tid; i < kFeConstFloats; i += kFeThreads) {
    float v;
    if (i < 32) v = p.stem_scale[i];
    else if (i < 64) v = p.stem_shift[i - 32];
    else if (i < 64 + 288) v = p.dw1_w[i - 64];
    else if (i < 64 + 320) v = p.dw1_scale[i - 352];
    else if (i < 416) v = p.dw1_shift[i - 384];
    else if (i < 480) v = p.pw_scale[i - 416];
    else if (i < 544) v = p.pw_shift[i - 480];
    else if (i < 544 + 576) v = p.dw2_w[i - 544];
    else if (i < 544 + 640) v = p.dw2_scale[i - 1120];
    else v = p.dw2_shift[i - 1184];
    cst[i] = v;
  }
  if constexpr (kU8) {   // val.normalize of every possible byte, rounded as the stem rounds its inputs: double -> float -> bf16
    for (int i = tid; i < 768; i += kFeThreads) {
      const int ci = i >> 8, b = i & 255;
      const float f = __double2float_rn(__dmul_rn(__dsub_rn((double)b, p.mean[ci]), p.img_scale));
      const __nv_bfloat16 h = __float2bfloat16_rn(f);
      lut[i] = *reinterpret_cast<const uint16_t *>(&h);
    }
  }
  ptx::fence_proxy_async();   // w0 / w1 were written through the generic proxy, the tensor core reads them through the async proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // the output buffer may still be read by the previous step's kernels

  const uint32_t smem_s = ptx::smem_u32(smem);
  const uint64_t db0 = ptx::umma_desc_k_sw64(smem_s + kFeOffW0), db1 = ptx::umma_desc_k_sw64(smem_s + kFeOffW1);
  const uint8_t *x8 = reinterpret_cast<const uint8_t *>(p.x);
  const float *xf = reinterpret_cast<const float *>(p.x);
  const int per_img = p.tiles_x * p.tiles_y;

  // prefetch registers: the next tile's patch as aligned vectors (see kFeF32Vec / kFeU8Vec); vectors outside the frame are
  // not loaded (pre_ok bit clear) and become the zero padding of the normalised image
  uint4 pre4[kU8 ? 1 : kFeF32Pre];
  uint32_t pre1[kU8 ? kFeU8Pre : 1];
  uint32_t pre_ok = 0;
  auto prefetch = [&](int t) {
    const int img = t / per_img, rem = t - img * per_img, ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int iy0 = 4 * ty * kFeQH - 5, ix0 = 4 * tx * kFeQW - 5;
    pre_ok = 0;
    if constexpr (kU8) {
#pragma unroll
      for (int i = 0; i < kFeU8Pre; ++i) {
        const int v = tid + i * kFeThreads, row = v / kFeU8Vec, wi = v - row * kFeU8Vec;
        const int gy = iy0 + row, gb = ix0 * 3 - 1 + wi * 4;   // first byte of the word inside the frame row
        pre1[i] = 0u;
        if (row < kFeIH && gy >= 0 && gy < p.H && gb >= 0 && gb < p.W * 3) {
          pre1[i] = __ldg(reinterpret_cast<const uint32_t *>(x8 + ((size_t)img * p.H + gy) * (size_t)(p.W * 3) + gb));
          pre_ok |= 1u << i;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kFeF32Pre; ++i) {
        const int v = tid + i * kFeThreads, line = v / kFeF32Vec, qi = v - line * kFeF32Vec;
        const int ci = line / kFeIH, row = line - ci * kFeIH;
        const int gy = iy0 + row, gx = ix0 - 3 + qi * 4;
        pre4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (line < kFeF32Lines && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
          pre4[i] = __ldg(reinterpret_cast<const uint4 *>(xf + (((size_t)img * 3 + ci) * p.H + gy) * (size_t)p.W + gx));
          pre_ok |= 1u << i;
        }
      }
    }
  };

  int t = blockIdx.x;
  if (t < p.tiles) prefetch(t);
  uint32_t parity = 0;
#ifdef LWP_TIMING_EXPERIMENTS
  long long fe_ts = clock64();
  if (blockIdx.x == 0 && tid == 0) for (int k = 0; k < 16; ++k) g_fe_prof[k] = 0;
#endif
  for (; t < p.tiles; t += gridDim.x, parity ^= 1u) {
    const int img = t / per_img, rem = t - img * per_img, ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int qy0 = ty * kFeQH, qx0 = tx * kFeQW;
    const int py0 = 2 * qy0 - 1, px0 = 2 * qx0 - 1, sy0 = py0 - 1, sx0 = px0 - 1;

    // ---- 0: prefetched patch -> bf16 [ci][row][72] ----
    if constexpr (kU8) {
#pragma unroll
      for (int i = 0; i < kFeU8Pre; ++i) {
        const int v = tid + i * kFeThreads, row = v / kFeU8Vec, wi = v - row * kFeU8Vec;
        if (row < kFeIH) {
          const bool okv = (pre_ok >> i) & 1u;
          int ci = (wi + 2) % 3;             // channel of byte 4 wi - 1 of the row
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rc = wi * 4 - 1 + k;   // byte (= halfword) index inside the 213-element patch row
            if (rc >= 0 && rc < kFeIW * 3)
              patch[row * kFeIBp + rc] = okv ? lut[ci * 256 + (int)((pre1[i] >> (8 * k)) & 0xffu)] : (uint16_t)0;
            ci = ci == 2 ? 0 : ci + 1;
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kFeF32Pre; ++i) {
        const int v = tid + i * kFeThreads, line = v / kFeF32Vec, qi = v - line * kFeF32Vec;
        if (line < kFeF32Lines) {
          const uint32_t w4[4] = {pre4[i].x, pre4[i].y, pre4[i].z, pre4[i].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int col = qi * 4 - 3 + k;
            if (col >= 0 && col < kFeIW) {
              const __nv_bfloat16 bb = __float2bfloat16_rn(__uint_as_float(w4[k]));   // 0 for vectors outside the frame
              patch[line * kFeIWp + col] = *reinterpret_cast<const uint16_t *>(&bb);
            }
          }
        }
      }
    }
    __syncthreads();
    FE_T(0);
    if (t + (int)gridDim.x < p.tiles) prefetch(t + gridDim.x);   // in flight during the whole tile
    FE_T(1);

    // ---- 1: im2col, one stem pixel per thread and round ----
#pragma unroll 1
    for (int sp = tid; sp < kFeSPix; sp += kFeThreads) {
      const int sy = sp / kFeSW, sx = sp - sy * kFeSW;
      uint32_t h[32];
#pragma unroll
      for (int k = 27; k < 32; ++k) h[k] = 0u;
      if constexpr (kU8) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {   // the 3 x 3 (kx, ci) values of a window row are 9 consecutive halfwords (4-byte aligned: 12 sx bytes)
          const uint32_t *src = reinterpret_cast<const uint32_t *>(patch + (2 * sy + ky) * kFeIBp + 6 * sx);
          uint32_t w[5];
#pragma unroll
          for (int q = 0; q < 5; ++q) w[q] = src[q];
#pragma unroll
          for (int e = 0; e < 9; ++e) {     // e = kx * 3 + ci
            const uint32_t hv = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu);
            h[(e % 3) * 9 + ky * 3 + e / 3] = hv;
          }
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint16_t *src = patch + (ci * kFeIH + 2 * sy + ky) * kFeIWp + 2 * sx;   // 4-byte aligned (even column)
            const uint32_t ab = *reinterpret_cast<const uint32_t *>(src);
            h[ci * 9 + ky * 3 + 0] = ab & 0xffffu;
            h[ci * 9 + ky * 3 + 1] = ab >> 16;
            h[ci * 9 + ky * 3 + 2] = src[2];
          }
      }
      uint8_t *rowp = stemA + (sp >> 7) * kFeTileBytes + (sp & 127) * 64;
      const int sw = (sp >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4 *>(rowp + ((j ^ sw) << 4)) =
            make_uint4(h[8 * j] | (h[8 * j + 1] << 16), h[8 * j + 2] | (h[8 * j + 3] << 16), h[8 * j + 4] | (h[8 * j + 5] << 16), h[8 * j + 6] | (h[8 * j + 7] << 16));
    }
    ptx::fence_proxy_async();
    __syncthreads();
    FE_T(2);

    // ---- 2: stem GEMM ----
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int m = 0; m < kFeSTiles; ++m) {
        const uint64_t da = ptx::umma_desc_k_sw64(smem_s + kFeOffStemA + m * kFeTileBytes);
        ptx::umma<false>(tmem_base + (uint32_t)(m * 32), da, db0, p.idesc0, 0u);
        ptx::umma<false>(tmem_base + (uint32_t)(m * 32), da + 2u, db0 + 2u, p.idesc0, 1u);
      }
      ptx::umma_commit(&bars[0]);
    }
    if (!ptx::mbar_wait(&bars[0], parity)) { atomicExch(p.err_flag, 61); break; }
    ptx::tc_fence_after();
    FE_T(3);

    // ---- 3: stem epilogue -> S tile (pitch 40 pixels, 64-byte pixels, 16-byte chunks XOR-swizzled with (column >> 1) & 3) ----
    // 48 half-tasks (M tile, lane quarter, 16-column half): warp w takes lane quarter w & 3 and every 4th of its 12
#pragma unroll 1
    for (int ht = warp >> 2; ht < kFeSTiles * 2; ht += 4) {
      const int m = ht >> 1, hh = ht & 1, qq = warp & 3;
      const int sp = m * 128 + qq * 32 + lane;
      uint32_t r[16];
      ptx::tmem_ld_32x16(tmem_base + ((uint32_t)(qq * 32) << 16) + (uint32_t)(m * 32 + hh * 16), r);
      ptx::tmem_ld_wait(r);
      if (sp < kFeSPix) {
        const int sy = sp / kFeSW, sx = sp - sy * kFeSW;
        const int gy = sy0 + sy, gx = sx0 + sx;
        const bool inside = gy >= 0 && gy < p.Ho && gx >= 0 && gx < p.Wo;
        uint8_t *rowp = stile + (sy * kFePitch + sx) * 64;
        const int sw = (sx >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (inside) {
            const int cg = hh * 16 + j * 8;
            const float4 sc0 = *reinterpret_cast<const float4 *>(c_s0 + cg), sc1 = *reinterpret_cast<const float4 *>(c_s0 + cg + 4);
            const float4 sh0 = *reinterpret_cast<const float4 *>(c_s0 + 32 + cg), sh1 = *reinterpret_cast<const float4 *>(c_s0 + 32 + cg + 4);
            pk.x = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y)));
            pk.y = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w)));
            pk.z = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y)));
            pk.w = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w)));
          }
          *reinterpret_cast<uint4 *>(rowp + (((hh * 2 + j) ^ sw) << 4)) = pk;
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    FE_T(4);

    // ---- 4: depthwise 3x3 on the S tile -> A tiles of the pointwise GEMM ----
    {
      const int cq = tid & 7;   // 4 channels
      float2 wk[9][2];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 w4 = *reinterpret_cast<const float4 *>(c_dw1 + k * 32 + cq * 4);
        wk[k][0] = make_float2(w4.x, w4.y); wk[k][1] = make_float2(w4.z, w4.w);
      }
      const float4 sc4 = *reinterpret_cast<const float4 *>(c_dw1 + 288 + cq * 4), sh4 = *reinterpret_cast<const float4 *>(c_dw1 + 320 + cq * 4);
      const float2 sc[2] = {make_float2(sc4.x, sc4.y), make_float2(sc4.z, sc4.w)}, sh[2] = {make_float2(sh4.x, sh4.y), make_float2(sh4.z, sh4.w)};
      // a thread computes a COLUMN segment (6, 6, 5 rows) of 4 channels: the three window columns are three base addresses
      // (their chunk swizzle depends on the column only), every row is an immediate offset.  The four 8-lane groups of a
      // warp take adjacent columns: adjacent 64-byte pixels sit in different halves of the 128-byte bank line.
#pragma unroll 1
      for (int item = tid; item < kFePW * 3 * 8; item += kFeThreads) {
        const int g = item >> 3, vs = g / kFePW, col = g - vs * kFePW, row0 = vs * 6;
        const int nrows = vs == 2 ? 5 : 6;
        const uint8_t *cb[3];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int sx = col + kx;
          cb[kx] = stile + (row0 * kFePitch + sx) * 64 + ((((cq >> 1) ^ (sx >> 1)) & 3) << 4) + (cq & 1) * 8;
        }
        float2 acc[6][2];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int iy = 0; iy < 8; ++iy) {
          if (iy == 7 && nrows == 5) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint2 raw = *reinterpret_cast<const uint2 *>(cb[kx] + iy * (kFePitch * 64));
            const float2 v0 = fe_bf16x2_to_f32x2(raw.x), v1 = fe_bf16x2_to_f32x2(raw.y);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              const int ky = iy - r;
              if (ky >= 0 && ky < 3) {
                acc[r][0] = __ffma2_rn(v0, wk[ky * 3 + kx][0], acc[r][0]);
                acc[r][1] = __ffma2_rn(v1, wk[ky * 3 + kx][1], acc[r][1]);
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          if (r < nrows) {
            const int dp = (row0 + r) * kFePW + col, rr = dp & 127;
            const uint2 o = make_uint2(fe_pack_relu(__ffma2_rn(acc[r][0], sc[0], sh[0])), fe_pack_relu(__ffma2_rn(acc[r][1], sc[1], sh[1])));
            *reinterpret_cast<uint2 *>(dtile + (dp >> 7) * kFeTileBytes + rr * 64 + ((((cq >> 1) ^ (rr >> 1)) & 3) << 4) + (cq & 1) * 8) = o;
          }
        }
      }
    }
    ptx::fence_proxy_async();
    __syncthreads();
    FE_T(5);

    // ---- 5: pointwise GEMM ----
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int m = 0; m < kFePTiles; ++m) {
        const uint64_t da = ptx::umma_desc_k_sw64(smem_s + kFeOffD + m * kFeTileBytes);
        ptx::umma<false>(tmem_base + (uint32_t)(192 + m * 64), da, db1, p.idesc1, 0u);
        ptx::umma<false>(tmem_base + (uint32_t)(192 + m * 64), da + 2u, db1 + 2u, p.idesc1, 1u);
      }
      ptx::umma_commit(&bars[1]);
    }
    if (!ptx::mbar_wait(&bars[1], parity)) { atomicExch(p.err_flag, 62); break; }
    ptx::tc_fence_after();
    FE_T(6);

    // ---- 6: pointwise epilogue -> P tile (128-byte rows, 16-byte chunks XOR-swizzled with row & 7) ----
    // 40 half-tasks (M tile, lane quarter, 32-column half); warp w takes lane quarter w & 3 and every 4th of its 10
#pragma unroll 1
    for (int ht = warp >> 2; ht < kFePTiles * 2; ht += 4) {
      const int m = ht >> 1, half = ht & 1, qq = warp & 3;
      const int pp = m * 128 + qq * 32 + lane;
      const int py = pp / kFePW, px = pp - py * kFePW;
      const int gy = py0 + py, gx = px0 + px;
      const bool inside = pp < kFePPix && gy >= 0 && gy < p.Ho && gx >= 0 && gx < p.Wo;
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(qq * 32) << 16) + (uint32_t)(192 + m * 64 + half * 32), r);
      ptx::tmem_ld_wait(r);
      if (pp < kFePPix) {
        uint8_t *rowp = ptile + (py * kFePitch + px) * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (inside) {
            const int cg = half * 32 + j * 8;
            const float4 sc0 = *reinterpret_cast<const float4 *>(c_pw + cg), sc1 = *reinterpret_cast<const float4 *>(c_pw + cg + 4);
            const float4 sh0 = *reinterpret_cast<const float4 *>(c_pw + 64 + cg), sh1 = *reinterpret_cast<const float4 *>(c_pw + 64 + cg + 4);
            pk.x = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y)));
            pk.y = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w)));
            pk.z = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y)));
            pk.w = fe_pack_relu(__ffma2_rn(make_float2(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w)));
          }
          *reinterpret_cast<uint4 *>(rowp + (((half * 4 + j) ^ (px & 7)) << 4)) = pk;
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    FE_T(7);

    // ---- 7: depthwise 3x3 stride 2 on the P tile -> global ----
    {
      const int cq = tid & 15;   // 4 channels
      float2 wk[9][2];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 w4 = *reinterpret_cast<const float4 *>(c_dw2 + k * 64 + cq * 4);
        wk[k][0] = make_float2(w4.x, w4.y); wk[k][1] = make_float2(w4.z, w4.w);
      }
      const float4 sc4 = *reinterpret_cast<const float4 *>(c_dw2 + 576 + cq * 4), sh4 = *reinterpret_cast<const float4 *>(c_dw2 + 640 + cq * 4);
      const float2 sc[2] = {make_float2(sc4.x, sc4.y), make_float2(sc4.z, sc4.w)}, sh[2] = {make_float2(sh4.x, sh4.y), make_float2(sh4.z, sh4.w)};
      // one item per thread: 4 channels of 4 vertically adjacent outputs (a 9 x 3 window: three column base addresses, rows
      // as immediates); the 16 lanes of a half-warp read the 128 bytes of one pixel
      {
        const int g = tid >> 4, qx = g & 15, qyl = (g >> 4) * 4;
        const uint8_t *cb[3];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int px = 2 * qx + kx;
          cb[kx] = ptile + (2 * qyl * kFePitch + px) * 128 + ((((cq >> 1) ^ px) & 7) << 4) + (cq & 1) * 8;
        }
        float2 acc[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int iy = 0; iy < 9; ++iy)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint2 raw = *reinterpret_cast<const uint2 *>(cb[kx] + iy * (kFePitch * 128));
            const float2 v0 = fe_bf16x2_to_f32x2(raw.x), v1 = fe_bf16x2_to_f32x2(raw.y);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int ky = iy - 2 * r;
              if (ky >= 0 && ky < 3) {
                acc[r][0] = __ffma2_rn(v0, wk[ky * 3 + kx][0], acc[r][0]);
                acc[r][1] = __ffma2_rn(v1, wk[ky * 3 + kx][1], acc[r][1]);
              }
            }
          }
        const int gx = qx0 + qx;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int gy = qy0 + qyl + r;
          if (gy < p.H2 && gx < p.W2) {
            const uint2 o = make_uint2(fe_pack_relu(__ffma2_rn(acc[r][0], sc[0], sh[0])), fe_pack_relu(__ffma2_rn(acc[r][1], sc[1], sh[1])));
            *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + (((size_t)img * p.H2 + gy) * p.W2 + gx) * 64 + cq * 4) = o;
          }
        }
      }
    }
    __syncthreads();   // the P tile is dead: the next tile's patch and A tiles may overwrite it
    FE_T(8);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

int frontend_fused_launch(const FrontendArgs &a, int *err_flag, cudaStream_t st) {
  FrontendParams p;
  p.x = a.x; p.stem_w = a.stem_w; p.stem_scale = a.stem_scale; p.stem_shift = a.stem_shift;
  p.dw1_w = a.dw1_w; p.dw1_scale = a.dw1_scale; p.dw1_shift = a.dw1_shift;
  p.pw_w = a.pw_w; p.pw_scale = a.pw_scale; p.pw_shift = a.pw_shift;
  p.dw2_w = a.dw2_w; p.dw2_scale = a.dw2_scale; p.dw2_shift = a.dw2_shift;
  p.out = a.out;
  p.n = a.n; p.H = a.H; p.W = a.W; p.Ho = a.H / 2; p.Wo = a.W / 2; p.H2 = a.H / 4; p.W2 = a.W / 4;
  p.tiles_x = ceil_div(p.W2, kFeQW); p.tiles_y = ceil_div(p.H2, kFeQH);
  p.tiles = a.n * p.tiles_x * p.tiles_y;
  p.idesc0 = make_umma_idesc(false, kBlockM, 32);
  p.idesc1 = make_umma_idesc(false, kBlockM, 64);
  for (int i = 0; i < 3; ++i) p.mean[i] = a.mean[i];
  p.img_scale = a.img_scale;
  p.err_flag = err_flag;
  const size_t smem = (size_t)kFeSmemBytes + 1024;
  static DeviceOnce attr;
  int attr_slot;
  if (attr.pending(&attr_slot)) {
    LWP_CUDA_CHECK(cudaFuncSetAttribute(frontend_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LWP_CUDA_CHECK(cudaFuncSetAttribute(frontend_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.done[attr_slot] = true;
  }
  const int grid = p.tiles < net_sms() ? p.tiles : net_sms();
  if (a.x_is_u8) LWP_CUDA_CHECK(launch_pdl(frontend_fused_kernel<true>, grid, kFeThreads, smem, st, 1, p));
  else LWP_CUDA_CHECK(launch_pdl(frontend_fused_kernel<false>, grid, kFeThreads, smem, st, 1, p));
  return LWP_OK;
}

#ifdef LWP_TIMING_EXPERIMENTS
extern "C" int lwp_debug_frontend_prof(long long *out_host) { return cudaMemcpyFromSymbol(out_host, g_fe_prof, sizeof(long long) * 16) == cudaSuccess ? 0 : 1; }
#endif

size_t frontend_fused_smem_bytes() { return (size_t)kFeSmemBytes + 1024; }

}  // namespace lwp
