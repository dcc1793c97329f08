// Stem convolution (3x3, stride 2, pad 1, 3 -> 32, BN + ReLU folded) as an im2col GEMM on tcgen05 tensor cores.
//
//   reference: conv(3, 32, stride=2, bias=False) of models/with_mobilenet.py:93 (modules/conv.py:4-10: Conv2d +
//   BatchNorm2d + ReLU); with kU8 also val.normalize (val.py:30-33) of the raw BGR frame.
//
// K = 27 is far too thin for a TMA-fed implicit GEMM (the taps of a stride-2 window are not a box of the NCHW
// input), but it is a perfectly good tensor-core problem once the A operand exists: every thread gathers the 27
// inputs of ONE output pixel, rounds them to the plan dtype and writes that pixel's K row straight into the
// K-major SWIZZLE_128B shared-memory tile (27 real + 5 zero columns = one 64-byte half row in bf16 / one whole
// 128-byte row in tf32); one thread issues tcgen05.mma (M = 128 pixels, N = 32 channels, K = 32); the epilogue
// reads the fp32 accumulators back from TMEM (thread = pixel) and stores the 32 channels of its pixel.  The CUDA
// cores move data (about 400 instructions per output pixel, mostly addressing, packing and the epilogue, instead of ~550 for the direct FFMA form), the 864
// MACs per pixel go to the tensor pipe, and the kernel becomes a streaming read of the frames and write of the
// first activation.  Several small CTAs per SM (each with its own 64 TMEM columns) hide the load latency, and inside a
// CTA the gather is software-pipelined: the loads of the tile after next are in flight while the current tile's
// accumulator is drained and stored (gather() / commit() below).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_direct.cuh"
#include "conv_gemm.cuh"
#include "tcgen05.cuh"

namespace lwp {

constexpr int kStemThreads = 128;         // one output pixel per thread
constexpr int kStemStages = 2;            // A tiles / accumulators in flight per CTA
constexpr int kStemN = 32, kStemK = 27;
constexpr int kStemTmemCols = kStemStages * kStemN;   // 64 (power of two >= 32)

struct StemGemmParams {
  const void *x;            // NCHW float32 [n][3][H][W]  or  uint8 [n][H][W][3]
  const float *w;           // [32][27] (co, ci*9 + ky*3 + kx)
  const float *scale, *shift;
  void *out;                // NHWC [n][H/2][W/2][32] in the plan dtype
  int n, H, W, Ho, Wo;
  long long total;          // n * Ho * Wo output pixels
  int tiles;
  uint32_t idesc;
  double mean[3], img_scale;
  // uint8 frames whose normalisation is exact in float32 (integer means, power-of-two scale -- the reference's 128 and
  // 1/256): v = fma(float(8388608 + b), nscale, noff[ci]) with the byte OR-ed into the mantissa of 2^23, no conversions
  int fast_norm;
  float nscale, noff[3];
  int *err_flag;
};

template <bool kTf32>
__device__ __forceinline__ void stem_store_row(uint8_t *tile, int r, const float (&v)[32]) {
  uint8_t *row = tile + r * kKBlockBytes;
  if constexpr (kTf32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4 *>(row + ((j ^ (r & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 pk;
      __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
      *reinterpret_cast<uint4 *>(row + ((j ^ (r & 7)) << 4)) = pk;
    }
  }
}

template <bool kTf32, int kU8>   // kU8: 0 float32 NCHW frames, 1 uint8 frames (float64 normalisation), 2 uint8 frames (exact float32 short cut)
__global__ void __launch_bounds__(kStemThreads, 5)
stem_gemm_kernel(const StemGemmParams p) {
  extern __shared__ uint8_t stem_smem_raw[];
  uint8_t *smem = stem_smem_raw + ((1024u - (ptx::smem_u32(stem_smem_raw) & 1023u)) & 1023u);
  uint8_t *a_tiles = smem;                                          // kStemStages x 16 KB
  uint8_t *b_tile = smem + kStemStages * kATileBytes;               // 32 rows x 128 B
  float *s_scale = reinterpret_cast<float *>(b_tile + kStemN * kKBlockBytes);
  float *s_shift = s_scale + kStemN;
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_shift + kStemN);  // one per stage: "these MMAs have completed"
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kStemStages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStemStages; ++s) ptx::mbar_init(&bars[s], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, kStemTmemCols);
  if (tid < kStemN) {  // weights row `tid`: 27 taps + zero padding, K-major, same swizzle as the A rows
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < kStemK ? p.w[tid * kStemK + k] : 0.f;
    stem_store_row<kTf32>(b_tile, tid, v);
    if constexpr (!kTf32) {  // the unused upper half of the 128-byte rows (never read: K = 32 of 64) stays defined
#pragma unroll
      for (int j = 4; j < 8; ++j) *reinterpret_cast<uint4 *>(b_tile + tid * kKBlockBytes + ((j ^ (tid & 7)) << 4)) = make_uint4(0, 0, 0, 0);
    }
    s_scale[tid] = p.scale[tid];
    s_shift[tid] = p.shift[tid];
  }
  if constexpr (!kTf32) {  // same for the A tiles
    for (int s = 0; s < kStemStages; ++s)
#pragma unroll
      for (int j = 4; j < 8; ++j)
        *reinterpret_cast<uint4 *>(a_tiles + s * kATileBytes + tid * kKBlockBytes + ((j ^ (tid & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // (the output buffer may still be read by the previous step's kernels)
  const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(b_tile));
  const float *xf = reinterpret_cast<const float *>(p.x);
  const uint8_t *x8 = reinterpret_cast<const uint8_t *>(p.x);
  const int HoWo = p.Ho * p.Wo;

  // The A row of a tile is built in two steps so that the gather's memory latency overlaps the previous tile's epilogue:
  // gather() issues the 27 loads of this thread's pixel into registers (raw values, taps outside the image stay 0 and
  // are flagged), commit() normalises / rounds them and writes the pixel's K row into stage s.
  struct Taps {
    float f[27];     // float32 frames: the values; uint8 frames: the bytes (as integers in float registers' bit patterns)
    uint32_t skip;   // bit 0: no pixel (past the end), bit 1: window row 0 is outside, bit 2: window column 0 is outside
  };
  auto gather = [&](int t, Taps &g) {
    const long long pix = (long long)t * kBlockM + tid;
#pragma unroll
    for (int k = 0; k < 27; ++k) g.f[k] = 0.f;
    g.skip = 1u;
    if (pix < p.total) {
      const int img = (int)(pix / HoWo), rem = (int)(pix - (long long)img * HoWo);
      const int yo = rem / p.Wo, xo = rem - yo * p.Wo;
      const int yi0 = 2 * yo - 1, xi0 = 2 * xo - 1;
      // H and W are even, so only the first row / column of the window can fall outside (yo == 0 / xo == 0)
      const bool top = yo == 0, left = xo == 0;
      g.skip = (top ? 2u : 0u) | (left ? 4u : 0u);
      if constexpr (kU8 != 0) {
        const uint8_t *base = x8 + (((long long)img * p.H + yi0) * p.W + xi0) * 3;   // may point before the frame: guarded
        const int rowb = p.W * 3;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          if (ky == 0 && top) continue;
          const uint8_t *row = base + ky * rowb;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (kx == 0 && left) continue;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) g.f[ci * 9 + ky * 3 + kx] = __uint_as_float((uint32_t)__ldg(row + kx * 3 + ci));
          }
        }
      } else {
        const float *base = xf + (((long long)img * 3) * p.H + yi0) * p.W + xi0;
        const int HW = p.H * p.W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            if (ky == 0 && top) continue;
            const float *row = base + (ci * HW + ky * p.W);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              if (kx == 0 && left) continue;
              g.f[ci * 9 + ky * 3 + kx] = __ldg(row + kx);
            }
          }
        }
      }
    }
  };
  auto commit = [&](const Taps &g, int s) {
    float v[32];
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int k = ci * 9 + ky * 3 + kx;
          if constexpr (kU8 != 0) {
            // zero padding applies to the NORMALISED image (pad value 0 after normalize)
            const bool out = (g.skip & 1u) || (ky == 0 && (g.skip & 2u)) || (kx == 0 && (g.skip & 4u));
            const uint32_t b = __float_as_uint(g.f[k]);
            // val.normalize in float64 like the reference (two conversions + two FP64 operations per tap: the
            // conversion pipe made this variant 30 us slower than the float32 one) -- or, when that is exact in
            // float32, one LOP3 + one FFMA with the same bits
            float nv;
            if constexpr (kU8 == 2) nv = __fmaf_rn(__uint_as_float(0x4B000000u | b), p.nscale, p.noff[ci]);
            else nv = __double2float_rn(__dmul_rn(__dsub_rn((double)b, p.mean[ci]), p.img_scale));
            v[k] = out ? 0.f : nv;
          } else {
            v[k] = g.f[k];
          }
        }
    stem_store_row<kTf32>(a_tiles + s * kATileBytes, tid, v);
  };
  // one thread: the K = 32 product of stage s into accumulator stage s, completion on bars[s]
  auto mma = [&](int s) {
    const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(a_tiles + s * kATileBytes));
    constexpr int kSteps = kTf32 ? 4 : 2;   // 32 bytes of K per instruction: 8 tf32 / 16 bf16 elements
#pragma unroll
    for (int k = 0; k < kSteps; ++k)
      ptx::umma<kTf32>(tmem_base + (uint32_t)(s * kStemN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (uint32_t)(k != 0));
    ptx::umma_commit(&bars[s]);
  };

  int t = blockIdx.x;
  Taps g;
  if (t < p.tiles) {
    gather(t, g);
    commit(g, 0);
    ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) { ptx::tc_fence_after(); mma(0); }
    if (t + (int)gridDim.x < p.tiles) gather(t + gridDim.x, g);   // in flight during the first epilogue
  }
  uint32_t phases = 0u;   // bit s: parity the next wait on bars[s] expects
  for (int s = 0; t < p.tiles; t += gridDim.x, s ^= 1) {
    const int tn = t + gridDim.x;
    // ---- epilogue of tile t (its MMA was issued one iteration ago); the loads of tile tn are in flight ----
    if (!ptx::mbar_wait(&bars[s], (phases >> s) & 1u)) { atomicExch(p.err_flag, 21); break; }
    phases ^= 1u << s;
    ptx::tc_fence_after();
    uint32_t r[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * kStemN), r);
    ptx::tmem_ld_wait(r);
    const long long pix = (long long)t * kBlockM + tid;
    if (kTf32 && pix < p.total) {
      float y[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) y[c] = fmaxf(fmaf(__uint_as_float(r[c]), s_scale[c], s_shift[c]), 0.f);
      float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + pix * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) op[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
    }
    if constexpr (!kTf32) {
      // bf16: the warp's 32 pixels are 2 KB of contiguous output.  They are staged in the warp's own rows of the A tile
      // the MMA has just finished reading (only this warp rewrites those rows, in commit() two tiles later) and leave
      // as ONE bulk copy instead of 4 x 32 scattered 16-byte stores.
      uint8_t *stg = a_tiles + s * kATileBytes + warp * 4096;
      float y[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) y[c] = fmaxf(fmaf(__uint_as_float(r[c]), s_scale[c], s_shift[c]), 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(y[8 * j + 2 * q], y[8 * j + 2 * q + 1]);
        *reinterpret_cast<uint4 *>(stg + lane * 64 + j * 16) = pk;
      }
      ptx::fence_proxy_async();
      __syncwarp();
      const long long pix0 = (long long)t * kBlockM + warp * 32;
      if (lane == 0 && pix0 < p.total) {
        const long long left = p.total - pix0;
        ptx::bulk_store_1d(reinterpret_cast<__nv_bfloat16 *>(p.out) + pix0 * 32, stg, (uint32_t)(left < 32 ? left : 32) * 64u);
        ptx::bulk_commit();
      }
      // stage s ^ 1 is written next: the bulk store of tile t - 1 (everything but the store just committed) must have
      // finished reading its staging rows there
      if (lane == 0) ptx::bulk_wait_read<1>();
      __syncwarp();
    }
    // ---- next tile: its A rows into stage s ^ 1 (released: MMA waited for, accumulator drained, staging read), MMA ----
    if (tn < p.tiles) {
      commit(g, s ^ 1);
      ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (tn < p.tiles && tid == 0) { ptx::tc_fence_after(); mma(s ^ 1); }
    if (tn + (int)gridDim.x < p.tiles) gather(tn + gridDim.x, g);   // the tile after next: in flight during the next epilogue
  }
  if constexpr (!kTf32) {
    if (lane == 0) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, kStemTmemCols);
}

int stem_gemm_launch(bool f32, const void *x, bool x_is_u8, const double *mean3, double img_scale, const float *w,
                     const float *scale, const float *shift, void *out, int n, int H, int W, int *err_flag,
                     cudaStream_t st) {
  StemGemmParams p;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.out = out;
  p.n = n; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2;
  p.total = (long long)n * p.Ho * p.Wo;
  p.tiles = (int)((p.total + kBlockM - 1) / kBlockM);
  p.idesc = make_umma_idesc(f32, kBlockM, kStemN);
  p.mean[0] = mean3 ? mean3[0] : 0; p.mean[1] = mean3 ? mean3[1] : 0; p.mean[2] = mean3 ? mean3[2] : 0;
  p.img_scale = img_scale;
  {
    int e = 0;
    bool ok = img_scale > 0 && frexp(img_scale, &e) == 0.5 && e > -100 && e < 100;
    for (int c = 0; c < 3; ++c) ok = ok && p.mean[c] == floor(p.mean[c]) && p.mean[c] >= 0 && p.mean[c] <= 1048576.0;
    if (getenv("LWP_STEM_F64_NORM") != nullptr) ok = false;   // (tests: the float64 path on the default constants)
    p.fast_norm = ok ? 1 : 0;
    p.nscale = (float)img_scale;
    for (int c = 0; c < 3; ++c) p.noff[c] = (float)(-(8388608.0 + p.mean[c]) * img_scale);
  }
  p.err_flag = err_flag;
  const size_t smem = 1024 + kStemStages * kATileBytes + kStemN * kKBlockBytes + 2 * kStemN * sizeof(float) + 64;
  static DeviceOnce attr;
  int attr_slot;
  if (attr.pending(&attr_slot)) {
    LWP_CUDA_CHECK(cudaFuncSetAttribute(stem_gemm_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LWP_CUDA_CHECK(cudaFuncSetAttribute(stem_gemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LWP_CUDA_CHECK(cudaFuncSetAttribute(stem_gemm_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LWP_CUDA_CHECK(cudaFuncSetAttribute(stem_gemm_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LWP_CUDA_CHECK(cudaFuncSetAttribute(stem_gemm_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LWP_CUDA_CHECK(cudaFuncSetAttribute(stem_gemm_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.done[attr_slot] = true;
  }
  int per_sm = 5;   // 5 x (38 KB smem, 64 TMEM columns, 128 threads) per SM
  if (const char *e = getenv("LWP_STEM_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 8) per_sm = v; }
  int grid = net_sms() * per_sm;
  if (grid > p.tiles) grid = p.tiles;
  const int mode = !x_is_u8 ? 0 : (p.fast_norm ? 2 : 1);
  if (f32) {
    if (mode == 2) LWP_CUDA_CHECK(launch_pdl(stem_gemm_kernel<true, 2>, grid, kStemThreads, smem, st, 1, p));
    else if (mode == 1) LWP_CUDA_CHECK(launch_pdl(stem_gemm_kernel<true, 1>, grid, kStemThreads, smem, st, 1, p));
    else LWP_CUDA_CHECK(launch_pdl(stem_gemm_kernel<true, 0>, grid, kStemThreads, smem, st, 1, p));
  } else {
    if (mode == 2) LWP_CUDA_CHECK(launch_pdl(stem_gemm_kernel<false, 2>, grid, kStemThreads, smem, st, 1, p));
    else if (mode == 1) LWP_CUDA_CHECK(launch_pdl(stem_gemm_kernel<false, 1>, grid, kStemThreads, smem, st, 1, p));
    else LWP_CUDA_CHECK(launch_pdl(stem_gemm_kernel<false, 0>, grid, kStemThreads, smem, st, 1, p));
  }
  return LWP_OK;
}

}  // namespace lwp
