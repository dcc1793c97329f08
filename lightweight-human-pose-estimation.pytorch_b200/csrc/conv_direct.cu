// CUDA-core kernels for the bandwidth-bound layers: stem 3x3/s2 (K = 27 is too thin for tensor cores),
// depthwise 3x3 (stride 1/2, dilation 1/2) with fused BN + ReLU / ELU, and the NHWC -> NCHW hand-off.
//
// Reference: modules/conv.py:4-10 (stem via conv(3, 32, stride=2, bias=False)), :13-32 (conv_dw, conv_dw_no_bn);
// models/with_mobilenet.py:93-105.
#include "common.cuh"
#include "conv_direct.cuh"
#include "tcgen05.cuh"

namespace lwp {

template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
    uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p));
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __bfloat1622float2(h[j]);
      v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8]) {
    uint4 pk;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4 *>(p) = pk;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float *p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float *p, const float (&v)[8]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == LWP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == LWP_ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}

// ------------------------------------------------------------------------------------------------
// stem: NCHW fp32 [n][3][H][W] -> NHWC [n][H/2][W/2][32]; 3x3, stride 2, pad 1; BN folded; ReLU.
// One thread = two horizontally adjacent output pixels x all 32 output channels: every weight pair read
// from shared memory (LDS.128 = two pairs) feeds two packed FFMA2s, the 5 input columns the two pixels
// share are loaded once.
// ------------------------------------------------------------------------------------------------
struct StemNorm {  // (pixel - mean[c]) * scale of val.normalize, evaluated in double like NumPy does, then rounded to fp32
  double mean[3];
  double scale;
};

template <typename T, bool kU8>
__global__ void __launch_bounds__(128)
stem_kernel(const void *__restrict__ xin, const float *__restrict__ w, const float *__restrict__ scale,
            const float *__restrict__ shift, T *__restrict__ out, int n, int H, int W, int Ho, int Wo,
            const StemNorm nrm) {
  const float *x = reinterpret_cast<const float *>(xin);          // kU8 == false: NCHW float32
  const uint8_t *x8 = reinterpret_cast<const uint8_t *>(xin);     // kU8 == true: [n][H][W][3] uint8 (BGR frame)
  __shared__ __align__(16) float s_w[27][32];  // [ci*9 + ky*3 + kx][co]
  __shared__ __align__(16) float s_scale[32], s_shift[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
    int co = i & 31, k = i >> 5;
    s_w[k][co] = w[co * 27 + k];
  }
  if (threadIdx.x < 32) { s_scale[threadIdx.x] = scale[threadIdx.x]; s_shift[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  const int Wp = (Wo + 1) / 2;  // pixel pairs per output row
  const long long total = (long long)n * Ho * Wp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(idx % Wp);
    long long t = idx / Wp;
    const int yo = (int)(t % Ho);
    const int img = (int)(t / Ho);
    const int xo = 2 * xp;                 // first of the two output pixels
    const bool has2 = xo + 1 < Wo;
    float2 acc[2][16];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[q][j] = make_float2(0.f, 0.f);
    const float *xb = x + (size_t)img * 3 * H * W;
    const uint8_t *xb8 = x8 + (size_t)img * 3 * H * W;
    const int xi0 = 2 * xo - 1;            // leftmost input column (pixel 0 taps xi0..xi0+2, pixel 1 taps xi0+2..xi0+4)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yi = 2 * yo - 1 + ky;
        if (yi < 0 || yi >= H) continue;
        const float *row = xb + ((size_t)ci * H + yi) * W;
        const uint8_t *row8 = xb8 + (size_t)yi * W * 3 + ci;
        float v[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const int xi = xi0 + c;
          if constexpr (kU8) {  // zero padding applies to the NORMALISED image, as in the reference (pad_value 0 after normalize)
            v[c] = (xi >= 0 && xi < W)
                       ? __double2float_rn(__dmul_rn(__dsub_rn((double)__ldg(row8 + (size_t)xi * 3), nrm.mean[ci]), nrm.scale))
                       : 0.f;
          } else {
            v[c] = (xi >= 0 && xi < W) ? __ldg(row + xi) : 0.f;
          }
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 *wr = reinterpret_cast<const float4 *>(&s_w[ci * 9 + ky * 3 + kx][0]);
          const float2 a = make_float2(v[kx], v[kx]), b = make_float2(v[kx + 2], v[kx + 2]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w4 = wr[j];
            acc[0][2 * j] = __ffma2_rn(a, make_float2(w4.x, w4.y), acc[0][2 * j]);
            acc[0][2 * j + 1] = __ffma2_rn(a, make_float2(w4.z, w4.w), acc[0][2 * j + 1]);
            acc[1][2 * j] = __ffma2_rn(b, make_float2(w4.x, w4.y), acc[1][2 * j]);
            acc[1][2 * j + 1] = __ffma2_rn(b, make_float2(w4.z, w4.w), acc[1][2 * j + 1]);
          }
        }
      }
    }
    const size_t pix = ((size_t)img * Ho + yo) * Wo + xo;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q == 1 && !has2) break;
      T *op = out + (pix + q) * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g * 8 + 2 * j;
          const float2 r = __ffma2_rn(acc[q][g * 4 + j], make_float2(s_scale[c], s_scale[c + 1]),
                                      make_float2(s_shift[c], s_shift[c + 1]));
          o[2 * j] = fmaxf(r.x, 0.f);
          o[2 * j + 1] = fmaxf(r.y, 0.f);
        }
        Vec8<T>::store(op + g * 8, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, NHWC.  One thread = 8 channels x TW consecutive output columns; the input columns
// of a row are streamed through registers once and scattered to the outputs they touch, so every
// input vector is loaded once per thread (3 rows x ((TW-1)*S + 2*D + 1) loads for TW outputs).
// Weights arrive tap-major [9][C] so the 8 channels of a thread are one 32-byte read.
// ------------------------------------------------------------------------------------------------
template <typename T, int S, int D, int TW>
__global__ void __launch_bounds__(256)
depthwise3x3_kernel(const T *__restrict__ in, T *__restrict__ out, const float *__restrict__ w9c,
                    const float *__restrict__ scale, const float *__restrict__ shift, int n, int H, int W, int C,
                    int Ho, int Wo, int act) {
  constexpr int NCOL = (TW - 1) * S + 2 * D + 1;
  const int cvecs = C >> 3;
  const int xtiles = (Wo + TW - 1) / TW;
  const long long total = (long long)n * Ho * xtiles * cvecs;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvecs);
    long long t = idx / cvecs;
    const int xt = (int)(t % xtiles);
    t /= xtiles;
    const int yo = (int)(t % Ho);
    const int img = (int)(t / Ho);
    const int c0 = cv * 8;
    const int xo0 = xt * TW;
    float acc[TW][8];
#pragma unroll
    for (int a = 0; a < TW; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    const int xi0 = xo0 * S - D;  // leftmost input column of the window
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yi = yo * S - D + ky * D;
      if (yi < 0 || yi >= H) continue;
      float wk[3][8];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) Vec8<float>::load(w9c + (size_t)(ky * 3 + kx) * C + c0, wk[kx]);
      const T *rowp = in + (((size_t)img * H + yi) * W) * C + c0;
#pragma unroll
      for (int ci = 0; ci < NCOL; ++ci) {
        const int xi = xi0 + ci;
        float v[8];
        if (xi >= 0 && xi < W) Vec8<T>::load(rowp + (size_t)xi * C, v);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int a = 0; a < TW; ++a) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (a * S + kx * D == ci) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[a][j] = fmaf(v[j], wk[kx][j], acc[a][j]);
            }
          }
        }
      }
    }
    float sc[8], sh[8];
    Vec8<float>::load(scale + c0, sc);
    Vec8<float>::load(shift + c0, sh);
#pragma unroll
    for (int a = 0; a < TW; ++a) {
      const int xo = xo0 + a;
      if (xo >= Wo) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_apply(fmaf(acc[a][j], sc[j], sh[j]), act);
      Vec8<T>::store(out + (((size_t)img * Ho + yo) * Wo + xo) * C + c0, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, NHWC, shared-memory halo tiles fed by TMA (reference: the groups=Cin Conv2d of
// modules/conv.py:15,27 + BatchNorm/ReLU or ELU).
// A tile = TH x TW output pixels x CB channels (CB*sizeof(T) <= 128 bytes per pixel).  A CTA keeps ONE channel
// block for its whole life, so the 9 taps and the folded BN of a thread's 4 channels live in registers and
// shared-memory bandwidth is spent on activations only.  One elected thread keeps a ring of NS halo boxes in
// flight by TMA (out-of-image parts are zero-filled = the convolution's padding); all 256 threads compute.
// A thread owns 4 channels x (2 rows x 4 columns) of outputs: the (2-1)*S+2*D+1 by (4-1)*S+2*D+1 input
// window is read once from shared memory (LDS.64 for bf16) and feeds packed fp32 FFMA2s in the same tap
// order (ky-major) for every output, i.e. the same bits whatever the tile shape.
// ------------------------------------------------------------------------------------------------
struct DwTileParams {
  int n, H, W, C, Ho, Wo;
  int tw, th;          // output tile (tw multiple of 4, th multiple of 2, (tw/4)*(th/2)*cq == 256)
  int iw, ih;          // input box
  int cb, cq;          // channels per tile, 4-channel groups per pixel (cb/4)
  int tiles_x, tiles_y, cblocks, sp_tiles;   // sp_tiles = n * tiles_x * tiles_y spatial tiles per channel block
  int act, stages;
  int reverse;         // 1: walk the tiles from the last to the first: the previous layer's GEMM wrote its tiles in
                       // ascending order, so its LAST tiles are the ones still in L2 when this kernel starts reading
  int tma_out;         // 1: outputs staged in smem and written by one TMA tensor store per tile (2 staging buffers)
  uint32_t stg_bytes;  // tw * th * cb * sizeof(T)
  int debug;           // timing experiments only (LWP_DEBUG_DW): bit 0 skip the window loads + FMAs, bit 1 skip the stores
  uint32_t stage_bytes;
};

__device__ __forceinline__ float2 bf16x2_to_f32x2(uint32_t x) {  // PRMT + LOP3: keeps the FMA pipe for the FFMA2s
  return make_float2(__uint_as_float(__byte_perm(x, 0u, 0x1044)), __uint_as_float(x & 0xffff0000u));
}
__device__ __forceinline__ float elu1(float v) { return lwp_elu(v); }  // ELU(alpha = 1) as in the fused block: abs error ~1e-7
#ifndef LWP_DW_STCS
#define LWP_DW_STCS 0
#endif
template <typename T> struct SmemVec4;  // 4 channels from shared memory as two packed fp32 pairs; activation + store of 4 channels
template <> struct SmemVec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(uint32_t saddr, float2 (&v)[2]) {
    uint2 raw;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "r"(saddr));
    v[0] = bf16x2_to_f32x2(raw.x); v[1] = bf16x2_to_f32x2(raw.y);
  }
  template <int ACT>
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, float2 a, float2 b) {
    if constexpr (ACT == LWP_ACT_ELU) { a.x = elu1(a.x); a.y = elu1(a.y); b.x = elu1(b.x); b.y = elu1(b.y); }
    __nv_bfloat162 lo = __float22bfloat162_rn(a), hi = __float22bfloat162_rn(b);
    if constexpr (ACT == LWP_ACT_RELU) {  // ReLU after rounding == rounding after ReLU
      const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
      lo = __hmax2(lo, zero2); hi = __hmax2(hi, zero2);
    }
    if (LWP_DW_STCS) __stcs(reinterpret_cast<uint2 *>(p), make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi)));
    else *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
  }
};
template <> struct SmemVec4<float> {
  static __device__ __forceinline__ void load(uint32_t saddr, float2 (&v)[2]) {
    float4 a;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(saddr));
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
  }
  template <int ACT>
  static __device__ __forceinline__ void store(float *p, float2 a, float2 b) {
    if constexpr (ACT == LWP_ACT_ELU) { a.x = elu1(a.x); a.y = elu1(a.y); b.x = elu1(b.x); b.y = elu1(b.y); }
    if constexpr (ACT == LWP_ACT_RELU) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); }
    *reinterpret_cast<float4 *>(p) = make_float4(a.x, a.y, b.x, b.y);
  }
};

constexpr int kDwMaxStages = 4;

struct DwTileCursor {  // (image, tile row, tile column) of a spatial tile index, advanced by a fixed stride without divisions
  int img, ty, tx;
  __device__ __forceinline__ void init(int j, int per_img, int tiles_x) {
    img = j / per_img;
    const int rem = j - img * per_img;
    ty = rem / tiles_x;
    tx = rem - ty * tiles_x;
  }
  __device__ __forceinline__ void retreat(const DwTileCursor &d, int tiles_x, int tiles_y) {   // the same walk backwards
    tx -= d.tx;
    if (tx < 0) { tx += tiles_x; --ty; }
    ty -= d.ty;
    if (ty < 0) { ty += tiles_y; --img; }
    img -= d.img;
  }
  __device__ __forceinline__ void move(const DwTileCursor &d, int tiles_x, int tiles_y, bool backwards) {
    if (backwards) retreat(d, tiles_x, tiles_y); else advance(d, tiles_x, tiles_y);
  }
  __device__ __forceinline__ void advance(const DwTileCursor &d, int tiles_x, int tiles_y) {
    tx += d.tx;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    ty += d.ty;
    if (ty >= tiles_y) { ty -= tiles_y; ++img; }
    img += d.img;
  }
};

template <typename T, int S, int D, int ACT, int PIXB>   // PIXB: bytes of one pixel of the smem tile (cb * sizeof(T): 64 or 128)
__global__ void __launch_bounds__(256, 2)
depthwise3x3_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                        T *__restrict__ out,
                        const float *__restrict__ w9c, const float *__restrict__ scale,
                        const float *__restrict__ shift, const DwTileParams p) {
  constexpr int R = 2, CC = 4;                            // outputs per thread: R rows x CC columns (x 4 channels)
  // Dilated (D = 2, stride 1): a thread's outputs are D pixels apart, (y0 + D r, x0 + D c), and the D x D threads of a
  // group interleave over a (R D) x (CC D) block -- the taps of neighbouring outputs then coincide exactly as in the
  // undilated case (4 x 6 shared-memory loads per thread instead of the 6 x 8 of a contiguous 2 x 4 block: the dilated
  // 512-channel layer ran 18 % slower than its undilated siblings on twice the loads and bf16 unpacks).
  constexpr bool IL = (D == 2 && S == 1);
  constexpr int SP = IL ? D : 1;                          // spacing of a thread's outputs
  constexpr int PS = IL ? D : 1;                          // spacing of the window positions a thread reads
  constexpr int NROW = ((R - 1) * S * SP + 2 * D) / PS + 1, NCOL = ((CC - 1) * S * SP + 2 * D) / PS + 1;
  extern __shared__ uint8_t dw_smem_raw[];
  uint8_t *smem = dw_smem_raw + ((128u - (ptx::smem_u32(dw_smem_raw) & 127u)) & 127u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);   // one "full" barrier per stage
  uint8_t *bufs = smem + 128;
  const int tid = threadIdx.x;
  if (tid == 0) {
    ptx::prefetch_tmap(&tm_in);
    if (p.tma_out) ptx::prefetch_tmap(&tm_out);
    for (int s = 0; s < p.stages; ++s) ptx::mbar_init(&bars[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();   // before the first halo load (previous kernel's output) and the first store (its input)
  const int cq = tid % p.cq, pb = tid / p.cq;
  const int xgroups = p.tw / CC;
  const int xg = pb % xgroups, yg = pb / xgroups;
  // first output of this thread inside the tile
  const int x0 = IL ? (xg / D) * (CC * D) + xg % D : xg * CC;
  const int y0 = IL ? (yg / D) * (R * D) + yg % D : yg * R;
  const int cblk = blockIdx.x % p.cblocks;               // neighbouring CTAs: same pixels, adjacent channel blocks
  const int j0 = blockIdx.x / p.cblocks, jstride = gridDim.x / p.cblocks;
  const int per_img = p.tiles_x * p.tiles_y;
  const int c0 = cblk * p.cb + cq * 4;
  DwTileCursor step, cur, nxt;                            // nxt: the tile the producer thread loads next
  const bool rev = p.reverse != 0;
  step.init(jstride, per_img, p.tiles_x);
  cur.init(rev ? p.sp_tiles - 1 - j0 : j0, per_img, p.tiles_x);
  nxt = cur;

  auto issue = [&](const DwTileCursor &t, int buf) {
    ptx::mbar_arrive_expect_tx(&bars[buf], p.stage_bytes);
    ptx::tma_load_4d(bufs + (size_t)buf * p.stage_bytes, &tm_in, &bars[buf], cblk * p.cb, t.tx * p.tw * S - D,
                     t.ty * p.th * S - D, t.img);
  };
  int jn = j0;                                            // index of nxt
  if (tid == 0) {
    for (int s = 0; s < p.stages - 1; ++s) {
      if (jn < p.sp_tiles) issue(nxt, s);
      nxt.move(step, p.tiles_x, p.tiles_y, rev);
      jn += jstride;
    }
  }

  // this thread's constants: 9 taps, folded BN scale/shift of 4 channels
  float2 wk[9][2], sc[2], sh[2];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(w9c + (size_t)k * p.C + c0));
    wk[k][0] = make_float2(a.x, a.y); wk[k][1] = make_float2(a.z, a.w);
  }
  {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(scale + c0)), b = __ldg(reinterpret_cast<const float4 *>(shift + c0));
    sc[0] = make_float2(a.x, a.y); sc[1] = make_float2(a.z, a.w);
    sh[0] = make_float2(b.x, b.y); sh[1] = make_float2(b.z, b.w);
  }
  const int row_bytes = p.iw * PIXB;
  const uint32_t win_off = (uint32_t)((y0 * S) * row_bytes + x0 * S * PIXB + cq * 4 * (int)sizeof(T));
  const uint32_t bufs_s = ptx::smem_u32(bufs);
  const size_t opix = (size_t)p.C;                        // elements between horizontally adjacent output pixels
  const size_t orow = (size_t)p.Wo * p.C;

  int buf = 0, it = 0;
  uint32_t phase = 0;
  for (int j = j0; j < p.sp_tiles; j += jstride) {
    if (tid == 0) {  // refill the buffer every thread left at the end of the previous iteration
      if (jn < p.sp_tiles) issue(nxt, buf == 0 ? p.stages - 1 : buf - 1);
      nxt.move(step, p.tiles_x, p.tiles_y, rev);
      jn += jstride;
    }
    if (!ptx::mbar_wait(&bars[buf], phase)) return;  // never spin forever
    const uint32_t win = bufs_s + (uint32_t)buf * p.stage_bytes + win_off;
    float2 acc[R][CC][2];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < CC; ++c) acc[r][c][0] = acc[r][c][1] = make_float2(0.f, 0.f);
    if (!(LWP_DBG(p.debug) & 1))
#pragma unroll
    for (int iy = 0; iy < NROW; ++iy) {
      const uint32_t rowp = win + (uint32_t)(iy * PS * row_bytes);
#pragma unroll
      for (int ic = 0; ic < NCOL; ++ic) {
        float2 v[2];
        SmemVec4<T>::load(rowp + ic * PS * PIXB, v);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
            if (r * S * SP + ky * D == iy * PS) {
#pragma unroll
              for (int c = 0; c < CC; ++c)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                  if (c * S * SP + kx * D == ic * PS) {
                    acc[r][c][0] = __ffma2_rn(v[0], wk[ky * 3 + kx][0], acc[r][c][0]);
                    acc[r][c][1] = __ffma2_rn(v[1], wk[ky * 3 + kx][1], acc[r][c][1]);
                  }
            }
      }
    }
    const int yo0 = cur.ty * p.th + y0, xo0 = cur.tx * p.tw + x0;
    if (p.tma_out) {
      // outputs -> staging tile [th][tw][cb] in shared memory -> one TMA tensor store of the tile (rows / columns past the
      // image are clipped by the hardware).  Two staging buffers: the store of tile i-1 may still read the other one.
      uint8_t *stg = bufs + (size_t)p.stages * p.stage_bytes + (size_t)(it & 1) * p.stg_bytes;
      __syncthreads();   // thread 0 has waited for the store that last read this staging buffer (two tiles ago)
      T *sp = reinterpret_cast<T *>(stg + (y0 * p.tw + x0) * PIXB) + cq * 4;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CC; ++c)
          SmemVec4<T>::template store<ACT>(sp + (r * SP * p.tw + c * SP) * (PIXB / (int)sizeof(T)), __ffma2_rn(acc[r][c][0], sc[0], sh[0]),
                                           __ffma2_rn(acc[r][c][1], sc[1], sh[1]));
      ptx::fence_proxy_async();
      __syncthreads();   // also: everyone is done with input buffer `buf`
      if (tid == 0) {
        if (!(LWP_DBG(p.debug) & 2)) {
          ptx::tma_store_4d(&tm_out, stg, cblk * p.cb, cur.tx * p.tw, cur.ty * p.th, cur.img);
          ptx::bulk_commit();
        }
        ptx::bulk_wait_read<1>();   // the previous tile's store has finished reading the other staging buffer
      }
      cur.move(step, p.tiles_x, p.tiles_y, rev);
      ++it;
      if (++buf == p.stages) { buf = 0; phase ^= 1u; }
      continue;
    }
    T *op = out + (((size_t)cur.img * p.Ho + yo0) * p.Wo + xo0) * p.C + c0;
    const bool full = yo0 + (R - 1) * SP + 1 <= p.Ho && xo0 + (CC - 1) * SP + 1 <= p.Wo;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      T *oq = op;
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        if (!(LWP_DBG(p.debug) & 2) && (full || (yo0 + r * SP < p.Ho && xo0 + c * SP < p.Wo)))
          SmemVec4<T>::template store<ACT>(oq, __ffma2_rn(acc[r][c][0], sc[0], sh[0]), __ffma2_rn(acc[r][c][1], sc[1], sh[1]));
        oq += opix * SP;
      }
      op += orow * SP;
    }
    cur.move(step, p.tiles_x, p.tiles_y, rev);
    __syncthreads();  // everyone is done with buffer `buf` before it is refilled at the top of the next iteration
    if (++buf == p.stages) { buf = 0; phase ^= 1u; }
  }
  if (p.tma_out && tid == 0) ptx::bulk_wait<0>();   // all tensor stores have landed
}

// NHWC (T or float, pixel stride ld, channels [c0, c0 + c)) -> NCHW float32 [n][c][H][W]
template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T *__restrict__ in, int ld, int c0, int c, float *__restrict__ out, int n, int HW) {
  const long long total = (long long)n * c * HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % HW);
    long long t = idx / HW;
    const int ch = (int)(t % c);
    const int img = (int)(t / c);
    out[idx] = (float)in[((size_t)img * HW + p) * ld + c0 + ch];
  }
}

static int grid_for(long long total, int block, int per_sm) {
  long long b = (total + block - 1) / block;
  long long cap = (long long)net_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int stem_launch(bool f32, const void *x, bool x_is_u8, const double *mean3, double img_scale, const float *w,
                const float *scale, const float *shift, void *out, int n, int H, int W, cudaStream_t st) {
  const int Ho = H / 2, Wo = W / 2;
  long long total = (long long)n * Ho * ((Wo + 1) / 2);
  int grid = grid_for(total, 128, 16);
  StemNorm nrm;
  nrm.mean[0] = mean3 ? mean3[0] : 0; nrm.mean[1] = mean3 ? mean3[1] : 0; nrm.mean[2] = mean3 ? mean3[2] : 0;
  nrm.scale = img_scale;
  if (x_is_u8) {
    if (f32) stem_kernel<float, true><<<grid, 128, 0, st>>>(x, w, scale, shift, (float *)out, n, H, W, Ho, Wo, nrm);
    else stem_kernel<__nv_bfloat16, true><<<grid, 128, 0, st>>>(x, w, scale, shift, (__nv_bfloat16 *)out, n, H, W, Ho, Wo, nrm);
  } else {
    if (f32) stem_kernel<float, false><<<grid, 128, 0, st>>>(x, w, scale, shift, (float *)out, n, H, W, Ho, Wo, nrm);
    else stem_kernel<__nv_bfloat16, false><<<grid, 128, 0, st>>>(x, w, scale, shift, (__nv_bfloat16 *)out, n, H, W, Ho, Wo, nrm);
  }
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

template <typename T>
static int depthwise_launch_t(const T *in, T *out, const float *w9c, const float *scale, const float *shift, int n,
                              int H, int W, int C, int stride, int dil, int act, cudaStream_t st) {
  constexpr int TW = 4;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;  // k = 3, pad = dil
  long long total = (long long)n * Ho * ((Wo + TW - 1) / TW) * (C / 8);
  int grid = grid_for(total, 256, 8);
  if (stride == 1 && dil == 1)
    depthwise3x3_kernel<T, 1, 1, TW><<<grid, 256, 0, st>>>(in, out, w9c, scale, shift, n, H, W, C, Ho, Wo, act);
  else if (stride == 2 && dil == 1)
    depthwise3x3_kernel<T, 2, 1, TW><<<grid, 256, 0, st>>>(in, out, w9c, scale, shift, n, H, W, C, Ho, Wo, act);
  else if (stride == 1 && dil == 2)
    depthwise3x3_kernel<T, 1, 2, TW><<<grid, 256, 0, st>>>(in, out, w9c, scale, shift, n, H, W, C, Ho, Wo, act);
  else {
    set_error("depthwise: unsupported stride %d / dilation %d", stride, dil);
    return LWP_EINVAL;
  }
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

int depthwise_launch(bool f32, const void *in, void *out, const float *w9c, const float *scale, const float *shift,
                     int n, int H, int W, int C, int stride, int dil, int act, cudaStream_t st) {
  if (f32)
    return depthwise_launch_t<float>((const float *)in, (float *)out, w9c, scale, shift, n, H, W, C, stride, dil, act,
                                     st);
  return depthwise_launch_t<__nv_bfloat16>((const __nv_bfloat16 *)in, (__nv_bfloat16 *)out, w9c, scale, shift, n, H, W,
                                           C, stride, dil, act, st);
}


// Tile geometry for the TMA depthwise kernel (shared by the plan, which builds the tensor map, and the launcher).
int depthwise_tma_geometry(bool f32, int n, int H, int W, int C, int stride, int dil, DwTileGeom *g) {
  const int es = f32 ? 4 : 2;
  const int cb = C * es <= 128 ? C : 128 / es;
  if (C % cb != 0 || (cb * es != 64 && cb * es != 128)) return LWP_EINVAL;   // other widths: the direct kernel
  const int cq = cb / 4, blocks = 256 / cq;   // 2x4-pixel blocks per tile
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  long long best = -1;
  for (int bx = 1; bx <= blocks; bx <<= 1) {
    const int tw = 4 * bx, th = 2 * (blocks / bx);
    if (dil == 2 && stride == 1 && (tw % 8 != 0 || th % 4 != 0)) continue;   // dilated: 2 x 2 thread groups interleave over 4 x 8 blocks
    const int iw = (tw - 1) * stride + 2 * dil + 1, ih = (th - 1) * stride + 2 * dil + 1;
    if (iw > 256 || ih > 256) continue;
    if ((long long)iw * ih * cb * es > 100 * 1024) continue;
    // cost ~ bytes moved: padded input boxes + padded outputs
    long long tiles = (long long)ceil_div(Wo, tw) * ceil_div(Ho, th);
    long long cost = tiles * ((long long)iw * ih + (long long)tw * th);
    if (best < 0 || cost < best) { best = cost; g->tw = tw; g->th = th; g->iw = iw; g->ih = ih; }
  }
  if (best < 0) return LWP_EINVAL;
  g->cb = cb; g->cv = cq; g->Ho = Ho; g->Wo = Wo;
  g->tiles_x = ceil_div(Wo, g->tw); g->tiles_y = ceil_div(Ho, g->th); g->cblocks = C / cb;
  g->num_tiles = n * g->tiles_x * g->tiles_y;   // spatial tiles per channel block
  g->stage_bytes = (uint32_t)(g->iw * g->ih * cb * es);
  return LWP_OK;
}

int depthwise_tma_init() {
  static DeviceOnce once;
  int slot;
  if (!once.pending(&slot)) return LWP_OK;
#define LWP_DW_ATTR1(T, S, D, A) \
  LWP_CUDA_CHECK(cudaFuncSetAttribute(depthwise3x3_tma_kernel<T, S, D, A, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024)); \
  LWP_CUDA_CHECK(cudaFuncSetAttribute(depthwise3x3_tma_kernel<T, S, D, A, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024))
#define LWP_DW_ATTR(T, S, D) \
  LWP_DW_ATTR1(T, S, D, LWP_ACT_NONE); LWP_DW_ATTR1(T, S, D, LWP_ACT_RELU); LWP_DW_ATTR1(T, S, D, LWP_ACT_ELU)
  LWP_DW_ATTR(float, 1, 1); LWP_DW_ATTR(float, 2, 1); LWP_DW_ATTR(float, 1, 2);
  LWP_DW_ATTR(__nv_bfloat16, 1, 1); LWP_DW_ATTR(__nv_bfloat16, 2, 1); LWP_DW_ATTR(__nv_bfloat16, 1, 2);
#undef LWP_DW_ATTR
#undef LWP_DW_ATTR1
  once.done[slot] = true;
  return LWP_OK;
}

template <typename T>
static int depthwise_tma_launch_t(const CUtensorMap &tm, const CUtensorMap &tm_out, T *out, const float *w9c, const float *scale,
                                  const float *shift, DwTileParams p, int stride, int dil, cudaStream_t st) {
  // two CTAs per SM with a ring of up to 4 halo boxes each; big (stride-2) boxes: one CTA per SM
  int per_sm = 2;
  p.stg_bytes = (uint32_t)(p.tw * p.th * p.cb * (int)sizeof(T));
  const size_t stg = p.tma_out ? 2 * (size_t)p.stg_bytes : 0;
  int stages = (int)((100 * 1024 - stg) / p.stage_bytes);
  if (stages < 2) { per_sm = 1; stages = (int)((200 * 1024 - stg) / p.stage_bytes); }
  if (stages > kDwMaxStages) stages = kDwMaxStages;
  if (stages < 2) { set_error("depthwise: halo box of %u bytes does not fit twice in shared memory", p.stage_bytes); return LWP_ECAP; }
  if (const char *e = getenv("LWP_DW_STAGES")) { int v = atoi(e); if (v >= 2 && v <= stages) stages = v; }
  p.stages = stages;
  const size_t smem = 128 + 128 + (size_t)stages * p.stage_bytes + stg;
  int per_cblk = net_sms() * per_sm / p.cblocks;
  if (per_cblk < 1) per_cblk = 1;
  if (per_cblk > p.sp_tiles) per_cblk = p.sp_tiles;
  const int grid = per_cblk * p.cblocks;
  const int pixb = p.cb * (int)sizeof(T);
  if (pixb != 64 && pixb != 128) { set_error("depthwise: %d bytes per tile pixel (need 64 or 128)", pixb); return LWP_EINVAL; }
#define LWP_DW_GO(S_, D_)                                                                                             \
  do {                                                                                                              \
    if (pixb == 128) {                                                                                              \
      if (p.act == LWP_ACT_RELU) LWP_CUDA_CHECK(launch_pdl(depthwise3x3_tma_kernel<T, S_, D_, LWP_ACT_RELU, 128>, grid, 256, smem, st, 1, tm, tm_out, out, w9c, scale, shift, p)); \
      else if (p.act == LWP_ACT_ELU) LWP_CUDA_CHECK(launch_pdl(depthwise3x3_tma_kernel<T, S_, D_, LWP_ACT_ELU, 128>, grid, 256, smem, st, 1, tm, tm_out, out, w9c, scale, shift, p)); \
      else LWP_CUDA_CHECK(launch_pdl(depthwise3x3_tma_kernel<T, S_, D_, LWP_ACT_NONE, 128>, grid, 256, smem, st, 1, tm, tm_out, out, w9c, scale, shift, p)); \
    } else {                                                                                                        \
      if (p.act == LWP_ACT_RELU) LWP_CUDA_CHECK(launch_pdl(depthwise3x3_tma_kernel<T, S_, D_, LWP_ACT_RELU, 64>, grid, 256, smem, st, 1, tm, tm_out, out, w9c, scale, shift, p)); \
      else if (p.act == LWP_ACT_ELU) LWP_CUDA_CHECK(launch_pdl(depthwise3x3_tma_kernel<T, S_, D_, LWP_ACT_ELU, 64>, grid, 256, smem, st, 1, tm, tm_out, out, w9c, scale, shift, p)); \
      else LWP_CUDA_CHECK(launch_pdl(depthwise3x3_tma_kernel<T, S_, D_, LWP_ACT_NONE, 64>, grid, 256, smem, st, 1, tm, tm_out, out, w9c, scale, shift, p)); \
    }                                                                                                               \
  } while (0)
  if (stride == 1 && dil == 1) LWP_DW_GO(1, 1);
  else if (stride == 2 && dil == 1) LWP_DW_GO(2, 1);
  else if (stride == 1 && dil == 2) LWP_DW_GO(1, 2);
#undef LWP_DW_GO
  else { set_error("depthwise: unsupported stride %d / dilation %d", stride, dil); return LWP_EINVAL; }
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

int depthwise_tma_launch(bool f32, const CUtensorMap &tm, const CUtensorMap *tm_out, void *out, const float *w9c, const float *scale,
                         const float *shift, int n, int H, int W, int C, int stride, int dil, int act,
                         const DwTileGeom &g, cudaStream_t st) {
  DwTileParams p;
  p.n = n; p.H = H; p.W = W; p.C = C; p.Ho = g.Ho; p.Wo = g.Wo; p.tw = g.tw; p.th = g.th; p.iw = g.iw; p.ih = g.ih;
  p.cb = g.cb; p.cq = g.cv; p.tiles_x = g.tiles_x; p.tiles_y = g.tiles_y; p.cblocks = g.cblocks;
  p.sp_tiles = g.num_tiles; p.act = act; p.stage_bytes = g.stage_bytes; p.stages = 2;
  p.debug = debug_env("LWP_DEBUG_DW");
  // default on (measured: 1.128 -> 1.098 ms over the 14 depthwise launches of a 64-frame step); LWP_DW_REVERSE=0 walks forwards
  p.reverse = (getenv("LWP_DW_REVERSE") != nullptr && atoi(getenv("LWP_DW_REVERSE")) == 0) ? 0 : 1;
  p.tma_out = tm_out != nullptr ? 1 : 0;
  const CUtensorMap &tmo = tm_out != nullptr ? *tm_out : tm;
  if (f32) return depthwise_tma_launch_t<float>(tm, tmo, (float *)out, w9c, scale, shift, p, stride, dil, st);
  return depthwise_tma_launch_t<__nv_bfloat16>(tm, tmo, (__nv_bfloat16 *)out, w9c, scale, shift, p, stride, dil, st);
}

int nhwc_to_nchw_launch(bool in_f32, const void *in, int ld, int c0, int c, float *out, int n, int HW,
                        cudaStream_t st) {
  long long total = (long long)n * c * HW;
  int grid = grid_for(total, 256, 8);
  if (in_f32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float *)in, ld, c0, c, out, n, HW);
  else nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)in, ld, c0, c, out, n, HW);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

}  // namespace lwp
