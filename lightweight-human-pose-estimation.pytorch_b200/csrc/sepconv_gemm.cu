// Weight-resident fused depthwise-separable block for the THIN layers of the network (Cin * Cout small enough
// for the whole 1x1 weight matrix to live in shared memory for the CTA's life):
//
//   reference: conv_dw / conv_dw_no_bn of modules/conv.py:13-32 -- depthwise 3x3 (stride 1 or 2, pad 1) + BN + ReLU
//   (or ELU) followed by 1x1 conv + BN + ReLU (or ELU), as used by models/with_mobilenet.py:94-99 (backbone blocks
//   2, 3, 5) and :12-16 (the three Cpm trunk blocks, with the residual of :20 in the last one).
//
// On these layers the two-kernel form is bound by the HBM round trip of the depthwise output (written once, read
// once; arithmetic intensity of the 1x1 21-124 FLOP/B).  Here the depthwise result only ever exists as the
// shared-memory A operand of the tcgen05 GEMM.  What the first fused kernel (dwpw_gemm.cu) got wrong for these
// layers is fixed by construction:
//   * the 1x1 weights (<= 128 KB) and the depthwise constants are loaded ONCE per CTA, not per 128-pixel tile;
//   * 16 depthwise warps (the CUDA-core stencil, not the tensor pipe, is the critical resource), working on two
//     (bf16) or four (tf32) K blocks concurrently, each thread 4 channels x (2 rows x 4 columns) with a register
//     window over the halo box (24 LDS.64 per 32 outputs at stride 1) -- the mapping of the stand-alone kernel;
//   * stride 2 is supported (backbone block 2), so the block's input is read once at full resolution and only
//     the quarter-size output is written.
// Roles (18 warps = 576 threads -> 96 registers per thread): warp 0 TMA producer (weights once, then one halo box
// per (tile, K block)), warp 1 MMA issuer + TMEM owner, warps 2-17 COMPUTE warps.  Both CUDA-core jobs of this kernel
// -- the depthwise stencil that produces the A operand and the epilogue that drains the accumulator -- run on the
// same 16 warps, so they balance by construction (a first version with 4 dedicated epilogue warps was bound by them:
// ncu showed the depthwise warps waiting for a free A stage behind an MMA issuer that waited for a free accumulator).
// A "work item" is one K block of one tile, numbered consecutively over the CTA's tiles; item i is produced by
// depthwise group i % G (G = 2 groups of 8 warps for bf16, 4 groups of 4 warps for tf32), uses halo stage
// i % in_stages and A stage i % a_stages.  After each of its items a warp catches up on the epilogue of all EARLIER
// tiles: for its TMEM lane quarter (warp % 4) it owns one quarter of the output columns ((warp / 4) % 4), which it
// reads with tcgen05.ld, scales / shifts / activates (+ residual), stages in swizzled shared memory and writes with
// one TMA tensor store per tile.
#include "common.cuh"
#include "conv_gemm.cuh"
#include "gemm_epilogue.cuh"
#include "sepconv_gemm.cuh"
#include "tcgen05.cuh"

namespace lwp {

// outputs per depthwise thread: 4 channels x (kSepR rows x kSepC columns).  2 x 2 (16 accumulator registers, 16 LDS.64 per
// 16 outputs at stride 1) instead of the stand-alone kernel's 2 x 4: the item then needs ~55 registers, which leaves room
// for the kernel's loop-carried state -- with 2 x 4 a dozen values were spilled to local memory around every item, and
// with 227 KB of shared memory there is no L1 left to catch them.
constexpr int kSepR = 2, kSepC = 2, kSepPixBlocks = 128 / (kSepR * kSepC);
constexpr int kSepDwWarp0 = 2;
constexpr int kSepDwWarps = 16;
constexpr int kSepDwThreads = kSepDwWarps * 32;
constexpr int kSepThreads = (kSepDwWarp0 + kSepDwWarps) * 32;   // 576
template <int N> __device__ __forceinline__ void sep_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void sep_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int kSepMaxStages = 8;

struct SepSmem {
  uint32_t w_off, a_off, staging_off, in_off, const_off, scale_off, shift_off, bars_off, total;
};

__host__ __device__ inline SepSmem sep_smem_layout(const SepParams &p) {
  SepSmem L;
  L.w_off = 0;
  L.a_off = L.w_off + (uint32_t)p.kblocks * (uint32_t)p.cout_pad * kKBlockBytes;
  L.staging_off = L.a_off + (uint32_t)p.a_stages * kATileBytes;
  L.in_off = L.staging_off + (uint32_t)kSepDwWarps * 32u * (uint32_t)p.slice_bytes;   // one staging buffer per warp
  L.const_off = L.in_off + (uint32_t)p.in_stages * p.in_stage_bytes;
  L.scale_off = L.const_off + (uint32_t)p.kblocks * 11u * (uint32_t)p.kb_ch * 4u;
  L.shift_off = L.scale_off + (uint32_t)p.cout_pad * 4u;
  L.bars_off = (L.shift_off + (uint32_t)p.cout_pad * 4u + 15u) & ~15u;
  L.total = L.bars_off + (4 * kSepMaxStages + 2 * 4 + 1) * 8 + 16;
  return L;
}

size_t sepconv_smem_bytes(const SepParams &p) { return (size_t)sep_smem_layout(p).total + 1024; }

__device__ __forceinline__ float2 sep_bf16x2_to_f32x2(uint32_t x) {   // PRMT + LOP3 (ALU pipe): the FMA pipe is for the FFMA2s
  return make_float2(__uint_as_float(__byte_perm(x, 0u, 0x1044)), __uint_as_float(x & 0xffff0000u));
}

template <bool kTf32>
__device__ __forceinline__ void sep_load4(uint32_t saddr, float2 (&v)[2]) {   // 4 channels of one halo pixel
  if constexpr (kTf32) {
    float4 a;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(saddr));
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
  } else {
    uint2 raw;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "r"(saddr));
    v[0] = sep_bf16x2_to_f32x2(raw.x); v[1] = sep_bf16x2_to_f32x2(raw.y);
  }
}

// One work item of one depthwise thread: 4 channels x (kSepR rows x kSepC columns) of the tile's depthwise output for
// one K block, accumulated in packed fp32 in tap order (ky-major: the same products and order as the stand-alone
// kernel, so the same bits), folded BN + activation, rounded to the plan dtype and stored into the K-major
// SWIZZLE_128B A tile.  cst: this thread's 4 channels of the K block's constants [9 taps | scale | shift][kb_ch];
// a_off[r][c]: byte offset of output pixel (r, c) of this thread inside an A tile (swizzle included; the same for
// every item, so it is computed once outside the item loop).
template <bool kTf32, int S, int ACT>
__device__ __forceinline__ void sep_dw_item(uint32_t win, uint32_t row_bytes, const float *cst, uint32_t abuf_s,
                                            const uint32_t (&a_off)[kSepR][kSepC]) {
  constexpr int R = kSepR, CC = kSepC, NROW = (R - 1) * S + 3, NCOL = (CC - 1) * S + 3;
  constexpr int KB = kTf32 ? 32 : 64;   // channels per K block: constant offsets into the constants block
  float2 wk[9][2];   // tap row ky is loaded when the first input row that uses it arrives
  float2 acc[R][CC][2];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < CC; ++c) acc[r][c][0] = acc[r][c][1] = make_float2(0.f, 0.f);
#pragma unroll
  for (int iy = 0; iy < NROW; ++iy) {
    const uint32_t rowp = win + (uint32_t)iy * row_bytes;
    if (iy < 3) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 a = *reinterpret_cast<const float4 *>(cst + (iy * 3 + kx) * KB);
        wk[iy * 3 + kx][0] = make_float2(a.x, a.y); wk[iy * 3 + kx][1] = make_float2(a.z, a.w);
      }
    }
#pragma unroll
    for (int ic = 0; ic < NCOL; ++ic) {
      float2 v[2];
      sep_load4<kTf32>(rowp + ic * kKBlockBytes, v);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
          if (r * S + ky == iy) {
#pragma unroll
            for (int c = 0; c < CC; ++c)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
                if (c * S + kx == ic) {
                  acc[r][c][0] = __ffma2_rn(v[0], wk[ky * 3 + kx][0], acc[r][c][0]);
                  acc[r][c][1] = __ffma2_rn(v[1], wk[ky * 3 + kx][1], acc[r][c][1]);
                }
          }
    }
  }
  const float4 s4 = *reinterpret_cast<const float4 *>(cst + 9 * KB), b4 = *reinterpret_cast<const float4 *>(cst + 10 * KB);
  const float2 sc0 = make_float2(s4.x, s4.y), sc1 = make_float2(s4.z, s4.w);
  const float2 sh0 = make_float2(b4.x, b4.y), sh1 = make_float2(b4.z, b4.w);
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      float2 y0 = __ffma2_rn(acc[r][c][0], sc0, sh0), y1 = __ffma2_rn(acc[r][c][1], sc1, sh1);
      if constexpr (ACT == LWP_ACT_ELU) {
        y0.x = lwp_elu(y0.x); y0.y = lwp_elu(y0.y);
        y1.x = lwp_elu(y1.x); y1.y = lwp_elu(y1.y);
      }
      const uint32_t dst = abuf_s + a_off[r][c];
      if constexpr (kTf32) {
        if constexpr (ACT == LWP_ACT_RELU) { y0.x = fmaxf(y0.x, 0.f); y0.y = fmaxf(y0.y, 0.f); y1.x = fmaxf(y1.x, 0.f); y1.y = fmaxf(y1.y, 0.f); }
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(y0.x), "f"(y0.y), "f"(y1.x), "f"(y1.y) : "memory");
      } else {
        __nv_bfloat162 lo = __float22bfloat162_rn(y0), hi = __float22bfloat162_rn(y1);
        if constexpr (ACT == LWP_ACT_RELU) {   // ReLU after rounding == rounding after ReLU
          const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
          lo = __hmax2(lo, zero2); hi = __hmax2(hi, zero2);
        }
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(*reinterpret_cast<uint32_t *>(&lo)),
                     "r"(*reinterpret_cast<uint32_t *>(&hi)) : "memory");
      }
    }
}

// Epilogue share of one warp for one tile: its 32 TMEM lanes (= 32 pixel rows) x its quarter of the output columns
// (`units` units of 16 columns), with the next unit's tcgen05.ld and residual fetch in flight during the current
// unit's math.  The 32 x slice_bytes box is staged in shared memory with the TMA swizzle of that row pitch and
// written by one tensor store.
template <bool kTf32>
__device__ __forceinline__ void sep_unit16(const uint32_t (&r)[16], uint8_t *srow /* this lane's staged row */, uint32_t swz,
                                           int c16 /* first 16-byte chunk of the unit inside the row */, int cg0,
                                           const float *s_scale, const float *s_shift, int act, bool res_ok,
                                           const uint4 (&res)[kTf32 ? 4 : 2]) {
#pragma unroll
  for (int g8 = 0; g8 < 2; ++g8) {
    const int cg = cg0 + g8 * 8;
    const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + cg), sc1 = *reinterpret_cast<const float4 *>(s_scale + cg + 4);
    const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + cg), sh1 = *reinterpret_cast<const float4 *>(s_shift + cg + 4);
    float2 a2[4];
    a2[0] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 0]), __uint_as_float(r[g8 * 8 + 1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
    a2[1] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 2]), __uint_as_float(r[g8 * 8 + 3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
    a2[2] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 4]), __uint_as_float(r[g8 * 8 + 5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
    a2[3] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 6]), __uint_as_float(r[g8 * 8 + 7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
    if constexpr (!kTf32) {
      if (act == LWP_ACT_RELU && !res_ok) {   // round first, then one packed max per two channels
        uint4 pk;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
        const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __hmax2(__float22bfloat162_rn(a2[j]), zero2);
        *reinterpret_cast<uint4 *>(srow + (((uint32_t)(c16 + g8) ^ swz) << 4)) = pk;
        continue;
      }
    }
    float v[8] = {a2[0].x, a2[0].y, a2[1].x, a2[1].y, a2[2].x, a2[2].y, a2[3].x, a2[3].y};
    if (act == LWP_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == LWP_ACT_ELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = lwp_elu(v[j]);
    }
    if constexpr (kTf32) {
      if (res_ok) {
        const float4 a = *reinterpret_cast<const float4 *>(&res[g8 * 2]), b = *reinterpret_cast<const float4 *>(&res[g8 * 2 + 1]);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
      }
      *reinterpret_cast<float4 *>(srow + (((uint32_t)(c16 + 2 * g8) ^ swz) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4 *>(srow + (((uint32_t)(c16 + 2 * g8 + 1) ^ swz) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      if (res_ok) {
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&res[g8]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
      }
      uint4 pk;
      __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4 *>(srow + (((uint32_t)(c16 + g8) ^ swz) << 4)) = pk;
    }
  }
}

template <bool kTf32>
__device__ __forceinline__ void sep_res_fetch(uint4 (&res)[kTf32 ? 4 : 2], const void *residual, int res_ld, bool res_ok, size_t pix, int cg0) {
  if (res_ok) {
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(residual) + (pix * res_ld + cg0) * (kTf32 ? 4 : 2));
#pragma unroll
    for (int i = 0; i < (kTf32 ? 4 : 2); ++i) res[i] = __ldg(src + i);
  }
}

template <bool kTf32>
__device__ __forceinline__ void sep_epilogue_share(const CUtensorMap *tmC, uint8_t *sbuf /* this warp's free staging buffer */,
                                                   uint32_t swz, int slice_bytes, uint32_t t_col0 /* TMEM address of the slice */,
                                                   int col0, int units, const float *s_scale, const float *s_shift, int act,
                                                   const void *residual, int res_ld, bool valid, size_t pix, int lane,
                                                   int store_x, int store_y, int img) {
  constexpr int kChunksPerUnit = kTf32 ? 4 : 2;   // 16 columns = 64 / 32 bytes
  const bool res_ok = residual != nullptr && valid;
  uint8_t *srow = sbuf + lane * slice_bytes;
  uint32_t ra[16], rb[16];
  uint4 pa[kTf32 ? 4 : 2], pb[kTf32 ? 4 : 2];
  ptx::tmem_ld_32x16(t_col0, ra);
  sep_res_fetch<kTf32>(pa, residual, res_ld, res_ok, pix, col0);
  for (int u = 0; u < units; u += 2) {
    ptx::tmem_ld_wait(ra);
    if (u + 1 < units) {
      ptx::tmem_ld_32x16(t_col0 + (uint32_t)((u + 1) << 4), rb);
      sep_res_fetch<kTf32>(pb, residual, res_ld, res_ok, pix, col0 + ((u + 1) << 4));
    }
    sep_unit16<kTf32>(ra, srow, swz, u * kChunksPerUnit, col0 + (u << 4), s_scale, s_shift, act, res_ok, pa);
    if (u + 1 < units) {
      ptx::tmem_ld_wait(rb);
      if (u + 2 < units) {
        ptx::tmem_ld_32x16(t_col0 + (uint32_t)((u + 2) << 4), ra);
        sep_res_fetch<kTf32>(pa, residual, res_ld, res_ok, pix, col0 + ((u + 2) << 4));
      }
      sep_unit16<kTf32>(rb, srow, swz, (u + 1) * kChunksPerUnit, col0 + ((u + 1) << 4), s_scale, s_shift, act, res_ok, pb);
    }
  }
  ptx::fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_4d(tmC, sbuf, col0, store_x, store_y, img);
    ptx::bulk_commit();
  }
}

struct SepTileCursor {   // (image, tile row, tile column) advanced by the grid stride without divisions
  int img, ty, tx, dimg, dty, dtx;
  __device__ __forceinline__ void init(const SepParams &p, int t, int stride) {
    const int per_img = p.tiles_x * p.tiles_y;
    img = t / per_img; int rem = t - img * per_img; ty = rem / p.tiles_x; tx = rem - ty * p.tiles_x;
    dimg = stride / per_img; rem = stride - dimg * per_img; dty = rem / p.tiles_x; dtx = rem - dty * p.tiles_x;
  }
  __device__ __forceinline__ void advance(const SepParams &p) {
    tx += dtx;
    if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    ty += dty;
    if (ty >= p.tiles_y) { ty -= p.tiles_y; ++img; }
    img += dimg;
  }
};

template <bool kTf32, int S>
__global__ void __launch_bounds__(kSepThreads, 1)
sepconv_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const SepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const SepSmem L = sep_smem_layout(p);
  float *s_const = reinterpret_cast<float *>(smem + L.const_off);
  float *s_scale = reinterpret_cast<float *>(smem + L.scale_off);
  float *s_shift = reinterpret_cast<float *>(smem + L.shift_off);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  uint64_t *in_full = bars, *in_empty = bars + kSepMaxStages, *a_full = bars + 2 * kSepMaxStages, *a_empty = bars + 3 * kSepMaxStages;
  uint64_t *tfull = bars + 4 * kSepMaxStages, *tempty = tfull + 4, *w_full = tempty + 4;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cqn = p.kb_ch >> 2;                 // 4-channel groups per K block: 16 (bf16) / 8 (tf32)
  const int tasks = cqn * kSepPixBlocks;        // depthwise threads per work item: 512 (bf16) / 256 (tf32)
  const int groups = kSepDwThreads / tasks;     // 1 / 2 items in flight
  const int warps_per_group = tasks >> 5;
  const uint32_t w_block_bytes = (uint32_t)p.cout_pad * kKBlockBytes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmIn);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < p.in_stages; ++s) { ptx::mbar_init(&in_full[s], 1); ptx::mbar_init(&in_empty[s], warps_per_group); }
    for (int s = 0; s < p.a_stages; ++s) { ptx::mbar_init(&a_full[s], warps_per_group); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.acc_stages; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], kSepDwWarps); }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.cout_pad; i += kSepThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  for (int i = threadIdx.x; i < p.kblocks * 11 * p.kb_ch; i += kSepThreads) s_const[i] = p.dw_consts[i];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {   // the whole 1x1 weight matrix, once (constants: may precede the dependency wait)
    ptx::mbar_arrive_expect_tx(w_full, (uint32_t)p.kblocks * w_block_bytes);
    for (int kb = 0; kb < p.kblocks; ++kb)
      ptx::tma_load_2d(smem + L.w_off + (size_t)kb * w_block_bytes, &tmB, w_full, kb * p.kb_ch, 0);
  }
  pdl_trigger();
  pdl_wait();   // activations of the previous kernel from here on

  const int my_tiles = blockIdx.x < p.m_tiles ? (p.m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ===================== TMA producer: one halo box per (tile, K block) =====================
    if (lane == 0) {
      int is = 0;
      uint32_t iph = 0;
      SepTileCursor cur;
      cur.init(p, blockIdx.x, gridDim.x);
      for (int t = 0; t < my_tiles; ++t, cur.advance(p)) {
        const int x0 = cur.tx * p.tile_w * S - 1, y0 = cur.ty * p.tile_h * S - 1;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (!ptx::mbar_wait(&in_empty[is], iph ^ 1u)) { atomicExch(p.err_flag, 51); break; }
          ptx::mbar_arrive_expect_tx(&in_full[is], p.in_stage_bytes);
          ptx::tma_load_4d(smem + L.in_off + (size_t)is * p.in_stage_bytes, &tmIn, &in_full[is], kb * p.kb_ch, x0, y0, cur.img);
          if (++is == p.in_stages) { is = 0; iph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int as = 0, acc = 0;
      uint32_t aph = 0, acc_phase = 0;
      bool ok = ptx::mbar_wait(w_full, 0);
      const uint32_t w_addr = ptx::smem_u32(smem + L.w_off), a_addr = ptx::smem_u32(smem + L.a_off);
      for (int t = 0; t < my_tiles && ok; ++t) {
        if (!ptx::mbar_wait(&tempty[acc], acc_phase ^ 1u)) { atomicExch(p.err_flag, 53); break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.cout_pad);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (!ptx::mbar_wait(&a_full[as], aph)) { atomicExch(p.err_flag, 54); ok = false; break; }
          ptx::tc_fence_after();
          const uint64_t da = ptx::umma_desc_k_sw128(a_addr + (uint32_t)as * kATileBytes);
          const uint64_t db = ptx::umma_desc_k_sw128(w_addr + (uint32_t)kb * w_block_bytes);
#pragma unroll
          for (int k = 0; k < kKBlockBytes / 32; ++k)
            ptx::umma<kTf32>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (uint32_t)((kb | k) != 0));
          ptx::umma_commit(&a_empty[as]);
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
        ptx::umma_commit(&tfull[acc]);
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== compute warps: depthwise producers of the A operand + epilogue shares ===============
    // (index arithmetic is kept off the per-item path: thread-dependent offsets are computed once -- all divisors are
    // powers of two --, ring positions and the epilogue's tile cursor advance incrementally; with runtime divisions per
    // item the kernel was instruction-bound: 600 instead of ~250 instructions per item.)
    constexpr int ES = kTf32 ? 4 : 2;
    const int total_items = my_tiles * p.kblocks;
    const int cw = warp - kSepDwWarp0;
    const int dwtid = (int)threadIdx.x - kSepDwWarp0 * 32;
    const int tasks_log2 = kTf32 ? 8 : 9, cqn_log2 = kTf32 ? 3 : 4;      // tasks = cqn * 32 pixel blocks
    const int g = dwtid >> tasks_log2, local = dwtid & (tasks - 1);
    const int cq = local & (cqn - 1), pb = local >> cqn_log2;             // 4-channel group, 2x2 pixel block (0..31)
    const int xg_log2 = 31 - __clz(p.tile_w / kSepC);
    const int xg = pb & ((1 << xg_log2) - 1), yg = pb >> xg_log2;
    const uint32_t win_off = (uint32_t)((yg * kSepR * S) * p.iw + xg * kSepC * S) * kKBlockBytes + (uint32_t)(cq * 4 * ES);
    uint32_t a_off[kSepR][kSepC];   // byte offset of each of this thread's output pixels inside an A tile (swizzled)
#pragma unroll
    for (int r = 0; r < kSepR; ++r)
#pragma unroll
      for (int c = 0; c < kSepC; ++c) {
        const int rr = (yg * kSepR + r) * p.tile_w + xg * kSepC + c;   // row of the A tile == pixel of the tile
        a_off[r][c] = (uint32_t)rr * kKBlockBytes +
                      (kTf32 ? (uint32_t)((cq ^ (rr & 7)) << 4) : (uint32_t)((((cq >> 1) ^ (rr & 7)) << 4) + ((cq & 1) << 3)));
      }
    // epilogue share: TMEM lane quarter q (hardware: warp id % 4), column slice sl; everything that does not depend on the
    // tile is computed here, once
    const int q = warp & 3, sl = cw >> 2;
    const int tw_log2 = 31 - __clz(p.tile_w);
    const int slice_cols = p.n_store >> 2, col0 = sl * slice_cols, e_units = slice_cols >> 4;
    const uint32_t swz = p.slice_bytes == 128 ? (uint32_t)(lane & 7) : p.slice_bytes == 64 ? (uint32_t)((lane >> 1) & 3)
                                                                    : p.slice_bytes == 32 ? (uint32_t)((lane >> 2) & 1) : 0u;
    uint8_t *e_stage = smem + L.staging_off + (size_t)cw * 32 * p.slice_bytes;
    const uint32_t e_tmem = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
    const int e_dx = (q * 32) & (p.tile_w - 1), e_dy = (q * 32) >> tw_log2;     // origin of this warp's 32 rows inside the tile
    const int e_px = (q * 32 + lane) & (p.tile_w - 1), e_py = (q * 32 + lane) >> tw_log2;   // this lane's pixel inside the tile
    const bool has_res = p.residual != nullptr;
    SepTileCursor ecur;
    ecur.init(p, blockIdx.x, gridDim.x);
    int epi_done = 0, eacc = 0;
    uint32_t eacc_phase = 0;
    auto epilogue_next = [&]() -> bool {   // this warp's share of tile number `epi_done` of this CTA
      if (!ptx::mbar_wait(&tfull[eacc], eacc_phase)) { atomicExch(p.err_flag, 55); return false; }
      ptx::tc_fence_after();
      const int x0 = ecur.tx << tw_log2, y0 = ecur.ty * p.tile_h;
      bool valid = false;
      size_t pix = 0;
      if (has_res) {
        const int y = y0 + e_py, x = x0 + e_px;
        valid = y < p.Ho && x < p.Wo;
        pix = ((size_t)ecur.img * p.Ho + y) * (size_t)p.Wo + x;
      }
      if (lane == 0) ptx::bulk_wait_read<0>();   // the previous tile's store (issued a whole tile ago) is done reading the staging buffer
      __syncwarp();
      sep_epilogue_share<kTf32>(&tmC, e_stage, swz, p.slice_bytes, e_tmem + (uint32_t)(eacc * p.cout_pad), col0, e_units, s_scale,
                                s_shift, p.act, p.residual, p.res_ld, valid, pix, lane, x0 + e_dx, y0 + e_dy, ecur.img);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[eacc]);
      if (++eacc == p.acc_stages) { eacc = 0; eacc_phase ^= 1u; }
      ecur.advance(p);
      ++epi_done;
      return true;
    };
    const uint32_t in_base = ptx::smem_u32(smem + L.in_off) + win_off, a_base = ptx::smem_u32(smem + L.a_off);
    const uint32_t row_bytes = (uint32_t)p.iw * kKBlockBytes;
    int is = g % p.in_stages, as = g % p.a_stages, kb = g % p.kblocks, tile_of_item = g / p.kblocks;
    uint32_t iph = (uint32_t)(g / p.in_stages) & 1u, aph = (uint32_t)(g / p.a_stages) & 1u;
    bool ok = true;
    for (int i = g; i < total_items && ok; i += groups) {
      if (!ptx::mbar_wait(&in_full[is], iph) || !ptx::mbar_wait(&a_empty[as], aph ^ 1u)) { atomicExch(p.err_flag, 56); break; }
      const uint32_t win = in_base + (uint32_t)is * p.in_stage_bytes;
      const uint32_t abuf_s = a_base + (uint32_t)as * kATileBytes;
      const float *cst = s_const + kb * 11 * (kTf32 ? 32 : 64) + cq * 4;
      if (p.dw_act == LWP_ACT_RELU) sep_dw_item<kTf32, S, LWP_ACT_RELU>(win, row_bytes, cst, abuf_s, a_off);
      else if (p.dw_act == LWP_ACT_ELU) sep_dw_item<kTf32, S, LWP_ACT_ELU>(win, row_bytes, cst, abuf_s, a_off);
      else sep_dw_item<kTf32, S, LWP_ACT_NONE>(win, row_bytes, cst, abuf_s, a_off);
      ptx::fence_proxy_async();  // A-tile writes (generic proxy) -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&a_full[as]);
        ptx::mbar_arrive(&in_empty[is]);
      }
      while (epi_done < tile_of_item && ok) ok = epilogue_next();   // catch up on the epilogue of all earlier tiles
      for (int s = 0; s < groups; ++s) {   // advance the ring / K block / tile positions by `groups` items
        if (++is == p.in_stages) { is = 0; iph ^= 1u; }
        if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        if (++kb == p.kblocks) { kb = 0; ++tile_of_item; }
      }
    }
    while (epi_done < my_tiles && ok) ok = epilogue_next();
    if (lane == 0) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

int sepconv_init() {
  static DeviceOnce once;
  int slot;
  if (!once.pending(&slot)) return LWP_OK;
#define LWP_SEP_ATTR(TF, S) \
  LWP_CUDA_CHECK(cudaFuncSetAttribute(sepconv_kernel<TF, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))
  LWP_SEP_ATTR(false, 1); LWP_SEP_ATTR(false, 2); LWP_SEP_ATTR(true, 1); LWP_SEP_ATTR(true, 2);
#undef LWP_SEP_ATTR
  once.done[slot] = true;
  return LWP_OK;
}

int sepconv_launch(bool tf32, const CUtensorMap &tmIn, const CUtensorMap &tmB, const CUtensorMap &tmC, const SepParams &p,
                   int grid, cudaStream_t st) {
  const size_t smem = sepconv_smem_bytes(p);
  if (p.stride == 1) {
    if (tf32) LWP_CUDA_CHECK(launch_pdl(sepconv_kernel<true, 1>, grid, kSepThreads, smem, st, 1, tmIn, tmB, tmC, p));
    else LWP_CUDA_CHECK(launch_pdl(sepconv_kernel<false, 1>, grid, kSepThreads, smem, st, 1, tmIn, tmB, tmC, p));
  } else if (p.stride == 2) {
    if (tf32) LWP_CUDA_CHECK(launch_pdl(sepconv_kernel<true, 2>, grid, kSepThreads, smem, st, 1, tmIn, tmB, tmC, p));
    else LWP_CUDA_CHECK(launch_pdl(sepconv_kernel<false, 2>, grid, kSepThreads, smem, st, 1, tmIn, tmB, tmC, p));
  } else {
    set_error("sepconv: unsupported stride %d", p.stride);
    return LWP_EINVAL;
  }
  return LWP_OK;
}

}  // namespace lwp
