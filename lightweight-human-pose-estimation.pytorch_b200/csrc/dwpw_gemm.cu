// Fused depthwise-separable block: depthwise 3x3 (stride 1, dilation 1/2) + BN + ReLU/ELU written
// STRAIGHT INTO the shared-memory A operand of the following 1x1 convolution's tcgen05 GEMM.
//
//   reference: conv_dw / conv_dw_no_bn of modules/conv.py:13-32 (two Conv2d + BN + activation each), as used by
//   models/with_mobilenet.py:94-105 (backbone) and :12-16 (Cpm trunk, with the residual of :20 in the last one).
//
// Per 128-pixel tile and per 128-byte K block (64 bf16 / 32 tf32 channels):
//   warp 0      TMA: the input halo box (tile + 2*dil rows/cols, zero-filled outside the image = the depthwise
//               padding), the K block's depthwise constants (9 taps, folded BN scale/shift; a 1-D bulk copy into the
//               same stage -- with the dynamic smem carve-out at its maximum there is no L1 left to serve them from
//               global memory) and the K block of the 1x1 weights
//   warps 6-13  depthwise: 8 channels x 4 columns per thread with a register window over the smem halo tile,
//               fp32 FFMA2 accumulate, scale/shift/activation, round to the plan dtype, store into the K-major
//               SWIZZLE_128B A tile (same bits the unfused depthwise kernel would have written to HBM)
//   warp 14     TMA: the 1x1 weights ring (its own thread: the two rings must not throttle each other)
//   warp 1      one thread issues tcgen05.mma (M = 128, N = Cout up to 512 as 1-2 instructions), fp32 in TMEM
//   warps 2-5   epilogue: TMEM -> scale/shift/activation (+ residual) -> swizzled smem -> TMA tensor store
// The depthwise output never exists in global memory: per block the HBM traffic is input + output only.
#include "common.cuh"
#include "conv_gemm.cuh"
#include "dwpw_gemm.cuh"
#include "gemm_epilogue.cuh"
#include "tcgen05.cuh"

namespace lwp {

constexpr int kDwpwEpiWarps = 4;   // the depthwise warps need the registers: one epilogue warp per TMEM lane quarter here
constexpr int kDwWarp0 = 2 + kDwpwEpiWarps, kDwWarps = 8;   // warps 0/1: TMA / MMA, then epilogue warps, then depthwise warps
constexpr int kDwpwBWarp = kDwWarp0 + kDwWarps;            // last warp: TMA producer of the 1x1 weights
constexpr int kDwpwThreads = (kDwpwBWarp + 1) * 32;        // 480

struct DwpwSmem {
  uint32_t a_off, b_off, staging_off, in_off, scale_off, shift_off, bars_off, total;
};

__host__ __device__ inline DwpwSmem dwpw_smem_layout(const DwpwParams &p) {
  DwpwSmem L;
  L.a_off = 0;
  L.b_off = L.a_off + (uint32_t)p.a_stages * kATileBytes;
  L.staging_off = L.b_off + (uint32_t)p.b_stages * p.b_stage_bytes;
  L.in_off = L.staging_off + (uint32_t)kDwpwEpiWarps * (uint32_t)p.staging_bufs * kStageOutBytes;
  L.scale_off = (L.in_off + (uint32_t)p.in_stages * p.in_stage_bytes + 127u) & ~127u;
  L.shift_off = L.scale_off + (uint32_t)p.cout_pad * 4;
  L.bars_off = L.shift_off + (uint32_t)p.cout_pad * 4;
  L.total = L.bars_off + 40 * 8 + 16;
  return L;
}

size_t dwpw_smem_bytes(const DwpwParams &p) { return (size_t)dwpw_smem_layout(p).total + 1024; }

__device__ __forceinline__ float dw_act(float v, int act) {
  if (act == LWP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == LWP_ACT_ELU) return lwp_elu(v);
  return v;
}

template <bool kTf32> struct TileIn;  // 8 channels of the smem halo tile as four fp32 pairs
// bf16 pair -> fp32 pair with two ALU-pipe instructions (PRMT, LOP3): the compiler's own choice (IMAD.U32 for the
// shift) lands on the FMA pipe, which the FFMA2 stream needs
__device__ __forceinline__ float2 bf16x2_to_f32x2(uint32_t x) {
  return make_float2(__uint_as_float(__byte_perm(x, 0u, 0x1044)), __uint_as_float(x & 0xffff0000u));
}
template <> struct TileIn<false> {
  static __device__ __forceinline__ void load(const uint8_t *p, float2 (&v)[4]) {
    const uint4 raw = *reinterpret_cast<const uint4 *>(p);
    v[0] = bf16x2_to_f32x2(raw.x); v[1] = bf16x2_to_f32x2(raw.y);
    v[2] = bf16x2_to_f32x2(raw.z); v[3] = bf16x2_to_f32x2(raw.w);
  }
};
template <> struct TileIn<true> {
  static __device__ __forceinline__ void load(const uint8_t *p, float2 (&v)[4]) {
    float4 a = reinterpret_cast<const float4 *>(p)[0], b = reinterpret_cast<const float4 *>(p)[1];
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
    v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
  }
};
__device__ __forceinline__ void ldg4pairs(const float *p, float2 (&v)[4]) {  // p: the K block's constants in shared memory
  float4 a = reinterpret_cast<const float4 *>(p)[0], b = reinterpret_cast<const float4 *>(p)[1];
  v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
  v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}

// Depthwise 3x3 of one thread's 8 channels x 4 consecutive output columns of one tile row, for one K block:
// reads the smem halo tile through a register window, accumulates in packed fp32, applies the folded BN +
// activation and writes the four 16-byte (bf16) / 32-byte (fp32) pieces into the swizzled A tile.
// All addressing is 32-bit; the activation is a compile-time choice.
template <bool kTf32, int D, int ACT>
__device__ __forceinline__ void dw_block(const uint8_t *in0, int row_bytes, const float *w0, int cin, uint8_t *abuf,
                                         int r0, int cv) {
  const float *sc0 = w0 + 9 * cin, *sh0 = w0 + 10 * cin;  // (cin = row stride of the constants block = kb_ch)
  constexpr int TWT = 4, NCOL = (TWT - 1) + 2 * D + 1;
  float2 acc[TWT][4];
#pragma unroll
  for (int a = 0; a < TWT; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[a][j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    float2 wk[3][4];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) ldg4pairs(w0 + (ky * 3 + kx) * cin, wk[kx]);
    const uint8_t *rowp = in0 + ky * D * row_bytes;
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      float2 v[4];
      TileIn<kTf32>::load(rowp + ci * kKBlockBytes, v);
#pragma unroll
      for (int a = 0; a < TWT; ++a) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if (a + kx * D == ci) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[a][j] = __ffma2_rn(v[j], wk[kx][j], acc[a][j]);
          }
        }
      }
    }
  }
  float2 sc[4], sh[4];
  ldg4pairs(sc0, sc);
  ldg4pairs(sh0, sh);
#pragma unroll
  for (int a = 0; a < TWT; ++a) {
    const int r = r0 + a;  // row of the A tile == pixel of the tile
    float2 y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      y[j] = __ffma2_rn(acc[a][j], sc[j], sh[j]);
      if constexpr (ACT == LWP_ACT_ELU) {
        y[j].x = lwp_elu(y[j].x);
        y[j].y = lwp_elu(y[j].y);
      } else if constexpr (ACT == LWP_ACT_RELU && kTf32) {
        y[j].x = fmaxf(y[j].x, 0.f);
        y[j].y = fmaxf(y[j].y, 0.f);
      }
    }
    uint8_t *rowa = abuf + r * kKBlockBytes;
    if constexpr (kTf32) {
      *reinterpret_cast<float4 *>(rowa + (((2 * cv) ^ (r & 7)) << 4)) = make_float4(y[0].x, y[0].y, y[1].x, y[1].y);
      *reinterpret_cast<float4 *>(rowa + (((2 * cv + 1) ^ (r & 7)) << 4)) = make_float4(y[2].x, y[2].y, y[3].x, y[3].y);
    } else {
      uint4 pk;
      __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
      const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __float22bfloat162_rn(y[j]);
        if constexpr (ACT == LWP_ACT_RELU) h[j] = __hmax2(h[j], zero2);  // ReLU after rounding == rounding after ReLU
      }
      *reinterpret_cast<uint4 *>(rowa + ((cv ^ (r & 7)) << 4)) = pk;
    }
  }
}

template <bool kTf32, int D>
__global__ void __launch_bounds__(kDwpwThreads, 1)
dwpw_gemm_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const DwpwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const DwpwSmem L = dwpw_smem_layout(p);
  float *s_scale = reinterpret_cast<float *>(smem + L.scale_off);
  float *s_shift = reinterpret_cast<float *>(smem + L.shift_off);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  uint64_t *in_full = bars, *in_empty = bars + 4, *a_full = bars + 8, *a_empty = bars + 12;
  uint64_t *b_full = bars + 16, *b_empty = bars + 20, *tfull = bars + 24, *tempty = bars + 32;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 40);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmIn);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < p.in_stages; ++s) { ptx::mbar_init(&in_full[s], 1); ptx::mbar_init(&in_empty[s], kDwWarps); }
    for (int s = 0; s < p.a_stages; ++s) { ptx::mbar_init(&a_full[s], kDwWarps); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < p.acc_stages; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], kDwpwEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.cout_pad; i += kDwpwThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer: input halo tiles + depthwise constants =====================
    // (its own thread, so that input tiles run ahead of the depthwise warps independently of the weights ring)
    if (lane == 0) {
      int is = 0;
      uint32_t iph = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x) {
        const int img = t / per_img, rem = t - img * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (!ptx::mbar_wait(&in_empty[is], iph ^ 1u)) { ok = false; atomicExch(p.err_flag, 11); break; }
          ptx::mbar_arrive_expect_tx(&in_full[is], p.in_stage_bytes);
          uint8_t *stage = smem + L.in_off + (size_t)is * p.in_stage_bytes;
          ptx::tma_load_4d(stage, &tmIn, &in_full[is], kb * p.kb_ch, x0 - D, y0 - D, img);
          ptx::bulk_load_1d(stage + p.in_tile_bytes, reinterpret_cast<const uint8_t *>(p.dw_consts) + (size_t)kb * p.dw_const_bytes,
                            p.dw_const_bytes, &in_full[is]);
          if (++is == p.in_stages) { is = 0; iph ^= 1u; }
        }
      }
    }
  } else if (warp == kDwpwBWarp) {
    // ===================== TMA producer: 1x1 weights (ring in halves of N, <= 256 rows each) =====================
    if (lane == 0) {
      int bs = 0;
      uint32_t bph = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x) {
        for (int kb = 0; kb < p.kblocks && ok; ++kb) {
          for (int h = 0; h < p.n_mma; ++h) {
            if (!ptx::mbar_wait(&b_empty[bs], bph ^ 1u)) { ok = false; atomicExch(p.err_flag, 12); break; }
            ptx::mbar_arrive_expect_tx(&b_full[bs], p.b_stage_bytes);
            ptx::tma_load_2d(smem + L.b_off + (size_t)bs * p.b_stage_bytes, &tmB, &b_full[bs], kb * p.kb_ch,
                             h * p.n_per_mma);
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_phase = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x) {
        if (!ptx::mbar_wait(&tempty[acc], acc_phase ^ 1u)) { atomicExch(p.err_flag, 13); break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.cout_pad);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (!ptx::mbar_wait(&a_full[as], aph)) { ok = false; atomicExch(p.err_flag, 14); break; }
          const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + L.a_off + (size_t)as * kATileBytes));
          for (int h = 0; h < p.n_mma; ++h) {
            if (!ptx::mbar_wait(&b_full[bs], bph)) { ok = false; atomicExch(p.err_flag, 14); break; }
            ptx::tc_fence_after();
            if (!(LWP_DBG(p.debug) & 4)) {
              const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + L.b_off + (size_t)bs * p.b_stage_bytes));
#pragma unroll
              for (int k = 0; k < kKBlockBytes / 32; ++k)
                ptx::umma<kTf32>(d_tmem + (uint32_t)(h * p.n_per_mma), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k),
                                 p.idesc, (uint32_t)((kb | k) != 0));
            }
            ptx::umma_commit(&b_empty[bs]);
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
          }
          if (!ok) break;
          ptx::umma_commit(&a_empty[as]);
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
        if (!ok) break;
        ptx::umma_commit(&tfull[acc]);
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp < kDwWarp0) {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = row / p.tile_w, tx = row - ty * p.tile_w;
    int acc = 0, sbuf_idx = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
      if (!ptx::mbar_wait(&tfull[acc], acc_phase)) { atomicExch(p.err_flag, 15); break; }
      ptx::tc_fence_after();
      const int img = t / per_img, rem = t - img * per_img;
      const int tyy = rem / p.tiles_x, txx = rem - tyy * p.tiles_x;
      const int x0 = txx * p.tile_w, y0 = tyy * p.tile_h;
      const int y = y0 + ty, x = x0 + tx;
      const bool valid = y < p.H && x < p.W;
      const size_t pix = ((size_t)img * p.H + y) * (size_t)p.W + x;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.cout_pad);
      if (!(LWP_DBG(p.debug) & 2))
        staged_epilogue_tile<kTf32>(&tmC, smem + L.staging_off + (size_t)(warp - 2) * p.staging_bufs * kStageOutBytes,
                                    p.staging_bufs, sbuf_idx, t_row, 0, p.cout_pad, p.n_store, s_scale, s_shift, p.act,
                                    p.residual, p.res_ld, valid, pix, lane, x0 + (q * 32) % p.tile_w,
                                    y0 + (q * 32) / p.tile_w, img, (warp - 2) >> 2, kDwpwEpiWarps / 4, LWP_DBG(p.debug) >> 3);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) ptx::bulk_wait<0>();
  } else {
    // ===================== depthwise producers of the A operand (8 warps) =====================
    constexpr int TWT = 4;
    constexpr int ES = kTf32 ? 4 : 2;
    const int dwtid = threadIdx.x - kDwWarp0 * 32;
    const int cvn = p.kb_ch / 8;                       // 8-channel vectors per pixel in a K block
    const int tasks = cvn * (kBlockM / TWT);
    const bool active = dwtid < tasks;
    const int cv = dwtid % cvn, pt = dwtid / cvn;
    const int xgroups = p.tile_w / TWT;
    const int xg = pt % xgroups, ty = pt / xgroups;
    const int row_bytes = p.iw * kKBlockBytes;
    const int in_off0 = (ty * p.iw + xg * TWT) * kKBlockBytes + cv * 8 * ES;  // this thread's window origin in a stage
    const int r0 = ty * p.tile_w + xg * TWT;
    int is = 0, as = 0;
    uint32_t iph = 0, aph = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x) {
      for (int kb = 0; kb < p.kblocks; ++kb) {
        if (!ptx::mbar_wait(&in_full[is], iph) || !ptx::mbar_wait(&a_empty[as], aph ^ 1u)) {
          ok = false; atomicExch(p.err_flag, 16); break;
        }
        if (active && !(LWP_DBG(p.debug) & 1)) {
          const uint8_t *sbuf = smem + L.in_off + is * (int)p.in_stage_bytes;
          uint8_t *abuf = smem + L.a_off + as * kATileBytes;
          const int c0 = kb * p.kb_ch + cv * 8;
          if (c0 < p.cin) {
            const uint8_t *in0 = sbuf + in_off0;
            const float *wc = reinterpret_cast<const float *>(sbuf + p.in_tile_bytes) + cv * 8;
            if (p.dw_act == LWP_ACT_RELU)
              dw_block<kTf32, D, LWP_ACT_RELU>(in0, row_bytes, wc, p.kb_ch, abuf, r0, cv);
            else if (p.dw_act == LWP_ACT_ELU)
              dw_block<kTf32, D, LWP_ACT_ELU>(in0, row_bytes, wc, p.kb_ch, abuf, r0, cv);
            else
              dw_block<kTf32, D, LWP_ACT_NONE>(in0, row_bytes, wc, p.kb_ch, abuf, r0, cv);
          } else {
            // channels past Cin (a layer thinner than one K block): the GEMM weights there are zero-filled by TMA;
            // write zeros so no NaN garbage can reach the tensor core
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const int r = r0 + a;
              uint8_t *rowa = abuf + r * kKBlockBytes;
              if constexpr (kTf32) {
                *reinterpret_cast<float4 *>(rowa + (((2 * cv) ^ (r & 7)) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4 *>(rowa + (((2 * cv + 1) ^ (r & 7)) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
              } else {
                *reinterpret_cast<uint4 *>(rowa + ((cv ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
        }
        ptx::fence_proxy_async();  // A-tile writes (generic proxy) -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(&a_full[as]);
          ptx::mbar_arrive(&in_empty[is]);
        }
        if (++is == p.in_stages) { is = 0; iph ^= 1u; }
        if (++as == p.a_stages) { as = 0; aph ^= 1u; }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

int dwpw_init() {
  static DeviceOnce once;
  int slot;
  if (!once.pending(&slot)) return LWP_OK;
#define LWP_DWPW_ATTR(TF, D) \
  LWP_CUDA_CHECK(cudaFuncSetAttribute(dwpw_gemm_kernel<TF, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))
  LWP_DWPW_ATTR(false, 1); LWP_DWPW_ATTR(false, 2); LWP_DWPW_ATTR(true, 1); LWP_DWPW_ATTR(true, 2);
#undef LWP_DWPW_ATTR
  once.done[slot] = true;
  return LWP_OK;
}

int dwpw_launch(bool tf32, const CUtensorMap &tmIn, const CUtensorMap &tmB, const CUtensorMap &tmC, const DwpwParams &p,
                int grid, cudaStream_t st) {
  const size_t smem = dwpw_smem_bytes(p);
  if (p.dil == 1) {
    if (tf32) dwpw_gemm_kernel<true, 1><<<grid, kDwpwThreads, smem, st>>>(tmIn, tmB, tmC, p);
    else dwpw_gemm_kernel<false, 1><<<grid, kDwpwThreads, smem, st>>>(tmIn, tmB, tmC, p);
  } else if (p.dil == 2) {
    if (tf32) dwpw_gemm_kernel<true, 2><<<grid, kDwpwThreads, smem, st>>>(tmIn, tmB, tmC, p);
    else dwpw_gemm_kernel<false, 2><<<grid, kDwpwThreads, smem, st>>>(tmIn, tmB, tmC, p);
  } else {
    set_error("dwpw: unsupported dilation %d", p.dil);
    return LWP_EINVAL;
  }
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

}  // namespace lwp
