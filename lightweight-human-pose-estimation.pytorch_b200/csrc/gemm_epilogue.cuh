// Epilogue of the tcgen05 GEMM kernels for plain single-output layers: one epilogue warp drains its 32 TMEM
// lanes (= 32 pixel rows of the tile), applies y = act(acc * scale[c] + shift[c]) (+ residual), stages
// 128 output bytes per row per chunk in SWIZZLE_128B shared memory and lets one lane issue a 4-D TMA
// tensor store of the box (out-of-image rows are clipped by the hardware).
//
// The accumulator is read in units of 32 columns with the NEXT unit's tcgen05.ld already in flight while
// the current unit is converted (tcgen05.wait::ld waits for all outstanding loads, so the next load is
// issued right after the wait and before the math).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace lwp {

constexpr int kStageOutBytes = 32 * 128;  // one warp, one 128-byte column chunk

// one unit = 32 accumulator columns of this thread's row -> 64 (bf16) / 128 (fp32) bytes of the staged row
// bf16 residual of one unit (32 columns = 64 bytes of this thread's pixel row), fetched ahead of the unit's math so its
// latency hides behind the previous unit / the TMEM load (the wait on tcgen05.ld is a compiler barrier: loads
// issued inside the unit could not be hoisted above it)
struct ResPrefetch {
  uint4 v[4];
};
__device__ __forceinline__ void res_prefetch(ResPrefetch &rp, const void *residual, int res_ld, bool res_ok, size_t pix,
                                             int cg0) {
  if (res_ok) {
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(residual) + pix * res_ld + cg0);
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) rp.v[g8] = __ldg(src + g8);
  }
}

template <bool kTf32>
__device__ __forceinline__ void epilogue_unit(const uint32_t (&r)[32], uint8_t *sbuf, int lane, int col_in_chunk, int cg0,
                                              const float *s_scale, const float *s_shift, int act, bool fast_relu,
                                              const void *residual, int res_ld, bool res_ok, size_t pix,
                                              const ResPrefetch *pre = nullptr) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    const int cg = cg0 + g8 * 8;
    // y = acc * scale + shift as packed fp32 FMAs (FFMA2), two channels per instruction
    const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + cg), sc1 = *reinterpret_cast<const float4 *>(s_scale + cg + 4);
    const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + cg), sh1 = *reinterpret_cast<const float4 *>(s_shift + cg + 4);
    float2 a2[4];
    a2[0] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 0]), __uint_as_float(r[g8 * 8 + 1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
    a2[1] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 2]), __uint_as_float(r[g8 * 8 + 3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
    a2[2] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 4]), __uint_as_float(r[g8 * 8 + 5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
    a2[3] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 6]), __uint_as_float(r[g8 * 8 + 7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
    const int cc = col_in_chunk + g8 * 8;  // column inside the 128-byte chunk
    if constexpr (!kTf32) {
      if (fast_relu) {  // bf16, ReLU, no residual: round first, then one packed max per two channels
        uint4 pk;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
        const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __hmax2(__float22bfloat162_rn(a2[j]), zero2);
        *reinterpret_cast<uint4 *>(sbuf + lane * 128 + (((cc >> 3) ^ (lane & 7)) << 4)) = pk;
        continue;
      }
    }
    float v[8] = {a2[0].x, a2[0].y, a2[1].x, a2[1].y, a2[2].x, a2[2].y, a2[3].x, a2[3].y};
    if (act == LWP_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == LWP_ACT_ELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = lwp_elu(v[j]);  // ELU(alpha=1); abs error ~1e-7
    }
    if (res_ok) {
      if constexpr (kTf32) {
        const float4 *rp = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(residual) + pix * res_ld + cg);
        float4 a = __ldg(rp), b = __ldg(rp + 1);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
        v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
      } else {
        const uint4 raw = pre != nullptr ? pre->v[g8]
                                         : __ldg(reinterpret_cast<const uint4 *>(
                                               reinterpret_cast<const __nv_bfloat16 *>(residual) + pix * res_ld + cg));
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h[j]);
          v[2 * j] += f.x; v[2 * j + 1] += f.y;
        }
      }
    }
    // 16-byte pieces of this row's 128-byte line, XOR-swizzled like SWIZZLE_128B expects
    if constexpr (kTf32) {
      const int j0 = cc >> 2;  // two 16-byte pieces
      *reinterpret_cast<float4 *>(sbuf + lane * 128 + (((j0) ^ (lane & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4 *>(sbuf + lane * 128 + (((j0 + 1) ^ (lane & 7)) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint4 pk;
      __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4 *>(sbuf + lane * 128 + (((cc >> 3) ^ (lane & 7)) << 4)) = pk;
    }
  }
}

// `part` of `nparts` epilogue warps share one TMEM lane quarter: each takes every nparts-th 128-byte chunk.
template <bool kTf32>
__device__ __forceinline__ void staged_epilogue_tile(const CUtensorMap *tmC, uint8_t *stage_base, int nbuf, int &sbuf_idx,
                                                     uint32_t t_row, int n0, int block_n, int n_store,
                                                     const float *s_scale, const float *s_shift, int act,
                                                     const void *residual, int res_ld, bool valid, size_t pix, int lane,
                                                     int store_x, int store_y, int img, int part = 0, int nparts = 1,
                                                     int dbg = 0) {
  constexpr int kUnitsPerChunk = kTf32 ? 1 : 2;   // 128 output bytes = 32 fp32 / 64 bf16 columns
  constexpr int kChunkCols = 32 * kUnitsPerChunk;
  const bool fast_relu = act == LWP_ACT_RELU && residual == nullptr;
  const bool res_ok = residual != nullptr && valid;
  int cols = n_store - n0;                          // columns of this tile that are stored (whole chunks)
  if (cols > block_n) cols = block_n;
  if (cols <= 0) return;
  const int chunks = cols / kChunkCols;
  uint32_t ra[32], rb[32];
  ResPrefetch pa, pb;
  if (part < chunks) {
    ptx::tmem_ld_32x32(t_row + (uint32_t)(part * kChunkCols), ra);
    if constexpr (!kTf32) res_prefetch(pa, residual, res_ld, res_ok, pix, n0 + part * kChunkCols);
    ptx::tmem_ld_wait(ra);
  }
  for (int c = part; c < chunks; c += nparts) {
    const int col0 = c * kChunkCols;
    uint8_t *sbuf = stage_base + (size_t)sbuf_idx * kStageOutBytes;
    if (lane == 0) {  // the tensor store that last read this staging buffer must be done reading it
      if (nbuf == 2) ptx::bulk_wait_read<1>(); else ptx::bulk_wait_read<0>();
    }
    __syncwarp();
    const int cn = c + nparts;  // this warp's next chunk
    if constexpr (kUnitsPerChunk == 2) {
      ptx::tmem_ld_32x32(t_row + (uint32_t)(col0 + 32), rb);          // in flight during the first unit's math
      res_prefetch(pb, residual, res_ld, res_ok, pix, n0 + col0 + 32);
      if (!(dbg & 2))
        epilogue_unit<kTf32>(ra, sbuf, lane, 0, n0 + col0, s_scale, s_shift, act, fast_relu, residual, res_ld, res_ok, pix,
                             &pa);
      ptx::tmem_ld_wait(rb);
      if (cn < chunks) {
        ptx::tmem_ld_32x32(t_row + (uint32_t)(cn * kChunkCols), ra);  // next chunk, in flight
        res_prefetch(pa, residual, res_ld, res_ok, pix, n0 + cn * kChunkCols);
      }
      if (!(dbg & 2))
        epilogue_unit<kTf32>(rb, sbuf, lane, 32, n0 + col0 + 32, s_scale, s_shift, act, fast_relu, residual, res_ld, res_ok,
                             pix, &pb);
    } else {
      if (cn < chunks) ptx::tmem_ld_32x32(t_row + (uint32_t)(cn * kChunkCols), rb);
      if (!(dbg & 2))
        epilogue_unit<kTf32>(ra, sbuf, lane, 0, n0 + col0, s_scale, s_shift, act, fast_relu, residual, res_ld, res_ok, pix);
    }
    ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0 && !(dbg & 1)) {  // one TMA tensor store of the warp's 32 x 128-byte box
      ptx::tma_store_4d(tmC, sbuf, n0 + col0, store_x, store_y, img);
      ptx::bulk_commit();
    }
    sbuf_idx = nbuf == 2 ? (sbuf_idx ^ 1) : 0;
    if (cn < chunks) {
      if constexpr (kUnitsPerChunk == 2) {
        ptx::tmem_ld_wait(ra);
      } else {
        ptx::tmem_ld_wait(rb);
#pragma unroll
        for (int i = 0; i < 32; ++i) ra[i] = rb[i];
      }
    }
  }
}

}  // namespace lwp
