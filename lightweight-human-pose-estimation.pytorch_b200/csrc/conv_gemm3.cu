// Dense 3x3 convolution (stride 1, dilation d, Cout = 128) on CTA pairs with COLUMN-STRIP reuse of the activations.
//
//   reference: the nn.Conv2d 3x3 layers of models/with_mobilenet.py:16,28-31,52-54 (+ bias / BatchNorm / ReLU and the
//   residual adds of :20,60 in the epilogue), same semantics as conv_gemm_kernel.
//
// Why another kernel: the tap-by-tap implicit GEMM (conv_gemm.cu) re-fetches a shifted copy of the 128-pixel tile for
// every one of the 9 taps and the whole 128 x 1152 weight matrix for every tile -- 576 KB of L2 -> shared-memory
// traffic per 4608 tensor-pipe cycles, about twice what one SM can pull from L2, so those layers ran at the SM's L2
// read bandwidth, not at the tensor pipe's.  Here
//   * a tile is 8 pixels wide and 16 high, so one 8-pixel row of the tile is exactly one 8-row SWIZZLE_128B atom of the
//     K-major A operand.  For each horizontal tap dx the producer loads ONE strip of 8 x (16 + 2d) pixels; the three
//     vertical taps read it at start addresses dy * d * 1024 bytes (whole atoms: plain descriptors, no base offset).
//     Activation traffic per tile: 3 strips instead of 9 tiles per K block (110 KB instead of 288 KB);
//   * CTA pairs (tcgen05.mma.cta_group::2, M = 256): each CTA loads only half of every weight tile (72 KB per tile
//     instead of 288 KB).
// Together 182 KB per tile per SM instead of 576 KB: the layer becomes tensor-pipe bound.
//
//   warp 0   TMA producer of activation strips        a_full / a_empty ring (leader's a_full counts both CTAs' bytes)
//   warp 10  TMA producer of weight tiles             b_full / b_empty ring
//   warp 1   MMA issuer (leader CTA): for (kb, dx): wait strip; for dy: wait weights, 4 MMAs, release weights; release strip
//   warps 2-9 epilogue (as conv_gemm2_kernel)
// K is accumulated in (kb, dx, dy) order instead of (dy, dx, kb): same products, fp32 accumulate.
//
// kFusePw (bf16): the 1x1 conv that FOLLOWS the 3x3 (the `initial` conv of the next RefinementStageBlock,
// models/with_mobilenet.py:52,57-60) runs in the same kernel as a back-to-back GEMM: the epilogue warps convert the 3x3
// accumulator (BN + ReLU + residual, rounded to bf16 exactly as the stand-alone kernel would store it) into the TMEM A
// operand of a second cta_group::2 MMA (K = 128, N = 128, W2 resident in shared memory, half per CTA), and only the
// 1x1's output is written.  Saves a launch and the 2 x 62 MB round trip of the block output per RefinementStageBlock.
//   TMEM: 2 main accumulator stages (columns 0..255), two A2 buffers (256..383), acc2 (384..511);  warp 11: W2 load +
//   GEMM2 issuer.  The tensor pipe executes in issue order and the main loop runs far ahead, so GEMM2 of tile t completes
//   long after it was issued: the epilogue warps therefore convert tile t (pass 1) FIRST and only then drain the second
//   accumulator of tile t - 1 (pass 2) -- they never wait for a GEMM2 (A2 is double-buffered for the same reason).
#include "common.cuh"
#include "conv_gemm.cuh"
#include "gemm_epilogue.cuh"
#include "tcgen05.cuh"

#include <string.h>

namespace lwp {

constexpr int kStripW = 8, kStripTileH = 16;      // tile = 8 x 16 pixels
constexpr int kC3N = 128;                          // Cout (padded) handled by this kernel
constexpr int kC3BBytes = (kC3N / 2) * kKBlockBytes;   // one CTA's half of a weight tile: 64 rows x 128 B
constexpr int kC3MaxStages = 8;

constexpr int kC3W2Bytes = 2 * kC3BBytes;         // fused 1x1: this CTA's 64 rows of W2, two K blocks
constexpr int kC3A2Col = 256, kC3Acc2Col = 384;   // fused 1x1: TMEM columns of the A operand / accumulator of the second GEMM
constexpr int kC3PwThreads = kGemmThreads + 32;   // + warp 11

struct Smem3Layout {
  uint32_t strip_bytes, a_off, b_off, staging_off, scale_off, shift_off, w2_off, scale2_off, shift2_off, bars_off, total;
};

__host__ __device__ inline Smem3Layout smem3_layout(int dil, int a_stages, int b_stages, int fuse_pw = 0) {
  Smem3Layout L;
  L.strip_bytes = (uint32_t)(kStripW * (kStripTileH + 2 * dil) * kKBlockBytes);   // multiple of 1024
  L.a_off = 0;
  L.b_off = L.strip_bytes * (uint32_t)a_stages;
  L.staging_off = L.b_off + (uint32_t)kC3BBytes * (uint32_t)b_stages;
  L.scale_off = L.staging_off + kStagingBytes;
  L.shift_off = L.scale_off + kC3N * 4;
  L.scale2_off = L.shift_off + kC3N * 4;
  L.shift2_off = L.scale2_off + (fuse_pw ? kC3N * 4 : 0);
  L.w2_off = (L.shift2_off + (fuse_pw ? kC3N * 4 : 0) + 1023u) & ~1023u;
  L.bars_off = fuse_pw ? L.w2_off + kC3W2Bytes : ((L.shift_off + kC3N * 4 + 15u) & ~15u);
  L.total = L.bars_off + (4 * kC3MaxStages + 2 * kMaxAccStages + 8) * 8 + 16;
  return L;
}

size_t conv_gemm3_smem_bytes(const GemmParams &p) {
  return (size_t)smem3_layout(p.dil, p.c3_a_stages, p.c3_b_stages, p.fuse_pw).total + 1024;
}

struct Tile3 {
  int img, y0, x0;
};
// this CTA's M tile of pair tile pt (may lie past the last tile: then img == NIMG, loads are zero-filled, stores clipped)
__device__ __forceinline__ Tile3 decode_tile3(const GemmParams &p, int pt, int rank) {
  Tile3 c;
  const int m_tile = 2 * pt + rank;
  const int per_img = p.tiles_x * p.tiles_y;
  c.img = m_tile / per_img;
  const int rem = m_tile - c.img * per_img;
  const int ty = rem / p.tiles_x;
  c.y0 = ty * kStripTileH;
  c.x0 = (rem - ty * p.tiles_x) * kStripW;
  return c;
}

template <bool kTf32, bool kFusePw>
__global__ void __launch_bounds__(kFusePw ? kC3PwThreads : kGemmBoundThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmW2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int a_stages = p.c3_a_stages, b_stages = p.c3_b_stages;
  const Smem3Layout L = smem3_layout(p.dil, a_stages, b_stages, kFusePw ? 1 : 0);
  float *s_scale = reinterpret_cast<float *>(smem + L.scale_off);
  float *s_shift = reinterpret_cast<float *>(smem + L.shift_off);
  uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  uint64_t *a_empty = a_full + kC3MaxStages, *b_full = a_empty + kC3MaxStages, *b_empty = b_full + kC3MaxStages;
  uint64_t *tfull_bar = b_empty + kC3MaxStages;
  uint64_t *tempty_bar = tfull_bar + kMaxAccStages;
  // fused 1x1: w2_full (leader), a2_full (leader: 16 converter warps), a2_free / acc2_full (multicast commits, both CTAs),
  // acc2_empty (leader: 16 epilogue warps)
  uint64_t *xbar = tempty_bar + kMaxAccStages;
  uint64_t *w2_full = xbar, *a2_full = xbar + 1, *a2_free = xbar + 3, *acc2_full = xbar + 5, *acc2_empty = xbar + 6;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xbar + 8);
  float *s_scale2 = reinterpret_cast<float *>(smem + L.scale2_off), *s_shift2 = reinterpret_cast<float *>(smem + L.shift2_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_pt = (p.m_tiles + 1) / 2;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < a_stages; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < b_stages; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < p.acc_stages; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);                 // multicast commit
      ptx::mbar_init(&tempty_bar[a], 2 * kEpiWarps);    // leader's copy: epilogue warps of both CTAs
    }
    if constexpr (kFusePw) {
      ptx::mbar_init(w2_full, 1);
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(&a2_full[b], 2 * kEpiWarps); ptx::mbar_init(&a2_free[b], 1); }
      ptx::mbar_init(acc2_full, 1); ptx::mbar_init(acc2_empty, 2 * kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair(tmem_slot, p.tmem_cols);
  for (int i = threadIdx.x; i < kC3N; i += kGemmThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
    if constexpr (kFusePw) { s_scale2[i] = p.scale2[i]; s_shift2[i] = p.shift2[i]; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();   // both CTAs' barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t afull0 = ptx::smem_u32(a_full), aempty0 = ptx::smem_u32(a_empty);
  const uint32_t bfull0 = ptx::smem_u32(b_full), bempty0 = ptx::smem_u32(b_empty);
  const int kblocks = p.kblocks_per_tap;

  if (warp == 0) {
    // ===================== TMA producer: activation strips (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0, dst = smem_base + L.a_off;
    bool ok = true;
    for (int pt = cluster_id; pt < num_pt && ok; pt += num_clusters) {
      const Tile3 tc = decode_tile3(p, pt, rank);
      for (int kb = 0; kb < kblocks && ok; ++kb) {
        for (int tx = 0; tx < 3; ++tx) {
          if (!ptx::mbar_wait_u32(aempty0 + 8u * stage, phase ^ 1u)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 31); break; }
          if (ptx::elect_one()) {
            if (leader) ptx::mbar_arrive_expect_tx_u32(afull0 + 8u * stage, 2u * L.strip_bytes);
            ptx::tma_load_4d_pair_u32(dst, &tmA, afull0 + 8u * stage, kb * p.kb_elems, tc.x0 + (tx - 1) * p.dil, tc.y0 - p.dil,
                                      tc.img);
          }
          dst += L.strip_bytes;
          if (++stage == a_stages) { stage = 0; phase ^= 1u; dst = smem_base + L.a_off; }
        }
      }
    }
  } else if (warp == kBProducerWarp) {
    // ===================== TMA producer: this CTA's half of the weight tiles =====================
    const int nb = rank * (kC3N / 2);
    int stage = 0;
    uint32_t phase = 0, dst = smem_base + L.b_off;
    bool ok = true;
    for (int pt = cluster_id; pt < num_pt && ok; pt += num_clusters) {
      for (int kb = 0; kb < kblocks && ok; ++kb) {
        for (int tx = 0; tx < 3 && ok; ++tx) {
          for (int ty = 0; ty < 3; ++ty) {
            if (!ptx::mbar_wait_u32(bempty0 + 8u * stage, phase ^ 1u)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 32); break; }
            if (ptx::elect_one()) {
              if (leader) ptx::mbar_arrive_expect_tx_u32(bfull0 + 8u * stage, 2u * kC3BBytes);
              ptx::tma_load_2d_pair_u32(dst, &tmB, bfull0 + 8u * stage, (ty * 3 + tx) * p.cin + kb * p.kb_elems, nb);
            }
            dst += kC3BBytes;
            if (++stage == b_stages) { stage = 0; phase ^= 1u; dst = smem_base + L.b_off; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint64_t desc_hi = ptx::umma_desc_k_sw128(0);
      const uint32_t a16_0 = ((smem_base + L.a_off) & 0x3FFFFu) >> 4, b16_0 = ((smem_base + L.b_off) & 0x3FFFFu) >> 4;
      const uint32_t strip16 = L.strip_bytes >> 4, btile16 = kC3BBytes >> 4;
      const uint32_t dy16 = (uint32_t)(p.dil * kStripW * kKBlockBytes) >> 4;   // one vertical tap = d rows of 8 pixels
      const uint32_t idesc = p.idesc;
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_phase = 0, a16 = a16_0, b16 = b16_0;
      bool ok = true;
      for (int pt = cluster_id; pt < num_pt && ok; pt += num_clusters) {
        if (!ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u)) { if (lane == 0) atomicExch(p.err_flag, 33); break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kC3N);
        uint32_t first = 0u;   // 0 for the very first MMA of the tile
        for (int kx = 0; kx < 3 * kblocks && ok; ++kx) {   // (kb, dx) pairs: one strip each
          if (!ptx::mbar_wait_u32(afull0 + 8u * as, aph)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 34); break; }
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
            if (!ptx::mbar_wait_u32(bfull0 + 8u * bs, bph)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 35); break; }
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint64_t da = desc_hi | (uint64_t)(a16 + (uint32_t)ty * dy16), db = desc_hi | (uint64_t)b16;
              ptx::umma_pair<kTf32>(d_tmem, da, db, idesc, first);
              ptx::umma_pair<kTf32>(d_tmem, da + 2u, db + 2u, idesc, 1u);
              ptx::umma_pair<kTf32>(d_tmem, da + 4u, db + 4u, idesc, 1u);
              ptx::umma_pair<kTf32>(d_tmem, da + 6u, db + 6u, idesc, 1u);
              ptx::umma_commit_pair_u32(bempty0 + 8u * bs);   // weight stage free in BOTH CTAs
              if (ty == 2) ptx::umma_commit_pair_u32(aempty0 + 8u * as);   // strip free in BOTH CTAs
            }
            first = 1u;
            b16 += btile16;
            if (++bs == b_stages) { bs = 0; bph ^= 1u; b16 = b16_0; }
          }
          a16 += strip16;
          if (++as == a_stages) { as = 0; aph ^= 1u; a16 = a16_0; }
        }
        if (!ok) break;
        if (ptx::elect_one()) ptx::umma_commit_pair(&tfull_bar[acc]);       // accumulators of both CTAs complete
        __syncwarp();
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (kFusePw && warp == kBProducerWarp + 1) {
    // ===================== fused 1x1: W2 load (both CTAs), GEMM2 issuer (leader) =====================
    if (ptx::elect_one()) {
      if (leader) ptx::mbar_arrive_expect_tx(w2_full, 2u * kC3W2Bytes);
      for (int kb = 0; kb < 2; ++kb)
        ptx::tma_load_2d_pair(smem + L.w2_off + kb * kC3BBytes, &tmW2, w2_full, kb * 64, rank * (kC3N / 2));
    }
    __syncwarp();
    if (leader) {
      const uint64_t desc_hi = ptx::umma_desc_k_sw128(0);
      const uint32_t w16 = ((smem_base + L.w2_off) & 0x3FFFFu) >> 4;
      const uint32_t idesc = p.idesc;   // M = 256, N = 128
      int it = 0;
      if (!ptx::mbar_wait(w2_full, 0u)) { if (lane == 0) atomicExch(p.err_flag, 37); }
      for (int pt = cluster_id; pt < num_pt; pt += num_clusters, ++it) {
        const int b = it & 1;
        if (!ptx::mbar_wait(&a2_full[b], (uint32_t)(it >> 1) & 1u) || !ptx::mbar_wait(acc2_empty, ((uint32_t)it & 1u) ^ 1u)) { if (lane == 0) atomicExch(p.err_flag, 38); break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t db = desc_hi | (uint64_t)(w16 + (uint32_t)kb * (kC3BBytes >> 4));
            const uint32_t a = tmem_base + (uint32_t)kC3A2Col + (uint32_t)b * 64u + (uint32_t)kb * 32u;
            ptx::umma_pair_ts(tmem_base + (uint32_t)kC3Acc2Col, a, db, idesc, kb == 0 ? 0u : 1u);
            ptx::umma_pair_ts(tmem_base + (uint32_t)kC3Acc2Col, a + 8u, db + 2u, idesc, 1u);
            ptx::umma_pair_ts(tmem_base + (uint32_t)kC3Acc2Col, a + 16u, db + 4u, idesc, 1u);
            ptx::umma_pair_ts(tmem_base + (uint32_t)kC3Acc2Col, a + 24u, db + 6u, idesc, 1u);
          }
          ptx::umma_commit_pair(&a2_free[b]);  // this A2 buffer may be rewritten (both CTAs)
          ptx::umma_commit_pair(acc2_full);    // the second accumulator is complete (both CTAs)
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2 && warp < 2 + kEpiWarps) {
    // ===================== epilogue (8 warps per CTA, its own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = row / kStripW, tx = row - ty * kStripW;
    int acc = 0, sbuf_idx = 0;
    uint32_t acc_phase = 0;
    int pw_it = 0;
    Tile3 prev_tc = {0, 0, 0};
    bool prev_valid = false;
    size_t prev_pix = 0;
    (void)pw_it; (void)prev_tc; (void)prev_valid; (void)prev_pix;
    for (int pt = cluster_id; pt < num_pt; pt += num_clusters) {
      if (!ptx::mbar_wait(&tfull_bar[acc], acc_phase)) { atomicExch(p.err_flag, 36); break; }
      ptx::tc_fence_after();
      const Tile3 tc = decode_tile3(p, pt, rank);
      const int y = tc.y0 + ty, x = tc.x0 + tx;
      const bool valid = tc.img < p.NIMG && y < p.H && x < p.W;
      const size_t pix = ((size_t)tc.img * p.H + y) * (size_t)p.W + x;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kC3N);
      if constexpr (kFusePw) {
        // ---- pass 1: this warp's 64 columns of the 3x3 result -> BN + activation + residual -> bf16 -> A2[it & 1] (TMEM) ----
        const int part = (warp - 2) >> 2;
        const bool res_ok = p.residual != nullptr && valid;
        const int b2 = pw_it & 1;
        if (!ptx::mbar_wait(&a2_free[b2], (((uint32_t)pw_it >> 1) & 1u) ^ 1u)) { atomicExch(p.err_flag, 39); break; }   // GEMM2 of two tiles ago has read this buffer
        ptx::tc_fence_after();
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          const int col0 = part * 64 + u * 32;
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_row + (uint32_t)col0, r);
          ResPrefetch rp;
          res_prefetch(rp, p.residual, p.res_ld, res_ok, pix, col0);
          ptx::tmem_ld_wait(r);
          uint32_t o[16];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int cg = col0 + g8 * 8;
            const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + cg), sc1 = *reinterpret_cast<const float4 *>(s_scale + cg + 4);
            const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + cg), sh1 = *reinterpret_cast<const float4 *>(s_shift + cg + 4);
            float2 a2[4];
            a2[0] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 0]), __uint_as_float(r[g8 * 8 + 1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
            a2[1] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 2]), __uint_as_float(r[g8 * 8 + 3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
            a2[2] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 4]), __uint_as_float(r[g8 * 8 + 5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
            a2[3] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 6]), __uint_as_float(r[g8 * 8 + 7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
            float v[8] = {a2[0].x, a2[0].y, a2[1].x, a2[1].y, a2[2].x, a2[2].y, a2[3].x, a2[3].y};
            if (p.act == LWP_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
            } else if (p.act == LWP_ACT_ELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = lwp_elu(v[j]);
            }
            if (res_ok) {
              const __nv_bfloat162 *hres = reinterpret_cast<const __nv_bfloat162 *>(&rp.v[g8]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(hres[j]);
                v[2 * j] += f.x; v[2 * j + 1] += f.y;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 hv = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              o[g8 * 4 + j] = *reinterpret_cast<const uint32_t *>(&hv);
            }
          }
          ptx::tmem_st_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kC3A2Col + (uint32_t)b2 * 64u + (uint32_t)(col0 >> 1), o);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_leader(&tempty_bar[acc]);   // the main accumulator stage is free
          ptx::mbar_arrive_leader(&a2_full[b2]);       // this warp's part of the A operand is in place
        }
        // ---- pass 2 (deferred by one tile): the 1x1's accumulator of the PREVIOUS tile -> BN + activation -> TMA store ----
        if (pw_it > 0) {
          if (!ptx::mbar_wait(acc2_full, (uint32_t)(pw_it - 1) & 1u)) { atomicExch(p.err_flag, 40); break; }
          ptx::tc_fence_after();
          staged_epilogue_tile<false>(&tmC, smem + L.staging_off + (size_t)(warp - 2) * kStageOutBytes, 1, sbuf_idx,
                                      tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kC3Acc2Col, 0, kC3N, p.n_store, s_scale2,
                                      s_shift2, p.act2, nullptr, 0, prev_valid, prev_pix, lane, prev_tc.x0,
                                      prev_tc.y0 + q * (32 / kStripW), prev_tc.img, part, kEpiWarps / 4);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_leader(acc2_empty);
        }
        prev_tc = tc; prev_valid = valid; prev_pix = pix;
        ++pw_it;
      } else {
      staged_epilogue_tile<kTf32>(&tmC, smem + L.staging_off + (size_t)(warp - 2) * kStageOutBytes, 1, sbuf_idx, t_row, 0,
                                  kC3N, p.n_store, s_scale, s_shift, p.act, p.residual, p.res_ld, valid, pix, lane,
                                  tc.x0, tc.y0 + q * (32 / kStripW), tc.img, (warp - 2) >> 2, kEpiWarps / 4);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_leader(&tempty_bar[acc]);  // the leader's MMA thread waits for both CTAs
      }
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
    if constexpr (kFusePw) {
      if (pw_it > 0 && ptx::mbar_wait(acc2_full, (uint32_t)(pw_it - 1) & 1u)) {   // the last tile's second accumulator
        ptx::tc_fence_after();
        staged_epilogue_tile<false>(&tmC, smem + L.staging_off + (size_t)(warp - 2) * kStageOutBytes, 1, sbuf_idx,
                                    tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kC3Acc2Col, 0, kC3N, p.n_store, s_scale2,
                                    s_shift2, p.act2, nullptr, 0, prev_valid, prev_pix, lane, prev_tc.x0,
                                    prev_tc.y0 + q * (32 / kStripW), prev_tc.img, (warp - 2) >> 2, kEpiWarps / 4);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_leader(acc2_empty);
      }
    }
    if (lane == 0) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();   // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, p.tmem_cols);
}

int conv_gemm3_init() {
  static DeviceOnce once;
  int slot;
  if (!once.pending(&slot)) return LWP_OK;
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  once.done[slot] = true;
  return LWP_OK;
}

int conv_gemm3_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                      const GemmParams &p, int grid, cudaStream_t st) {
  const size_t smem = conv_gemm3_smem_bytes(p);
  if (tf32) LWP_CUDA_CHECK(launch_pdl(conv3x3_pair_kernel<true, false>, grid, kGemmThreads, smem, st, 2, tmA, tmB, tmC, tmC, p));
  else LWP_CUDA_CHECK(launch_pdl(conv3x3_pair_kernel<false, false>, grid, kGemmThreads, smem, st, 2, tmA, tmB, tmC, tmC, p));
  return LWP_OK;
}

int conv_gemm3_pw_launch(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const CUtensorMap &tmW2,
                         const GemmParams &p, int grid, cudaStream_t st) {
  const size_t smem = conv_gemm3_smem_bytes(p);
  LWP_CUDA_CHECK(launch_pdl(conv3x3_pair_kernel<false, true>, grid, kC3PwThreads, smem, st, 2, tmA, tmB, tmC, tmW2, p));
  return LWP_OK;
}

}  // namespace lwp
