// Thin 1x1 convolutions (K <= 2 K blocks, Cout_pad <= 128) as a WEIGHT-RESIDENT tcgen05 GEMM with several M tiles
// per pipeline stage.
//
// conv_gemm_kernel treats every 128-pixel tile as one pipeline transaction: two producer arrivals, one MMA issue, one
// accumulator hand-off -- and re-fetches the (tiny) weight tile from L2 for every tile.  On the full-resolution
// 32 -> 64 layer (30 176 tiles of 8 KB, 204 per SM) those per-tile round trips, not HBM, set the time: with the
// epilogue skipped the load / MMA side alone took 126 us of the layer's 161 us, 108 us with the activation loads
// skipped as well (timing experiments).  Here
//   * the whole weight matrix (<= 32 KB) is loaded ONCE per CTA, before griddepcontrol.wait (weights do not depend
//     on the previous kernel), and stays in shared memory;
//   * a pipeline stage is a SUPER-TILE of S = 256 / N consecutive 128-pixel tiles (S * nkb TMA boxes on one
//     mbarrier), the MMA lane issues all S * nkb * (kb_bytes / 32) instructions behind one barrier wait into S
//     accumulator blocks of one 256-column TMEM stage (two stages), and the epilogue drains a whole super-tile per
//     hand-off: every barrier round trip is amortised over S tiles.
// Operand layouts, tensor maps, instruction descriptor and the TMA-store epilogue are those of conv_gemm_kernel.
//
// Replaces the same reference layers as conv_gemm.cu: the 1x1 half of conv_dw / conv_dw_no_bn (modules/conv.py:13-32)
// at 32 -> 64, 64 -> 128, 128 -> 128 channels and Cpm's 1x1s (models/with_mobilenet.py:10-21).
#include "common.cuh"
#include "conv_gemm.cuh"
#include "tcgen05.cuh"
#include "gemm_epilogue.cuh"

namespace lwp {

namespace {
constexpr int kWarpProd = kEpiWarps, kWarpMma = kEpiWarps + 2;   // warp kEpiWarps + 1 has no role (thread count as conv_gemm_kernel)
constexpr int kAccCols = 256;                                   // one TMEM accumulator stage = S * block_n columns

struct WresLayout {
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t b_off, staging_off, scale_off, shift_off, bars_off, total;
};
__host__ __device__ inline WresLayout wres_layout(const GemmParams &p) {
  WresLayout L;
  const uint32_t nkb = (uint32_t)p.kblocks_per_tap;
  L.a_bytes = (uint32_t)kBlockM * (uint32_t)p.kb_bytes;
  L.b_bytes = (uint32_t)p.block_n * (uint32_t)p.kb_bytes;
  L.stage_bytes = (uint32_t)p.wres_sub * nkb * L.a_bytes;
  L.b_off = L.stage_bytes * (uint32_t)p.num_stages;
  L.staging_off = L.b_off + ((nkb * L.b_bytes + 1023u) & ~1023u);
  L.scale_off = L.staging_off + kStagingBytes;
  L.shift_off = L.scale_off + (uint32_t)p.cout_pad * 4;
  L.bars_off = (L.shift_off + (uint32_t)p.cout_pad * 4 + 15u) & ~15u;
  L.total = L.bars_off + (2 * kMaxStages + 2 * 2 + 1) * 8 + 16;
  return L;
}
}  // namespace

size_t conv_gemm_wres_smem_bytes(const GemmParams &p) { return (size_t)wres_layout(p).total + 1024; }

template <bool kTf32>
__global__ void __launch_bounds__(kGemmBoundThreads, 1)
conv_gemm_wres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const WresLayout L = wres_layout(p);
  float *s_scale = reinterpret_cast<float *>(smem + L.scale_off);
  float *s_shift = reinterpret_cast<float *>(smem + L.shift_off);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  uint64_t *empty_bar = full_bar + kMaxStages;
  uint64_t *tfull_bar = empty_bar + kMaxStages;
  uint64_t *tempty_bar = tfull_bar + 2;
  uint64_t *bfull_bar = tempty_bar + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.wres_sub, nkb = p.kblocks_per_tap;
  const int num_super = (p.m_tiles + S - 1) / S;

  if (warp == kWarpProd && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], kEpiWarps);
    }
    ptx::mbar_init(bfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == kWarpMma) ptx::tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.cout_pad; i += kGemmThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  if (warp == kWarpProd && lane == 0) {   // the weights do not depend on the previous kernel: fetch them before the wait
    ptx::mbar_arrive_expect_tx(bfull_bar, (uint32_t)nkb * L.b_bytes);
    for (int kb = 0; kb < nkb; ++kb)
      ptx::tma_load_2d(smem + L.b_off + (size_t)kb * L.b_bytes, &tmB, bfull_bar, kb * p.kb_elems, 0);
  }
  pdl_wait();

  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
  if (warp == kWarpProd) {
    // ===================== TMA producer: S activation tiles per stage =====================
    const bool skip = (LWP_DBG(p.debug) & 1) != 0;
    int stage = 0;
    uint32_t phase = 0, dst = smem_base;
    for (int st = blockIdx.x; st < num_super; st += gridDim.x) {
      const int t0 = st * S;
      const int n_sub = p.m_tiles - t0 < S ? p.m_tiles - t0 : S;
      if (!ptx::mbar_wait_u32(empty0 + 8u * stage, phase ^ 1u)) break;
      if (ptx::elect_one()) {
        if (skip) {
          ptx::mbar_arrive_u32(full0 + 8u * stage);
        } else {
          ptx::mbar_arrive_expect_tx_u32(full0 + 8u * stage, (uint32_t)(n_sub * nkb) * L.a_bytes);
          uint32_t d = dst;
          for (int j = 0; j < n_sub; ++j)
            for (int kb = 0; kb < nkb; ++kb, d += L.a_bytes)
              ptx::tma_load_4d_u32(d, &tmA, full0 + 8u * stage, kb * p.kb_elems, (t0 + j) * kBlockM, 0, 0);
        }
      }
      dst += L.stage_bytes;
      if (++stage == p.num_stages) { stage = 0; phase ^= 1u; dst = smem_base; }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    const bool do_mma = (LWP_DBG(p.debug) & 4) == 0;
    const bool thin = p.kb_bytes != kKBlockBytes;
    const uint64_t desc_hi = thin ? ptx::umma_desc_k_sw64(0) : ptx::umma_desc_k_sw128(0);
    const uint32_t a16 = L.a_bytes >> 4, b16 = L.b_bytes >> 4, stage16 = L.stage_bytes >> 4;
    const uint32_t base16 = (smem_base & 0x3FFFFu) >> 4, bbase16 = ((smem_base + L.b_off) & 0x3FFFFu) >> 4;
    const uint32_t idesc = p.idesc;
    const int ksteps = p.kb_bytes / 32;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0, sa16 = base16;
    ptx::mbar_wait(bfull_bar, 0);
    for (int st = blockIdx.x; st < num_super; st += gridDim.x) {
      const int t0 = st * S;
      const int n_sub = p.m_tiles - t0 < S ? p.m_tiles - t0 : S;
      if (!ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u)) break;
      if (!ptx::mbar_wait_u32(full0 + 8u * stage, phase)) break;
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        if (do_mma) {
          uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
          uint32_t a = sa16;
          for (int j = 0; j < n_sub; ++j, d_tmem += (uint32_t)p.block_n) {
            uint32_t b = bbase16;
            for (int kb = 0; kb < nkb; ++kb, a += a16, b += b16) {
              const uint64_t da = desc_hi | (uint64_t)a, db = desc_hi | (uint64_t)b;
              for (int k = 0; k < ksteps; ++k)
                ptx::umma<kTf32>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            }
          }
        }
        ptx::umma_commit_u32(empty0 + 8u * stage);   // the stage's S tiles are free once these MMAs retire
        ptx::umma_commit(&tfull_bar[acc]);           // ... and the S accumulator blocks are complete
      }
      __syncwarp();
      sa16 += stage16;
      if (++stage == p.num_stages) { stage = 0; phase ^= 1u; sa16 = base16; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (8 warps, two per TMEM lane quarter) =====================
    const int q = warp & 3, part = warp >> 2;
    int cols = p.n_store < p.block_n ? p.n_store : p.block_n;
    const int chunks = cols * (kTf32 ? 4 : 2) / kKBlockBytes;   // 128-byte chunks of one tile's row
    uint8_t *my_staging = smem + L.staging_off + (size_t)warp * kStageOutBytes;
    int acc = 0, sbuf_idx = 0;
    uint32_t acc_phase = 0;
    for (int st = blockIdx.x; st < num_super; st += gridDim.x) {
      const int t0 = st * S;
      const int n_sub = p.m_tiles - t0 < S ? p.m_tiles - t0 : S;
      if (!ptx::mbar_wait(&tfull_bar[acc], acc_phase)) break;
      ptx::tc_fence_after();
      if (!(LWP_DBG(p.debug) & 8)) {
        for (int j = 0; j < n_sub; ++j) {
          // a tile with two or more chunks is shared by the quarter's two warps chunk by chunk; single-chunk tiles
          // (N = 64 bf16) alternate between them tile by tile
          int pt = part, np = 2;
          if (chunks < 2) {
            if ((j & 1) != part) continue;
            pt = 0; np = 1;
          }
          const int x0 = (t0 + j) * kBlockM + q * 32;
          const size_t pix = (size_t)x0 + lane;
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + j * p.block_n);
          staged_epilogue_tile<kTf32>(&tmC, my_staging, 1, sbuf_idx, t_row, 0, p.block_n, p.n_store, s_scale, s_shift, p.act,
                                      p.residual, p.res_ld, pix < (size_t)p.W, pix, lane, x0, 0, 0, pt, np, 0);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) ptx::bulk_wait<0>();   // all tensor stores of this warp have landed
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) ptx::tmem_dealloc(tmem_base, 512);
}

int conv_gemm_wres_init() {
  static DeviceOnce once;
  int slot;
  if (!once.pending(&slot)) return LWP_OK;
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_wres_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_wres_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  once.done[slot] = true;
  return LWP_OK;
}

int conv_gemm_wres_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                          const GemmParams &p, int grid, cudaStream_t st) {
  const size_t smem = conv_gemm_wres_smem_bytes(p);
  if (tf32)
    LWP_CUDA_CHECK(launch_pdl(conv_gemm_wres_kernel<true>, grid, kGemmThreads, smem, st, 1, tmA, tmB, tmC, p));
  else
    LWP_CUDA_CHECK(launch_pdl(conv_gemm_wres_kernel<false>, grid, kGemmThreads, smem, st, 1, tmA, tmB, tmC, p));
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

}  // namespace lwp
