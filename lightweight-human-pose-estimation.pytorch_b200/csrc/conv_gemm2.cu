// CTA-pair variant of the tcgen05 implicit-GEMM convolution (conv_gemm.cu): clusters of two CTAs on one TPC issue
// tcgen05.mma.cta_group::2 with M = 256 (128 pixel rows per CTA) and N = block_n.  Each CTA TMA-loads its own A tile
// and only HALF of the weight tile (block_n / 2 rows); the tensor cores of the pair read both halves, so the
// shared-memory operand traffic per MMA drops from 8 KB to 6 KB per 64 cycles (N = 128) and the L2 -> SM weight
// traffic halves.  Same layer semantics, tile geometry and epilogue as conv_gemm_kernel.
//
//   leader CTA (cluster rank 0): producer + MMA issuer; its full barrier counts the bytes of BOTH CTAs' loads
//   peer CTA (rank 1): producer only signals the leader's barrier (cp.async.bulk.tensor ... cta_group::2)
//   smem stages are released in both CTAs by tcgen05.commit ... multicast::cluster (mask 0b11)
//   accumulator stages: full -> multicast commit; empty -> epilogue warps of both CTAs arrive on the leader's barrier
#include "common.cuh"
#include "conv_gemm.cuh"
#include "gemm_epilogue.cuh"
#include "tcgen05.cuh"

namespace lwp {

struct Smem2Layout {
  uint32_t stage_bytes, b_bytes, staging_off, scale_off, shift_off, bars_off, total;
};

__host__ __device__ inline Smem2Layout smem2_layout(int block_n, int num_stages, int cout_pad) {
  Smem2Layout L;
  L.b_bytes = (uint32_t)(block_n / 2) * kKBlockBytes;
  L.stage_bytes = kATileBytes + L.b_bytes;
  L.staging_off = L.stage_bytes * (uint32_t)num_stages;
  L.scale_off = L.staging_off + kStagingBytes;
  L.shift_off = L.scale_off + (uint32_t)cout_pad * 4;
  L.bars_off = (L.shift_off + (uint32_t)cout_pad * 4 + 15u) & ~15u;
  L.total = L.bars_off + (2 * kMaxStages + 2 * kMaxAccStages) * 8 + 16;
  return L;
}

size_t conv_gemm2_smem_bytes(const GemmParams &p) {
  return (size_t)smem2_layout(p.block_n, p.num_stages, p.cout_pad).total + 1024;
}

struct Tile2 {
  int img, y0, x0, n0;
};
// pair tile pt = (m pair, n tile); this CTA takes M tile 2 * mpair + rank (may lie past the last tile: then img == NIMG,
// every load is zero-filled and every store clipped)
__device__ __forceinline__ Tile2 decode_tile2(const GemmParams &p, int pt, int rank) {
  Tile2 c;
  const int mpair = pt / p.n_tiles;
  c.n0 = (pt - mpair * p.n_tiles) * p.block_n;
  const int m_tile = 2 * mpair + rank;
  const int per_img = p.tiles_x * p.tiles_y;
  c.img = m_tile / per_img;
  const int rem = m_tile - c.img * per_img;
  const int ty = rem / p.tiles_x;
  c.y0 = ty * p.tile_h;
  c.x0 = (rem - ty * p.tiles_x) * p.tile_w;
  return c;
}

template <bool kTf32>
__global__ void __launch_bounds__(kGemmBoundThreads, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const Smem2Layout L = smem2_layout(p.block_n, p.num_stages, p.cout_pad);
  float *s_scale = reinterpret_cast<float *>(smem + L.scale_off);
  float *s_shift = reinterpret_cast<float *>(smem + L.shift_off);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  uint64_t *empty_bar = full_bar + kMaxStages;
  uint64_t *tfull_bar = empty_bar + kMaxStages;
  uint64_t *tempty_bar = tfull_bar + kMaxAccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + kMaxAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int num_pt = m_pairs * p.n_tiles;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int k_iters = p.taps * p.kblocks_per_tap;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);    // leader's copy is the live one: its producer's arrive + both CTAs' bytes
      ptx::mbar_init(&empty_bar[s], 1);   // multicast commit of the leader's MMA thread
    }
    for (int a = 0; a < p.acc_stages; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);                 // multicast commit
      ptx::mbar_init(&tempty_bar[a], 2 * kEpiWarps);    // leader's copy: epilogue warps of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair(tmem_slot, p.tmem_cols);
  for (int i = threadIdx.x; i < p.cout_pad; i += kGemmThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();   // both CTAs' barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  // pipeline roles run warp-uniform with an elected issuing lane (see conv_gemm.cu: the loop length of these
  // single-lane roles bounds the kernel)
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const int taps_y = p.taps == 1 ? 1 : 3;
    const int nb = rank * (p.block_n / 2);
    int stage = 0;
    uint32_t phase = 0, dst = smem_base;
    bool ok = true;
    for (int pt = cluster_id; pt < num_pt && ok; pt += num_clusters) {
      const Tile2 tc = decode_tile2(p, pt, rank);
      int kcoord = 0;
      for (int ty = 0; ty < taps_y && ok; ++ty) {
        const int cy = tc.y0 + (p.taps == 1 ? 0 : (ty - 1) * p.dil);
        for (int tx = 0; tx < taps_y && ok; ++tx) {
          const int cx = tc.x0 + (p.taps == 1 ? 0 : (tx - 1) * p.dil);
          for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
            if (!ptx::mbar_wait_u32(empty0 + 8u * stage, phase ^ 1u)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 21); break; }
            if (ptx::elect_one()) {
              if (leader) ptx::mbar_arrive_expect_tx_u32(full0 + 8u * stage, 2u * L.stage_bytes);
              ptx::tma_load_4d_pair_u32(dst, &tmA, full0 + 8u * stage, kb * p.kb_elems, cx, cy, tc.img);
              ptx::tma_load_2d_pair_u32(dst + kATileBytes, &tmB, full0 + 8u * stage, kcoord + kb * p.kb_elems, tc.n0 + nb);
            }
            dst += L.stage_bytes;
            if (++stage == p.num_stages) { stage = 0; phase ^= 1u; dst = smem_base; }
          }
          kcoord += p.cin;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint64_t desc_hi = ptx::umma_desc_k_sw128(0);
      const uint32_t a_off16 = kATileBytes >> 4, stage16 = L.stage_bytes >> 4, base16 = (smem_base & 0x3FFFFu) >> 4;
      const uint32_t idesc = p.idesc;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, sa16 = base16;
      bool ok = true;
      for (int pt = cluster_id; pt < num_pt && ok; pt += num_clusters) {
        if (!ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u)) { if (lane == 0) atomicExch(p.err_flag, 22); break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
        for (int it = 0; it < k_iters; ++it) {
          if (!ptx::mbar_wait_u32(full0 + 8u * stage, phase)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 23); break; }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t da = desc_hi | (uint64_t)sa16, db = desc_hi | (uint64_t)(sa16 + a_off16);
            ptx::umma_pair<kTf32>(d_tmem, da, db, idesc, (uint32_t)(it != 0));
            ptx::umma_pair<kTf32>(d_tmem, da + 2u, db + 2u, idesc, 1u);
            ptx::umma_pair<kTf32>(d_tmem, da + 4u, db + 4u, idesc, 1u);
            ptx::umma_pair<kTf32>(d_tmem, da + 6u, db + 6u, idesc, 1u);
            ptx::umma_commit_pair_u32(empty0 + 8u * stage);   // frees this stage in BOTH CTAs
          }
          sa16 += stage16;
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; sa16 = base16; }
        }
        if (!ok) break;
        if (ptx::elect_one()) ptx::umma_commit_pair(&tfull_bar[acc]);       // accumulators of both CTAs complete
        __syncwarp();
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue (8 warps per CTA, its own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = row / p.tile_w, tx = row - ty * p.tile_w;
    int acc = 0, sbuf_idx = 0;
    uint32_t acc_phase = 0;
    for (int pt = cluster_id; pt < num_pt; pt += num_clusters) {
      if (!ptx::mbar_wait(&tfull_bar[acc], acc_phase)) { atomicExch(p.err_flag, 24); break; }
      ptx::tc_fence_after();
      const Tile2 tc = decode_tile2(p, pt, rank);
      const int y = tc.y0 + ty, x = tc.x0 + tx;
      const bool valid = tc.img < p.NIMG && y < p.H && x < p.W;
      const size_t pix = ((size_t)tc.img * p.H + y) * (size_t)p.W + x;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
      staged_epilogue_tile<kTf32>(&tmC, smem + L.staging_off + (size_t)(warp - 2) * kStageOutBytes, 1, sbuf_idx, t_row,
                                  tc.n0, p.block_n, p.n_store, s_scale, s_shift, p.act, p.residual, p.res_ld, valid, pix,
                                  lane, tc.x0 + (q * 32) % p.tile_w, tc.y0 + (q * 32) / p.tile_w, tc.img,
                                  (warp - 2) >> 2, kEpiWarps / 4);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_leader(&tempty_bar[acc]);  // the leader's MMA thread waits for both CTAs
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();   // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, p.tmem_cols);
}

int conv_gemm2_init() {
  static DeviceOnce once;
  int slot;
  if (!once.pending(&slot)) return LWP_OK;
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  once.done[slot] = true;
  return LWP_OK;
}

int conv_gemm2_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                      const GemmParams &p, int grid, cudaStream_t st) {
  const size_t smem = conv_gemm2_smem_bytes(p);
  if (tf32) LWP_CUDA_CHECK(launch_pdl(conv_gemm2_kernel<true>, grid, kGemmThreads, smem, st, 2, tmA, tmB, tmC, p));
  else LWP_CUDA_CHECK(launch_pdl(conv_gemm2_kernel<false>, grid, kGemmThreads, smem, st, 2, tmA, tmB, tmC, p));
  return LWP_OK;
}

}  // namespace lwp
