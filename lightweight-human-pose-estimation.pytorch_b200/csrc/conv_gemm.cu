// Dense convolution (1x1 and 3x3, stride 1, any dilation) as an implicit GEMM on tcgen05 tensor cores.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + offset(tap), cin] * Wt[cout, tap, cin]
//
// * A (activations, NHWC) is fetched by TMA through a 4-D tensor map (C, W, H, N): a tile is a
//   tile_h x tile_w rectangle of 128 pixels, a tap is the same rectangle shifted by (dy, dx); TMA
//   zero-fills whatever falls outside the image, which IS the convolution's zero padding.  1x1 convs
//   use the same path with the whole batch flattened to one row of pixels.
// * B (weights, [Cout][taps*Cin], K-major) is fetched by a 2-D tensor map.
// * Both land in shared memory in the canonical K-major SWIZZLE_128B layout; one elected thread issues
//   tcgen05.mma (M = 128, N = block_n, K = 32 bytes per instruction), accumulating fp32 in TMEM.
// * Persistent CTAs, warp-specialised: warps 0-7 epilogue, warp 8 / 9 TMA producers (activations / weights), warp 10
//   MMA issuer;
//   a ring of smem stages (full/empty mbarriers) and two TMEM accumulator stages (tmem_full/empty) so
//   the epilogue of tile i overlaps the main loop of tile i+1.
// * Epilogue: tcgen05.ld -> y = act(acc*scale[c] + shift[c]) (+ residual) -> plan dtype and/or fp32.
//
// Replaces nn.Conv2d (+ BatchNorm2d + ReLU/ELU + residual add) of modules/conv.py:4-10,19,30 and
// models/with_mobilenet.py:10,16,20,28-38,51-54,60,74-79 of the reference.
#include "common.cuh"
#include "conv_gemm.cuh"
#include "tcgen05.cuh"
#include "gemm_epilogue.cuh"

namespace lwp {

// Warp roles of conv_gemm_kernel.  The warp scheduler favours higher warp ids, and the three single-lane pipeline roles
// are latency-critical (their loop length bounds the short-K layers), so they sit ABOVE the eight epilogue warps, whose
// bulk conversion work would otherwise delay every barrier poll and MMA issue.  Epilogue warp w drains TMEM lane
// quarter w % 4.
constexpr int kWarpA = kEpiWarps, kWarpB = kEpiWarps + 1, kWarpMma = kEpiWarps + 2;

struct SmemLayout {
  uint32_t stage_bytes;
  uint32_t stages_off;   // 0 (after 1024-alignment)
  uint32_t staging_off;  // output staging of the TMA-store epilogue (1024-aligned: stages are multiples of 1024)
  uint32_t scale_off, shift_off, bars_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(int block_n, int num_stages, int cout_pad, int kb_bytes, int staging_bufs,
                                                  int kbps) {
  SmemLayout L;
  L.stage_bytes = (uint32_t)(kBlockM + block_n) * (uint32_t)kb_bytes * (uint32_t)kbps;
  L.stages_off = 0;
  L.staging_off = L.stage_bytes * (uint32_t)num_stages;
  L.scale_off = L.staging_off + (uint32_t)staging_bufs * kStagingBytes;
  L.shift_off = L.scale_off + (uint32_t)cout_pad * 4;
  L.bars_off = (L.shift_off + (uint32_t)cout_pad * 4 + 15u) & ~15u;
  L.total = L.bars_off + (2 * kMaxStages + 2 * kMaxAccStages) * 8 + 16;
  return L;
}

size_t conv_gemm_smem_bytes(const GemmParams &p) {
  return (size_t)smem_layout(p.block_n, p.num_stages, p.cout_pad, p.kb_bytes, p.staging_bufs, p.kbps).total + 1024;  // + alignment slack
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == LWP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == LWP_ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}

struct TileCoord {
  int img, y0, x0, n0;
};
// (n tile, tile column, tile row, image) of a tile index, advanced by the grid stride without divisions: on the thin
// layers a tile is a few hundred cycles of work and three integer divisions per role per tile were a third of it
struct TileCursor {
  int n, tx, ty, img;            // position
  int dn, dtx, dty, dimg;        // decomposition of the stride
  __device__ __forceinline__ void init(const GemmParams &p, int t, int stride) {
    decompose(p, t, n, tx, ty, img);
    decompose(p, stride, dn, dtx, dty, dimg);
  }
  static __device__ __forceinline__ void decompose(const GemmParams &p, int t, int &n_, int &tx_, int &ty_, int &img_) {
    const int m_tile = t / p.n_tiles;
    n_ = t - m_tile * p.n_tiles;
    const int per_img = p.tiles_x * p.tiles_y;
    img_ = m_tile / per_img;
    const int rem = m_tile - img_ * per_img;
    ty_ = rem / p.tiles_x;
    tx_ = rem - ty_ * p.tiles_x;
  }
  __device__ __forceinline__ void advance(const GemmParams &p) {   // mixed-radix add with carries
    n += dn;
    int c = 0;
    if (n >= p.n_tiles) { n -= p.n_tiles; c = 1; }
    tx += dtx + c; c = 0;
    if (tx >= p.tiles_x) { tx -= p.tiles_x; c = 1; }
    ty += dty + c; c = 0;
    if (ty >= p.tiles_y) { ty -= p.tiles_y; c = 1; }
    img += dimg + c;
  }
  __device__ __forceinline__ TileCoord coord(const GemmParams &p) const {
    TileCoord tc;
    tc.img = img; tc.y0 = ty * p.tile_h; tc.x0 = tx * p.tile_w; tc.n0 = n * p.block_n;
    return tc;
  }
};

template <bool kTf32>
__global__ void __launch_bounds__(kGemmBoundThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemLayout L = smem_layout(p.block_n, p.num_stages, p.cout_pad, p.kb_bytes, p.staging_bufs, p.kbps);
  const uint32_t a_bytes = (uint32_t)kBlockM * (uint32_t)p.kb_bytes;
  float *s_scale = reinterpret_cast<float *>(smem + L.scale_off);
  float *s_shift = reinterpret_cast<float *>(smem + L.shift_off);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  uint64_t *empty_bar = full_bar + kMaxStages;
  uint64_t *tfull_bar = empty_bar + kMaxStages;
  uint64_t *tempty_bar = tfull_bar + kMaxAccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + kMaxAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.kblocks_per_tap / p.kbps;
  const uint32_t mini_bytes = (uint32_t)(kBlockM + p.block_n) * (uint32_t)p.kb_bytes;   // one [A tile | B tile] pair

  if (warp == kWarpA && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (p.tma_store) ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);   // the activation producer and the weight producer each arrive once (+ their bytes)
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < p.acc_stages; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], kEpiWarps);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == kWarpMma) ptx::tmem_alloc(tmem_slot, p.tmem_cols);
  for (int i = threadIdx.x; i < p.cout_pad; i += kGemmThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();   // the next kernel of the stream may be scheduled as SMs free up ...
  pdl_wait();      // ... and this one touches activations only after the previous kernel has completed

  // The three pipeline roles below run WARP-UNIFORM (all 32 lanes execute the loops, one elected lane issues the
  // asynchronous instruction): every loop variable is then provably uniform, so the compiler keeps stage counters,
  // barrier / shared-memory addresses and TMA / UMMA operands in the uniform datapath instead of converting vector
  // registers for each issue.  A lone thread retires ~1 instruction per 6 cycles; at 256 tensor-pipe cycles per
  // K block (N = 128) the length of these loops IS the speed of the short-K layers.
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
  if (warp == kWarpA) {
    // ===================== TMA producer: activation tiles =====================
    const bool skip = (LWP_DBG(p.debug) & 1) != 0;
    const int taps_y = p.taps == 1 ? 1 : 3;
    int stage = 0;
    uint32_t phase = 0, dst = smem_base;
    bool ok = true;
    TileCursor cur;
    cur.init(p, blockIdx.x, gridDim.x);
    for (int t = blockIdx.x; t < num_tiles && ok; t += gridDim.x, cur.advance(p)) {
      const TileCoord tc = cur.coord(p);
      for (int ty = 0; ty < taps_y && ok; ++ty) {
        const int cy = tc.y0 + (p.taps == 1 ? 0 : (ty - 1) * p.dil);
        for (int tx = 0; tx < taps_y && ok; ++tx) {
          const int cx = tc.x0 + (p.taps == 1 ? 0 : (tx - 1) * p.dil);
          for (int kb = 0; kb < p.kblocks_per_tap; kb += p.kbps) {
            if (!ptx::mbar_wait_u32(empty0 + 8u * stage, phase ^ 1u)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 1); break; }
            if (ptx::elect_one()) {
              if (skip) {
                ptx::mbar_arrive_u32(full0 + 8u * stage);
              } else {
                ptx::mbar_arrive_expect_tx_u32(full0 + 8u * stage, a_bytes * (uint32_t)p.kbps);
                ptx::tma_load_4d_u32(dst, &tmA, full0 + 8u * stage, kb * p.kb_elems, cx, cy, tc.img);
                if (p.kbps == 2) ptx::tma_load_4d_u32(dst + mini_bytes, &tmA, full0 + 8u * stage, (kb + 1) * p.kb_elems, cx, cy, tc.img);
              }
            }
            dst += L.stage_bytes;
            if (++stage == p.num_stages) { stage = 0; phase ^= 1u; dst = smem_base; }
          }
        }
      }
    }
  } else if (warp == kWarpB) {
    // ===================== TMA producer: weight tiles =====================
    const bool skip = (LWP_DBG(p.debug) & 2) != 0;
    int stage = 0;
    uint32_t phase = 0, dst = smem_base + a_bytes;
    bool ok = true;
    const uint32_t b_bytes = mini_bytes - a_bytes;
    int nidx = blockIdx.x % p.n_tiles;
    const int dnidx = gridDim.x % p.n_tiles;
    for (int t = blockIdx.x; t < num_tiles && ok; t += gridDim.x) {
      const int n0 = nidx * p.block_n;
      nidx += dnidx;
      if (nidx >= p.n_tiles) nidx -= p.n_tiles;
      int kcoord = 0;
      for (int tap = 0; tap < p.taps && ok; ++tap) {
        for (int kb = 0; kb < p.kblocks_per_tap; kb += p.kbps) {
          if (!ptx::mbar_wait_u32(empty0 + 8u * stage, phase ^ 1u)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 5); break; }
          if (ptx::elect_one()) {
            if (skip) {
              ptx::mbar_arrive_u32(full0 + 8u * stage);
            } else {
              ptx::mbar_arrive_expect_tx_u32(full0 + 8u * stage, b_bytes * (uint32_t)p.kbps);
              ptx::tma_load_2d_u32(dst, &tmB, full0 + 8u * stage, kcoord + kb * p.kb_elems, n0);
              if (p.kbps == 2) ptx::tma_load_2d_u32(dst + mini_bytes, &tmB, full0 + 8u * stage, kcoord + (kb + 1) * p.kb_elems, n0);
            }
          }
          dst += L.stage_bytes;
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; dst = smem_base + a_bytes; }
        }
        kcoord += p.cin;
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    const bool do_mma = (LWP_DBG(p.debug) & 4) == 0;
    const bool thin = p.kb_bytes != kKBlockBytes;
    // descriptor without the start address; the address field is added per stage / per 32-byte K step (+2)
    const uint64_t desc_hi = thin ? ptx::umma_desc_k_sw64(0) : ptx::umma_desc_k_sw128(0);
    const uint32_t a_off16 = a_bytes >> 4, stage16 = L.stage_bytes >> 4, base16 = (smem_base & 0x3FFFFu) >> 4;
    const uint32_t idesc = p.idesc, mini16 = mini_bytes >> 4;
    const bool two = p.kbps == 2;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0, sa16 = base16;
    bool ok = true;
    for (int t = blockIdx.x; t < num_tiles && ok; t += gridDim.x) {
      if (!ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u)) { if (lane == 0) atomicExch(p.err_flag, 2); break; }
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
      for (int it = 0; it < k_iters; ++it) {
        if (!ptx::mbar_wait_u32(full0 + 8u * stage, phase)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 3); break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (do_mma) {
            const uint64_t da = desc_hi | (uint64_t)sa16, db = desc_hi | (uint64_t)(sa16 + a_off16);
            ptx::umma<kTf32>(d_tmem, da, db, idesc, (uint32_t)(it != 0));
            ptx::umma<kTf32>(d_tmem, da + 2u, db + 2u, idesc, 1u);   // 32 bytes of K per instruction: start address + 2
            if (!thin) {
              ptx::umma<kTf32>(d_tmem, da + 4u, db + 4u, idesc, 1u);
              ptx::umma<kTf32>(d_tmem, da + 6u, db + 6u, idesc, 1u);
            }
            if (two) {   // second K block of the stage
              const uint64_t da2 = da + mini16, db2 = db + mini16;
              ptx::umma<kTf32>(d_tmem, da2, db2, idesc, 1u);
              ptx::umma<kTf32>(d_tmem, da2 + 2u, db2 + 2u, idesc, 1u);
              ptx::umma<kTf32>(d_tmem, da2 + 4u, db2 + 4u, idesc, 1u);
              ptx::umma<kTf32>(d_tmem, da2 + 6u, db2 + 6u, idesc, 1u);
            }
          }
          ptx::umma_commit_u32(empty0 + 8u * stage);  // frees the smem stage once these MMAs retire
        }
        sa16 += stage16;
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; sa16 = base16; }
      }
      if (!ok) break;
      if (ptx::elect_one()) ptx::umma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (8 warps, two per TMEM lane quarter) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = row / p.tile_w, tx = row - ty * p.tile_w;
    int acc = 0, sbuf_idx = 0, tile_it = 0;
    uint32_t acc_phase = 0;
    TileCursor cur;
    cur.init(p, blockIdx.x, gridDim.x);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_it, cur.advance(p)) {
      if (!ptx::mbar_wait(&tfull_bar[acc], acc_phase)) { atomicExch(p.err_flag, 4); break; }
      ptx::tc_fence_after();
      const TileCoord tc = cur.coord(p);
      const int y = tc.y0 + ty, x = tc.x0 + tx;
      const bool valid = y < p.H && x < p.W;
      const size_t pix = ((size_t)tc.img * p.H + y) * (size_t)p.W + x;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
      if (p.tma_store) {
        // the two warps of a lane quarter alternate 128-byte chunks; a tile with a single chunk (N = 64 bf16) would leave
        // one of them idle, so there they alternate TILES instead (the store latency of one hides behind the other)
        int part = warp >> 2, nparts = kEpiWarps / 4;
        int cols = p.n_store - tc.n0;
        if (cols > p.block_n) cols = p.block_n;
        bool mine = true;
        if (cols * (kTf32 ? 4 : 2) < nparts * kKBlockBytes) {
          mine = (tile_it % nparts) == part;
          part = 0; nparts = 1;
        }
        if (mine && !(LWP_DBG(p.debug) & 8))
          staged_epilogue_tile<kTf32>(&tmC, smem + L.staging_off + (size_t)warp * p.staging_bufs * kStageOutBytes,
                                      p.staging_bufs, sbuf_idx, t_row,
                                      tc.n0, p.block_n, p.n_store, s_scale, s_shift, p.act, p.residual, p.res_ld, valid, pix,
                                      lane, tc.x0 + (q * 32) % p.tile_w, tc.y0 + (q * 32) / p.tile_w, tc.img, part, nparts,
                                      LWP_DBG(p.debug) >> 4);
      } else {
        for (int c = (warp >> 2) * 32; c < p.block_n; c += 32 * (kEpiWarps / 4)) {  // the quarter's two warps alternate
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_row + (uint32_t)c, r);
          ptx::tmem_ld_wait();
          if (valid) {
  #pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int cg = tc.n0 + c + g8 * 8;
              if (cg + 8 <= p.n_store) {
                float v[8];
  #pragma unroll
                for (int j = 0; j < 8; ++j)
                  v[j] = apply_act(fmaf(__uint_as_float(r[g8 * 8 + j]), s_scale[cg + j], s_shift[cg + j]), p.act);
                if (p.residual != nullptr) {
                  if constexpr (kTf32) {
                    const float4 *rp = reinterpret_cast<const float4 *>(
                        reinterpret_cast<const float *>(p.residual) + pix * p.res_ld + cg);
                    float4 a = __ldg(rp), b = __ldg(rp + 1);
                    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
                    v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
                  } else {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(
                        reinterpret_cast<const __nv_bfloat16 *>(p.residual) + pix * p.res_ld + cg));
                    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
  #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      float2 f = __bfloat1622float2(h[j]);
                      v[2 * j] += f.x; v[2 * j + 1] += f.y;
                    }
                  }
                }
                if (p.out != nullptr) {
                  if constexpr (kTf32) {
                    float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + pix * p.out_ld + cg);
                    op[0] = make_float4(v[0], v[1], v[2], v[3]);
                    op[1] = make_float4(v[4], v[5], v[6], v[7]);
                  } else {
                    uint4 pk;
                    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
  #pragma unroll
                    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                    *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + pix * p.out_ld + cg) = pk;
                  }
                }
                if (p.out_f32 != nullptr) {
                  float4 *op = reinterpret_cast<float4 *>(p.out_f32 + pix * p.out_f32_ld + cg);
                  op[0] = make_float4(v[0], v[1], v[2], v[3]);
                  op[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);  // accumulator stage drained
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
    if (p.tma_store && lane == 0) ptx::bulk_wait<0>();  // all tensor stores of this warp have landed
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

int conv_gemm_init() {
  static DeviceOnce once;
  int slot;
  if (!once.pending(&slot)) return LWP_OK;
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  LWP_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  once.done[slot] = true;
  return LWP_OK;
}

int conv_gemm_launch(bool tf32, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC,
                     const GemmParams &p, int grid, cudaStream_t st) {
  size_t smem = conv_gemm_smem_bytes(p);
  if (tf32)
    LWP_CUDA_CHECK(launch_pdl(conv_gemm_kernel<true>, grid, kGemmThreads, smem, st, 1, tmA, tmB, tmC, p));
  else
    LWP_CUDA_CHECK(launch_pdl(conv_gemm_kernel<false>, grid, kGemmThreads, smem, st, 1, tmA, tmB, tmC, p));
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

}  // namespace lwp
