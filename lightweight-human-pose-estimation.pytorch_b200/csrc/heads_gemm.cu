// The two 1x1 layers of a stage's heads as ONE back-to-back GEMM kernel (bf16 plans):
//
//   reference: InitialStage.heatmaps / .pafs and RefinementStage.heatmaps / .pafs of models/with_mobilenet.py:33-38,
//   74-79 -- conv(128, 512 | 128, 1x1) + ReLU followed by conv(512 | 128, 19 | 38, 1x1, relu=False); both heads read the
//   same trunk features, so their first layers are stacked (N1 = 1024 | 256) and their second layers are one
//   block-diagonal matrix (N2 = 64: 19 heat-maps | 38 PAFs | 7 zero rows), as in the two-kernel form.
//
// The two-kernel form writes the 1024-channel intermediate to HBM and reads it back (2 x 494 MB per step for the
// initial stage, more than every other tensor of the network).  Here it never leaves the SM.  Per 128-pixel tile, for
// each 64-column chunk c of the intermediate:
//   GEMM1   acc1[c & 1] (64 TMEM columns) = X[128 x 128] * W1[c]^T                      8 MMAs (K = 128), N = 64
//   convert acc1 -> bias + ReLU -> bf16 -> shared memory, written directly as the K-major SWIZZLE_128B A operand of
//           the second GEMM (one 128-byte row per pixel: exactly one K block), by the eight epilogue warps
//   GEMM2   acc2 += A2[c & 1] * W2[:, chunk c]^T                                         4 MMAs (K = 64), N = 64
// and after the last chunk the epilogue warps write acc2 + bias as float32 heads (and a bf16 copy into the
// refinement stage's concat buffer).  GEMM2 of chunk c-1 is issued after GEMM1 of chunk c, so the tensor pipe works
// on the next chunk while the CUDA cores convert the previous one.
//
//   warp 0   TMA producer: the X tile (two K blocks, double-buffered across tiles)
//   warp 10  TMA producer: ring of [W1 chunk (64 x 128) | W2 chunk (64 x 64)] stages (re-streamed from L2 per tile)
//   warp 1   MMA issuer of GEMM1, warp 11 MMA issuer of GEMM2
//   warps 2-9 converter / epilogue: warp w handles TMEM lane quarter w % 4 and column half (w - 2) / 4
#include "common.cuh"
#include "conv_gemm.cuh"
#include "tcgen05.cuh"

namespace lwp {

constexpr int kHdChunk = 64;                                  // intermediate columns per chunk = one bf16 K block
constexpr int kHdWStages = 5;
constexpr int kHdW1Bytes = kHdChunk * kKBlockBytes;           // one K block of a W1 chunk: 64 rows x 128 B
constexpr int kHdW2Bytes = 64 * kKBlockBytes;                 // W2 chunk: 64 output rows x 64 K
constexpr int kHdBufs = 2;                                    // acc1 / A2 buffers in flight between the two GEMMs
constexpr int kHdAcc2Col = kHdBufs * 64;                      // TMEM: acc1 buffers, then two acc2 buffers
constexpr int kHdTmemCols = 512;
constexpr int kHdWarpMma2 = kBProducerWarp + 1;               // warp 11: issuer of the second GEMM
constexpr int kHdThreads = kGemmThreads + 32;

struct HeadsParams {
  int n_px, m_tiles;
  int k1_blocks;           // Cin / 64
  int chunks;              // Cmid / 64
  uint32_t idesc;          // M = 128, N = 64
  const float *scale1, *shift1, *scale2, *shift2;
  float *out_f32; int out_f32_ld;
  void *out_bf16; int out_ld;      // optional bf16 copy
  int *err_flag;
  int debug;               // timing experiments only (LWP_DEBUG_HEADS): bit 0 skip GEMM1 MMAs, 1 skip GEMM2 MMAs,
                           // 2 skip the conversion, 3 skip the weight loads
};

struct HeadsSmem {
  uint32_t x_off, w_off, a2_off, s1_off, s2_off, bars_off, total, x_bytes, w_stage_bytes;
};
__host__ __device__ inline HeadsSmem heads_smem(int k1_blocks, int chunks) {
  HeadsSmem L;
  L.x_bytes = (uint32_t)k1_blocks * kATileBytes;
  L.w_stage_bytes = (uint32_t)k1_blocks * kHdW1Bytes + kHdW2Bytes;
  L.x_off = 0;
  L.w_off = 2 * L.x_bytes;
  L.a2_off = L.w_off + kHdWStages * L.w_stage_bytes;
  L.s1_off = L.a2_off + kHdBufs * kATileBytes;
  L.s2_off = L.s1_off + (uint32_t)chunks * kHdChunk * 8;      // scale1 | shift1
  L.bars_off = L.s2_off + 64 * 8;                             // scale2 | shift2
  L.total = L.bars_off + 40 * 8 + 16;   // 40 barrier slots + the TMEM address
  return L;
}

__global__ void __launch_bounds__(kHdThreads, 1)
heads_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const HeadsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const HeadsSmem L = heads_smem(p.k1_blocks, p.chunks);
  float *s_scale1 = reinterpret_cast<float *>(smem + L.s1_off), *s_shift1 = s_scale1 + p.chunks * kHdChunk;
  float *s_scale2 = reinterpret_cast<float *>(smem + L.s2_off), *s_shift2 = s_scale2 + 64;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  uint64_t *x_full = bars, *x_empty = bars + 2, *w_full = bars + 4, *w_empty = bars + 12;   // up to 8 weight stages
  uint64_t *acc1_full = bars + 20, *acc1_empty = bars + 24, *a2_full = bars + 28, *a2_empty = bars + 32;
  uint64_t *acc2_full = bars + 36, *acc2_empty = bars + 38;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 40);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW1);
    ptx::prefetch_tmap(&tmW2);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 1);
      ptx::mbar_init(&acc2_full[s], 1); ptx::mbar_init(&acc2_empty[s], kEpiWarps);
    }
    for (int s = 0; s < kHdBufs; ++s) {
      ptx::mbar_init(&acc1_full[s], 1); ptx::mbar_init(&acc1_empty[s], kEpiWarps);
      ptx::mbar_init(&a2_full[s], kEpiWarps); ptx::mbar_init(&a2_empty[s], 1);
    }
    for (int s = 0; s < kHdWStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kHdTmemCols);
  for (int i = threadIdx.x; i < p.chunks * kHdChunk; i += kHdThreads) { s_scale1[i] = p.scale1[i]; s_shift1[i] = p.shift1[i]; }
  if (threadIdx.x < 64) { s_scale2[threadIdx.x] = p.scale2[threadIdx.x]; s_shift2[threadIdx.x] = p.shift2[threadIdx.x]; }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t wfull0 = ptx::smem_u32(w_full), wempty0 = ptx::smem_u32(w_empty);

  if (warp == 0) {
    // ===================== TMA producer: X tiles =====================
    int it = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      if (!ptx::mbar_wait(&x_empty[s], ph ^ 1u)) { if (lane == 0) atomicExch(p.err_flag, 41); break; }
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&x_full[s], L.x_bytes);
        for (int kb = 0; kb < p.k1_blocks; ++kb)
          ptx::tma_load_2d(smem + L.x_off + (size_t)s * L.x_bytes + (size_t)kb * kATileBytes, &tmX, &x_full[s], kb * 64, t * kBlockM);
      }
      __syncwarp();
    }
  } else if (warp == kBProducerWarp) {
    // ===================== TMA producer: [W1 chunk | W2 chunk] ring =====================
    int stage = 0;
    uint32_t phase = 0, dst = smem_base + L.w_off;
    bool ok = true;
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x) {
      for (int c = 0; c < p.chunks; ++c) {
        if (!ptx::mbar_wait_u32(wempty0 + 8u * stage, phase ^ 1u)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 42); break; }
        if (ptx::elect_one()) {
          if (LWP_DBG(p.debug) & 8) { ptx::mbar_arrive_u32(wfull0 + 8u * stage); } else {
          ptx::mbar_arrive_expect_tx_u32(wfull0 + 8u * stage, L.w_stage_bytes);
          for (int kb = 0; kb < p.k1_blocks; ++kb)
            ptx::tma_load_2d_u32(dst + (uint32_t)kb * kHdW1Bytes, &tmW1, wfull0 + 8u * stage, kb * 64, c * kHdChunk);
          ptx::tma_load_2d_u32(dst + (uint32_t)p.k1_blocks * kHdW1Bytes, &tmW2, wfull0 + 8u * stage, c * kHdChunk, 0);
          }
        }
        dst += L.w_stage_bytes;
        if (++stage == kHdWStages) { stage = 0; phase ^= 1u; dst = smem_base + L.w_off; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer of the first GEMM =====================
    // (the two GEMMs have their own issuing warps: one thread's instruction latency would otherwise serialise them)
    const uint64_t desc_hi = ptx::umma_desc_k_sw128(0);
    const uint32_t x16 = ((smem_base + L.x_off) & 0x3FFFFu) >> 4, w16 = ((smem_base + L.w_off) & 0x3FFFFu) >> 4;
    const uint32_t xbuf16 = L.x_bytes >> 4, wst16 = L.w_stage_bytes >> 4;
    const uint32_t idesc = p.idesc;
    const uint32_t a1f0 = ptx::smem_u32(acc1_full), a1e0 = ptx::smem_u32(acc1_empty);
    int it = 0, ws = 0;
    uint32_t wph = 0, b = 0, ph = 0;    // acc1 buffer of the next chunk and the parity of its current use
    bool ok = true;
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x, ++it) {
      const int xs = it & 1;
      if (!ptx::mbar_wait(&x_full[xs], (uint32_t)(it >> 1) & 1u)) { if (lane == 0) atomicExch(p.err_flag, 43); break; }
      const uint32_t xa16 = x16 + (uint32_t)xs * xbuf16;
      for (int c = 0; c < p.chunks; ++c) {
        if (!ptx::mbar_wait_u32(wfull0 + 8u * ws, wph) || !ptx::mbar_wait_u32(a1e0 + 8u * b, ph ^ 1u)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 44); break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d1 = tmem_base + b * 64u;
          const uint32_t wb16 = w16 + (uint32_t)ws * wst16;
          for (int kb = 0; kb < p.k1_blocks && !(LWP_DBG(p.debug) & 1); ++kb) {
            const uint64_t da = desc_hi | (uint64_t)(xa16 + (uint32_t)kb * (kATileBytes >> 4));
            const uint64_t db = desc_hi | (uint64_t)(wb16 + (uint32_t)kb * (kHdW1Bytes >> 4));
            ptx::umma<false>(d1, da, db, idesc, kb == 0 ? 0u : 1u);
            ptx::umma<false>(d1, da + 2u, db + 2u, idesc, 1u);
            ptx::umma<false>(d1, da + 4u, db + 4u, idesc, 1u);
            ptx::umma<false>(d1, da + 6u, db + 6u, idesc, 1u);
          }
          ptx::umma_commit_u32(a1f0 + 8u * b);
          if (c == p.chunks - 1) ptx::umma_commit(&x_empty[xs]);   // last GEMM1 of the tile: the X tile is free
        }
        __syncwarp();
        if (++ws == kHdWStages) { ws = 0; wph ^= 1u; }
        if (++b == kHdBufs) { b = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kHdWarpMma2) {
    // ===================== MMA issuer of the second GEMM =====================
    // chunk n: A2[n & 1] (its a2_full implies the chunk's weight stage has landed) x the W2 part of stage n % 4
    const uint64_t desc_hi = ptx::umma_desc_k_sw128(0);
    const uint32_t w216 = ((smem_base + L.w_off + (uint32_t)p.k1_blocks * kHdW1Bytes) & 0x3FFFFu) >> 4, wst16 = L.w_stage_bytes >> 4;
    const uint32_t a216 = ((smem_base + L.a2_off) & 0x3FFFFu) >> 4;
    const uint32_t idesc = p.idesc;
    const uint32_t a2f0 = ptx::smem_u32(a2_full), a2e0 = ptx::smem_u32(a2_empty);
    int it = 0, ws = 0;
    uint32_t b = 0, ph = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x, ++it) {
      const int xs = it & 1;
      if (!ptx::mbar_wait(&acc2_empty[xs], ((uint32_t)(it >> 1) & 1u) ^ 1u)) { if (lane == 0) atomicExch(p.err_flag, 45); break; }
      const uint32_t d2 = tmem_base + (uint32_t)kHdAcc2Col + (uint32_t)xs * 64u;
      for (int c = 0; c < p.chunks; ++c) {
        if (!ptx::mbar_wait_u32(a2f0 + 8u * b, ph)) { ok = false; if (lane == 0) atomicExch(p.err_flag, 46); break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t da = desc_hi | (uint64_t)(a216 + b * (kATileBytes >> 4));
          const uint64_t db = desc_hi | (uint64_t)(w216 + (uint32_t)ws * wst16);
          if (!(LWP_DBG(p.debug) & 2)) {
          ptx::umma<false>(d2, da, db, idesc, c == 0 ? 0u : 1u);
          ptx::umma<false>(d2, da + 2u, db + 2u, idesc, 1u);
          ptx::umma<false>(d2, da + 4u, db + 4u, idesc, 1u);
          ptx::umma<false>(d2, da + 6u, db + 6u, idesc, 1u);
          }
          ptx::umma_commit_u32(a2e0 + 8u * b);                 // the A2 buffer and ...
          ptx::umma_commit_u32(wempty0 + 8u * (uint32_t)ws);   // ... the weight stage are free
          if (c == p.chunks - 1) ptx::umma_commit(&acc2_full[xs]);
        }
        __syncwarp();
        if (++ws == kHdWStages) ws = 0;
        if (++b == kHdBufs) { b = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== converter / epilogue =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    uint8_t *a2_row = smem + L.a2_off + row * kKBlockBytes;
    uint32_t b = 0, ph = 0;
    int it = 0;
    bool ok = true;
    const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x, ++it) {
      for (int c = 0; c < p.chunks; ++c) {
        if (!ptx::mbar_wait(&acc1_full[b], ph) || !ptx::mbar_wait(&a2_empty[b], ph ^ 1u)) { atomicExch(p.err_flag, 47); ok = false; break; }
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + b * 64u + (uint32_t)half * 32u, r);
        ptx::tmem_ld_wait(r);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc1_empty[b]);   // the accumulator is in registers: GEMM1 may reuse it
        const int cg0 = c * kHdChunk + half * 32;
        uint8_t *arow = a2_row + b * kATileBytes;
        if (!(LWP_DBG(p.debug) & 4))
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int cg = cg0 + g8 * 8;
          const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale1 + cg), sc1 = *reinterpret_cast<const float4 *>(s_scale1 + cg + 4);
          const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift1 + cg), sh1 = *reinterpret_cast<const float4 *>(s_shift1 + cg + 4);
          float2 a2[4];
          a2[0] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 0]), __uint_as_float(r[g8 * 8 + 1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
          a2[1] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 2]), __uint_as_float(r[g8 * 8 + 3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
          a2[2] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 4]), __uint_as_float(r[g8 * 8 + 5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
          a2[3] = __ffma2_rn(make_float2(__uint_as_float(r[g8 * 8 + 6]), __uint_as_float(r[g8 * 8 + 7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
          uint4 pk;
          __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __hmax2(__float22bfloat162_rn(a2[j]), zero2);   // ReLU after rounding == rounding after ReLU
          *reinterpret_cast<uint4 *>(arow + (((half * 4 + g8) ^ (row & 7)) << 4)) = pk;
        }
        ptx::fence_proxy_async();   // A2 writes (generic proxy) -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a2_full[b]);
        if (++b == kHdBufs) { b = 0; ph ^= 1u; }
      }
      // ---- tile epilogue: acc2 + bias -> float32 heads (+ bf16 copy) ----
      if (!ok) break;
      const int xs = it & 1;
      if (!ptx::mbar_wait(&acc2_full[xs], (uint32_t)(it >> 1) & 1u)) { atomicExch(p.err_flag, 48); break; }
      ptx::tc_fence_after();
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kHdAcc2Col + (uint32_t)xs * 64u + (uint32_t)half * 32u, r);
      ptx::tmem_ld_wait(r);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc2_empty[xs]);
      const long long pix = (long long)t * kBlockM + row;
      if (pix < p.n_px) {
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = fmaf(__uint_as_float(r[j]), s_scale2[half * 32 + j], s_shift2[half * 32 + j]);
        float4 *of = reinterpret_cast<float4 *>(p.out_f32 + pix * p.out_f32_ld + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) of[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        if (p.out_bf16 != nullptr) {
          uint4 *ob = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out_bf16) + pix * p.out_ld + half * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(y[8 * j + 2 * k], y[8 * j + 2 * k + 1]);
            ob[j] = pk;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kHdTmemCols);
}

size_t heads_fused_smem_bytes(int k1_blocks, int chunks) { return (size_t)heads_smem(k1_blocks, chunks).total + 1024; }

int heads_fused_launch(const CUtensorMap &tmX, const CUtensorMap &tmW1, const CUtensorMap &tmW2, int n_px, int c_in,
                       int c_mid, const float *scale1, const float *shift1, const float *scale2, const float *shift2,
                       float *out_f32, int out_f32_ld, void *out_bf16, int out_ld, int *err_flag, cudaStream_t st) {
  static DeviceOnce attr;
  int attr_slot;
  if (attr.pending(&attr_slot)) {
    LWP_CUDA_CHECK(cudaFuncSetAttribute(heads_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr.done[attr_slot] = true;
  }
  HeadsParams p;
  p.n_px = n_px; p.m_tiles = ceil_div(n_px, kBlockM);
  p.k1_blocks = c_in / 64; p.chunks = c_mid / kHdChunk;
  p.idesc = make_umma_idesc(false, kBlockM, 64);
  p.scale1 = scale1; p.shift1 = shift1; p.scale2 = scale2; p.shift2 = shift2;
  p.out_f32 = out_f32; p.out_f32_ld = out_f32_ld; p.out_bf16 = out_bf16; p.out_ld = out_ld;
  p.err_flag = err_flag;
  p.debug = debug_env("LWP_DEBUG_HEADS");
  const size_t smem = heads_fused_smem_bytes(p.k1_blocks, p.chunks);
  if (smem > 232448) { set_error("heads_fused: %zu bytes of shared memory", smem); return LWP_ECAP; }
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  LWP_CUDA_CHECK(launch_pdl(heads_fused_kernel, grid, kHdThreads, smem, st, 1, tmX, tmW1, tmW2, p));
  return LWP_OK;
}

}  // namespace lwp
