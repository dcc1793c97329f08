// The two 1x1 layers of a stage's heads as ONE back-to-back GEMM kernel (bf16 plans):
//
//   reference: InitialStage.heatmaps / .pafs and RefinementStage.heatmaps / .pafs of models/with_mobilenet.py:33-38,
//   74-79 -- conv(128, 512 | 128, 1x1) + ReLU followed by conv(512 | 128, 19 | 38, 1x1, relu=False); both heads read the
//   same trunk features, so their first layers are stacked (N1 = 1024 | 256) and their second layers are one
//   block-diagonal matrix (N2 = 64: 19 heat-maps | 38 PAFs | 7 zero rows), as in the two-kernel form.
//
// The two-kernel form writes the 1024-channel intermediate to HBM and reads it back (2 x 494 MB per step for the
// initial stage, more than every other tensor of the network).  Here it never leaves the SM -- not even for shared
// memory: the intermediate goes TMEM -> registers -> TMEM and the second GEMM takes its A operand from tensor memory
// (tcgen05.mma with a TMEM A operand, as attention kernels do for P = softmax(S)).  Per 128-pixel tile, for each chunk n
// of CC = 128 intermediate columns (n runs over all chunks of all tiles of the CTA, buffer b = n % 2):
//   GEMM1   acc1[b] (128 fp32 TMEM columns) = X[128 x Cin] * W1[chunk]^T            Cin / 16 MMAs, N = 128
//   convert acc1[b] -> bias + ReLU -> bf16 pairs -> A2[b] (64 TMEM columns = 128 packed bf16 per pixel row: the K-major A
//           operand of the second GEMM), one 32 x 32 piece per converter warp (tcgen05.ld -> FFMA2 / F2FP / HMNMX2 -> tcgen05.st)
//   GEMM2   acc2 += A2[b] (TMEM) * W2[:, chunk]^T (smem)                             8 MMAs (K = 128), N = 64
// and after the last chunk of a tile acc2 + bias leaves as float32 heads (and a bf16 copy into the refinement stage's
// concat buffer), written by four warps of their own.
//
// History (64 frames @46x82, initial / refinement stage): 147 / 61 us for the first version (64-column chunks, the
// intermediate through a shared-memory A tile, converters also doing the tile epilogue) -> 100 / 40 us.  What the
// experiments of round 2 showed (LWP_TIMING_EXPERIMENTS build: per-role cycle accounting via lwp_debug_heads_prof, event
// traces, and scripts/microbench/*.cu):
//  * no role was waiting much: every role's loop iteration (two barrier waits, elect, MMAs or TMA issue, commit) costs
//    500-750 cycles whatever it moves, against 384 tensor-pipe cycles per 64-column chunk -> 128-column chunks;
//  * deeper buffering, fewer commits per chunk and more converter groups changed nothing (hand-offs are cheap: 130
//    cycles per mbarrier hop, 170 through tcgen05.commit; tcgen05.ld sustains ~480 B/clk/SM, tcgen05.st ~380);
//  * the generic->async proxy hand-off of a shared-memory A2 (fence.proxy.async per converter warp and chunk) and its
//    32 KB of shared-memory traffic per chunk went away with A2 in TMEM;
//  * the tile epilogue's global stores (48 KB per tile, all CTAs at about the same time) block the issuing warp for
//    ~4000 cycles when the memory system pushes back; done by the converters that stalled the whole pipeline once per
//    tile (30 % of the kernel) -> dedicated epilogue warps, acc2 double-buffered;
//  * what remains is the converter stage: its 16 warps run in lock-step through tcgen05.ld -> FFMA2 (1.7 warp
//    instructions/clk/SM) -> F2FP.BF16.PACK_AB (1.0/clk/SM: conversions are quarter rate) -> HMNMX2 -> tcgen05.st,
//    ~1100 cycles per chunk with nothing overlapping inside the stage; TMEM (512 columns = 2 x 128 acc1 + 2 x 64 A2 +
//    2 x 64 acc2) has no room for a third chunk in flight.
//
//   warp 0      TMA producer: the X tile (Cin / 64 K blocks, double-buffered across tiles)
//   warp 1      MMA issuer of GEMM1 (+ TMEM owner), warp 23 MMA issuer of GEMM2
//   warps 2-17  converters: TMEM lane quarter w % 4, 32-column piece (w - 2) / 4 of the chunk
//   warps 18-21 tile epilogue (TMEM lane quarter w % 4)
//   warp 22     TMA producer: W1-chunk ring (CC x Cin) and W2-chunk ring (64 x CC), 3 stages each, re-streamed from L2 per tile
#include "common.cuh"
#include "conv_gemm.cuh"
#include "tcgen05.cuh"

namespace lwp {

constexpr int kHdBufs = 2;                                    // chunks in flight between the GEMMs: acc1 / A2 TMEM buffers, indexed n % 2
constexpr int kHdWStages = 3;                                 // W1 and W2 rings, indexed n % 3
constexpr int kHdBufCols = 128;                               // TMEM columns of an acc1 buffer (fp32)
constexpr int kHdA2Col = kHdBufs * kHdBufCols;                // TMEM: acc1[2] (128 columns each), A2[2] (64: 128 packed bf16), acc2[2] (64)
constexpr int kHdAcc2Col = kHdA2Col + kHdBufs * 64;
constexpr int kHdTmemCols = 512;
constexpr int kHdConvWarps = 16;                              // TMEM lane quarter x 32-column piece of a 128-column chunk
constexpr int kHdEpiWarp0 = 2 + kHdConvWarps;                 // warps 18-21: tile epilogue (one per TMEM lane quarter)
constexpr int kHdWarpW = kHdEpiWarp0 + 4;                     // warp 22: TMA producer of both weight rings
constexpr int kHdWarpMma2 = kHdWarpW + 1;                     // warp 23: issuer of the second GEMM
constexpr int kHdThreads = 32 * (kHdWarpMma2 + 1);
constexpr int kHdW2Lag = 1;                                   // the producer loads W2 of chunk n after W1 of chunk n + 1

// Per-role cycle accounting (experiments build only): total cycles of the role's loop and the part of it spent in
// barrier waits, per CTA; read back with lwp_debug_heads_prof().
#ifdef LWP_TIMING_EXPERIMENTS
__device__ long long g_heads_prof[160 * 8 * 2];
#define HD_T0() long long hd_t0 = clock64(), hd_tw = 0
#define HD_WAIT(e) ([&]() -> bool { const long long a_ = clock64(); const bool r_ = (e); hd_tw += clock64() - a_; return r_; }())
#define HD_END(role) do { if (lane == 0 && blockIdx.x < 160) { g_heads_prof[(blockIdx.x * 8 + (role)) * 2] = clock64() - hd_t0; g_heads_prof[(blockIdx.x * 8 + (role)) * 2 + 1] = hd_tw; } } while (0)
#else
#define HD_T0() do {} while (0)
#define HD_WAIT(e) (e)
#define HD_END(role) do {} while (0)
#endif

struct HeadsParams {
  int n_px, m_tiles;
  int k1_blocks;           // Cin / 64
  int chunks;              // Cmid / chunk_cols
  int kbc;                 // 64-column K blocks per chunk: 2 (128-column chunks) or 1 (Cmid not a multiple of 128)
  uint32_t idesc1, idesc2; // M = 128, N = chunk_cols / M = 128, N = 64
  const float *scale1, *shift1, *scale2, *shift2;
  float *out_f32; int out_f32_ld;
  void *out_bf16; int out_ld;      // optional bf16 copy
  int *err_flag;
  int debug;               // timing experiments only (LWP_DEBUG_HEADS): bit 0 skip GEMM1 MMAs, 1 skip GEMM2 MMAs,
                           // 2 skip the conversion, 3 skip the weight loads
};

struct HeadsSmem {
  uint32_t x_off, w1_off, w2_off, s1_off, s2_off, bars_off, total, x_bytes, w1_kb_bytes, w1_stage_bytes, w2_stage_bytes;
};
__host__ __device__ inline HeadsSmem heads_smem(int k1_blocks, int chunks, int kbc) {
  HeadsSmem L;
  L.x_bytes = (uint32_t)k1_blocks * kATileBytes;
  L.w1_kb_bytes = (uint32_t)kbc * 64 * kKBlockBytes;          // one K block of a W1 chunk: chunk_cols rows x 128 B
  L.w1_stage_bytes = (uint32_t)k1_blocks * L.w1_kb_bytes;
  L.w2_stage_bytes = (uint32_t)kbc * 64 * kKBlockBytes;       // kbc K blocks of 64 output rows x 128 B
  L.x_off = 0;
  L.w1_off = 2 * L.x_bytes;
  L.w2_off = L.w1_off + kHdWStages * L.w1_stage_bytes;
  L.s1_off = L.w2_off + kHdWStages * L.w2_stage_bytes;
  L.s2_off = L.s1_off + (uint32_t)chunks * kbc * 64 * 8;      // scale1 | shift1
  L.bars_off = L.s2_off + 64 * 8;                             // scale2 | shift2
  L.total = L.bars_off + 40 * 8;   // 32 barrier slots + the TMEM address
  return L;
}

__global__ void __launch_bounds__(kHdThreads, 1)
heads_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ HeadsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const HeadsSmem L = heads_smem(p.k1_blocks, p.chunks, p.kbc);
  const int cmid = p.chunks * p.kbc * 64;
  float *s_scale1 = reinterpret_cast<float *>(smem + L.s1_off), *s_shift1 = s_scale1 + cmid;
  float *s_scale2 = reinterpret_cast<float *>(smem + L.s2_off), *s_shift2 = s_scale2 + 64;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars_off);
  // g1_done[b]: GEMM1 of the chunk in buffer b has completed (acc1[b] full); acc1_empty[b]: the converters have read it;
  // a2_full[b]: the converted chunk is in A2[b]; g2_done[b]: GEMM2 has read A2[b].  The weight rings have their own
  // full / empty pairs (three stages against two buffers).
  uint64_t *x_full = bars, *x_empty = bars + 2, *w1_full = bars + 4, *w1_empty = bars + 8, *w2_full = bars + 12, *w2_empty = bars + 16;
  uint64_t *g1_done = bars + 20, *acc1_empty = bars + 22, *a2_full = bars + 24, *g2_done = bars + 26;
  uint64_t *acc2_full = bars + 28, *acc2_empty = bars + 30;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW1);
    ptx::prefetch_tmap(&tmW2);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 1);
      ptx::mbar_init(&acc2_full[s], 1); ptx::mbar_init(&acc2_empty[s], 4);
      ptx::mbar_init(&g1_done[s], 1); ptx::mbar_init(&acc1_empty[s], 8 * p.kbc);   // one arrival per converting warp
      ptx::mbar_init(&a2_full[s], 8 * p.kbc); ptx::mbar_init(&g2_done[s], 1);
    }
    for (int s = 0; s < kHdWStages; ++s) {
      ptx::mbar_init(&w1_full[s], 1); ptx::mbar_init(&w1_empty[s], 1);
      ptx::mbar_init(&w2_full[s], 1); ptx::mbar_init(&w2_empty[s], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kHdTmemCols);
  for (int i = threadIdx.x; i < cmid; i += kHdThreads) { s_scale1[i] = p.scale1[i]; s_shift1[i] = p.shift1[i]; }
  if (threadIdx.x < 64) { s_scale2[threadIdx.x] = p.scale2[threadIdx.x]; s_shift2[threadIdx.x] = p.shift2[threadIdx.x]; }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  const uint32_t smem_base = ptx::smem_u32(smem);

  if (warp == 0) {
    // ===================== TMA producer: X tiles =====================
    HD_T0();
    int it = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      if (!HD_WAIT(ptx::mbar_wait(&x_empty[s], ph ^ 1u))) { if (lane == 0) atomicExch(p.err_flag, 41); break; }
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&x_full[s], L.x_bytes);
        for (int kb = 0; kb < p.k1_blocks; ++kb)
          ptx::tma_load_2d(smem + L.x_off + (size_t)s * L.x_bytes + (size_t)kb * kATileBytes, &tmX, &x_full[s], kb * 64, t * kBlockM);
      }
      __syncwarp();
    }
    HD_END(0);
  } else if (warp == kHdWarpW) {
    // ===================== TMA producer: W1 chunks and (kHdW2Lag chunks later) W2 chunks =====================
    const uint32_t w1f0 = ptx::smem_u32(w1_full), w2f0 = ptx::smem_u32(w2_full), g1d0 = ptx::smem_u32(w1_empty), g2d0 = ptx::smem_u32(w2_empty);
    const int my_tiles = blockIdx.x < p.m_tiles ? (p.m_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int total = my_tiles * p.chunks;
    const int cc = p.kbc * 64;
    int c1 = 0, c2 = 0;
    uint32_t b1 = 0, ph1 = 0, b2 = 0, ph2 = 0;
    HD_T0();
    for (int n = 0; n < total + kHdW2Lag; ++n) {
      if (n < total) {
        if (!HD_WAIT(ptx::mbar_wait_u32(g1d0 + 8u * b1, ph1 ^ 1u))) { if (lane == 0) atomicExch(p.err_flag, 42); break; }
        if (ptx::elect_one()) {
          if (LWP_DBG(p.debug) & 8) { ptx::mbar_arrive_u32(w1f0 + 8u * b1); } else {
          const uint32_t dst = smem_base + L.w1_off + b1 * L.w1_stage_bytes;
          ptx::mbar_arrive_expect_tx_u32(w1f0 + 8u * b1, L.w1_stage_bytes);
          for (int kb = 0; kb < p.k1_blocks; ++kb)
            ptx::tma_load_2d_u32(dst + (uint32_t)kb * L.w1_kb_bytes, &tmW1, w1f0 + 8u * b1, kb * 64, c1 * cc);
          }
        }
        __syncwarp();
        if (++c1 == p.chunks) c1 = 0;
        if (++b1 == kHdWStages) { b1 = 0; ph1 ^= 1u; }
      }
      if (n >= kHdW2Lag) {
        if (!HD_WAIT(ptx::mbar_wait_u32(g2d0 + 8u * b2, ph2 ^ 1u))) { if (lane == 0) atomicExch(p.err_flag, 49); break; }
        if (ptx::elect_one()) {
          if (LWP_DBG(p.debug) & 8) { ptx::mbar_arrive_u32(w2f0 + 8u * b2); } else {
          const uint32_t dst = smem_base + L.w2_off + b2 * L.w2_stage_bytes;
          ptx::mbar_arrive_expect_tx_u32(w2f0 + 8u * b2, L.w2_stage_bytes);
          for (int kb = 0; kb < p.kbc; ++kb)
            ptx::tma_load_2d_u32(dst + (uint32_t)kb * (64 * kKBlockBytes), &tmW2, w2f0 + 8u * b2, c2 * cc + kb * 64, 0);
          }
        }
        __syncwarp();
        if (++c2 == p.chunks) c2 = 0;
        if (++b2 == kHdWStages) { b2 = 0; ph2 ^= 1u; }
      }
    }
    HD_END(2);
  } else if (warp == 1) {
    // ===================== MMA issuer of the first GEMM =====================
    // (the two GEMMs have their own issuing warps: one thread's instruction latency would otherwise serialise them)
    const uint64_t desc_hi = ptx::umma_desc_k_sw128(0);
    const uint32_t x16 = ((smem_base + L.x_off) & 0x3FFFFu) >> 4, w16 = ((smem_base + L.w1_off) & 0x3FFFFu) >> 4;
    const uint32_t xbuf16 = L.x_bytes >> 4, wst16 = L.w1_stage_bytes >> 4, wkb16 = L.w1_kb_bytes >> 4;
    const uint32_t idesc = p.idesc1;
    const uint32_t g1d0 = ptx::smem_u32(g1_done), a1e0 = ptx::smem_u32(acc1_empty), w1f0 = ptx::smem_u32(w1_full), w1e0 = ptx::smem_u32(w1_empty);
    int it = 0;
    uint32_t b = 0, ph = 0;    // buffer of the next chunk and the parity of its current use
    uint32_t ws = 0, wph = 0;  // W1 ring stage and parity
    bool ok = true;
    HD_T0();
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x, ++it) {
      const int xs = it & 1;
      if (!HD_WAIT(ptx::mbar_wait(&x_full[xs], (uint32_t)(it >> 1) & 1u))) { if (lane == 0) atomicExch(p.err_flag, 43); break; }
      const uint32_t xa16 = x16 + (uint32_t)xs * xbuf16;
      for (int c = 0; c < p.chunks; ++c) {
        if (!HD_WAIT(ptx::mbar_wait_u32(w1f0 + 8u * ws, wph)) || !HD_WAIT(ptx::mbar_wait_u32(a1e0 + 8u * b, ph ^ 1u))) { ok = false; if (lane == 0) atomicExch(p.err_flag, 44); break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d1 = tmem_base + b * (uint32_t)kHdBufCols;
          const uint32_t wb16 = w16 + ws * wst16;
          for (int kb = 0; kb < p.k1_blocks && !(LWP_DBG(p.debug) & 1); ++kb) {
            const uint64_t da = desc_hi | (uint64_t)(xa16 + (uint32_t)kb * (kATileBytes >> 4));
            const uint64_t db = desc_hi | (uint64_t)(wb16 + (uint32_t)kb * wkb16);
            ptx::umma<false>(d1, da, db, idesc, kb == 0 ? 0u : 1u);
            ptx::umma<false>(d1, da + 2u, db + 2u, idesc, 1u);
            ptx::umma<false>(d1, da + 4u, db + 4u, idesc, 1u);
            ptx::umma<false>(d1, da + 6u, db + 6u, idesc, 1u);
          }
          ptx::umma_commit_u32(g1d0 + 8u * b);                      // acc1[b] full
          ptx::umma_commit_u32(w1e0 + 8u * ws);                     // the W1 stage is free
          if (c == p.chunks - 1) ptx::umma_commit(&x_empty[xs]);   // last GEMM1 of the tile: the X tile is free
        }
        __syncwarp();
        if (++ws == kHdWStages) { ws = 0; wph ^= 1u; }
        if (++b == kHdBufs) { b = 0; ph ^= 1u; }
      }
    }
    HD_END(1);
  } else if (warp == kHdWarpMma2) {
    // ===================== MMA issuer of the second GEMM: A from tensor memory, B (W2 chunk) from shared memory =====================
    const uint64_t desc_hi = ptx::umma_desc_k_sw128(0);
    const uint32_t w216 = ((smem_base + L.w2_off) & 0x3FFFFu) >> 4, wst16 = L.w2_stage_bytes >> 4;
    const uint32_t idesc = p.idesc2;
    const uint32_t a2f0 = ptx::smem_u32(a2_full), g2d0 = ptx::smem_u32(g2_done), w2f0 = ptx::smem_u32(w2_full), w2e0 = ptx::smem_u32(w2_empty);
    int it = 0;
    uint32_t b = 0, ph = 0, ws = 0, wph = 0;
    bool ok = true;
    HD_T0();
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x, ++it) {
      const int xs = it & 1;
      if (!HD_WAIT(ptx::mbar_wait(&acc2_empty[xs], ((uint32_t)(it >> 1) & 1u) ^ 1u))) { if (lane == 0) atomicExch(p.err_flag, 45); break; }
      const uint32_t d2 = tmem_base + (uint32_t)kHdAcc2Col + (uint32_t)xs * 64u;
      for (int c = 0; c < p.chunks; ++c) {
        if (!HD_WAIT(ptx::mbar_wait_u32(w2f0 + 8u * ws, wph)) || !HD_WAIT(ptx::mbar_wait_u32(a2f0 + 8u * b, ph))) { ok = false; if (lane == 0) atomicExch(p.err_flag, 46); break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t ta = tmem_base + (uint32_t)kHdA2Col + b * 64u;   // packed bf16: 8 TMEM columns per K = 16
          const uint64_t db0 = desc_hi | (uint64_t)(w216 + ws * wst16);
          if (!(LWP_DBG(p.debug) & 2))
          for (int kb = 0; kb < p.kbc; ++kb) {
            const uint64_t db = db0 + (uint64_t)((uint32_t)kb * ((64 * kKBlockBytes) >> 4));
            const uint32_t a = ta + (uint32_t)kb * 32u;
            ptx::umma_ts(d2, a, db, idesc, (c | kb) == 0 ? 0u : 1u);
            ptx::umma_ts(d2, a + 8u, db + 2u, idesc, 1u);
            ptx::umma_ts(d2, a + 16u, db + 4u, idesc, 1u);
            ptx::umma_ts(d2, a + 24u, db + 6u, idesc, 1u);
          }
          ptx::umma_commit_u32(g2d0 + 8u * b);                 // A2[b] and ...
          ptx::umma_commit_u32(w2e0 + 8u * ws);                // ... the W2 stage are free
          if (c == p.chunks - 1) ptx::umma_commit(&acc2_full[xs]);
        }
        __syncwarp();
        if (++ws == kHdWStages) { ws = 0; wph ^= 1u; }
        if (++b == kHdBufs) { b = 0; ph ^= 1u; }
      }
    }
    HD_END(3);
  } else if (warp < 2 + kHdConvWarps) {
    // ===================== converters =====================
    // All sixteen warps work on the SAME chunk: warp (q, s) converts fp32 columns [32 s, 32 s + 32) of TMEM lane quarter q
    // into the packed bf16 columns [16 s, 16 s + 16) of A2 (64-column chunks: only the warps with s < 2).  A buffer is held
    // for one piece per warp only, so GEMM1 refills one acc1 buffer in the shadow of the conversion of the other.
    const int q = warp & 3, s = (warp - 2) >> 2;
    const uint32_t tm_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t n = 0;
    bool ok = s < 2 * p.kbc;
    const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
    HD_T0();
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x) {
      for (int c = 0; c < p.chunks; ++c, ++n) {
        const uint32_t b = n & 1u, ph = (n >> 1) & 1u;
        if (!HD_WAIT(ptx::mbar_wait(&g1_done[b], ph)) || !HD_WAIT(ptx::mbar_wait(&g2_done[b], ph ^ 1u))) { atomicExch(p.err_flag, 47); ok = false; break; }
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_32x32(tm_lane + b * (uint32_t)kHdBufCols + (uint32_t)s * 32u, r);
        ptx::tmem_ld_wait(r);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc1_empty[b]);   // this warp's part of acc1[b] is in registers
        const int cg0 = c * p.kbc * 64 + s * 32;
        uint32_t o[16];
        if (!(LWP_DBG(p.debug) & 4)) {
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
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 hv = __hmax2(__float22bfloat162_rn(a2[k]), zero2);   // ReLU after rounding == rounding after ReLU
            o[g8 * 4 + k] = *reinterpret_cast<const uint32_t *>(&hv);
          }
        }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = r[k];
        }
        ptx::tmem_st_32x16(tm_lane + (uint32_t)kHdA2Col + b * 64u + (uint32_t)s * 16u, o);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a2_full[b]);
      }
    }
    if (q == 0 && s == 0) HD_END(4);
  } else if (warp < kHdEpiWarp0 + 4) {
    // ===================== tile epilogue: acc2 + bias -> float32 heads (+ bf16 copy) =====================
    // Its own warps: the stores of a tile (48 KB per CTA, every CTA at about the same time) block the issuing warp for
    // thousands of cycles when the memory system pushes back -- with the converters doing them, that stalled the
    // whole pipeline once per tile (30 % of the kernel).  acc2 is double-buffered, so these warps have a tile's time.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tm_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    HD_T0();
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
      const int xs = it & 1;
      if (!HD_WAIT(ptx::mbar_wait(&acc2_full[xs], (uint32_t)(it >> 1) & 1u))) { atomicExch(p.err_flag, 48); break; }
      ptx::tc_fence_after();
      const long long pix = (long long)t * kBlockM + row;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tm_lane + (uint32_t)kHdAcc2Col + (uint32_t)xs * 64u + (uint32_t)half * 32u, r);
        ptx::tmem_ld_wait(r);
        if (half == 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc2_empty[xs]);
        }
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
              __nv_bfloat162 *hh = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
              for (int k = 0; k < 4; ++k) hh[k] = __floats2bfloat162_rn(y[8 * j + 2 * k], y[8 * j + 2 * k + 1]);
              ob[j] = pk;
            }
          }
        }
      }
    }
    if (q == 0) HD_END(5);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kHdTmemCols);
}

#ifdef LWP_TIMING_EXPERIMENTS
extern "C" int lwp_debug_heads_prof(long long *out_host, int n) {   // experiments build only: [cta][role][total, waiting] cycles
  return cudaMemcpyFromSymbol(out_host, g_heads_prof, sizeof(long long) * (size_t)(n < 160 * 16 ? n : 160 * 16)) == cudaSuccess ? 0 : 1;
}
#endif

// c_mid: multiple of 64; chunks are 128 columns wide when c_mid is a multiple of 128
static inline int heads_kbc(int c_mid) { return c_mid % 128 == 0 ? 2 : 1; }
int heads_fused_chunk_cols(int c_mid) { return 64 * heads_kbc(c_mid); }
size_t heads_fused_smem_bytes(int c_in, int c_mid) {
  const int kbc = heads_kbc(c_mid);
  return (size_t)heads_smem(c_in / 64, c_mid / (64 * kbc), kbc).total + 1024;
}

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
  p.kbc = heads_kbc(c_mid);
  p.k1_blocks = c_in / 64; p.chunks = c_mid / (64 * p.kbc);
  p.idesc1 = make_umma_idesc(false, kBlockM, 64 * p.kbc);
  p.idesc2 = make_umma_idesc(false, kBlockM, 64);
  p.scale1 = scale1; p.shift1 = shift1; p.scale2 = scale2; p.shift2 = shift2;
  p.out_f32 = out_f32; p.out_f32_ld = out_f32_ld; p.out_bf16 = out_bf16; p.out_ld = out_ld;
  p.err_flag = err_flag;
  p.debug = debug_env("LWP_DEBUG_HEADS");
  const size_t smem = heads_fused_smem_bytes(c_in, c_mid);
  if (smem > 232448) { set_error("heads_fused: %zu bytes of shared memory", smem); return LWP_ECAP; }
  const int grid = p.m_tiles < net_sms() ? p.m_tiles : net_sms();
  LWP_CUDA_CHECK(launch_pdl(heads_fused_kernel, grid, kHdThreads, smem, st, 1, tmX, tmW1, tmW2, p));
  return LWP_OK;
}

}  // namespace lwp
