// lwp_plan: a recorded list of layer launches (stem, depthwise, tcgen05 implicit-GEMM convs, layout
// hand-off) with pre-built TMA tensor maps for one (batch, H, W, dtype).  Replaces
// PoseEstimationWithMobileNet.forward (reference models/with_mobilenet.py:114-123); the Python mirror
// walks its own module tree (same state_dict as the reference) and records one op per layer.
#include "common.cuh"
#include "conv_direct.cuh"
#include "conv_gemm.cuh"
#include "dwpw_gemm.cuh"
#include "frontend_fused.cuh"
#include "sepconv_gemm.cuh"
#include "tcgen05.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace lwp {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Shared-memory budget of one persistent GEMM CTA.  The whole 227 KB by default; LWP_GEMM_SMEM (bytes) leaves room for
// a block of another stream's kernel (the post-processing of the previous batch) to co-reside on the SM.
static int gemm_smem_cap() {
  static int cap = -1;
  if (cap < 0) {
    cap = 232448;
    if (const char *e = getenv("LWP_GEMM_SMEM")) { int v = atoi(e); if (v >= 96 * 1024 && v <= 232448) cap = v; }
  }
  return cap;
}

enum OpKind { OP_STEM, OP_DW, OP_GEMM, OP_NCHW, OP_DWPW, OP_HEADS, OP_SEP, OP_FRONTEND };

struct Op {
  OpKind kind;
  // stem / dw / nchw
  const void *in = nullptr;
  void *out = nullptr;
  const float *w = nullptr, *scale = nullptr, *shift = nullptr;
  int n = 0, H = 0, W = 0, C = 0, stride = 1, dil = 1, act = 0;
  int ld = 0, c0 = 0, in_f32 = 0;
  bool stem_u8 = false;          // stem reads a uint8 HWC frame and normalises it on the fly
  double mean[3] = {0, 0, 0}, img_scale = 1.0;
  DwTileGeom dwg;
  bool dw_tma = false, dw_tma_out = false;   // (tmB = output map of the depthwise kernel)
  // gemm (tmA is also the input map of the TMA depthwise kernel)
  CUtensorMap tmA, tmB, tmC;
  GemmParams gp;
  bool two_cta = false;   // conv_gemm2_kernel: CTA pairs, tcgen05.mma.cta_group::2
  bool strips = false;    // conv3x3_pair_kernel: CTA pairs + column-strip reuse of the activations (3x3, Cout 128)
  DwpwParams fp;
  SepParams sp;
  int grid = 0;
  // fused heads (tmA = X, tmB = W1, tmC = W2)
  int hd_px = 0, hd_cin = 0, hd_cmid = 0, hd_out_ld = 0, hd_f32_ld = 0;
  const float *hd_scale1 = nullptr, *hd_shift1 = nullptr, *hd_scale2 = nullptr, *hd_shift2 = nullptr;
  float *hd_out_f32 = nullptr;
  void *hd_out = nullptr;
  FrontendArgs fe;   // fused front end
  CUtensorMap tmD;   // fused 3x3 + 1x1: the 1x1's weights
};

}  // namespace lwp

struct lwp_plan {
  int dtype;
  std::vector<lwp::Op> ops;
  int *err_flag = nullptr;
  bool stem_direct = false;
  std::vector<void *> owned;   // device blobs built while recording (re-laid-out constants), freed with the plan
};

using namespace lwp;

extern "C" int lwp_plan_create(int dtype, lwp_plan **out) {
  LWP_REQUIRE(out != nullptr, "lwp_plan_create: null out");
  LWP_REQUIRE(dtype == LWP_DTYPE_BF16 || dtype == LWP_DTYPE_TF32, "lwp_plan_create: bad dtype %d", dtype);
  int dev = 0;
  LWP_CUDA_CHECK(cudaGetDevice(&dev));
  int rc = lwp_check_device(dev);
  if (rc != LWP_OK) return rc;
  if (get_encode_fn() == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return LWP_ECUDA; }
  rc = conv_gemm_init();
  if (rc != LWP_OK) return rc;
  lwp_plan *p = new lwp_plan();
  p->dtype = dtype;
  p->stem_direct = getenv("LWP_STEM_DIRECT") != nullptr && atoi(getenv("LWP_STEM_DIRECT")) != 0;
  cudaError_t e = cudaMalloc(&p->err_flag, sizeof(int));
  if (e != cudaSuccess) { delete p; set_error("cudaMalloc err_flag: %s", cudaGetErrorString(e)); return LWP_ECUDA; }
  cudaMemset(p->err_flag, 0, sizeof(int));
  *out = p;
  return LWP_OK;
}

extern "C" void lwp_plan_destroy(lwp_plan *p) {
  if (p == nullptr) return;
  if (p->err_flag) cudaFree(p->err_flag);
  for (void *d : p->owned) cudaFree(d);
  delete p;
}

extern "C" int lwp_plan_num_ops(const lwp_plan *p) { return p ? (int)p->ops.size() : 0; }

extern "C" int lwp_plan_num_launches(const lwp_plan *p) { return p ? (int)p->ops.size() : 0; }  // one kernel per op

extern "C" int lwp_plan_error_flag(lwp_plan *p) {
  LWP_REQUIRE(p != nullptr, "lwp_plan_error_flag: null plan");
  int v = 0;
  LWP_CUDA_CHECK(cudaMemcpy(&v, p->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  return v;
}

extern "C" int lwp_plan_add_stem(lwp_plan *p, const float *w, const float *scale, const float *shift, void *out, int n,
                                 int H, int W) {
  LWP_REQUIRE(p && w && scale && shift && out, "lwp_plan_add_stem: null pointer");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "lwp_plan_add_stem: bad shape %dx%dx%d", n, H, W);
  Op op;
  op.kind = OP_STEM;
  op.w = w; op.scale = scale; op.shift = shift; op.out = out; op.n = n; op.H = H; op.W = W;
  p->ops.push_back(op);
  return LWP_OK;
}

extern "C" int lwp_plan_add_stem_u8(lwp_plan *p, const float *w, const float *scale, const float *shift, void *out,
                                    int n, int H, int W, const double *img_mean3, double img_scale) {
  LWP_REQUIRE(img_mean3 != nullptr, "lwp_plan_add_stem_u8: null mean");
  int rc = lwp_plan_add_stem(p, w, scale, shift, out, n, H, W);
  if (rc != LWP_OK) return rc;
  Op &op = p->ops.back();
  op.stem_u8 = true;
  op.mean[0] = img_mean3[0]; op.mean[1] = img_mean3[1]; op.mean[2] = img_mean3[2];
  op.img_scale = img_scale;
  return LWP_OK;
}

extern "C" int lwp_plan_add_frontend(lwp_plan *p, const float *stem_w, const float *stem_scale, const float *stem_shift,
                                     const float *dw1_w, const float *dw1_scale, const float *dw1_shift, const void *pw_w,
                                     const float *pw_scale, const float *pw_shift, const float *dw2_w,
                                     const float *dw2_scale, const float *dw2_shift, void *out, int n, int H, int W,
                                     int input_u8, const double *img_mean3, double img_scale) {
  LWP_REQUIRE(p && stem_w && stem_scale && stem_shift && dw1_w && dw1_scale && dw1_shift && pw_w && pw_scale && pw_shift &&
                  dw2_w && dw2_scale && dw2_shift && out, "lwp_plan_add_frontend: null pointer");
  LWP_REQUIRE(p->dtype == LWP_DTYPE_BF16, "lwp_plan_add_frontend: bf16 plans only");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, "lwp_plan_add_frontend: H and W must be multiples of 4");
  LWP_REQUIRE(((uintptr_t)pw_w % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dw1_w % 16) == 0 && ((uintptr_t)dw2_w % 16) == 0,
              "lwp_plan_add_frontend: unaligned pointer");
  LWP_REQUIRE(!input_u8 || img_mean3 != nullptr, "lwp_plan_add_frontend: null mean");
  Op op;
  op.kind = OP_FRONTEND;
  FrontendArgs &a = op.fe;
  a.x_is_u8 = input_u8 != 0;
  if (input_u8) { a.mean[0] = img_mean3[0]; a.mean[1] = img_mean3[1]; a.mean[2] = img_mean3[2]; a.img_scale = img_scale; }
  a.stem_w = stem_w; a.stem_scale = stem_scale; a.stem_shift = stem_shift;
  a.dw1_w = dw1_w; a.dw1_scale = dw1_scale; a.dw1_shift = dw1_shift;
  a.pw_w = pw_w; a.pw_scale = pw_scale; a.pw_shift = pw_shift;
  a.dw2_w = dw2_w; a.dw2_scale = dw2_scale; a.dw2_shift = dw2_shift;
  a.out = out; a.n = n; a.H = H; a.W = W;
  p->ops.push_back(op);
  return LWP_OK;
}

extern "C" int lwp_plan_add_depthwise(lwp_plan *p, const void *in, void *out, const float *w, const float *scale,
                                      const float *shift, int n, int H, int W, int C, int stride, int dilation,
                                      int act) {
  LWP_REQUIRE(p && in && out && w && scale && shift, "lwp_plan_add_depthwise: null pointer");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "lwp_plan_add_depthwise: bad shape");
  LWP_REQUIRE((stride == 1 && (dilation == 1 || dilation == 2)) || (stride == 2 && dilation == 1),
              "lwp_plan_add_depthwise: unsupported stride %d dilation %d", stride, dilation);
  Op op;
  op.kind = OP_DW;
  op.in = in; op.out = out; op.w = w; op.scale = scale; op.shift = shift;
  op.n = n; op.H = H; op.W = W; op.C = C; op.stride = stride; op.dil = dilation; op.act = act;
  const bool f32 = p->dtype == LWP_DTYPE_TF32;
  const int es = f32 ? 4 : 2;
  if (depthwise_tma_geometry(f32, n, H, W, C, stride, dilation, &op.dwg) == LWP_OK && ((uintptr_t)in % 16) == 0 &&
      depthwise_tma_init() == LWP_OK) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)C * es * W, (cuuint64_t)C * es * W * H};
    cuuint32_t box[4] = {(cuuint32_t)op.dwg.cb, (cuuint32_t)op.dwg.iw, (cuuint32_t)op.dwg.ih, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode_fn()(&op.tmA, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                                 const_cast<void *>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(depthwise) failed: %d", (int)r); return LWP_ECUDA; }
    op.dw_tma = true;
    // output tensor map: one TMA store per tile (LWP_DW_TMA_STORE=0: register stores)
    op.dw_tma_out = false;
    if (getenv("LWP_DW_TMA_STORE") == nullptr || atoi(getenv("LWP_DW_TMA_STORE")) != 0) {
      cuuint64_t odims[4] = {(cuuint64_t)C, (cuuint64_t)op.dwg.Wo, (cuuint64_t)op.dwg.Ho, (cuuint64_t)n};
      cuuint64_t ostr[3] = {(cuuint64_t)C * es, (cuuint64_t)C * es * op.dwg.Wo, (cuuint64_t)C * es * op.dwg.Wo * op.dwg.Ho};
      cuuint32_t obox[4] = {(cuuint32_t)op.dwg.cb, (cuuint32_t)op.dwg.tw, (cuuint32_t)op.dwg.th, 1};
      CUresult ro = get_encode_fn()(&op.tmB, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, odims,
                                    ostr, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      op.dw_tma_out = ro == CUDA_SUCCESS && ((uintptr_t)out % 16) == 0;
    }
  }
  p->ops.push_back(op);
  return LWP_OK;
}

extern "C" int lwp_plan_add_nhwc_to_nchw(lwp_plan *p, const void *in, int in_ld, int in_is_f32, int c0, int c,
                                         float *out, int n, int H, int W) {
  LWP_REQUIRE(p && in && out && n > 0 && H > 0 && W > 0 && c > 0 && c0 >= 0 && c0 + c <= in_ld,
              "lwp_plan_add_nhwc_to_nchw: bad arguments");
  Op op;
  op.kind = OP_NCHW;
  op.in = in; op.out = out; op.ld = in_ld; op.in_f32 = in_is_f32 || p->dtype == LWP_DTYPE_TF32; op.c0 = c0; op.C = c;
  op.n = n; op.H = H; op.W = W;
  p->ops.push_back(op);
  return LWP_OK;
}

extern "C" int lwp_plan_add_conv_gemm(lwp_plan *p, const void *in, int in_ld, const void *w, const float *scale,
                                      const float *shift, const void *residual, int res_ld, void *out, int out_ld,
                                      float *out_f32, int out_f32_ld, int n, int H, int W, int Cin, int Cout, int taps,
                                      int dilation, int act) {
  LWP_REQUIRE(p && in && w && scale && shift && (out || out_f32), "lwp_plan_add_conv_gemm: null pointer");
  LWP_REQUIRE(taps == 1 || taps == 9, "lwp_plan_add_conv_gemm: taps must be 1 or 9");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && dilation >= 1, "lwp_plan_add_conv_gemm: bad shape");
  const bool tf32 = p->dtype == LWP_DTYPE_TF32;
  const int es = tf32 ? 4 : 2;
  // thin 1x1 layer whose whole K is 64 bytes: 64-byte K block (SWIZZLE_64B) instead of a half-empty 128-byte one
  const bool thin64 = taps == 1 && Cin * es == 64 && getenv("LWP_NO_SW64") == nullptr;
  const int kb_bytes = thin64 ? 64 : kKBlockBytes;
  const int kb_elems = kb_bytes / es;
  LWP_REQUIRE(Cin % 8 == 0 && in_ld % 8 == 0 && in_ld >= Cin, "lwp_plan_add_conv_gemm: Cin/in_ld must be multiples of 8");
  LWP_REQUIRE(taps == 1 || Cin % kb_elems == 0, "lwp_plan_add_conv_gemm: 3x3 needs Cin %% %d == 0", kb_elems);
  const int cout_pad = (Cout + 63) / 64 * 64;
  LWP_REQUIRE(cout_pad <= kMaxCout, "lwp_plan_add_conv_gemm: Cout too large");
  LWP_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)w % 16) == 0, "lwp_plan_add_conv_gemm: unaligned pointer");
  LWP_REQUIRE(!out || (out_ld % 8 == 0 && (uintptr_t)out % 16 == 0), "lwp_plan_add_conv_gemm: bad out_ld/alignment");
  LWP_REQUIRE(!out_f32 || (out_f32_ld % 4 == 0 && (uintptr_t)out_f32 % 16 == 0), "lwp_plan_add_conv_gemm: bad out_f32_ld");
  LWP_REQUIRE(!residual || (res_ld % 8 == 0 && (uintptr_t)residual % 16 == 0), "lwp_plan_add_conv_gemm: bad res_ld");

  Op op;
  op.kind = OP_GEMM;
  GemmParams &g = op.gp;
  g.fuse_pw = 0; g.scale2 = nullptr; g.shift2 = nullptr; g.act2 = 0; g.wres_sub = 0;
  g.taps = taps; g.dil = dilation; g.cin = Cin; g.kb_elems = kb_elems; g.kb_bytes = kb_bytes;
  g.kblocks_per_tap = (Cin + kb_elems - 1) / kb_elems;
  g.cout_pad = cout_pad;
  g.block_n = cout_pad % 256 == 0 ? 256 : cout_pad % 128 == 0 ? 128 : 64;
  if (const char *bn = getenv("LWP_BLOCK_N")) {  // tuning knob: wider N tiles halve the A re-reads of wide layers
    int v = atoi(bn);
    if ((v == 64 || v == 128 || v == 256) && cout_pad % v == 0) g.block_n = v;
  }
  g.n_tiles = cout_pad / g.block_n;
  // columns actually written: whole 8-groups that fit the narrowest destination row
  int n_store = cout_pad;
  if (out && out_ld < n_store) n_store = out_ld / 8 * 8;
  if (out_f32 && out_f32_ld < n_store) n_store = out_f32_ld / 8 * 8;
  if (residual) {
    int lim = (Cout + 7) / 8 * 8;  // residual only covers real channels
    if (lim < n_store) n_store = lim;
  }
  g.n_store = n_store;
  g.idesc = make_umma_idesc(tf32, kBlockM, g.block_n);
  g.act = act;
  g.scale = scale; g.shift = shift;
  g.residual = residual; g.res_ld = res_ld;
  g.out = out; g.out_ld = out_ld; g.out_f32 = out_f32; g.out_f32_ld = out_f32_ld;
  g.err_flag = p->err_flag;
  g.debug = debug_env("LWP_DEBUG_GEMM");
  const int smem_budget = gemm_smem_cap() - 1024 - kStagingBytes - 2 * cout_pad * 4 - 512;
  const int stage_bytes = (kBlockM + g.block_n) * kb_bytes;
  int stages = smem_budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  // two K blocks per stage when that still leaves 3 stages: halves the barrier round trips per MMA (the pipeline
  // loops of the single-lane producer / issuer roles, not the tensor pipe, bound the N = 128 layers)
  g.kbps = 1;
  {
    const char *e = getenv("LWP_KBPS");
    const int want = e ? atoi(e) : 2;
    if (want == 2 && !thin64 && g.kblocks_per_tap % 2 == 0 && smem_budget / (2 * stage_bytes) >= 3) {
      g.kbps = 2;
      stages = smem_budget / (2 * stage_bytes);
      if (stages > kMaxStages) stages = kMaxStages;
    }
  }
  // opt-in (LWP_STAGING=2) second staging buffer per epilogue warp; measured: no gain on any layer of this network
  g.staging_bufs = 1;
  if (const char *e = getenv("LWP_STAGING")) {
    const int st2 = (smem_budget - kStagingBytes) / (stage_bytes * g.kbps);
    if (atoi(e) == 2 && st2 >= 2) { g.staging_bufs = 2; stages = st2 > kMaxStages ? kMaxStages : st2; }
  }
  if (const char *sv = getenv("LWP_GEMM_STAGES")) { int v = atoi(sv); if (v >= 2 && v < stages) stages = v; }
  g.num_stages = stages;
  g.acc_stages = 512 / g.block_n;  // the CTA owns the SM (smem > half), so it can take all 512 TMEM columns
  if (g.acc_stages > kMaxAccStages) g.acc_stages = kMaxAccStages;
  if (const char *as = getenv("LWP_ACC_STAGES")) {
    int v = atoi(as);
    if (v >= 2 && v <= g.acc_stages) g.acc_stages = v;
  }
  g.tmem_cols = 512;

  // geometry: 1x1 -> whole batch flattened to one row of pixels; 3x3 -> best rectangular tile
  if (taps == 1) {
    g.H = 1; g.NIMG = 1; g.W = n * H * W;
    g.tile_w = 128; g.tile_h = 1;
  } else {
    g.H = H; g.W = W; g.NIMG = n;
    long long best = -1;
    for (int th = 1; th <= 128; th <<= 1) {
      int tw = 128 / th;
      long long area = (long long)ceil_div(H, th) * th * (long long)ceil_div(W, tw) * tw;
      if (best < 0 || area < best) { best = area; g.tile_h = th; g.tile_w = tw; }
    }
  }
  g.tiles_x = ceil_div(g.W, g.tile_w);
  g.tiles_y = ceil_div(g.H, g.tile_h);
  g.m_tiles = g.NIMG * g.tiles_x * g.tiles_y;
  long long total_tiles = (long long)g.m_tiles * g.n_tiles;
  op.grid = (int)(total_tiles < net_sms() ? total_tiles : net_sms());
  // CTA pairs (cta_group::2, M = 256): layers whose epilogue is the plain TMA-store one and that have enough tiles
  {
    // default (measured, 64 x 368x656 bf16): pairs win on the 1x1 layers with 256-wide N tiles (512->512: 138 -> 122 us,
    // the weight tile is half of the smem traffic there) and lose on the 128-wide 3x3 layers (90 -> 98 us)
    const char *e2 = getenv("LWP_GEMM_2CTA");
    const int mode = e2 ? atoi(e2) : -1;   // -1 default policy, 0 off, 1: 3x3 layers, 2: every eligible layer, 3: also small layers
    const bool plain = out != nullptr && out_f32 == nullptr && g.n_store % (kKBlockBytes / es) == 0 &&
                       getenv("LWP_NO_TMA_STORE") == nullptr;   // the pair kernel only has the TMA-store epilogue
    const bool want = thin64 ? false : mode < 0 ? (taps == 1 && g.block_n == 256 && g.m_tiles >= 2 * net_sms())
                               : (mode > 0 && g.block_n >= (mode >= 4 ? 64 : 128) && (mode >= 3 || g.m_tiles >= 2 * net_sms()) && (mode >= 2 || taps == 9));
    if (want && plain && conv_gemm2_init() == LWP_OK) {
      op.two_cta = true;
      g.kbps = 1;   // conv_gemm2_kernel: one K block per stage
      g.idesc = make_umma_idesc(tf32, 2 * kBlockM, g.block_n);
      const int stage2 = kATileBytes + (g.block_n / 2) * kKBlockBytes;
      int st2 = ((gemm_smem_cap() < 200 * 1024 ? gemm_smem_cap() - 4096 : 200 * 1024) - kStagingBytes) / stage2;
      g.num_stages = st2 > kMaxStages ? kMaxStages : st2;
      if (const char *sv = getenv("LWP_GEMM2_STAGES")) { int v = atoi(sv); if (v >= 2 && v < g.num_stages) g.num_stages = v; }   // experiment
      const long long pairs = (long long)((g.m_tiles + 1) / 2) * g.n_tiles;
      long long gr = 2 * pairs;
      const int cap = net_sms() / 2 * 2;
      op.grid = (int)(gr < cap ? gr : cap);
    }
  }

  // thin 1x1 layers (whole weight matrix <= 32 KB, one N tile): weights resident in shared memory, S = 256 / N tiles
  // per pipeline stage (conv_gemm_wres.cu).  LWP_GEMM_WRES=0 never, 1 every eligible layer, unset: single-K-block layers
  // with at least four super-tiles per SM, so that the coarser work unit costs little tail (measured, 64 x 368x656 bf16:
  // 32->64 165 -> 129 us, 64->128 72 -> 65 us; two-K-block layers +-1 us: 128->128 @92x164 88 -> 90, Cpm trunk 32 -> 31)
  {
    const char *ew = getenv("LWP_GEMM_WRES");
    const int mode = ew ? atoi(ew) : -1;
    const bool plain = out != nullptr && out_f32 == nullptr && g.n_store % (kKBlockBytes / es) == 0 &&
                       getenv("LWP_NO_TMA_STORE") == nullptr;
    if (mode != 0 && taps == 1 && plain && !op.two_cta && g.n_tiles == 1 && g.block_n <= 128 && g.kblocks_per_tap <= 2 &&
        Cin % kb_elems == 0) {
      GemmParams w = g;
      w.wres_sub = 256 / g.block_n;
      w.kbps = 1;
      w.acc_stages = 2;
      const int wres_stage = w.wres_sub * w.kblocks_per_tap * kBlockM * kb_bytes;
      const int b_bytes = (w.kblocks_per_tap * g.block_n * kb_bytes + 1023) / 1024 * 1024;
      int st = (gemm_smem_cap() - 2048 - kStagingBytes - b_bytes - 2 * cout_pad * 4 - 512) / wres_stage;
      if (st > kMaxStages) st = kMaxStages;
      if (const char *sv = getenv("LWP_GEMM_STAGES")) { int v = atoi(sv); if (v >= 2 && v < st) st = v; }
      w.num_stages = st;
      const long long supers = ((long long)g.m_tiles + w.wres_sub - 1) / w.wres_sub;
      const int min_per_sm = mode > 1 ? mode : 4;
      if (st >= 2 && (mode == 1 || (w.kblocks_per_tap == 1 && supers >= (long long)min_per_sm * net_sms())) &&
          conv_gemm_wres_init() == LWP_OK) {
        g = w;
        op.grid = (int)(supers < net_sms() ? supers : net_sms());
      }
    }
  }

  // 3x3, Cout 128, plain epilogue: CTA pairs + column strips (conv_gemm3.cu); LWP_CONV3=0 keeps the tap-by-tap kernels
  {
    const char *e3 = getenv("LWP_CONV3");
    const bool plain = out != nullptr && out_f32 == nullptr && g.n_store % (kKBlockBytes / es) == 0 &&
                       getenv("LWP_NO_TMA_STORE") == nullptr;
    if ((e3 == nullptr || atoi(e3) != 0) && taps == 9 && cout_pad == 128 && plain && Cin % kb_elems == 0 && dilation <= 2 &&
        H * W >= 128 && conv_gemm3_init() == LWP_OK) {
      op.strips = true; op.two_cta = false;
      g.block_n = 128; g.n_tiles = 1; g.kbps = 1;
      g.tile_w = 8; g.tile_h = 16;
      g.tiles_x = ceil_div(g.W, g.tile_w); g.tiles_y = ceil_div(g.H, g.tile_h);
      g.m_tiles = g.NIMG * g.tiles_x * g.tiles_y;
      g.idesc = make_umma_idesc(tf32, 2 * kBlockM, 128);
      g.c3_a_stages = 4; g.c3_b_stages = 8;
      while (g.c3_b_stages > 3 && conv_gemm3_smem_bytes(g) > (size_t)gemm_smem_cap()) --g.c3_b_stages;
      while (g.c3_a_stages > 2 && conv_gemm3_smem_bytes(g) > (size_t)gemm_smem_cap()) --g.c3_a_stages;
      g.acc_stages = 4; g.tmem_cols = 512;
      const long long pairs = (g.m_tiles + 1) / 2;
      const int cap = net_sms() / 2 * 2;
      op.grid = (int)(2 * pairs < cap ? 2 * pairs : cap);
    }
  }

  EncodeTiledFn enc = get_encode_fn();
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.NIMG};
    cuuint64_t strides[3] = {(cuuint64_t)in_ld * es, (cuuint64_t)in_ld * es * g.W, (cuuint64_t)in_ld * es * g.W * g.H};
    cuuint32_t box[4] = {(cuuint32_t)kb_elems, (cuuint32_t)g.tile_w, (cuuint32_t)(op.strips ? g.tile_h + 2 * dilation : g.tile_h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&op.tmA, dt, 4, const_cast<void *>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     thin64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return LWP_ECUDA; }
  }
  {
    const cuuint64_t ktot = (cuuint64_t)taps * Cin;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {ktot * es};
    cuuint32_t box[2] = {(cuuint32_t)kb_elems, (cuuint32_t)(op.two_cta || op.strips ? g.block_n / 2 : g.block_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&op.tmB, dt, 2, const_cast<void *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     thin64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return LWP_ECUDA; }
  }
  // epilogue through smem + TMA tensor store for plain single-output layers whose stored width is whole 128-byte chunks
  memset(&op.tmC, 0, sizeof(op.tmC));
  g.tma_store = 0; g.store_bw = 32; g.store_bh = 1;
  const int chunk_cols = kKBlockBytes / es;
  if (out != nullptr && out_f32 == nullptr && g.n_store % chunk_cols == 0 && getenv("LWP_NO_TMA_STORE") == nullptr) {
    g.store_bw = g.tile_w < 32 ? g.tile_w : 32;
    g.store_bh = 32 / g.store_bw;
    cuuint64_t dims[4] = {(cuuint64_t)g.n_store, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.NIMG};
    cuuint64_t strides[3] = {(cuuint64_t)out_ld * es, (cuuint64_t)out_ld * es * g.W, (cuuint64_t)out_ld * es * g.W * g.H};
    cuuint32_t box[4] = {(cuuint32_t)chunk_cols, (cuuint32_t)g.store_bw, (cuuint32_t)g.store_bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&op.tmC, dt, 4, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(C) failed: %d", (int)r); return LWP_ECUDA; }
    g.tma_store = 1;
  }
  p->ops.push_back(op);
  return LWP_OK;
}

extern "C" int lwp_plan_add_conv3x3_pw(lwp_plan *p, const void *in, int in_ld, const void *w, const float *scale,
                                       const float *shift, const void *residual, int res_ld, int act, const void *w2,
                                       const float *scale2, const float *shift2, int act2, void *out, int out_ld, int n,
                                       int H, int W, int Cin, int dilation) {
  LWP_REQUIRE(p && w2 && scale2 && shift2 && out, "lwp_plan_add_conv3x3_pw: null pointer");
  LWP_REQUIRE(p->dtype == LWP_DTYPE_BF16, "lwp_plan_add_conv3x3_pw: bf16 plans only");
  LWP_REQUIRE(((uintptr_t)w2 % 16) == 0, "lwp_plan_add_conv3x3_pw: unaligned pointer");
  // record the 3x3 as usual (its output map then describes the 1x1's output), and turn the op into the fused form
  const size_t before = p->ops.size();
  int rc = lwp_plan_add_conv_gemm(p, in, in_ld, w, scale, shift, residual, res_ld, out, out_ld, nullptr, 0, n, H, W, Cin, 128, 9,
                                  dilation, act);
  if (rc != LWP_OK) return rc;
  Op &op = p->ops.back();
  if (!op.strips || op.gp.n_store != 128 || op.gp.tma_store == 0) {   // not the strip kernel: the caller records two ops
    p->ops.resize(before);
    set_error("lwp_plan_add_conv3x3_pw: the layer does not run on the strip kernel");
    return LWP_ECAP;
  }
  GemmParams &g = op.gp;
  g.fuse_pw = 1; g.scale2 = scale2; g.shift2 = shift2; g.act2 = act2;
  g.acc_stages = 2;   // TMEM: 2 x 128 (3x3 accumulators) + 64 (A2) + 128 (1x1 accumulator)
  g.c3_a_stages = 4; g.c3_b_stages = 8;
  while (g.c3_b_stages > 3 && conv_gemm3_smem_bytes(g) > (size_t)gemm_smem_cap()) --g.c3_b_stages;
  while (g.c3_a_stages > 2 && conv_gemm3_smem_bytes(g) > (size_t)gemm_smem_cap()) --g.c3_a_stages;
  if (conv_gemm3_smem_bytes(g) > (size_t)gemm_smem_cap()) {
    p->ops.resize(before);
    set_error("lwp_plan_add_conv3x3_pw: tiles do not fit in shared memory");
    return LWP_ECAP;
  }
  EncodeTiledFn enc = get_encode_fn();
  cuuint64_t dims[2] = {128, 128};
  cuuint64_t strides[1] = {128 * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&op.tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w2), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { p->ops.resize(before); set_error("cuTensorMapEncodeTiled(W2) failed: %d", (int)r); return LWP_ECUDA; }
  return LWP_OK;
}

extern "C" int lwp_plan_add_dwpw(lwp_plan *p, const void *in, const float *dw_w, const float *dw_scale,
                                 const float *dw_shift, int dw_act, int dilation, const void *w, const float *scale,
                                 const float *shift, int act, const void *residual, int res_ld, void *out, int out_ld,
                                 int n, int H, int W, int Cin, int Cout) {
  LWP_REQUIRE(p && in && dw_w && dw_scale && dw_shift && w && scale && shift && out, "lwp_plan_add_dwpw: null pointer");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "lwp_plan_add_dwpw: bad shape");
  LWP_REQUIRE(dilation == 1 || dilation == 2, "lwp_plan_add_dwpw: dilation must be 1 or 2");
  const bool tf32 = p->dtype == LWP_DTYPE_TF32;
  const int es = tf32 ? 4 : 2;
  const int kb_ch = kKBlockBytes / es;
  const int cout_pad = (Cout + 63) / 64 * 64;
  LWP_REQUIRE(Cin % 8 == 0 && (Cin % kb_ch == 0 || Cin < kb_ch), "lwp_plan_add_dwpw: Cin must be a multiple of %d (or a multiple of 8 below it)", kb_ch);
  LWP_REQUIRE(cout_pad <= 512 && (512 % cout_pad == 0), "lwp_plan_add_dwpw: Cout (padded %d) must divide 512", cout_pad);
  LWP_REQUIRE(out_ld % 8 == 0 && out_ld >= Cout && ((uintptr_t)out % 16) == 0 && ((uintptr_t)in % 16) == 0 &&
                  ((uintptr_t)w % 16) == 0,
              "lwp_plan_add_dwpw: bad out_ld / alignment");
  LWP_REQUIRE(!residual || (res_ld % 8 == 0 && (uintptr_t)residual % 16 == 0), "lwp_plan_add_dwpw: bad res_ld");
  int rc = dwpw_init();
  if (rc != LWP_OK) return rc;
  Op op;
  op.kind = OP_DWPW;
  DwpwParams &f = op.fp;
  f.H = H; f.W = W; f.NIMG = n; f.dil = dilation; f.cin = Cin; f.kb_ch = kb_ch; f.kblocks = (Cin + kb_ch - 1) / kb_ch;
  f.cout_pad = cout_pad;
  const int chunk_cols = kKBlockBytes / es;
  int n_store = cout_pad;
  if (out_ld < n_store) n_store = out_ld / 8 * 8;
  if (residual) { int lim = (Cout + 7) / 8 * 8; if (lim < n_store) n_store = lim; }
  LWP_REQUIRE(n_store % chunk_cols == 0, "lwp_plan_add_dwpw: stored width %d is not whole 128-byte chunks", n_store);
  f.n_store = n_store;
  f.n_mma = cout_pad > 256 ? 2 : 1;
  f.n_per_mma = cout_pad / f.n_mma;
  f.idesc = make_umma_idesc(tf32, kBlockM, f.n_per_mma);
  f.acc_stages = 512 / cout_pad;
  if (f.acc_stages > 4) f.acc_stages = 4;
  f.dw_act = dw_act; f.act = act;
  f.scale = scale; f.shift = shift;
  {
    // depthwise constants re-laid-out per K block: [kblocks][9 taps | scale | shift][kb_ch] fp32, zero beyond Cin, so
    // that one 1-D bulk copy per K block brings them into the input stage
    f.dw_const_bytes = (uint32_t)(11 * kb_ch * sizeof(float));
    const size_t blob = (size_t)f.kblocks * f.dw_const_bytes;
    float *d = nullptr;
    LWP_CUDA_CHECK(cudaMalloc(&d, blob));
    p->owned.push_back(d);
    LWP_CUDA_CHECK(cudaMemset(d, 0, blob));
    for (int kb = 0; kb < f.kblocks; ++kb) {
      const int c0 = kb * kb_ch, nc = Cin - c0 < kb_ch ? Cin - c0 : kb_ch;
      float *dst = d + (size_t)kb * 11 * kb_ch;
      LWP_CUDA_CHECK(cudaMemcpy2D(dst, kb_ch * sizeof(float), dw_w + c0, (size_t)Cin * sizeof(float), nc * sizeof(float), 9,
                                  cudaMemcpyDeviceToDevice));
      LWP_CUDA_CHECK(cudaMemcpy(dst + 9 * kb_ch, dw_scale + c0, nc * sizeof(float), cudaMemcpyDeviceToDevice));
      LWP_CUDA_CHECK(cudaMemcpy(dst + 10 * kb_ch, dw_shift + c0, nc * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    f.dw_consts = d;
  }
  f.residual = residual; f.res_ld = res_ld; f.err_flag = p->err_flag;
  f.debug = debug_env("LWP_DEBUG_DWPW");
  // tile: rectangle of 128 pixels, width a multiple of 4 (a depthwise thread owns 4 consecutive columns)
  long long best = -1;
  for (int th = 1; th <= 32; th <<= 1) {
    const int tw = 128 / th;
    if (tw < 4) continue;
    const int iw = tw + 2 * dilation, ih = th + 2 * dilation;
    if (iw > 256 || ih > 256 || (long long)iw * ih * kKBlockBytes > 40 * 1024) continue;
    long long cost = (long long)ceil_div(H, th) * ceil_div(W, tw) * ((long long)iw * ih + 128);
    if (best < 0 || cost < best) { best = cost; f.tile_h = th; f.tile_w = tw; f.iw = iw; f.ih = ih; }
  }
  LWP_REQUIRE(best >= 0, "lwp_plan_add_dwpw: no tile shape fits");
  f.tiles_x = ceil_div(W, f.tile_w); f.tiles_y = ceil_div(H, f.tile_h);
  f.m_tiles = n * f.tiles_x * f.tiles_y;
  f.in_tile_bytes = (uint32_t)(f.iw * f.ih * kKBlockBytes);
  f.in_stage_bytes = f.in_tile_bytes + f.dw_const_bytes;
  f.b_stage_bytes = (uint32_t)(f.n_per_mma * kKBlockBytes);  // the weights ring works in halves of N when N > 256
  // shared-memory budget: prefer deep B / A rings, then a second input stage, then double-buffered staging
  f.b_stages = f.n_mma + 1; f.a_stages = 2; f.in_stages = 1; f.staging_bufs = 1;
  const size_t limit = 232448;
  // upgrade order (LWP_DWPW_ORDER, default "iisab i b"): the input ring first -- its TMA latency is what the depthwise
  // warps wait for -- then double-buffered staging, a third A stage, deeper weight ring
  const char *order = getenv("LWP_DWPW_ORDER") ? getenv("LWP_DWPW_ORDER") : "iisabib";
  for (const char *c = order; *c; ++c) {
    DwpwParams t = f;
    if (*c == 'i') { if (t.in_stages >= 4) continue; t.in_stages++; }
    else if (*c == 's') { if (t.staging_bufs >= 2) continue; t.staging_bufs = 2; }
    else if (*c == 'a') { if (t.a_stages >= 4) continue; t.a_stages++; }
    else if (*c == 'b') { if (t.b_stages >= 4) continue; t.b_stages++; }
    else continue;
    if (dwpw_smem_bytes(t) <= limit) f = t;
  }
  if (dwpw_smem_bytes(f) > limit) { set_error("lwp_plan_add_dwpw: shared memory budget exceeded"); return LWP_ECAP; }
  op.grid = f.m_tiles < net_sms() ? f.m_tiles : net_sms();

  EncodeTiledFn enc = get_encode_fn();
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * es, (cuuint64_t)Cin * es * W, (cuuint64_t)Cin * es * W * H};
    cuuint32_t box[4] = {(cuuint32_t)kb_ch, (cuuint32_t)f.iw, (cuuint32_t)f.ih, 1};
    CUresult r = enc(&op.tmA, dt, 4, const_cast<void *>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(dwpw in) failed: %d", (int)r); return LWP_ECUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * es};
    cuuint32_t box[2] = {(cuuint32_t)kb_ch, (cuuint32_t)f.n_per_mma};
    CUresult r = enc(&op.tmB, dt, 2, const_cast<void *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(dwpw B) failed: %d", (int)r); return LWP_ECUDA; }
  }
  {
    const int bw = f.tile_w < 32 ? f.tile_w : 32, bh = 32 / bw;
    cuuint64_t dims[4] = {(cuuint64_t)n_store, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)out_ld * es, (cuuint64_t)out_ld * es * W, (cuuint64_t)out_ld * es * W * H};
    cuuint32_t box[4] = {(cuuint32_t)chunk_cols, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    CUresult r = enc(&op.tmC, dt, 4, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(dwpw C) failed: %d", (int)r); return LWP_ECUDA; }
  }
  p->ops.push_back(op);
  return LWP_OK;
}

// dw constants re-laid-out per K block: [kblocks][9 taps | scale | shift][kb_ch] fp32, zero beyond Cin
static int build_dw_consts(lwp_plan *p, const float *dw_w, const float *dw_scale, const float *dw_shift, int Cin, int kb_ch,
                           int kblocks, const float **out) {
  const size_t blob = (size_t)kblocks * 11 * kb_ch * sizeof(float);
  float *d = nullptr;
  LWP_CUDA_CHECK(cudaMalloc(&d, blob));
  p->owned.push_back(d);
  LWP_CUDA_CHECK(cudaMemset(d, 0, blob));
  for (int kb = 0; kb < kblocks; ++kb) {
    const int c0 = kb * kb_ch, nc = Cin - c0 < kb_ch ? Cin - c0 : kb_ch;
    float *dst = d + (size_t)kb * 11 * kb_ch;
    LWP_CUDA_CHECK(cudaMemcpy2D(dst, kb_ch * sizeof(float), dw_w + c0, (size_t)Cin * sizeof(float), nc * sizeof(float), 9,
                                cudaMemcpyDeviceToDevice));
    LWP_CUDA_CHECK(cudaMemcpy(dst + 9 * kb_ch, dw_scale + c0, nc * sizeof(float), cudaMemcpyDeviceToDevice));
    LWP_CUDA_CHECK(cudaMemcpy(dst + 10 * kb_ch, dw_shift + c0, nc * sizeof(float), cudaMemcpyDeviceToDevice));
  }
  *out = d;
  return LWP_OK;
}

extern "C" int lwp_plan_add_sepconv(lwp_plan *p, const void *in, const float *dw_w, const float *dw_scale,
                                    const float *dw_shift, int dw_act, int stride, const void *w, const float *scale,
                                    const float *shift, int act, const void *residual, int res_ld, void *out, int out_ld,
                                    int n, int H, int W, int Cin, int Cout) {
  LWP_REQUIRE(p && in && dw_w && dw_scale && dw_shift && w && scale && shift && out, "lwp_plan_add_sepconv: null pointer");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "lwp_plan_add_sepconv: bad shape");
  LWP_REQUIRE(stride == 1 || stride == 2, "lwp_plan_add_sepconv: stride must be 1 or 2");
  const bool tf32 = p->dtype == LWP_DTYPE_TF32;
  const int es = tf32 ? 4 : 2;
  const int kb_ch = kKBlockBytes / es;
  const int cout_pad = (Cout + 63) / 64 * 64;
  LWP_REQUIRE(Cin % kb_ch == 0, "lwp_plan_add_sepconv: Cin must be a multiple of %d", kb_ch);
  LWP_REQUIRE(out_ld % 8 == 0 && out_ld >= Cout && ((uintptr_t)out % 16) == 0 && ((uintptr_t)in % 16) == 0 &&
                  ((uintptr_t)w % 16) == 0,
              "lwp_plan_add_sepconv: bad out_ld / alignment");
  LWP_REQUIRE(!residual || (res_ld % 8 == 0 && (uintptr_t)residual % 16 == 0), "lwp_plan_add_sepconv: bad res_ld");
  if (cout_pad > 256) { set_error("lwp_plan_add_sepconv: Cout %d needs more than one 256-column MMA", Cout); return LWP_ECAP; }
  int rc = sepconv_init();
  if (rc != LWP_OK) return rc;
  Op op;
  op.kind = OP_SEP;
  SepParams &f = op.sp;
  f.H = H; f.W = W; f.NIMG = n; f.stride = stride;
  f.Ho = (H - 1) / stride + 1; f.Wo = (W - 1) / stride + 1;
  f.cin = Cin; f.kb_ch = kb_ch; f.kblocks = Cin / kb_ch;
  f.cout_pad = cout_pad;
  int n_store = cout_pad;
  if (out_ld < n_store) n_store = out_ld / 8 * 8;
  if (residual) { int lim = (Cout + 7) / 8 * 8; if (lim < n_store) n_store = lim; }
  LWP_REQUIRE(n_store % 64 == 0, "lwp_plan_add_sepconv: stored width %d is not a multiple of 64 columns", n_store);
  f.n_store = n_store;
  f.slice_bytes = n_store / 4 * es;   // one compute warp stages 32 rows x a quarter of the columns
  if (f.slice_bytes > 128) { set_error("lwp_plan_add_sepconv: %d-byte output slices (fp32, N = 256)", f.slice_bytes); return LWP_ECAP; }
  f.idesc = make_umma_idesc(tf32, kBlockM, cout_pad);
  f.acc_stages = 512 / cout_pad > 4 ? 4 : 512 / cout_pad;
  f.dw_act = dw_act; f.act = act; f.scale = scale; f.shift = shift;
  f.residual = residual; f.res_ld = res_ld; f.err_flag = p->err_flag;
  // output tile: 128 pixels, width a multiple of 4 and at most 32 (one epilogue warp = a bw x bh box of 32 rows), height
  // a multiple of 2; cheapest = fewest (halo + output) pixels over the map
  long long best = -1;
  for (int tw = 4; tw <= 32; tw <<= 1) {
    const int th = 128 / tw;
    const int iw = (tw - 1) * stride + 3, ih = (th - 1) * stride + 3;
    if (iw > 256 || ih > 256) continue;
    long long cost = (long long)ceil_div(f.Ho, th) * ceil_div(f.Wo, tw) * ((long long)iw * ih + 2 * 128);
    if (best < 0 || cost < best) { best = cost; f.tile_h = th; f.tile_w = tw; f.iw = iw; f.ih = ih; }
  }
  LWP_REQUIRE(best >= 0, "lwp_plan_add_sepconv: no tile shape fits");
  if (const char *e = getenv("LWP_SEP_TW")) { int tw = atoi(e); if (tw == 4 || tw == 8 || tw == 16 || tw == 32) { f.tile_w = tw; f.tile_h = 128 / tw; f.iw = (tw - 1) * stride + 3; f.ih = (f.tile_h - 1) * stride + 3; } }
  f.tiles_x = ceil_div(f.Wo, f.tile_w); f.tiles_y = ceil_div(f.Ho, f.tile_h);
  f.m_tiles = n * f.tiles_x * f.tiles_y;
  f.in_stage_bytes = (uint32_t)(f.iw * f.ih * kKBlockBytes);
  // rings: as deep as shared memory allows (the depthwise groups work on 2 / 4 items at once), halo boxes first
  // (in_stages must be a multiple of the number of depthwise groups: a group refills the stages it reads)
  // ring depths are multiples of the number of depthwise groups: a stage is then always used by the same group, which
  // therefore sees every phase of its barriers (a group skipping a phase of a shared stage would pass a parity wait early)
  const int groups = tf32 ? 2 : 1;
  f.in_stages = 2; f.a_stages = 2;
  if (sepconv_smem_bytes(f) > (size_t)gemm_smem_cap()) { set_error("lwp_plan_add_sepconv: weights + rings do not fit in shared memory"); return LWP_ECAP; }
  for (int round = 0; round < 8; ++round) {
    SepParams t = f;
    if (t.in_stages <= t.a_stages && t.in_stages + groups <= 8) t.in_stages += groups; else if (t.a_stages + groups <= 8) t.a_stages += groups; else break;
    if (sepconv_smem_bytes(t) <= (size_t)gemm_smem_cap()) f = t; else break;
  }
  rc = build_dw_consts(p, dw_w, dw_scale, dw_shift, Cin, kb_ch, f.kblocks, &f.dw_consts);
  if (rc != LWP_OK) return rc;
  op.grid = f.m_tiles < net_sms() ? f.m_tiles : net_sms();

  EncodeTiledFn enc = get_encode_fn();
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * es, (cuuint64_t)Cin * es * W, (cuuint64_t)Cin * es * W * H};
    cuuint32_t box[4] = {(cuuint32_t)kb_ch, (cuuint32_t)f.iw, (cuuint32_t)f.ih, 1};
    CUresult r = enc(&op.tmA, dt, 4, const_cast<void *>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(sepconv in) failed: %d", (int)r); return LWP_ECUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * es};
    cuuint32_t box[2] = {(cuuint32_t)kb_ch, (cuuint32_t)cout_pad};
    CUresult r = enc(&op.tmB, dt, 2, const_cast<void *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(sepconv B) failed: %d", (int)r); return LWP_ECUDA; }
  }
  {
    const int bw = f.tile_w < 32 ? f.tile_w : 32, bh = 32 / bw;
    cuuint64_t dims[4] = {(cuuint64_t)n_store, (cuuint64_t)f.Wo, (cuuint64_t)f.Ho, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)out_ld * es, (cuuint64_t)out_ld * es * f.Wo, (cuuint64_t)out_ld * es * f.Wo * f.Ho};
    cuuint32_t box[4] = {(cuuint32_t)(n_store / 4), (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const CUtensorMapSwizzle sw = f.slice_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : f.slice_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(&op.tmC, dt, 4, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(sepconv C) failed: %d", (int)r); return LWP_ECUDA; }
  }
  p->ops.push_back(op);
  return LWP_OK;
}

extern "C" int lwp_plan_add_heads_fused(lwp_plan *p, const void *in, int in_ld, const void *w1, const float *scale1,
                                        const float *shift1, int c_mid, const void *w2, const float *scale2,
                                        const float *shift2, void *out, int out_ld, float *out_f32, int out_f32_ld,
                                        int n_pixels, int c_in) {
  LWP_REQUIRE(p && in && w1 && scale1 && shift1 && w2 && scale2 && shift2 && out_f32, "lwp_plan_add_heads_fused: null pointer");
  LWP_REQUIRE(p->dtype == LWP_DTYPE_BF16, "lwp_plan_add_heads_fused: bf16 plans only");
  LWP_REQUIRE(n_pixels > 0 && c_in > 0 && c_in % 64 == 0 && c_mid > 0 && c_mid % 64 == 0 && in_ld >= c_in && in_ld % 8 == 0,
              "lwp_plan_add_heads_fused: Cin / Cmid must be multiples of 64");
  LWP_REQUIRE(out_f32_ld >= 64 && out_f32_ld % 4 == 0 && (uintptr_t)out_f32 % 16 == 0, "lwp_plan_add_heads_fused: bad out_f32");
  LWP_REQUIRE(!out || (out_ld >= 64 && out_ld % 8 == 0 && (uintptr_t)out % 16 == 0), "lwp_plan_add_heads_fused: bad out");
  LWP_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)w1 % 16) == 0 && ((uintptr_t)w2 % 16) == 0, "lwp_plan_add_heads_fused: unaligned pointer");
  if (heads_fused_smem_bytes(c_in, c_mid) > 232448) { set_error("lwp_plan_add_heads_fused: tiles do not fit in shared memory"); return LWP_ECAP; }
  Op op;
  op.kind = OP_HEADS;
  op.hd_px = n_pixels; op.hd_cin = c_in; op.hd_cmid = c_mid; op.hd_out_ld = out_ld; op.hd_f32_ld = out_f32_ld;
  op.hd_scale1 = scale1; op.hd_shift1 = shift1; op.hd_scale2 = scale2; op.hd_shift2 = shift2;
  op.hd_out_f32 = out_f32; op.hd_out = out;
  EncodeTiledFn enc = get_encode_fn();
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)n_pixels};
    cuuint64_t strides[1] = {(cuuint64_t)in_ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)kBlockM};
    CUresult r = enc(&op.tmA, dt, 2, const_cast<void *>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(heads X) failed: %d", (int)r); return LWP_ECUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)c_mid};
    cuuint64_t strides[1] = {(cuuint64_t)c_in * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)heads_fused_chunk_cols(c_mid)};
    CUresult r = enc(&op.tmB, dt, 2, const_cast<void *>(w1), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(heads W1) failed: %d", (int)r); return LWP_ECUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)c_mid, 64};
    cuuint64_t strides[1] = {(cuuint64_t)c_mid * 2};
    cuuint32_t box[2] = {64, 64};
    CUresult r = enc(&op.tmC, dt, 2, const_cast<void *>(w2), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(heads W2) failed: %d", (int)r); return LWP_ECUDA; }
  }
  p->ops.push_back(op);
  return LWP_OK;
}

extern "C" int lwp_plan_run_range(lwp_plan *p, const void *x, int first, int last, void *stream) {
  LWP_REQUIRE(p != nullptr, "lwp_plan_run: null plan");
  LWP_REQUIRE(first >= 0 && last <= (int)p->ops.size() && first <= last, "lwp_plan_run_range: bad range");
  cudaStream_t st = (cudaStream_t)stream;
  const bool f32 = p->dtype == LWP_DTYPE_TF32;
  for (int i = first; i < last; ++i) {
    const Op &op = p->ops[i];
    int rc = LWP_OK;
    switch (op.kind) {
      case OP_STEM:
        LWP_REQUIRE(x != nullptr, "lwp_plan_run: the plan has a stem but x is NULL");
        if (p->stem_direct)   // LWP_STEM_DIRECT=1: the CUDA-core FFMA2 kernel (fp32 inputs and weights, not rounded to the plan dtype)
          rc = stem_launch(f32, x, op.stem_u8, op.mean, op.img_scale, op.w, op.scale, op.shift, op.out, op.n, op.H, op.W, st);
        else
          rc = stem_gemm_launch(f32, x, op.stem_u8, op.mean, op.img_scale, op.w, op.scale, op.shift, op.out, op.n, op.H,
                                op.W, p->err_flag, st);
        break;
      case OP_DW:
        if (op.dw_tma)
          rc = depthwise_tma_launch(f32, op.tmA, op.dw_tma_out ? &op.tmB : nullptr, op.out, op.w, op.scale, op.shift, op.n, op.H, op.W, op.C, op.stride,
                                    op.dil, op.act, op.dwg, st);
        else
          rc = depthwise_launch(f32, op.in, op.out, op.w, op.scale, op.shift, op.n, op.H, op.W, op.C, op.stride,
                                op.dil, op.act, st);
        break;
      case OP_GEMM:
        rc = op.gp.fuse_pw ? conv_gemm3_pw_launch(op.tmA, op.tmB, op.tmC, op.tmD, op.gp, op.grid, st)
             : op.strips  ? conv_gemm3_launch(f32, op.tmA, op.tmB, op.tmC, op.gp, op.grid, st)
             : op.two_cta ? conv_gemm2_launch(f32, op.tmA, op.tmB, op.tmC, op.gp, op.grid, st)
             : op.gp.wres_sub ? conv_gemm_wres_launch(f32, op.tmA, op.tmB, op.tmC, op.gp, op.grid, st)
                          : conv_gemm_launch(f32, op.tmA, op.tmB, op.tmC, op.gp, op.grid, st);
        break;
      case OP_DWPW:
        rc = dwpw_launch(f32, op.tmA, op.tmB, op.tmC, op.fp, op.grid, st);
        break;
      case OP_SEP:
        rc = sepconv_launch(f32, op.tmA, op.tmB, op.tmC, op.sp, op.grid, st);
        break;
      case OP_HEADS:
        rc = heads_fused_launch(op.tmA, op.tmB, op.tmC, op.hd_px, op.hd_cin, op.hd_cmid, op.hd_scale1, op.hd_shift1,
                                op.hd_scale2, op.hd_shift2, op.hd_out_f32, op.hd_f32_ld, op.hd_out, op.hd_out_ld, p->err_flag,
                                st);
        break;
      case OP_FRONTEND: {
        LWP_REQUIRE(x != nullptr, "lwp_plan_run: the plan has a front end but x is NULL");
        FrontendArgs a = op.fe;
        a.x = x;
        rc = frontend_fused_launch(a, p->err_flag, st);
        break;
      }
      case OP_NCHW:
        rc = nhwc_to_nchw_launch(op.in_f32 != 0, op.in, op.ld, op.c0, op.C, (float *)op.out, op.n, op.H * op.W, st);
        break;
    }
    if (rc != LWP_OK) return rc;
  }
  return LWP_OK;
}

extern "C" int lwp_plan_run(lwp_plan *p, const void *x, void *stream) {
  LWP_REQUIRE(p != nullptr, "lwp_plan_run: null plan");
  return lwp_plan_run_range(p, x, 0, (int)p->ops.size(), stream);
}
