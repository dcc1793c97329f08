// Frame preparation on the GPU: what infer_fast does to the camera frame before the
// network sees it -- cv2.resize(img, (0, 0), fx=scale, fy=scale, INTER_CUBIC) on the uint8 BGR frame (demo.py:59) and the
// centred pad to the network size (demo.py:61-62 -> val.pad_width, val.py:36-49).  The normalisation (demo.py:60) is
// already fused into the stem kernel (lwp_plan_add_stem_u8), so the pad value in the uint8 domain is the mean (128):
// (128 - 128) * (1 / 256) == 0, the reference's pad value in the normalised image.
//
// Bit-exact restatement of OpenCV's generic uint8 cubic path (imgproc/src/resize.cpp: fixed-point coefficients
// saturate_cast<short>(c * 2048), HResizeCubic<uchar, int, short> in int32, then VResizeCubicVec_32s8u -- float32
// S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) with b = beta * 2^-22, round-to-nearest-even, saturating packs -- for the body of
// each row and the scalar VResizeCubic + FixedPtCast<int, uchar, 22> ((v + 2^21) >> 22) for the last (W * 3) % 8
// elements).  OpenCV builds with IPP enabled route 3-channel uint8 cubic resizes through IPP instead, whose results
// differ from this path by +-1 in about 5 % of the pixels (measured with cv2 4.13); the golden fixtures are taken with
// cv2.ipp.setUseIPP(False).
#include "common.cuh"

namespace lwp {

__device__ __forceinline__ void pp_cubic_coeffs(float t, float c[4]) {   // interpolateCubic, A = -0.75, every op rounded
  const float A = -0.75f;
  const float t1 = __fadd_rn(t, 1.f);
  float v = __fsub_rn(__fmul_rn(A, t1), -3.75f);
  v = __fadd_rn(__fmul_rn(v, t1), -6.0f);
  c[0] = __fsub_rn(__fmul_rn(v, t1), -3.0f);
  v = __fsub_rn(__fmul_rn(1.25f, t), 2.25f);
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(v, t), t), 1.f);
  const float u = __fsub_rn(1.f, t);
  v = __fsub_rn(__fmul_rn(1.25f, u), 2.25f);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(v, u), u), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

// tap 1 source index and the four fixed-point weights of destination coordinate d
__device__ __forceinline__ int pp_axis(int d, double scale, int (&w)[4]) {
  const float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
  const float fl = floorf(f);
  float c[4];
  pp_cubic_coeffs(__fsub_rn(f, fl), c);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int v = __float2int_rn(__fmul_rn(c[k], 2048.f));   // saturate_cast<short>(cbuf[k] * INTER_RESIZE_COEF_SCALE)
    w[k] = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
  }
  return (int)fl;
}

__device__ __forceinline__ int pp_clamp(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256)
resize_pad_u8_kernel(const uint8_t *__restrict__ src, int h, int w, uint8_t *__restrict__ dst, int Hp, int Wp, int H, int W,
                     int top, int left, double scale_x, double scale_y, uchar3 pad, long long total) {
  const int body = W * 3 - (W * 3) % 8;   // elements of a resized row taken by the 8-lane SIMD body
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(idx % Wp);
    const long long t = idx / Wp;
    const int yp = (int)(t % Hp), img = (int)(t / Hp);
    uint8_t *o = dst + idx * 3;
    const int x = xp - left, y = yp - top;
    if (x < 0 || x >= W || y < 0 || y >= H) { o[0] = pad.x; o[1] = pad.y; o[2] = pad.z; continue; }
    int ax[4], ay[4];
    const int sx = pp_axis(x, scale_x, ax), sy = pp_axis(y, scale_y, ay);
    int ix[4], T[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ix[k] = pp_clamp(sx - 1 + k, w - 1) * 3;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint8_t *row = src + ((size_t)img * h + pp_clamp(sy - 1 + r, h - 1)) * (size_t)w * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        T[c][r] = (int)__ldg(row + ix[0] + c) * ax[0] + (int)__ldg(row + ix[1] + c) * ax[1] + (int)__ldg(row + ix[2] + c) * ax[2] +
                  (int)__ldg(row + ix[3] + c) * ax[3];
    }
    const float s22 = 1.f / (2048.f * 2048.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int v;
      if (x * 3 + c < body) {   // VResizeCubicVec_32s8u
        float acc = __fmul_rn((float)T[c][3], __fmul_rn((float)ay[3], s22));
        acc = __fadd_rn(__fmul_rn((float)T[c][2], __fmul_rn((float)ay[2], s22)), acc);
        acc = __fadd_rn(__fmul_rn((float)T[c][1], __fmul_rn((float)ay[1], s22)), acc);
        acc = __fadd_rn(__fmul_rn((float)T[c][0], __fmul_rn((float)ay[0], s22)), acc);
        v = __float2int_rn(acc);
      } else {                  // scalar tail: integer sum, FixedPtCast<int, uchar, 22>
        v = (T[c][0] * ay[0] + T[c][1] * ay[1] + T[c][2] * ay[2] + T[c][3] * ay[3] + (1 << 21)) >> 22;
      }
      o[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
}

}  // namespace lwp

extern "C" int lwp_resize_pad_u8(const uint8_t *src, int n, int h, int w, uint8_t *dst, int Hp, int Wp, int H, int W, int top,
                                 int left, double inv_scale_x, double inv_scale_y, int pad_b, int pad_g, int pad_r, void *stream) {
  LWP_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && H > 0 && W > 0, "lwp_resize_pad_u8: bad arguments");
  LWP_REQUIRE(top >= 0 && left >= 0 && top + H <= Hp && left + W <= Wp, "lwp_resize_pad_u8: the resized frame does not fit the padded one");
  LWP_REQUIRE(inv_scale_x > 0 && inv_scale_y > 0, "lwp_resize_pad_u8: bad scale");
  const long long total = (long long)n * Hp * Wp;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)lwp::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  uchar3 pad = make_uchar3((unsigned char)pad_b, (unsigned char)pad_g, (unsigned char)pad_r);
  lwp::resize_pad_u8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, h, w, dst, Hp, Wp, H, W, top, left, 1.0 / inv_scale_x,
                                                                          1.0 / inv_scale_y, pad, total);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}
