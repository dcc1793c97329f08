// CUDA-core kernels for the bandwidth-bound layers: stem 3x3/s2 (K = 27 is too thin for tensor cores),
// depthwise 3x3 (stride 1/2, dilation 1/2) with fused BN + ReLU / ELU, and the NHWC -> NCHW hand-off.
//
// Reference: modules/conv.py:4-10 (stem via conv(3, 32, stride=2, bias=False)), :13-32 (conv_dw, conv_dw_no_bn);
// models/with_mobilenet.py:93-105.
#include "common.cuh"
#include "conv_direct.cuh"

namespace lwp {

template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
    uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p));
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __bfloat1622float2(h[j]);
      v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8]) {
    uint4 pk;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4 *>(p) = pk;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float *p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float *p, const float (&v)[8]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == LWP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == LWP_ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}

// ------------------------------------------------------------------------------------------------
// stem: NCHW fp32 [n][3][H][W] -> NHWC [n][H/2][W/2][32]; 3x3, stride 2, pad 1; BN folded; ReLU.
// One thread = one output pixel x 16 output channels (two threads per pixel).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
stem_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ scale,
            const float *__restrict__ shift, T *__restrict__ out, int n, int H, int W, int Ho, int Wo) {
  __shared__ float s_w[27][32];  // [ci*9 + ky*3 + kx][co]
  __shared__ float s_scale[32], s_shift[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
    int co = i & 31, k = i >> 5;
    s_w[k][co] = w[co * 27 + k];
  }
  if (threadIdx.x < 32) { s_scale[threadIdx.x] = scale[threadIdx.x]; s_shift[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  const long long total = (long long)n * Ho * Wo * 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int half = (int)(idx & 1);
    long long pix = idx >> 1;
    const int xo = (int)(pix % Wo);
    long long t = pix / Wo;
    const int yo = (int)(t % Ho);
    const int img = (int)(t / Ho);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    const float *xb = x + (size_t)img * 3 * H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yi = 2 * yo - 1 + ky;
        if (yi < 0 || yi >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xi = 2 * xo - 1 + kx;
          if (xi < 0 || xi >= W) continue;
          const float v = __ldg(xb + ((size_t)ci * H + yi) * W + xi);
          const float *wr = &s_w[ci * 9 + ky * 3 + kx][half * 16];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
        }
      }
    }
    T *op = out + (size_t)pix * 32 + half * 16;
    float o[8];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = half * 16 + g * 8 + j;
        o[j] = fmaxf(fmaf(acc[g * 8 + j], s_scale[c], s_shift[c]), 0.f);
      }
      Vec8<T>::store(op + g * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, NHWC.  One thread = 8 channels x TW consecutive output columns; the input columns
// of a row are streamed through registers once and scattered to the outputs they touch, so every
// input vector is loaded once per thread (3 rows x ((TW-1)*S + 2*D + 1) loads for TW outputs).
// Weights arrive tap-major [9][C] so the 8 channels of a thread are one 32-byte read.
// ------------------------------------------------------------------------------------------------
template <typename T, int S, int D, int TW>
__global__ void __launch_bounds__(256)
depthwise3x3_kernel(const T *__restrict__ in, T *__restrict__ out, const float *__restrict__ w9c,
                    const float *__restrict__ scale, const float *__restrict__ shift, int n, int H, int W, int C,
                    int Ho, int Wo, int act) {
  constexpr int NCOL = (TW - 1) * S + 2 * D + 1;
  const int cvecs = C >> 3;
  const int xtiles = (Wo + TW - 1) / TW;
  const long long total = (long long)n * Ho * xtiles * cvecs;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvecs);
    long long t = idx / cvecs;
    const int xt = (int)(t % xtiles);
    t /= xtiles;
    const int yo = (int)(t % Ho);
    const int img = (int)(t / Ho);
    const int c0 = cv * 8;
    const int xo0 = xt * TW;
    float acc[TW][8];
#pragma unroll
    for (int a = 0; a < TW; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    const int xi0 = xo0 * S - D;  // leftmost input column of the window
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yi = yo * S - D + ky * D;
      if (yi < 0 || yi >= H) continue;
      float wk[3][8];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) Vec8<float>::load(w9c + (size_t)(ky * 3 + kx) * C + c0, wk[kx]);
      const T *rowp = in + (((size_t)img * H + yi) * W) * C + c0;
#pragma unroll
      for (int ci = 0; ci < NCOL; ++ci) {
        const int xi = xi0 + ci;
        float v[8];
        if (xi >= 0 && xi < W) Vec8<T>::load(rowp + (size_t)xi * C, v);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int a = 0; a < TW; ++a) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (a * S + kx * D == ci) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[a][j] = fmaf(v[j], wk[kx][j], acc[a][j]);
            }
          }
        }
      }
    }
    float sc[8], sh[8];
    Vec8<float>::load(scale + c0, sc);
    Vec8<float>::load(shift + c0, sh);
#pragma unroll
    for (int a = 0; a < TW; ++a) {
      const int xo = xo0 + a;
      if (xo >= Wo) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_apply(fmaf(acc[a][j], sc[j], sh[j]), act);
      Vec8<T>::store(out + (((size_t)img * Ho + yo) * Wo + xo) * C + c0, o);
    }
  }
}

// NHWC (T or float, pixel stride ld, channels [c0, c0 + c)) -> NCHW float32 [n][c][H][W]
template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T *__restrict__ in, int ld, int c0, int c, float *__restrict__ out, int n, int HW) {
  const long long total = (long long)n * c * HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % HW);
    long long t = idx / HW;
    const int ch = (int)(t % c);
    const int img = (int)(t / c);
    out[idx] = (float)in[((size_t)img * HW + p) * ld + c0 + ch];
  }
}

static int grid_for(long long total, int block, int per_sm) {
  long long b = (total + block - 1) / block;
  long long cap = (long long)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int stem_launch(bool f32, const float *x, const float *w, const float *scale, const float *shift, void *out, int n,
                int H, int W, cudaStream_t st) {
  const int Ho = H / 2, Wo = W / 2;
  long long total = (long long)n * Ho * Wo * 2;
  int grid = grid_for(total, 256, 8);
  if (f32) stem_kernel<float><<<grid, 256, 0, st>>>(x, w, scale, shift, (float *)out, n, H, W, Ho, Wo);
  else stem_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, w, scale, shift, (__nv_bfloat16 *)out, n, H, W, Ho, Wo);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

template <typename T>
static int depthwise_launch_t(const T *in, T *out, const float *w9c, const float *scale, const float *shift, int n,
                              int H, int W, int C, int stride, int dil, int act, cudaStream_t st) {
  constexpr int TW = 4;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;  // k = 3, pad = dil
  long long total = (long long)n * Ho * ((Wo + TW - 1) / TW) * (C / 8);
  int grid = grid_for(total, 256, 8);
  if (stride == 1 && dil == 1)
    depthwise3x3_kernel<T, 1, 1, TW><<<grid, 256, 0, st>>>(in, out, w9c, scale, shift, n, H, W, C, Ho, Wo, act);
  else if (stride == 2 && dil == 1)
    depthwise3x3_kernel<T, 2, 1, TW><<<grid, 256, 0, st>>>(in, out, w9c, scale, shift, n, H, W, C, Ho, Wo, act);
  else if (stride == 1 && dil == 2)
    depthwise3x3_kernel<T, 1, 2, TW><<<grid, 256, 0, st>>>(in, out, w9c, scale, shift, n, H, W, C, Ho, Wo, act);
  else {
    set_error("depthwise: unsupported stride %d / dilation %d", stride, dil);
    return LWP_EINVAL;
  }
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

int depthwise_launch(bool f32, const void *in, void *out, const float *w9c, const float *scale, const float *shift,
                     int n, int H, int W, int C, int stride, int dil, int act, cudaStream_t st) {
  if (f32)
    return depthwise_launch_t<float>((const float *)in, (float *)out, w9c, scale, shift, n, H, W, C, stride, dil, act,
                                     st);
  return depthwise_launch_t<__nv_bfloat16>((const __nv_bfloat16 *)in, (__nv_bfloat16 *)out, w9c, scale, shift, n, H, W,
                                           C, stride, dil, act, st);
}

int nhwc_to_nchw_launch(bool in_f32, const void *in, int ld, int c0, int c, float *out, int n, int HW,
                        cudaStream_t st) {
  long long total = (long long)n * c * HW;
  int grid = grid_for(total, 256, 8);
  if (in_f32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float *)in, ld, c0, c, out, n, HW);
  else nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)in, ld, c0, c, out, n, HW);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

}  // namespace lwp
