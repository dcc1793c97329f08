// Post-processing kernels: bit-exact cubic upsample, key-point extraction, PAF grouping.
//
// Everything here must produce the reference's bits, so all floating point goes through the
// explicitly rounded intrinsics (__fmul_rn, __dadd_rn, ...) which nvcc never contracts into FMAs.
//
// Reference (paths relative to the reference root):
//   upsample_cubic_kernel      cv2.resize INTER_CUBIC at demo.py:72,76; val.py:98-107
//   peak_candidates_kernel,
//   peak_nms_kernel,
//   keypoint_ids_kernel        modules/keypoints.py:16-48 (extract_keypoints) + id bookkeeping demo.py:95-98
//   paf_score_kernel           modules/keypoints.py:94-139 (line integral over the PAF)
//   limb_match_kernel          modules/keypoints.py:140-157 (sort + greedy one-to-one assignment)
//   pose_assemble_kernel       modules/keypoints.py:63-92,159-200 (sequential pose assembly + filter)
#include <type_traits>

#include "common.cuh"

#include <float.h>
#include <limits.h>
#include <string.h>

namespace lwp {

// modules/keypoints.py:5-8
__constant__ int c_kpt_ids[LWP_NUM_LIMBS][2] = {
    {1, 2}, {1, 5}, {2, 3}, {3, 4}, {5, 6}, {6, 7}, {1, 8}, {8, 9}, {9, 10}, {1, 11},
    {11, 12}, {12, 13}, {1, 0}, {0, 14}, {14, 16}, {0, 15}, {15, 17}, {2, 16}, {5, 17}};
__constant__ int c_paf_ids[LWP_NUM_LIMBS][2] = {
    {12, 13}, {20, 21}, {14, 15}, {16, 17}, {22, 23}, {24, 25}, {0, 1}, {2, 3}, {4, 5},
    {6, 7}, {8, 9}, {10, 11}, {28, 29}, {30, 31}, {34, 35}, {32, 33}, {36, 37}, {18, 19}, {26, 27}};

// ------------------------------------------------------------------------------------------------
// cubic upsample
// ------------------------------------------------------------------------------------------------

// OpenCV interpolateCubic(): A = -0.75, float, each operation rounded on its own.
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
  const float A = -0.75f;
  float t1 = __fadd_rn(t, 1.f);
  float v = __fsub_rn(__fmul_rn(A, t1), -3.75f);         // A*(t+1) - 5A
  v = __fadd_rn(__fmul_rn(v, t1), -6.0f);                // (..)*(t+1) + 8A
  c[0] = __fsub_rn(__fmul_rn(v, t1), -3.0f);             // (..)*(t+1) - 4A
  v = __fsub_rn(__fmul_rn(1.25f, t), 2.25f);             // (A+2)*t - (A+3)
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(v, t), t), 1.f);  // (..)*t*t + 1
  float u = __fsub_rn(1.f, t);
  v = __fsub_rn(__fmul_rn(1.25f, u), 2.25f);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(v, u), u), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

// source tap index (un-clamped, tap 1) and weights of destination coordinate d
__device__ __forceinline__ int cubic_axis(int d, double scale, float c[4]) {
  float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
  float fl = floorf(f);
  cubic_coeffs(__fsub_rn(f, fl), c);
  return (int)fl;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Where an up-sampled map comes from when it is not materialised: the stride-8 source + the resize geometry.
struct UpSrc {
  const float *src;   // [n][h][w] pixels, pixel stride ld floats; the resized channels start at the pointer
  int h, w, ld;
  int c_layout;       // channel count of the cv2.resize call being reproduced (decides the SIMD-tail columns)
  int H, W;           // up-sampled size
  double scale_x, scale_y;  // 1 / inv_scale
};

// horizontal 4-tap sum of HResizeCubic for one source row pointer (already offset to the channel)
__device__ __forceinline__ float cubic_hsum(const float *p, int i0, int i1, int i2, int i3, const float (&cx)[4],
                                            bool border) {
  float p0 = __fmul_rn(__ldg(p + i0), cx[0]);
  float p1 = __fmul_rn(__ldg(p + i1), cx[1]);
  float p2 = __fmul_rn(__ldg(p + i2), cx[2]);
  float p3 = __fmul_rn(__ldg(p + i3), cx[3]);
  float v = border ? __fadd_rn(0.f, p0) : p0;  // border columns start from v = 0 and add the products in turn
  v = __fadd_rn(v, p1);
  v = __fadd_rn(v, p2);
  return __fadd_rn(v, p3);
}

// vertical combination: VResizeCubicVec_32f body order, or the scalar-tail order for the last (W*C)%4 floats of a row
__device__ __forceinline__ float cubic_vsum(const float (&T)[4], const float (&cy)[4], bool simd_body) {
  float o;
  if (simd_body) {
    o = __fadd_rn(__fmul_rn(T[2], cy[2]), __fmul_rn(T[3], cy[3]));
    o = __fadd_rn(__fmul_rn(T[1], cy[1]), o);
    o = __fadd_rn(__fmul_rn(T[0], cy[0]), o);
  } else {
    o = __fadd_rn(__fmul_rn(T[0], cy[0]), __fmul_rn(T[1], cy[1]));
    o = __fadd_rn(o, __fmul_rn(T[2], cy[2]));
    o = __fadd_rn(o, __fmul_rn(T[3], cy[3]));
  }
  return o;
}

// NCH (1 or 2) channels of up-sampled pixel (e, d) of image img, computed from the source with OpenCV's bits.
template <int NCH>
__device__ __forceinline__ void upsampled_at(const UpSrc &u, int img, int e, int d, const int (&ch)[NCH],
                                             float (&out)[NCH]) {
  float cx[4], cy[4];
  const int sx = cubic_axis(d, u.scale_x, cx);
  const int sy = cubic_axis(e, u.scale_y, cy);
  const bool border = (sx < 1) || (sx + 2 >= u.w);
  const int i0 = clampi(sx - 1, 0, u.w - 1) * u.ld, i1 = clampi(sx, 0, u.w - 1) * u.ld;
  const int i2 = clampi(sx + 1, 0, u.w - 1) * u.ld, i3 = clampi(sx + 2, 0, u.w - 1) * u.ld;
  const int rowlen = u.W * u.c_layout;
  const int body = rowlen - (rowlen & 3);
  float T[NCH][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = clampi(sy - 1 + r, 0, u.h - 1);
    const float *p = u.src + ((size_t)img * u.h + row) * (size_t)u.w * u.ld;
#pragma unroll
    for (int c = 0; c < NCH; ++c) T[c][r] = cubic_hsum(p + ch[c], i0, i1, i2, i3, cx, border);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) out[c] = cubic_vsum(T[c], cy, d * u.c_layout + ch[c] < body);
}

// Same arithmetic as upsampled_at<2>, source = this block's copy of the two PAF channels in shared memory
// ([h][w] float2): one LDS.64 per tap serves both channels.
// cw4 != nullptr: the map is an exact x4 up-sampling in both directions, so coordinate d has source index (d - 2) >> 2 and
// the weights of phase (d + 2) & 3 -- cw4[phase] holds cubic_coeffs(0.125 + 0.25 phase), the very values cubic_axis()
// computes (its fraction is exactly that) -- instead of the float64 geometry + coefficient polynomials per sample.
__device__ __forceinline__ void upsampled_at_smem2(const UpSrc &u, const float2 *s, int e, int d, int ch0, int ch1,
                                                   float (&out)[2], const float4 *cw4 = nullptr) {
  float cx[4], cy[4];
  int sx, sy;
  if (cw4 != nullptr) {
    const float4 wx = cw4[(d + 2) & 3], wy = cw4[(e + 2) & 3];
    cx[0] = wx.x; cx[1] = wx.y; cx[2] = wx.z; cx[3] = wx.w;
    cy[0] = wy.x; cy[1] = wy.y; cy[2] = wy.z; cy[3] = wy.w;
    sx = (d - 2) >> 2; sy = (e - 2) >> 2;
  } else {
    sx = cubic_axis(d, u.scale_x, cx);
    sy = cubic_axis(e, u.scale_y, cy);
  }
  const bool border = (sx < 1) || (sx + 2 >= u.w);
  const int i0 = clampi(sx - 1, 0, u.w - 1), i1 = clampi(sx, 0, u.w - 1);
  const int i2 = clampi(sx + 1, 0, u.w - 1), i3 = clampi(sx + 2, 0, u.w - 1);
  const int rowlen = u.W * u.c_layout;
  const int body = rowlen - (rowlen & 3);
  float T[2][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float2 *p = s + clampi(sy - 1 + r, 0, u.h - 1) * u.w;
    const float2 a0 = p[i0], a1 = p[i1], a2 = p[i2], a3 = p[i3];
    float x0 = __fmul_rn(a0.x, cx[0]), x1 = __fmul_rn(a1.x, cx[1]), x2 = __fmul_rn(a2.x, cx[2]), x3 = __fmul_rn(a3.x, cx[3]);
    float y0 = __fmul_rn(a0.y, cx[0]), y1 = __fmul_rn(a1.y, cx[1]), y2 = __fmul_rn(a2.y, cx[2]), y3 = __fmul_rn(a3.y, cx[3]);
    float vx = border ? __fadd_rn(0.f, x0) : x0, vy = border ? __fadd_rn(0.f, y0) : y0;
    vx = __fadd_rn(vx, x1); vx = __fadd_rn(vx, x2); T[0][r] = __fadd_rn(vx, x3);
    vy = __fadd_rn(vy, y1); vy = __fadd_rn(vy, y2); T[1][r] = __fadd_rn(vy, y3);
  }
  out[0] = cubic_vsum(T[0], cy, d * u.c_layout + ch0 < body);
  out[1] = cubic_vsum(T[1], cy, d * u.c_layout + ch1 < body);
}

// Materialising resize: a block takes one output row and a span of kUpCols columns with all channels.  The column
// geometry (source index + 4 weights: double arithmetic + the coefficient polynomials) is computed once per column of the
// block instead of once per output float (19 / 38 channels share it), the row geometry once per block; the per-element
// work is 16 loads, 16 products and 15 sums in OpenCV's order -- same bits as upsampled_at<1>.
constexpr int kUpCols = 64;
__global__ void __launch_bounds__(256)
upsample_cubic_kernel(const UpSrc u, float *__restrict__ dst, int col_tiles, long long row_pitch, long long img_pitch,
                      float acc_div) {
  __shared__ __align__(16) float s_cx[kUpCols][4];
  __shared__ __align__(16) int s_i[kUpCols][4];      // clamped source columns, already multiplied by the pixel stride
  __shared__ unsigned char s_border[kUpCols];
  __shared__ float s_cy[4];
  __shared__ int s_row[4];
  const int tile = blockIdx.x % col_tiles;
  const long long re = blockIdx.x / col_tiles;   // img * H + e
  const int e = (int)(re % u.H), img = (int)(re / u.H);
  const int d0 = tile * kUpCols;
  const int ncol = min(kUpCols, u.W - d0);
  const int c = u.c_layout;
  if (threadIdx.x < ncol) {
    float cx[4];
    const int sx = cubic_axis(d0 + threadIdx.x, u.scale_x, cx);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s_cx[threadIdx.x][k] = cx[k];
      s_i[threadIdx.x][k] = clampi(sx - 1 + k, 0, u.w - 1) * u.ld;
    }
    s_border[threadIdx.x] = (sx < 1) || (sx + 2 >= u.w);
  } else if (threadIdx.x == 255) {
    float cy[4];
    const int sy = cubic_axis(e, u.scale_y, cy);
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_cy[k] = cy[k]; s_row[k] = clampi(sy - 1 + k, 0, u.h - 1); }
  }
  __syncthreads();
  const float cy[4] = {s_cy[0], s_cy[1], s_cy[2], s_cy[3]};
  // the source may be a cropped view of a larger map: rows row_pitch floats apart, images img_pitch floats apart
  const float *img_src = u.src + (size_t)img * img_pitch;
  const float *r0 = img_src + (size_t)s_row[0] * row_pitch, *r1 = img_src + (size_t)s_row[1] * row_pitch;
  const float *r2 = img_src + (size_t)s_row[2] * row_pitch, *r3 = img_src + (size_t)s_row[3] * row_pitch;
  const int rowlen = u.W * c, body = rowlen - (rowlen & 3);
  float *out = dst + ((size_t)re * u.W + d0) * c;
  const int q0 = d0 * c;
  for (int q = threadIdx.x; q < ncol * c; q += 256) {
    const int dl = q / c, ch = q - dl * c;
    const float4 cx = *reinterpret_cast<const float4 *>(s_cx[dl]);
    const int4 ii = *reinterpret_cast<const int4 *>(s_i[dl]);
    const bool border = s_border[dl] != 0;
    const float cxv[4] = {cx.x, cx.y, cx.z, cx.w};
    float T[4];
    const float *rows[4] = {r0 + ch, r1 + ch, r2 + ch, r3 + ch};
#pragma unroll
    for (int r = 0; r < 4; ++r) T[r] = cubic_hsum(rows[r], ii.x, ii.y, ii.z, ii.w, cxv, border);
    const float v = cubic_vsum(T, cy, q0 + q < body);
    // acc_div != 0: dst += v / acc_div, the running average of val.py:101,108 (avg += resized / len(scales)) in float32
    out[q] = acc_div != 0.f ? __fadd_rn(out[q], __fdiv_rn(v, acc_div)) : v;
  }
}

// Separable form of the same resize for c <= kUpMaxC channels: a block takes kUpRows output rows of a kUpCols-column
// span and streams down them.  Each source row the rows need is resized horizontally ONCE (HResizeCubic order) into a
// ring of four shared-memory rows -- the four rows of a vertical window r .. r + 3 fall into four different slots
// r & 3 -- and every output row is a vertical combination (VResizeCubicVec_32f / scalar-tail order) of the ring:
// 4 global loads per horizontally-resized value, amortised over the output rows that share the source row, plus 4
// shared-memory loads per output, instead of 16 global loads per output.  Operation for operation the arithmetic of
// upsample_cubic_kernel: same bits.  Used for up-scaling by 2 or more (val.py:98-108: the x8 map -> frame size resize of
// the scales below the frame size).
constexpr int kUpRows = 32, kUpMaxC = 38;
__global__ void __launch_bounds__(256)
upsample_cubic_rows_kernel(const UpSrc u, float *__restrict__ dst, int col_tiles, int row_tiles, long long row_pitch,
                           long long img_pitch, float acc_div) {
  __shared__ __align__(16) float s_cx[kUpCols][4];
  __shared__ __align__(16) int s_i[kUpCols][4];      // clamped source columns, already multiplied by the pixel stride
  __shared__ unsigned char s_border[kUpCols];
  __shared__ __align__(16) float s_cy[kUpRows][4];
  __shared__ int s_sy[kUpRows];
  __shared__ float s_T[4][kUpCols * kUpMaxC];          // ring of horizontally resized source rows
  const int tile = blockIdx.x % col_tiles;
  const int rt = (blockIdx.x / col_tiles) % row_tiles;
  const int img = blockIdx.x / (col_tiles * row_tiles);
  const int d0 = tile * kUpCols, e0 = rt * kUpRows;
  const int ncol = min(kUpCols, u.W - d0), nrow = min(kUpRows, u.H - e0);
  const int c = u.c_layout;
  if (threadIdx.x < ncol) {
    float cx[4];
    const int sx = cubic_axis(d0 + threadIdx.x, u.scale_x, cx);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s_cx[threadIdx.x][k] = cx[k];
      s_i[threadIdx.x][k] = clampi(sx - 1 + k, 0, u.w - 1) * u.ld;
    }
    s_border[threadIdx.x] = (sx < 1) || (sx + 2 >= u.w);
  } else if (threadIdx.x >= 128 && threadIdx.x < 128 + nrow) {
    const int t = threadIdx.x - 128;
    s_sy[t] = cubic_axis(e0 + t, u.scale_y, s_cy[t]);
  }
  __syncthreads();
  const float *img_src = u.src + (size_t)img * img_pitch;   // (possibly a cropped view: rows row_pitch floats apart)
  const int rowlen = u.W * c, body = rowlen - (rowlen & 3);
  const int nq = ncol * c, q0 = d0 * c;
  int next_row = INT_MIN;   // first (un-clamped) source row that is not in the ring yet; block-uniform
  for (int t = 0; t < nrow; ++t) {
    const int sy = s_sy[t];
    // horizontal pass of the window rows sy - 1 .. sy + 2 that are new (source rows are needed in increasing order)
    for (int r = max(next_row, sy - 1); r <= sy + 2; ++r) {
      const float *rp = img_src + (size_t)clampi(r, 0, u.h - 1) * row_pitch;
      float *Tr = s_T[r & 3];
      for (int q = threadIdx.x; q < nq; q += 256) {
        const int dl = q / c, ch = q - dl * c;
        const float4 cx = *reinterpret_cast<const float4 *>(s_cx[dl]);
        const int4 ii = *reinterpret_cast<const int4 *>(s_i[dl]);
        const float cxv[4] = {cx.x, cx.y, cx.z, cx.w};
        Tr[q] = cubic_hsum(rp + ch, ii.x, ii.y, ii.z, ii.w, cxv, s_border[dl] != 0);
      }
    }
    next_row = sy + 3;
    __syncthreads();
    const float4 cy4 = *reinterpret_cast<const float4 *>(s_cy[t]);
    const float cy[4] = {cy4.x, cy4.y, cy4.z, cy4.w};
    const float *T0 = s_T[(sy - 1) & 3], *T1 = s_T[sy & 3], *T2 = s_T[(sy + 1) & 3], *T3 = s_T[(sy + 2) & 3];
    float *out = dst + (((size_t)img * u.H + e0 + t) * u.W + d0) * c;
    for (int q = threadIdx.x; q < nq; q += 256) {
      const float T[4] = {T0[q], T1[q], T2[q], T3[q]};
      const float v = cubic_vsum(T, cy, q0 + q < body);
      out[q] = acc_div != 0.f ? __fadd_rn(out[q], __fdiv_rn(v, acc_div)) : v;
    }
    __syncthreads();   // the next row's horizontal pass may overwrite a ring slot this row still read
  }
}

// Integer ratios that are powers of two (the x4 of demo.py:72,76 and the x8 of val.py:98,105): output pixel e = R k + R/2 + j
// (j = 0..R-1) has source index k and fraction (j + 0.5) / R exactly, whatever k -- the geometry is periodic, so a thread
// takes ONE channel of one source cell (ky, kx), loads its 4 x 4 source patch once and produces the R x R outputs that
// share it: 16 loads and ~11 flops per output instead of 16 loads and 31 flops + address arithmetic per output.  The
// weights come from the same cubic_axis() / cubic_coeffs() as everywhere else (evaluated for the R phases by the first
// threads of the block), the operation order per output is that of upsampled_at<1>: same bits.
template <int R>
__global__ void __launch_bounds__(256)
upsample_cubic_pow2_kernel(const UpSrc u, float *__restrict__ dst, long long row_pitch, long long img_pitch, long long total) {
  __shared__ float s_c[R][4];   // weights of phase j (identical for x and y: same ratio, same function)
  if (threadIdx.x < R) {
    float c[4];
    (void)cubic_axis(R / 2 + threadIdx.x, u.scale_x, c);   // output R/2 + j: source index 0, fraction (j + 0.5) / R
#pragma unroll
    for (int k = 0; k < 4; ++k) s_c[threadIdx.x][k] = c[k];
  }
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = u.c_layout;
  const int ch = (int)(idx % c);
  long long t = idx / c;
  const int cw = u.w + 1, chh = u.h + 1;
  const int kx = (int)(t % cw) - 1;
  t /= cw;
  const int ky = (int)(t % chh) - 1;
  const int img = (int)(t / chh);
  const bool border = (kx < 1) || (kx + 2 >= u.w);
  const float *img_src = u.src + (size_t)img * img_pitch + ch;
  int ix[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ix[k] = clampi(kx - 1 + k, 0, u.w - 1) * u.ld;
  // horizontal pass: T[r][j] for the 4 source rows and the R output columns of this cell
  float T[4][R];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float *p = img_src + (size_t)clampi(ky - 1 + r, 0, u.h - 1) * row_pitch;
    const float a0 = __ldg(p + ix[0]), a1 = __ldg(p + ix[1]), a2 = __ldg(p + ix[2]), a3 = __ldg(p + ix[3]);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const float p0 = __fmul_rn(a0, s_c[j][0]), p1 = __fmul_rn(a1, s_c[j][1]), p2 = __fmul_rn(a2, s_c[j][2]), p3 = __fmul_rn(a3, s_c[j][3]);
      float v = border ? __fadd_rn(0.f, p0) : p0;
      v = __fadd_rn(v, p1);
      v = __fadd_rn(v, p2);
      T[r][j] = __fadd_rn(v, p3);
    }
  }
  const int rowlen = u.W * c, body = rowlen - (rowlen & 3);
  const int e0 = R * ky + R / 2, d0 = R * kx + R / 2;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int e = e0 + i;
    if (e < 0 || e >= u.H) continue;
    const float cy[4] = {s_c[i][0], s_c[i][1], s_c[i][2], s_c[i][3]};
    float *orow = dst + ((size_t)img * u.H + e) * (size_t)rowlen + ch;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int d = d0 + j;
      if (d < 0 || d >= u.W) continue;
      const float Tc[4] = {T[0][j], T[1][j], T[2][j], T[3][j]};
      orow[(size_t)d * c] = cubic_vsum(Tc, cy, d * c + ch < body);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// extract_keypoints
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float thr01(float v) { return v < 0.1f ? 0.f : v; }  // heatmap[heatmap < 0.1] = 0

// One thread per (pixel, channel): thresholded value strictly greater than its 4 axial neighbours
// (zero outside the image).  Candidates go to an unordered per-(image, channel) list.
__global__ void __launch_bounds__(256)
peak_candidates_kernel(const float *__restrict__ hm, int H, int W, int ld, int n_ch,
                       unsigned long long *__restrict__ cand, int *__restrict__ cand_count, int cap,
                       long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(idx % n_ch);
    long long t = idx / n_ch;
    int x = (int)(t % W);
    t /= W;
    int y = (int)(t % H);
    int img = (int)(t / H);
    const float *p = hm + (((size_t)img * H + y) * W + x) * (size_t)ld + ch;
    float v = thr01(__ldg(p));
    if (!(v > 0.f)) continue;  // a peak must beat neighbours that are >= 0
    float l = x > 0 ? thr01(__ldg(p - ld)) : 0.f;
    float r = x + 1 < W ? thr01(__ldg(p + ld)) : 0.f;
    float u = y > 0 ? thr01(__ldg(p - (size_t)W * ld)) : 0.f;
    float dn = y + 1 < H ? thr01(__ldg(p + (size_t)W * ld)) : 0.f;
    if (v > l && v > r && v > u && v > dn) {
      int slot = atomicAdd(&cand_count[img * n_ch + ch], 1);
      if (slot < cap) {
        unsigned long long key = ((unsigned long long)(((unsigned)x << 16) | (unsigned)y) << 32) | __float_as_uint(v);
        cand[((size_t)img * n_ch + ch) * cap + slot] = key;
      }
    }
  }
}

// Fused variant: the up-sampled heat-map is never written to memory.  A block rebuilds a tile of
// 30 columns x 32 rows (+1 halo) of the up-sampled map for all scanned channels straight from the
// stride-8 source -- separable cubic with exactly the operation order of upsample_cubic_kernel, so the
// values are bit-identical to the materialised map.  Lanes own up-sampled columns (each lane keeps its
// column's cubic weights in registers), warps own channels: the horizontal pass goes source -> smem,
// the vertical pass streams down the rows keeping three consecutive rows in registers, so the up/down
// neighbours are registers and the left/right neighbours are warp shuffles.
constexpr int kPkCols = 30, kPkRows = 32, kPkSrcMax = 16, kPkWarps = 6, kPkMaxCh = 24;

__global__ void __launch_bounds__(kPkWarps * 32)
peak_candidates_fused_kernel(const UpSrc u, int n_ch, unsigned long long *__restrict__ cand,
                             int *__restrict__ cand_count, int cap, int *__restrict__ overflow) {
  __shared__ __align__(16) float s_cy[kPkRows + 2][4];
  __shared__ int s_sy[kPkRows + 2];
  __shared__ int s_win[4];  // sx_lo, ncols, sy_lo, nrows
  __shared__ int s_cmax[kPkMaxCh];  // per channel: max |source value| over the tile's source window (float bits)
  extern __shared__ float pk_smem[];
  float *s_src = pk_smem;                                   // [nrows][n_ch][kPkSrcMax]
  float *s_T = s_src + kPkSrcMax * n_ch * kPkSrcMax + (threadIdx.x >> 5) * (kPkSrcMax * 32);   // this warp's [nrows][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (u.W + kPkCols - 1) / kPkCols;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int img = blockIdx.y;
  const int ox0 = tx * kPkCols, oy0 = ty * kPkRows;
  constexpr int kInvalid = -(1 << 30);
  // An up-sampled value is a combination of 16 source values whose weights sum to at most 1.375^2 = 1.890625 in
  // absolute value (Keys cubic, A = -0.75, any phase), so a channel whose source window stays below 0.1 / 1.890625
  // cannot reach the 0.1 threshold anywhere in this tile: its two passes are skipped (same candidates, less work).
  constexpr float kHopeless = 0.0528f;   // a little under 0.052893 (rounding of the fp32 sums)
  if (tid < kPkMaxCh) s_cmax[tid] = 0;
  // this lane's up-sampled column (lanes 0 and 31 are halo only)
  const int d = ox0 - 1 + lane;
  const bool col_ok = d >= 0 && d < u.W;
  float cx[4] = {0.f, 0.f, 0.f, 0.f};
  const int sx = col_ok ? cubic_axis(d, u.scale_x, cx) : kInvalid;
  if (tid < kPkRows + 2) {
    const int e = oy0 - 1 + tid;
    int sy = kInvalid;
    if (e >= 0 && e < u.H) sy = cubic_axis(e, u.scale_y, s_cy[tid]);
    s_sy[tid] = sy;
  }
  // source window: sx is monotone in d, so the first / last valid lane give the range
  {
    const unsigned okm = __ballot_sync(0xffffffffu, col_ok);
    const int lo = __shfl_sync(0xffffffffu, sx, __ffs(okm) - 1) - 1;
    const int hi = __shfl_sync(0xffffffffu, sx, 31 - __clz(okm)) + 2;
    if (tid == 0) { s_win[0] = lo; s_win[1] = hi - lo + 1; }
  }
  __syncthreads();
  if (tid == 0) {
    const int if_ = oy0 == 0 ? 1 : 0, il = min(kPkRows + 1, u.H - oy0);
    s_win[2] = s_sy[if_] - 1; s_win[3] = s_sy[il] + 2 - s_win[2] + 1;
  }
  __syncthreads();
  const int sx_lo = s_win[0], ncols = s_win[1], sy_lo = s_win[2], nrows = s_win[3];
  if (ncols > kPkSrcMax || nrows > kPkSrcMax) {  // host guarantees ratio >= 3; never silently wrong
    if (tid == 0) overflow[img] = 1;
    return;
  }
  // 1. source window -> smem [ry][c][rx]; replicate border by clamping the coordinates.  A thread owns one (column,
  // channel) pair -- channel fastest, so a warp reads runs of n_ch consecutive floats -- and walks down the rows: one
  // clamp, one load, one store and one integer max per element (the flat-index version with its mixed-radix carries
  // and a conditional atomic per element was a fifth of the kernel's instructions).
  {
    const float *src_img = u.src + (size_t)img * u.h * u.w * u.ld;
    const int npairs = n_ch * ncols;
    const size_t row_pitch = (size_t)u.w * u.ld;
    for (int pr = tid; pr < npairs; pr += blockDim.x) {
      const int rx = pr / n_ch, c = pr - rx * n_ch;
      const float *colp = src_img + (size_t)clampi(sx_lo + rx, 0, u.w - 1) * u.ld + c;
      float *dst = s_src + c * kPkSrcMax + rx;
      int mbits = 0;   // max |value| as float bits: non-negative floats order like their bit patterns (NaN: never skipped)
#pragma unroll 4
      for (int ry = 0; ry < nrows; ++ry) {
        const float val = __ldg(colp + (size_t)clampi(sy_lo + ry, 0, u.h - 1) * row_pitch);
        dst[ry * n_ch * kPkSrcMax] = val;
        mbits = max(mbits, __float_as_int(fabsf(val)));
      }
      atomicMax(&s_cmax[c], mbits);
    }
  }
  __syncthreads();
  // 2. horizontal pass: T[ry][c][lane]
  const bool border = (sx < 1) || (sx + 2 >= u.w);
  const int rx0 = col_ok ? sx - 1 - sx_lo : 0;
  const int rowlen = u.W * u.c_layout, body = rowlen - (rowlen & 3);
  // x4 in y with an exact 1/4 scale: the row geometry is periodic (see the vertical pass); weights of the four phases
  const bool periodic4 = u.scale_y == 0.25 && u.H == 4 * u.h;
  float cy4p[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cubic_coeffs(0.125f + 0.25f * (float)j, cy4p[j]);
  for (int c = warp; c < n_ch; c += kPkWarps) {   // a warp takes its channels one at a time: both passes, then the next
    if (__int_as_float(s_cmax[c]) < kHopeless) continue;
    __syncwarp();   // the previous channel's vertical pass has finished reading s_T
    for (int ry = 0; ry < nrows; ++ry) {
      const float *p = s_src + ((size_t)ry * n_ch + c) * kPkSrcMax + rx0;
      float v = 0.f;
      if (col_ok) {
        float p0 = __fmul_rn(p[0], cx[0]), p1 = __fmul_rn(p[1], cx[1]);
        float p2 = __fmul_rn(p[2], cx[2]), p3 = __fmul_rn(p[3], cx[3]);
        v = border ? __fadd_rn(0.f, p0) : p0;
        v = __fadd_rn(v, p1);
        v = __fadd_rn(v, p2);
        v = __fadd_rn(v, p3);
      }
      s_T[ry * 32 + lane] = v;
    }
    __syncwarp();
    // 3. vertical pass streamed down the rows + strict 4-neighbour test on the row in the middle
    const bool simd_body = d * u.c_layout + c < body;
    float v0 = 0.f, v1 = 0.f;  // rows i-2, i-1
    // centre row e_c (values v0 / v1 / v2 = rows e_c - 1, e_c, e_c + 1): interior lanes only; rows without any value >= 0.1
    // (the common case) skip the neighbour exchange altogether
    auto test_centre = [&](int e_c, float up_v, float mid_v, float down_v) {
      if (__any_sync(0xffffffffu, mid_v > 0.f)) {
        const float vl = __shfl_up_sync(0xffffffffu, mid_v, 1), vr = __shfl_down_sync(0xffffffffu, mid_v, 1);
        if (lane >= 1 && lane <= kPkCols && col_ok && mid_v > 0.f && mid_v > vl && mid_v > vr && mid_v > up_v &&
            mid_v > down_v && e_c < u.H) {
          int slot = atomicAdd(&cand_count[img * n_ch + c], 1);
          if (slot < cap) {
            unsigned long long key =
                ((unsigned long long)(((unsigned)d << 16) | (unsigned)e_c) << 32) | __float_as_uint(mid_v);
            cand[((size_t)img * n_ch + c) * cap + slot] = key;
          }
        }
      }
    };
    if (periodic4 && __all_sync(0xffffffffu, simd_body || !col_ok)) {
      // x4 (demo.py:72): output row e = 4 k + 2 + j has source row k and fraction (j + 0.5) / 4 exactly, whatever k, so
      // the four weight sets are constants and the loop needs neither the per-row geometry tables nor a source-row
      // comparison: per source row one load shifts the window, then four output rows of 4 products + 3 sums each
      // (VResizeCubicVec_32f body order -- the scalar-tail columns, if the map has any, take the generic loop below).
      // Lanes outside the image hold T = 0 and so produce 0; rows outside the image are forced to 0 by the row test;
      // the neighbour exchange is voted on once per group of four rows.
      const int k0 = (oy0 >> 2) - 1;                      // source row of the tile's first row e = oy0 - 1 (phase 1)
      const float *tp = s_T + (k0 - 1 - sy_lo) * 32 + lane;
      float T0 = tp[0], T1 = tp[32], T2 = tp[64], T3 = tp[96];
      const int e_lo = oy0 > 0 ? oy0 - 1 : 0;
      const unsigned e_span = (unsigned)(min(u.H - 1, oy0 + kPkRows) - e_lo);   // valid rows: e_lo .. e_lo + e_span
      float pm2 = 0.f, pm1 = 0.f;                         // rows e - 2, e - 1 of the current group
      auto centre4 = [&](int e_c, float up_v, float mid_v, float down_v) {
        const float vl = __shfl_up_sync(0xffffffffu, mid_v, 1), vr = __shfl_down_sync(0xffffffffu, mid_v, 1);
        if (lane >= 1 && lane <= kPkCols && mid_v > 0.f && mid_v > vl && mid_v > vr && mid_v > up_v && mid_v > down_v &&
            e_c >= oy0 && e_c < oy0 + kPkRows) {
          int slot = atomicAdd(&cand_count[img * n_ch + c], 1);
          if (slot < cap) {
            unsigned long long key =
                ((unsigned long long)(((unsigned)d << 16) | (unsigned)e_c) << 32) | __float_as_uint(mid_v);
            cand[((size_t)img * n_ch + c) * cap + slot] = key;
          }
        }
      };
      // (the shared-memory address of the window's next row is kept as an integer: with a generic pointer the compiler
      // rebuilt the shared window base from SR_CgaCtaId in every iteration)
      const uint32_t tq0 = (uint32_t)__cvta_generic_to_shared(tp) + 3u * 32u * 4u;
      auto run_groups = [&](auto all_valid_tag) {
        constexpr bool kAllValid = decltype(all_valid_tag)::value;   // every row of the tile (+ halo) lies inside the image
        uint32_t tq = tq0;
#pragma unroll 1
        for (int kk = 0; kk < 9; ++kk) {                  // group kk: rows e .. e + 3, e = oy0 - 2 + 4 kk (source row k0 + kk)
          if (kk > 0) {
            T0 = T1; T1 = T2; T2 = T3;
            tq += 128u;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(T3) : "r"(tq));
          }
          const int e = oy0 - 2 + 4 * kk;
          float a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float o = __fadd_rn(__fmul_rn(T2, cy4p[j][2]), __fmul_rn(T3, cy4p[j][3]));
            o = __fadd_rn(__fmul_rn(T1, cy4p[j][1]), o);
            o = __fadd_rn(__fmul_rn(T0, cy4p[j][0]), o);
            // thr01 + zero outside the image
            a[j] = ((kAllValid || (unsigned)(e + j - e_lo) <= e_span) && !(o < 0.1f)) ? o : 0.f;
          }
          // centres e - 1 .. e + 2 (rows oy0 - 2 and oy0 + 33 are only ever neighbours of rows that are not centres)
          if (__any_sync(0xffffffffu, fmaxf(fmaxf(pm1, a[0]), fmaxf(a[1], a[2])) > 0.f)) {
            centre4(e - 1, pm2, pm1, a[0]);
            centre4(e, pm1, a[0], a[1]);
            centre4(e + 1, a[0], a[1], a[2]);
            centre4(e + 2, a[1], a[2], a[3]);
          }
          pm2 = a[2]; pm1 = a[3];
        }
      };
      if (oy0 >= 1 && oy0 + kPkRows <= u.H - 1) run_groups(std::true_type{});
      else run_groups(std::false_type{});
      continue;
    }
    // the four horizontally-resized source rows of the current output row stay in registers: consecutive output
    // rows share them (x4: four rows per source row), a step of one source row shifts the window by one load
    float T[4] = {0.f, 0.f, 0.f, 0.f};
    int cur_sy = kInvalid;
    for (int i = 0; i < kPkRows + 2; ++i) {
      const int sy = s_sy[i];                   // warp-uniform
      float v2 = 0.f;
      if (sy != kInvalid) {
        if (sy != cur_sy) {
          const float *t = s_T + (sy - 1 - sy_lo) * 32 + lane;
          if (sy == cur_sy + 1) {
            T[0] = T[1]; T[1] = T[2]; T[2] = T[3]; T[3] = t[3 * 32];
          } else {
            T[0] = t[0]; T[1] = t[32]; T[2] = t[2 * 32]; T[3] = t[3 * 32];
          }
          cur_sy = sy;
        }
        const float4 cy4 = *reinterpret_cast<const float4 *>(s_cy[i]);
        const float cyv[4] = {cy4.x, cy4.y, cy4.z, cy4.w};
        if (col_ok) v2 = thr01(cubic_vsum(T, cyv, simd_body));
      }
      // centre = row i-1 (tile rows are i-1 in [1, kPkRows])
      if (i >= 2) test_centre(oy0 - 1 + (i - 1), v0, v1, v2);
      v0 = v1; v1 = v2;
    }
  }
}

template <typename T, typename Less>
__device__ void bitonic_sort_smem(T *a, int n_pow2, Less less) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          T x = a[i], y = a[ixj];
          bool up = (i & k) == 0;
          if (up ? less(y, x) : less(x, y)) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

struct U64Less {
  __device__ bool operator()(unsigned long long a, unsigned long long b) const { return a < b; }
};

// One block per (image, channel): sort candidates by (x, y), greedy radius-6 suppression in that
// order (suppressed points never suppress), emit survivors in order.
__global__ void __launch_bounds__(256)
peak_nms_kernel(const unsigned long long *__restrict__ cand, const int *__restrict__ cand_count, int cap_cand,
                int n_ch, lwp_keypoint *__restrict__ kpts, int *__restrict__ counts, int cap_kpts,
                int *__restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int slot = blockIdx.x;  // img * n_ch + ch
  const int img = slot / n_ch;
  int cnt = cand_count[slot];
  if (cnt > cap_cand) {
    if (threadIdx.x == 0) overflow[img] = 1;
    cnt = cap_cand;
  }
  int P = 32;
  while (P < cnt) P <<= 1;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
  unsigned char *sup = reinterpret_cast<unsigned char *>(keys + P);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    keys[i] = i < cnt ? cand[(size_t)slot * cap_cand + i] : ~0ull;
    sup[i] = 0;
  }
  __syncthreads();
  bitonic_sort_smem(keys, P, U64Less());
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  for (int i = 0; i < cnt; ++i) {
    if (sup[i]) continue;  // warp-uniform: written before the last __syncwarp
    unsigned ki = (unsigned)(keys[i] >> 32);
    int xi = ki >> 16, yi = ki & 0xffff;
    for (int j = i + 1 + lane; j < cnt; j += 32) {
      unsigned kj = (unsigned)(keys[j] >> 32);
      int dx = (int)(kj >> 16) - xi;
      if (dx >= 6) break;  // sorted by x: nothing further can be within radius 6
      int dy = (int)(kj & 0xffff) - yi;
      if (dx * dx + dy * dy < 36) sup[j] = 1;
    }
    __syncwarp();
  }
  int out_n = 0;
  lwp_keypoint *out = kpts + (size_t)slot * cap_kpts;
  for (int base = 0; base < cnt; base += 32) {
    int j = base + lane;
    bool alive = j < cnt && !sup[j];
    unsigned m = __ballot_sync(0xffffffffu, alive);
    int pos = out_n + __popc(m & ((1u << lane) - 1));
    if (alive && pos < cap_kpts) {
      unsigned long long k = keys[j];
      unsigned kk = (unsigned)(k >> 32);
      lwp_keypoint kp;
      kp.x = kk >> 16; kp.y = kk & 0xffff; kp.score = __uint_as_float((unsigned)k); kp.id = 0;
      out[pos] = kp;
    }
    out_n += __popc(m);
  }
  if (lane == 0) {
    if (out_n > cap_kpts) { overflow[img] = 1; out_n = cap_kpts; }
    counts[slot] = out_n;
  }
}

// One block per image: exclusive prefix of the per-channel counts, then global ids.
__global__ void keypoint_ids_kernel(lwp_keypoint *__restrict__ kpts, const int *__restrict__ counts,
                                    int *__restrict__ kpt_start, int n_ch, int cap_kpts) {
  __shared__ int s_start[65];
  const int img = blockIdx.x;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < n_ch; ++c) { s_start[c] = acc; acc += counts[img * n_ch + c]; }
    s_start[n_ch] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c <= n_ch; c += blockDim.x) kpt_start[img * (n_ch + 1) + c] = s_start[c];
  for (int c = 0; c < n_ch; ++c) {
    int cnt = s_start[c + 1] - s_start[c];
    for (int j = threadIdx.x; j < cnt; j += blockDim.x)
      kpts[((size_t)img * n_ch + c) * cap_kpts + j].id = s_start[c] + j;
  }
}

// ------------------------------------------------------------------------------------------------
// group_keypoints
// ------------------------------------------------------------------------------------------------

struct Conn {   // candidate connection of one limb: indices into the A / B key-point lists
  double ratio;
  int i, j;
};
struct Match {  // accepted connection: global key-point ids, their scores, the PAF ratio
  double ratio;
  int ida, idb;
  float sa, sb;
};

struct ConnBefore {  // ratio descending; ties keep (i, j) generation order (stable sort, reverse=True)
  __device__ bool operator()(const Conn &a, const Conn &b) const {
    if (a.ratio > b.ratio) return true;
    if (a.ratio < b.ratio) return false;
    if (a.i != b.i) return a.i < b.i;
    return a.j < b.j;
  }
};

// One sample of the line integral: point k of linspace2d(a, b, 10) (modules/keypoints.py:11-13,119-131), the PAF
// vector there (on-the-fly cubic from the stride-8 map when kFused) projected on the unit limb vector.
template <bool kFused, bool kSmem>
__device__ __forceinline__ double paf_sample(const float *paf, int W, int ld, const UpSrc &up, const float2 *s_src, int img,
                                             int demo, int ax, int ay, int vx, int vy, double ux, double uy, int k, int cx,
                                             int cy, const float4 *cw4 = nullptr) {
  const double fx = __dadd_rn(__dmul_rn(__dmul_rn(1.0 / 9, (double)vx), (double)k), (double)ax);
  const double fy = __dadd_rn(__dmul_rn(__dmul_rn(1.0 / 9, (double)vy), (double)k), (double)ay);
  const int ix = demo ? __double2int_rz(fx) : __double2int_rn(fx);
  const int iy = demo ? __double2int_rz(fy) : __double2int_rn(fy);
  float pv[2];
  if constexpr (kFused) {  // the up-sampled PAF pixel is computed on the fly, bit-identical to the materialised map
    if constexpr (kSmem) {
      upsampled_at_smem2(up, s_src, iy, ix, cx, cy, pv, cw4);
    } else {
      const int ch[2] = {cx, cy};
      upsampled_at<2>(up, img, iy, ix, ch, pv);
    }
  } else {
    const float *pp = paf + ((size_t)iy * W + ix) * ld;
    pv[0] = __ldg(pp + cx); pv[1] = __ldg(pp + cy);
  }
  return __dadd_rn(__dmul_rn(ux, (double)pv[0]), __dmul_rn(uy, (double)pv[1]));
}

// Re-pack of the stride-8 PAF channels for paf_score_kernel<true, true>: [img][pixel][ld] float (38 PAF channels among the
// 64 of a head row) -> [img][limb][pixel] float2 = the limb's (x, y) channel pair.  Every (limb, image) block of the
// scoring kernel stages its two channels in shared memory; read straight from the head rows that is 8 useful bytes
// per 32-byte sector at a 256-byte pitch, 1 216 blocks x 241 KB of sector traffic per 64-frame batch, and it was 29 %
// of that kernel's stall samples.  Packed, the same staging is one contiguous 30 KB copy.
constexpr int kPackPx = 128;
__global__ void __launch_bounds__(256)
paf_pack_kernel(const UpSrc up, float2 *__restrict__ packed, int plane_stride) {
  __shared__ float s[kPackPx][39];   // 38 channels + 1 pad: conflict-free column reads
  const int img = blockIdx.y, px0 = blockIdx.x * kPackPx, npx = up.h * up.w;
  const float *src = up.src + ((size_t)img * npx + px0) * up.ld;
  for (int idx = threadIdx.x; idx < kPackPx * 38; idx += blockDim.x) {
    const int p = idx / 38, c = idx - p * 38;
    if (px0 + p < npx) s[p][c] = __ldg(src + (size_t)p * up.ld + c);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kPackPx * LWP_NUM_LIMBS; idx += blockDim.x) {
    const int limb = idx / kPackPx, p = idx - limb * kPackPx;
    if (px0 + p < npx)
      packed[((size_t)img * LWP_NUM_LIMBS + limb) * plane_stride + px0 + p] =
          make_float2(s[p][c_paf_ids[limb][0]], s[p][c_paf_ids[limb][1]]);
  }
}

// PAF line integral in two phases.  A connection needs at least 9 of its 10 samples above min_paf_score
// (success_ratio > 0.8, keypoints.py:137), so a pair with two failing samples can never be one: phase 1 gives every lane
// one candidate pair and evaluates up to four probe samples (k = 3, 6, 1, 8), stopping at the second failure (on sparse
// PAFs that is after two probes; on the noisy maps of a random-init network four probes leave ~13 % of the pairs
// where two left ~58 %);
// phase 2 scores the surviving pairs in full, 3 at a time, 10 lanes per pair, lane k taking sample k of linspace2d;
// the sum is then re-done in k order so the float64 result is the reference's.  Which pairs are connections and
// their ratios are exactly the reference's; only work on hopeless pairs is skipped.
// grid = (blocks per limb, 19 limbs, images).
template <bool kFused, bool kSmem>   // kSmem (fused only): the limb's two stride-8 PAF channels are staged in shared memory
__global__ void __launch_bounds__(256)
paf_score_kernel(const lwp_keypoint *__restrict__ kpts, const int *__restrict__ counts, int cap_kpts,
                 const float *__restrict__ pafs, int H, int W, int ld, const UpSrc up, int demo,
                 double min_paf_score, Conn *__restrict__ conn, int *__restrict__ conn_count, int cap_conn,
                 const float2 *__restrict__ packed, int plane_stride) {
  const int limb = blockIdx.y, img = blockIdx.z;
  const int ka = c_kpt_ids[limb][0], kb = c_kpt_ids[limb][1];
  const int nA = counts[img * LWP_NUM_KPT_TYPES + ka], nB = counts[img * LWP_NUM_KPT_TYPES + kb];
  if (nA == 0 || nB == 0) return;
  const lwp_keypoint *A = kpts + ((size_t)img * LWP_NUM_KPT_TYPES + ka) * cap_kpts;
  const lwp_keypoint *B = kpts + ((size_t)img * LWP_NUM_KPT_TYPES + kb) * cap_kpts;
  const int cx = c_paf_ids[limb][0], cy = c_paf_ids[limb][1];
  const float *paf = kFused ? nullptr : pafs + (size_t)img * H * W * ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int g = lane / 10, k = lane - g * 10;
  const int total = nA * nB;
  const double height_n = (double)(H / 2);  // pafs.shape[0] // 2
  Conn *out = conn + ((size_t)img * LWP_NUM_LIMBS + limb) * cap_conn;
  int *out_count = conn_count + img * LWP_NUM_LIMBS + limb;
  if (blockIdx.x * wpb * 32 >= total) return;   // no pair for this block
  extern __shared__ __align__(16) unsigned char ps_smem[];
  const float2 *s_src = reinterpret_cast<const float2 *>(ps_smem);
  __shared__ float4 s_cw4[4];
  const float4 *cw4 = nullptr;
  if constexpr (kFused && kSmem) {
    if (up.scale_x == 0.25 && up.scale_y == 0.25 && up.W == 4 * up.w && up.H == 4 * up.h) {   // x4: periodic geometry
      if (threadIdx.x < 4) {
        float c[4];
        cubic_coeffs(0.125f + 0.25f * (float)threadIdx.x, c);
        s_cw4[threadIdx.x] = make_float4(c[0], c[1], c[2], c[3]);
      }
      cw4 = s_cw4;
    }
    float2 *s_w = reinterpret_cast<float2 *>(ps_smem);
    if (packed != nullptr) {   // the limb's plane as paf_pack_kernel wrote it: one contiguous, fully coalesced copy
      const float4 *src4 = reinterpret_cast<const float4 *>(packed + ((size_t)img * LWP_NUM_LIMBS + limb) * plane_stride);
      float4 *dst4 = reinterpret_cast<float4 *>(ps_smem);
      for (int idx = threadIdx.x; idx < plane_stride / 2; idx += blockDim.x) dst4[idx] = __ldg(src4 + idx);
    } else {
      const float *src = up.src + (size_t)img * up.h * up.w * up.ld;
      for (int idx = threadIdx.x; idx < up.h * up.w; idx += blockDim.x) {
        const float *pp = src + (size_t)idx * up.ld;
        s_w[idx] = make_float2(__ldg(pp + cx), __ldg(pp + cy));
      }
    }
    __syncthreads();
  }

  for (int base = (blockIdx.x * wpb + warp) * 32; base < total; base += gridDim.x * wpb * 32) {
    // ---- phase 1: lane = pair, two probe samples ----
    const int p1 = base + lane;
    bool survive = false;
    // this lane's pair: kept for phase 2, whose ten lanes per pair fetch them by shuffle instead of redoing the two
    // loads, the integer division and the float64 square root + two divisions
    int p_i = 0, p_j = 0, p_ax = 0, p_ay = 0, p_vx = 0, p_vy = 0;
    double p_norm = 0.0, p_ux = 0.0, p_uy = 0.0;
    if (p1 < total) {
      const int i = p1 / nB, j = p1 - i * nB;
      const lwp_keypoint a = A[i], b = B[j];
      const int vx = b.x - a.x, vy = b.y - a.y;
      const double norm = __dsqrt_rn((double)((long long)vx * vx + (long long)vy * vy));
      p_i = i; p_j = j; p_ax = a.x; p_ay = a.y; p_vx = vx; p_vy = vy; p_norm = norm;
      if (norm != 0.0) {
        const double ux = __ddiv_rn((double)vx, norm), uy = __ddiv_rn((double)vy, norm);
        p_ux = ux; p_uy = uy;
        // up to four probes (k = 3, 6, 1, 8); the pair is dropped at its second failing sample
        int fails = 0;
#pragma unroll 1
        for (int q = 0; q < 4 && fails < 2; ++q) {
          const int kq = (0x8163 >> (4 * q)) & 15;
          if (!(paf_sample<kFused, kSmem>(paf, W, ld, up, s_src, img, demo, a.x, a.y, vx, vy, ux, uy, kq, cx, cy, cw4) > min_paf_score))
            ++fails;
        }
        survive = fails < 2;
      }
    }
    unsigned alive = __ballot_sync(0xffffffffu, survive);
    // ---- phase 2: full score of the survivors, 3 pairs x 10 lanes per pass ----
    while (alive != 0u) {
      int src = -1;   // lane of phase 1 that holds this group's pair
      {
        unsigned m = alive;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int bpos = m ? __ffs(m) - 1 : -1;
          if (q == g) src = bpos;
          if (m) m &= m - 1;
        }
        alive = m;
      }
      const bool active = g < 3 && src >= 0;
      const int sl = src & 31;
      const int i = __shfl_sync(0xffffffffu, p_i, sl), j = __shfl_sync(0xffffffffu, p_j, sl);
      const int ax = __shfl_sync(0xffffffffu, p_ax, sl), ay = __shfl_sync(0xffffffffu, p_ay, sl);
      const int vx = __shfl_sync(0xffffffffu, p_vx, sl), vy = __shfl_sync(0xffffffffu, p_vy, sl);
      const double norm = __shfl_sync(0xffffffffu, p_norm, sl);
      const double ux = __shfl_sync(0xffffffffu, p_ux, sl), uy = __shfl_sync(0xffffffffu, p_uy, sl);
      double v = 0.0;
      bool pass = false;
      if (active) {
        v = paf_sample<kFused, kSmem>(paf, W, ld, up, s_src, img, demo, ax, ay, vx, vy, ux, uy, k, cx, cy, cw4);
        pass = v > min_paf_score;
      }
      double sum = 0.0;
      int cnt = 0;
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) {
        int srcl = (g * 10 + kk) & 31;
        double vk = __shfl_sync(0xffffffffu, v, srcl);
        int pk = __shfl_sync(0xffffffffu, (int)pass, srcl);
        if (pk) { sum = __dadd_rn(sum, vk); ++cnt; }
      }
      if (active && k == 0) {
        double ratio = cnt > 0 ? __ddiv_rn(sum, (double)cnt) : 0.0;
        double pen = __dsub_rn(__ddiv_rn(height_n, norm), 1.0);
        if (pen < 0.0) ratio = __dadd_rn(ratio, pen);
        if (ratio > 0.0 && cnt >= 9) {  // success_ratio = cnt / 10 > 0.8
          int slot = atomicAdd(out_count, 1);
          if (slot < cap_conn) { Conn c; c.ratio = ratio; c.i = i; c.j = j; out[slot] = c; }
        }
      }
    }
  }
}

// One block per (limb, image): order the candidate connections, then greedy one-to-one matching
// with warp ballots.  grid = (19, images).
__global__ void __launch_bounds__(256)
limb_match_kernel(const Conn *__restrict__ conn, const int *__restrict__ conn_count, int cap_conn,
                  const lwp_keypoint *__restrict__ kpts, const int *__restrict__ counts,
                  const int *__restrict__ kpt_start, int cap_kpts, Match *__restrict__ match,
                  int *__restrict__ match_count, int *__restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int limb = blockIdx.x, img = blockIdx.y;
  const int slot = img * LWP_NUM_LIMBS + limb;
  const int ka = c_kpt_ids[limb][0], kb = c_kpt_ids[limb][1];
  const int nA = counts[img * LWP_NUM_KPT_TYPES + ka], nB = counts[img * LWP_NUM_KPT_TYPES + kb];
  int cnt = (nA == 0 || nB == 0) ? 0 : conn_count[slot];
  if (cnt > cap_conn) {
    if (threadIdx.x == 0) overflow[img] = 1;
    cnt = cap_conn;
  }
  if (cnt == 0) {
    if (threadIdx.x == 0) match_count[slot] = 0;
    return;
  }
  int P = 32;
  while (P < cnt) P <<= 1;
  Conn *cs = reinterpret_cast<Conn *>(smem_raw);
  unsigned char *usedA = reinterpret_cast<unsigned char *>(cs + P);
  unsigned char *usedB = usedA + cap_kpts;
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    Conn c;
    if (t < cnt) c = conn[(size_t)slot * cap_conn + t];
    else { c.ratio = -DBL_MAX; c.i = INT_MAX; c.j = INT_MAX; }
    cs[t] = c;
  }
  for (int t = threadIdx.x; t < cap_kpts; t += blockDim.x) { usedA[t] = 0; usedB[t] = 0; }
  __syncthreads();
  bitonic_sort_smem(cs, P, ConnBefore());
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const lwp_keypoint *A = kpts + ((size_t)img * LWP_NUM_KPT_TYPES + ka) * cap_kpts;
  const lwp_keypoint *B = kpts + ((size_t)img * LWP_NUM_KPT_TYPES + kb) * cap_kpts;
  const int a0 = kpt_start[img * (LWP_NUM_KPT_TYPES + 1) + ka], b0 = kpt_start[img * (LWP_NUM_KPT_TYPES + 1) + kb];
  Match *out = match + (size_t)slot * cap_kpts;
  const int limit = nA < nB ? nA : nB;
  int accepted = 0;
  for (int base = 0; base < cnt && accepted < limit; base += 32) {
    int idx = base + lane;
    bool valid = idx < cnt;
    Conn c;
    c.ratio = 0; c.i = 0; c.j = 0;
    if (valid) c = cs[idx];
    bool freeij = valid && !usedA[c.i] && !usedB[c.j];
    unsigned m = __ballot_sync(0xffffffffu, freeij);
    while (m != 0 && accepted < limit) {
      int b = __ffs(m) - 1;
      m &= m - 1;
      int ok = 0;
      if (lane == b) ok = !usedA[c.i] && !usedB[c.j];  // earlier lanes of this chunk may have taken i or j
      ok = __shfl_sync(0xffffffffu, ok, b);
      if (ok) {
        if (lane == b) {
          usedA[c.i] = 1; usedB[c.j] = 1;
          cs[accepted] = c;   // accepted connections are compacted in place (accepted <= the index just read)
        }
        ++accepted;
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // the key-point scores are fetched for all accepted connections at once (a dependent global load inside the greedy
  // loop would serialise one memory round trip per connection)
  for (int t = lane; t < accepted; t += 32) {
    const Conn c = cs[t];
    Match mt;
    mt.ratio = c.ratio; mt.ida = a0 + c.i; mt.idb = b0 + c.j;
    mt.sa = A[c.i].score; mt.sb = B[c.j].score;
    out[t] = mt;
  }
  if (lane == 0) match_count[slot] = accepted;
}

// Pose assembly (modules/keypoints.py:63-92,159-200), one block per image.  The reference walks the accepted connections
// of a limb one after the other and scans every pose for each of them; limbs must stay in table order, but WITHIN a limb
// the connections are independent: greedy matching is one-to-one, so the key-point ids on either side of a limb's
// connections are pairwise distinct, a pose holds one id per slot and therefore matches at most one connection, and a
// pose appended for an unmatched connection (its slot ka holds that connection's own id) can never match a later one.
// So every limb is a constant number of block-wide steps: (1) an id -> connection lookup table, (2) all poses updated in
// parallel through it, (3) unmatched connections appended in connection order by a prefix sum -- the same pose table,
// bit for bit (every pose still receives its single update `score += (score_b + ratio)` with the same operands), in a
// few microseconds instead of ~100 for the one-warp sequential walk.  Quirks kept: limb 0 replaces the list, limbs
// 17 / 18 only fill a missing end and touch neither score nor count, singleton poses for one-sided limbs.
constexpr int kPaThreads = 128;

// exclusive prefix over the block of one flag per thread; *total = number of flags set (uniform); two barriers inside
__device__ __forceinline__ int block_excl_scan_flag(bool flag, int *s_warp /* kPaThreads / 32 + 1 ints */, int *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_warp[w] = __popc(m);
  __syncthreads();
  int base = 0, sum = 0;
#pragma unroll
  for (int i = 0; i < kPaThreads / 32; ++i) {
    const int v = s_warp[i];
    if (i < w) base += v;
    sum += v;
  }
  __syncthreads();   // s_warp may be rewritten by the next call
  *total = sum;
  return base + __popc(m & ((1u << lane) - 1u));
}

template <bool kSmem>
__global__ void __launch_bounds__(kPaThreads)
pose_assemble_kernel(const lwp_keypoint *__restrict__ kpts, const int *__restrict__ counts,
                     const int *__restrict__ kpt_start, int cap_kpts, const Match *__restrict__ match,
                     const int *__restrict__ match_count, double *__restrict__ scratch,
                     double *__restrict__ pose_entries, int *__restrict__ n_poses, int cap_poses,
                     int *__restrict__ overflow) {
  const int img = blockIdx.x, tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char pa_smem[];
  // layout: [poses (kSmem only)] [matches of the current limb: cap_kpts] [lutA, lutB, found: 3 x cap_kpts ints]
  double *poses = kSmem ? reinterpret_cast<double *>(pa_smem) : scratch + (size_t)img * cap_poses * LWP_POSE_ENTRY;
  Match *s_match = reinterpret_cast<Match *>(pa_smem + (kSmem ? (size_t)cap_poses * LWP_POSE_ENTRY * sizeof(double) : 0));
  int *lutA = reinterpret_cast<int *>(s_match + cap_kpts), *lutB = lutA + cap_kpts, *found = lutB + cap_kpts;
  __shared__ int s_cnt[LWP_NUM_KPT_TYPES], s_start[LWP_NUM_KPT_TYPES + 1], s_m[LWP_NUM_LIMBS], s_warp[kPaThreads / 32 + 1];
#define POSE(jj) (poses + (size_t)(jj) * LWP_POSE_ENTRY)
  if (tid < LWP_NUM_KPT_TYPES) s_cnt[tid] = counts[img * LWP_NUM_KPT_TYPES + tid];
  if (tid < LWP_NUM_KPT_TYPES + 1) s_start[tid] = kpt_start[img * (LWP_NUM_KPT_TYPES + 1) + tid];
  if (tid < LWP_NUM_LIMBS) s_m[tid] = match_count[img * LWP_NUM_LIMBS + tid];
  __syncthreads();
  int np = 0;          // uniform over the block
  bool ovf = false;
  const Match *Mimg = match + (size_t)img * LWP_NUM_LIMBS * cap_kpts;

  // write one pose holding up to two key-points (one thread writes all 20 fields)
  auto write_pose = [&](int pos, int slot_a, double id_a, int slot_b, double id_b, double score, double count) {
    double *p = POSE(pos);
#pragma unroll
    for (int q = 0; q < LWP_NUM_KPT_TYPES; ++q) p[q] = -1.0;
    p[slot_a] = id_a;
    if (slot_b >= 0) p[slot_b] = id_b;
    p[18] = score;
    p[19] = count;
  };

  for (int limb = 0; limb < LWP_NUM_LIMBS; ++limb) {
    const int ka = c_kpt_ids[limb][0], kb = c_kpt_ids[limb][1];
    const int nA = s_cnt[ka], nB = s_cnt[kb];
    if (nA == 0 && nB == 0) continue;
    if (nA == 0 || nB == 0) {  // :66-92 singleton poses, in key-point order, for key-points no pose holds yet
      const int slot = nA == 0 ? kb : ka, cnt = nA == 0 ? nB : nA, start = s_start[slot];
      const lwp_keypoint *K = kpts + ((size_t)img * LWP_NUM_KPT_TYPES + slot) * cap_kpts;
      for (int i = tid; i < cnt; i += kPaThreads) found[i] = 0;
      __syncthreads();
      for (int j = tid; j < np; j += kPaThreads) {
        const double v = POSE(j)[slot];
        if (v != -1.0) {
          const int idx = (int)v - start;
          if (idx >= 0 && idx < cnt) found[idx] = 1;
        }
      }
      __syncthreads();
      for (int base = 0; base < cnt; base += kPaThreads) {   // ordered append of the key-points not found
        const int i = base + tid;
        const bool add = i < cnt && !found[i];
        int total;
        const int pos = np + block_excl_scan_flag(add, s_warp, &total);
        if (add) {
          if (pos < cap_poses) write_pose(pos, slot, (double)(start + i), -1, 0.0, (double)K[i].score, 1.0);
        }
        np += total;
      }
      if (np > cap_poses) { ovf = true; np = cap_poses; }
      __syncthreads();
      continue;
    }
    const int m = s_m[limb];
    if (m == 0) continue;
    const Match *Mg = Mimg + (size_t)limb * cap_kpts;
    const int startA = s_start[ka], startB = s_start[kb];
    for (int c = tid; c < m; c += kPaThreads) s_match[c] = Mg[c];
    if (limb == 0) {  // :159-165 replaces the list with one pose per accepted connection
      __syncthreads();
      np = m < cap_poses ? m : cap_poses;
      if (m > cap_poses) ovf = true;
      for (int c = tid; c < np; c += kPaThreads) {
        const Match mt = s_match[c];
        write_pose(c, ka, (double)mt.ida, kb, (double)mt.idb, __dadd_rn(__dadd_rn((double)mt.sa, (double)mt.sb), mt.ratio), 2.0);
      }
      __syncthreads();
      continue;
    }
    // id -> connection tables (ids of one key-point type are consecutive: local index = id - start of the type)
    for (int i = tid; i < nA; i += kPaThreads) lutA[i] = -1;
    for (int i = tid; i < nB; i += kPaThreads) lutB[i] = -1;
    for (int c = tid; c < m; c += kPaThreads) found[c] = 0;
    __syncthreads();
    for (int c = tid; c < m; c += kPaThreads) {
      lutA[s_match[c].ida - startA] = c;
      lutB[s_match[c].idb - startB] = c;
    }
    __syncthreads();
    if (limb == 17 || limb == 18) {  // :166-175 fill a missing end only; no score / count change, no new pose
      for (int j = tid; j < np; j += kPaThreads) {
        double *p = POSE(j);
        const double va = p[ka], vb = p[kb];
        if (va != -1.0 && vb == -1.0) {
          const int c = lutA[(int)va - startA];
          if (c >= 0) p[kb] = (double)s_match[c].idb;
        } else if (va == -1.0 && vb != -1.0) {
          const int c = lutB[(int)vb - startB];
          if (c >= 0) p[ka] = (double)s_match[c].ida;
        }
      }
      __syncthreads();
      continue;
    }
    // :176-193 every pose whose slot ka holds a connection's a gets its b, count += 1, score += (score_b + ratio)
    for (int j = tid; j < np; j += kPaThreads) {
      double *p = POSE(j);
      const double va = p[ka];
      if (va != -1.0) {
        const int c = lutA[(int)va - startA];
        if (c >= 0) {
          const Match mt = s_match[c];
          p[kb] = (double)mt.idb;
          p[19] = __dadd_rn(p[19], 1.0);
          p[18] = __dadd_rn(p[18], __dadd_rn((double)mt.sb, mt.ratio));
          found[c] = 1;
        }
      }
    }
    __syncthreads();
    for (int base = 0; base < m; base += kPaThreads) {   // connections no pose matched: new poses, in connection order
      const int c = base + tid;
      const bool add = c < m && !found[c];
      int total;
      const int pos = np + block_excl_scan_flag(add, s_warp, &total);
      if (add && pos < cap_poses) {
        const Match mt = s_match[c];
        write_pose(pos, ka, (double)mt.ida, kb, (double)mt.idb, __dadd_rn(__dadd_rn((double)mt.sa, (double)mt.sb), mt.ratio), 2.0);
      }
      np += total;
    }
    if (np > cap_poses) { ovf = true; np = cap_poses; }
    __syncthreads();
  }
  __syncthreads();
  // :195-200 final filter, order preserved
  double *outp = pose_entries + (size_t)img * cap_poses * LWP_POSE_ENTRY;
  int kept = 0;
  for (int base = 0; base < np; base += kPaThreads) {
    const int j = base + tid;
    bool keep = false;
    if (j < np) {
      const double cnt = POSE(j)[19], sc = POSE(j)[18];
      keep = !(cnt < 3.0 || __ddiv_rn(sc, cnt) < 0.2);
    }
    int total;
    const int pos = kept + block_excl_scan_flag(keep, s_warp, &total);
    if (keep)
      for (int q = 0; q < LWP_POSE_ENTRY; ++q) outp[(size_t)pos * LWP_POSE_ENTRY + q] = POSE(j)[q];
    kept += total;
  }
  if (tid == 0) {
    n_poses[img] = kept;
    if (ovf) overflow[img] = 1;
  }
#undef POSE
}

static int grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace lwp

using namespace lwp;

extern "C" int lwp_upsample_cubic_ex(const float *src, int n, int h, int w, int c, int src_ld, long long src_row_pitch,
                                     long long src_img_pitch, float *dst, int H, int W, double inv_scale_x,
                                     double inv_scale_y, float accumulate_divisor, void *stream) {
  LWP_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c > 0 && src_ld >= c && H > 0 && W > 0,
              "lwp_upsample_cubic: bad shape");
  LWP_REQUIRE(inv_scale_x > 0 && inv_scale_y > 0, "lwp_upsample_cubic: bad scale");
  LWP_REQUIRE(src_row_pitch >= (long long)w * src_ld && src_img_pitch >= (long long)h * src_row_pitch, "lwp_upsample_cubic: bad pitch");
  LWP_REQUIRE((long long)W * c < INT_MAX, "lwp_upsample_cubic: row too long");
  UpSrc u;
  u.src = src; u.h = h; u.w = w; u.ld = src_ld; u.c_layout = c; u.H = H; u.W = W;
  u.scale_x = 1. / inv_scale_x; u.scale_y = 1. / inv_scale_y;
  // power-of-two ratio in both directions, plain store: the periodic-geometry kernel
  const int ratio = (H == 2 * h && W == 2 * w) ? 2 : (H == 4 * h && W == 4 * w) ? 4 : (H == 8 * h && W == 8 * w) ? 8 : 0;
  if (ratio != 0 && inv_scale_x == (double)ratio && inv_scale_y == (double)ratio && accumulate_divisor == 0.f &&
      getenv("LWP_UPSAMPLE_GENERIC") == nullptr) {
    const long long tot = (long long)n * (h + 1) * (w + 1) * c;
    const long long blk = (tot + 255) / 256;
    LWP_REQUIRE(blk < INT_MAX, "lwp_upsample_cubic: too many blocks");
    if (ratio == 2) upsample_cubic_pow2_kernel<2><<<(unsigned)blk, 256, 0, (cudaStream_t)stream>>>(u, dst, src_row_pitch, src_img_pitch, tot);
    else if (ratio == 4) upsample_cubic_pow2_kernel<4><<<(unsigned)blk, 256, 0, (cudaStream_t)stream>>>(u, dst, src_row_pitch, src_img_pitch, tot);
    else upsample_cubic_pow2_kernel<8><<<(unsigned)blk, 256, 0, (cudaStream_t)stream>>>(u, dst, src_row_pitch, src_img_pitch, tot);
    LWP_LAUNCH_CHECK();
    return LWP_OK;
  }
  const int col_tiles = ceil_div(W, kUpCols);
  const long long blocks = (long long)n * H * col_tiles;
  LWP_REQUIRE(blocks < INT_MAX, "lwp_upsample_cubic: too many blocks");
  // the row-streaming kernel wins where several output rows share a source row (measured, 16 x 57 channels -> 480x640:
  // from 184x248 1.81 -> 1.33 ms; from 368x491 1.81 -> 1.78; from 552x736 / 736x984 1.86 / 1.89 -> 2.23 / 2.67 ms: with
  // fewer output rows per source row its horizontal passes and two barriers per row cost more than 16 L1-resident loads)
  if (c <= kUpMaxC && u.scale_y <= 0.5 && getenv("LWP_UPSAMPLE_NO_ROWS") == nullptr) {
    const int row_tiles = ceil_div(H, kUpRows);
    const long long blocks_r = (long long)n * row_tiles * col_tiles;
    LWP_REQUIRE(blocks_r < INT_MAX, "lwp_upsample_cubic: too many blocks");
    upsample_cubic_rows_kernel<<<(unsigned)blocks_r, 256, 0, (cudaStream_t)stream>>>(u, dst, col_tiles, row_tiles, src_row_pitch,
                                                                                   src_img_pitch, accumulate_divisor);
    LWP_LAUNCH_CHECK();
    return LWP_OK;
  }
  upsample_cubic_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(u, dst, col_tiles, src_row_pitch, src_img_pitch,
                                                                            accumulate_divisor);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

extern "C" int lwp_upsample_cubic(const float *src, int n, int h, int w, int c, int src_ld, float *dst, int H, int W,
                                  double inv_scale_x, double inv_scale_y, void *stream) {
  return lwp_upsample_cubic_ex(src, n, h, w, c, src_ld, (long long)w * src_ld, (long long)h * w * src_ld, dst, H, W, inv_scale_x,
                               inv_scale_y, 0.f, stream);
}

extern "C" size_t lwp_extract_workspace_bytes(int n, int n_ch, int cap_candidates) {
  size_t slots = (size_t)n * n_ch;
  return align_up(slots * sizeof(int), 256) + slots * (size_t)cap_candidates * sizeof(unsigned long long);
}

static int extract_common(bool fused, const float *hm, const UpSrc *up, int n, int H, int W, int ld, int n_ch,
                          lwp_keypoint *kpts, int32_t *counts, int32_t *kpt_start, int cap_kpts, int cap_candidates,
                          void *workspace, size_t workspace_bytes, int32_t *overflow, void *stream) {
  LWP_REQUIRE(hm && kpts && counts && kpt_start && workspace && overflow, "lwp_extract_keypoints: null pointer");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && H < 65536 && W < 65536 && n_ch > 0 && n_ch <= 64 && ld >= n_ch,
              "lwp_extract_keypoints: bad shape");
  LWP_REQUIRE(cap_kpts > 0 && cap_candidates >= cap_kpts, "lwp_extract_keypoints: bad capacities");
  LWP_REQUIRE(workspace_bytes >= lwp_extract_workspace_bytes(n, n_ch, cap_candidates),
              "lwp_extract_keypoints: workspace too small");
  size_t smem = (size_t)next_pow2(cap_candidates) * 9;
  if (smem > 200 * 1024) { set_error("lwp_extract_keypoints: cap_candidates %d too large", cap_candidates); return LWP_ECAP; }
  cudaStream_t st = (cudaStream_t)stream;
  size_t slots = (size_t)n * n_ch;
  int *cand_count = (int *)workspace;
  unsigned long long *cand = (unsigned long long *)((char *)workspace + align_up(slots * sizeof(int), 256));
  LWP_CUDA_CHECK(cudaMemsetAsync(cand_count, 0, slots * sizeof(int), st));
  LWP_CUDA_CHECK(cudaMemsetAsync(overflow, 0, (size_t)n * sizeof(int), st));
  static DeviceOnce attr_set;
  int attr_set_slot;
  if (attr_set.pending(&attr_set_slot)) {
    LWP_CUDA_CHECK(cudaFuncSetAttribute(peak_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LWP_CUDA_CHECK(cudaFuncSetAttribute(peak_candidates_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        100 * 1024));
    attr_set.done[attr_set_slot] = true;
  }
  if (fused) {
    LWP_REQUIRE(n_ch <= kPkMaxCh, "lwp_extract_keypoints_fused: at most %d channels", kPkMaxCh);
    const size_t pk_smem = (size_t)(kPkSrcMax * n_ch * kPkSrcMax + kPkWarps * kPkSrcMax * 32) * sizeof(float);
    dim3 grid(ceil_div(W, kPkCols) * ceil_div(H, kPkRows), n);
    peak_candidates_fused_kernel<<<grid, kPkWarps * 32, pk_smem, st>>>(*up, n_ch, cand, cand_count, cap_candidates,
                                                                      overflow);
  } else {
    long long total = (long long)n * H * W * n_ch;
    peak_candidates_kernel<<<grid_for(total, 256), 256, 0, st>>>(hm, H, W, ld, n_ch, cand, cand_count, cap_candidates,
                                                                 total);
  }
  LWP_LAUNCH_CHECK();
  peak_nms_kernel<<<(unsigned)slots, 256, smem, st>>>(cand, cand_count, cap_candidates, n_ch, kpts, counts, cap_kpts,
                                                      overflow);
  LWP_LAUNCH_CHECK();
  keypoint_ids_kernel<<<n, 128, 0, st>>>(kpts, counts, kpt_start, n_ch, cap_kpts);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

extern "C" int lwp_extract_keypoints(const float *hm, int n, int H, int W, int ld, int n_ch, lwp_keypoint *kpts,
                                     int32_t *counts, int32_t *kpt_start, int cap_kpts, int cap_candidates,
                                     void *workspace, size_t workspace_bytes, int32_t *overflow, void *stream) {
  return extract_common(false, hm, nullptr, n, H, W, ld, n_ch, kpts, counts, kpt_start, cap_kpts, cap_candidates,
                        workspace, workspace_bytes, overflow, stream);
}

static int make_upsrc(const float *src, int h, int w, int ld, int c_layout, int H, int W, double inv_scale_x,
                      double inv_scale_y, UpSrc *u) {
  LWP_REQUIRE(src && h > 0 && w > 0 && ld >= c_layout && c_layout > 0 && H > 0 && W > 0, "fused source: bad shape");
  LWP_REQUIRE(inv_scale_x >= 3.0 && inv_scale_y >= 3.0,
              "fused post-processing needs an up-sampling factor >= 3 (got %g, %g); use the materialising path",
              inv_scale_x, inv_scale_y);
  u->src = src; u->h = h; u->w = w; u->ld = ld; u->c_layout = c_layout; u->H = H; u->W = W;
  u->scale_x = 1. / inv_scale_x; u->scale_y = 1. / inv_scale_y;
  return LWP_OK;
}

extern "C" int lwp_extract_keypoints_fused(const float *src, int n, int h, int w, int ld, int n_ch, int c_layout, int H,
                                           int W, double inv_scale_x, double inv_scale_y, lwp_keypoint *kpts,
                                           int32_t *counts, int32_t *kpt_start, int cap_kpts, int cap_candidates,
                                           void *workspace, size_t workspace_bytes, int32_t *overflow, void *stream) {
  UpSrc u;
  int rc = make_upsrc(src, h, w, ld, c_layout, H, W, inv_scale_x, inv_scale_y, &u);
  if (rc != LWP_OK) return rc;
  LWP_REQUIRE(n_ch <= c_layout, "lwp_extract_keypoints_fused: n_ch > c_layout");
  return extract_common(true, src, &u, n, H, W, ld, n_ch, kpts, counts, kpt_start, cap_kpts, cap_candidates, workspace,
                        workspace_bytes, overflow, stream);
}

namespace {
struct GroupWs {
  int *conn_count;
  int *match_count;
  Conn *conn;
  Match *match;
  double *poses;
  size_t bytes;
};
GroupWs carve_group_ws(void *base, int n, int cap_kpts, int cap_conn, int cap_poses) {
  GroupWs w;
  size_t off = 0;
  size_t slots = (size_t)n * LWP_NUM_LIMBS;
  char *b = (char *)base;
  w.conn_count = (int *)(b + off); off += align_up(slots * sizeof(int), 256);
  w.match_count = (int *)(b + off); off += align_up(slots * sizeof(int), 256);
  w.conn = (Conn *)(b + off); off += align_up(slots * (size_t)cap_conn * sizeof(Conn), 256);
  w.match = (Match *)(b + off); off += align_up(slots * (size_t)cap_kpts * sizeof(Match), 256);
  w.poses = (double *)(b + off); off += align_up((size_t)n * cap_poses * LWP_POSE_ENTRY * sizeof(double), 256);
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" size_t lwp_group_workspace_bytes(int n, int cap_kpts, int cap_connections, int cap_poses) {
  return carve_group_ws(nullptr, n, cap_kpts, cap_connections, cap_poses).bytes;
}

extern "C" size_t lwp_paf_pack_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  const size_t plane_stride = ((size_t)h * w + 1) & ~(size_t)1;
  return 256 + align_up((size_t)n * LWP_NUM_LIMBS * plane_stride * sizeof(float2), 256);   // 256: alignment of the part before it
}

static int group_common(bool fused, const lwp_keypoint *kpts, const int32_t *counts, const int32_t *kpt_start,
                        int cap_kpts, const float *pafs, const UpSrc *up, int n, int H, int W, int paf_ld, int demo,
                        double min_paf_score, double *pose_entries, int32_t *n_poses, int cap_poses,
                        int cap_connections, void *workspace, size_t workspace_bytes, int32_t *overflow, void *stream) {
  LWP_REQUIRE(kpts && counts && kpt_start && pafs && pose_entries && n_poses && workspace && overflow,
              "lwp_group_keypoints: null pointer");
  LWP_REQUIRE(n > 0 && H > 0 && W > 0 && paf_ld >= 38 && cap_kpts > 0 && cap_poses > 0 && cap_connections > 0,
              "lwp_group_keypoints: bad shape");
  LWP_REQUIRE(workspace_bytes >= lwp_group_workspace_bytes(n, cap_kpts, cap_connections, cap_poses),
              "lwp_group_keypoints: workspace too small");
  size_t smem = (size_t)next_pow2(cap_connections) * sizeof(Conn) + 2 * (size_t)cap_kpts;
  if (smem > 200 * 1024) { set_error("lwp_group_keypoints: cap_connections %d too large", cap_connections); return LWP_ECAP; }
  cudaStream_t st = (cudaStream_t)stream;
  GroupWs w = carve_group_ws(workspace, n, cap_kpts, cap_connections, cap_poses);
  LWP_CUDA_CHECK(cudaMemsetAsync(w.conn_count, 0, (size_t)n * LWP_NUM_LIMBS * sizeof(int), st));
  int bx = 2368 / (LWP_NUM_LIMBS * n);  // frames differ a lot in pair count: spread each limb over several blocks
  bx = bx < 8 ? 8 : (bx > 32 ? 32 : bx);
  UpSrc u0;
  memset(&u0, 0, sizeof(u0));
  if (fused) {
    const size_t src_bytes = (size_t)up->h * up->w * sizeof(float2);   // the limb's two PAF channels of one image
    static DeviceOnce ps_attr;
    int ps_attr_slot;
    if (ps_attr.pending(&ps_attr_slot)) {
      LWP_CUDA_CHECK(cudaFuncSetAttribute(paf_score_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      ps_attr.done[ps_attr_slot] = true;
    }
    // staged variant: every block copies the limb's two channels first, so use as few blocks per (limb, image) as still
    // fill the GPU (batches: one block of 8 warps walks all pairs; a single image: up to 8 blocks per limb)
    int bs = ceil_div(4 * num_sms(), LWP_NUM_LIMBS * n);
    bs = bs < 1 ? 1 : (bs > 8 ? 8 : bs);
    // a workspace with lwp_paf_pack_bytes(n, h, w) extra bytes behind the lwp_group_workspace_bytes() part enables the
    // packed staging (paf_pack_kernel): same results, the limb's channels then arrive as one coalesced copy, which also
    // makes it affordable to spread a (limb, image) with many candidate pairs over several blocks (blocks without pairs
    // return before staging)
    const int plane_stride = (up->h * up->w + 1) & ~1;
    const size_t base_bytes = align_up(w.bytes, 256);
    float2 *packed = nullptr;
    if (src_bytes <= 200 * 1024 && workspace_bytes >= base_bytes + lwp_paf_pack_bytes(n, up->h, up->w) &&
        getenv("LWP_NO_PAF_PACK") == nullptr) {
      packed = reinterpret_cast<float2 *>((char *)workspace + base_bytes);
      paf_pack_kernel<<<dim3(ceil_div(up->h * up->w, kPackPx), n), 256, 0, st>>>(*up, packed, plane_stride);
      LWP_LAUNCH_CHECK();
      bs = ceil_div(24 * num_sms(), LWP_NUM_LIMBS * n);   // 64 frames: 3 (measured 1 / 2 / 3 / 4 / 8 blocks: 205 / 188 / 183 / 186 / 187 us for the grouping)
      bs = bs < 1 ? 1 : (bs > 8 ? 8 : bs);
      if (const char *e = getenv("LWP_PAF_BLOCKS")) { int v = atoi(e); if (v >= 1 && v <= 32) bs = v; }
    }
    if (src_bytes <= 200 * 1024)
      paf_score_kernel<true, true><<<dim3(bs, LWP_NUM_LIMBS, n), 256, (size_t)plane_stride * sizeof(float2), st>>>(
          kpts, counts, cap_kpts, nullptr, H, W, paf_ld, *up, demo, min_paf_score, w.conn, w.conn_count, cap_connections,
          packed, plane_stride);
    else
      paf_score_kernel<true, false><<<dim3(bx, LWP_NUM_LIMBS, n), 128, 0, st>>>(
          kpts, counts, cap_kpts, nullptr, H, W, paf_ld, *up, demo, min_paf_score, w.conn, w.conn_count, cap_connections,
          nullptr, 0);
  } else {
    paf_score_kernel<false, false><<<dim3(bx, LWP_NUM_LIMBS, n), 128, 0, st>>>(kpts, counts, cap_kpts, pafs, H, W, paf_ld, u0,
                                                                               demo, min_paf_score, w.conn, w.conn_count,
                                                                               cap_connections, nullptr, 0);
  }
  LWP_LAUNCH_CHECK();
  static DeviceOnce attr_set;
  int attr_set_slot;
  if (attr_set.pending(&attr_set_slot)) {
    LWP_CUDA_CHECK(cudaFuncSetAttribute(limb_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.done[attr_set_slot] = true;
  }
  limb_match_kernel<<<dim3(LWP_NUM_LIMBS, n), 256, smem, st>>>(w.conn, w.conn_count, cap_connections, kpts, counts,
                                                              kpt_start, cap_kpts, w.match, w.match_count, overflow);
  LWP_LAUNCH_CHECK();
  const size_t pa_tables = (size_t)cap_kpts * (sizeof(Match) + 3 * sizeof(int));   // current limb's matches + lutA / lutB / found
  const size_t pa_smem = (size_t)cap_poses * LWP_POSE_ENTRY * sizeof(double) + pa_tables;
  if (pa_smem <= 160 * 1024) {
    static DeviceOnce pa_attr;
    int pa_attr_slot;
    if (pa_attr.pending(&pa_attr_slot)) {
      LWP_CUDA_CHECK(cudaFuncSetAttribute(pose_assemble_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          160 * 1024));
      pa_attr.done[pa_attr_slot] = true;
    }
    pose_assemble_kernel<true><<<n, kPaThreads, pa_smem, st>>>(kpts, counts, kpt_start, cap_kpts, w.match, w.match_count,
                                                       w.poses, pose_entries, n_poses, cap_poses, overflow);
  } else {
    if (pa_tables > 200 * 1024) { set_error("lwp_group_keypoints: cap_kpts %d too large", cap_kpts); return LWP_ECAP; }
    static DeviceOnce pb_attr;
    int pb_attr_slot;
    if (pb_attr.pending(&pb_attr_slot)) {
      LWP_CUDA_CHECK(cudaFuncSetAttribute(pose_assemble_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      pb_attr.done[pb_attr_slot] = true;
    }
    pose_assemble_kernel<false><<<n, kPaThreads, pa_tables, st>>>(
        kpts, counts, kpt_start, cap_kpts, w.match, w.match_count, w.poses, pose_entries, n_poses, cap_poses, overflow);
  }
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

extern "C" int lwp_group_keypoints(const lwp_keypoint *kpts, const int32_t *counts, const int32_t *kpt_start,
                                   int cap_kpts, const float *pafs, int n, int H, int W, int paf_ld, int demo,
                                   double min_paf_score, double *pose_entries, int32_t *n_poses, int cap_poses,
                                   int cap_connections, void *workspace, size_t workspace_bytes, int32_t *overflow,
                                   void *stream) {
  return group_common(false, kpts, counts, kpt_start, cap_kpts, pafs, nullptr, n, H, W, paf_ld, demo, min_paf_score,
                      pose_entries, n_poses, cap_poses, cap_connections, workspace, workspace_bytes, overflow, stream);
}

extern "C" int lwp_group_keypoints_fused(const lwp_keypoint *kpts, const int32_t *counts, const int32_t *kpt_start,
                                         int cap_kpts, const float *src, int n, int h, int w, int ld, int H, int W,
                                         double inv_scale_x, double inv_scale_y, int demo, double min_paf_score,
                                         double *pose_entries, int32_t *n_poses, int cap_poses, int cap_connections,
                                         void *workspace, size_t workspace_bytes, int32_t *overflow, void *stream) {
  UpSrc u;
  int rc = make_upsrc(src, h, w, ld, 38, H, W, inv_scale_x, inv_scale_y, &u);
  if (rc != LWP_OK) return rc;
  return group_common(true, kpts, counts, kpt_start, cap_kpts, src, &u, n, H, W, ld, demo, min_paf_score, pose_entries,
                      n_poses, cap_poses, cap_connections, workspace, workspace_bytes, overflow, stream);
}

// ------------------------------------------------------------------------------------------------
// Result post-conversion: what demo.py:101-115 does with the pose table -- key-point coordinates back in the original
// frame, (x * stride / upsample_ratio - pad) / scale in float64 with every operation rounded separately, Python int()
// truncation -- plus Pose.get_bbox (modules/pose.py:30-39: cv2.boundingRect of the found key-points).  One thread per pose.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pose_convert_kernel(const double *__restrict__ pose_entries, const int *__restrict__ n_poses, int cap_poses,
                    const lwp_keypoint *__restrict__ kpts, const int *__restrict__ kpt_start, int cap_kpts,
                    double stride, double upsample_ratio, const double *__restrict__ xform, int32_t *__restrict__ pose_kpts,
                    int32_t *__restrict__ bbox, double *__restrict__ confidence) {
  const int img = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_poses[img] || j >= cap_poses) return;
  const double pad_left = xform[img * 3 + 0], pad_top = xform[img * 3 + 1], scale = xform[img * 3 + 2];
  const double *pe = pose_entries + ((size_t)img * cap_poses + j) * LWP_POSE_ENTRY;
  int32_t *pk = pose_kpts + ((size_t)img * cap_poses + j) * LWP_NUM_KPT_TYPES * 2;
  const int *starts = kpt_start + img * (LWP_NUM_KPT_TYPES + 1);
  int minx = INT_MAX, miny = INT_MAX, maxx = INT_MIN, maxy = INT_MIN;
  for (int k = 0; k < LWP_NUM_KPT_TYPES; ++k) {
    int x = -1, y = -1;
    const double v = pe[k];
    if (v != -1.0) {
      const int id = (int)v;                     // global id -> (channel k, index id - start of the channel)
      const lwp_keypoint kp = kpts[((size_t)img * LWP_NUM_KPT_TYPES + k) * cap_kpts + (id - starts[k])];
      x = __double2int_rz(__ddiv_rn(__dsub_rn(__ddiv_rn(__dmul_rn((double)kp.x, stride), upsample_ratio), pad_left), scale));
      y = __double2int_rz(__ddiv_rn(__dsub_rn(__ddiv_rn(__dmul_rn((double)kp.y, stride), upsample_ratio), pad_top), scale));
      minx = min(minx, x); maxx = max(maxx, x); miny = min(miny, y); maxy = max(maxy, y);
    }
    pk[2 * k] = x; pk[2 * k + 1] = y;
  }
  int32_t *bb = bbox + ((size_t)img * cap_poses + j) * 4;
  if (maxx >= minx) { bb[0] = minx; bb[1] = miny; bb[2] = maxx - minx + 1; bb[3] = maxy - miny + 1; }
  else { bb[0] = bb[1] = bb[2] = bb[3] = 0; }
  confidence[(size_t)img * cap_poses + j] = pe[18];
}

extern "C" int lwp_pose_convert(const double *pose_entries, const int32_t *n_poses, int cap_poses, const lwp_keypoint *kpts,
                                const int32_t *kpt_start, int cap_kpts, int n, double stride, double upsample_ratio,
                                const double *xform, int32_t *pose_kpts, int32_t *bbox, double *confidence, void *stream) {
  LWP_REQUIRE(pose_entries && n_poses && kpts && kpt_start && xform && pose_kpts && bbox && confidence,
              "lwp_pose_convert: null pointer");
  LWP_REQUIRE(n > 0 && cap_poses > 0 && cap_kpts > 0 && upsample_ratio != 0.0, "lwp_pose_convert: bad arguments");
  dim3 grid(ceil_div(cap_poses, 128), n);
  pose_convert_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(pose_entries, n_poses, cap_poses, kpts, kpt_start, cap_kpts,
                                                              stride, upsample_ratio, xform, pose_kpts, bbox, confidence);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

// Copy item i of `src` to item i of `dst` only where flags[i] != 0 (16-byte granularity).  The pipeline keeps the heads of
// frames whose fixed-capacity tables overflowed, so that exactly those frames can be re-processed with larger tables.
__global__ void __launch_bounds__(256)
copy_flagged_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, const int *__restrict__ flags, size_t vec_per_item) {
  const int item = blockIdx.y;
  if (flags[item] == 0) return;
  const uint4 *s = src + (size_t)item * vec_per_item;
  uint4 *d = dst + (size_t)item * vec_per_item;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_item; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

extern "C" int lwp_copy_flagged(const void *src, void *dst, const int32_t *flags, int n, size_t bytes_per_item, void *stream) {
  LWP_REQUIRE(src && dst && flags && n > 0, "lwp_copy_flagged: bad arguments");
  LWP_REQUIRE(bytes_per_item % 16 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0, "lwp_copy_flagged: 16-byte alignment");
  dim3 grid(16, n);
  copy_flagged_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4 *)src, (uint4 *)dst, flags, bytes_per_item / 16);
  LWP_LAUNCH_CHECK();
  return LWP_OK;
}

// ------------------------------------------------------------------------------------------------
// One call for the whole post-processing of a batch (demo.py:72-100 / val.py:129-134 without the maps ever being
// materialised): lwp_extract_keypoints_fused followed by lwp_group_keypoints_fused on one workspace.
// ------------------------------------------------------------------------------------------------
extern "C" size_t lwp_postprocess_workspace_bytes(int n, int cap_kpts, int cap_candidates, int cap_connections, int cap_poses) {
  return align_up(lwp_extract_workspace_bytes(n, LWP_NUM_KPT_TYPES, cap_candidates), 256) +
         lwp_group_workspace_bytes(n, cap_kpts, cap_connections, cap_poses);
}

extern "C" int lwp_postprocess(const float *heads, int n, int h, int w, int ld, int upsample_ratio, int demo, double min_paf_score,
                               lwp_keypoint *kpts, int32_t *counts, int32_t *kpt_start, int cap_kpts, int cap_candidates,
                               double *pose_entries, int32_t *n_poses, int cap_poses, int cap_connections, void *workspace,
                               size_t workspace_bytes, int32_t *overflow, void *stream) {
  LWP_REQUIRE(heads && workspace && ld >= 57 && upsample_ratio >= 3, "lwp_postprocess: bad arguments (needs the 19 + 38 head channels and a ratio >= 3)");
  LWP_REQUIRE(workspace_bytes >= lwp_postprocess_workspace_bytes(n, cap_kpts, cap_candidates, cap_connections, cap_poses),
              "lwp_postprocess: workspace too small");
  const size_t ews = align_up(lwp_extract_workspace_bytes(n, LWP_NUM_KPT_TYPES, cap_candidates), 256);
  const int H = h * upsample_ratio, W = w * upsample_ratio;
  int rc = lwp_extract_keypoints_fused(heads, n, h, w, ld, LWP_NUM_KPT_TYPES, 19, H, W, (double)upsample_ratio, (double)upsample_ratio,
                                       kpts, counts, kpt_start, cap_kpts, cap_candidates, workspace, ews, overflow, stream);
  if (rc != LWP_OK) return rc;
  return lwp_group_keypoints_fused(kpts, counts, kpt_start, cap_kpts, heads + 19, n, h, w, ld, H, W, (double)upsample_ratio,
                                   (double)upsample_ratio, demo, min_paf_score, pose_entries, n_poses, cap_poses, cap_connections,
                                   (char *)workspace + ews, workspace_bytes - ews, overflow, stream);
}
