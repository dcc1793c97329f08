/*
 * lwpose_b200.h -- C ABI of the B200-native (sm_100a) Lightweight OpenPose inference hot path.
 *
 * The reference (vivek87799/lightweight-human-pose-estimation.pytorch) is pure Python and defines no
 * FFI; every entry point below names the reference call it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates device
 *     memory after lwp_plan_create / lwp_plan_add_* returned;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never synchronise;
 *   - return value 0 = ok, otherwise an LWP_E* code; lwp_last_error() gives a thread-local message;
 *   - activations are NHWC ("pixels x channels") inside the library; the public tensors at the module
 *     boundary keep the reference's NCHW float32 layout.
 */
#ifndef LWPOSE_B200_H
#define LWPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LWP_OK 0
#define LWP_EINVAL 1   /* bad argument */
#define LWP_ECUDA 2    /* CUDA runtime / driver error (see lwp_last_error) */
#define LWP_ECAP 3     /* a capacity given by the caller is too small for the kernels' shared memory etc. */
#define LWP_EARCH 4    /* device is not sm_100 */

#define LWP_NUM_KPT_TYPES 18   /* Pose.num_kpts, modules/pose.py:9 */
#define LWP_NUM_LIMBS 19       /* len(BODY_PARTS_PAF_IDS), modules/keypoints.py:7-8 */
#define LWP_POSE_ENTRY 20      /* pose_entry_size, modules/keypoints.py:51 */

/* element type of activations/weights inside a plan */
#define LWP_DTYPE_BF16 0   /* bf16 storage, tcgen05 kind::f16, fp32 accumulate */
#define LWP_DTYPE_TF32 1   /* fp32 storage, tcgen05 kind::tf32, fp32 accumulate ("fp32 mode") */

/* epilogue activation */
#define LWP_ACT_NONE 0
#define LWP_ACT_RELU 1     /* nn.ReLU, modules/conv.py:9,17,21 */
#define LWP_ACT_ELU 2      /* nn.ELU(alpha=1), modules/conv.py:28,31 */

int lwp_version(void);
const char *lwp_last_error(void);
/* 1 if the library was built with -DLWP_TIMING_EXPERIMENTS (kernels honour the LWP_DEBUG_* work-skipping switches and
 * results may be wrong by design); 0 for a release build, where those code paths do not exist. */
int lwp_timing_experiments(void);
/* 0 if device `dev` can run this library (compute capability 10.x). */
int lwp_check_device(int dev);

/* ---------------------------------------------------------------------------------------------
 * Frame preparation
 * ------------------------------------------------------------------------------------------- */

/*
 * Replaces cv2.resize(img, (0, 0), fx=scale, fy=scale, INTER_CUBIC) on the uint8 BGR camera frame (demo.py:59) and the
 * centred pad of val.pad_width (demo.py:61-62, val.py:36-49) for n frames at once.  Bit-exact with OpenCV's generic
 * uint8 cubic path (fixed-point coefficients; builds that route this call through IPP differ by +-1 in ~5 % of pixels).
 * src: [n][h][w][3] uint8; dst: [n][Hp][Wp][3] uint8: the resized H x W frame at rows [top, top + H), columns
 * [left, left + W), the pad value everywhere else (use the mean, 128: the stem's normalisation turns it into the 0 the
 * reference pads the normalised image with).  inv_scale_* is OpenCV's inv_scale: fx (fy) when the caller gives factors
 * (then W = round(w * fx)), (double)W / w when the caller gives dsize.  dst feeds lwp_plan_add_stem_u8 plans directly.
 */
int lwp_resize_pad_u8(const uint8_t *src, int n, int h, int w, uint8_t *dst, int Hp, int Wp, int H, int W, int top, int left,
                      double inv_scale_x, double inv_scale_y, int pad_b, int pad_g, int pad_r, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Post-processing
 * ------------------------------------------------------------------------------------------- */

/* One detected key-point: the tuple (x, y, score, id) that extract_keypoints appends
 * (modules/keypoints.py:43-45). */
typedef struct lwp_keypoint {
  int32_t x;
  int32_t y;
  float score;
  int32_t id; /* global id, consecutive over the 18 channels of one image (demo.py:95-98) */
} lwp_keypoint;

/*
 * Replaces cv2.resize(maps, (0,0), fx=r, fy=r, INTER_CUBIC) at demo.py:72,76 / val.py:98,105 and
 * cv2.resize(maps, (W,H), INTER_CUBIC) at val.py:100,107 -- bit-exact with OpenCV's float path.
 * src: [n][h][w] pixels, `c` used channels per pixel, pixel stride src_ld floats.
 * dst: [n][H][W][c] contiguous.  inv_scale_* is OpenCV's inv_scale: fx (fy) when the caller gives
 * factors, (double)W / w ((double)H / h) when the caller gives dsize.
 */
int lwp_upsample_cubic(const float *src, int n, int h, int w, int c, int src_ld, float *dst, int H, int W,
                       double inv_scale_x, double inv_scale_y, void *stream);
/*
 * The same resize with (a) a source that is a cropped VIEW of a larger map -- rows src_row_pitch floats apart, images
 * src_img_pitch floats apart (val.py:99,106 crop the x8 maps before resizing them to the image size) -- and (b) an optional
 * running-average destination: accumulate_divisor != 0 makes it dst += resized / accumulate_divisor in float32, one
 * rounding per operation (val.py:101,108: avg_heatmaps = avg_heatmaps + heatmaps / len(scales)); 0 = plain store.
 */
int lwp_upsample_cubic_ex(const float *src, int n, int h, int w, int c, int src_ld, long long src_row_pitch,
                          long long src_img_pitch, float *dst, int H, int W, double inv_scale_x, double inv_scale_y,
                          float accumulate_divisor, void *stream);

/* bytes of scratch lwp_extract_keypoints needs */
size_t lwp_extract_workspace_bytes(int n, int n_ch, int cap_candidates);

/*
 * Replaces the loop `for k in range(18): extract_keypoints(heatmaps[:, :, k], ...)`
 * (demo.py:95-98, val.py:129-132; modules/keypoints.py:16-48) for n images at once.
 * hm: [n][H][W] pixels, channels 0..n_ch-1 scanned, pixel stride ld floats (read only: the
 *     reference's in-place threshold is re-done by the Python wrapper on its own host array).
 * kpts: [n][n_ch][cap_kpts]; counts: [n][n_ch]; kpt_start: [n][n_ch+1] exclusive prefix of counts
 * (global id of key-point j of channel c = kpt_start[c] + j).
 * overflow: [n] ints, set non-zero when an image exceeded cap_candidates (peak candidates per channel
 * before suppression) or cap_kpts (key-points per channel) -- results for that image are then invalid.
 */
int lwp_extract_keypoints(const float *hm, int n, int H, int W, int ld, int n_ch, lwp_keypoint *kpts,
                          int32_t *counts, int32_t *kpt_start, int cap_kpts, int cap_candidates,
                          void *workspace, size_t workspace_bytes, int32_t *overflow, void *stream);

/*
 * Same as lwp_extract_keypoints, but the up-sampled heat-maps are never materialised: peaks are found
 * on tiles rebuilt in shared memory straight from the stride-8 network output, with the bits
 * lwp_upsample_cubic would have produced (demo.py:72 + :95-98 in one pass).
 * src: [n][h][w] pixels, pixel stride ld floats, heat-map channels start at the pointer; c_layout = channel
 * count of the cv2.resize call being reproduced (19); H, W = up-sampled size; inv_scale_* as for
 * lwp_upsample_cubic (must be >= 3).
 */
int lwp_extract_keypoints_fused(const float *src, int n, int h, int w, int ld, int n_ch, int c_layout, int H, int W,
                                double inv_scale_x, double inv_scale_y, lwp_keypoint *kpts, int32_t *counts,
                                int32_t *kpt_start, int cap_kpts, int cap_candidates, void *workspace,
                                size_t workspace_bytes, int32_t *overflow, void *stream);

size_t lwp_group_workspace_bytes(int n, int cap_kpts, int cap_connections, int cap_poses);
/*
 * Optional extra workspace of lwp_group_keypoints_fused / lwp_postprocess for n images with h x w stride-8 maps:
 * when the workspace handed in is at least lwp_group_workspace_bytes(...) (resp. lwp_postprocess_workspace_bytes(...))
 * + lwp_paf_pack_bytes(n, h, w) bytes, the PAF channel pairs are first re-packed per limb ([n][19][h*w] float2) so
 * that the line-integral kernel (modules/keypoints.py:94-139) stages a limb's two channels as one contiguous copy and
 * spreads crowded frames over several blocks.  Results are identical with and without it.
 */
size_t lwp_paf_pack_bytes(int n, int h, int w);

/*
 * Replaces group_keypoints(all_keypoints_by_type, pafs, 20, min_paf_score, demo)
 * (modules/keypoints.py:51-201) for n images at once.
 * kpts/counts/kpt_start: as written by lwp_extract_keypoints (18 channels).
 * pafs: [n][H][W] pixels x 38 used channels, pixel stride paf_ld floats (the UPSAMPLED maps).
 * demo != 0: sample coordinates are truncated (demo.py:100 passes demo=True), else rounded half-to-even.
 * pose_entries: [n][cap_poses][20] doubles (18 key-point ids or -1, [18] score, [19] count);
 * n_poses: [n].  overflow: [n], non-zero when cap_connections / cap_poses was exceeded.
 */
int lwp_group_keypoints(const lwp_keypoint *kpts, const int32_t *counts, const int32_t *kpt_start, int cap_kpts,
                        const float *pafs, int n, int H, int W, int paf_ld, int demo, double min_paf_score,
                        double *pose_entries, int32_t *n_poses, int cap_poses, int cap_connections,
                        void *workspace, size_t workspace_bytes, int32_t *overflow, void *stream);

/*
 * Same as lwp_group_keypoints, but every PAF sample is computed on the fly from the stride-8 network
 * output (4x4 source patch, OpenCV's operation order) instead of being read from a materialised
 * up-sampled map (demo.py:76 + :100 in one pass).  src: [n][h][w] pixels, pixel stride ld floats, the 38 PAF
 * channels start at the pointer.  lwp_group_keypoints* only SET overflow flags; lwp_extract_keypoints* clear them.
 */
int lwp_group_keypoints_fused(const lwp_keypoint *kpts, const int32_t *counts, const int32_t *kpt_start, int cap_kpts,
                              const float *src, int n, int h, int w, int ld, int H, int W, double inv_scale_x,
                              double inv_scale_y, int demo, double min_paf_score, double *pose_entries,
                              int32_t *n_poses, int cap_poses, int cap_connections, void *workspace,
                              size_t workspace_bytes, int32_t *overflow, void *stream);

/*
 * Replaces the result post-conversion of demo.py:101-115 for n images at once: key-point coordinates mapped back to
 * the original frame, (x * stride / upsample_ratio - pad) / scale evaluated in float64 exactly like the reference's
 * Python expression, truncated by int(), gathered per pose into an [18][2] int32 table (-1, -1 for a missing
 * key-point); plus Pose.get_bbox (modules/pose.py:30-39: cv2.boundingRect of the found key-points = x, y, w, h) and the
 * pose confidence pose_entries[n][18] (demo.py:114).
 * xform: [n][3] doubles (pad_left = pad[1], pad_top = pad[0], scale) as returned by infer_fast for each image.
 * pose_kpts: [n][cap_poses][18][2]; bbox: [n][cap_poses][4]; confidence: [n][cap_poses]; entries >= n_poses[i] untouched.
 */
int lwp_pose_convert(const double *pose_entries, const int32_t *n_poses, int cap_poses, const lwp_keypoint *kpts,
                     const int32_t *kpt_start, int cap_kpts, int n, double stride, double upsample_ratio,
                     const double *xform, int32_t *pose_kpts, int32_t *bbox, double *confidence, void *stream);

/* dst[i] = src[i] (items of bytes_per_item bytes, multiple of 16) for every i < n with flags[i] != 0.  The batched
 * pipeline uses it to keep the network output of exactly those frames whose fixed-capacity tables overflowed (flags = the
 * overflow array of lwp_extract_keypoints* / lwp_group_keypoints*), so that they can be re-processed with larger tables. */
int lwp_copy_flagged(const void *src, void *dst, const int32_t *flags, int n, size_t bytes_per_item, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Network forward: a "plan" is a recorded list of layer launches with pre-built TMA descriptors
 * for one (batch, height, width, dtype).  It replaces PoseEstimationWithMobileNet.forward
 * (models/with_mobilenet.py:114-123); the layer list itself is built by the Python module mirror
 * from the same state_dict the reference uses.
 * ------------------------------------------------------------------------------------------- */
typedef struct lwp_plan lwp_plan;

int lwp_plan_create(int dtype, lwp_plan **out);
void lwp_plan_destroy(lwp_plan *p);
int lwp_plan_num_ops(const lwp_plan *p);

/*
 * Stem conv(3, 32, stride=2, bias=False) + BN + ReLU (models/with_mobilenet.py:93, modules/conv.py:4-10).
 * Input is taken at run time (lwp_plan_run's x): NCHW float32 [n][3][H][W].
 * w: [32][3][3][3] float32 (OIHW, as in the state_dict); scale/shift: folded BN, float32 [32].
 * out: NHWC [n][H/2][W/2][32] of the plan dtype.
 */
int lwp_plan_add_stem(lwp_plan *p, const float *w, const float *scale, const float *shift, void *out, int n, int H,
                      int W);

/*
 * Same stem, but fed by the raw frame: x at run time is uint8 [n][H][W][3] (BGR, already at the network size) and
 * the kernel applies val.normalize ((pixel - mean[c]) * img_scale, val.py:30-33 / demo.py:60) on the fly, evaluated in
 * double and rounded to float32 exactly like the reference's NumPy expression followed by .float().  Zero padding of
 * the convolution applies to the normalised values.  4x fewer input bytes over PCIe / HBM than float32 NCHW.
 */
int lwp_plan_add_stem_u8(lwp_plan *p, const float *w, const float *scale, const float *shift, void *out, int n, int H,
                         int W, const double *img_mean3, double img_scale);

/*
 * The full-resolution front end of the backbone as one kernel (bf16 plans): model[0] conv(3, 32, stride=2), model[1]
 * conv_dw(32, 64) and the depthwise half of model[2] conv_dw(64, 128, stride=2) of models/with_mobilenet.py:92-95
 * (Conv2d + BatchNorm2d + ReLU each, modules/conv.py:4-22); with input_u8 also val.normalize (val.py:30-33) of the raw
 * uint8 [n][H][W][3] frame, as lwp_plan_add_stem_u8.  The 32- and 64-channel maps at H/2 x W/2 never reach HBM; the
 * result is bit-identical to the four separate ops.  stem_w: [32][27] float32 (OIHW flattened); dw*_w: [9][C] float32
 * tap-major (C = 32, 64); pw_w: [64][32] in the plan dtype, K-major; *_scale / *_shift: folded BN, float32.
 * out: NHWC [n][H/4][W/4][64] of the plan dtype (the input of model[2]'s 1x1 conv).  H, W multiples of 4.
 */
int lwp_plan_add_frontend(lwp_plan *p, const float *stem_w, const float *stem_scale, const float *stem_shift,
                          const float *dw1_w, const float *dw1_scale, const float *dw1_shift, const void *pw_w,
                          const float *pw_scale, const float *pw_shift, const float *dw2_w, const float *dw2_scale,
                          const float *dw2_shift, void *out, int n, int H, int W, int input_u8, const double *img_mean3,
                          double img_scale);

/*
 * Depthwise 3x3 conv (groups == channels, bias=False) + per-channel scale/shift + activation
 * (modules/conv.py:15-17 conv_dw, :27-28 conv_dw_no_bn).  NHWC in / NHWC out, plan dtype.
 * w: [9][C] float32, tap-major (tap = ky*3 + kx; the host mirror transposes the state_dict's [C][1][3][3]);
 * pad == dilation (the reference always uses padding=dilation for 3x3).  C must be a multiple of 8.
 */
int lwp_plan_add_depthwise(lwp_plan *p, const void *in, void *out, const float *w, const float *scale,
                           const float *shift, int n, int H, int W, int C, int stride, int dilation, int act);

/*
 * Dense convolution as an implicit GEMM on tcgen05 tensor cores: 1x1 (taps == 1) or 3x3 (taps == 9,
 * pad == dilation), stride 1 (modules/conv.py:4-10,19,30; models/with_mobilenet.py:10,16,28-38,51-54,74-79).
 *   in : NHWC [n][H][W] pixels, Cin channels (multiple of the 128-byte K block), pixel stride in_ld elements
 *   w  : [Cout_pad][taps][Cin] elements of the plan dtype, K-major (packed by the host mirror), Cout_pad
 *        a multiple of 16
 *   y = act(acc * scale[c] + shift[c]) (+ residual[pixel][c] if residual != NULL), c < Cout
 *   out: plan dtype, pixel stride out_ld elements (may be NULL)
 *   out_f32: optional float32 copy of the result, pixel stride out_f32_ld floats (heads feed the
 *        post-processing in float32 whatever the plan dtype)
 */
int lwp_plan_add_conv_gemm(lwp_plan *p, const void *in, int in_ld, const void *w, const float *scale,
                           const float *shift, const void *residual, int res_ld, void *out, int out_ld,
                           float *out_f32, int out_f32_ld, int n, int H, int W, int Cin, int Cout, int taps,
                           int dilation, int act);

/*
 * A RefinementStageBlock's second 3x3 conv (+ BN + ReLU + the residual add, models/with_mobilenet.py:53-60) and the
 * `initial` 1x1 conv of the NEXT block (:52,57) as one back-to-back tcgen05 kernel (bf16 plans, 128 -> 128 -> 128): the
 * block output stays in tensor memory, only the 1x1's output [pixels][out_ld] is written.  Bit-identical to the two
 * separate lwp_plan_add_conv_gemm ops.  w: [128][9*Cin], w2: [128][128], both K-major in the plan dtype.  Returns
 * LWP_ECAP (nothing recorded) when the 3x3 cannot run on the CTA-pair strip kernel; the caller then records two ops.
 */
int lwp_plan_add_conv3x3_pw(lwp_plan *p, const void *in, int in_ld, const void *w, const float *scale, const float *shift,
                            const void *residual, int res_ld, int act, const void *w2, const float *scale2,
                            const float *shift2, int act2, void *out, int out_ld, int n, int H, int W, int Cin,
                            int dilation);

/* Both 1x1 layers of a stage's heads (models/with_mobilenet.py:33-38,74-79: conv 128->512|128 + ReLU, conv ->19|38,
   the two heads stacked / block-diagonal) as ONE back-to-back GEMM kernel: the c_mid-channel intermediate never leaves
   the SM.  bf16 plans only.  in: [n_pixels][in_ld] plan dtype; w1: [c_mid][c_in], w2: [64][c_mid] (K-major, plan
   dtype); scale/shift: folded bias (+ ReLU after the first layer, none after the second); out_f32: [n_pixels][out_f32_ld]
   float32 heads (64 columns written); out: optional plan-dtype copy, pixel stride out_ld, or NULL. */
int lwp_plan_add_heads_fused(lwp_plan *p, const void *in, int in_ld, const void *w1, const float *scale1,
                             const float *shift1, int c_mid, const void *w2, const float *scale2, const float *shift2,
                             void *out, int out_ld, float *out_f32, int out_f32_ld, int n_pixels, int c_in);

/*
 * Fused depthwise-separable block (modules/conv.py:13-32 conv_dw / conv_dw_no_bn): depthwise 3x3 stride 1
 * (pad == dilation, 1 or 2) + scale/shift + dw_act, whose result is written straight into the shared-memory A
 * operand of the following 1x1 convolution's tcgen05 GEMM (+ scale/shift + act (+ residual)).  The depthwise
 * output never goes to global memory.  in: NHWC [n][H][W][Cin] (dense), dw_w: [9][Cin] float32,
 * w: [Cout_pad][Cin] plan dtype; Cin a multiple of the 128-byte K block, Cout_pad (multiple of 64) must divide 512.
 */
int lwp_plan_add_dwpw(lwp_plan *p, const void *in, const float *dw_w, const float *dw_scale, const float *dw_shift,
                      int dw_act, int dilation, const void *w, const float *scale, const float *shift, int act,
                      const void *residual, int res_ld, void *out, int out_ld, int n, int H, int W, int Cin, int Cout);

/*
 * Weight-resident fused depthwise-separable block for the THIN layers (modules/conv.py:13-32 conv_dw / conv_dw_no_bn as
 * used by models/with_mobilenet.py:94-99 and :12-16): depthwise 3x3, pad 1, stride 1 or 2, + scale/shift + dw_act,
 * written straight into the shared-memory A operand of the 1x1 convolution's tcgen05 GEMM (+ scale/shift + act
 * (+ residual)); the whole 1x1 weight matrix stays in shared memory for the CTA's life.  Same argument meaning as
 * lwp_plan_add_dwpw (stride instead of dilation); out is [n][(H-1)/stride+1][(W-1)/stride+1] pixels.
 * Cin a multiple of the 128-byte K block, Cout <= 256.  Returns LWP_ECAP when the weights + rings do not fit in
 * shared memory (the caller then records the two-kernel form).
 */
int lwp_plan_add_sepconv(lwp_plan *p, const void *in, const float *dw_w, const float *dw_scale, const float *dw_shift,
                         int dw_act, int stride, const void *w, const float *scale, const float *shift, int act,
                         const void *residual, int res_ld, void *out, int out_ld, int n, int H, int W, int Cin, int Cout);

/* NHWC (plan dtype or float32) -> NCHW float32, for the tensors `forward` returns at the module boundary. */
int lwp_plan_add_nhwc_to_nchw(lwp_plan *p, const void *in, int in_ld, int in_is_f32, int c0, int c, float *out, int n,
                              int H, int W);

/* Enqueue every recorded op on `stream`.  x: input of the stem -- NCHW float32, or uint8 NHWC for a plan built with
 * lwp_plan_add_stem_u8 (may be NULL if the plan has no stem). */
int lwp_plan_run(lwp_plan *p, const void *x, void *stream);
/* Enqueue ops [first, last) only (per-layer timing / parity tests). */
int lwp_plan_run_range(lwp_plan *p, const void *x, int first, int last, void *stream);
/* number of kernel launches one lwp_plan_run issues */
int lwp_plan_num_launches(const lwp_plan *p);
/* Synchronising read of the plan's device error flag: 0 = ok, else the id of the pipeline wait that timed
 * out inside a tcgen05 GEMM kernel (the kernels never spin forever). */
int lwp_plan_error_flag(lwp_plan *p);

/* ---------------------------------------------------------------------------------------------
 * The whole path in three calls, for hosts that are not Python.  lwp_net_load rebuilds a runnable network from a
 * self-contained blob (written once from a reference state_dict by the Python mirror: Plan.export_blob(); BatchNorm
 * already folded, weights already packed for the tcgen05 kernels, the layer list with every argument) -- it replaces
 * PoseEstimationWithMobileNet(...) + load_state(net, checkpoint) + net.cuda() (models/with_mobilenet.py:89-112,
 * modules/load_state.py:4-15) for one (dtype, batch, height, width); lwp_net_forward replaces forward (:114-123);
 * lwp_postprocess replaces the cubic up-sampling + 18 x extract_keypoints + group_keypoints of demo.py:72-100.
 * The library owns the network's device memory (allocated in lwp_net_load, freed in lwp_net_destroy).
 * ------------------------------------------------------------------------------------------- */
typedef struct lwp_net lwp_net;
int lwp_net_load(const void *blob_host, size_t bytes, lwp_net **out);
void lwp_net_destroy(lwp_net *net);
int lwp_net_info(const lwp_net *net, int *dtype, int *n, int *H, int *W, int *n_stages);
/* x: NCHW float32 [n][3][H][W] (or uint8 [n][H][W][3] for a blob exported with the uint8 stem), device memory.
 * with_nchw != 0 also fills the NCHW float32 tensors `forward` returns (lwp_net_output_nchw). */
int lwp_net_forward(lwp_net *net, const void *x, int with_nchw, void *stream);
/* float32 heads of stage `stage` (-1 = last): [n * H/8 * W/8] pixels x *ld floats, 19 heat-maps + 38 PAFs + zero pad */
int lwp_net_heads(const lwp_net *net, int stage, const float **heads, int *ld);
/* index-th entry of the list `forward` returns: [hm_0, paf_0, ..., hm_R, paf_R], NCHW float32 */
int lwp_net_output_nchw(const lwp_net *net, int index, const float **out);

size_t lwp_postprocess_workspace_bytes(int n, int cap_kpts, int cap_candidates, int cap_connections, int cap_poses);
/* heads: [n][h][w] pixels x ld floats (19 heat-map + 38 PAF channels first), e.g. lwp_net_heads(net, -1, ...).
 * Outputs as for lwp_extract_keypoints_fused + lwp_group_keypoints_fused (up-sampled size = h, w times upsample_ratio). */
int lwp_postprocess(const float *heads, int n, int h, int w, int ld, int upsample_ratio, int demo, double min_paf_score,
                    lwp_keypoint *kpts, int32_t *counts, int32_t *kpt_start, int cap_kpts, int cap_candidates,
                    double *pose_entries, int32_t *n_poses, int cap_poses, int cap_connections, void *workspace,
                    size_t workspace_bytes, int32_t *overflow, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LWPOSE_B200_H */
