/* calib_b200.h - C ABI of the B200-native calibration hot path.
 *
 * Drop-in boundary for the per-frame calibration path of
 * NikolasEnt/soccernet-calibration-sportlight.  The reference is pure Python with
 * no FFI of its own: its plug-in points are Python classes
 * selected by Hydra `_target_` strings.  The Python package in
 * soccernet_calibration_sportlight_b200/ mirrors those classes and binds the entry
 * points below with ctypes.  Each entry point cites the
 * reference interface it replaces as `file:line` under /root/reference.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name starts with `h_` (host);
 *   - the caller owns every buffer; nothing is allocated behind the ABI;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all work is enqueued asynchronously on it, no host synchronisation;
 *   - return value: 0 = CAL_OK, negative = CalStatus; cal_last_error() returns a
 *     thread-local human-readable message for the last failure;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns CAL_E_CUDA.
 */
#ifndef CALIB_B200_H_
#define CALIB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAL_ABI_VERSION 1

typedef enum CalStatus {
  CAL_OK = 0,
  CAL_E_INVALID = -1,     /* bad argument (shape, alignment, null pointer) */
  CAL_E_UNSUPPORTED = -2, /* shape outside what the kernels were built for */
  CAL_E_CUDA = -3,        /* CUDA runtime / driver error (message has details) */
} CalStatus;

int cal_abi_version(void);
const char* cal_last_error(void);

/* Shared memory (bytes per SM, 0..98304) the persistent tcgen05 kernels leave unused so that blocks of
 * cal_camera_solve (one frame each: ~39 KB, 64 threads x <= 168 registers) can be co-resident with
 * them: the reference solves the cameras of a batch in a 16-process CPU pool while the GPU already
 * runs the next batch (src/utils/make_submit.py:53-73); here the solve of batch i runs UNDER the
 * networks of batch i+1 on a second stream.  Default 0, or the environment variable CAL_SMEM_HEADROOM. */
int cal_set_smem_headroom(int bytes);

/* ------------------------------------------------------------------ decode -- */

/* Keypoint heat-map decode.
 * Replaces HRNetPredictionTransform.__call__ (src/models/hrnet/transforms.py:228-239).
 *   logp : (B, C, h, w) fp32 log-probabilities, contiguous NCHW
 *   out  : (B, C-1, 3) fp32 [x, y, conf]; the last channel (background) is dropped
 *   x = first argmax over columns of the column maxima of exp(logp), scaled by
 *   W_img / w ; y likewise over rows, scaled by H_img / h ; conf = min of the two
 *   maxima.  exp is the correctly rounded fp32 exponential; ties -> first index. */
int cal_kp_decode(const float* logp, int B, int C, int h, int w, int H_img, int W_img,
                  float* out, void* stream);

/* Line heat-map decode (two peaks per channel with Gaussian suppression).
 * Replaces EHMPredictionTransform.mask_heat_points_gauss and __call__
 * (src/models/line/transforms.py:216-280).
 *   heat : (B, C, h, w) fp32 probabilities
 *   out  : (B, C, 2, 3) fp32 [x*scale, y*scale, value] for peak 1 and peak 2
 *   scale = 1 reproduces mask_heat_points_gauss; __call__ uses its `scale`. */
int cal_line_decode(const float* heat, int B, int C, int h, int w, double sigma, float scale,
                    float* out, void* stream);

/* ------------------------------------------------------- HRNet building ops -- */
/* Activations are fp16 NHWC with the channel count padded to a multiple of 64
 * (zero-filled pad lanes); accumulation is fp32 on tcgen05 tensor cores.  These
 * ops together replace HighResolutionNet.forward (src/models/hrnet/hrnet.py:437-511,
 * src/models/line/hrnet.py:185-249); the layer schedule itself lives in the Python
 * mirror (hrnet.py in the package), one call per fused conv+BN(+add)(+ReLU). */

typedef struct CalConvArgs {
  const void* x;     /* fp16 NHWC (B, Hin, Win, Cin_pad) */
  const void* w;     /* fp16 (Cout_rows, taps, Cin_pad), BN folded, K-major */
  const float* bias; /* fp32 (Cout_pad), BN folded; pad lanes zero */
  const void* res;   /* optional fp16 NHWC (B, Hout, Wout, Cout_pad) added before ReLU */
  void* y;           /* mode 0: fp16 NHWC (B, Hout, Wout, Cout_pad);
                        mode 1/2: fp32 NCHW (B, n_classes, Hout, Wout) */
  int32_t B, Hin, Win, Cin_pad;
  int32_t Hout, Wout, Cout_pad;
  int32_t Cout_rows; /* rows in w: real Cout rounded up to 16 */
  int32_t ksize;     /* 1 or 3 (padding = ksize/2) */
  int32_t stride;    /* 1 or 2 */
  int32_t relu;      /* mode 0 only */
  int32_t mode;      /* 0 = fp16 NHWC; 1 = LogSoftmax over n_classes -> fp32 NCHW
                        (hrnet.py:329); 2 = Softmax (line/hrnet.py:101) */
  int32_t n_classes; /* modes 1/2: 58 / 23 */
  int32_t Cin;       /* real input channels (<= Cin_pad; 0 = Cin_pad): K steps over pad lanes are skipped */
  int32_t w_slices;  /* 0: w is (Cout_rows, taps*Cin_pad) K-major;
                        1: w is slice-major (taps*Cin_pad/64, Cout_rows, 64): every (tap, 64-channel chunk)
                           slice is one contiguous block - what a streamed weight ring wants (3x3 stride-1 only) */
} CalConvArgs;

/* conv (+folded BN) (+residual) (+ReLU) as an implicit GEMM on tcgen05/TMEM with
 * TMA-staged operands.  Replaces nn.Conv2d + BatchNorm2d + add + ReLU groups
 * (hrnet.py:42-58, 79-99, 184-213, 260-266, 316-330, 366-388). */
int cal_conv2d(const CalConvArgs* h_args, void* stream);

typedef struct CalBasicBlockArgs {
  const void* x;      /* fp16 NHWC (B, H, W, 64): the block input, also its residual */
  const void* w1;     /* fp16 slice-major (9, rows, 64): conv1 + bn1 folded (CalConvArgs.w with w_slices = 1) */
  const float* bias1; /* fp32 (64) */
  const void* w2;     /* conv2 + bn2, same layout */
  const float* bias2;
  void* y;            /* fp16 NHWC (B, H, W, 64) */
  int32_t B, H, W;
  int32_t C_pad;      /* 64 */
  int32_t rows;       /* weight rows per tap: the channel count rounded up to 16 (16, 32 or 48) */
  int32_t C;          /* real channels (K steps over pad lanes are skipped) */
} CalBasicBlockArgs;

/* BasicBlock.forward (src/models/hrnet/hrnet.py:29-58) with inplanes == planes <= 48 and no downsample, in one
 * kernel: y = relu(bn2(conv2(relu(bn1(conv1(x))))) + x), the intermediate tensor held in shared memory (half the
 * HBM traffic of two cal_conv2d launches, bit-identical results).  CAL_E_UNSUPPORTED for other shapes: the
 * caller runs the two convs. */
int cal_basicblock(const CalBasicBlockArgs* h_args, void* stream);

/* Stem conv1: 3x3 stride-2 conv 3->64 + BN + ReLU straight from the fp32 NCHW
 * frame tensor (hrnet.py:450-452).  x: (B,3,H,W) fp32 in [0,1] BGR;
 * w: fp32 (64, 27) BN folded [co][ci*9+ky*3+kx]; y: fp16 NHWC (B,Ho,Wo,64). */
int cal_stem_conv(const float* x, const float* w, const float* bias, void* y,
                  int B, int H, int W, int Ho, int Wo, void* stream);
/* The same from the frame as cv2.imread leaves it: x (B,H,W,3) uint8 HWC BGR; T.ToTensor's float32
 * division by 255 (src/models/hrnet/transforms.py:59-68, src/utils/make_submit.py:66) is folded into the
 * load - bit-identical to cal_stem_conv on ToTensor's output, a quarter of the host->device bytes. */
int cal_stem_conv_u8(const uint8_t* x, const float* w, const float* bias, void* y,
                     int B, int H, int W, int Ho, int Wo, void* stream);

#define CAL_MAX_SOURCES 6
typedef struct CalCombineArgs {
  void* y;        /* fp16 NHWC (B, H, W, C_pad) */
  int32_t B, H, W, C_pad;
  int32_t n_src;
  const void* src[CAL_MAX_SOURCES]; /* fp16 NHWC (B, h_i, w_i, C_pad) */
  int32_t src_h[CAL_MAX_SOURCES];   /* == H,W: plain add; else bilinear, align_corners=True */
  int32_t src_w[CAL_MAX_SOURCES];
  const float* bias;                /* optional fp32 (C_pad) */
  int32_t relu;
  int32_t C;                        /* real channels (0 = C_pad): the pad lanes are written as zeros without being gathered */
} CalCombineArgs;

/* y = [relu]( bias + sum_i up_i(src_i) ): the multi-resolution fuse of
 * HighResolutionModule.forward (hrnet.py:229-246) and the head's upsample+concat
 * (hrnet.py:489-509, line/hrnet.py:236-245) after commuting the 1x1 conv with the
 * bilinear interpolation. */
int cal_fuse_combine(const CalCombineArgs* h_args, void* stream);

#define CAL_MAX_LOW 4
typedef struct CalHeadArgs {
  const void* full;      /* fp16 NHWC (B, H, W, 64): the full-resolution source of the head (x_stem for the
                            keypoint net, branch 0 for the line net), channels padded to 64 */
  const void* w_full;    /* fp16 (Cout_rows, 64) K-major: the columns of the first head conv that multiply
                            `full`, BN folded */
  const void* low[CAL_MAX_LOW];  /* fp16 NHWC (B, h_i, w_i, Cout_pad): p_i = W1_i * y_i, the same conv applied
                            to every lower-resolution branch at ITS resolution (no bias) */
  int32_t low_h[CAL_MAX_LOW], low_w[CAL_MAX_LOW];
  int32_t n_low;
  const float* bias;     /* fp32 (Cout_pad): conv bias with BN folded */
  void* z;               /* fp16 NHWC (B, H, W, Cout_pad); unused (may be NULL) when w2 is given */
  int32_t B, H, W, Cf_pad, Cout_pad, Cout_rows;
  /* optional chained tail: the final 1x1 conv + (Log)Softmax (hrnet.py:325-329, line/hrnet.py:97-101)
     applied to the ReLU'd tile while it is still on chip, z never reaches HBM */
  const void* w2;        /* fp16 (64, Cout_pad) K-major, rows >= n_classes zero; NULL = write z */
  const float* bias2;    /* fp32 (64) */
  float* heat;           /* fp32 NCHW (B, n_classes, H, W) */
  int32_t n_classes;     /* <= 64 */
  int32_t mode;          /* 1 = LogSoftmax, 2 = Softmax */
} CalHeadArgs;

/* z = ReLU( W1_full * full + sum_i bilinear_up(p_i) + bias ): the upsample + concat + first 1x1 conv
 * + BN + ReLU of the head (src/models/hrnet/hrnet.py:489-511 with last_layer[0:3] :316-324;
 * src/models/line/hrnet.py:236-249) as one tensor-core accumulation per tile, the bilinear
 * interpolation (align_corners=True) expressed as a small GEMM against the source patches.
 * Returns CAL_E_UNSUPPORTED when a source's footprint does not fit (caller falls back to
 * cal_fuse_combine + cal_conv2d). */
int cal_head_fused(const CalHeadArgs* h_args, void* stream);

/* ------------------------------------------------------------------ engine -- */
/* The whole network behind one call: what `HRNetHeatmap(hrnet_config).forward(x)` is in the reference
 * (src/models/hrnet/model.py:130-150 -> HighResolutionNet.__init__/forward, src/models/hrnet/hrnet.py:255-330,
 * 437-511; src/models/line/hrnet.py:30-249), for a host written in any language.  The library walks the
 * architecture, folds eval-mode BatchNorm, packs the weights (cal_hrnet_create) and owns the layer schedule
 * (cal_hrnet_forward: one launch per fused conv, intermediate tensors from the CUDA stream-ordered
 * allocator - the one exception to "nothing is allocated behind the ABI").  The fields mirror the reference's
 * model_config/hrnet_*.yaml. */
typedef struct CalHrnetStage {
  int32_t num_modules, num_branches;
  int32_t block_type;            /* 0 = BASIC, 1 = BOTTLENECK */
  int32_t num_blocks[4], num_channels[4];
} CalHrnetStage;

typedef struct CalHrnetConfig {
  int32_t kind;                  /* 0 = keypoint net (stem skip, `upscale`, LogSoftmax); 1 = line net (Softmax) */
  int32_t num_classes;           /* 58 / 23 (<= 64) */
  int32_t stem_width;            /* 64 */
  int32_t upscale;               /* keypoint net: head resolution / first-branch resolution (2, or 4 for w48x4) */
  CalHrnetStage stage[4];        /* stage1 .. stage4 */
} CalHrnetConfig;

/* Number of floats of the weight blob: every tensor of the reference's `nn_state_dict` in its own order
 * (conv weight, [conv bias], BN weight, bias, running_mean, running_var; `num_batches_tracked` skipped),
 * each flattened row-major, concatenated. */
int cal_hrnet_weight_count(const CalHrnetConfig* h_cfg, size_t* h_n_floats);
int cal_hrnet_create(const CalHrnetConfig* h_cfg, const float* h_weights, size_t n_floats, void** h_handle);
/* x: (B,3,H,W) fp32 NCHW in [0,1] BGR (x_is_u8 = 0: the reference's input) or (B,H,W,3) uint8 HWC BGR frames as
 * cv2.imread leaves them (x_is_u8 = 1); heat: (B, num_classes, h, w) fp32 NCHW, (h, w) from cal_hrnet_output_shape:
 * log-probabilities (keypoints, H/2 x W/2) or probabilities (lines, H/4 x W/4). */
int cal_hrnet_forward(void* handle, const void* x, int x_is_u8, int B, int H, int W, float* heat, void* stream);
int cal_hrnet_output_shape(void* handle, int H, int W, int* h_n_classes, int* h_h, int* h_w);
long cal_hrnet_launches(void* handle);    /* kernels launched through this handle so far */
int cal_hrnet_destroy(void* handle);

/* ------------------------------------------------------------ camera solve -- */

#define CAL_NUM_KEYPOINTS 57

typedef struct CalSolveParams {
  double pitch_xyz[CAL_NUM_KEYPOINTS * 3]; /* world coordinates (metres) of keypoint ids 0..56: the
                         `pitch` constructor argument of CameraCreator looked up through
                         INTERSECTON_TO_PITCH_POINTS (prediction.py:44-45, ellipse.py:99-157) */
  int32_t algorithm;  /* 0 opencv_calibration, 1 opencv_calibration_multiplane,
                         2 original_voter, 3 voter, 4 iterative_voter
                         (src/models/hrnet/prediction.py:90-96) */
  int32_t img_w, img_h;
  float conf_thresh;
  float conf_threshs[8];
  int32_t n_conf_threshs;
  int32_t min_points, min_points_per_plane, min_points_for_refinement, reliable_thresh;
  float min_focal_length, max_rmse, max_rmse_rel;
} CalSolveParams;

typedef struct CalCameraRecord { /* 128 bytes, the all-gather payload */
  double position[3];
  double rotation[9]; /* row-major world->camera */
  double fx, fy;
  double rmse;        /* mean reprojection L2 of the selected camera (Camera.projection_rmse) */
  int32_t valid;      /* 0 = the reference returns None */
  int32_t branch;     /* which heuristic produced it (solve_cascade.cuh, enum Branch) */
} CalCameraRecord;

/* Batched CameraCreator.__call__ (src/models/hrnet/prediction.py:130-136 and the
 * algorithms it dispatches to, :138-437, with baseline/camera.py:92-119, 366-426):
 * one thread block per frame.
 *   preds    : (B, 57, 3) fp32 [x, y, conf]
 *   line_pts : optional (B, 57, 2) fp64 keypoints from line intersections, NaN = absent
 *   out      : (B) records */
int cal_camera_solve(const float* preds, const double* line_pts, const CalSolveParams* h_params,
                     int B, CalCameraRecord* out, void* stream);

/* Single-camera helpers used by the Camera class mirror.
 * cal_pnp_refine  replaces Camera.refine_camera (baseline/camera.py:105-119);
 * cal_pnp_solve   replaces Camera.solve_pnp     (baseline/camera.py:92-103), i.e. the pose
 *                  cv2.solvePnPRansac(obj, img, K, None) returns: P3P on 4 matches, EPnP on 5,
 *                  seeded 5-point RANSAC (8 px, 100 iterations, 0.99) + refit on the consensus set
 *                  on more (csrc/solve_pnp_cv.cuh).  *ok = 1 reproduced; 0 OpenCV's RANSAC fails
 *                  on these matches (the reference then reads uninitialised memory; rvec/tvec hold
 *                  the least-squares pose); -1 no finite pose (rvec/tvec untouched).
 *   obj (n,3) fp64, img (n,2) fp64, K (9) fp64 row-major, rvec/tvec (3) fp64 in/out; n <= 57 */
int cal_pnp_refine(const double* obj, const double* img, int n, const double* K,
                   double* rvec, double* tvec, void* stream);
int cal_pnp_solve(const double* obj, const double* img, int n, const double* K,
                  double* rvec, double* tvec, int32_t* ok, void* stream);

/* Keypoints from the line model's decoded peaks: get_line_data (src/utils/export_line_result.py:
 * 85-131, slope/intercept per line class with both peaks at p >= prob_thre) followed by the
 * line-pair intersections of CameraCreator.__init__ (prediction.py:110-124, 643-653), in fp64
 * on the fp32 peaks, as the reference's pinned numpy 1.24 evaluates them (scalar promotion).
 *   peaks    : (B, 23, 2, 3) fp32 [x, y, p], already in image pixels (cal_line_decode's scale)
 *   pair_a/b : (57) int32 line-class channel indices whose intersection is keypoint i, -1 = none
 *              (LINE_INTERSECTIONS, src/datatools/intersections.py:13-44)
 *   out      : (B, 57, 2) fp64, NaN where the keypoint is not produced */
int cal_line_points(const float* peaks, int B, int n_lines, const int32_t* pair_a, const int32_t* pair_b,
                    float prob_thre, double* out, void* stream);

/* ------------------------------------------------------------------- debug -- */
/* Dumps the shared-memory image of one TMA box load (used by tests to pin the
 * tensor-map conventions the conv kernel relies on). */
int cal_debug_tma_probe(const void* x, int B, int H, int W, int C, int box_w, int box_h,
                        int estride, int c0, int x0, int y0, int n0, void* out_smem_16k,
                        void* stream);

/* Experiment pinning a hardware convention the 3x3 kernel relies on: a SWIZZLE_128B K-major
 * operand descriptor whose start address is shifted by whole 128-byte rows.
 * D(128x64 fp32) = X[shift : shift+128, :] * W^T with X (256,64), W (64,64) fp16. */
int cal_debug_shift_mma(const void* x_256x64, const void* w_64x64, int shift,
                        int base_offset_mode, float* out_128x64, void* stream);

/* Experiment pinning the MN-major (N contiguous) B-operand descriptor convention used by the
 * fused head's interpolation GEMM: D(128x128 fp32) = X(128x64) * Y(64x128), Y row-major. */
int cal_debug_mn_mma(const void* x_128x64, const void* y_64x128, int mode, float* out_128x128,
                     void* stream);

/* Experiment: cycles for `iters` back-to-back tcgen05.mma (M = 128, K = 16) as a function of N, of the
 * row shift of the A operand's start address and of `flags`: bit 0 = MN-major B operand, bits 8-15 =
 * a tcgen05.commit after every that many groups of 4 MMAs (a power of two; 0: only at the end), bit 1 =
 * also an mbarrier wait after each such commit, bit 3 = also a tcgen05.fence::after_thread_sync, bit 4 / bit 5 =
 * that wait is an mbarrier.test_wait / a plain shared-memory flag poll instead of mbarrier.try_wait, bit 6 = the wait goes before the commit, bit 2 = issue no MMAs (commits only),
 * bits 16-23 = CTAs launched (0: one); out_cycles has one
 * entry per CTA. */
int cal_debug_mma_rate(int N, int shift_rows, int iters, int flags, long long* out_cycles, void* stream);

/* Experiment: cycles per tile of the tcgen05.mma trains the 3x3 kernels issue (csrc/probe.cu), in isolation:
 * pattern 0 = filter-row grouping (N = 2n and N = n per filter row), 1 = the same with a fixed A address,
 * 2 = nine N = n taps, 3 = three taps side by side (N = 3n), 4 = N = 2n throughout, 5 = N = 256, 6 = pattern 0
 * with the N = 2n trains first; nk = K steps per chunk; flags: bit 0 = random operands (else zeros), bit 1 = two
 * tcgen05.commit per tile, bit 2 = four warps reading TMEM meanwhile, bits 8-11 = accumulators rotated over. */
int cal_debug_mma_pattern(int pattern, int n, int nk, int iters, int flags, int ctas, long long* out_cycles,
                          void* stream);

/* ------------------------------------------------------------------ metric -- */

/* The official camera-calibration metric on the GPU, one thread block per frame: the step after the path,
 * what EvalAImetric runs per validation batch (src/models/hrnet/metrics.py:107-137, 181-211).
 * Replaces get_polylines (baseline/evaluate_camera.py:14-107: the sampled pitch model projected by the
 * camera, clipped at the image border), distance_to_polyline (:110-160) and evaluate_camera_prediction
 * (:163-229) for the annotation and for its mirrored labelling (evaluate_extremities.py:24-34), and the
 * choice between the two of Evaluator.__call__ (metrics.py:112-135).
 *   cams        : (B) camera records (cal_camera_solve's output); valid = 0 -> out.valid = 0 ("missed")
 *   field_pts   : (n_pts, 3) fp64 sampled pitch points (SoccerPitch.sample_field_points), grouped by class
 *   class_off   : (n_proj + 1) int32 offsets of each projectable class into field_pts
 *   class_id    : (n_proj) int32 index of that class among the CAL_EVAL_CLASSES dataset classes
 *   mirror      : (CAL_EVAL_CLASSES) int32 class index under the point reflection through the pitch centre
 *   is_circle   : (CAL_EVAL_CLASSES) uint8 ('Circle' in the class name: 9 false positives instead of 2)
 *   gt_pts      : (B, CAL_EVAL_CLASSES, max_gt, 2) fp64 annotated points in pixels
 *   gt_count    : (B, CAL_EVAL_CLASSES) int32, -1 = class absent from the annotation
 *   poly        : (B, n_proj, max_poly, 2) fp64 polylines (output, or input when from_polylines & 1;
 *                 from_polylines & 2: report the labelling as annotated instead of the better one)
 *   poly_count  : (B, n_proj) int32 (ditto)
 *   dist        : (B, 2, CAL_EVAL_CLASSES, max_gt) fp64 scratch: point-to-polyline distances per labelling */
#define CAL_EVAL_CLASSES 28
typedef struct CalEvalRecord {
  double accuracy;                         /* confusion[0][0] / sum of the chosen labelling */
  double confusion[4];                     /* [[tp, fp], [fn, 0]] over classes */
  double l2_sum;                           /* sum of the point-to-polyline distances of the common classes */
  int32_t l2_count;
  int32_t labelling;                       /* 0 = as annotated, 1 = mirrored */
  int32_t valid;
  int32_t pad;
  double per_class[CAL_EVAL_CLASSES][4];   /* per-class confusion matrices, rows of zeros for untouched classes */
  uint8_t touched[CAL_EVAL_CLASSES];       /* class has an entry in the reference's per-class dictionary */
  uint8_t pad2[4];
} CalEvalRecord;

int cal_evaluate_cameras(const CalCameraRecord* cams, int B, const double* field_pts, const int32_t* class_off,
                         const int32_t* class_id, int n_proj, const int32_t* mirror, const uint8_t* is_circle,
                         const double* gt_pts, const int32_t* gt_count, int max_gt, int img_w, int img_h,
                         double threshold, double* poly, int32_t* poly_count, int max_poly, int from_polylines,
                         double* dist, CalEvalRecord* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CALIB_B200_H_ */
