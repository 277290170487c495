/*
 * epos_b200.h -- C ABI of the B200-native EPOS inference hot path.
 *
 * The reference (thodan/epos) has no FFI for this path except the pybind11 module
 * `pyprogressivex`; everything else is Python calling TensorFlow-1.12 graph ops.  Each entry point
 * below replaces the reference interface cited next to it and is what a maintainer would bind
 * (ctypes stub in INTEGRATION.md).  Conventions:
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless named *_host;
 *   - the caller owns every buffer, the library never frees caller memory;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), except where noted;
 *   - return 0 on success, negative epos_status otherwise; never throws; epos_last_error() gives text;
 *   - activations are NHWC, fp32, or "split-bf16": two bf16 planes [2][rows][ld] (hi, lo) with
 *     value = float(hi) + float(lo) (the operand format of the error-compensated tensor-core GEMM).
 */
#ifndef EPOS_B200_H_
#define EPOS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  EPOS_OK = 0,
  EPOS_ERR_INVALID_ARG = -1,
  EPOS_ERR_CUDA = -2,
  EPOS_ERR_UNSUPPORTED = -3,
  EPOS_ERR_NO_DEVICE = -4
} epos_status;

/* Text of the last error on the calling thread. */
const char* epos_last_error(void);
/* Library/ABI version, compiled arch (100 for sm_100a). */
int epos_version(void);
int epos_compiled_arch(void);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
uint64_t epos_launch_count(void);

/* ---- CNN ops: replace the TF ops built by model.predict (epos_lib/model.py:629-687) ---------------- */

/* entry_flow/conv1_1 (Xception, Cout = 32) and conv1_1 of the ResNet beta root (Cout = 64,
 * net_resnet_v1_beta.py:108): (2/255)x-1 preprocessing (feature.py:171-174) + 3x3 stride-2 conv with
 * explicit padding (resnet_utils.conv2d_same, external/slim/nets/resnet_utils.py:77-122) + folded BN +
 * ReLU.  x [B,H,W,3] f32 in [0,255]; w [3][3][3][Cout] (HWIO, BN scale folded); outputs
 * y [B,H/2,W/2,Cout] f32 and/or y_split [2][B*H/2*W/2][Cout] bf16 (either may be NULL). */
int epos_conv3x3_rgb_s2(const float* x, const float* w, const float* bias, float* y, uint16_t* y_split,
                        int B, int H, int W, int Cout, void* stream);

/* slim.max_pool2d(3, stride 2, padding='SAME') of the ResNet root (net_resnet_v1_beta.py:187).
 * x [B,H,W,C] f32 -> y_f32 [B,ceil(H/2),ceil(W/2),C] and/or y_split (split-bf16 planes). */
int epos_maxpool3x3_s2(const float* x, float* y_f32, uint16_t* y_split, int B, int H, int W, int C,
                       void* stream);

/* resnet_utils.subsample (external/slim/nets/resnet_utils.py:59-74): every factor-th pixel of
 * x [B,H,W,ldx] (first C channels) -> y [B,ceil(H/f),ceil(W/f),C] f32 (identity shortcut of a strided
 * bottleneck unit, net_resnet_v1_beta.py:69-70). */
int epos_subsample_f32(const float* x, int ldx, float* y, int B, int H, int W, int C, int factor,
                       void* stream);

/* Dense 3x3 stride-1 SAME conv + folded BN + ReLU (entry_flow/conv1_2).  w [3][3][Cin][Cout]. */
int epos_conv3x3_dense(const float* x, const float* w, const float* bias, float* y,
                       int B, int H, int W, int Cin, int Cout, void* stream);

/* Depthwise 3x3 conv (net_xception.py:167-177 separable_conv2d_same depthwise half; model.py:80-89):
 * optional ReLU on the input (xception_module pre-activation, net_xception.py:276), dilation `rate`,
 * stride 1 (TF SAME) or 2 (fixed_padding then VALID), folded BN bias, optional ReLU on the output.
 * x [B,H,W,ldx] f32 (first C channels used); w [9][C] f32; y_f32 [B,Ho,Wo,C] and/or
 * y_split [2][B*Ho*Wo][ldy_split] (either may be NULL).  ldy_split >= C lets the caller pad the bf16 row
 * pitch to a multiple of 32 bytes (C = 728: 736), which keeps the rows sector-aligned for this kernel's
 * stores and for the TMA loads of the GEMM that consumes them. */
int epos_dwconv3x3(const float* x, int ldx, const float* w, const float* bias,
                   float* y_f32, uint16_t* y_split, int ldy_split,
                   int B, int H, int W, int C, int stride, int rate, int relu_in, int relu_out,
                   void* stream);

/* Pointwise (1x1) convolution = GEMM on tcgen05 tensor cores with split-bf16 operands (3 MMAs per
 * product: hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM):
 *   D[m][n] = act( sum_k A[m][k] * Wt[n][k] + bias[(m / bias_group_rows)][n] ) (+ residual[m][n])
 * Replaces slim.conv2d 1x1 + BatchNorm (+ReLU) (+ residual add), net_xception.py:178-182,297-313,
 * model.py:90-97,223-258,350-352,448-456.
 * a_split [2][M][lda] bf16; w_split [2][N][ldw] bf16 (ldw >= K; a row pitch that is a multiple of 128 bytes
 * -- K = 728 stored with ldw = 768 -- keeps the TMA rows aligned: measured 131 -> 109 us on the 728-channel layers); bias [groups][N] f32 (bias_group_rows = 0: one row);
 * residual f32 with leading dim ldr or NULL; outputs: d_f32 (leading dim ldd) and/or d_split
 * ([2][M][ldd_split], plane stride = split_plane_stride elements), either may be NULL.
 * relu: 0 = identity, 1 = ReLU (after the residual add, as the ResNet bottleneck needs,
 * net_resnet_v1_beta.py:88; the Xception units never combine the two), 2 = softmax over aligned groups of 64 output columns fused into the
 * epilogue (tf.nn.softmax over the fragment axis, model.py:676-678; needs N % 64 == 0, f32 output only,
 * no residual). */
int epos_pwconv_gemm(const uint16_t* a_split, int lda, size_t a_plane_stride,
                     const uint16_t* w_split, int ldw, const float* bias, int bias_group_rows,
                     const float* residual, int ldr,
                     float* d_f32, int ldd,
                     uint16_t* d_split, int ldd_split, size_t d_plane_stride,
                     int M, int N, int K, int relu, void* stream);

/* Dense 3x3 (atrous) stride-1 'SAME' convolution + folded BN (+ residual) (+ReLU) as an implicit GEMM on
 * the same tcgen05 kernel: K = 9 taps x C, the A tile of tap (ky,kx) is an 8x16 pixel block shifted by
 * ((ky-1) rate, (kx-1) rate) fetched with a 5-D TMA box (zero fill outside the image = TF SAME padding).
 * Replaces resnet_utils.conv2d_same (external/slim/nets/resnet_utils.py:77-122) at stride 1 as used by
 * net_resnet_v1_beta.py:85,108-110; a stride-2 conv2d_same equals this followed by subsampling
 * (resnet_v1_test.py:72-149).  x_split [2][B,H,W,ldx] bf16 (C % 64 == 0);
 * w_split [2][N][ldw] bf16 (ldw >= 9*C) with k = (ky*3+kx)*C + c; outputs and residual are indexed by the output
 * pixel (b*H + y)*W + x as in epos_pwconv_gemm.  relu: 0 / 1 (applied after the residual add). */
int epos_conv3x3_gemm(const uint16_t* x_split, int ldx, size_t x_plane_stride, const uint16_t* w_split, int ldw,
                      const float* bias, const float* residual, int ldr, float* d_f32, int ldd,
                      uint16_t* d_split, int ldd_split, size_t d_plane_stride,
                      int B, int H, int W, int C, int N, int rate, int relu, void* stream);

/* Host-side view of the GEMM kernel's work distribution (no GPU needed): pieces of tile_m rows x n_cols
 * columns, full 128 x block_n tiles first, then -- when the tile count leaves at most half a wave over
 * num_ctas CTAs (CTA pairs for tile_m = 256) -- the remainder cut into 64-column blocks.  out [cap][3] =
 * (m0, n0, n_cols); returns the number of pieces, negative on error. */
int epos_gemm_pieces(int m_tiles, int N, int block_n, int tile_m, int num_ctas, int32_t* out, int cap);

/* Same contract computed by an fp32 SIMT kernel from f32 operands (validation / tiny shapes):
 * a [M][lda] f32, w [N][K] f32. */
int epos_pwconv_simt(const float* a, int lda, const float* w, const float* bias, int bias_group_rows,
                     const float* residual, int ldr, float* d, int ldd,
                     int M, int N, int K, int relu, void* stream);

/* f32 -> split-bf16 conversion, optional spatial stride-2 subsample (for the stride-2 1x1 shortcut
 * convs, net_xception.py:297-302) and optional ReLU.  x [B,H,W,ldx] (first C channels). */
int epos_split_bf16(const float* x, int ldx, uint16_t* y_split, int ldy, size_t y_plane_stride,
                    int B, int H, int W, int C, int subsample, int relu, void* stream);

/* Global average pool over the spatial axes (model.py:220). x [B,HW,C] -> y [B,C]. */
int epos_global_mean(const float* x, float* y, int B, int HW, int C, void* stream);

/* tf.image.resize_bilinear(align_corners=True) (misc.py:94-107) of x [B,Hi,Wi,C] into channels
 * [0,C) of y [B,Ho,Wo,ldy]. */
int epos_resize_bilinear(const float* x, float* y, int ldy, int B, int Hi, int Wi, int Ho, int Wo, int C,
                         void* stream);

/* Softmax over the last axis, in place (model.py:676-678); rows x n.  If labels != NULL also writes
 * argmax as int64 (model.py:683). */
int epos_softmax_rows(float* x, int64_t* labels, size_t rows, int n, void* stream);

/* Engine path: tf.nn.softmax over the fragment axis (model.py:676-678) of x [pixels][num_objs][num_frags] (in place) only for
 * the (pixel, object) pairs with obj_conf [pixels][num_objs + 1] (softmaxed, channel 0 = background) above min_obj_conf --
 * the pairs establish_many_to_many reads (corresp.py:46-60).  Other rows keep their logits. */
int epos_softmax_rows_masked(float* x, const float* obj_conf, size_t pixels, int num_objs, int num_frags, float min_obj_conf,
                             void* stream);

/* Input side (datagen.py:424-476 _parse_and_preprocess, misc.py:75-91 resize_image_tf, misc.py:110-147 crop_image): a
 * decoded uint8 RGB image src [in_h][src_pitch bytes] (device) is resized so that its height is
 * min(max_height_before_crop, in_h) -- tf.image.resize_area(align_corners=True) when not enlarged, resize_bilinear
 * otherwise -- and cropped at (off_y, off_x) to dst [crop_h][crop_w][3] f32 in [0,255] (the layout model.predict takes).
 * K_in / K_out [9] (HOST, row-major; may be NULL): the intrinsics follow the image, f' = f s, c' = c s - offset with
 * s = new_h / in_h (datagen.py:461-467).  The reference draws the crop offset uniformly (datagen.py:451-455); the
 * caller passes it (for a 640x480 source it is 0).  Image file decoding stays on the host. */
int epos_preprocess_u8(const uint8_t* src, int in_h, int in_w, size_t src_pitch, int max_height_before_crop,
                       int crop_h, int crop_w, int off_y, int off_x, const double* K_in, float* dst, double* K_out,
                       void* stream);

/* ---- correspondences: replaces corresp.establish_many_to_many (epos_lib/corresp.py:9-101) and the
 * top-K selection of scripts/infer.py:425-440 ------------------------------------------------------ */

/* For every (image b, object slot j): pixels with obj_conf[b,p,obj_ids[j]] > min_obj_conf, fragments with
 * frag_conf > min_frag_rel_conf * max_f; emits rows in row-major pixel then fragment order (the reference's order).
 * Inputs are model.predict's maps: obj_conf [B,h,w,O+1], frag_conf [B,h,w,O,F], frag_loc [B,h,w,O,F,3] (f32).
 * Outputs per (b,j) segment of capacity `cap` (segment s = b*J+j starts at row s*cap): coord_2d [cap][2] f64,
 * coord_3d [cap][3] f64, conf / conf_obj / conf_frag [cap] f32, px [cap] i32 (linear output pixel y*w+x),
 * frag [cap] i32; counts[s] = rows written, totals[s] (may be NULL) = rows the reference would emit.
 * If max_corr > 0 and totals[s] > max_corr the segment holds the max_corr most confident rows in descending
 * confidence (ties: descending emission index) = np.argsort(conf)[::-1][:max_corr] (infer.py:431-440);
 * max_corr <= 4096.  Without top-K, rows beyond cap are dropped (counts[s] = cap < totals[s]).
 * frag_centers [num_objs][F][3] f64, frag_sizes [num_objs][F] f64, indexed by obj_id-1 (datagen.py:93-124).
 * obj_ids outside [1, num_objs] give an empty segment. */
int epos_corresp(const float* obj_conf, const float* frag_conf, const float* frag_loc,
                 int B, int h, int w, int num_objs, int num_frags,
                 const int32_t* obj_ids, int J,
                 const double* frag_centers, const double* frag_sizes,
                 double output_scale, float min_obj_conf, float min_frag_rel_conf,
                 int cap, int max_corr,
                 double* coord_2d, double* coord_3d, float* conf, float* conf_obj, float* conf_frag,
                 int32_t* px, int32_t* frag, int32_t* counts, int32_t* totals,
                 void* workspace, size_t workspace_bytes, void* stream);
size_t epos_corresp_workspace_bytes(int B, int J, int h, int w);

/* Same contract with a LAZY localisation head: pred_frag_loc [B,h,w,O,F,3] (model.py:448-456; 2.4 GB per image at
 * O = 30, F = 256) is never materialised.  The three local coordinates of each surviving (pixel, object, fragment) row
 * (at most max_corr per segment) are computed in place from the decoder features -- feat_split [2][B*h*w][ldf] bf16
 * (hi, lo planes `feat_plane_stride` elements apart, feat_channels = 256 used) -- and the 1x1 logit weights
 * w_loc [O*F*3][feat_channels] f32 (row c = (o*F+f)*3+k, the channel order of model.py:133-147) + bias b_loc [O*F*3]
 * (may be NULL), in fp32.  px / frag / coord_2d / conf* are bit-identical to epos_corresp on the materialised maps;
 * coord_3d agrees to fp32 rounding of the 256-term dot product (the dense head computes the same sum on the tensor
 * cores in a different order). */
int epos_corresp_lazy_loc(const float* obj_conf, const float* frag_conf,
                          const uint16_t* feat_split, int ldf, size_t feat_plane_stride, int feat_channels,
                          const float* w_loc, const float* b_loc,
                          int B, int h, int w, int num_objs, int num_frags,
                          const int32_t* obj_ids, int J,
                          const double* frag_centers, const double* frag_sizes,
                          double output_scale, float min_obj_conf, float min_frag_rel_conf,
                          int cap, int max_corr,
                          double* coord_2d, double* coord_3d, float* conf, float* conf_obj, float* conf_frag,
                          int32_t* px, int32_t* frag, int32_t* counts, int32_t* totals,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- pose fitting: replaces pyprogressivex.find6DPoses
 * (external/progressive-x/src/pyprogressivex/src/bindings.cpp:9-118, progressivex_python.cpp:36-336),
 * single-instance branch (GC-RANSAC + final LM), batched over P independent problems. --------------- */
typedef struct {
  double threshold;                  /* inlier threshold in px (infer.py inlier_thresh, 4.0) */
  double spatial_coherence_weight;   /* 0.1 */
  double neighborhood_ball_radius;   /* 20.0 */
  double scaling_from_millimeters;   /* 0.1 */
  double min_triangle_area;          /* 0.0 */
  double min_coverage;               /* 0.5 */
  int32_t max_iters;                 /* 400 */
  int32_t min_iters;                 /* 10  (progressivex_python.cpp:230) */
  int32_t min_iters_before_lo;       /* 20  (settings.h:80) */
  int32_t max_lo_trials;             /* 20  (progressivex_python.cpp:228) */
  int32_t max_graph_cuts;            /* 10  (settings.h:78) */
  int32_t max_lsq_iters;             /* 10  (settings.h:79) */
  int32_t max_unsuccessful;          /* 100 (settings.h:85) */
  int32_t max_neighbors;             /* 5 (<= 8): deterministic stand-in for FLANN checks=6 */
  int32_t apply_numerical_optimization; /* 1 */
  int32_t reserved;
} epos_fit_params;

void epos_fit_params_default(epos_fit_params* p);

/* One record per problem: pose[12] row-major [R|t], then n_inliers, iterations, valid (1 = pose found,
 * 0 = no pose: fewer than 6 correspondences or no model with > 3 inliers, -1 = more than epos_fit_max_points()
 * correspondences), graph_cuts.  The reference returns an uninitialised matrix when nothing is found
 * (model.h:84-89, progressivex_python.cpp:326-335); this library reports valid = 0 instead. */
#define EPOS_POSE_RECORD_DOUBLES 16

/* P problems; problem i owns rows [offsets[i], offsets[i]+counts[i]) of coord_2d [*,2] / coord_3d [*,3]
 * (f64, device).  K [P][9] f64 row-major.  seeds [P] u64: RANSAC stream key (counter-based generator).
 * Outputs: poses [P][16] f64, labeling i32 (1 = inlier) at the same row positions as the inputs.
 * proposal_engine_conf is fixed at 1.0 (scripts/infer.py:90 default), i.e. the iteration bound is max_iters. */
int epos_fit_poses(const double* coord_2d, const double* coord_3d,
                   const int32_t* offsets, const int32_t* counts, int P,
                   const double* K, const uint64_t* seeds, const epos_fit_params* params,
                   double* poses, int32_t* labeling,
                   void* workspace, size_t workspace_bytes, void* stream);
size_t epos_fit_workspace_bytes(int P, int max_points, const epos_fit_params* params);
/* Largest number of correspondences per problem (shared-memory resident point set): 4096. */
int epos_fit_max_points(void);
/* Profiling aid (synchronous): per-problem counters of the last epos_fit_poses on `workspace`.
 * out [P][EPOS_FIT_DEBUG_COLS] i64 (host): N, used_pixels, iterations, passes, graph_cuts, lo_runs, phase, best_inliers,
 * then clock64() totals: main phase sampling+P3P, scoring, replay, whole; cut, trials, final phases; fits inside trials
 * (warp 0); then the number of models scored over all N points in the main loop / the LO trials / the final phase
 * (the "hypotheses" of SURVEY.md 8d's algorithmic-bytes figure: each costs N x 40 B of points + N x 8 B of pixel ids),
 * then the number of points the main loop's early-out did not visit (to be subtracted from models x N). */
#define EPOS_FIT_DEBUG_COLS 20
int epos_fit_debug_state(const void* workspace, int P, long long* out);

/* ---- multi-instance fitting: the Progressive-X branch of find6DPoses (max_model_number in 2 ..
 * max_model_number_for_optimization; progressivex_python.cpp:136-221, progressive_x.h:397-649, PEARL.h:391-536) for P
 * independent problems, one persistent CTA each: proposal (GC-RANSAC with the compound-model score,
 * scoring_function_with_compound_model.h:127-266) -> Tanimoto validation -> PEARL (alpha-expansion with label costs,
 * refits, rejections) -> compound update -> unseen-inlier termination. -------------------------------------------- */
typedef struct {
  int32_t max_model_number_for_pearl;   /* maximum_model_number_to_optimize (infer.py max_model_number_for_pearl, 5; <= 5) */
  int32_t min_point_number;             /* 6: minimum inliers of an instance and PEARL's label cost (infer.py:486) */
  double confidence;                    /* conf = required_progx_confidence (0.5) */
  double max_tanimoto_similarity;       /* 0.9 */
} epos_multi_params;

/* Inputs as epos_fit_poses plus max_models [P] i32 (device): the instance bound of each problem.  2 ..
 * max_model_number_for_pearl: Progressive-X with PEARL.  Larger, or -1 ("all instances", DETECTION): sequential
 * propose-and-remove fitting without PEARL (spedUpFitting, progressive_x.h:265-391; neighbourhood rebuilt over the 7
 * columns of the remaining rows; labeling and scores stay zero as in the reference).  For -1 the reference's loop never
 * terminates (a size_t counter compared with an int holding -1, progressive_x.h:280); here it stops when no model is
 * found, when a proposal has fewer than min_point_number inliers or when the unseen-inlier test of ProgressiveX::run
 * (:589-611) fires on the remaining points; at most epos_fit_max_instances() instances either way.
 * Outputs: multi_counts [P] (instances found; -1 = max_models is 0, 1 or < -1), multi_poses
 * [P][epos_fit_max_instances()][12] row-major [R|t] per instance (no final LM in this branch, as in the reference),
 * multi_scores [P][epos_fit_max_instances()] (sum of the instance's preferences, progressive_x.h:781-790), labeling
 * (instance index per correspondence; outliers = number of instances; with one instance 0 = inlier, 1 = outlier),
 * poses [P][16] = first instance + (points of instance 0, total RANSAC iterations, instances, proposals). */
int epos_fit_poses_multi(const double* coord_2d, const double* coord_3d,
                         const int32_t* offsets, const int32_t* counts, int P,
                         const double* K, const uint64_t* seeds, const epos_fit_params* params,
                         const epos_multi_params* mparams, const int32_t* max_models,
                         double* poses, int32_t* labeling,
                         double* multi_poses, double* multi_scores, int32_t* multi_counts,
                         void* workspace, size_t workspace_bytes, void* stream);
size_t epos_fit_multi_workspace_bytes(int P);
int epos_fit_max_instances(void);
/* Profiling aid (synchronous): out [P][8] i64 = proposals, accepted, RANSAC iterations, PEARL iterations, expansion
 * moves, instances, rejected proposals, 0. */
int epos_fit_multi_debug_state(const void* workspace, int P, long long* out);

/* Debugging aid (synchronous): local-optimisation rounds of problem p of the last epos_fit_poses.
 * out [16][72] i32: graph-cut number, labelled inliers, updated, LO value, LO inliers, then (ok, inliers, pixels) of
 * the 20 inner fits (ok = -1: not evaluated). */
int epos_fit_debug_trace(const void* workspace, int P, int p, int32_t* out);

/* Measurement aid: when enabled, epos_fit_poses records CUDA events on its stream around the set-up kernel and the
 * persistent fitting kernel; epos_fit_last_kernel_ms waits for the last launch and returns both durations (the
 * RANSAC roofline line of bench.py divides SURVEY.md 8d's algorithmic bytes by fit_ms). */
int epos_fit_enable_timing(int on);
int epos_fit_last_kernel_ms(float* prep_ms, float* fit_ms);

#ifdef __cplusplus
}
#endif
#endif /* EPOS_B200_H_ */
