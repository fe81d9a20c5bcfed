/* vpho_b200 -- C ABI of the B200-native VPHO evaluation hot path.
 *
 * Every entry point takes raw device pointers, explicit sizes and a CUDA stream handle (passed as void* so the
 * header needs no CUDA include) and returns an int status: 0 = ok, -1 = invalid argument, -2 = kernel launch
 * failure, -3 = allocation failure.  No exceptions cross this boundary, nothing is allocated per call, inputs are
 * never written.  All calls are stream-ordered and issue no host synchronisation unless stated.
 *
 * The reference (zhoujun-7/VPHO) has no FFI layer: the seam is the set of Python methods wired in
 * `vpho_net.__init__` (lib/model/VPHO.py:56-76).  Each function below names the reference interface it replaces;
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 */
#ifndef VPHO_B200_H_
#define VPHO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vpho_mano_t;      /* packed MANO model on the device          */
typedef void* vpho_denoiser_t;  /* packed score-network weights on the device */
typedef void* vpho_assets_t;    /* force-anchor tables + object point tables  */

int vpho_version(void);

/* Measured peaks of THIS device at the clock a short kernel runs at, in TFLOP/s: back-to-back FP32 FMAs on every SM (the
 * bound of the contact scans) and back-to-back tcgen05 kind::f16 UMMAs M128 N256 K16 with FP32 accumulation on every SM
 * (the bound of the 3xFP16 score network).  Best of `reps` launches, timed with CUDA events on `stream`. */
int vpho_measure_peaks(float* fp32_fma_tflops, float* f16_umma_tflops, int reps, void* stream);

/* Programmatic dependent launch (kernel N+1's prologue overlapping kernel N's tail) is on by default; 0 turns it off so that
 * per-kernel CUDA-event brackets measure isolated durations (bench.py's serialised breakdown pass). */
int vpho_set_pdl(int enabled);

/* Bookkeeping for benchmarks: number of kernels this library has launched since it was loaded, and optional CUDA-event
 * brackets around tagged kernels (recorded on the launching stream).  vpho_profile_collect synchronises on the recorded
 * events of `tag`, returns their summed duration and count, and clears them.
 * tags: 0 hand head-GEMM, 1 object head-GEMM, 2 pose encoder (incl. 6), 3 MANO skinning, 4 physics3 scan, 5 hand heat-map
 * scorer, 6 stage-input + time-term, 7 feat-term, 8 RK error norm / controller / dense output, 9 whole vpho_hoi_aggregate,
 * 10 hand-physics contact scan, 11 trajectory post-processing (6D -> axis-angle) */
unsigned long long vpho_launch_count(void);
int vpho_profile_reserve(int n_events);  /* pre-create the event pool (keeps event creation out of timed regions) */
int vpho_profile_enable(int tag_mask);   /* bit t set: record tag t; 0 = off; -1 = every tag */
int vpho_profile_collect(int tag, double* total_ms, int* n_launches);
/* same, and the individual launch durations (ms) of the first `cap` recorded launches in each_ms (HOST pointer) */
int vpho_profile_collect_list(int tag, double* total_ms, int* n_launches, float* each_ms, int cap);

/* ------------------------------------------------------------------------------------------------ MANO ---- */
/* Packs the MANO tensors (HOST pointers, float32, manopth layouts: v_template [778][3], shapedirs [778][3][10],
 * posedirs [778][3][135], J_regressor [16][778], weights [778][16]) into the device layout.
 * Replaces the `ManoLayer(...)` construction at lib/model/head_mano.py:48-55. */
int vpho_mano_create(const float* v_template, const float* shapedirs, const float* posedirs,
                     const float* J_regressor, const float* weights, vpho_mano_t* out);
int vpho_mano_destroy(vpho_mano_t h);

/* verts [n][778][3] (may be NULL: joints only), joints [n][21][3]; pose [n][48] axis-angle, shape [n][10];
 * metres, wrist-centred.  Replaces `HeadMano.get_hand_verts` (lib/model/head_mano.py:78-87). */
int vpho_mano_forward(vpho_mano_t h, const float* pose, const float* shape, int n, float* verts, float* joints,
                      void* stream);
/* Same with an explicit numerical path: flags = 0 is vpho_mano_forward (tcgen05 blend when vertices are materialised);
 * VPHO_MANO_STRICT_FP32 runs the FP32 SIMT kernel (cross-check of the tensor-core kernel; also what joints-only calls use). */
#define VPHO_MANO_STRICT_FP32 1
#define VPHO_MANO_DEBUG_BLEND 2   /* diagnostics: verts receives the blended rest pose (template + shape + pose correctives) */
int vpho_mano_forward_ex(vpho_mano_t h, const float* pose, const float* shape, int n, float* verts, float* joints, int flags,
                         void* stream);

/* --------------------------------------------------------------------------------------------- sampler ---- */
/* Packs a BaseDenoiser state dict (HOST pointers, float32, reference layouts -- lib/model/denoiser.py:33-66,
 * lib/model/parallel_linear.py:15-16):
 *   fourier_W [64]; t_w [128][128], t_b [128]; p1_w [256][D], p1_b [256]; p2_w [256][256], p2_b [256];
 *   ha_w [n][1408][256], ha_b [n][256]; hb_w [n][256][3], hb_b [n][3];  D = 3n (96 hand / 9 object). */
int vpho_denoiser_create(int n_heads, const float* fourier_W, const float* t_w, const float* t_b, const float* p1_w,
                         const float* p1_b, const float* p2_w, const float* p2_b, const float* ha_w,
                         const float* ha_b, const float* hb_w, const float* hb_b, vpho_denoiser_t* out);
/* Same with an explicit numerical path, fixed for the life of the handle:
 *   flags = 0                          tcgen05 tensor-core path (3xTF32 / 3xFP16 splitting, FP32-class accuracy).  Every
 *                                      operand plane and TMA descriptor it needs is REQUIRED: if one cannot be built the
 *                                      call returns VPHO_ERR_ALLOC / VPHO_ERR_LAUNCH -- it never degrades to another path.
 *   flags = VPHO_DENOISER_STRICT_FP32  FP32 SIMT kernels only (cross-check of the tensor-core kernels). */
#define VPHO_DENOISER_STRICT_FP32 1
int vpho_denoiser_create_ex(int n_heads, const float* fourier_W, const float* t_w, const float* t_b, const float* p1_w,
                            const float* p1_b, const float* p2_w, const float* p2_b, const float* ha_w,
                            const float* ha_b, const float* hb_w, const float* hb_b, int flags, vpho_denoiser_t* out);
int vpho_denoiser_destroy(vpho_denoiser_t h);

/* One score-network evaluation: out[N][D] = denoiser(x[N][D], t, feat) / (sigma(t)+1e-7), t a single float shared
 * by all rows (as on the sampling path).  feat is [N/rows_per_feat][1024]; row r uses feat[r / rows_per_feat].
 * workspace: vpho_sample_workspace_bytes(...) bytes.  Replaces `BaseDenoiser.forward`
 * (lib/model/denoiser.py:68-82) for the sampling call pattern. */
int vpho_score_eval(vpho_denoiser_t h, const float* x, float t, const float* feat, int n_rows, int rows_per_feat,
                    float* out, void* workspace, size_t workspace_bytes, void* stream);

size_t vpho_sample_workspace_bytes(int n_heads, int n_rows, int rows_per_feat, int n_eval);

/* Probability-flow ODE sampler with SciPy's adaptive RK45 controller run entirely on the device.
 * Replaces `ScoreBasedModelAgent.sample` -> `cond_ode_sampler` (lib/model/score_based_model.py:45-105,130-146).
 *   init_x   [N][D] f32  prior draw randn*sigma(T0) (lib/model/sde.py:26-28) -- passed in so both sides share noise
 *   t_eval   [n_eval] f64 DEVICE pointer, or NULL for numpy.linspace(T0, eps, n_eval) computed on the device
 *   num_steps            the reference's `num_steps` (scale of the final predictor step; = n_eval on the eval path)
 *   xs       [n_eval][N][D] f64 (may be NULL), x [N][D] f64 (written by vpho_sample_finish)
 *   counters [8] int32 device (may be NULL): {status, nfev, accepted, rejected, nan_seen, attempts, -, -}
 *            status 1 = finished, 0 = needs more attempts (vpho_sample_continue), -1 = step size too small.
 * `max_attempts` RK step attempts (6 network calls each) are enqueued; attempts after the integration has
 * finished are skipped on the device.  No host synchronisation. */
int vpho_sample_begin(vpho_denoiser_t h, const float* feat, int n_rows, int rows_per_feat, const float* init_x,
                      double T0, double eps, const double* t_eval, int n_eval, double rtol, double atol,
                      double max_step, int num_steps, int max_attempts, double* xs, double* x, int32_t* counters,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Enqueues `max_attempts` more attempts on the same workspace (same n_rows / rows_per_feat / n_eval as begin). */
int vpho_sample_continue(vpho_denoiser_t h, int n_rows, int rows_per_feat, int n_eval, int max_attempts,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Final "denoise" predictor step (score_based_model.py:95-104) -> x.  A no-op on the device while status != 1,
 * so it can be enqueued right behind begin/continue and repeated after a continue. */
int vpho_sample_finish(vpho_denoiser_t h, int n_rows, int rows_per_feat, int n_eval, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Two samplers in lock-step.  `vpho_net.forward(mode='predict')` integrates the hand ODE and the object ODE of the same
 * batch back to back (lib/model/VPHO.py:239-262); both issue the same sequence of network calls (2 + 6 per RK attempt
 * + 1), so the pair entry points serve both integrations with ONE launch per kernel of that sequence -- the object's work
 * fills the SMs the hand's leaves idle instead of competing for them from another stream.  Each sampler keeps its own
 * controller, workspace and outputs, and its results are bit-identical to vpho_sample_begin/continue/finish on its own
 * arguments.  The structure holds exactly the per-sampler arguments of vpho_sample_begin. */
typedef struct {
  vpho_denoiser_t denoiser;
  const float* feat;       /* [ceil(n_rows / rows_per_feat)][1024] */
  int n_rows, rows_per_feat;
  const float* init_x;     /* [n_rows][D] prior draw */
  double T0, eps;
  const double* t_eval;    /* device, or NULL for linspace(T0, eps, n_eval) */
  int n_eval;
  double rtol, atol, max_step;
  int num_steps;
  double* xs;              /* [n_eval][n_rows][D] or NULL */
  double* x;               /* [n_rows][D] */
  int32_t* counters;       /* [8] status word */
  void* workspace;
  size_t workspace_bytes;
  float* xs_f32;           /* optional [n_eval][n_rows][D]: the dense output rounded to float32, i.e. what the predict branch keeps
                            * of it (`.float()`, lib/model/VPHO.py:243); `xs` may then be NULL and no float64 copy is written */
} vpho_sample_args;
int vpho_sample_pair_begin(const vpho_sample_args* a, const vpho_sample_args* b, int max_attempts, void* stream);
int vpho_sample_pair_continue(const vpho_sample_args* a, const vpho_sample_args* b, int max_attempts, void* stream);
int vpho_sample_pair_finish(const vpho_sample_args* a, const vpho_sample_args* b, void* stream);

/* 6D -> axis-angle for the hand finals (+ regressed shape): `vpho_net.postprocess_diffusion_hand`
 * branch 'mano_pose' (lib/model/VPHO.py:318-326).  x6d [n][16][6] f32 -> pose_aa [n][48] f32. */
int vpho_rot6d_to_axis_angle(const float* x6d, int n_rot, float* aa, void* stream);

/* The whole of `vpho_net.postprocess_diffusion_hand`, branch 'mano_pose' (lib/model/VPHO.py:306-331) in one pass over the
 * sampler's float64 output: xs [n_steps][n_rows][96] f64 (storage order of vpho_sample_begin's `xs`; n_steps = 1 for the
 * final state `x`) -> out [n_rows][n_steps][58] f32 = 48 axis-angle parameters (after the float64 -> float32 rounding of
 * `.float()`) + the 10 regressed shape coefficients of the row's image (shape [n_rows / rows_per_shape][10]). */
int vpho_postprocess_hand(const double* xs, int n_steps, int n_rows, int rows_per_shape, const float* shape, float* out,
                          void* stream);
/* Same for a float32 trajectory (vpho_sample_args.xs_f32). */
int vpho_postprocess_hand_f32(const float* xs, int n_steps, int n_rows, int rows_per_shape, const float* shape, float* out,
                              void* stream);

/* ---------------------------------------------------------------------------------- assets / aggregation ---- */
/* Force-anchor tables (lib/utils/physics_fn.py:121-257, lib/utils/hand_fn.py:427-448) and per-object point tables
 * (lib/model/head_object.py:9-34).  HOST pointers: face_vertex_idx [32][3] int32, anchor_weight [32][2] f32,
 * vert2joint [21][778] f32, kpt3d [n_obj][27][3], verts [n_obj][n_pts][3], com [n_obj][3]. */
int vpho_assets_create(const int32_t* face_vertex_idx, const float* anchor_weight, const float* vert2joint,
                       int n_obj, int n_pts, const float* kpt3d, const float* verts, const float* com,
                       vpho_assets_t* out);
int vpho_assets_destroy(vpho_assets_t h);

/* `HeadObject.forward` + `flip_pt3d` (lib/model/head_object.py:36-67): out[b][c][v][3] = R(pose6d[b][c]) p_v + t,
 * x negated where !is_right.  which: 0 keypoints(27) 1 sampled verts 2 CoM.  pose6d f32 [bs][C][9]. */
int vpho_object_points(vpho_assets_t h, const float* pose6d, const int32_t* obj_id, const uint8_t* is_right, int bs,
                       int C, int which, int flip, float* out, void* stream);

/* `from_local_to_global` (lib/model/physics.py:362-371) over `VERT2ANCHOR` : verts [n][778][3] (camera frame),
 * force_local [n/group][32][3] -> force_point [n][32][3], force_global [n][32][3]. */
int vpho_force_anchors(vpho_assets_t h, const float* verts, const float* force_local, int n, int group,
                       float* force_point, float* force_global, void* stream);

/* Contact scoring of n posed hands against one object point cloud per `group` consecutive hands (BASELINE config 3).
 * vpho_anchor_contact: nearest distance of each of the 32 force anchors to obj_points [n/group][n_pts][3] (exact
 * Euclidean, `nn_for_r_memory_save2` lib/model/aggregation.py:1145-1158) -> dist [n][32] (may be NULL) and the per-finger
 * physics scores -(w * d * |sum f_hat|) summed over the finger's anchors (`select_by_physics`, aggregation.py:553-590)
 * -> finger_score [n][5] (may be NULL).
 * vpho_vertex_contact: nearest distance of every MANO vertex, verts [n][778][3] -> dist [n][778]; a dense stress variant
 * with no counterpart in the reference. */
int vpho_anchor_contact(const float* force_point, const float* force_global, const float* obj_points, int n, int group,
                        int n_pts, float* dist, float* finger_score, void* stream);
int vpho_vertex_contact(const float* verts, const float* obj_points, int n, int group, int n_pts, float* dist, void* stream);

/* Pseudo-force evaluation (BASELINE config 5): the forward math of one iteration of `ForceOptimizer.optimize_batch`
 * (lib/engine/force_optimization.py:141-171) for n posed hands.  verts [n][778][3] camera frame; scale [n][32];
 * weight [n][32][8] (pre-softmax); contact_mask [n][32] u8 or NULL; force_contact [n][32] or NULL; cone_anchor [8][3] the
 * `HeadPhysics.anchor` buffer with its xy already scaled by the friction coefficient (lib/model/physics.py:549-550,692-698);
 * gravity, com [n/group][3].  terms [n][4] = {|sum f + g|, (sum f).(-g), |sum (p - CoM) x f|, mean_j (log|c_j/s_j| mask_j)^2};
 * optional outputs force_local / force_point / force_global [n][32][3]. */
int vpho_force_eval(vpho_assets_t h, const float* verts, const float* scale, const float* weight, const uint8_t* contact_mask,
                    const float* force_contact, const float* cone_anchor, const float* gravity, const float* com, int n, int group,
                    float* terms, float* force_local, float* force_point, float* force_global, void* stream);

/* `ForceOptimizer.optimize_batch` for one batch (lib/engine/force_optimization.py:110-207): n_iter iterations (reference 3000)
 * of the two AdamW optimisers (:33-37; lr, betas (0.9, 0.999), eps 1e-8, weight decay 0.01) over scale [n][32] (initialised
 * 0.05) and weight [n][32][8] (0): gravity loss on the weights for the first switch_iter iterations (reference 300), then
 * force + moment + contact-distribution loss on both (:150-176), with an analytic backward pass, in ONE persistent kernel.
 * verts [n][778][3], force_contact [n][32], gravity / com [n][3] in the flipped frame (:134-137); cone_anchor [8][3] as
 * for vpho_force_eval.  Out: the parameters after the last step; force_local / force_global [n][32][3] of the last
 * iteration's forward pass, zeroed for hands with is_grasped == 0 (:191-194; NULL: none); losses [n_iter][5] =
 * {loss, force, gravity, moment, dist} (optional).  The optimiser state starts fresh (the reference's first batch). */
size_t vpho_force_optimize_workspace_bytes(int n);
int vpho_force_optimize(vpho_assets_t h, const float* verts, const float* force_contact, const float* gravity, const float* com,
                        const uint8_t* is_grasped, const float* cone_anchor, int n, int n_iter, int switch_iter, float lr,
                        float* scale, float* weight, float* force_local, float* force_global, float* losses, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Final pose error per image, in millimetres: metrics [n][4] = {MJE, MVE, ADD, ADD-S}.  MJE / MVE: mean Euclidean
 * joint / vertex distance of `TesterHand` (lib/engine/test.py:657-679); ADD / ADD-S of
 * `TesterObject.criterion_ADD_REP` (lib/engine/test.py:413-442) on the object's sampled vertices posed by the predicted /
 * ground-truth 6D poses (rot6d + translation, float64 [n][9]). */
int vpho_pose_metrics(vpho_assets_t h, const float* pd_joint, const float* gt_joint, const float* pd_vert,
                      const float* gt_vert, const double* pd_obj6d, const double* gt_obj6d, const int32_t* obj_id, int n,
                      float* metrics, void* stream);

/* The whole per-image row of `TesterHand.__call__` (lib/engine/test.py:589-654), in millimetres:
 * metrics [n][25] = {MJE, PA_MJE, MVE, PAMVE, JE[21]} for one prediction per image. */
int vpho_hand_metrics(const float* pd_joint, const float* gt_joint, const float* pd_vert, const float* gt_vert, int n,
                      float* metrics, void* stream);

/* Object-pose metrics of `TesterObject` (lib/engine/test.py:196-584) per (image, candidate), on the device -- what
 * Trainer.evaluate computes from `diff_final_obj_rt` / `agg_obj_rt` on the host (lib/engine/train_diff_hand_obj.py:249-257).
 * Tables (HOST pointers): bbox3d [n_obj][8][3] and diameter [n_obj] (YCB_MESHES[name]['bbox3d' / 'diameter']), the padded
 * symmetry transforms of TesterObject.__init__ (test.py:208-232) sym_R [n_obj][sym_k][3][3], sym_t [n_obj][sym_k][3] in
 * metres, sym_count [n_obj] (may be NULL), and the cloud of the F-score / Chamfer terms fverts [n_obj][n_fpts][3]
 * (NULL: the assets' sampled surface, as `verts_sampled`).
 * vpho_object_metrics: pd_rt [n][C][3][4], gt_rt [n][3][4] float64 ([R | t] rows), obj_id [n], cam_intr [n][3][3] ->
 * out [n][C][VPHO_OBJ_METRIC_COLS] float64 = MCE, OCE, MCE2, SMCE, ADD, ADD-S (m), REP (px), CD (m),
 * FSCORE@{2,5,10 mm, 2,5,10 cm}, ADD01d, ADDS01d, REP5 (0/1).  Replaces criterion_MCE_OCE :354-375, criterion_SMCE
 * :377-399, criterion_MCE2 :401-417, criterion_ADD_REP :419-451, criterion_FSCORE :453-503, cal_ADD01d / cal_REP5 :505-521. */
#define VPHO_OBJ_METRIC_COLS 17
typedef void* vpho_objmetrics_t;
int vpho_objmetrics_create(int n_obj, const float* bbox3d, const float* diameter, int sym_k, const double* sym_R,
                           const double* sym_t, const int32_t* sym_count, int n_fpts, const float* fverts,
                           vpho_objmetrics_t* out);
int vpho_objmetrics_destroy(vpho_objmetrics_t h);
int vpho_object_metrics(vpho_assets_t assets, vpho_objmetrics_t tables, const double* pd_rt, const double* gt_rt,
                        const int32_t* obj_id, const float* cam_intr, int n, int C, double* out, void* stream);

/* The metric step of Trainer.evaluate for one batch (lib/engine/train_diff_hand_obj.py:224-269) in one call: the
 * post-processing of :578-602 (un-flip left hands and add the root joint; rot6d + translation -> [R | t + root]),
 * `TesterHand` on the aggregated hand, the first diffusion candidate and (when reg_hand_* are given) the regression hand,
 * `TesterObject` on the aggregated and the first candidate object pose.  All pointers are DEVICE pointers.
 * out [bs][n_sets * 25 + 2 * VPHO_OBJ_METRIC_COLS] float64, n_sets = 2 (3 with the regression hand): per hand set
 * {MJE, PA_MJE, MVE, PAMVE, JE[21]} (mm), then per object set the columns of vpho_object_metrics.  This row is the
 * payload that replaces the pickled dicts of `gather_for_metrics(use_gather_object=True)` (:333-357). */
typedef struct {
  int bs;                         /* images                                                             */
  int S;                          /* candidates per image in the cand_* arrays (row 0 of each image is read) */
  const float* agg_hand_joint;    /* [bs][21][3]      wrist-relative, flipped frame (vpho_hoi_aggregate)  */
  const float* agg_hand_vert;     /* [bs][778][3]                                                       */
  const float* cand_hand_joint;   /* [bs][S][21][3]   diff_final_hand_joint                             */
  const float* cand_hand_vert;    /* [bs][S][778][3]  diff_final_hand_vert                              */
  const float* reg_hand_joint;    /* [bs][21][3] or NULL                                                */
  const float* reg_hand_vert;     /* [bs][778][3] or NULL                                               */
  const double* agg_obj_6d;       /* [bs][9]     rot6d + root-relative translation                      */
  const double* cand_obj_6d;      /* [bs][S][9]                                                         */
  const float* root_joint;        /* [bs][3]                                                            */
  const uint8_t* is_right;        /* [bs]                                                               */
  const float* gt_joint;          /* [bs][21][3]  camera frame                                          */
  const float* gt_vert;           /* [bs][778][3]                                                       */
  const double* gt_obj_rt;        /* [bs][3][4]                                                         */
  const float* cam_intr;          /* [bs][3][3]                                                         */
  const int32_t* obj_id;          /* [bs]                                                               */
  double* out;                    /* [bs][n_sets * 25 + 34]                                             */
} vpho_eval_record_args;
size_t vpho_eval_record_workspace_bytes(int bs);
int vpho_eval_record(vpho_assets_t assets, vpho_objmetrics_t tables, const vpho_eval_record_args* args, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Procrustes-aligned hand errors per image, in millimetres: metrics [n][23] = {PA-MJE, PA-MVE, JE[21]} of
 * `TesterHand.criterion_MJE_PAMJE` (lib/engine/test.py:657-679): the prediction is aligned to the ground truth by the
 * similarity transform of `rigid_align_AtoB` (lib/utils/transform_fn.py:43-66; SVD of the 3x3 cross-covariance, reflection
 * fix, scale = sum(s) / var) before the mean joint / vertex distance is taken; JE are the unaligned per-joint errors. */
int vpho_hand_pa_metrics(const float* pd_joint, const float* gt_joint, const float* pd_vert, const float* gt_vert, int n,
                         float* metrics, void* stream);

typedef struct {
  int bs;          /* images in this batch                         */
  int S;           /* sample_num: diffusion candidates per image   */
  int topk_hand;   /* <= 64                                        */
  int topk_obj;    /* <= 16                                        */
  int phy_topk;    /* hard-coded 5 in the reference (aggregation.py:1246) */
  /* inputs (device) */
  const float* cam_intrinsic;    /* [bs][3][3]   cam_intr_crop_flip */
  const float* root_joint_flip;  /* [bs][3] */
  const float* root_joint;       /* [bs][3] */
  const uint8_t* is_right;       /* [bs] */
  const uint8_t* is_grasped;     /* [bs] */
  const float* force_local;      /* [bs][32][3] */
  const float* hand_pose_diff;   /* [bs*S][48] */
  const float* hand_pose_reg;    /* [bs][48] */
  const float* hand_shape;       /* [bs*S][10] */
  const float* hand_heatmap;     /* [bs][21][64][64] */
  const float* hand_bbox;        /* [bs][4] */
  const double* obj_pose6d;      /* [bs][S][9] float64 */
  const float* obj_heatmap;      /* [bs][27][64][64] */
  const float* obj_bbox;         /* [bs][4] */
  const int32_t* obj_id;         /* [bs] index into the object tables */
  /* outputs (device) -- keys of the dict returned by HOI_Aggregator.__call__ (aggregation.py:1339-1348) */
  double* obj_agg_6d;            /* [bs][9] f64 */
  double* pose6d_candidate;      /* [bs][topk_obj^2][9] f64 */
  float* agg_obj_vert;           /* [bs][n_pts][3] */
  float* hand_agg_mano;          /* [bs][58] */
  float* hand_agg_vert;          /* [bs][778][3] */
  float* hand_agg_joint;         /* [bs][21][3] */
  /* optional diagnostics (may be NULL): scores / indices of every selection, for stage-wise parity tests */
  float* dbg_hand_score;         /* [4][bs][2S][5]  (level 0 uses [..][0]) */
  int32_t* dbg_hand_topk;        /* [4][bs][5][topk_hand] */
  float* dbg_cascade_pose;       /* [bs][48] fused pose after the cascade */
  float* dbg_obj_score;          /* [4][bs*max(S,topk_obj^2)]: transl heat, rot heat (each [bs][S]), physics3, final heat
                                    (each [bs][topk_obj^2]), packed at the start of their block */
  int32_t* dbg_obj_topk;         /* [4][bs][max(topk_obj,phy_topk)] */
  float* dbg_finger_score;       /* [bs][5][topk_hand+1] */
  int32_t* dbg_finger_topk;      /* [bs][5][phy_topk] */
  float* dbg_force_point;        /* [bs][32][3] */
  float* dbg_force_global;       /* [bs][32][3] */
} vpho_hoi_args;

size_t vpho_hoi_workspace_bytes(int bs, int S, int topk_hand, int topk_obj, int n_pts);

/* Whole `HOI_Aggregator.__call__` (lib/model/aggregation.py:1167-1353): hand heat-map cascade, object
 * heat-map / physics selection, hand physics refinement. */
int vpho_hoi_aggregate(vpho_mano_t mano, vpho_assets_t assets, const vpho_hoi_args* args, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------------
 * N4: the aggregation modes the predict branch does not use -- `HandAggregator.__call__` modes 'heatmap',
 * 'heatmap_cascade_n_level', '2D_pt_pose', '2D_pt_joint', 'average_all', 'random' (lib/model/aggregation.py:63-113,286-535)
 * and `ObjectAggregator` 'heatmap' / the non-physics branch of 'heatmap_cascade' / '2D_pt_pose' / 'average_all' / 'random'
 * (:646-722, 1001-1113) -- as device primitives; the host
 * mirror vpho_b200/aggregation_modes.py composes them the way the reference's methods do. */
typedef struct {
  int bs, n, n_joints;            /* images, candidates per image, joints per candidate (21)                         */
  const float* joint;             /* [bs][n][n_joints][3] wrist-relative MANO joints of the candidates                  */
  const float* root_joint;        /* [bs][3]                                                                            */
  const float* cam_intrinsic;     /* [bs][3][3]                                                                         */
  const float* bbox;              /* [bs][4]                                                                            */
  const float* heatmap;           /* [bs][n_joints][64][64]                                                             */
  float* heat;                    /* out or NULL [bs][n][n_joints]: bicubic grid_sample of joint j's map at its projection
                                     (aggregation.py:196-210)                                                           */
  float* dist2d;                  /* out or NULL [bs][n][n_joints]: -||projection - argmax position|| (:313-326)        */
} vpho_joint_scores_args;
/* workspace: bs * n_joints * 2 floats when dist2d is requested (the maps' peak positions), else may be NULL */
int vpho_joint_scores(const vpho_joint_scores_args* args, void* workspace, size_t workspace_bytes, void* stream);

/* `select_topk_hand_by_observed_heatmap_and_fuse_by_index` (lib/model/aggregation.py:180-284) for any index sets. */
typedef struct {
  int bs, n, K, n_joints;
  const float* score;             /* [bs][n][n_joints] per-joint scores (vpho_joint_scores heat or dist2d)              */
  float* pose;                    /* [bs][n][48] axis-angle candidates; fuse_kind 0; overwritten at fuse_index when
                                     write_back (fused_pose[:, :, fuse_index] = ... :235-236)                           */
  const float* joint;             /* [bs][n][n_joints][3]; fuse_kind 1 only                                             */
  const int32_t* observe_index;   /* HOST [n_observe] joints whose scores enter                                         */
  int n_observe;
  const int32_t* fuse_index;      /* HOST [n_fuse] pose parameters to fuse, whole joints (3j, 3j+1, 3j+2)               */
  int n_fuse;
  int independent;                /* 0: one list (sum over observe_index); 1: one list per fused joint (mean, :242-245) */
  int is_weight;                  /* heat-value weights or a plain average                                              */
  int fuse_kind;                  /* 0: quaternion average of the winners' parameters; 1: mean of the winners' joint
                                     positions, joint fuse_index[3l]/3 of list l ('2D_pt_joint' :357-362)               */
  int write_back;
  float* val;                     /* out or NULL [bs][K][lists]                                                         */
  int32_t* topk;                  /* out or NULL [bs][K][lists]                                                         */
  float* fused;                   /* out [bs][n_fuse] (fuse_kind 0) or [bs][lists][3] (fuse_kind 1)                     */
} vpho_hand_level_args;
int vpho_hand_level(const vpho_hand_level_args* args, void* stream);

/* `average_all` (:400-404): unweighted quaternion average of every joint over all n candidates; pose [bs][n][n_joints][3]
 * axis-angle -> out [bs][n_joints][3] */
int vpho_quat_average_all(const float* pose, int bs, int n, int n_joints, float* out, void* stream);

/* `ObjectAggregator.select_topk_object_by_heatmap` + `fuse_topk` (:729-781) */
typedef struct {
  int bs, n, K;
  const double* pose6d;           /* [bs][n][9] rot6d + root-relative translation                                       */
  const float* root_joint;        /* [bs][3]                                                                            */
  const float* cam_intrinsic;     /* [bs][3][3]                                                                         */
  const float* bbox;              /* [bs][4]                                                                            */
  const float* heatmap;           /* [bs][27][64][64]                                                                   */
  const uint8_t* is_right;        /* [bs]                                                                               */
  const int32_t* obj_id;          /* [bs]                                                                               */
  int is_weight;                  /* fuse with the heat weights or with a plain mean                                    */
  int score_kind;                 /* 0: summed heat values (:752-776); 1: minus the summed 2D distances between the projected
                                     key-points and their maps' argmax positions ('2D_pt_pose', :1013-1036; plain mean)   */
  const int32_t* topk_in;         /* NULL, or [bs][K] winners chosen by an earlier call: only fuse_topk runs (plain mean) --
                                     the reference's non-physics cascade fuses one selection on another pose set (:713-715) */
  int32_t* topk;                  /* out or NULL [bs][K]                                                                */
  float* weight;                  /* out or NULL [bs][K]                                                                */
  double* fused;                  /* out or NULL [bs][9]                                                                */
} vpho_obj_select_args;
/* workspace: bs * n floats (the candidates' scores) + bs * 27 * 2 floats (score_kind 1: the maps' peak positions) */
int vpho_obj_select(vpho_assets_t assets, const vpho_obj_select_args* args, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------------
 * N1: the modules that PRODUCE the hot path's inputs, from the RoI-aligned feature maps to the encodings,
 * heat-maps, regression pose and local contact forces -- `vpho_net.forward` lib/model/VPHO.py:129-178:
 *   head_hm_hand / head_hm_obj   HeadHeatmap2.forward              lib/model/head_inplane.py:42-107
 *   align_hm_to_bbox_rectangle, flip_tensor_by_mask_index, F.interpolate   lib/model/VPHO.py:136-148,333-357
 *   encoder_hand / encoder_obj   Encoder.forward, Residual.forward lib/model/encoding.py:5-73
 *   head_mano                    HeadMano.forward                  lib/model/head_mano.py:61-76
 *   cross_hand / cross_obj       CrossModule.forward               lib/model/cross_module.py:91-137
 *   head_physics                 HeadPhysics.forward               lib/model/physics.py:648-721 (+ get_local_force :546-557)
 * eval mode (BatchNorm running statistics, no dropout).  Weights are handed over as the module tree's state dict: a table of
 * named float32 HOST tensors whose names are the reference's state-dict keys (`head_hm_hand.conv_layers.0.weight`, ...,
 * as `accel.save_state` writes them, lib/engine/base_trainer.py:85-89).  Every dimension (feature channels, RoI size, hidden
 * widths, joint / key-point counts, d_model, feed-forward width) is read from the tensors' shapes. */
typedef struct vpho_named_tensor {
  const char* name;        /* reference state-dict key, without any `module.` wrapper prefix */
  const float* data;       /* host pointer, float32, C-contiguous */
  int32_t ndim;            /* <= 4 */
  int64_t shape[4];
} vpho_named_tensor;
typedef struct vpho_heads* vpho_heads_t;
/* VPHO_ERR_INVALID when a key is missing or a shape is inconsistent with the module definitions. */
int vpho_heads_create(const vpho_named_tensor* tensors, int n_tensors, vpho_heads_t* out);
int vpho_heads_destroy(vpho_heads_t h);
size_t vpho_heads_workspace_bytes(vpho_heads_t h, int bs, int roi_size);

typedef struct vpho_heads_args {
  int32_t bs, roi_size;            /* images; RoI side (cfg.roi_size = 32); heat-maps are 2*roi_size */
  /* inputs (device, f32 unless noted) */
  const float* hf_hr;              /* [bs][C][roi][roi]  hand features, tight hand box   (VPHO.py:123) */
  const float* of_or_rect;         /* [bs][C][roi][roi]  object features, square box     (VPHO.py:126) */
  const float* hf_hr_rect;         /* [bs][C][roi][roi]  hand features, square hand box  (VPHO.py:125) */
  const float* bbox_hand;          /* [bs][4] */
  const float* bbox_hand_rect;     /* [bs][4] */
  const float* bbox_obj;           /* [bs][4] */
  const float* bbox_obj_rect;      /* [bs][4] */
  const uint8_t* is_right;         /* [bs] bool */
  const float* gravity;            /* [bs][3] (data['gravity'], un-flipped) */
  /* outputs (device) */
  float* hand_heatmap;             /* [bs][Jh][2roi][2roi]  pd_hm_hand */
  float* obj_heatmap;              /* [bs][Jo][2roi][2roi]  pd_hm_obj */
  float* encoding_hand;            /* [bs][enc_dim] */
  float* encoding_obj;             /* [bs][enc_dim] */
  float* mano_pose;                /* [bs][48] axis-angle  pd_mano_pose */
  float* mano_shape;               /* [bs][10] */
  float* force_local;              /* [bs][32][3] */
  float* force_scale;              /* [bs][32]    (optional) */
  float* force_weight;             /* [bs][32][8] (optional; after fc_weight's Softmax) */
  float* CoM;                      /* [bs][32][3] (optional) */
  float* enc_phy_hand;             /* [bs][32][d_model] (optional diagnostics) */
  float* enc_phy_obj;              /* [bs][32][d_model] (optional diagnostics) */
  int32_t flags;                   /* 0: tcgen05 path (FP16 hi/lo operand planes, 3 UMMAs per product, ~2^-22 relative);
                                    * VPHO_HEADS_STRICT_FP32: FP32 SIMT path (cross-check; the only one of the emulator build) */
} vpho_heads_args;
#define VPHO_HEADS_STRICT_FP32 1
int vpho_heads_forward(vpho_heads_t h, const vpho_heads_args* args, void* workspace, size_t workspace_bytes, void* stream);
/* dims[8] = {C, Jh, Jo, enc_dim, d_model, n_force(32), heat hidden, encoder hidden} */
int vpho_heads_dims(vpho_heads_t h, int32_t* dims);
/* *flag = 1 when an activation of the last tensor-core forward left the FP16 range of the operand planes (results invalid:
 * re-run with VPHO_HEADS_STRICT_FP32).  Synchronises `stream`. */
int vpho_heads_overflow(vpho_heads_t h, int32_t* flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VPHO_B200_H_ */
