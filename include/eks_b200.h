/* eks_b200.h -- C ABI of libeks_b200.so: the B200-native (sm_100a) EKS smoothing hot path.
 *
 * The reference (paninski-lab/eks v4.6.2) has no native/FFI layer: its seam is the Python call
 * signatures of eks/core.py.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless its name ends in _host or the comment says "host".
 *  - dtype arguments take EKS_F32 / EKS_F64 ("real" below = that type).  Production precision of the
 *    reference is float32; float64 is the parity mode.
 *  - The caller owns every buffer; the library never allocates or frees caller memory.  Scratch is
 *    passed as (workspace, workspace_bytes); sizes come from the *_workspace_bytes queries.
 *  - Calls ENQUEUE work on `stream` (a cudaStream_t cast to void*) and return; the caller synchronises.
 *    Exceptions, stated at the entry point: eks_pupil_optimize, eks_filter_smooth (sequences of >= 512 frames) and
 *    eks_optimize_s / eks_nll_grad for generic models with >= 512 frames (everything except EKS_STRUCT_DIAG*) read a
 *    completion / verification flag and therefore synchronise `stream` themselves.
 *  - State kept by the library: a thread-local error string and two counters (eks_last_error,
 *    eks_last_launch_count, eks_last_unverified_count), and -- only for EKS_STRUCT_DIAG_STREAM -- a per-device table of two internal streams,
 *    created on first use under a mutex and forked from / joined to the caller's stream by events.  Nothing else
 *    persists between calls.  Entry points may be called concurrently from several host threads (different streams).
 *  - Return value: 0 ok; <0 invalid argument; >0 a cudaError_t.  eks_last_error() returns a
 *    thread-local message for the last non-zero return.
 *  - Per-frame data are "channel planes": element (sequence b, channel o, frame t) of a view
 *    (base, seq_stride, chan_off[O]) lives at base[b*seq_stride + chan_off[o] + t], i.e. frames are
 *    contiguous (frame-major), strides/offsets are in ELEMENTS.  chan_off is a HOST array.
 */
#ifndef EKS_B200_H
#define EKS_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define EKS_F32 0
#define EKS_F64 1
#define EKS_MAX_CHAN 16    /* max observation channels (2 x cameras) */
#define EKS_MAX_STATE 6    /* max latent dimension */
#define EKS_STRUCT_GENERAL 0 /* model_structure: no assumption */
#define EKS_STRUCT_DIAG 1    /* caller asserts D == O == 2 and diagonal A, C, Q, S0 (singlecam model) */
#define EKS_STRUCT_DIAG_STREAM 2 /* same model; force one streaming pass over the observations per evaluation */
#define EKS_CAM_STRIDE 29  /* R(9 row-major) t(3) fx fy cx cy skew k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 */

const char* eks_last_error(void);
int eks_version(void);   /* 202 for this header; the Python binding refuses any other value */
/* Number of kernels the most recent eks_optimize_s / eks_diag_smooth / eks_const_R_median call of this thread
 * enqueued (bench.py's gpu_launches). */
int eks_last_launch_count(void);
/* Evaluations of the last run-parallel eks_optimize_s / eks_nll_grad call on this thread that were accepted although a
 * run boundary still disagreed at the float32 warm-up cap (16384 frames): 0 means every accepted loss was verified. */
int eks_last_unverified_count(void);

/* ---- ensemble statistics: replaces eks.core.ensemble / compute_stats (eks/core.py:25-101) --------
 * raw: [n_sessions][M][V][T][K][3] in the reference MarkerArray layout (eks/marker_array.py:15-30),
 * raw_dtype f32|f64 (values are cast to out_dtype before any arithmetic, as core.py:90-92 does).
 * Writes 5 planes [x_avg, y_avg, var_x, var_y, likelihood] per (session, camera, keypoint) at
 *   out[s*sess_stride + v*cam_stride + k*kp_stride + plane_off[f] + t].
 * moment_partials (nullable): [n_sessions*V*K][ceil(T/eks_ensemble_tile_frames(M,K,raw_dtype,out_dtype))][4] doubles
 * receiving per-tile sums (x, y, x^2, y^2) of the averaged coordinates, consumed by eks_center_moments.
 * avg_median: 1 median | 0 mean.  var_mode: 1 confidence_weighted_var | 0 var. */
int eks_ensemble_tile_frames(int M, int K, int raw_dtype, int out_dtype);
int eks_ensemble_stats(const void* raw, int raw_dtype, long long raw_sess_stride, int n_sessions, int M, int V,
                       int T, int K, int avg_median, int var_mode, double nan_replacement, void* out, int out_dtype,
                       long long sess_stride, long long cam_stride, long long kp_stride,
                       const long long* plane_off_host, double* moment_partials, void* stream);

/* ---- centring moments: the all-frames case of center_predictions (eks/utils.py:293-365 with
 * quantile 100, eks/singlecam_smoother.py:155-157) and S0 = diag(nanvar) (singlecam_smoother.py:262-266).
 * mean_out/var_out: [n_seq][2] real. */
int eks_center_moments(const double* moment_partials, int n_seq, int n_tiles, int T, void* mean_out, void* var_out,
                       int dtype, void* stream);

/* ---- initial guess: compute_initial_guesses + caller fallback (eks/core.py:104-133, :233-236) and
 * the float32 log seed (core.py:612-613, :622).  var view = raw ensemble variances (unclipped).
 * guess_out: [B] double.  s_log0_out (nullable): [B] real. */
int eks_initial_guess(const void* var_base, long long seq_stride, const long long* chan_off_host, int dtype, int B,
                      int O, int T, double* guess_out, void* s_log0_out, void* stream);

/* ---- constant R for the loss path: constant_R_from_timevarying on cropped frames
 * (eks/core.py:599-602, :702-709; crop_frames eks/utils.py:235-290).  Exact nanmedian over the
 * frames in the spans (n_spans == 0: all frames; spans sorted, non-overlapping, host arrays),
 * floored at max(1e-12, min_var).  Rconst_out: [B][O] real.
 * Sequences of >= 131072 frames: the median is bracketed from 4096 samples, ONE pass over the planes counts the keys
 * below the bracket and compacts those inside it, and the exact radix select runs on the candidates (bit-identical
 * order statistics); problems whose bracket missed are redone by the three-pass radix select used for short ones. */
size_t eks_const_R_median_workspace_bytes(int dtype, int B, int O, int T);
int eks_const_R_median(const void* var_base, long long seq_stride, const long long* chan_off_host, int dtype, int B,
                       int O, int T, int n_spans, const int* span_start_host, const int* span_end_host,
                       double min_var, void* Rconst_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- state-space model shared by the calls below (params_nlgssm_for_keypoint, eks/core.py:136-155)
 * Per-sequence arrays, real, row-major: m0 [B][D], S0 [B][D][D], A [B][D][D], Q [B][D][D] (scaled by s
 * inside), C [B][O][D] (linear emission; ignored when ncam > 0).  ncam > 0 selects the calibrated
 * pinhole emission of make_projection_from_camgroup (eks/multicam_smoother.py:806-885): cams is
 * [ncam][EKS_CAM_STRIDE] real, O must equal 2*ncam and D 3.
 * Observations: y view (+ optional per-sequence offset ymean [B][O] subtracted on load, i.e. the
 * centring of center_predictions) and either Rconst [B][O] or a var view (clipped at 1e-12 on load,
 * build_R_from_vars eks/utils.py:368-377). */

/* NLL of the EKF filter and d NLL / d s for each sequence at s[b]: the loss of
 * _vmap_optimize_singletons._optimize_one.loss (eks/core.py:640-650) incl. the non-finite -> 1e12 rule.
 * Frames restricted to the spans (n_spans == 0: all).  nll_out, dnll_ds_out: [B] real.
 * Sequences of 512 frames or more are evaluated run-parallel (verified, see eks_optimize_s); the scratch for that
 * is taken from and returned to the stream's memory pool (cudaMallocAsync / cudaFreeAsync), the call stays async. */
int eks_nll_grad(int dtype, int B, int D, int O, int T, const void* m0, const void* S0, const void* A, const void* Q,
                 const void* C, int ncam, const void* cams, const void* y_base, long long y_seq_stride,
                 const long long* y_chan_off_host, const void* ymean, const void* Rconst, int n_spans,
                 const int* span_start_host, const int* span_end_host, const void* s, void* nll_out,
                 void* dnll_ds_out, void* stream);

/* Device-resident Adam loop on log s: replaces optimize_smooth_param / _vmap_optimize_singletons and
 * the block slow path (eks/core.py:306-401, 403-559, 562-699).  Block j owns members
 * members[block_off[j] .. block_off[j+1]) (sequence indices; device int arrays) and shares one s;
 * loss = sum of member NLLs.  s_log0: [n_blocks] real (float32-rounded seed).  Outputs [n_blocks]:
 * s_log_out (real, the value AFTER the last update, core.py:675), last_loss_out (real), iters_out (int).
 * The caller forms s = exp(clip(s_log, lo, hi)) (core.py:694).  trace (nullable): [n_blocks][trace_cap][3]
 * real rows (s_log, loss, lr*grad) per iteration.  model_structure: EKS_STRUCT_GENERAL (T < 512: one block per thread,
 * sequential in time; T >= 512: linear models with A = I, D = 3, O in {4, 6, 8} and one span use the lag statistics of
 * the stationary signal in ONE pass over the observations and a persistent Adam warp per block with the closed-form
 * NLL -- eks_b200/csrc/lin_lag.cu; every other model, and blocks for which the closed form is not exact, use the
 * verified run-parallel evaluation of generic_runs.cu, enqueued in chunks of 32 evaluations.  Both synchronise
 * `stream`), EKS_STRUCT_DIAG (needs <= 1 span: lag statistics of the increments in ONE pass over the
 * observations, then the whole Adam loop of a block in one persistent CTA evaluating the NLL in closed form from them,
 * streaming an evaluation only where the closed form is not exact to rounding -- eks_b200/csrc/diag_lag.cu) or
 * EKS_STRUCT_DIAG_STREAM (the exact time-parallel streaming evaluation once per Adam iteration, one launch each --
 * eks_b200/csrc/diag.cu; also selected by the environment variable EKS_OPT_MODE=stream). */
size_t eks_optimize_s_workspace_bytes(int dtype, int n_blocks, int B, int D, int O, int T);
int eks_optimize_s(int dtype, int B, int D, int O, int T, const void* m0, const void* S0, const void* A,
                   const void* Q, const void* C, int ncam, const void* cams, const void* y_base,
                   long long y_seq_stride, const long long* y_chan_off_host, const void* ymean, const void* Rconst,
                   int n_spans, const int* span_start_host, const int* span_end_host, int n_blocks,
                   const int* block_off, const int* members, const void* s_log0, double lr, double s_log_lo,
                   double s_log_hi, double tol, int safety_cap, void* s_log_out, void* last_loss_out,
                   int* iters_out, void* trace, int trace_cap, int model_structure, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Final pass: EKF filter + RTS smoother over ALL frames with time-varying diagonal R_t from the var
 * view: replaces vmap(_smooth_one) / extended_kalman_smoother (eks/core.py:274-295).
 * s: [B] real.  ms_out [B][T][D], Vs_out [B][T][D][D] real (reference return layout, core.py:296-297).
 * The filter represents dynamax's 1e-9 gain boost exactly (updates with R + eps, then P_f += eps P_f (H^T R'^-2 H) P_f).
 * T >= 512: run-parallel execution with verified warm-up; this entry point then SYNCHRONISES `stream` (it reads two
 * verification flags and repeats the pass with a longer warm-up if a run boundary disagreed). */
size_t eks_filter_smooth_workspace_bytes(int dtype, int B, int D, int T);
int eks_filter_smooth(int dtype, int B, int D, int O, int T, const void* m0, const void* S0, const void* A,
                      const void* Q, const void* C, int ncam, const void* cams, const void* y_base,
                      long long y_seq_stride, const long long* y_chan_off_host, const void* ymean,
                      const void* var_base, long long var_seq_stride, const long long* var_chan_off_host,
                      const void* s, void* ms_out, void* Vs_out, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Decoupled (single-camera) final pass: EKS_STRUCT_DIAG form of eks_filter_smooth fused with the
 * reprojection epilogue of eks/singlecam_smoother.py:189-217 (x = C m + mean, posterior variance =
 * diag(C V C^T)).  Writes, per sequence b, the planes out[b*out_seq_stride + out_off[j] + t], j = 0..3: smoothed x,
 * smoothed y, posterior var x, posterior var y.
 * flags: bit 0 (EKS_SMOOTH_LATENT) writes the latent smoothed means / variances themselves (the ms, diag Vs of
 * eks_filter_smooth); bit 1 (EKS_SMOOTH_EXACT_SCAN) skips the fused pass.
 * Default execution: ONE fused, time-segmented kernel (forward filter and RTS recursion of a segment in registers, halo
 * frames on both sides absorb the unknown boundary states; the contraction of both recursions over the halos is
 * bounded from the data and VERIFIED per segment), then the exact scan kernels (Moebius / affine block scans over the
 * whole sequence, filtered moments through the workspace) redo only the sequences the fused kernel flagged. */
#define EKS_SMOOTH_LATENT 1
#define EKS_SMOOTH_EXACT_SCAN 2
size_t eks_diag_smooth_workspace_bytes(int dtype, int B, int T);
int eks_diag_smooth(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q,
                    const void* C, const void* y_base, long long y_seq_stride, const long long* y_chan_off_host,
                    const void* ymean, const void* var_base, long long var_seq_stride,
                    const long long* var_chan_off_host, const void* s, void* out, long long out_seq_stride,
                    const long long* out_off_host, int flags, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Reprojection epilogue for generic models: replaces the loops of eks/multicam_smoother.py:450-511 and
 * project_3d_covariance_to_2d (:914-946).  ms [B][T][D], Vs [B][T][D][D].  For camera c of sequence b writes
 * out[b*out_seq_stride + c*out_cam_stride + plane_off[j] + t], j = 0..3: x, y, posterior var x, posterior var y.
 * Linear (ncam == 0): (x,y) = C[2c..2c+1] m + ymean, var = diag(C V C^T) (+ var view channel 2c / 2c+1).
 * Pinhole (ncam == V): (x,y) = h_c(m), var = diag(J V J^T) + var view channel; pinhole_var_quirk = 1
 * reproduces the reference, which adds channels 0 / 1 for every camera (multicam_smoother.py:943-944). */
int eks_reproject(int dtype, int B, int T, int D, int V, const void* ms, const void* Vs, const void* C,
                  const void* ymean, int ncam, const void* cams, const void* var_base, long long var_seq_stride,
                  const long long* var_chan_off_host, int pinhole_var_quirk, void* out, long long out_seq_stride,
                  long long out_cam_stride, const long long* plane_off_host, void* stream);

/* IBL pupil model: replaces pupil_optimize_smooth (eks/ibl_pupil_smoother.py:452-607).  Per session b: 3 latent
 * states [diameter, com_x, com_y], A = diag(s_d, s_c, s_c), Q = diag(var3 * (1 - s^2)), 8 observation channels
 * through C [B][8][3], time-varying diagonal R_t from the var view (clipped at 1e-12), frames restricted to the
 * spans.  Adam(lr) on u (s = sigmoid(u)(1 - 2e-3) + 1e-3, seed s = [0.99, 0.98]) with the reference's stop rule.
 * Outputs per session: u_out [B][2], s_out [B][2] (the optimised s_d, s_c), last_loss_out [B], iters_out [B].
 * trace (nullable): [B][trace_cap][3] rows (u0, u1, loss).  The final smoothing pass is eks_filter_smooth with
 * A = diag(s), Q = diag(var3 (1 - s^2)) and s = 1.  This entry point synchronises the stream between chunks of
 * 64 evaluations (the reference's cap is 5000). */
size_t eks_pupil_optimize_workspace_bytes(int dtype, int B, int T);
int eks_pupil_optimize(int dtype, int B, int T, const void* m0, const void* S0, const void* C, const void* var3,
                       const void* y_base, long long y_seq_stride, const long long* y_chan_off_host,
                       const void* ymean, const void* var_base, long long var_seq_stride,
                       const long long* var_chan_off_host, int n_spans, const int* span_start_host,
                       const int* span_end_host, double lr, double tol, int safety_cap, void* u_out, void* s_out,
                       void* last_loss_out, int* iters_out, void* trace, int trace_cap, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Multi-camera pre-stage on the device.  A problem b = session * K + keypoint owns O = 2 * cameras
 * channel planes of ensemble means (y view) and ensemble variances (var view).  The three calls share ONE workspace
 * (eks_mc_prestage_workspace_bytes) and must be issued in this order on one stream:
 *
 * eks_mc_center: replaces center_predictions (eks/utils.py:293-365).  max over channels of the variance per frame,
 *   threshold = np.percentile(., quantile_keep) per keypoint (numpy 'linear' method, evaluated in the working
 *   precision exactly as numpy does), good frames = max-variance <= threshold, n = min over the session's keypoints
 *   of the good-frame counts, ymean_out [B][O] = mean over the FIRST n good frames.  n_good_out (nullable) [B][2] =
 *   (good frames, n).  The centred predictions themselves are never materialised: every consumer takes ymean.
 * eks_mc_pca_moments: the data pass of sklearn PCA.fit on those n centred frames (eks/stats.py:9-64):
 *   moments_out [B][1 + O + O*O] (double) = n, sum_t x_t, sum_t x_t x_t^T with x = y - ymean.  The O x O
 *   eigen-decomposition (and sklearn's sign convention) is host work on these moments.
 * eks_mc_latent_init: replaces pca.transform + initialize_kalman_filter_pca (eks/multicam_smoother.py:554-597).
 *   z_t = ((y_t - ymean) - pca_mean) @ components for ALL good frames; S0_out [B][L][L] = diag(np.var(z)),
 *   Q_out [B][L][L] = np.cov of the differences of consecutive good frames, divided by its largest |entry|.
 *   components [B][O][L] (= pca.components_.T, the observation matrix C), pca_mean [B][O]. */
size_t eks_mc_prestage_workspace_bytes(int dtype, int B, int O, int T);
int eks_mc_center(int dtype, int S, int K, int O, int T, const void* y_base, long long y_seq_stride,
                  const long long* y_chan_off_host, const void* var_base, long long var_seq_stride,
                  const long long* var_chan_off_host, double quantile_keep, void* ymean_out, int* n_good_out,
                  void* workspace, size_t workspace_bytes, void* stream);
int eks_mc_pca_moments(int dtype, int B, int O, int T, const void* y_base, long long y_seq_stride,
                       const long long* y_chan_off_host, const void* ymean, double* moments_out, void* workspace,
                       size_t workspace_bytes, void* stream);
int eks_mc_latent_init(int dtype, int B, int O, int L, int T, const void* y_base, long long y_seq_stride,
                       const long long* y_chan_off_host, const void* ymean, const void* pca_mean,
                       const void* components, void* S0_out, void* Q_out, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Mahalanobis variance inflation, one iteration of the while-loop of mA_compute_maha
 * (eks/multicam_smoother.py:653-725) for B problems at once; V cameras, O = 2 V channels.
 *
 * eks_mc_valid_moments: the data pass of the FactorAnalysis fit inside compute_mahalanobis (eks/stats.py:103-124):
 *   rows with max_o var < np.percentile(max_o var, v_quantile) (strict; v_quantile < 0 disables) and, when a
 *   likelihood view (V channels) is given, min likelihood >= lik_threshold.  moments_out [B][1 + O + O*O] (double) =
 *   n, sum x, sum x x^T with x = y - ymean.  Problems with active[b] == 0 (nullable: all active) are skipped.
 *   Uses the workspace of eks_mc_prestage_workspace_bytes.  The O x O factor-analysis iteration is host work.
 * eks_mc_inflate_step: stats.py:126-157 + inflate_variance (multicam_smoother.py:728-764), fp64 per-frame algebra:
 *   B_t = (W^T diag(1/(v_t + epsilon)) W)^-1, z_t, reconstruction, per-view 2x2 posterior predictive covariance and
 *   Mahalanobis distance; the variances of views with distance > threshold (both views when V == 2 and either is
 *   flagged) are multiplied by `scalar` IN PLACE in the var view; flags_out[b] = 1 if problem b inflated anything.
 *   loading [B][O][L] and mean [B][O] are device arrays of doubles (FactorAnalysis components_.T and mean_). */
int eks_mc_valid_moments(int dtype, int B, int V, int T, const void* y_base, long long y_seq_stride,
                         const long long* y_chan_off_host, const void* ymean, const void* var_base,
                         long long var_seq_stride, const long long* var_chan_off_host, const void* lik_base,
                         long long lik_seq_stride, const long long* lik_chan_off_host, double lik_threshold,
                         double v_quantile, const int* active, double* moments_out, void* workspace,
                         size_t workspace_bytes, void* stream);
int eks_mc_inflate_step(int dtype, int B, int V, int L, int T, const void* y_base, long long y_seq_stride,
                        const long long* y_chan_off_host, const void* ymean, void* var_base, long long var_seq_stride,
                        const long long* var_chan_off_host, const double* loading, const double* mean, double epsilon,
                        double threshold, double scalar, const int* active, int* flags_out, void* stream);

/* Triangulation of the calibrated multi-camera path: triangulate_3d_models(...).mean(axis=0)
 * (eks/multicam_smoother.py:888-911, :385-386; aniposelib CameraGroup.triangulate(fast=True) restated as
 * cv2.undistortPoints + pairwise cv2.triangulatePoints + nan-median over the camera pairs).
 * raw: the (M, V, T, K, 3) seed predictions (pixels; float or double), cams [V][EKS_CAM_STRIDE] DOUBLE camera
 * parameters (same packing as the pinhole emission), out [K][T][3] double = mean over the M ensemble members of the
 * triangulated points.  fp64 arithmetic. */
int eks_triangulate_mean(const void* raw, int raw_dtype, int M, int V, int T, int K, const double* cams, double* out,
                         void* stream);

/* Geometric initialisation of the calibrated model: replaces initialize_kalman_filter_geometric
 * (eks/multicam_smoother.py:600-650).  tri [B][T][3] DOUBLE (eks_triangulate_mean's layout, B = sessions x keypoints).
 * Outputs, DOUBLE: m0_out [B][3] = mean of the first min(10, T) frames; S0_diag_out [B][3] = nanvar over time + 1e-4;
 * Q_diag_out [B][3] = max((1.4826 (median|dx - median dx| + 1e-12))^2, 1e-8) of the lag-1 differences dx (NaN if the
 * column holds a NaN: np.median is not NaN-aware).  A = C = I are the caller's.  Exact radix-select medians. */
size_t eks_geometric_init_workspace_bytes(int B, int T);
int eks_geometric_init(int B, int T, const void* tri, void* m0_out, void* S0_diag_out, void* Q_diag_out,
                       void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EKS_B200_H */
