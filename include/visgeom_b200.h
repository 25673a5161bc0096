/*
 * visgeom_b200.h -- C ABI of the B200-native reprojection-residual engine.
 *
 * This is the drop-in boundary for visgeom's calibration hot path.  Every entry
 * point names the reference interface it replaces (paths relative to the visgeom
 * tree).  Plain pointers and sizes only; no C++/torch types; nothing throws across
 * the boundary.  All functions return VG_OK (0) or a negative VG_ERR_* code and
 * leave a message retrievable with vg_last_error() (thread local).
 *
 * Two nested boundaries:
 *   inner  -- the Ceres cost-function contract of GenericProjectionJac::Evaluate
 *             (include/calibration/calib_cost_functions.h:53-54,
 *              src/calibration/calib_cost_functions.cpp:28-117), batched over the
 *             images of one dataset: vg_eval_chain / vg_eval_chain_dev.
 *   outer  -- what GenericCameraCalibration asks of ceres::Problem / ceres::Solve
 *             (src/calibration/unified_calibration.cpp:514-630, :39-53):
 *             vg_problem_*.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and
 * fails with VG_ERR_CUDA when none is usable.
 */
#ifndef VISGEOM_B200_H
#define VISGEOM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_OK               0
#define VG_ERR_INVALID     -1   /* bad argument (the reference throws runtime_error for config errors) */
#define VG_ERR_CUDA        -2   /* CUDA runtime failure / no device */
#define VG_ERR_NOMEM       -3
#define VG_ERR_UNSUPPORTED -4   /* e.g. board too large for one CTA's shared memory */
#define VG_ERR_NUMERIC     -5   /* linear solve failed repeatedly */
#define VG_ERR_PEER        -6   /* a rank of the peer-memory exchange never posted its block (died, or issued a different
                                   sequence of evaluations): the sums of that exchange are NaN */

/* camera models: include/projection/eucm.h, ucm.h, mei.h; parameter order as there:
 * EUCM [alpha,beta,fu,fv,u0,v0]  UCM [xi,fu,fv,u0,v0]  MEI [xi,k1..k5,fu,fv,u0,v0] */
#define VG_MODEL_EUCM 0
#define VG_MODEL_UCM  1
#define VG_MODEL_MEI  2

/* enum TransformationStatus, include/calibration/calib_cost_functions.h:25 */
#define VG_TRANSFORM_DIRECT  0
#define VG_TRANSFORM_INVERSE 1

#define VG_MAX_CHAIN 5          /* unified_calibration.cpp:567 */
#define VG_MAX_INTRINSICS 10

/* residual written when a projection fails: DOUBLE_BIG, include/std.h:71,
 * calib_cost_functions.cpp:66-70 */
#define VG_DOUBLE_BIG 1e15

const char *vg_last_error(void);
int vg_version(void);
/* number of usable CUDA devices (0 = none; compute calls will fail) */
int vg_device_count(void);

/* ICamera::numParams / lowerBound / upperBound -- generic_camera.h:116-119,
 * eucm.h:228-246, ucm.h:199-215, mei.h:287-313 */
int vg_model_num_params(int model);
int vg_model_bounds(int model, int idx, double *lower, double *upper);

/* number of doubles of one per-image normal-equation block: the packed upper
 * triangle of [J r]^T [J r], W = K + 6*chain_len + 1, (W)(W+1)/2; column order
 * [intrinsics(K), chain element 0 (6), .., chain element L-1 (6), residual]. */
int vg_hessian_entries(int model, int chain_len);

/* ------------------------------------------------------------------------- *
 * Inner boundary: GenericProjectionJac::Evaluate batched over n_img images.
 * Replaces, per image: calib_cost_functions.cpp:28-117 (+ InterJacobian
 * jacobian.h:136-171, ICamera::{projectPoint,projectionJacobian,
 * intrinsicJacobian}, Transformation::compose/composeInverse/transform).
 *
 *   intr      K doubles                      (params[0] of the functor)
 *   board     P x 3 doubles                  (_grid)
 *   obs       n_img x P x 2 doubles          (_proj of each image's functor)
 *   status    chain_len x {DIRECT,INVERSE}   (_transformStatusVec)
 *   is_global chain_len flags; xi[e] points at 6 doubles if global, else at
 *             n_img x 6 doubles [tx,ty,tz,rx,ry,rz] (params[1+e] of image i)
 *   r         n_img x 2P doubles, [u0,v0,u1,v1..] = projected - observed, or
 *             (1e15,1e15) where the projection fails.           May be NULL.
 *   J_intr    n_img x 2P x K, row-major rows u_i, v_i.          May be NULL.
 *   J_xi      chain_len pointers (array may be NULL, entries may be NULL),
 *             each n_img x 2P x 6, columns [d/dt(3), d/dr(3)].
 *   H         n_img x vg_hessian_entries() per-image normal-equation blocks
 *             (what Ceres accumulates from the block's Jacobians). May be NULL.
 * Host variant: all pointers are host memory; the call copies in, launches,
 * copies out and synchronises.  Device variant: all pointers are device memory
 * on the current device (16-byte aligned), the launch is asynchronous on
 * `stream` (a cudaStream_t passed as void*).
 * ------------------------------------------------------------------------- */
int vg_eval_chain(int model, const double *intr, int n_img, int P,
                  const double *board, const double *obs,
                  int chain_len, const int *status, const int *is_global,
                  const double *const *xi,
                  double *r, double *J_intr, double *const *J_xi, double *H);
/* Optional: page-lock a caller-owned array (cudaHostRegister) so that vg_eval_chain's DMA writes it directly instead of
 * going through a staging slot and a host memcpy -- worth it for arrays that live across many calls, as the residual /
 * Jacobian arrays Ceres hands to CostFunction::Evaluate do.  vg_eval_chain recognises page-locked output arrays by
 * itself (each output independently); results are identical either way.  Unregister before freeing the memory. */
int vg_host_register(void *p, size_t bytes);
int vg_host_unregister(void *p);

int vg_eval_chain_dev(int model, const double *intr, int n_img, int P,
                      const double *board, const double *obs,
                      int chain_len, const int *status, const int *is_global,
                      const double *const *xi, const int *seq_index /* nullable, device */,
                      double *r, double *J_intr, double *const *J_xi, double *H,
                      void *stream);

/* vg_eval_chain keeps a grow-only device workspace between calls; this frees it */
void vg_release_workspace(void);

/* Number of kernel launches this library has issued in the calling process
 * (monotonic counter; used by bench.py for its gpu_launches claim). */
unsigned long long vg_launch_count(void);

/* ------------------------------------------------------------------------- *
 * Outer boundary: the problem GenericCameraCalibration builds and solves.
 * ------------------------------------------------------------------------- */
typedef struct vg_problem vg_problem;

/* Solver::Options as the reference sets them (unified_calibration.cpp:42-53);
 * remaining fields are Ceres' documented trust-region defaults. */
typedef struct {
    int max_num_iterations;        /* 1000 */
    double function_tolerance;     /* 1e-15 */
    double gradient_tolerance;     /* 1e-15 */
    double parameter_tolerance;    /* 1e-15 */
    double initial_radius;         /* 1e4 */
    double max_radius;             /* 1e16 */
    double min_radius;             /* 1e-32 */
    double min_relative_decrease;  /* 1e-3 */
    double min_lm_diagonal;        /* 1e-6 */
    double max_lm_diagonal;        /* 1e32 */
    int jacobi_scaling;            /* 1 */
    int max_consecutive_invalid;   /* 5 */
    int verbose;                   /* minimizer_progress_to_stdout */
    int reserved;
} vg_solve_options;

/* Solver::Summary subset */
typedef struct {
    int iterations;
    int num_successful;
    int num_unsuccessful;
    int termination;               /* 0 function tol, 1 gradient tol, 2 parameter tol, 3 max iterations, 4 radius, 5 failure */
    double initial_cost;
    double final_cost;
    double seconds_total;
    double seconds_evaluate;       /* device time in the fused residual+Jacobian+normal-equation kernels */
    int num_evaluations;
} vg_solve_summary;

void vg_solve_options_default(vg_solve_options *o);

/* ---- inner level, the other functors of the global problem -------------------------------
 * Batched over n independent blocks, host buffers, Ceres layouts (residual 6, Jacobians row-major 6 x 6).
 *
 * TransformationPrior(stiffness, xi_prior)::Evaluate(xi)   calib_cost_functions.h:83-108, .cpp:215-228
 *   r = A (R e_t, R e_r), e = xi_prior^-1 o xi, A = diag(stiffness) with its rotation block times
 *   interOmegaRot(prior rotation); the Jacobian the functor reports is A itself.  J nullable. */
int vg_eval_transformation_prior(int n, const double *stiffness /* n x 6 */, const double *xi_prior /* n x 6 */,
                                 const double *xi /* n x 6 */, double *r /* n x 6 */, double *J /* n x 36 */);
/* OdometryPrior(errV, errW, lambda, odom1, odom2)::Evaluate(xi1, xi2)   calib_cost_functions.cpp:119-213
 *   zeta_prior = odom1^-1 o odom2, r = A (zeta_prior^-1 o (xi1^-1 o xi2)); J1, J2 nullable. */
int vg_eval_odometry_prior(int n, double errV, double errW, double lambda,
                           const double *odom1 /* n x 6 */, const double *odom2 /* n x 6 */,
                           const double *xi1 /* n x 6 */, const double *xi2 /* n x 6 */,
                           double *r /* n x 6 */, double *J1 /* n x 36 */, double *J2 /* n x 36 */);
/* OdometryCost (include/calibration/odometry_cost_function.h:33-59, src/calibration/odometry_cost_function.cpp:144-267:
 * the residual block of "odometry_intrinsic" datasets, unified_calibration.cpp:718-731) for n blocks: the motion of a
 * differential-drive platform integrated from wheel-angle increments against xi1^-1 o xi2.  Block b owns the pairs
 * (left, right) dq[2 dq_offset[b] .. 2 dq_offset[b + 1]) (at least one); intr_prior = the (r1, r2, g) the constructor
 * builds _zetaPrior and _A from, intr = the parameter block; r 6 per block, J1 / J2 (6 x 6, d r / d xi1, xi2) and J3
 * (6 x 3, d r / d intr) row-major, any of the three may be NULL.  (The reference declares the functor
 * SizedCostFunction<6, 6> while handing it three blocks, so its own "odometry_intrinsic" datasets cannot be added to a
 * Ceres problem; Evaluate itself is well defined and is what this entry point offers.) */
int vg_eval_odometry_cost(int n, double errV, double errW, double lambda, const int *dq_offset, const double *dq,
                          const double *intr_prior, const double *xi1, const double *xi2, const double *intr, double *r,
                          double *J1, double *J2, double *J3);

/* TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206), batched over n camera poses:
 * the covariance of a camera pose localised on the board, from dP/dX of the model at every board point:
 *   [t, r] = camPose^-1 o xiBoard, X_i = R b_i + t, J_i = dpdx(X_i) [-I | hat(X_i)],
 *   JtJ = sum J_i^T J_i, JtCJ = sum J_i^T S J_i with S = U of the Cholesky factorisation of diag(1 / feature_variance),
 *   cov = JtCJ^-T JtJ JtCJ^-1   (6 x 6 row-major per pose). */
int vg_visual_cov(int model, const double *intr, const double *xi_board /* 6 */, int P, const double *board /* P x 3 */,
                  double feature_variance, int n, const double *cam_poses /* n x 6 */, double *cov /* n x 36 */);

/* device < 0 -> current device */
vg_problem *vg_problem_create(int device);
void vg_problem_destroy(vg_problem *p);

/* parseCameras, unified_calibration.cpp:134-180: returns camera id >= 0.
 * Bounds default to the model's (SetParameterLower/UpperBound, :621-626);
 * constant -> SetParameterBlockConstant (:614-617). */
int vg_problem_add_camera(vg_problem *p, int model, const double *value, int constant);
int vg_problem_set_bounds(vg_problem *p, int camera, int idx, double lower, double upper);

/* parseTransforms, :91-132: a global transform (n == 1) or a sequence of n
 * transforms; values n x 6 [t,r]; constant -> SetParameterBlockConstant (:604-610).
 * Returns transform id >= 0. */
int vg_problem_add_transform(vg_problem *p, int is_global, int constant, int n, const double *values);

/* addResiduals "transformation_prior", unified_calibration.cpp:808-829: a TransformationPrior block on element
 * `index` of a transform (0 for a global one; the reference always takes element 0, unified_calibration.h:161-165).
 * xi_prior NULL -> the element's value at the time of the call, which is what the reference passes (:826-827).
 * Returns the block id >= 0. */
int vg_problem_add_transformation_prior(vg_problem *p, int transform, int index, const double *stiffness /* 6 */,
                                        const double *xi_prior /* 6 or NULL */);
/* addResiduals "odometry", :742-807: one OdometryPrior block per pair of consecutive elements of a sequence
 * transform, built from n = its length odometry readings (n x 6).  The pose part of the normal equations becomes
 * block tridiagonal along that sequence.  Returns the number of blocks added. */
int vg_problem_add_odometry(vg_problem *p, int transform, double errV, double errW, double lambda, int n,
                            const double *odom /* n x 6 */);
/* The LossFunction argument of AddResidualBlock for every block of a dataset: a > 0 -> SoftLOneLoss(a) (the
 * initialisation solves use it, unified_calibration.cpp:379-404 with a = 1, :1143 with a = 25), 0 -> NULL (the
 * global problem, :539-565).  Ceres applies the loss to the squared norm of the whole block, i.e. per image. */
int vg_problem_set_loss(vg_problem *p, int dataset, double a);
/* "anchor" (:803-806): SetParameterBlockConstant on ONE element of a sequence transform */
int vg_problem_set_pose_constant(vg_problem *p, int transform, int index, int constant);

/* addGridResidualBlocks, :514-568: one residual block per image of the dataset.
 * obs n_img x P x 2 (host).  seq_index (nullable -> identity) maps image i to the
 * element of the chain's sequence transform; images without an extracted board
 * are simply not listed (:520).  Exactly one chain element must be a sequence
 * (:223-228), chain_len <= 5 (:567).  Returns dataset id >= 0. */
int vg_problem_add_dataset(vg_problem *p, int camera, int P, const double *board,
                           int n_img, const double *obs, const int *seq_index,
                           int chain_len, const int *transform_ids, const int *status);

/* Route the cross-GPU sum of the reduced normal equations through the host
 * application (one process per GPU): fn must sum `count` doubles at device
 * pointer `buf` in place across all ranks, ordered after the work already queued
 * on `stream`.  NULL -> single GPU.  Images are sharded by the caller: each rank
 * adds only its own images/poses; shared parameters are replicated. */
typedef int (*vg_allreduce_fn)(void *ctx, double *buf, int count, void *stream);
int vg_problem_set_allreduce(vg_problem *p, vg_allreduce_fn fn, void *ctx, int rank, int nranks);

/* The same sum over peer memory instead (GPUs of one NVLink / NVSwitch domain, one process per GPU): the kernel that
 * assembles the reduced normal equations also exchanges them with the other ranks -- peer stores into every rank's
 * inbox, a flag, a sum in rank order (bit-identical on every rank) -- with no collective launch and no host in the
 * loop.  Each rank exports the CUDA IPC handle of its inbox (64 bytes), the application gathers the handles of all
 * ranks (any transport) and hands the whole array to every rank.  Call before the first evaluation; every rank must
 * then issue the same sequence of evaluations / solves on this problem. */
#define VG_IPC_HANDLE_BYTES 64
int vg_problem_peer_export(vg_problem *p, void *ipc_handle_out /* VG_IPC_HANDLE_BYTES */);
int vg_problem_peer_connect(vg_problem *p, int rank, int nranks, const void *ipc_handles /* nranks x VG_IPC_HANDLE_BYTES */);
/* The same for problems that live in ONE process (one host thread driving several GPUs, or several problems sharing a
 * GPU): every problem reports the device address of its inbox, and is handed the addresses of all of them in rank
 * order.  Devices other than the problem's own must be peer-accessible (cudaDeviceEnablePeerAccess is attempted). */
int vg_problem_peer_inbox(vg_problem *p, void **inbox_out, int *device_out);
int vg_problem_peer_connect_local(vg_problem *p, int rank, int nranks, void *const *inboxes, const int *devices);
/* A collect gives up after `polls` reads of a word that never arrives (0: the default, more than ten seconds); the
 * call that next synchronises with the device then returns VG_ERR_PEER. */
int vg_problem_set_peer_timeout(vg_problem *p, long long polls);

/* When enabled, every evaluation also materialises r and all Jacobian blocks in
 * device memory in the Ceres layout (what GenericProjectionJac::Evaluate hands to
 * Ceres); the LM loop itself only needs the per-image normal-equation blocks and
 * runs with this off. */
int vg_problem_materialize_jacobians(vg_problem *p, int enable);
/* device pointers of dataset buffers (valid until the problem is destroyed):
 * which = 0 residual, 1 J_intr, 2+e J_xi[e], -1 observations, -2 normal-equation blocks */
int vg_problem_device_buffer(vg_problem *p, int dataset, int which, void **ptr, size_t *bytes);
/* the CUDA stream (cudaStream_t) all of this problem's work is queued on; set_stream
 * makes the problem use a caller-owned stream instead of its own */
void *vg_problem_stream(vg_problem *p);
int vg_problem_set_stream(vg_problem *p, void *stream);
/* vg_problem_evaluate without the read-back: queues the fused kernels, the shared-block
 * reduction and (multi-GPU) its all-reduce on the problem's stream and returns.
 * vg_problem_fetch_reduced then copies cost / reduced (nullable) to the host and waits.
 * Consecutive evaluations queued on one stream overlap on the device (a launch's main loop does not wait for the
 * evaluation ahead of it); stream order holds for everything else: a copy or a kernel queued after an evaluation sees its
 * complete result, and parameters uploaded between two evaluations are seen by the second one only. */
int vg_problem_evaluate_async(vg_problem *p);
int vg_problem_fetch_reduced(vg_problem *p, double *cost, double *reduced);

/* ceres::Solve, :53.  Parameters are updated in place inside the handle. */
int vg_problem_solve(vg_problem *p, const vg_solve_options *o, vg_solve_summary *s);

/* One residual + Jacobian + normal-equation pass at the current parameters
 * (what one Ceres evaluation costs): cost = 1/2 sum r^2 over all datasets.
 * reduced (nullable, host) receives Ks*Ks + Ks doubles: J^T J and J^T r of the
 * shared (intrinsic + global transform) block. */
int vg_problem_evaluate(vg_problem *p, double *cost, double *reduced);
int vg_problem_num_shared(vg_problem *p);

int vg_problem_get_camera(vg_problem *p, int camera, double *out);
int vg_problem_set_camera(vg_problem *p, int camera, const double *value);
int vg_problem_get_transform(vg_problem *p, int transform, double *out /* n x 6 */);
int vg_problem_set_transform(vg_problem *p, int transform, const double *values);
/* replace the observations / poses of a dataset from (pinned) host memory -- the
 * per-step input upload of the end-to-end benchmark */
int vg_problem_update_observations(vg_problem *p, int dataset, const double *obs);
/* same for a sequence transform's poses (n x 6); both calls are asynchronous on the problem's
 * stream: the host buffers must stay valid until the next call that waits (fetch / evaluate /
 * solve / get_*) */
int vg_problem_update_poses(vg_problem *p, int transform, const double *values);
/* residuals of one dataset at the current parameters (writeImageResidual,
 * :1186-1213 needs err = -r and proj = r + obs), n_img x 2P doubles to host */
int vg_problem_residuals(vg_problem *p, int dataset, double *r);

/* ---- per-image initialisation solves (unified_calibration.cpp:1131-1155) ---------------------------------------------
 * estimateInitialGrid's refinement, for n_img images at once but as n_img INDEPENDENT problems: image i's board pose
 * poses[6 i .. 6 i + 5] = [t, r] (camera <- board, a single DIRECT transform) is the only free block of its own
 * Levenberg-Marquardt solve -- camera constant (:1145), the block under SoftLOneLoss(loss_a) (:1143; 0: no loss), own
 * trust region, own accept / reject, own termination -- one warp per image, no host synchronisation in the loop.
 * opt: NULL = vg_solve_options_default (the calibration's 1e-15 tolerances); the reference's per-image solve runs
 * with Ceres' defaults (function 1e-6, gradient 1e-10, parameter 1e-8) and max_num_iterations = 500 (:1148).
 * iterations / final_cost / termination (vg_solve_summary's codes): n_img entries each, may be NULL. */
int vg_refine_poses(int model, const double *intr, int n_img, int P, const double *board, const double *obs,
                    double *poses, double loss_a, const vg_solve_options *opt, int *iterations, double *final_cost,
                    int *termination);

/* ---- the ICamera point API, batched (include/projection/generic_camera.h:36-113) -------------------------------------
 * What a visgeom caller does with a camera object outside the calibration functor: projectPoint (:39),
 * projectionJacobian (:46), intrinsicJacobian (:50), reconstructPoint (:36) and the *PointCloud loops around them
 * (:64-113), for n points in one launch.  Layouts: X n x 3, uv n x 2; dPdX n x 6 = [du/dX (3), dv/dX (3)] per point
 * (the two arrays projectionJacobian fills); dPdintr n x 2K = [du/dintr (K), dv/dintr (K)] (intrinsicJacobian);
 * ok n bytes = the bool the reference's calls return.  A failed point (EUCM only: eucm.h:46-54, :100) keeps the
 * caller's uv / X entry and gets zero Jacobians.  Any output pointer may be NULL.  _dev: device pointers, asynchronous
 * on `stream`. */
int vg_project_points(int model, const double *intr, long long n, const double *X, double *uv, double *dPdX,
                      double *dPdintr, unsigned char *ok);
int vg_project_points_dev(int model, const double *intr, long long n, const double *X, double *uv, double *dPdX,
                          double *dPdintr, unsigned char *ok, void *stream);
int vg_reconstruct_points(int model, const double *intr, long long n, const double *uv, double *X, unsigned char *ok);
int vg_reconstruct_points_dev(int model, const double *intr, long long n, const double *uv, double *X, unsigned char *ok,
                              void *stream);

/* ---- checkerboard detector, first stage (src/calibration/corner_detector.cpp:262-329) ------------------------------
 * CornerDetector::computeResponse for n_img 8-bit images of width x height (row-major, one after the other): the two
 * Gaussian blurs (cv::GaussianBlur of an 8-bit image: OpenCV's bit-exact fixed-point path), the sharp gradient maps
 * _gradx, _grady, _imgrad and the saddle response _resp (all float, one value per pixel, zero on the one-pixel border),
 * avg = _avgVal (mean of the responses kept, :318) and count of them per image.  sigma1 / sigma2 are computeResponse's
 * arguments (detectPattern calls it with 0.7 and 1.4, 2, 1: :229-233); filter sizes 3 and 1 + 2 ceil(sigma2).  Every
 * float equals the CPU restatement's bit for bit; avg is summed in a different, fixed order (last-bit differences).
 * avg / count may be NULL. */
int vg_corner_response(const unsigned char *img, int n_img, int width, int height, double sigma1, double sigma2,
                       float *resp, float *gradx, float *grady, float *imgrad, double *avg, long long *count);
int vg_corner_response_dev(const unsigned char *img, int n_img, int width, int height, double sigma1, double sigma2,
                           float *resp, float *gradx, float *grady, float *imgrad, double *avg, long long *count,
                           void *stream);

/* ---- checkerboard detector (src/calibration/corner_detector.cpp:223-260; include/calibration/corner_detector.h:49-60) --
 * CornerDetector(Nx, Ny, 3, improve).setImage(img) + detectPattern(ptVec) for n_img 8-bit images of width x height (row
 * major, one after the other), as GenericCameraCalibration::extractGridProjections runs it per image
 * (unified_calibration.cpp:995,1031-1033).  found[i] = detectPattern's return value; corners + i * Nx * Ny * 2 = ptVec
 * (u, v per corner, board order) when found.  The scales 1.4 / 2 / 1 are tried in turn per image (:225-245).  On the
 * GPU: blurs, gradients, response (:262-329), the scan for local maxima (:494-534) and, with improve != 0, the sub-pixel
 * refinement (:162-198: SubpixelCorner minimised with ceres::GradientProblemSolver's defaults, one warp per corner);
 * on host threads, one image each: candidate tests, the flood-fill graph and the pattern search (:331-492, :536-1076).
 * Integer corner positions equal the reference's exactly; refined positions to a tolerance (the minimiser is a
 * restatement of Ceres' line search). */
int vg_detect_pattern(const unsigned char *img, int n_img, int width, int height, int Nx, int Ny, int improve,
                      double *corners, unsigned char *found);
/* The host stages of one scale on their own (no GPU needed): image, its two blurred copies (_src1, _src2), the local
 * maxima of the response (value, u v) in any order, INIT_RADIUS -> candidates in the order constructGraph numbers them
 * (cand, at most cand_cap; n_cand = how many there are), the grid (Nx Ny integer corners), initPoin's start values
 * (5 per corner, :1261-1298) and improveCorners' radMax per corner (:164-175).  Outputs may be NULL.  Returns 1 when the
 * pattern was found, 0 when not, a negative error code otherwise. */
int vg_detector_host_stages(const unsigned char *img, const unsigned char *s1, const unsigned char *s2, int width, int height,
                            const float *max_val, const int *max_uv, int n_max, int Nx, int Ny, int init_radius, int *cand,
                            int cand_cap, int *n_cand, int *grid, double *start, double *reach);
/* SubpixelCorner(gradu, gradv, prior_i, 7, length_i).Evaluate(params_i) (:47-100) for n parameter vectors (5 doubles
 * each) on one pair of gradient maps: cost[n], gradient[5 n].  Host buffers. */
int vg_subpixel_evaluate(const float *gradx, const float *grady, int width, int height, int n, const double *prior,
                         const double *length, const double *params, double *cost, double *gradient);
/* improveCorners' minimisation alone (:176-196) for n corners on one pair of gradient maps: start = initPoin's values,
 * refined[2 n] = the minimiser's u, v; iterations may be NULL. */
int vg_subpixel_refine(const float *gradx, const float *grady, int width, int height, int n, const double *prior,
                       const double *length, const double *start, double *refined, int *iterations);

#ifdef __cplusplus
}
#endif
#endif
