/*
 * visgeom_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C99, fp64 restatement of the visgeom calibration hot path, written
 * from the reference's behaviour (file:line citations are relative to
 * /root/reference).  It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the
 * CUDA path against the reference algorithm.  Nothing in visgeom_b200/ may
 * include, link or call it.
 *
 * Parity status: the reference ships no golden vectors for this path
 * (SURVEY.md section 8c).  The oracle is pinned against the reference's own
 * source compiled with stand-in Eigen/Ceres headers (oracle/_ref, see
 * oracle/Makefile) and against finite differences / high precision numpy.
 */
#ifndef VISGEOM_ORACLE_H
#define VISGEOM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* include/std.h:71 */
#define VGO_DOUBLE_BIG 1e15

/* camera model ids; parameter order as in eucm.h:34-39, ucm.h:37-41, mei.h:33-42 */
enum { VGO_EUCM = 0, VGO_UCM = 1, VGO_MEI = 2 };
/* calib_cost_functions.h:25 */
enum { VGO_TRANSFORM_DIRECT = 0, VGO_TRANSFORM_INVERSE = 1 };

int vgo_num_params(int model);                                 /* 6 / 5 / 10 */
double vgo_lower_bound(int model, int idx);                    /* eucm.h:238-246 etc */
double vgo_upper_bound(int model, int idx);                    /* eucm.h:228-236 etc */

/* ---- geometry (include/geometry) ---- */
void vgo_rotation_matrix(const double v[3], double R[9]);      /* geometry_core.h:40-76, row-major */
void vgo_inter_omega_rot(const double v[3], double B[9]);      /* geometry_core.h:158-180 */
void vgo_quat_from_rotvec(const double r[3], double q[4]);     /* quaternion.h:31-50 (x,y,z,w) */
void vgo_quat_to_rotvec(const double q[4], double r[3]);       /* quaternion.h:84-98 */
void vgo_quat_rotate(const double q[4], const double v[3], double out[3]); /* quaternion.h:61-82 */
void vgo_quat_mul(const double a[4], const double b[4], double out[4]);    /* quaternion.h:105-118 */
/* transforms are 6 doubles [t(3), r(3)] (transformation.h:46) */
void vgo_compose(const double a[6], const double b[6], double out[6]);          /* transformation.h:80-88 */
void vgo_compose_inverse(const double a[6], const double b[6], double out[6]);  /* transformation.h:101-110 */
void vgo_inverse_compose(const double a[6], const double b[6], double out[6]);  /* transformation.h:90-99 */
void vgo_transform_point(const double xi[6], const double src[3], double dst[3]); /* transformation.h:147-155 */

/* ---- cameras (include/projection) ---- */
/* return 1 = projected, 0 = failed (outputs untouched for project, zero for jacobians) */
int vgo_project(int model, const double *params, const double X[3], double uv[2]);
int vgo_projection_jacobian(int model, const double *params, const double X[3],
                            double dudx[3], double dvdx[3]);
int vgo_intrinsic_jacobian(int model, const double *params, const double X[3],
                           double *dudalpha, double *dvdalpha);
int vgo_reconstruct(int model, const double *params, const double uv[2], double X[3]);

/* ---- InterJacobian (projection/jacobian.h:136-194) ---- */
typedef struct {
    int model;
    const double *params;
    double R12[9], M12[9], t13[3];
} vgo_inter_jacobian;
void vgo_inter_jacobian_init(vgo_inter_jacobian *ij, int model, const double *params,
                             const double xi13[6], const double xi23[6], int inverted);
void vgo_dpdxi(const vgo_inter_jacobian *ij, const double X1[3], double dudxi[6], double dvdxi[6]);

/* ---- GenericProjectionJac::Evaluate (calib_cost_functions.cpp:28-117) ----
 * params[0] = K intrinsics, params[1+j] = chain element j (6 doubles).
 * residual: 2P doubles.  jacobian may be NULL; jacobian[b] may be NULL,
 * else row-major 2P x blocksize(b).  Always returns 1 (the reference returns true). */
int vgo_evaluate(int model, int P, const double *obs /*P x 2*/, const double *board /*P x 3*/,
                 int chain_len, const int *status,
                 double const *const *params, double *residual, double **jacobian);

/* Batched driver over images (what Ceres does per iteration, serial or OpenMP over
 * images).  xi[e] points at n_img x 6 doubles (sequence) or 6 doubles (global).
 * Outputs may be NULL.  H (optional) receives, per image, the packed upper triangle
 * of [J r]^T [J r] with column order [intr(K), e0(6), .., e(L-1)(6), r]
 * -- ne = (D+1)(D+2)/2 doubles, D = K + 6 L.  threads<=1 -> serial. */
int vgo_evaluate_batch(int model, const double *intr, int n_img, int P,
                       const double *board, const double *obs,
                       int chain_len, const int *status, const int *is_global,
                       const double *const *xi,
                       double *r, double *J_intr, double *const *J_xi, double *H,
                       int threads);

/* ---- the other residual types of the global problem ----
 * TransformationPrior: calib_cost_functions.h:83-108 (ctor), calib_cost_functions.cpp:215-228 (Evaluate).
 * 6 residuals on ONE transform; the Jacobian the functor hands Ceres is the constant matrix A. */
typedef struct {
    double xi_prior[6];
    double A[36];            /* row-major (Matrix6drm) */
    double R[9];             /* rotMat of the prior */
} vgo_transformation_prior;
void vgo_transformation_prior_init(vgo_transformation_prior *tp, const double stiffness[6], const double xi_prior[6]);
/* r: 6 doubles; J: 36 doubles row-major or NULL */
void vgo_transformation_prior_eval(const vgo_transformation_prior *tp, const double xi[6], double r[6], double *J);

/* OdometryPrior: calib_cost_functions.h:64-81, calib_cost_functions.cpp:119-213.
 * 6 residuals between two consecutive elements of a sequence transform. */
typedef struct {
    double zeta_prior[6];
    double A[36];            /* row-major restatement of the column-major Matrix6d _A */
} vgo_odometry_prior;
void vgo_odometry_prior_init(vgo_odometry_prior *op, double errV, double errW, double lambda,
                             const double xi1[6], const double xi2[6]);
/* r: 6 doubles; J1, J2: 36 doubles row-major (Matrix6drm maps, .cpp:195,206) or NULL */
void vgo_odometry_prior_eval(const vgo_odometry_prior *op, const double xi1[6], const double xi2[6],
                             double r[6], double *J1, double *J2);
/* OdometryCost (src/calibration/odometry_cost_function.cpp): same record (prior motion + _A); m pairs of wheel-angle
 * increments dq, odometry intrinsics (r1, r2, g) */
int vgo_odometry_cost_init(vgo_odometry_prior *oc, double errV, double errW, double lambda, int m, const double *dq,
                           const double intr_prior[3]);
void vgo_odometry_cost_eval(const vgo_odometry_prior *oc, int m, const double *dq, const double xi1[6], const double xi2[6],
                            const double intr[3], double r[6], double *J1, double *J2, double *J3);

/* TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206): 6 x 6 covariance of each camera pose
 * localised on the board; cam_poses n x 6, out n x 36 row-major */
void vgo_visual_cov(int model, const double *intr, const double xi_board[6], int P, const double *board,
                    double feature_variance, int n, const double *cam_poses, double *out);

/* The tuned CPU variant of the batched driver (oracle_tuned.c): same outputs, the chain composed once, no per-image
 * allocation, one pass per corner sharing rho / eta between projection and both Jacobians. */
int vgo_evaluate_batch_tuned(int model, const double *intr, int n_img, int P,
                             const double *board, const double *obs,
                             int chain_len, const int *status, const int *is_global,
                             const double *const *xi,
                             double *r, double *J_intr, double *const *J_xi, double *H,
                             int threads);

int vgo_hessian_entries(int K, int chain_len);   /* (D+1)(D+2)/2 */
int vgo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
