/*
 * oracle_lm.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Problem assembly + Levenberg-Marquardt solve that stands in for what the
 * reference does with Ceres:
 *   GenericCameraCalibration::addGridResidualBlocks  unified_calibration.cpp:514-630
 *   GenericCameraCalibration::compute / ceres::Solve unified_calibration.cpp:39-53
 * Ceres itself is an absent, unpinned third-party dependency (README.md:19), so
 * the loop below restates its documented trust-region LM defaults:
 * initial radius 1e4, Jacobi column scaling fixed at iteration 0, LM diagonal
 * clamp(diag(J^T J), 1e-6, 1e32)/radius, step accepted when the relative decrease
 * exceeds 1e-3, radius /= max(1/3, 1-(2 rho-1)^3) on accept, radius /= nu with nu
 * doubling on reject, box bounds by projection, function / gradient / parameter
 * tolerances.  The linear algebra uses the arrowhead structure (per-pose 6x6 Schur
 * elimination) -- mathematically the same normal equations Ceres solves.
 */
#ifndef VISGEOM_ORACLE_LM_H
#define VISGEOM_ORACLE_LM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgo_problem vgo_problem;

typedef struct {
    int max_num_iterations;        /* unified_calibration.cpp:46 -> 1000 */
    double function_tolerance;     /* :47 -> 1e-15 */
    double gradient_tolerance;     /* :48 -> 1e-15 */
    double parameter_tolerance;    /* :49 -> 1e-15 */
    double initial_radius;         /* Ceres default 1e4 */
    double max_radius;             /* 1e16 */
    double min_radius;             /* 1e-32 */
    double min_relative_decrease;  /* 1e-3 */
    double min_lm_diagonal;        /* 1e-6 */
    double max_lm_diagonal;        /* 1e32 */
    int jacobi_scaling;            /* 1 */
    int max_consecutive_invalid;   /* 5 */
    int verbose;                   /* minimizer_progress_to_stdout, :51 */
    int threads;                   /* CPU threads for the per-image loops (reference: 1) */
} vgo_solve_options;

typedef struct {
    int iterations;                /* LM iterations run (successful + unsuccessful) */
    int num_successful;
    int num_unsuccessful;
    int termination;               /* 0 convergence(function) 1 (gradient) 2 (parameter) 3 max iter 4 radius 5 failure */
    double initial_cost;
    double final_cost;
    double seconds_total;
    double seconds_evaluate;       /* time inside residual+Jacobian evaluation */
    int num_evaluations;
} vgo_solve_summary;

void vgo_solve_options_default(vgo_solve_options *o);

vgo_problem *vgo_problem_create(void);
void vgo_problem_destroy(vgo_problem *p);
/* returns camera id >= 0 or < 0 on error; bounds default to the model's (eucm.h:228-246 ..) */
int vgo_problem_add_camera(vgo_problem *p, int model, const double *value, int constant);
int vgo_problem_set_bounds(vgo_problem *p, int cam, int idx, double lo, double hi);
/* n = 1 for a global transform, = sequence length otherwise; values n x 6 [t,r] */
int vgo_problem_add_transform(vgo_problem *p, int is_global, int constant, int n, const double *values);
/* seq_index (nullable -> identity) maps image i of this dataset to the element of its
 * sequence transform (images with no extracted board are simply absent, :520) */
int vgo_problem_add_dataset(vgo_problem *p, int cam, int P, const double *board,
                            int n_img, const double *obs, const int *seq_index,
                            int chain_len, const int *transform_ids, const int *status);
/* TransformationPrior on element `index` of a transform (0 for a global one); xi_prior NULL -> the element's
 * current value, as unified_calibration.cpp:826-828 constructs it.  Returns the block id. */
int vgo_problem_add_transformation_prior(vgo_problem *p, int tr, int index, const double *stiffness, const double *xi_prior);
/* OdometryPrior between every pair of consecutive elements of a sequence transform, built from n odometry
 * readings (unified_calibration.cpp:793-802) */
int vgo_problem_add_odometry(vgo_problem *p, int tr, double errV, double errW, double lambda, int n, const double *odom);
/* the LossFunction of every block of a dataset: SoftLOneLoss(a) for a > 0, NULL for 0 (unified_calibration.cpp:379,1143) */
int vgo_problem_set_loss(vgo_problem *p, int dataset, double a);
/* SetParameterBlockConstant on one element of a sequence ("anchor", unified_calibration.cpp:803-806) */
int vgo_problem_set_pose_constant(vgo_problem *p, int tr, int index, int constant);
int vgo_problem_solve(vgo_problem *p, const vgo_solve_options *o, vgo_solve_summary *s);
int vgo_problem_get_camera(const vgo_problem *p, int cam, double *out);
int vgo_problem_get_transform(const vgo_problem *p, int tr, double *out);
int vgo_problem_set_camera(vgo_problem *p, int cam, const double *value);
int vgo_problem_set_transform(vgo_problem *p, int tr, const double *values);
/* residuals of one dataset at the current parameters, n_img x 2P */
int vgo_problem_residuals(vgo_problem *p, int dataset, double *r);
/* one residual+Jacobian+normal-equation pass at the current parameters; cost = 1/2 sum r^2 */
int vgo_problem_evaluate(vgo_problem *p, int threads, double *cost);

#ifdef __cplusplus
}
#endif
#endif
