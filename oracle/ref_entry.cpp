// ref_entry.cpp -- C entry points around the REFERENCE's own hot-path classes (TEST
// INFRASTRUCTURE ONLY).  Compiled by `make -C oracle ref` together with
// $(REFERENCE)/src/calibration/calib_cost_functions.cpp against oracle/shim; the output
// oracle/_ref/libvisgeom_ref.so is git-ignored.  It drives GenericProjectionJac::Evaluate
// (calib_cost_functions.cpp:28-117) exactly as Ceres does: one functor per image, Evaluate
// with all Jacobian blocks requested.  No reference source is copied into this repository.
#include "calibration/calib_cost_functions.h"
#include "calibration/trajectory_generation.h"
#include "calibration/odometry_cost_function.h"
#include "projection/eucm.h"
#include "projection/ucm.h"
#include "projection/mei.h"

#include <cstring>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

static ICamera *make_camera(int model, const double *intr)
{
    switch (model) {
    case 0: return new EnhancedCamera(intr);   // unified_calibration.cpp:153
    case 1: return new UnifiedCamera(intr);    // :162
    case 2: return new MeiCamera(intr);        // :171
    default: return NULL;
    }
}

extern "C" {

int vgref_num_params(int model) { return model == 0 ? 6 : (model == 1 ? 5 : (model == 2 ? 10 : -1)); }

int vgref_evaluate_batch(int model, const double *intr, int n_img, int P,
                         const double *board, const double *obs,
                         int chain_len, const int *status, const int *is_global,
                         const double *const *xi,
                         double *r, double *J_intr, double *const *J_xi, double *H, int threads)
{
    const int K = vgref_num_params(model);
    if (K < 0 || chain_len < 1 || chain_len > 5) return -1;
    const int L = chain_len, D = K + 6 * L, W = D + 1, ne = W * (W + 1) / 2;
    Vector3dVec grid;
    for (int i = 0; i < P; i++) grid.emplace_back(board[3 * i], board[3 * i + 1], board[3 * i + 2]);
    vector<TransformationStatus> statusVec;
    for (int e = 0; e < L; e++) statusVec.push_back(status[e] ? TRANSFORM_INVERSE : TRANSFORM_DIRECT);
    std::unique_ptr<ICamera> cam(make_camera(model, intr));
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int img = 0; img < n_img; img++) {
        Vector2dVec proj;
        for (int i = 0; i < P; i++) proj.emplace_back(obs[((size_t)img * P + i) * 2], obs[((size_t)img * P + i) * 2 + 1]);
        // one cost function per image, as addGridResidualBlocks does (unified_calibration.cpp:532-533)
        GenericProjectionJac cost(proj, grid, cam.get(), statusVec);
        const double *params[6];
        double *jac[6];
        std::vector<double> scratch;
        size_t need = (r ? 0 : (size_t)2 * P) + (J_intr ? 0 : (size_t)2 * P * K);
        for (int e = 0; e < L; e++) need += (J_xi && J_xi[e]) ? 0 : (size_t)2 * P * 6;
        scratch.resize(need + 1);
        double *sp = scratch.data();
        params[0] = intr;
        double *rr = r ? r + (size_t)img * 2 * P : sp; if (!r) sp += (size_t)2 * P;
        jac[0] = J_intr ? J_intr + (size_t)img * 2 * P * K : sp; if (!J_intr) sp += (size_t)2 * P * K;
        for (int e = 0; e < L; e++) {
            params[1 + e] = xi[e] + (is_global[e] ? 0 : (size_t)img * 6);
            if (J_xi && J_xi[e]) jac[1 + e] = J_xi[e] + (size_t)img * 2 * P * 6;
            else { jac[1 + e] = sp; sp += (size_t)2 * P * 6; }
        }
        cost.Evaluate(params, rr, jac);
        if (H) {
            // the J^T J / J^T r product Ceres forms from the block (not reference code)
            double *h = H + (size_t)img * ne;
            for (int i = 0; i < ne; i++) h[i] = 0.0;
            double row[64];
            for (int k = 0; k < 2 * P; k++) {
                for (int c = 0; c < K; c++) row[c] = jac[0][(size_t)k * K + c];
                for (int e = 0; e < L; e++)
                    for (int c = 0; c < 6; c++) row[K + 6 * e + c] = jac[1 + e][(size_t)k * 6 + c];
                row[D] = rr[k];
                int idx = 0;
                for (int a = 0; a < W; a++)
                    for (int b = a; b < W; b++) h[idx++] += row[a] * row[b];
            }
        }
    }
    return 0;
}

// single-function probes used to pin the oracle's geometry / camera restatement
void vgref_rotation_matrix(const double *v, double *R)
{
    Matrix3d M = rotationMatrix<double>(Vector3d(v[0], v[1], v[2]));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = M(i, j);
}
void vgref_inter_omega_rot(const double *v, double *R)
{
    Matrix3d M = interOmegaRot<double>(Vector3d(v[0], v[1], v[2]));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = M(i, j);
}
void vgref_compose(const double *a, const double *b, double *out, int kind)
{
    Transformation<double> A(a), B(b), Cc;
    if (kind == 0) Cc = A.compose(B);
    else if (kind == 1) Cc = A.composeInverse(B);
    else Cc = A.inverseCompose(B);
    Cc.toArray(out);
}
int vgref_reconstruct(int model, const double *intr, const double *uv, double *X)
{
    std::unique_ptr<ICamera> cam(make_camera(model, intr));
    Vector3d Xv;
    bool ok = cam->reconstructPoint(Vector2d(uv[0], uv[1]), Xv);
    X[0] = Xv(0); X[1] = Xv(1); X[2] = Xv(2);
    return ok ? 1 : 0;
}
// the ICamera calls a visgeom user makes on a camera object (generic_camera.h:39-50): bit 0 projectPoint, bit 1
// projectionJacobian, bit 2 intrinsicJacobian returned true.  Outputs are zeroed first: the reference leaves them
// untouched on some failure paths.
int vgref_project_point(int model, const double *intr, const double *X, double *uv, double *dudx, double *dvdx,
                        double *dudalpha, double *dvdalpha)
{
    std::unique_ptr<ICamera> cam(make_camera(model, intr));
    const int K = cam->numParams();
    Vector3d Xv(X[0], X[1], X[2]);
    Vector2d p(0, 0);
    for (int i = 0; i < 3; i++) dudx[i] = dvdx[i] = 0;
    for (int i = 0; i < K; i++) dudalpha[i] = dvdalpha[i] = 0;
    int flags = 0;
    if (cam->projectPoint(Xv, p)) flags |= 1;
    uv[0] = p(0); uv[1] = p(1);
    if (cam->projectionJacobian(Xv, dudx, dvdx)) flags |= 2;
    if (cam->intrinsicJacobian(Xv, dudalpha, dvdalpha)) flags |= 4;
    return flags;
}
double vgref_bound(int model, const double *intr, int idx, int upper)
{
    std::unique_ptr<ICamera> cam(make_camera(model, intr));
    return upper ? cam->upperBound(idx) : cam->lowerBound(idx);
}

// the other residual types of the global problem (calib_cost_functions.h:64-108): constructed and evaluated
// exactly as addResiduals does (unified_calibration.cpp:795-803, 827-828)
void vgref_transformation_prior(const double *stiffness, const double *xi_prior, const double *xi, double *r, double *J)
{
    TransformationPrior cost(stiffness, xi_prior);
    const double *params[1] = {xi};
    double *jac[1] = {J};
    cost.Evaluate(params, r, J ? jac : NULL);
}
void vgref_odometry_prior(double errV, double errW, double lambda, const double *odom1, const double *odom2,
                          const double *xi1, const double *xi2, double *r, double *J1, double *J2)
{
    OdometryPrior cost(errV, errW, lambda, Transformation<double>(odom1), Transformation<double>(odom2));
    const double *params[2] = {xi1, xi2};
    double *jac[2] = {J1, J2};
    cost.Evaluate(params, r, (J1 || J2) ? jac : NULL);
}

// OdometryCost (odometry_cost_function.cpp:144-267): constructor from the prior intrinsics, then Evaluate with three
// parameter blocks, as unified_calibration.cpp:718-731 hands them over.  zeta_prior / A (optional) receive what the
// constructor left in the object.
void vgref_odometry_cost(double errV, double errW, double lambda, int m, const double *dq, const double *intr_prior,
                         const double *xi1, const double *xi2, const double *intr, double *r, double *J1, double *J2,
                         double *J3, double *zeta_prior, double *A)
{
    vector<Vector2d> deltaQ;
    for (int i = 0; i < m; i++) deltaQ.emplace_back(dq[2 * i], dq[2 * i + 1]);
    OdometryCost cost(errV, errW, lambda, deltaQ, intr_prior);
    const double *params[3] = {xi1, xi2, intr};
    double *jac[3] = {J1, J2, J3};
    cost.Evaluate(params, r, (J1 || J2 || J3) ? jac : NULL);
    if (zeta_prior) cost._zetaPrior.toArray(zeta_prior);
    if (A) for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) A[6 * i + j] = cost._A(i, j);
}

// TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206): the object is built through its own
// constructor from an in-memory property tree (the keys it reads, :85-118), no trajectories attached
static ptree leaf(double v) { ptree t; t.put_value(v); return t; }
static ptree vec(const double *v, int n) { ptree t; for (int i = 0; i < n; i++) t.add_child("", leaf(v[i])); return t; }

void vgref_visual_cov(int model, const double *intr, const double *xi_board, int nx, int ny, double step,
                      double feature_variance, int n, const double *cam_poses, double *out)
{
    const double zero6[6] = {0, 0, 0, 0, 0, 0}, one6[6] = {1, 1, 1, 1, 1, 1};
    ptree params, board;
    params.add_child("xiBaseCam", vec(zero6, 6));
    params.add_child("xiOrigBoard", vec(xi_board, 6));
    params.add_child("turn_radius", leaf(1.0));
    board.add_child("cols", leaf(nx)); board.add_child("rows", leaf(ny)); board.add_child("step", leaf(step));
    params.add_child("board", board);
    params.add_child("min_camera_dist", leaf(0.1));
    params.add_child("min_margin", leaf(10.0));
    params.add_child("min_cell_size", leaf(1.0));
    params.add_child("prior_variance", vec(one6, 6));
    params.add_child("feature_variance", leaf(feature_variance));
    std::unique_ptr<ICamera> cam(make_camera(model, intr));
    TrajectoryVisualQuality quality(vector<ITrajectory *>(), params, cam.get());
    for (int k = 0; k < n; k++) {
        Matrix6d C = quality.visualCov(Transformation<double>(cam_poses + 6 * k));
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) out[36 * k + 6 * i + j] = C(i, j);
    }
}

}  // extern "C"
