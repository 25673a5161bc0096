// vg_refine.cu -- N INDEPENDENT pose refinements on the GPU: the per-image initialisation solve of the calibration
// front end,
//   GenericCameraCalibration::estimateInitialGrid        src/calibration/unified_calibration.cpp:1131-1155
// which builds ONE ceres::Problem per image -- the board pose is the only free block (the camera is constant, :1145),
// the block is under SoftLOneLoss(25) (:1143), 500 iterations at most (:1148), Ceres' default tolerances -- and solves
// it with its own trust region, its own accept / reject decisions and its own termination.  Here: one warp per image
// runs that whole Levenberg-Marquardt loop in registers (the lanes take the board's corners; the 6 x 6 normal equations,
// the gradient and the squared norm are summed with shuffles in a fixed order), with no host synchronisation inside
// the loop and nothing shared between images: an image that converges late, or fails, cannot disturb another one's
// trust region -- which the batched solve this replaces (one radius, one accept / reject for the whole batch) could.
// Loop constants and decisions follow vg_problem_solve / Ceres' TrustRegionMinimizer (see vg_problem.cu's header).
#include "vg_common.h"
#include "vg_math.cuh"

#include <cstring>

namespace vg {
namespace {

struct RefineOpts {
    int max_iter, jacobi, max_invalid;
    double ftol, gtol, ptol, radius0, max_radius, min_radius, min_rel_dec, min_diag, max_diag, loss_b;
};

__device__ __forceinline__ constexpr int ut(int i, int j) { return i * 6 - i * (i - 1) / 2 + (j - i); }   // i <= j, 6 x 6 upper triangle

// [J r]^T [J r] pieces of one image at pose x: C (21, upper triangle of J^T J), b (J^T r), cost -- under the loss
template <int MODEL>
__device__ __forceinline__ void evaluate(const double (&intr)[Camera<MODEL>::K], const typename Camera<MODEL>::Consts &cc,
                                         const double (&x)[6], const int P, const double *__restrict__ board,
                                         const double *__restrict__ obs, const int lane, const double loss_b,
                                         double (&C)[21], double (&b)[6], double &cost)
{
    using CAM = Camera<MODEL>;
    double R[9], Jl[9];
    rodrigues_and_left_jacobian(x[3], x[4], x[5], R, Jl);
#pragma unroll
    for (int i = 0; i < 21; i++) C[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) b[i] = 0.0;
    double s = 0.0;
    for (int c = lane; c < P; c += 32) {
        const double bx = board[3 * c], by = board[3 * c + 1], bz = board[3 * c + 2];
        const double w0 = fma(R[2], bz, fma(R[1], by, R[0] * bx)), w1 = fma(R[5], bz, fma(R[4], by, R[3] * bx)),
                     w2 = fma(R[8], bz, fma(R[7], by, R[6] * bx));
        double u, v, Pu[3], Pv[3], Ju[CAM::K], Jv[CAM::K];
        const bool ok = CAM::eval(intr, cc, w0 + x[0], w1 + x[1], w2 + x[2], u, v, Pu, Pv, Ju, Jv);
        double ru = DOUBLE_BIG, rv = DOUBLE_BIG, ju[6] = {0, 0, 0, 0, 0, 0}, jv[6] = {0, 0, 0, 0, 0, 0};
        if (ok) {       // (a failed projection: the 1e15 sentinel and zero Jacobian rows, calib_cost_functions.cpp:66-70)
            ru = u - obs[2 * c]; rv = v - obs[2 * c + 1];
            const double cu0 = w1 * Pu[2] - w2 * Pu[1], cu1 = w2 * Pu[0] - w0 * Pu[2], cu2 = w0 * Pu[1] - w1 * Pu[0];
            const double cv0 = w1 * Pv[2] - w2 * Pv[1], cv1 = w2 * Pv[0] - w0 * Pv[2], cv2 = w0 * Pv[1] - w1 * Pv[0];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                ju[q] = Pu[q]; jv[q] = Pv[q];                       // d/dt: R12 = I for a single DIRECT element
                ju[3 + q] = fma(cu2, Jl[6 + q], fma(cu1, Jl[3 + q], cu0 * Jl[q]));      // (w x p)^T M12, M12 = J_l(r)
                jv[3 + q] = fma(cv2, Jl[6 + q], fma(cv1, Jl[3 + q], cv0 * Jl[q]));
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
#pragma unroll
            for (int j = i; j < 6; j++) C[ut(i, j)] = fma(ju[i], ju[j], fma(jv[i], jv[j], C[ut(i, j)]));
            b[i] = fma(ju[i], ru, fma(jv[i], rv, b[i]));
        }
        s = fma(ru, ru, fma(rv, rv, s));
    }
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 21; i++) C[i] += __shfl_xor_sync(0xffffffffu, C[i], off);
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] += __shfl_xor_sync(0xffffffffu, b[i], off);
        s += __shfl_xor_sync(0xffffffffu, s, off);
    }
    if (loss_b > 0.0) {
        // SoftLOneLoss(a), b = a^2: rho(s) = 2 b (sqrt(1 + s / b) - 1); rho'' < 0, so Ceres' corrector scales the block's
        // residuals and Jacobians by sqrt(rho') and the block costs rho / 2
        const double q = sqrt(1.0 + s / loss_b), w = 1.0 / q;
#pragma unroll
        for (int i = 0; i < 21; i++) C[i] *= w;
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] *= w;
        cost = loss_b * (q - 1.0);
    } else {
        cost = 0.5 * s;
    }
}

template <int MODEL>
__global__ void __launch_bounds__(128)
refine_poses_kernel(const double *__restrict__ intr_g, const int n_img, const int P, const double *__restrict__ board,
                    const double *__restrict__ obs_all, double *__restrict__ poses, const RefineOpts o,
                    int *__restrict__ iterations, double *__restrict__ final_cost, int *__restrict__ termination)
{
    using CAM = Camera<MODEL>;
    const int lane = threadIdx.x & 31;
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= n_img) return;
    double intr[CAM::K];
#pragma unroll
    for (int i = 0; i < CAM::K; i++) intr[i] = __ldg(intr_g + i);
    const typename CAM::Consts cc = CAM::prepare(intr);
    const double *obs = obs_all + (size_t)img * 2 * P;
    double x[6];
#pragma unroll
    for (int k = 0; k < 6; k++) x[k] = poses[(size_t)img * 6 + k];
    double C[21], b[6], cost;
    evaluate<MODEL>(intr, cc, x, P, board, obs, lane, o.loss_b, C, b, cost);
    double radius = o.radius0, dec = 2.0, scale[6] = {1, 1, 1, 1, 1, 1};
    int iter = 0, invalid_run = 0, term = 3;
    bool init_scale = true;
    for (;;) {
        double gmax = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) gmax = fmax(gmax, fabs(b[k]));
        if (gmax <= o.gtol) { term = 1; break; }
        if (iter >= o.max_iter) { term = 3; break; }
        if (radius < o.min_radius) { term = 4; break; }
        iter++;
        if (init_scale) {
#pragma unroll
            for (int k = 0; k < 6; k++) scale[k] = o.jacobi ? 1.0 / (1.0 + sqrt(C[ut(k, k)])) : 1.0;
            init_scale = false;
        }
        // damped factorisation (lower triangle in Lm), step d = -(C + D)^-1 b
        double Lm[21], d[6];
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) Lm[i * (i + 1) / 2 + j] = C[ut(j, i)];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double s2 = scale[k] * scale[k];
            Lm[k * (k + 1) / 2 + k] += fmin(fmax(s2 * C[ut(k, k)], o.min_diag), o.max_diag) / (radius * s2);
        }
#pragma unroll
        for (int j = 0; j < 6; j++) {
            double s = Lm[j * (j + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s = fma(-Lm[j * (j + 1) / 2 + k], Lm[j * (j + 1) / 2 + k], s);
            if (!(s > 0.0)) { ok = false; s = 1.0; }
            s = sqrt(s);
            Lm[j * (j + 1) / 2 + j] = s;
#pragma unroll
            for (int i = j + 1; i < 6; i++) {
                double t = Lm[i * (i + 1) / 2 + j];
#pragma unroll
                for (int k = 0; k < j; k++) t = fma(-Lm[i * (i + 1) / 2 + k], Lm[j * (j + 1) / 2 + k], t);
                Lm[i * (i + 1) / 2 + j] = t / s;
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double s = -b[i];
#pragma unroll
            for (int k = 0; k < i; k++) s = fma(-Lm[i * (i + 1) / 2 + k], d[k], s);
            d[i] = s / Lm[i * (i + 1) / 2 + i];
        }
#pragma unroll
        for (int i = 5; i >= 0; i--) {
            double s = d[i];
#pragma unroll
            for (int k = i + 1; k < 6; k++) s = fma(-Lm[k * (k + 1) / 2 + i], d[k], s);
            d[i] = s / Lm[i * (i + 1) / 2 + i];
        }
        double gd = 0.0, dHd = 0.0, step2 = 0.0, x2 = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 6; j++) t = fma(C[i <= j ? ut(i, j) : ut(j, i)], d[j], t);
            gd = fma(b[i], d[i], gd);
            dHd = fma(d[i], t, dHd);
            step2 = fma(d[i], d[i], step2);
            x2 = fma(x[i], x[i], x2);
        }
        const double model_change = -gd - 0.5 * dHd;
        if (!ok || !(model_change > 0.0)) {
            invalid_run++;
            if (invalid_run >= o.max_invalid) { term = 5; break; }
            radius /= dec; dec *= 2.0;
            continue;
        }
        invalid_run = 0;
        if (sqrt(step2) <= o.ptol * (sqrt(x2) + o.ptol)) { term = 2; break; }     // the candidate is discarded
        double xc[6], Cc[21], bc[6], new_cost;
#pragma unroll
        for (int k = 0; k < 6; k++) xc[k] = x[k] + d[k];
        evaluate<MODEL>(intr, cc, xc, P, board, obs, lane, o.loss_b, Cc, bc, new_cost);
        const double rho = (cost - new_cost) / model_change;
        if (rho > o.min_rel_dec) {
            const double cost_change = cost - new_cost, old_cost = cost;
#pragma unroll
            for (int k = 0; k < 6; k++) { x[k] = xc[k]; b[k] = bc[k]; }
#pragma unroll
            for (int k = 0; k < 21; k++) C[k] = Cc[k];
            cost = new_cost;
            const double t = 2.0 * rho - 1.0;
            radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
            if (radius > o.max_radius) radius = o.max_radius;
            dec = 2.0;
            if (fabs(cost_change) <= o.ftol * old_cost) { term = 0; break; }
        } else {
            radius /= dec; dec *= 2.0;
        }
    }
    if (lane < 6) poses[(size_t)img * 6 + lane] = x[0] * (lane == 0) + x[1] * (lane == 1) + x[2] * (lane == 2) + x[3] * (lane == 3) +
                                                  x[4] * (lane == 4) + x[5] * (lane == 5);
    if (lane == 0) {
        if (iterations) iterations[img] = iter;
        if (final_cost) final_cost[img] = cost;
        if (termination) termination[img] = term;
    }
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" int vg_refine_poses(int model, const double *intr, int n_img, int P, const double *board, const double *obs,
                               double *poses, double loss_a, const vg_solve_options *opt, int *iterations,
                               double *final_cost, int *termination)
{
    const int K = vg_model_num_params(model);
    if (K < 0) return K;
    if (n_img < 0 || P < 1 || !intr || !board || (n_img > 0 && (!obs || !poses)) || !(loss_a >= 0.0))
        return fail(VG_ERR_INVALID, "vg_refine_poses: bad arguments");
    if (vg_device_count() < 1) return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    if (n_img == 0) return VG_OK;
    vg_solve_options so;
    if (opt) so = *opt; else vg_solve_options_default(&so);
    RefineOpts o;
    o.max_iter = so.max_num_iterations; o.jacobi = so.jacobi_scaling; o.max_invalid = so.max_consecutive_invalid;
    o.ftol = so.function_tolerance; o.gtol = so.gradient_tolerance; o.ptol = so.parameter_tolerance;
    o.radius0 = so.initial_radius; o.max_radius = so.max_radius; o.min_radius = so.min_radius;
    o.min_rel_dec = so.min_relative_decrease; o.min_diag = so.min_lm_diagonal; o.max_diag = so.max_lm_diagonal;
    o.loss_b = loss_a * loss_a;
    const size_t nn = (size_t)n_img;
    const size_t b_intr = 0, b_board = 128, b_obs = b_board + (((size_t)P * 24 + 255) & ~size_t(255)),
                 b_pose = b_obs + ((nn * 2 * P * 8 + 255) & ~size_t(255)), b_cost = b_pose + ((nn * 48 + 255) & ~size_t(255)),
                 b_iter = b_cost + ((nn * 8 + 255) & ~size_t(255)), b_term = b_iter + ((nn * 4 + 255) & ~size_t(255)),
                 total = b_term + ((nn * 4 + 255) & ~size_t(255));
    char *d = nullptr;
    VG_CUDA(cudaMalloc(&d, total));
    cudaError_t e = cudaMemcpy(d + b_intr, intr, 8 * K, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + b_board, board, (size_t)P * 24, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + b_obs, obs, nn * 2 * P * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + b_pose, poses, nn * 48, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const int blocks = (n_img + 3) / 4;
        auto D = [&](size_t off) { return reinterpret_cast<double *>(d + off); };
        int *it = reinterpret_cast<int *>(d + b_iter), *tm = reinterpret_cast<int *>(d + b_term);
        if (model == VG_MODEL_EUCM) refine_poses_kernel<MODEL_EUCM><<<blocks, 128>>>(D(b_intr), n_img, P, D(b_board), D(b_obs), D(b_pose), o, it, D(b_cost), tm);
        else if (model == VG_MODEL_UCM) refine_poses_kernel<MODEL_UCM><<<blocks, 128>>>(D(b_intr), n_img, P, D(b_board), D(b_obs), D(b_pose), o, it, D(b_cost), tm);
        else refine_poses_kernel<MODEL_MEI><<<blocks, 128>>>(D(b_intr), n_img, P, D(b_board), D(b_obs), D(b_pose), o, it, D(b_cost), tm);
        count_launch(&launch_counter());
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(poses, d + b_pose, nn * 48, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && final_cost) e = cudaMemcpy(final_cost, d + b_cost, nn * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && iterations) e = cudaMemcpy(iterations, d + b_iter, nn * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && termination) e = cudaMemcpy(termination, d + b_term, nn * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_refine_poses");
}
