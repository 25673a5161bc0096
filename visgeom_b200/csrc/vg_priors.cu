// vg_priors.cu -- TransformationPrior / OdometryPrior blocks and the block-tridiagonal pose elimination
// (see vg_priors.cuh).  Definitions followed (reference paths relative to /root/reference):
//   Transformation::inverseCompose     include/geometry/transformation.h:90-99
//   Quaternion (rot-vec <-> quat)      include/geometry/quaternion.h:31-50,84-98
//   screwTransfInv                     include/geometry/transformation.h:234-243
//   TransformationPrior                include/calibration/calib_cost_functions.h:83-108, src/calibration/calib_cost_functions.cpp:215-228
//   OdometryPrior                      src/calibration/calib_cost_functions.cpp:119-213
#include "vg_common.h"
#include "vg_math.cuh"
#include "vg_priors.cuh"

#include <cmath>

namespace vg {
namespace {

__device__ __forceinline__ constexpr int lt(int i, int j) { return i * (i + 1) / 2 + j; }   // packed lower, i >= j

struct Quat { double x, y, z, w; };

// quaternion.h:31-50: the small-angle branch is taken below 1e-6
__device__ __forceinline__ Quat quat_from_rotvec(const double *r)
{
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    Quat q;
    if (th < 1e-6) { q.x = 0.5 * r[0]; q.y = 0.5 * r[1]; q.z = 0.5 * r[2]; q.w = 1.0; return q; }
    double s, c;
    sincos(0.5 * th, &s, &c);
    const double k = s / th;
    q.x = r[0] * k; q.y = r[1] * k; q.z = r[2] * k; q.w = c;
    return q;
}

// quaternion.h:84-98 with normalizeAngle (geometry_core.h:32-38): the angle is wrapped into (-pi, pi]
__device__ __forceinline__ void quat_to_rotvec(const Quat &q, double *r)
{
    const double s = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (s < 1e-5) { r[0] = 2.0 * q.x; r[1] = 2.0 * q.y; r[2] = 2.0 * q.z; return; }
    double th = 2.0 * atan2(s, q.w);
    constexpr double PI = 3.14159265358979323846;
    if (th > PI) th -= 2.0 * PI;
    else if (th < -PI) th += 2.0 * PI;
    const double k = th / s;
    r[0] = q.x * k; r[1] = q.y * k; r[2] = q.z * k;
}

__device__ __forceinline__ Quat quat_mul(const Quat &a, const Quat &b)
{
    Quat o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return o;
}

// v' = v + 2 w (u x v) + 2 u x (u x v)
__device__ __forceinline__ void quat_rotate(const Quat &q, const double *v, double *o)
{
    const double c0 = q.y * v[2] - q.z * v[1], c1 = q.z * v[0] - q.x * v[2], c2 = q.x * v[1] - q.y * v[0];
    const double d0 = q.y * c2 - q.z * c1, d1 = q.z * c0 - q.x * c2, d2 = q.x * c1 - q.y * c0;
    o[0] = v[0] + 2.0 * (q.w * c0 + d0);
    o[1] = v[1] + 2.0 * (q.w * c1 + d1);
    o[2] = v[2] + 2.0 * (q.w * c2 + d2);
}

// out = a^-1 o b as [t, r]
__device__ __forceinline__ void se3_inverse_compose(const double *a, const double *b, double *out)
{
    Quat qa = quat_from_rotvec(a + 3);
    const Quat qb = quat_from_rotvec(b + 3);
    qa.x = -qa.x; qa.y = -qa.y; qa.z = -qa.z;
    const double d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    quat_rotate(qa, d, out);
    quat_to_rotvec(quat_mul(qa, qb), out + 3);
}

__device__ __forceinline__ void forward6(const double (&Lm)[21], const double (&invd)[6], double (&x)[6])
{
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < i; k++) s = fma(-Lm[lt(i, k)], x[k], s);
        x[i] = s * invd[i];
    }
}

// ---- TransformationPrior ---------------------------------------------------------------------------------
__device__ void tp_setup(const double *stiffness, const double *xi_prior, double *rec)
{
    double R[9], M[9];
    rodrigues_and_left_jacobian(xi_prior[3], xi_prior[4], xi_prior[5], R, M);
    double A[36];
    for (int i = 0; i < 36; i++) A[i] = 0.0;
    for (int i = 0; i < 3; i++) {
        A[6 * i + i] = stiffness[i];
        for (int j = 0; j < 3; j++) A[6 * (i + 3) + j + 3] = stiffness[i + 3] * M[3 * i + j];   // diag * M, .h:95-96
    }
    for (int k = 0; k < 6; k++) rec[k] = xi_prior[k];
    for (int k = 0; k < 36; k++) rec[TP_OFF_A + k] = A[k];
    for (int k = 0; k < 9; k++) rec[TP_OFF_R + k] = R[k];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s = fma(A[6 * k + i], A[6 * k + j], s);
            rec[TP_OFF_AtA + lt(i, j)] = s;
        }
}

__device__ void tp_eval(const double *rec, const double *xi, double (&r)[6])
{
    double e[6], err[6];
    se3_inverse_compose(rec, xi, e);
    const double *R = rec + TP_OFF_R, *A = rec + TP_OFF_A;
    for (int i = 0; i < 3; i++) {
        err[i] = R[3 * i] * e[0] + R[3 * i + 1] * e[1] + R[3 * i + 2] * e[2];
        err[i + 3] = R[3 * i] * e[3] + R[3 * i + 1] * e[4] + R[3 * i + 2] * e[5];
    }
    for (int i = 0; i < 6; i++) {
        double s = 0.0;
        for (int k = 0; k < 6; k++) s = fma(A[6 * i + k], err[k], s);
        r[i] = s;
    }
}

// ---- OdometryPrior -----------------------------------------------------------------------------------------
__device__ void op_information(double errV, double errW, double lambda, const double *zp, double *rec);

__device__ void op_setup(double errV, double errW, double lambda, const double *o1, const double *o2, double *rec)
{
    double zp[6];
    se3_inverse_compose(o1, o2, zp);
    op_information(errV, errW, lambda, zp, rec);
}

// rec = [zeta_prior | A]: the information matrix from the prior motion (calib_cost_functions.cpp:124-172 and, word for
// word the same, odometry_cost_function.cpp:155-196)
__device__ void op_information(double errV, double errW, double lambda, const double *zp, double *rec)
{
    const double delta = fmax(sqrt(zp[3] * zp[3] + zp[4] * zp[4] + zp[5] * zp[5]), 0.01);
    const double l = fmax(sqrt(zp[0] * zp[0] + zp[1] * zp[1] + zp[2] * zp[2]), 0.01);
    double s, c;
    sincos(0.5 * delta, &s, &c);
    const double l2 = 0.5 * l;
    const double F[3][2] = {{c, l2 * s}, {-s, l2 * c}, {0.0, 1.0}};
    const double cu0 = fmax(errV * errV * l * l, 1e-4), cu1 = fmax(errW * errW * delta * delta, 1e-4);
    double Cx[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            Cx[3 * i + j] = F[i][0] * cu0 * F[j][0] + F[i][1] * cu1 * F[j][1] + (i == j ? lambda * lambda : 0.0);
    // inverse of the symmetric 3x3 by cofactors, then its upper Cholesky factor U (U^T U = Cx^-1)
    double inv[9];
    {
        const double c00 = Cx[4] * Cx[8] - Cx[5] * Cx[7], c01 = Cx[5] * Cx[6] - Cx[3] * Cx[8], c02 = Cx[3] * Cx[7] - Cx[4] * Cx[6];
        const double id = 1.0 / (Cx[0] * c00 + Cx[1] * c01 + Cx[2] * c02);
        inv[0] = c00 * id; inv[1] = (Cx[2] * Cx[7] - Cx[1] * Cx[8]) * id; inv[2] = (Cx[1] * Cx[5] - Cx[2] * Cx[4]) * id;
        inv[3] = c01 * id; inv[4] = (Cx[0] * Cx[8] - Cx[2] * Cx[6]) * id; inv[5] = (Cx[2] * Cx[3] - Cx[0] * Cx[5]) * id;
        inv[6] = c02 * id; inv[7] = (Cx[1] * Cx[6] - Cx[0] * Cx[7]) * id; inv[8] = (Cx[0] * Cx[4] - Cx[1] * Cx[3]) * id;
    }
    const double l00 = sqrt(inv[0]);
    const double l10 = inv[3] / l00, l20 = inv[6] / l00;
    const double l11 = sqrt(inv[4] - l10 * l10);
    const double l21 = (inv[7] - l20 * l10) / l11;
    const double l22 = sqrt(inv[8] - l20 * l20 - l21 * l21);
    double A[36];
    for (int i = 0; i < 36; i++) A[i] = 0.0;
    A[0] = l00; A[1] = l10;          // U(0,0), U(0,1)
    A[7] = l11;                      // U(1,1); U(1,0) = 0
    A[5] = l20; A[11] = l21;         // topRightCorner<2,1>() of the 6x6 is column 5 (.cpp:158)
    A[14] = 1.0 / lambda;
    A[21] = 1.0 / lambda; A[28] = 1.0 / lambda; A[35] = l22;
    for (int k = 0; k < 6; k++) rec[k] = zp[k];
    for (int k = 0; k < 36; k++) rec[6 + k] = A[k];
}

// r, J1 = d r / d xi1, J2 = d r / d xi2 (row-major 6x6) as the functor reports them
__device__ void op_eval(const double *rec, const double *x1, const double *x2, double (&r)[6], double (&J1)[36], double (&J2)[36])
{
    const double *A = rec + 6;
    double zeta[6], err[6];
    se3_inverse_compose(x1, x2, zeta);
    se3_inverse_compose(rec, zeta, err);
    for (int i = 0; i < 6; i++) {
        double s = 0.0;
        for (int k = 0; k < 6; k++) s = fma(A[6 * i + k], err[k], s);
        r[i] = s;
    }
    double R1[9], M1[9], R2[9], M2[9], Rz[9], Mz[9];
    rodrigues_and_left_jacobian(x1[3], x1[4], x1[5], R1, M1);
    rodrigues_and_left_jacobian(x2[3], x2[4], x2[5], R2, M2);
    rodrigues_and_left_jacobian(zeta[3], zeta[4], zeta[5], Rz, Mz);
    // B = screwTransfInv(zeta) * blockdiag(R10, R10 M1), R10 = R1^T, Rz^-1 = Rz^T
    double R10M[9], RzR10[9], RzR10M[9], T[9], TR[9], B[36];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0, q = 0.0;
            for (int k = 0; k < 3; k++) { s = fma(R1[3 * k + i], M1[3 * k + j], s); q = fma(Rz[3 * k + i], R1[3 * j + k], q); }
            R10M[3 * i + j] = s; RzR10[3 * i + j] = q;
        }
    // T = -Rz^T hat(t_zeta)
    const double h[9] = {0.0, -zeta[2], zeta[1], zeta[2], 0.0, -zeta[0], -zeta[1], zeta[0], 0.0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s = fma(-Rz[3 * k + i], h[3 * k + j], s);
            T[3 * i + j] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0, q = 0.0;
            for (int k = 0; k < 3; k++) { s = fma(Rz[3 * k + i], R10M[3 * k + j], s); q = fma(T[3 * i + k], R10M[3 * k + j], q); }
            RzR10M[3 * i + j] = s; TR[3 * i + j] = q;
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            B[6 * i + j] = RzR10[3 * i + j];
            B[6 * i + j + 3] = TR[3 * i + j];
            B[6 * (i + 3) + j] = 0.0;
            B[6 * (i + 3) + j + 3] = RzR10M[3 * i + j];
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s = fma(A[6 * i + k], B[6 * k + j], s);
            J1[6 * i + j] = -s;
        }
    // J2 = A * blockdiag(R20, R20 M2)
    double R20M[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s = fma(R2[3 * k + i], M2[3 * k + j], s);
            R20M[3 * i + j] = s;
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0, q = 0.0;
            for (int k = 0; k < 3; k++) { s = fma(A[6 * i + k], R2[3 * j + k], s); q = fma(A[6 * i + 3 + k], R20M[3 * k + j], q); }
            J2[6 * i + j] = s; J2[6 * i + j + 3] = q;
        }
}

__global__ void prior_setup_kernel(int n_tp, const double *tp_in, double *tp_const, int n_op, const double *op_in, double *op_const)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tp) tp_setup(tp_in + 12 * (size_t)i, tp_in + 12 * (size_t)i + 6, tp_const + (size_t)TP_CONST * i);
    else if (i < n_tp + n_op) {
        const double *in = op_in + 15 * (size_t)(i - n_tp);
        op_setup(in[0], in[1], in[2], in + 3, in + 9, op_const + (size_t)OP_CONST * (i - n_tp));
    }
}

// problem level: the normal-equation pieces of every prior block at one parameter set
__global__ void __launch_bounds__(64) prior_eval_kernel(PriorTables t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t.n_tp) {
        const double *rec = t.tp_const + (size_t)TP_CONST * i;
        double r[6];
        tp_eval(rec, t.tp_xi[i], r);
        double *out = t.tp_out + (size_t)TP_OUT * i;
        double cost = 0.0;
        for (int a = 0; a < 6; a++) {
            double g = 0.0;
            for (int k = 0; k < 6; k++) g = fma(rec[TP_OFF_A + 6 * k + a], r[k], g);   // the Jacobian is A itself (.cpp:224-227)
            out[a] = g;
            cost = fma(0.5 * r[a], r[a], cost);
        }
        out[6] = cost;
    } else if (i < t.n_tp + t.n_op) {
        const int e = i - t.n_tp;
        const double *x1 = t.op_xi[e];
        double r[6], J1[36], J2[36];
        op_eval(t.op_const + (size_t)OP_CONST * e, x1, x1 + 6, r, J1, J2);
        double *out = t.op_out + (size_t)OP_OUT * e;
        double cost = 0.0;
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) {
                double h11 = 0.0, h22 = 0.0, h21 = 0.0;
                for (int k = 0; k < 6; k++) {
                    h11 = fma(J1[6 * k + a], J1[6 * k + b], h11);
                    h22 = fma(J2[6 * k + a], J2[6 * k + b], h22);
                    h21 = fma(J2[6 * k + a], J1[6 * k + b], h21);
                }
                if (b <= a) { out[OP_OFF_H11 + lt(a, b)] = h11; out[OP_OFF_H22 + lt(a, b)] = h22; }
                out[OP_OFF_O + 6 * a + b] = h21;
            }
            double g1 = 0.0, g2 = 0.0;
            for (int k = 0; k < 6; k++) { g1 = fma(J1[6 * k + a], r[k], g1); g2 = fma(J2[6 * k + a], r[k], g2); }
            out[OP_OFF_G1 + a] = g1; out[OP_OFF_G2 + a] = g2;
            cost = fma(0.5 * r[a], r[a], cost);
        }
        out[OP_OFF_COST] = cost;
    }
}

// one block: costs summed in a fixed order, shared-block pieces added entry by entry
__global__ void __launch_bounds__(256) prior_reduce_kernel(PriorTables t, int Ks, double *red, int add_shared)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < t.n_tp + t.n_op; i += blockDim.x)
        s += i < t.n_tp ? t.tp_out[(size_t)TP_OUT * i + 6] : t.op_out[(size_t)OP_OUT * (i - t.n_tp) + OP_OFF_COST];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) red[red_off_cost(Ks)] += sh[0];
    if (add_shared && threadIdx.x < 42) {
        // thread = one entry of the 6x6 block (36) or of the gradient (6); priors on the same transform add in order
        for (int k = 0; k < t.n_tp_shared; k++) {
            const int rec = t.tp_shared_rec[k], off = t.tp_shared_off[k];
            if (threadIdx.x < 36) {
                const int a = threadIdx.x / 6, b = threadIdx.x % 6;
                red[red_off_A(Ks) + (off + a) * Ks + off + b] += t.tp_const[(size_t)TP_CONST * rec + TP_OFF_AtA + (a >= b ? lt(a, b) : lt(b, a))];
            } else {
                const int a = threadIdx.x - 36;
                red[red_off_g(Ks) + off + a] += t.tp_out[(size_t)TP_OUT * rec + a];
            }
        }
    }
}

// inner level (the functors' own contract, batched): residuals and Jacobians written out
__global__ void tp_functor_kernel(int n, const double *stiffness, const double *xi_prior, const double *xi, double *r_out, double *J_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double rec[TP_CONST], r[6];
    tp_setup(stiffness + 6 * (size_t)i, xi_prior + 6 * (size_t)i, rec);
    tp_eval(rec, xi + 6 * (size_t)i, r);
    for (int k = 0; k < 6; k++) r_out[6 * (size_t)i + k] = r[k];
    if (J_out)
        for (int k = 0; k < 36; k++) J_out[36 * (size_t)i + k] = rec[TP_OFF_A + k];
}

__global__ void op_functor_kernel(int n, double errV, double errW, double lambda, const double *o1, const double *o2,
                                  const double *x1, const double *x2, double *r_out, double *J1_out, double *J2_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double rec[OP_CONST], r[6], J1[36], J2[36];
    op_setup(errV, errW, lambda, o1 + 6 * (size_t)i, o2 + 6 * (size_t)i, rec);
    op_eval(rec, x1 + 6 * (size_t)i, x2 + 6 * (size_t)i, r, J1, J2);
    for (int k = 0; k < 6; k++) r_out[6 * (size_t)i + k] = r[k];
    if (J1_out)
        for (int k = 0; k < 36; k++) J1_out[36 * (size_t)i + k] = J1[k];
    if (J2_out)
        for (int k = 0; k < 36; k++) J2_out[36 * (size_t)i + k] = J2[k];
}


// ---- OdometryCost (src/calibration/odometry_cost_function.cpp; "odometry_intrinsic" blocks) ---------------------------
// The motion between two poses of a differential-drive platform integrated from m pairs of wheel-angle increments with
// the odometry intrinsics (r1, r2 wheel radii, g track gauge), held against xi1^-1 o xi2.

// out = a o b as [t, r] (transformation.h:80-88: through quaternions)
__device__ __forceinline__ void se3_compose(const double *a, const double *b, double *out)
{
    const Quat qa = quat_from_rotvec(a + 3), qb = quat_from_rotvec(b + 3);
    double t[3];
    quat_rotate(qa, b, t);
    out[0] = t[0] + a[0]; out[1] = t[1] + a[1]; out[2] = t[2] + a[2];
    quat_to_rotvec(quat_mul(qa, qb), out + 3);
}

// one increment (odom_zeta_i :10-35, zeta_i_jacobian :38-65): the planar step (v, 0, w) and d(v, -, w) / d(r1, r2, g)
__device__ __forceinline__ void oc_step(const double *dq, const double r1, const double r2, const double g, double *step, double *jz)
{
    step[0] = (r1 / 2) * dq[0] + (r2 / 2) * dq[1];
    step[1] = 0.0; step[2] = 0.0; step[3] = 0.0; step[4] = 0.0;
    step[5] = -(r1 / g) * dq[0] + (r2 / g) * dq[1];
    jz[0] = dq[0] / 2; jz[1] = dq[1] / 2; jz[2] = 0.0;
    jz[3] = 0.0; jz[4] = 0.0; jz[5] = 0.0;
    jz[6] = -dq[0] / g; jz[7] = dq[1] / g; jz[8] = (r1 * dq[0] - r2 * dq[1]) / (g * g);
}

// tf0n_jac_calc (:68-88): zeta_odo = the m steps composed
__device__ void oc_integrate(const int m, const double *dq, const double *in, double *zeta_odo)
{
    double cur[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < m; i++) {
        double step[6], jz[9], nxt[6];
        oc_step(dq + 2 * i, in[0], in[1], in[2], step, jz);
        se3_compose(cur, step, nxt);
        for (int k = 0; k < 6; k++) cur[k] = nxt[k];
    }
    for (int k = 0; k < 6; k++) zeta_odo[k] = cur[k];
}

// calc_acc (:90-141): ACC = sum_i R(0T(i-1)) [1 0 -y; 0 1 x; 0 0 1] d zeta_i / d(r1, r2, g), (x, y) the translation of
// iTn = (0Ti)^-1 o 0Tn.  The reference keeps every 0Ti in a vector; here the chain is walked a second time.
__device__ void oc_accumulate(const int m, const double *dq, const double *in, const double *zeta_odo, double *acc)
{
    for (int k = 0; k < 9; k++) acc[k] = 0.0;
    double prev[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < m; i++) {
        double step[6], jz[9], cur[6], R0[9], Mu[9], tin[6];
        oc_step(dq + 2 * i, in[0], in[1], in[2], step, jz);
        se3_compose(prev, step, cur);
        rodrigues_and_left_jacobian(prev[3], prev[4], prev[5], R0, Mu);
        // (0Ti)^-1 = (-R^T t, -r) (transformation.h:112-119), then composed with 0Tn
        double Ri[9], inv[6];
        rodrigues_and_left_jacobian(-cur[3], -cur[4], -cur[5], Ri, Mu);
        for (int a = 0; a < 3; a++) inv[a] = -(Ri[3 * a] * cur[0] + Ri[3 * a + 1] * cur[1] + Ri[3 * a + 2] * cur[2]);
        inv[3] = -cur[3]; inv[4] = -cur[4]; inv[5] = -cur[5];
        se3_compose(inv, zeta_odo, tin);
        const double J[9] = {1, 0, -tin[1], 0, 1, tin[0], 0, 0, 1};
        double RJ[9];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) RJ[3 * a + b] = R0[3 * a] * J[b] + R0[3 * a + 1] * J[3 + b] + R0[3 * a + 2] * J[6 + b];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) acc[3 * a + b] += RJ[3 * a] * jz[b] + RJ[3 * a + 1] * jz[3 + b] + RJ[3 * a + 2] * jz[6 + b];
        for (int k = 0; k < 6; k++) prev[k] = cur[k];
    }
}

// OdometryCost as a functor: one thread per block.  The first two Jacobian blocks are OdometryPrior's with the
// integrated motion in the prior's place (:231-253); the third (:256-264) is
// -A screwTransfInv(delta) blockdiag(R31, R31 M(zeta_odo)) [ACC rows x, y | 0 | ACC row theta].
__global__ void oc_functor_kernel(int n, double errV, double errW, double lambda, const int *__restrict__ dq_offset,
                                  const double *__restrict__ dq, const double *__restrict__ intr_prior, const double *x1,
                                  const double *x2, const double *__restrict__ intr, double *r_out, double *J1_out, double *J2_out,
                                  double *J3_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = dq_offset[i + 1] - dq_offset[i];
    const double *q = dq + 2 * (size_t)dq_offset[i];
    double rec[OP_CONST], zp[6], zo[6], r[6], J1[36], J2[36];
    oc_integrate(m, q, intr_prior, zp);                     // the constructor: prior motion -> A
    op_information(errV, errW, lambda, zp, rec);
    oc_integrate(m, q, intr, zo);
    for (int k = 0; k < 6; k++) rec[k] = zo[k];
    op_eval(rec, x1 + 6 * (size_t)i, x2 + 6 * (size_t)i, r, J1, J2);
    for (int k = 0; k < 6; k++) r_out[6 * (size_t)i + k] = r[k];
    if (J1_out)
        for (int k = 0; k < 36; k++) J1_out[36 * (size_t)i + k] = J1[k];
    if (J2_out)
        for (int k = 0; k < 36; k++) J2_out[36 * (size_t)i + k] = J2[k];
    if (!J3_out) return;
    const double *A = rec + 6;
    double acc[9], zeta[6], delta[6];
    oc_accumulate(m, q, intr, zo, acc);
    se3_inverse_compose(x1 + 6 * (size_t)i, x2 + 6 * (size_t)i, zeta);
    se3_inverse_compose(zo, zeta, delta);
    double R31[9], M3[9], Rd[9], Md[9], Mu[9];
    rodrigues_and_left_jacobian(-zo[3], -zo[4], -zo[5], R31, Mu);
    rodrigues_and_left_jacobian(zo[3], zo[4], zo[5], Mu, M3);
    rodrigues_and_left_jacobian(-delta[3], -delta[4], -delta[5], Rd, Md);
    // C = blockdiag(R31, R31 M3) * [ACC(0,:); ACC(1,:); 0; 0; 0; ACC(2,:)]   (6 x 3)
    double R31M[9], Cm[18];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) R31M[3 * a + b] = R31[3 * a] * M3[b] + R31[3 * a + 1] * M3[3 + b] + R31[3 * a + 2] * M3[6 + b];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            Cm[3 * a + b] = R31[3 * a] * acc[b] + R31[3 * a + 1] * acc[3 + b];            // the third row of the top block is zero
            Cm[3 * (a + 3) + b] = R31M[3 * a + 2] * acc[6 + b];                           // only theta's row is non-zero below
        }
    // D = screwTransfInv(delta) * C: [Rd, -Rd hat(t); 0, Rd]
    const double h[9] = {0.0, -delta[2], delta[1], delta[2], 0.0, -delta[0], -delta[1], delta[0], 0.0};
    double T[9], D[18];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) T[3 * a + b] = -(Rd[3 * a] * h[b] + Rd[3 * a + 1] * h[3 + b] + Rd[3 * a + 2] * h[6 + b]);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            double s = 0.0, w = 0.0;
            for (int k = 0; k < 3; k++) { s += Rd[3 * a + k] * Cm[3 * k + b] + T[3 * a + k] * Cm[3 * (k + 3) + b]; w += Rd[3 * a + k] * Cm[3 * (k + 3) + b]; }
            D[3 * a + b] = s; D[3 * (a + 3) + b] = w;
        }
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 3; b++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s += A[6 * a + k] * D[3 * k + b];
            J3_out[18 * (size_t)i + 3 * a + b] = -s;
        }
}


// ---- TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206) ----------------------
// One warp per camera pose: the lanes take the board points (X = R b + t with [t, r] = camPose^-1 o xiBoard, the
// model's dP/dX, rows dpdx [-I | hat(X)]), J^T J is summed over the lanes in a fixed order, lane 0 inverts
// J^T C J = stiffness * J^T J by LU with partial pivoting and forms inv^T J^T J inv.
template <int MODEL>
__global__ void __launch_bounds__(128) visual_cov_kernel(const double *intr_g, const double *xi_board, int P, const double *board,
                                                          double stiffness, int n, const double *cam_poses, double *cov)
{
    using CAM = Camera<MODEL>;
    constexpr int K = CAM::K;
    const int lane = threadIdx.x & 31, pose = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pose >= n) return;
    double intr[K];
#pragma unroll
    for (int i = 0; i < K; i++) intr[i] = intr_g[i];
    const typename CAM::Consts cc = CAM::prepare(intr);
    double xcb[6], R[9], M[9];
    se3_inverse_compose(cam_poses + 6 * (size_t)pose, xi_board, xcb);
    rodrigues_and_left_jacobian(xcb[3], xcb[4], xcb[5], R, M);
    double acc[21];
#pragma unroll
    for (int i = 0; i < 21; i++) acc[i] = 0.0;
    for (int i = lane; i < P; i += 32) {
        const double bx = board[3 * i], by = board[3 * i + 1], bz = board[3 * i + 2];
        const double X[3] = {fma(R[0], bx, fma(R[1], by, R[2] * bz)) + xcb[0], fma(R[3], bx, fma(R[4], by, R[5] * bz)) + xcb[1],
                             fma(R[6], bx, fma(R[7], by, R[8] * bz)) + xcb[2]};
        double u, v, Pu[3], Pv[3], Ju[K], Jv[K];
        const bool ok = CAM::eval(intr, cc, X[0], X[1], X[2], u, v, Pu, Pv, Ju, Jv);
        double d[2][6];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double *Pq = q ? Pv : Pu;
            const double p0 = ok ? Pq[0] : 0.0, p1 = ok ? Pq[1] : 0.0, p2 = ok ? Pq[2] : 0.0;   // zero rows on a failed projection
            d[q][0] = -p0; d[q][1] = -p1; d[q][2] = -p2;
            // P hat(X): columns (p1 X2 - p2 X1, p2 X0 - p0 X2, p0 X1 - p1 X0)
            d[q][3] = p1 * X[2] - p2 * X[1];
            d[q][4] = p2 * X[0] - p0 * X[2];
            d[q][5] = p0 * X[1] - p1 * X[0];
        }
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b <= a; b++) acc[lt(a, b)] += fma(d[0][a], d[0][b], d[1][a] * d[1][b]);
    }
    // lanes combined in a fixed (butterfly) order
#pragma unroll
    for (int i = 0; i < 21; i++)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    if (lane != 0) return;
    double JtJ[36], lu[36], inv[36];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b < 6; b++) {
            JtJ[6 * a + b] = acc[a >= b ? lt(a, b) : lt(b, a)];
            lu[6 * a + b] = stiffness * JtJ[6 * a + b];
        }
    int perm[6] = {0, 1, 2, 3, 4, 5};
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int i = c + 1; i < 6; i++) if (fabs(lu[6 * i + c]) > fabs(lu[6 * piv + c])) piv = i;
        if (piv != c) {
            for (int j = 0; j < 6; j++) { const double t = lu[6 * c + j]; lu[6 * c + j] = lu[6 * piv + j]; lu[6 * piv + j] = t; }
            const int t = perm[c]; perm[c] = perm[piv]; perm[piv] = t;
        }
        for (int i = c + 1; i < 6; i++) {
            lu[6 * i + c] /= lu[6 * c + c];
            for (int j = c + 1; j < 6; j++) lu[6 * i + j] = fma(-lu[6 * i + c], lu[6 * c + j], lu[6 * i + j]);
        }
    }
    for (int c = 0; c < 6; c++) {
        double x[6];
        for (int i = 0; i < 6; i++) {
            double s = perm[i] == c ? 1.0 : 0.0;
            for (int j = 0; j < i; j++) s = fma(-lu[6 * i + j], x[j], s);
            x[i] = s;
        }
        for (int i = 5; i >= 0; i--) {
            double s = x[i];
            for (int j = i + 1; j < 6; j++) s = fma(-lu[6 * i + j], x[j], s);
            x[i] = s / lu[6 * i + i];
        }
        for (int i = 0; i < 6; i++) inv[6 * i + c] = x[i];
    }
    double tmp[36];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s = fma(inv[6 * k + a], JtJ[6 * k + b], s);
            tmp[6 * a + b] = s;
        }
    double *out = cov + 36 * (size_t)pose;
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s = fma(tmp[6 * a + k], inv[6 * k + b], s);
            out[6 * a + b] = s;
        }
}

// ---- block-tridiagonal pose elimination --------------------------------------------------------------------
// One warp per segment (a run of consecutive sequence elements linked by odometry blocks, or a single element
// that carries a prior / is constant).  The 6x6 work of an element is done by every lane (same values, no
// shuffles); the Ks + 1 right-hand-side columns (E^T and the gradient) are spread over the lanes.
//   (C_i + D_i) - L(i,i-1) L(i,i-1)^T = L_ii L_ii^T,  L(i,i-1) = O_i L_(i-1,i-1)^-T
//   Z_i = L_ii^-1 (E_i^T - L(i,i-1) Z_(i-1)),  z_i likewise from the gradient
constexpr int CHAIN_WARPS = 4;

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
chain_factor_kernel(const DatasetDesc *desc_all, int Ks, const int *pose_start, const int *contrib_ds, const int *contrib_img,
                    double *scale, LmConsts lm, double *ws, ChainTables c, double *seg_gmax, int *fail_flag)
{
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * CHAIN_WARPS + (threadIdx.x >> 5);
    if (seg >= c.n_seg) return;
    const int p0 = c.seg_start[seg], n = c.seg_len[seg];
    const int stride = pose_ws_stride(Ks);
    double Lprev[21], invd_prev[6];
    bool prev_free = false, ok = true;
    double gmax = 0.0;
    for (int i = 0; i < n; i++) {
        const int p = p0 + i;
        const bool fixed = c.fixed[p] != 0;
        double C[21], b[6], lam[6], invd[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
        for (int k = 0; k < 21; k++) C[k] = 0.0;
        for (int k = 0; k < 6; k++) b[k] = 0.0;
        const int c0 = pose_start[p], c1 = pose_start[p + 1];
        if (!fixed) {
            for (int q = c0; q < c1; q++) {
                const DatasetDesc &d = desc_all[contrib_ds[q]];
                const double *H = d.H + (size_t)contrib_img[q] * d.ne;
                const int pc = d.pose_col, W = d.W;
                for (int a = 0; a < 6; a++) {
                    for (int j = 0; j <= a; j++) C[lt(a, j)] += H[(pc + j) * W - (pc + j) * (pc + j - 1) / 2 + (a - j)];
                    b[a] += H[(pc + a) * W - (pc + a) * (pc + a - 1) / 2 + (W - 1 - pc - a)];
                }
            }
            for (int x = c.extra_start[p]; x < c.extra_start[p + 1]; x++) {
                const int kind = c.extra_kind[x], rec = c.extra_rec[x];
                const double *Hx, *gx;
                if (kind == EXTRA_TP) { Hx = c.tp_const + (size_t)TP_CONST * rec + TP_OFF_AtA; gx = c.tp_out + (size_t)TP_OUT * rec; }
                else if (kind == EXTRA_OP_FIRST) { Hx = c.op_out + (size_t)OP_OUT * rec + OP_OFF_H11; gx = c.op_out + (size_t)OP_OUT * rec + OP_OFF_G1; }
                else { Hx = c.op_out + (size_t)OP_OUT * rec + OP_OFF_H22; gx = c.op_out + (size_t)OP_OUT * rec + OP_OFF_G2; }
                for (int k = 0; k < 21; k++) C[k] += Hx[k];
                for (int k = 0; k < 6; k++) b[k] += gx[k];
            }
        }
        bool empty = true;
        for (int k = 0; k < 6; k++) {
            const double ckk = C[lt(k, k)];
            if (ckk != 0.0) empty = false;
            double sc;
            if (lm.init_scale) {
                sc = lm.jacobi_scaling ? 1.0 / (1.0 + sqrt(ckk)) : 1.0;
                if (lane == 0) scale[(size_t)p * 6 + k] = sc;
            } else {
                sc = scale[(size_t)p * 6 + k];
            }
            const double s2 = sc * sc;
            lam[k] = fmin(fmax(s2 * ckk, lm.min_diag), lm.max_diag) / (lm.radius * s2);
            gmax = fmax(gmax, fabs(b[k]));
        }
        double Loff[36];
        for (int k = 0; k < 36; k++) Loff[k] = 0.0;
        const int edge = c.prev_edge[p];
        const bool coupled = i > 0 && edge >= 0 && !fixed && !empty && prev_free;
        if (empty) {
            // a constant element, or one nothing constrains: identity factor, zero step
            for (int a = 0; a < 6; a++)
                for (int j = 0; j <= a; j++) C[lt(a, j)] = (a == j) ? 1.0 : 0.0;
            for (int k = 0; k < 6; k++) lam[k] = 0.0;
        } else {
            for (int k = 0; k < 6; k++) C[lt(k, k)] += lam[k];
            if (coupled) {
                const double *O = c.op_out + (size_t)OP_OUT * edge + OP_OFF_O;
                for (int r = 0; r < 6; r++) {
                    double x[6];
                    for (int k = 0; k < 6; k++) x[k] = O[6 * r + k];
                    forward6(Lprev, invd_prev, x);
                    for (int k = 0; k < 6; k++) Loff[6 * r + k] = x[k];
                }
                for (int a = 0; a < 6; a++)
                    for (int j = 0; j <= a; j++) {
                        double s = C[lt(a, j)];
                        for (int k = 0; k < 6; k++) s = fma(-Loff[6 * a + k], Loff[6 * j + k], s);
                        C[lt(a, j)] = s;
                    }
            }
            for (int j = 0; j < 6; j++) {
                double s = C[lt(j, j)];
                for (int k = 0; k < j; k++) s = fma(-C[lt(j, k)], C[lt(j, k)], s);
                if (!(s > 0.0)) { ok = false; s = 1.0; }
                s = sqrt(s);
                C[lt(j, j)] = s;
                const double inv = 1.0 / s;
                invd[j] = inv;
                for (int a = j + 1; a < 6; a++) {
                    double t = C[lt(a, j)];
                    for (int k = 0; k < j; k++) t = fma(-C[lt(a, k)], C[lt(j, k)], t);
                    C[lt(a, j)] = t * inv;
                }
            }
        }
        double *w = ws + (size_t)p * stride;
        if (lane == 0) {
            for (int k = 0; k < 21; k++) w[k] = C[k];
            for (int k = 0; k < 6; k++) w[21 + k] = lam[k];
            for (int k = 0; k < 36; k++) c.off[(size_t)p * 36 + k] = Loff[k];
        }
        const double *wprev = w - stride;
        for (int col = lane; col <= Ks; col += 32) {
            double e[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (!fixed) {
                if (col == Ks) {
                    for (int k = 0; k < 6; k++) e[k] = b[k];
                } else {
                    for (int q = c0; q < c1; q++) {
                        const DatasetDesc &d = desc_all[contrib_ds[q]];
                        int lc = -1;
                        for (int s = 0; s < d.n_sl; s++)
                            if (d.sl_idx[s] == col) lc = d.sl_col[s];
                        if (lc < 0) continue;
                        const double *H = d.H + (size_t)contrib_img[q] * d.ne;
                        const int pc = d.pose_col, W = d.W;
                        for (int k = 0; k < 6; k++) {
                            const int a = min(lc, pc + k), bb = max(lc, pc + k);
                            e[k] += H[a * W - a * (a - 1) / 2 + (bb - a)];
                        }
                    }
                }
            }
            if (coupled) {
                double pz[6];
                for (int k = 0; k < 6; k++) pz[k] = col < Ks ? wprev[33 + k * Ks + col] : wprev[27 + k];
                for (int r = 0; r < 6; r++) {
                    double s = e[r];
                    for (int k = 0; k < 6; k++) s = fma(-Loff[6 * r + k], pz[k], s);
                    e[r] = s;
                }
            }
            forward6(C, invd, e);
            if (col < Ks) {
                for (int k = 0; k < 6; k++) w[33 + k * Ks + col] = e[k];
            } else {
                for (int k = 0; k < 6; k++) w[27 + k] = e[k];
            }
        }
        for (int k = 0; k < 21; k++) Lprev[k] = C[k];
        for (int k = 0; k < 6; k++) invd_prev[k] = invd[k];
        prev_free = !fixed && !empty;
    }
    if (lane == 0) {
        seg_gmax[seg] = gmax;
        if (!ok) atomicExch(fail_flag, 1);
    }
}

// delta = -L^-T w over a segment, last element first; candidate poses; model-decrease and norm partial sums.
// (-1/2 w^T w is added where w is formed, pose_backsub_kernel.)
__global__ void __launch_bounds__(32)
chain_backsub_kernel(int Ks, const double *const *seq_cur, double *const *seq_cand, const int *pose_seq, const int *pose_local,
                     const double *ws, ChainTables c, double *seg_partial)
{
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= c.n_seg) return;
    const int p0 = c.seg_start[seg], n = c.seg_len[seg];
    const int stride = pose_ws_stride(Ks);
    double y_next[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double m = 0.0, st2 = 0.0, x2 = 0.0;
    for (int i = n - 1; i >= 0; i--) {
        const int p = p0 + i;
        const double *w = ws + (size_t)p * stride;
        double y[6];
        for (int k = 0; k < 6; k++) y[k] = c.w[(size_t)p * 6 + k];
        if (i + 1 < n) {
            const double *Loff = c.off + (size_t)(p + 1) * 36;      // L(i+1, i)
            for (int k = 0; k < 6; k++) {
                double s = y[k];
                for (int r = 0; r < 6; r++) s = fma(-Loff[6 * r + k], y_next[r], s);
                y[k] = s;
            }
        }
        for (int a = 5; a >= 0; a--) {
            double s = y[a];
            for (int k = a + 1; k < 6; k++) s = fma(-w[lt(k, a)], y[k], s);
            y[a] = s / w[lt(a, a)];
        }
        const double *cur = seq_cur[pose_seq[p]] + (size_t)pose_local[p] * 6;
        double *cand = seq_cand[pose_seq[p]] + (size_t)pose_local[p] * 6;
        const bool fixed = c.fixed[p] != 0;
        for (int k = 0; k < 6; k++) {
            const double dlt = fixed ? 0.0 : -y[k];
            const double x = cur[k];
            cand[k] = x + dlt;
            m = fma(-0.5 * w[21 + k] * dlt, dlt, m);
            st2 = fma(dlt, dlt, st2);
            if (!fixed) x2 = fma(x, x, x2);
            y_next[k] = fixed ? 0.0 : y[k];
        }
    }
    seg_partial[(size_t)seg * 3] = m;
    seg_partial[(size_t)seg * 3 + 1] = st2;
    seg_partial[(size_t)seg * 3 + 2] = x2;
}

}  // namespace

cudaError_t launch_prior_setup(int n_tp, const double *tp_in, double *tp_const, int n_op, const double *op_in, double *op_const,
                               SolverLaunch sl)
{
    const int n = n_tp + n_op;
    if (n <= 0) return cudaSuccess;
    prior_setup_kernel<<<(n + 63) / 64, 64, 0, sl.stream>>>(n_tp, tp_in, tp_const, n_op, op_in, op_const);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

cudaError_t launch_prior_eval(const PriorTables &t, int Ks, double *red, int add_shared, double *, SolverLaunch sl)
{
    const int n = t.n_tp + t.n_op;
    if (n <= 0) return cudaSuccess;
    prior_eval_kernel<<<(n + 63) / 64, 64, 0, sl.stream>>>(t);
    prior_reduce_kernel<<<1, 256, 0, sl.stream>>>(t, Ks, red, add_shared);
    if (sl.launches) (*sl.launches) += 2;
    return cudaGetLastError();
}

cudaError_t launch_chain_factor(const DatasetDesc *d_desc, int Ks, const int *pose_start, const int *contrib_ds,
                                const int *contrib_img, double *scale, LmConsts lm, double *ws, const ChainTables &c,
                                double *seg_gmax, int *fail_flag, SolverLaunch sl)
{
    if (c.n_seg <= 0) return cudaSuccess;
    chain_factor_kernel<<<(c.n_seg + CHAIN_WARPS - 1) / CHAIN_WARPS, CHAIN_WARPS * 32, 0, sl.stream>>>(
        d_desc, Ks, pose_start, contrib_ds, contrib_img, scale, lm, ws, c, seg_gmax, fail_flag);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

cudaError_t launch_chain_backsub(int Ks, const double *const *seq_cur, double *const *seq_cand, const int *pose_seq,
                                 const int *pose_local, const double *ws, const ChainTables &c, double *seg_partial,
                                 SolverLaunch sl)
{
    if (c.n_seg <= 0) return cudaSuccess;
    chain_backsub_kernel<<<(c.n_seg + 31) / 32, 32, 0, sl.stream>>>(Ks, seq_cur, seq_cand, pose_seq, pose_local, ws, c, seg_partial);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

}  // namespace vg

// ---- C ABI, inner level: the prior functors' own contract, batched (include/visgeom_b200.h) -------------------
using namespace vg;

namespace {

struct DevBuf {
    double *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(double) * (n ? n : 1)); }
    cudaError_t up(const double *h, size_t n) { cudaError_t e = alloc(n); return e != cudaSuccess || !n ? e : cudaMemcpy(p, h, sizeof(double) * n, cudaMemcpyHostToDevice); }
    cudaError_t down(double *h, size_t n) const { return n ? cudaMemcpy(h, p, sizeof(double) * n, cudaMemcpyDeviceToHost) : cudaSuccess; }
};

int have_device()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
        cudaGetLastError();
        return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    }
    return VG_OK;
}

}  // namespace

extern "C" {

int vg_eval_transformation_prior(int n, const double *stiffness, const double *xi_prior, const double *xi, double *r, double *J)
{
    if (n < 0 || (n > 0 && (!stiffness || !xi_prior || !xi || !r))) return fail(VG_ERR_INVALID, "vg_eval_transformation_prior: bad arguments");
    int rc = have_device();
    if (rc || n == 0) return rc;
    DevBuf ds, dp, dx, dr, dj;
    const size_t m = (size_t)n;
    VG_CUDA(ds.up(stiffness, 6 * m)); VG_CUDA(dp.up(xi_prior, 6 * m)); VG_CUDA(dx.up(xi, 6 * m));
    VG_CUDA(dr.alloc(6 * m));
    if (J) VG_CUDA(dj.alloc(36 * m));
    tp_functor_kernel<<<(n + 63) / 64, 64>>>(n, ds.p, dp.p, dx.p, dr.p, J ? dj.p : nullptr);
    count_launch(&launch_counter());
    VG_CUDA(cudaGetLastError());
    VG_CUDA(dr.down(r, 6 * m));
    if (J) VG_CUDA(dj.down(J, 36 * m));
    return VG_OK;
}

int vg_eval_odometry_prior(int n, double errV, double errW, double lambda, const double *odom1, const double *odom2,
                           const double *xi1, const double *xi2, double *r, double *J1, double *J2)
{
    if (n < 0 || (n > 0 && (!odom1 || !odom2 || !xi1 || !xi2 || !r))) return fail(VG_ERR_INVALID, "vg_eval_odometry_prior: bad arguments");
    int rc = have_device();
    if (rc || n == 0) return rc;
    DevBuf o1, o2, x1, x2, dr, j1, j2;
    const size_t m = (size_t)n;
    VG_CUDA(o1.up(odom1, 6 * m)); VG_CUDA(o2.up(odom2, 6 * m)); VG_CUDA(x1.up(xi1, 6 * m)); VG_CUDA(x2.up(xi2, 6 * m));
    VG_CUDA(dr.alloc(6 * m));
    if (J1) VG_CUDA(j1.alloc(36 * m));
    if (J2) VG_CUDA(j2.alloc(36 * m));
    op_functor_kernel<<<(n + 63) / 64, 64>>>(n, errV, errW, lambda, o1.p, o2.p, x1.p, x2.p, dr.p, J1 ? j1.p : nullptr, J2 ? j2.p : nullptr);
    count_launch(&launch_counter());
    VG_CUDA(cudaGetLastError());
    VG_CUDA(dr.down(r, 6 * m));
    if (J1) VG_CUDA(j1.down(J1, 36 * m));
    if (J2) VG_CUDA(j2.down(J2, 36 * m));
    return VG_OK;
}

int vg_eval_odometry_cost(int n, double errV, double errW, double lambda, const int *dq_offset, const double *dq,
                          const double *intr_prior, const double *xi1, const double *xi2, const double *intr, double *r,
                          double *J1, double *J2, double *J3)
{
    if (n < 0 || (n > 0 && (!dq_offset || !dq || !intr_prior || !xi1 || !xi2 || !intr || !r)))
        return fail(VG_ERR_INVALID, "vg_eval_odometry_cost: bad arguments");
    for (int i = 0; i < n; i++)
        if (dq_offset[i + 1] - dq_offset[i] < 1 || dq_offset[i] < 0)
            return fail(VG_ERR_INVALID, "vg_eval_odometry_cost: every block needs at least one pair of increments");
    if (!(lambda > 0.0)) return fail(VG_ERR_INVALID, "vg_eval_odometry_cost: lambda must be positive");
    int rc = have_device();
    if (rc || n == 0) return rc;
    const size_t m = (size_t)n, total = (size_t)dq_offset[n];
    DevBuf q, ip, it, x1, x2, dr, j1, j2, j3;
    int *d_off = nullptr;
    VG_CUDA(q.up(dq, 2 * total)); VG_CUDA(ip.up(intr_prior, 3)); VG_CUDA(it.up(intr, 3));
    VG_CUDA(x1.up(xi1, 6 * m)); VG_CUDA(x2.up(xi2, 6 * m)); VG_CUDA(dr.alloc(6 * m));
    if (J1) VG_CUDA(j1.alloc(36 * m));
    if (J2) VG_CUDA(j2.alloc(36 * m));
    if (J3) VG_CUDA(j3.alloc(18 * m));
    VG_CUDA(cudaMalloc(&d_off, sizeof(int) * (m + 1)));
    cudaError_t e = cudaMemcpy(d_off, dq_offset, sizeof(int) * (m + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        oc_functor_kernel<<<(n + 63) / 64, 64>>>(n, errV, errW, lambda, d_off, q.p, ip.p, x1.p, x2.p, it.p, dr.p, J1 ? j1.p : nullptr,
                                                 J2 ? j2.p : nullptr, J3 ? j3.p : nullptr);
        count_launch(&launch_counter());
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = dr.down(r, 6 * m);
    if (e == cudaSuccess && J1) e = j1.down(J1, 36 * m);
    if (e == cudaSuccess && J2) e = j2.down(J2, 36 * m);
    if (e == cudaSuccess && J3) e = j3.down(J3, 18 * m);
    cudaFree(d_off);
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_eval_odometry_cost");
}

int vg_visual_cov(int model, const double *intr, const double *xi_board, int P, const double *board, double feature_variance,
                  int n, const double *cam_poses, double *cov)
{
    const int K = vg_model_num_params(model);
    if (K < 0) return K;
    if (n < 0 || P < 1 || !intr || !xi_board || !board || !(feature_variance > 0.0) || (n > 0 && (!cam_poses || !cov)))
        return fail(VG_ERR_INVALID, "vg_visual_cov: bad arguments");
    int rc = have_device();
    if (rc || n == 0) return rc;
    DevBuf di, dx, db, dp, dc;
    VG_CUDA(di.up(intr, (size_t)K)); VG_CUDA(dx.up(xi_board, 6)); VG_CUDA(db.up(board, 3 * (size_t)P));
    VG_CUDA(dp.up(cam_poses, 6 * (size_t)n)); VG_CUDA(dc.alloc(36 * (size_t)n));
    const double stiffness = std::sqrt(1.0 / feature_variance);      // U of the Cholesky factorisation of diag(1 / variance), :112-115
    const int blocks = (n + 3) / 4;
    if (model == VG_MODEL_EUCM) visual_cov_kernel<MODEL_EUCM><<<blocks, 128>>>(di.p, dx.p, P, db.p, stiffness, n, dp.p, dc.p);
    else if (model == VG_MODEL_UCM) visual_cov_kernel<MODEL_UCM><<<blocks, 128>>>(di.p, dx.p, P, db.p, stiffness, n, dp.p, dc.p);
    else visual_cov_kernel<MODEL_MEI><<<blocks, 128>>>(di.p, dx.p, P, db.p, stiffness, n, dp.p, dc.p);
    count_launch(&launch_counter());
    VG_CUDA(cudaGetLastError());
    VG_CUDA(dc.down(cov, 36 * (size_t)n));
    return VG_OK;
}

}  // extern "C"
