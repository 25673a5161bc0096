/*
 * oracle_tuned.c -- CPU ORACLE, tuned variant (test infrastructure, NOT product code).
 *
 * SURVEY.md 8(d) asks for two CPU baselines: the *faithful* restatement (visgeom_oracle.c keeps the reference's cost
 * profile: three camera calls per corner each recomputing rho / eta, the chain composed twice, a heap vector per
 * call, a camera clone per InterJacobian) and a *tuned* one "for an honest comparison".  This is the tuned one: the
 * same arithmetic definitions (calib_cost_functions.cpp:28-117, jacobian.h:136-171, eucm.h / ucm.h / mei.h), but
 * the chain is composed once, nothing is allocated per image, and projection, dP/dX and dP/dintr of a corner share
 * their sub-expressions.  Results agree with the faithful oracle to rounding (tests/test_oracle_math.py).
 */
#include "visgeom_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXK 10
#define MAXL 5

/* u, v, rows of dP/dX (Pu, Pv) and of dP/dintr (Ju, Jv) in one pass; returns 0 when the projection fails */
static int camera_all(int model, const double *p, const double X[3], double uv[2], double Pu[3], double Pv[3],
                      double *Ju, double *Jv)
{
    const double x = X[0], y = X[1], z = X[2];
    if (model == VGO_EUCM) {
        const double alpha = p[0], beta = p[1], fu = p[2], fv = p[3], u0 = p[4], v0 = p[5];
        const double r2 = x * x + y * y, rho = sqrt(z * z + beta * r2), gamma = 1. - alpha;
        const double eta = alpha * rho + gamma * z;
        if (eta < 1e-3) return 0;
        if (alpha > 0.5 && z / eta < (alpha - 1.) / (alpha + alpha - 1.)) return 0;
        const double ie = 1. / eta, k = ie * ie, irho = 1. / rho, a = alpha * beta * irho;
        uv[0] = fu * x * ie + u0; uv[1] = fv * y * ie + v0;
        const double jz = k * (gamma + alpha * z * irho), jxy = k * a * x * y;
        Pu[0] = fu * k * (eta - a * x * x); Pu[1] = -fu * jxy; Pu[2] = -fu * x * jz;
        Pv[0] = -fv * jxy; Pv[1] = fv * k * (eta - a * y * y); Pv[2] = -fv * y * jz;
        const double da = (rho - z) * k, db = alpha * r2 * k * 0.5 * irho;
        Ju[0] = -fu * x * da; Ju[1] = -fu * x * db; Ju[2] = x * ie; Ju[3] = 0; Ju[4] = 1; Ju[5] = 0;
        Jv[0] = -fv * y * da; Jv[1] = -fv * y * db; Jv[2] = 0; Jv[3] = y * ie; Jv[4] = 0; Jv[5] = 1;
        return 1;
    }
    /* UCM and MEI share the unified normalised point */
    const double xi = p[0];
    const double rho = sqrt(x * x + y * y + z * z), d = 1. / (z + xi * rho), d2 = d * d, irho = 1. / rho;
    const double xn = x * d, yn = y * d;
    const double n00 = (xi * rho + z - xi * x * x * irho) * d2, n01 = -xi * x * y * d2 * irho, n02 = -x * (1. + xi * z * irho) * d2;
    const double n11 = (xi * rho + z - xi * y * y * irho) * d2, n12 = -y * (1. + xi * z * irho) * d2;
    const double dxi_x = -xn * d * rho, dxi_y = -yn * d * rho;      /* d(xn, yn)/d xi */
    if (model == VGO_UCM) {
        const double fu = p[1], fv = p[2], u0 = p[3], v0 = p[4];
        uv[0] = fu * xn + u0; uv[1] = fv * yn + v0;
        Pu[0] = fu * n00; Pu[1] = fu * n01; Pu[2] = fu * n02;
        Pv[0] = fv * n01; Pv[1] = fv * n11; Pv[2] = fv * n12;
        Ju[0] = fu * dxi_x; Ju[1] = xn; Ju[2] = 0; Ju[3] = 1; Ju[4] = 0;
        Jv[0] = fv * dxi_y; Jv[1] = 0; Jv[2] = yn; Jv[3] = 0; Jv[4] = 1;
        return 1;
    }
    const double k1 = p[1], k2 = p[2], k3 = p[3], k4 = p[4], k5 = p[5], fu = p[6], fv = p[7], u0 = p[8], v0 = p[9];
    const double r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
    const double D = 1. + k1 * r2 + k2 * r4 + k3 * r6, dD = k1 + 2. * k2 * r2 + 3. * k3 * r4, xy = xn * yn;
    const double xd = xn * D + 2. * k4 * xy + k5 * (r2 + 2. * xn * xn);
    const double yd = yn * D + 2. * k5 * xy + k4 * (r2 + 2. * yn * yn);
    uv[0] = fu * xd + u0; uv[1] = fv * yd + v0;
    const double a00 = D + 2. * xn * xn * dD + 2. * k4 * yn + 6. * k5 * xn, a01 = 2. * xy * dD + 2. * k4 * xn + 2. * k5 * yn;
    const double a10 = 2. * xy * dD + 2. * k5 * yn + 2. * k4 * xn, a11 = D + 2. * yn * yn * dD + 2. * k5 * xn + 6. * k4 * yn;
    Pu[0] = fu * (a00 * n00 + a01 * n01); Pu[1] = fu * (a00 * n01 + a01 * n11); Pu[2] = fu * (a00 * n02 + a01 * n12);
    Pv[0] = fv * (a10 * n00 + a11 * n01); Pv[1] = fv * (a10 * n01 + a11 * n11); Pv[2] = fv * (a10 * n02 + a11 * n12);
    Ju[0] = fu * (a00 * dxi_x + a01 * dxi_y); Ju[1] = fu * xn * r2; Ju[2] = fu * xn * r4; Ju[3] = fu * xn * r6;
    Ju[4] = 2. * fu * xy; Ju[5] = fu * (r2 + 2. * xn * xn); Ju[6] = xd; Ju[7] = 0; Ju[8] = 1; Ju[9] = 0;
    Jv[0] = fv * (a10 * dxi_x + a11 * dxi_y); Jv[1] = fv * yn * r2; Jv[2] = fv * yn * r4; Jv[3] = fv * yn * r6;
    Jv[4] = fv * (r2 + 2. * yn * yn); Jv[5] = 2. * fv * xy; Jv[6] = 0; Jv[7] = yd; Jv[8] = 0; Jv[9] = 1;
    return 1;
}

static void mat3_mul_(const double A[9], const double B[9], double C[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

int vgo_evaluate_batch_tuned(int model, const double *intr, int n_img, int P, const double *board, const double *obs,
                             int chain_len, const int *status, const int *is_global, const double *const *xi,
                             double *r, double *J_intr, double *const *J_xi, double *H, int threads)
{
    const int K = vgo_num_params(model), L = chain_len;
    if (K < 0 || L < 1 || L > MAXL) return -1;
    const int D = K + 6 * L, W = D + 1, ne = W * (W + 1) / 2;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int img = 0; img < n_img; img++) {
        /* the chain, once: accumulated transform and per element the members of InterJacobian (jacobian.h:139-152) */
        double acc[6] = { 0, 0, 0, 0, 0, 0 }, tmp[6], R12[MAXL][9], M12[MAXL][9], t13[MAXL][3];
        for (int e = 0; e < L; e++) {
            const double *x23 = xi[e] + (is_global[e] ? 0 : (size_t)img * 6);
            double x13[6];
            if (status[e] == VGO_TRANSFORM_DIRECT) { vgo_compose(acc, x23, tmp); memcpy(acc, tmp, sizeof tmp); memcpy(x13, acc, sizeof acc); }
            else { memcpy(x13, acc, sizeof acc); vgo_compose_inverse(acc, x23, tmp); memcpy(acc, tmp, sizeof tmp); }
            double R13[9], R23i[9], M[9], nr[3] = { -x23[3], -x23[4], -x23[5] };
            vgo_rotation_matrix(x13 + 3, R13);
            vgo_rotation_matrix(nr, R23i);
            mat3_mul_(R13, R23i, R12[e]);
            vgo_inter_omega_rot(x23 + 3, M);
            mat3_mul_(R12[e], M, M12[e]);
            if (status[e] == VGO_TRANSFORM_INVERSE)
                for (int i = 0; i < 9; i++) { R12[e][i] = -R12[e][i]; M12[e][i] = -M12[e][i]; }
            t13[e][0] = x13[0]; t13[e][1] = x13[1]; t13[e][2] = x13[2];
        }
        double R[9];
        vgo_rotation_matrix(acc + 3, R);
        double h[(MAXK + 6 * MAXL + 1) * (MAXK + 6 * MAXL + 2) / 2];
        if (H) memset(h, 0, sizeof(double) * (size_t)ne);
        for (int i = 0; i < P; i++) {
            const double *Xb = board + 3 * i;
            const double X[3] = { R[0] * Xb[0] + R[1] * Xb[1] + R[2] * Xb[2] + acc[0],
                                  R[3] * Xb[0] + R[4] * Xb[1] + R[5] * Xb[2] + acc[1],
                                  R[6] * Xb[0] + R[7] * Xb[1] + R[8] * Xb[2] + acc[2] };
            double rows[2][MAXK + 6 * MAXL + 1], uv[2], Pu[3], Pv[3];
            memset(rows, 0, sizeof rows);
            const size_t o = ((size_t)img * P + i) * 2;
            if (camera_all(model, intr, X, uv, Pu, Pv, rows[0], rows[1])) {
                rows[0][D] = uv[0] - obs[o]; rows[1][D] = uv[1] - obs[o + 1];
                for (int e = 0; e < L; e++) {
                    const double w[3] = { X[0] - t13[e][0], X[1] - t13[e][1], X[2] - t13[e][2] };
                    for (int q = 0; q < 2; q++) {
                        const double *Pq = q ? Pv : Pu;
                        double *Jr = rows[q] + K + 6 * e;
                        /* -P hat(w) = (w x P)^T */
                        const double c[3] = { w[1] * Pq[2] - w[2] * Pq[1], w[2] * Pq[0] - w[0] * Pq[2], w[0] * Pq[1] - w[1] * Pq[0] };
                        for (int j = 0; j < 3; j++) {
                            Jr[j] = Pq[0] * R12[e][j] + Pq[1] * R12[e][3 + j] + Pq[2] * R12[e][6 + j];
                            Jr[3 + j] = c[0] * M12[e][j] + c[1] * M12[e][3 + j] + c[2] * M12[e][6 + j];
                        }
                    }
                }
            } else {
                memset(rows, 0, sizeof rows);
                rows[0][D] = rows[1][D] = VGO_DOUBLE_BIG;
            }
            for (int q = 0; q < 2; q++) {
                const size_t row = o + q;
                if (r) r[row] = rows[q][D];
                if (J_intr) memcpy(J_intr + row * K, rows[q], sizeof(double) * (size_t)K);
                for (int e = 0; e < L; e++)
                    if (J_xi && J_xi[e]) memcpy(J_xi[e] + row * 6, rows[q] + K + 6 * e, 6 * sizeof(double));
                if (H) {
                    int idx = 0;
                    for (int a = 0; a < W; a++) {
                        const double ra = rows[q][a];
                        for (int b = a; b < W; b++) h[idx++] += ra * rows[q][b];
                    }
                }
            }
        }
        if (H) memcpy(H + (size_t)img * ne, h, sizeof(double) * (size_t)ne);
    }
    return 0;
}
