/*
 * visgeom_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See visgeom_oracle.h.  Each function names the reference lines it follows
 * (paths relative to /root/reference).  The structure deliberately keeps the
 * reference's cost profile: the three per-corner camera calls each recompute
 * rho / eta (eucm.h:45,128-130,184-187), the chain is composed twice per
 * Evaluate (calib_cost_functions.cpp:32-46,76-92), the transformed point vector
 * is heap allocated per call (:49) and a camera clone is made per InterJacobian
 * (jacobian.h:141; the reference never frees it, we do).
 */
#include "visgeom_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* small 3x3 helpers (row-major), written in the order Eigen evaluates */
static void mat3_mul(const double A[9], const double B[9], double C[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

static void mat3_vec(const double A[9], const double v[3], double o[3])
{
    for (int i = 0; i < 3; i++)
        o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}

/* row(1x3) * M(3x3) */
static void row_mat3(const double r[3], const double M[9], double o[3])
{
    for (int j = 0; j < 3; j++)
        o[j] = r[0] * M[j] + r[1] * M[3 + j] + r[2] * M[6 + j];
}

static double norm3(const double v[3])
{
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

/* geometry_core.h:126-132 */
static void hat3(const double u[3], double M[9])
{
    M[0] = 0;      M[1] = -u[2];  M[2] = u[1];
    M[3] = u[2];   M[4] = 0;      M[5] = -u[0];
    M[6] = -u[1];  M[7] = u[0];   M[8] = 0;
}

/* geometry_core.h:24-30 */
static double sinc_(double x)
{
    if (x == 0.) return 1.;
    return sin(x) / x;
}

/* geometry_core.h:32-38 */
static double normalize_angle(double th)
{
    if (th > M_PI) return th - 2 * M_PI;
    else if (th < -M_PI) return th + 2 * M_PI;
    else return th;
}

/* ------------------------------------------------------------------ */
int vgo_num_params(int model)
{
    switch (model) {
    case VGO_EUCM: return 6;   /* eucm.h:64 */
    case VGO_UCM:  return 5;   /* ucm.h:61 */
    case VGO_MEI:  return 10;  /* mei.h:69 */
    default: return -1;
    }
}

double vgo_upper_bound(int model, int idx)
{
    if (model == VGO_EUCM) {            /* eucm.h:228-236 */
        if (idx == 0) return 1;
        if (idx == 1) return 10;
        return 1e5;
    }
    if (model == VGO_UCM) {             /* ucm.h:199-206 */
        if (idx == 0) return 3;
        return 1e5;
    }
    /* mei.h:287-299 */
    if (idx == 0) return 3;
    if (idx >= 1 && idx <= 5) return 10;
    return 1e5;
}

double vgo_lower_bound(int model, int idx)
{
    if (model == VGO_EUCM) {            /* eucm.h:238-246 */
        if (idx == 0) return 0;
        if (idx == 1) return 0.1;
        return 1;
    }
    if (model == VGO_UCM) {             /* ucm.h:208-215 */
        if (idx == 0) return 0;
        return 1;
    }
    /* mei.h:301-313 */
    if (idx == 0) return 0;
    if (idx >= 1 && idx <= 5) return -10;
    return 1;
}

/* ------------------------------------------------------------------ */
/* geometry_core.h:40-76 */
void vgo_rotation_matrix(const double v[3], double R[9])
{
    double th = norm3(v);
    if (th < 1e-5) {
        R[0] = 1.;     R[1] = -v[2];  R[2] = v[1];
        R[3] = v[2];   R[4] = 1.;     R[5] = -v[0];
        R[6] = -v[1];  R[7] = v[0];   R[8] = 1.;
    } else {
        double thInv = 1. / th;
        double u1 = v[0] * thInv;
        double u2 = v[1] * thInv;
        double u3 = v[2] * thInv;
        double sinth = sin(th);
        double costhVar = 1. - cos(th);

        R[0] = 1. + costhVar * (u1 * u1 - 1.);
        R[4] = 1. + costhVar * (u2 * u2 - 1.);
        R[8] = 1. + costhVar * (u3 * u3 - 1.);

        R[1] = -sinth * u3 + costhVar * u1 * u2;
        R[2] = sinth * u2 + costhVar * u1 * u3;
        R[5] = -sinth * u1 + costhVar * u2 * u3;

        R[3] = sinth * u3 + costhVar * u2 * u1;
        R[6] = -sinth * u2 + costhVar * u3 * u1;
        R[7] = sinth * u1 + costhVar * u3 * u2;
    }
}

/* geometry_core.h:158-180 */
void vgo_inter_omega_rot(const double v[3], double B[9])
{
    double theta = norm3(v);
    if (theta < 1e-5) {
        double h0 = v[0] / 2., h1 = v[1] / 2., h2 = v[2] / 2.;
        B[0] = 1.;   B[1] = -h2;  B[2] = h1;
        B[3] = h2;   B[4] = 1.;   B[5] = -h0;
        B[6] = -h1;  B[7] = h0;   B[8] = 1.;
    } else {
        double u[3] = { v[0] / theta, v[1] / theta, v[2] / theta };
        double uhat[9], uhat2[9];
        hat3(u, uhat);
        double thetaHalf = theta / 2.;
        double K1 = sinc_(thetaHalf);
        K1 = thetaHalf * K1 * K1;
        double K2 = 1. - sinc_(theta);
        mat3_mul(uhat, uhat, uhat2);
        for (int i = 0; i < 9; i++) {
            double I = (i == 0 || i == 4 || i == 8) ? 1. : 0.;
            B[i] = (I + K1 * uhat[i]) + K2 * uhat2[i];
        }
    }
}

/* quaternion.h:31-50 ; layout (x,y,z,w) */
void vgo_quat_from_rotvec(const double rot[3], double q[4])
{
    double theta = norm3(rot);
    if (fabs(theta) < 1e-6) {
        q[0] = rot[0] / 2.;
        q[1] = rot[1] / 2.;
        q[2] = rot[2] / 2.;
        q[3] = 1.;
    } else {
        double u0 = rot[0] / theta, u1 = rot[1] / theta, u2 = rot[2] / theta;
        double s = sin(theta / 2.);
        q[0] = u0 * s;
        q[1] = u1 * s;
        q[2] = u2 * s;
        q[3] = cos(theta / 2.);
    }
}

/* quaternion.h:84-98 */
void vgo_quat_to_rotvec(const double q[4], double r[3])
{
    double s = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (s < 1e-5) {
        r[0] = q[0] * 2.;
        r[1] = q[1] * 2.;
        r[2] = q[2] * 2.;
    } else {
        double th = 2. * atan2(s, q[3]);
        double nth = normalize_angle(th);
        r[0] = q[0] / s * nth;
        r[1] = q[1] / s * nth;
        r[2] = q[2] / s * nth;
    }
}

/* quaternion.h:61-82 */
void vgo_quat_rotate(const double q[4], const double v[3], double out[3])
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double t1 = w * x;
    double t2 = w * y;
    double t3 = w * z;
    double t4 = -x * x;
    double t5 = x * y;
    double t6 = x * z;
    double t7 = -y * y;
    double t8 = y * z;
    double t9 = -z * z;
    double v1 = v[0], v2 = v[1], v3 = v[2];
    out[0] = 2. * ((t7 + t9) * v1 + (t5 - t3) * v2 + (t2 + t6) * v3) + v1;
    out[1] = 2. * ((t3 + t5) * v1 + (t4 + t9) * v2 + (t8 - t1) * v3) + v2;
    out[2] = 2. * ((t6 - t2) * v1 + (t1 + t8) * v2 + (t4 + t7) * v3) + v3;
}

/* quaternion.h:105-118 */
void vgo_quat_mul(const double a[4], const double b[4], double o[4])
{
    double x = a[0], y = a[1], z = a[2], w = a[3];
    double x2 = b[0], y2 = b[1], z2 = b[2], w2 = b[3];
    double wn = w * w2 - x * x2 - y * y2 - z * z2;
    double xn = w * x2 + x * w2 + y * z2 - z * y2;
    double yn = w * y2 - x * z2 + y * w2 + z * x2;
    double zn = w * z2 + x * y2 - y * x2 + z * w2;
    o[0] = xn; o[1] = yn; o[2] = zn; o[3] = wn;
}

/* transformation.h:80-88 */
void vgo_compose(const double a[6], const double b[6], double out[6])
{
    double q1[4], q2[4], qres[4], rt[3];
    vgo_quat_from_rotvec(a + 3, q1);
    vgo_quat_from_rotvec(b + 3, q2);
    vgo_quat_rotate(q1, b, rt);
    double t0 = rt[0] + a[0], t1 = rt[1] + a[1], t2 = rt[2] + a[2];
    vgo_quat_mul(q1, q2, qres);
    double r[3];
    vgo_quat_to_rotvec(qres, r);
    out[0] = t0; out[1] = t1; out[2] = t2;
    out[3] = r[0]; out[4] = r[1]; out[5] = r[2];
}

/* transformation.h:101-110 */
void vgo_compose_inverse(const double a[6], const double b[6], double out[6])
{
    double q1[4], q2[4], q2inv[4], qres[4], rt[3], r[3];
    vgo_quat_from_rotvec(a + 3, q1);
    vgo_quat_from_rotvec(b + 3, q2);
    q2inv[0] = -q2[0]; q2inv[1] = -q2[1]; q2inv[2] = -q2[2]; q2inv[3] = q2[3]; /* quaternion.h:100-103 */
    vgo_quat_mul(q1, q2inv, qres);
    vgo_quat_rotate(qres, b, rt);
    double t0 = a[0] - rt[0], t1 = a[1] - rt[1], t2 = a[2] - rt[2];
    vgo_quat_to_rotvec(qres, r);
    out[0] = t0; out[1] = t1; out[2] = t2;
    out[3] = r[0]; out[4] = r[1]; out[5] = r[2];
}

/* transformation.h:90-99 */
void vgo_inverse_compose(const double a[6], const double b[6], double out[6])
{
    double q1[4], q2[4], q1inv[4], qres[4], d[3], rt[3], r[3];
    vgo_quat_from_rotvec(a + 3, q1);
    vgo_quat_from_rotvec(b + 3, q2);
    q1inv[0] = -q1[0]; q1inv[1] = -q1[1]; q1inv[2] = -q1[2]; q1inv[3] = q1[3];
    d[0] = b[0] - a[0]; d[1] = b[1] - a[1]; d[2] = b[2] - a[2];
    vgo_quat_rotate(q1inv, d, rt);
    vgo_quat_mul(q1inv, q2, qres);
    vgo_quat_to_rotvec(qres, r);
    out[0] = rt[0]; out[1] = rt[1]; out[2] = rt[2];
    out[3] = r[0]; out[4] = r[1]; out[5] = r[2];
}

/* transformation.h:165-169 (single point) == :147-155 + :185-193 (vector form) */
void vgo_transform_point(const double xi[6], const double src[3], double dst[3])
{
    double R[9], o[3];
    vgo_rotation_matrix(xi + 3, R);
    mat3_vec(R, src, o);
    dst[0] = o[0] + xi[0];
    dst[1] = o[1] + xi[1];
    dst[2] = o[2] + xi[2];
}

/* ------------------------------------------------------------------ */
/* EUCM  eucm.h:32-63 */
static int eucm_project(const double *params, const double *src, double *dst)
{
    double alpha = params[0], beta = params[1];
    double fu = params[2], fv = params[3], u0 = params[4], v0 = params[5];
    double x = src[0], y = src[1], z = src[2];

    double denom = alpha * sqrt(z * z + beta * (x * x + y * y)) + (1. - alpha) * z;
    if (denom < 1e-3) return 0;
    if (alpha > 0.5) {
        double zn = z / denom;
        double C = (alpha - 1.) / (alpha + alpha - 1.);
        if (zn < C) return 0;
    }
    double xn = x / denom;
    double yn = y / denom;
    dst[0] = fu * xn + u0;
    dst[1] = fv * yn + v0;
    return 1;
}

/* eucm.h:115-167 */
static int eucm_projection_jacobian(const double *params, const double *src,
                                    double *dudx, double *dvdx)
{
    double alpha = params[0], beta = params[1], fu = params[2], fv = params[3];
    double x = src[0], y = src[1], z = src[2];

    double rho = sqrt(z * z + beta * (x * x + y * y));
    double gamma = 1. - alpha;
    double eta = alpha * rho + gamma * z;

    int isProjected = 1;
    if (eta < 1e-3) isProjected = 0;
    else if (alpha > 0.5) {
        double zn = z / eta;
        double C = (alpha - 1.) / (alpha + alpha - 1.);
        if (zn < C) isProjected = 0;
    }
    if (!isProjected) {
        dudx[0] = dudx[1] = dudx[2] = 0;
        dvdx[0] = dvdx[1] = dvdx[2] = 0;
        return 0;
    }
    double k = 1. / eta / eta;
    double abrho = alpha * beta / rho;
    double Jxy = k * abrho * x * y;
    double Jz = k * (gamma + alpha * z / rho);
    double Jx = gamma * z + alpha * rho;
    dudx[0] = fu * k * (Jx - abrho * x * x);
    dudx[1] = -fu * Jxy;
    dudx[2] = -fu * x * Jz;
    dvdx[0] = -fv * Jxy;
    dvdx[1] = fv * k * (Jx - abrho * y * y);
    dvdx[2] = -fv * y * Jz;
    return 1;
}

/* eucm.h:169-226 */
static int eucm_intrinsic_jacobian(const double *params, const double *src,
                                   double *du, double *dv)
{
    double alpha = params[0], beta = params[1], fu = params[2], fv = params[3];
    double x = src[0], y = src[1], z = src[2];

    double x2y2 = x * x + y * y;
    double rho2 = z * z + beta * (x2y2);
    double rho = sqrt(rho2);
    double gamma = 1. - alpha;
    double eta = alpha * rho + gamma * z;

    int isProjected = 1;
    if (eta < 1e-3) isProjected = 0;
    else if (alpha > 0.5) {
        double zn = z / eta;
        double C = (alpha - 1.) / (alpha + alpha - 1.);
        if (zn < C) isProjected = 0;
    }
    if (!isProjected) {
        for (int i = 0; i < 6; i++) { du[i] = 0; dv[i] = 0; }
        return 0;
    }
    double eta2 = eta * eta;
    du[0] = -fu * x * (rho - z) / eta2;
    du[1] = -fu * x * alpha * x2y2 / (2 * eta2 * rho);
    du[2] = x / eta;
    du[3] = 0;
    du[4] = 1;
    du[5] = 0;
    dv[0] = -fv * y * (rho - z) / eta2;
    dv[1] = -fv * y * alpha * x2y2 / (2 * eta2 * rho);
    dv[2] = 0;
    dv[3] = y / eta;
    dv[4] = 0;
    dv[5] = 1;
    return 1;
}

/* eucm.h:85-106 */
static int eucm_reconstruct(const double *params, const double *src, double *dst)
{
    double alpha = params[0], beta = params[1];
    double fu = params[2], fv = params[3], u0 = params[4], v0 = params[5];
    double xn = (src[0] - u0) / fu;
    double yn = (src[1] - v0) / fv;
    double u2 = xn * xn + yn * yn;
    double gamma = 1. - alpha;
    double num = 1. - u2 * alpha * alpha * beta;
    double det = 1 - (alpha - gamma) * beta * u2;
    if (det < 0) return 0;
    double denom = gamma + alpha * sqrt(det);
    dst[0] = xn; dst[1] = yn; dst[2] = num / denom;
    return 1;
}

/* UCM  ucm.h:35-59 */
static int ucm_project(const double *params, const double *src, double *dst)
{
    double xi = params[0], fu = params[1], fv = params[2], u0 = params[3], v0 = params[4];
    double x = src[0], y = src[1], z = src[2];
    double rho = sqrt(z * z + x * x + y * y);
    double denominv = 1. / (z + xi * rho);
    double xn = x * denominv;
    double yn = y * denominv;
    dst[0] = fu * xn + u0;
    dst[1] = fv * yn + v0;
    return 1;
}

/* the normalised-point Jacobian shared by ucm.h:131-139 and mei.h:151-158 */
static void ucm_norm_jac(double xi, const double *src, double *jm /*2x3*/,
                         double *pxn, double *pyn)
{
    double x = src[0], y = src[1], z = src[2];
    double xx = x * x, yy = y * y, zz = z * z;
    double rho = sqrt(xx + yy + zz);
    double rhoinv = 1. / rho;
    double deninv = 1. / (xi * rho + z);
    double deninv2 = deninv * deninv;
    *pxn = x * deninv;
    *pyn = y * deninv;
    jm[0] = (xi * rho + z - xi * xx * rhoinv) * deninv2;
    jm[1] = -xi * x * y * rhoinv * deninv2;
    jm[2] = -x * (1 + xi * z * rhoinv) * deninv2;
    jm[3] = -xi * x * y * rhoinv * deninv2;
    jm[4] = (xi * rho + z - xi * yy * rhoinv) * deninv2;
    jm[5] = -y * (1 + xi * z * rhoinv) * deninv2;
}

/* ucm.h:106-151 */
static int ucm_projection_jacobian(const double *params, const double *src,
                                   double *dudx, double *dvdx)
{
    double xi = params[0], fu = params[1], fv = params[2];
    double jm[6], xn, yn;
    ucm_norm_jac(xi, src, jm, &xn, &yn);
    for (int i = 0; i < 3; i++) {
        dudx[i] = fu * jm[i];
        dvdx[i] = fv * jm[3 + i];
    }
    return 1;
}

/* ucm.h:153-197 */
static int ucm_intrinsic_jacobian(const double *params, const double *src,
                                  double *du, double *dv)
{
    double xi = params[0], fu = params[1], fv = params[2];
    double x = src[0], y = src[1], z = src[2];
    double rho = sqrt(x * x + y * y + z * z);
    double deninv = 1. / (xi * rho + z);
    double xn = x * deninv;
    double yn = y * deninv;
    du[0] = -fu * xn * deninv * rho;
    du[1] = xn;
    du[2] = 0;
    du[3] = 1;
    du[4] = 0;
    dv[0] = -fv * yn * deninv * rho;
    dv[1] = 0;
    dv[2] = yn;
    dv[3] = 0;
    dv[4] = 1;
    return 1;
}

/* ucm.h:81-103 and mei.h:90-112 (same formula; MEI ignores distortion) */
static int ucm_reconstruct_(double xi, double fu, double fv, double u0, double v0,
                            const double *src, double *dst)
{
    double xn = (src[0] - u0) / fu;
    double yn = (src[1] - v0) / fv;
    double u2 = xn * xn + yn * yn;
    double gamma = sqrt(1. + u2 * (1 - xi * xi));
    double etanum = -gamma - xi * u2;
    double etadenom = xi * xi * u2 - 1;
    dst[0] = xn; dst[1] = yn; dst[2] = etadenom / (etadenom + xi * etanum);
    return 1;
}

/* MEI  mei.h:31-66 */
static int mei_project(const double *p, const double *src, double *dst)
{
    double xi = p[0], k1 = p[1], k2 = p[2], k3 = p[3], k4 = p[4], k5 = p[5];
    double fu = p[6], fv = p[7], u0 = p[8], v0 = p[9];
    double x = src[0], y = src[1], z = src[2];
    double rho = sqrt(z * z + x * x + y * y);
    double denominv = 1. / (z + xi * rho);
    double xn = x * denominv;
    double yn = y * denominv;
    double xx = xn * xn, xy = xn * yn, yy = yn * yn;
    double r2 = xx + yy;
    double D = 1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2;
    double deltax = 2. * k4 * xy + k5 * (r2 + 2. * xx);
    double deltay = 2. * k5 * xy + k4 * (r2 + 2. * yy);
    dst[0] = fu * (xn * D + deltax) + u0;
    dst[1] = fv * (yn * D + deltay) + v0;
    return 1;
}

/* mei.h:121-191 */
static int mei_projection_jacobian(const double *p, const double *src,
                                   double *dudx, double *dvdx)
{
    double xi = p[0], k1 = p[1], k2 = p[2], k3 = p[3], k4 = p[4], k5 = p[5];
    double fu = p[6], fv = p[7];
    double jm[6], xn, yn;
    ucm_norm_jac(xi, src, jm, &xn, &yn);
    double xxn = xn * xn, yyn = yn * yn, xyn = xn * yn;
    double r2 = yn * yn + xn * xn;
    double D = 1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2;
    double dDdr2 = k1 + 2 * k2 * r2 + 3 * k3 * r2 * r2;
    double du0 = D + 2 * xxn * dDdr2 + 2 * k4 * yn + 6 * k5 * xn;
    double du1 = 2 * xyn * dDdr2 + 2 * k4 * xn + 2 * k5 * yn;
    double dv0 = 2 * xyn * dDdr2 + 2 * k5 * yn + 2 * k4 * xn;
    double dv1 = D + 2 * yyn * dDdr2 + 2 * k5 * xn + 6 * k4 * yn;
    du0 *= fu; du1 *= fu;
    dv0 *= fv; dv1 *= fv;
    for (int i = 0; i < 3; i++) {
        dudx[i] = du0 * jm[i] + du1 * jm[3 + i];
        dvdx[i] = dv0 * jm[i] + dv1 * jm[3 + i];
    }
    return 1;
}

/* mei.h:193-285 */
static int mei_intrinsic_jacobian(const double *p, const double *src,
                                  double *du, double *dv)
{
    double xi = p[0], k1 = p[1], k2 = p[2], k3 = p[3], k4 = p[4], k5 = p[5];
    double fu = p[6], fv = p[7];
    double x = src[0], y = src[1], z = src[2];
    double rho = sqrt(x * x + y * y + z * z);
    double deninv = 1. / (xi * rho + z);
    double xn = x * deninv;
    double yn = y * deninv;
    double xxn = xn * xn, yyn = yn * yn, xyn = xn * yn;
    double r2 = yn * yn + xn * xn;
    double D = 1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2;
    double dDdr2 = k1 + 2 * k2 * r2 + 3 * k3 * r2 * r2;
    double deltax = 2. * k4 * xyn + k5 * (r2 + 2. * xxn);
    double deltay = 2. * k5 * xyn + k4 * (r2 + 2. * yyn);
    double xd = xn * D + deltax;
    double yd = yn * D + deltay;
    double du0 = D + 2 * xxn * dDdr2 + 2 * k4 * yn + 6 * k5 * xn;
    double du1 = 2 * xyn * dDdr2 + 2 * k4 * xn + 2 * k5 * yn;
    double dv0 = 2 * xyn * dDdr2 + 2 * k5 * yn + 2 * k4 * xn;
    double dv1 = D + 2 * yyn * dDdr2 + 2 * k5 * xn + 6 * k4 * yn;
    du0 *= fu; du1 *= fu;
    dv0 *= fv; dv1 *= fv;
    double dxndxi = -xn * deninv * rho;
    double dyndxi = -yn * deninv * rho;

    du[0] = du0 * dxndxi + du1 * dyndxi;
    du[1] = fu * xn * r2;
    du[2] = fu * xn * r2 * r2;
    du[3] = fu * xn * r2 * r2 * r2;
    du[4] = 2. * fu * xyn;
    du[5] = fu * (r2 + 2. * xxn);
    du[6] = xd;
    du[7] = 0;
    du[8] = 1;
    du[9] = 0;

    dv[0] = dv0 * dxndxi + dv1 * dyndxi;
    dv[1] = fv * yn * r2;
    dv[2] = fv * yn * r2 * r2;
    dv[3] = fv * yn * r2 * r2 * r2;
    dv[4] = fv * (r2 + 2. * yyn);
    dv[5] = 2. * fv * xyn;
    dv[6] = 0;
    dv[7] = yd;
    dv[8] = 0;
    dv[9] = 1;
    return 1;
}

/* ---- virtual dispatch of ICamera (generic_camera.h:36-50) ---- */
int vgo_project(int model, const double *params, const double X[3], double uv[2])
{
    switch (model) {
    case VGO_EUCM: return eucm_project(params, X, uv);
    case VGO_UCM:  return ucm_project(params, X, uv);
    default:       return mei_project(params, X, uv);
    }
}

int vgo_projection_jacobian(int model, const double *params, const double X[3],
                            double dudx[3], double dvdx[3])
{
    switch (model) {
    case VGO_EUCM: return eucm_projection_jacobian(params, X, dudx, dvdx);
    case VGO_UCM:  return ucm_projection_jacobian(params, X, dudx, dvdx);
    default:       return mei_projection_jacobian(params, X, dudx, dvdx);
    }
}

int vgo_intrinsic_jacobian(int model, const double *params, const double X[3],
                           double *du, double *dv)
{
    switch (model) {
    case VGO_EUCM: return eucm_intrinsic_jacobian(params, X, du, dv);
    case VGO_UCM:  return ucm_intrinsic_jacobian(params, X, du, dv);
    default:       return mei_intrinsic_jacobian(params, X, du, dv);
    }
}

int vgo_reconstruct(int model, const double *p, const double uv[2], double X[3])
{
    switch (model) {
    case VGO_EUCM: return eucm_reconstruct(p, uv, X);
    case VGO_UCM:  return ucm_reconstruct_(p[0], p[1], p[2], p[3], p[4], uv, X);
    default:       return ucm_reconstruct_(p[0], p[6], p[7], p[8], p[9], uv, X);
    }
}

/* ------------------------------------------------------------------ */
/* jacobian.h:139-152 */
void vgo_inter_jacobian_init(vgo_inter_jacobian *ij, int model, const double *params,
                             const double xi13[6], const double xi23[6], int inverted)
{
    double R13[9], R23inv[9], nr[3], M[9];
    ij->model = model;
    ij->params = params;
    vgo_rotation_matrix(xi13 + 3, R13);
    nr[0] = -xi23[3]; nr[1] = -xi23[4]; nr[2] = -xi23[5];   /* rotMatInv, transformation.h:132 */
    vgo_rotation_matrix(nr, R23inv);
    mat3_mul(R13, R23inv, ij->R12);
    ij->t13[0] = xi13[0]; ij->t13[1] = xi13[1]; ij->t13[2] = xi13[2];
    vgo_inter_omega_rot(xi23 + 3, M);
    mat3_mul(ij->R12, M, ij->M12);
    if (inverted) {
        for (int i = 0; i < 9; i++) { ij->R12[i] *= -1; ij->M12[i] *= -1; }
    }
}

/* jacobian.h:155-171 */
void vgo_dpdxi(const vgo_inter_jacobian *ij, const double X1[3], double dudxi[6], double dvdxi[6])
{
    double pj[6];
    vgo_projection_jacobian(ij->model, ij->params, X1, pj, pj + 3);
    double t3X[3] = { X1[0] - ij->t13[0], X1[1] - ij->t13[1], X1[2] - ij->t13[2] };
    double H[9], nrow[3], tmp[3];
    hat3(t3X, H);

    row_mat3(pj, ij->R12, dudxi);
    nrow[0] = -pj[0]; nrow[1] = -pj[1]; nrow[2] = -pj[2];
    row_mat3(nrow, H, tmp);
    row_mat3(tmp, ij->M12, dudxi + 3);

    row_mat3(pj + 3, ij->R12, dvdxi);
    nrow[0] = -pj[3]; nrow[1] = -pj[4]; nrow[2] = -pj[5];
    row_mat3(nrow, H, tmp);
    row_mat3(tmp, ij->M12, dvdxi + 3);
}

/* ------------------------------------------------------------------ */
/* calib_cost_functions.cpp:28-117 */
int vgo_evaluate(int model, int P, const double *obs, const double *board,
                 int chain_len, const int *status,
                 double const *const *params, double *residual, double **jacobian)
{
    const int K = vgo_num_params(model);
    /* :32-46 chain */
    double xiAcc[6] = { 0, 0, 0, 0, 0, 0 }, tmp[6];
    for (int paramIdx = 1; paramIdx <= chain_len; paramIdx++) {
        if (status[paramIdx - 1] == VGO_TRANSFORM_DIRECT) {
            vgo_compose(xiAcc, params[paramIdx], tmp);
            memcpy(xiAcc, tmp, sizeof tmp);
        } else if (status[paramIdx - 1] == VGO_TRANSFORM_INVERSE) {
            vgo_compose_inverse(xiAcc, params[paramIdx], tmp);
            memcpy(xiAcc, tmp, sizeof tmp);
        }
    }
    /* :49-50 points in the camera frame (heap vector in the reference) */
    double *pointCam = (double *)malloc(sizeof(double) * 3 * (size_t)(P > 0 ? P : 1));
    {
        double R[9], o[3];
        vgo_rotation_matrix(xiAcc + 3, R);                   /* transformation.h:185-193 */
        for (int i = 0; i < P; i++) {
            mat3_vec(R, board + 3 * i, o);
            pointCam[3 * i] = o[0] + xiAcc[0];               /* transformation.h:151-154 */
            pointCam[3 * i + 1] = o[1] + xiAcc[1];
            pointCam[3 * i + 2] = o[2] + xiAcc[2];
        }
    }
    /* :54 setParameters copies the K doubles */
    double camParams[16];
    memcpy(camParams, params[0], sizeof(double) * (size_t)K);

    /* :57-71 */
    for (int i = 0; i < P; i++) {
        double modProj[2];
        if (vgo_project(model, camParams, pointCam + 3 * i, modProj)) {
            residual[2 * i] = modProj[0] - obs[2 * i];
            residual[2 * i + 1] = modProj[1] - obs[2 * i + 1];
        } else {
            residual[2 * i] = VGO_DOUBLE_BIG;
            residual[2 * i + 1] = VGO_DOUBLE_BIG;
        }
    }

    if (jacobian != NULL) {
        /* :76-92 chain again */
        double acc[6] = { 0, 0, 0, 0, 0, 0 };
        for (int paramIdx = 1; paramIdx <= chain_len; paramIdx++) {
            const int st = status[paramIdx - 1];
            const double *xi23 = params[paramIdx];
            double xi13[6];
            if (st == VGO_TRANSFORM_DIRECT) {
                vgo_compose(acc, xi23, tmp);
                memcpy(acc, tmp, sizeof tmp);
                memcpy(xi13, acc, sizeof acc);
            } else {
                memcpy(xi13, acc, sizeof acc);
                vgo_compose_inverse(acc, xi23, tmp);
                memcpy(acc, tmp, sizeof tmp);
            }
            if (jacobian[paramIdx] != NULL) {
                /* :95 ; the ctor clones the camera (jacobian.h:141) */
                double *clone = (double *)malloc(sizeof(double) * 16);
                memcpy(clone, camParams, sizeof(double) * (size_t)K);
                vgo_inter_jacobian ij;
                vgo_inter_jacobian_init(&ij, model, clone, xi13, xi23, st == VGO_TRANSFORM_INVERSE);
                for (int i = 0; i < P; i++)
                    vgo_dpdxi(&ij, pointCam + 3 * i, jacobian[paramIdx] + i * 12,
                              jacobian[paramIdx] + i * 12 + 6);
                free(clone);
            }
        }
        /* :105-114 */
        if (jacobian[0] != NULL) {
            for (int i = 0; i < P; i++)
                vgo_intrinsic_jacobian(model, camParams, pointCam + 3 * i,
                                       jacobian[0] + i * 2 * K, jacobian[0] + (i * 2 + 1) * K);
        }
    }
    free(pointCam);
    return 1;
}

/* ------------------------------------------------------------------ */
/* 6x6 row-major helpers for the prior functors */
static void mat6_mul(const double A[36], const double B[36], double C[36])
{
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = A[6 * i] * B[j];
            for (int k = 1; k < 6; k++) s += A[6 * i + k] * B[6 * k + j];
            C[6 * i + j] = s;
        }
}

static void mat6_vec(const double A[36], const double v[6], double o[6])
{
    for (int i = 0; i < 6; i++) {
        double s = A[6 * i] * v[0];
        for (int k = 1; k < 6; k++) s += A[6 * i + k] * v[k];
        o[i] = s;
    }
}

/* [[a, 0], [0, b]] with 3x3 blocks */
static void blockdiag6(const double a[9], const double b[9], double M[36])
{
    memset(M, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { M[6 * i + j] = a[3 * i + j]; M[6 * (i + 3) + j + 3] = b[3 * i + j]; }
}

/* calib_cost_functions.h:85-98 */
void vgo_transformation_prior_init(vgo_transformation_prior *tp, const double stiffness[6], const double xi_prior[6])
{
    memcpy(tp->xi_prior, xi_prior, 6 * sizeof(double));
    memset(tp->A, 0, sizeof tp->A);
    vgo_rotation_matrix(xi_prior + 3, tp->R);
    for (int i = 0; i < 6; i++) tp->A[6 * i + i] = stiffness[i];
    double M[9], D[9], DM[9];
    vgo_inter_omega_rot(xi_prior + 3, M);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) D[3 * i + j] = tp->A[6 * (i + 3) + j + 3];
    mat3_mul(D, M, DM);                                     /* bottom-right = bottom-right * M, :95-96 */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) tp->A[6 * (i + 3) + j + 3] = DM[3 * i + j];
}

/* calib_cost_functions.cpp:215-228 */
void vgo_transformation_prior_eval(const vgo_transformation_prior *tp, const double xi[6], double r[6], double *J)
{
    double e[6], err[6];
    vgo_inverse_compose(tp->xi_prior, xi, e);               /* :220 */
    mat3_vec(tp->R, e, err);                                /* :221 */
    mat3_vec(tp->R, e + 3, err + 3);                        /* :222 */
    mat6_vec(tp->A, err, r);                                /* :223 */
    if (J) memcpy(J, tp->A, 36 * sizeof(double));           /* :224-227: the Jacobian is A itself */
}

/* calib_cost_functions.cpp:119-173 */
static void odometry_information(vgo_odometry_prior *op, double errV, double errW, double lambda);

void vgo_odometry_prior_init(vgo_odometry_prior *op, double errV, double errW, double lambda,
                             const double xi1[6], const double xi2[6])
{
    vgo_inverse_compose(xi1, xi2, op->zeta_prior);          /* :122 */
    odometry_information(op, errV, errW, lambda);
}

/* the information matrix _A from the prior motion: calib_cost_functions.cpp:124-172 and, word for word the same,
 * odometry_cost_function.cpp:155-196 */
static void odometry_information(vgo_odometry_prior *op, double errV, double errW, double lambda)
{
    const double MIN_SIGMA_V = 0.01, MIN_SIGMA_W = 0.01, MIN_DELTA = 0.01, MIN_L = 0.01;
    memset(op->A, 0, sizeof op->A);
    const double delta = fmax(norm3(op->zeta_prior + 3), MIN_DELTA);
    const double l = fmax(norm3(op->zeta_prior), MIN_L);
    const double delta2 = delta / 2., l2 = l / 2.;
    const double s = sin(delta2), c = cos(delta2);
    const double dfdu[3][2] = { { c, l2 * s }, { -s, l2 * c }, { 0, 1 } };     /* :142-145 */
    double Cu[2];
    Cu[0] = fmax(errV * errV * l * l, MIN_SIGMA_V * MIN_SIGMA_V);              /* :147-151 */
    Cu[1] = fmax(errW * errW * delta * delta, MIN_SIGMA_W * MIN_SIGMA_W);
    /* Cx = dfdu Cu dfdu^T + lambda^2 I, evaluated left to right as (dfdu*Cu)*dfdu^T, :153 */
    double Cx[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double a0 = dfdu[i][0] * Cu[0] + dfdu[i][1] * 0.;
            const double a1 = dfdu[i][0] * 0. + dfdu[i][1] * Cu[1];
            Cx[3 * i + j] = (a0 * dfdu[j][0] + a1 * dfdu[j][1]) + (lambda * lambda) * (i == j ? 1. : 0.);
        }
    /* 3x3 inverse by cofactors (what Eigen's fixed-size inverse computes), :154 */
    double cof[3][3], inv[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            cof[i][j] = Cx[3 * i1 + j1] * Cx[3 * i2 + j2] - Cx[3 * i1 + j2] * Cx[3 * i2 + j1];
        }
    const double det = Cx[0] * cof[0][0] + Cx[1] * cof[0][1] + Cx[2] * cof[0][2];
    const double id = 1. / det;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) inv[3 * i + j] = cof[j][i] * id;
    /* LLT, U = L^T, :155-156 */
    double Lc[9] = { 0 };
    for (int j = 0; j < 3; j++) {
        double d = inv[3 * j + j];
        for (int k = 0; k < j; k++) d -= Lc[3 * j + k] * Lc[3 * j + k];
        Lc[3 * j + j] = sqrt(d);
        for (int i = j + 1; i < 3; i++) {
            double t = inv[3 * i + j];
            for (int k = 0; k < j; k++) t -= Lc[3 * i + k] * Lc[3 * j + k];
            Lc[3 * i + j] = t / Lc[3 * j + j];
        }
    }
#define U_(i, j) Lc[3 * (j) + (i)]
    op->A[0] = U_(0, 0); op->A[1] = U_(0, 1);               /* topLeftCorner<2,2>, :157 */
    op->A[6] = U_(1, 0); op->A[7] = U_(1, 1);
    op->A[5] = U_(0, 2); op->A[11] = U_(1, 2);              /* topRightCorner<2,1> of the 6x6: column 5, :158 */
    op->A[14] = 1. / lambda;                                /* (2,2), :159 */
    op->A[6 * 3 + 3] = 1. / lambda;                         /* B, :161-166 */
    op->A[6 * 4 + 4] = 1. / lambda;
    op->A[6 * 5 + 5] = U_(2, 2);
#undef U_
}

/* calib_cost_functions.cpp:177-213 */
void vgo_odometry_prior_eval(const vgo_odometry_prior *op, const double xi1[6], const double xi2[6],
                             double r[6], double *J1, double *J2)
{
    double zeta[6], err[6];
    vgo_inverse_compose(xi1, xi2, zeta);                    /* :182 */
    vgo_inverse_compose(op->zeta_prior, zeta, err);         /* :187 */
    mat6_vec(op->A, err, r);                                /* :188 */
    if (J1) {                                               /* :193-203 */
        double neg[3] = { -xi1[3], -xi1[4], -xi1[5] }, R10[9], M[9], R10M[9], Jb[36];
        vgo_rotation_matrix(neg, R10);
        vgo_inter_omega_rot(xi1 + 3, M);
        mat3_mul(R10, M, R10M);
        blockdiag6(R10, R10M, Jb);
        /* zeta.screwTransfInv(), transformation.h:234-243 */
        double nz[3] = { -zeta[3], -zeta[4], -zeta[5] }, Rz[9], th[9], Rth[9], TT[36];
        vgo_rotation_matrix(nz, Rz);
        hat3(zeta, th);
        double nRz[9];
        for (int i = 0; i < 9; i++) nRz[i] = -Rz[i];
        mat3_mul(nRz, th, Rth);
        memset(TT, 0, sizeof TT);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                TT[6 * i + j] = Rz[3 * i + j];
                TT[6 * i + j + 3] = Rth[3 * i + j];
                TT[6 * (i + 3) + j + 3] = Rz[3 * i + j];
            }
        double nA[36], nATT[36];
        for (int i = 0; i < 36; i++) nA[i] = -op->A[i];
        mat6_mul(nA, TT, nATT);                             /* (-_A * TT) * J1, left to right */
        mat6_mul(nATT, Jb, J1);
    }
    if (J2) {                                               /* :206-213 */
        double neg[3] = { -xi2[3], -xi2[4], -xi2[5] }, R20[9], M[9], R20M[9], Jb[36];
        vgo_rotation_matrix(neg, R20);
        vgo_inter_omega_rot(xi2 + 3, M);
        mat3_mul(R20, M, R20M);
        blockdiag6(R20, R20M, Jb);
        mat6_mul(op->A, Jb, J2);
    }
}

/* ------------------------------------------------------------------ */
/* OdometryCost, src/calibration/odometry_cost_function.cpp (the third residual block of the calibration,
 * "odometry_intrinsic" datasets): the motion between two poses of a differential-drive platform, integrated from m
 * pairs of wheel-angle increments (left, right) with the odometry intrinsics (r1, r2 wheel radii, g track gauge),
 * compared with the motion xi1^-1 o xi2. */

/* tf0n_jac_calc, :68-88, one step: zeta_i = (v, 0, w) from the increments (odom_zeta_i, :10-35) and its Jacobian
 * w.r.t. (r1, r2, g) (zeta_i_jacobian, :38-65) */
static void odometry_step(const double dq[2], const double in[3], double zeta_i[3], double jac[9])
{
    const double r1 = in[0], r2 = in[1], g = in[2];
    const double j2c[4] = { r1 / 2, r2 / 2, -(r1 / g), r2 / g };
    const double v = j2c[0] * dq[0] + j2c[1] * dq[1], w = j2c[2] * dq[0] + j2c[3] * dq[1];
    /* A = [1 0; 0 0; 0 1], zeta_i = A * v_w: every entry is a two-term product sum */
    zeta_i[0] = 1 * v + 0 * w; zeta_i[1] = 0 * v + 0 * w; zeta_i[2] = 0 * v + 1 * w;
    jac[0] = dq[0] / 2; jac[1] = dq[1] / 2; jac[2] = 0;
    jac[3] = 0; jac[4] = 0; jac[5] = 0;
    jac[6] = -dq[0] / g; jac[7] = dq[1] / g; jac[8] = (r1 * dq[0] - r2 * dq[1]) / (g * g);
}

/* tf0n_jac_calc (:68-88) + calc_acc (:90-141): zeta_odo = 0Tn and d(zeta_odo's x, y, theta) / d(r1, r2, g) as the 6 x 3
 * matrix calc_acc returns (rows x, y, -, -, -, theta).  tf = m transforms of scratch. */
static void odometry_integrate(int m, const double *dq, const double in[3], double *tf, double zeta_odo[6], double jac6x3[18])
{
    double cur[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < m; i++) {
        double z[3], jz[9], step[6], nxt[6];
        odometry_step(dq + 2 * i, in, z, jz);
        step[0] = z[0]; step[1] = z[1]; step[2] = 0; step[3] = 0; step[4] = 0; step[5] = z[2];
        vgo_compose(cur, step, nxt);                        /* :82 */
        memcpy(cur, nxt, sizeof cur);
        memcpy(tf + 6 * (size_t)i, cur, sizeof cur);
    }
    memcpy(zeta_odo, tf + 6 * (size_t)(m - 1), 6 * sizeof(double));
    if (!jac6x3) return;
    double ACC[9] = { 0 };
    for (int i = 0; i < m; i++) {
        double z[3], jz[9], prev[6] = { 0, 0, 0, 0, 0, 0 }, R0j[9];
        odometry_step(dq + 2 * i, in, z, jz);
        if (i > 0) memcpy(prev, tf + 6 * (size_t)(i - 1), sizeof prev);
        vgo_rotation_matrix(prev + 3, R0j);                 /* :110 */
        /* tfi_n = tf0_i.inverse().compose(tf0_n), :112-114; inverse() = (-R^T t, -rot), transformation.h:112-119 */
        const double *t0i = tf + 6 * (size_t)i;
        double neg[3] = { -t0i[3], -t0i[4], -t0i[5] }, Rinv[9], inv[6], tin[6];
        vgo_rotation_matrix(neg, Rinv);
        double nR[9];
        for (int k = 0; k < 9; k++) nR[k] = -Rinv[k];
        mat3_vec(nR, t0i, inv);
        inv[3] = neg[0]; inv[4] = neg[1]; inv[5] = neg[2];
        vgo_compose(inv, zeta_odo, tin);
        const double J[9] = { 1, 0, -tin[1], 0, 1, tin[0], 0, 0, 1 };       /* :119-122 */
        double RJ[9], RJz[9];
        mat3_mul(R0j, J, RJ);
        mat3_mul(RJ, jz, RJz);
        for (int k = 0; k < 9; k++) ACC[k] = ACC[k] + RJz[k];               /* :124 */
    }
    memset(jac6x3, 0, 18 * sizeof(double));
    for (int j = 0; j < 3; j++) { jac6x3[j] = ACC[j]; jac6x3[3 + j] = ACC[3 + j]; jac6x3[15 + j] = ACC[6 + j]; }
}

/* the constructor, :144-197: the prior motion from the prior intrinsics, then _A.  Returns 0, or -1 without increments. */
int vgo_odometry_cost_init(vgo_odometry_prior *oc, double errV, double errW, double lambda, int m, const double *dq,
                           const double intr_prior[3])
{
    if (m < 1) return -1;
    double *tf = (double *)malloc(sizeof(double) * 6 * (size_t)m);
    odometry_integrate(m, dq, intr_prior, tf, oc->zeta_prior, NULL);
    free(tf);
    odometry_information(oc, errV, errW, lambda);
    return 0;
}

/* Evaluate, :202-267.  J1, J2 6 x 6, J3 6 x 3, row-major; any may be NULL. */
void vgo_odometry_cost_eval(const vgo_odometry_prior *oc, int m, const double *dq, const double xi1[6], const double xi2[6],
                            const double intr[3], double r[6], double *J1, double *J2, double *J3)
{
    double *tf = (double *)malloc(sizeof(double) * 6 * (size_t)m);
    double zeta[6], zeta_odo[6], jin[18], delta[6];
    vgo_inverse_compose(xi1, xi2, zeta);                    /* :209 */
    odometry_integrate(m, dq, intr, tf, zeta_odo, jin);     /* :218-219 */
    free(tf);
    vgo_inverse_compose(zeta_odo, zeta, delta);             /* :224 */
    mat6_vec(oc->A, delta, r);                              /* :226 */
    /* the first two blocks are OdometryPrior's with the integrated motion in the prior's place, :231-253 */
    vgo_odometry_prior tmp = *oc;
    memcpy(tmp.zeta_prior, zeta_odo, sizeof tmp.zeta_prior);
    if (J1 || J2) { double r2[6]; vgo_odometry_prior_eval(&tmp, xi1, xi2, r2, J1, J2); }
    if (J3) {                                               /* :256-264 */
        double neg[3] = { -zeta_odo[3], -zeta_odo[4], -zeta_odo[5] }, R31[9], M[9], R31M[9], Jb[36];
        vgo_rotation_matrix(neg, R31);
        vgo_inter_omega_rot(zeta_odo + 3, M);
        mat3_mul(R31, M, R31M);
        blockdiag6(R31, R31M, Jb);
        double nd[3] = { -delta[3], -delta[4], -delta[5] }, Rd[9], th[9], nRd[9], Rth[9], TT[36];
        vgo_rotation_matrix(nd, Rd);
        hat3(delta, th);
        for (int i = 0; i < 9; i++) nRd[i] = -Rd[i];
        mat3_mul(nRd, th, Rth);
        memset(TT, 0, sizeof TT);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                TT[6 * i + j] = Rd[3 * i + j];
                TT[6 * i + j + 3] = Rth[3 * i + j];
                TT[6 * (i + 3) + j + 3] = Rd[3 * i + j];
            }
        double nA[36], a[36], b[36];
        for (int i = 0; i < 36; i++) nA[i] = -oc->A[i];
        mat6_mul(nA, TT, a);                                /* ((-_A * TT) * J3) * jac_intrinsic, left to right */
        mat6_mul(a, Jb, b);
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 3; j++) {
                double s = b[6 * i] * jin[j];
                for (int k = 1; k < 6; k++) s += b[6 * i + k] * jin[3 * k + j];
                J3[3 * i + j] = s;
            }
    }
}

/* ------------------------------------------------------------------ */
/* TrajectoryVisualQuality::visualCov, trajectory_generation.cpp:185-206: covariance of the camera pose
 * localised on the board, from dP/dX of every board point.  feature_variance -> _ptStiffness = U of the Cholesky
 * factorisation of diag(1 / variance) (:112-115).  out: 6 x 6 row-major per pose. */
void vgo_visual_cov(int model, const double *intr, const double xi_board[6], int P, const double *board,
                    double feature_variance, int n, const double *cam_poses, double *out)
{
    const double stiff = sqrt(1. / feature_variance);                 /* U(0,0) = U(1,1); U(0,1) = 0 */
    for (int k = 0; k < n; k++) {
        double xcb[6], R[9];
        vgo_inverse_compose(cam_poses + 6 * (size_t)k, xi_board, xcb);  /* :187 */
        vgo_rotation_matrix(xcb + 3, R);
        double JtJ[36], JtCJ[36];
        memset(JtJ, 0, sizeof JtJ); memset(JtCJ, 0, sizeof JtCJ);
        for (int i = 0; i < P; i++) {
            double X[3], o[3], pj[6], hx[9], d[2][6];
            mat3_vec(R, board + 3 * i, o);
            X[0] = o[0] + xcb[0]; X[1] = o[1] + xcb[1]; X[2] = o[2] + xcb[2];
            vgo_projection_jacobian(model, intr, X, pj, pj + 3);       /* :195 (zero rows when the projection fails) */
            hat3(X, hx);
            for (int q = 0; q < 2; q++)                                 /* dpdx * [-I | hat(X)], :197-200 */
                for (int j = 0; j < 3; j++) {
                    d[q][j] = -pj[3 * q + j];
                    d[q][3 + j] = pj[3 * q] * hx[j] + pj[3 * q + 1] * hx[3 + j] + pj[3 * q + 2] * hx[6 + j];
                }
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) {
                    JtJ[6 * a + b] += d[0][a] * d[0][b] + d[1][a] * d[1][b];                        /* :201 */
                    JtCJ[6 * a + b] += (d[0][a] * stiff) * d[0][b] + (d[1][a] * stiff) * d[1][b];   /* :202 */
                }
        }
        /* JtCJ^-1 by LU with partial pivoting (Eigen's fixed-size inverse beyond 4x4), :204 */
        double lu[6][6], inv[36];
        int perm[6];
        for (int i = 0; i < 6; i++) { perm[i] = i; for (int j = 0; j < 6; j++) lu[i][j] = JtCJ[6 * i + j]; }
        for (int c = 0; c < 6; c++) {
            int piv = c;
            for (int i = c + 1; i < 6; i++) if (fabs(lu[i][c]) > fabs(lu[piv][c])) piv = i;
            if (piv != c) {
                for (int j = 0; j < 6; j++) { double t = lu[c][j]; lu[c][j] = lu[piv][j]; lu[piv][j] = t; }
                int t = perm[c]; perm[c] = perm[piv]; perm[piv] = t;
            }
            for (int i = c + 1; i < 6; i++) {
                lu[i][c] /= lu[c][c];
                for (int j = c + 1; j < 6; j++) lu[i][j] -= lu[i][c] * lu[c][j];
            }
        }
        for (int c = 0; c < 6; c++) {
            double x[6];
            for (int i = 0; i < 6; i++) {
                double v = perm[i] == c ? 1. : 0.;
                for (int j = 0; j < i; j++) v -= lu[i][j] * x[j];
                x[i] = v;
            }
            for (int i = 5; i >= 0; i--) {
                double v = x[i];
                for (int j = i + 1; j < 6; j++) v -= lu[i][j] * x[j];
                x[i] = v / lu[i][i];
            }
            for (int i = 0; i < 6; i++) inv[6 * i + c] = x[i];
        }
        /* inv^T * JtJ * inv, :205 */
        double invT[36], tmp[36];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) invT[6 * i + j] = inv[6 * j + i];
        mat6_mul(invT, JtJ, tmp);
        mat6_mul(tmp, inv, out + 36 * (size_t)k);
    }
}

int vgo_hessian_entries(int K, int chain_len)
{
    int D = K + 6 * chain_len;
    return (D + 1) * (D + 2) / 2;
}

int vgo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Per-image normal-equation block: packed upper triangle of [J r]^T [J r],
 * column order [intr, e0, .., r] -- the dense product Ceres forms when it builds
 * J^T J / J^T r from a residual block. */
static void image_hessian(int P, int K, int L, const double *r, const double *Ja,
                          double *const *Je, double *H)
{
    const int D = K + 6 * L;
    const int W = D + 1;
    double row[64];
    const int ne = W * (W + 1) / 2;
    for (int i = 0; i < ne; i++) H[i] = 0.0;
    for (int k = 0; k < 2 * P; k++) {
        for (int c = 0; c < K; c++) row[c] = Ja[(size_t)k * K + c];
        for (int e = 0; e < L; e++)
            for (int c = 0; c < 6; c++) row[K + 6 * e + c] = Je[e][(size_t)k * 6 + c];
        row[D] = r[k];
        int idx = 0;
        for (int a = 0; a < W; a++)
            for (int b = a; b < W; b++)
                H[idx++] += row[a] * row[b];
    }
}

int vgo_evaluate_batch(int model, const double *intr, int n_img, int P,
                       const double *board, const double *obs,
                       int chain_len, const int *status, const int *is_global,
                       const double *const *xi,
                       double *r, double *J_intr, double *const *J_xi, double *H,
                       int threads)
{
    const int K = vgo_num_params(model);
    if (K < 0 || chain_len < 0 || chain_len > 5) return -1;
    const int L = chain_len;
    const int ne = vgo_hessian_entries(K, L);
    const int want_J = (J_intr != NULL) || (J_xi != NULL) || (H != NULL);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int img = 0; img < n_img; img++) {
        const double *params[6];
        double *jac[6];
        double *scratch = NULL;
        double *rr;
        params[0] = intr;
        for (int e = 0; e < L; e++)
            params[1 + e] = xi[e] + (is_global[e] ? 0 : (size_t)img * 6);
        /* scratch for outputs the caller did not ask for but H needs */
        size_t need = 0;
        if (r == NULL) need += (size_t)2 * P;
        if (H != NULL) {
            if (J_intr == NULL) need += (size_t)2 * P * K;
            for (int e = 0; e < L; e++)
                if (J_xi == NULL || J_xi[e] == NULL) need += (size_t)2 * P * 6;
        }
        if (need) scratch = (double *)malloc(sizeof(double) * need);
        double *sp = scratch;
        if (r != NULL) rr = r + (size_t)img * 2 * P; else { rr = sp; sp += (size_t)2 * P; }
        if (J_intr != NULL) jac[0] = J_intr + (size_t)img * 2 * P * K;
        else if (H != NULL) { jac[0] = sp; sp += (size_t)2 * P * K; }
        else jac[0] = NULL;
        for (int e = 0; e < L; e++) {
            if (J_xi != NULL && J_xi[e] != NULL) jac[1 + e] = J_xi[e] + (size_t)img * 2 * P * 6;
            else if (H != NULL) { jac[1 + e] = sp; sp += (size_t)2 * P * 6; }
            else jac[1 + e] = NULL;
        }
        vgo_evaluate(model, P, obs + (size_t)img * 2 * P, board, L, status,
                     params, rr, want_J ? jac : NULL);
        if (H != NULL)
            image_hessian(P, K, L, rr, jac[0], jac + 1, H + (size_t)img * ne);
        free(scratch);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Batch loops for the CPU-baseline timings of the rows around the hot path (tools/aux_timing.py): what a caller of the
 * reference does per point / per block, over n of them, OpenMP over the items. */

/* ICamera::projectPoint + projectionJacobian + intrinsicJacobian per point (generic_camera.h:39-50) */
void vgo_project_batch(int model, const double *params, long n, const double *X, double *uv, double *dPdX, double *dPdintr,
                       unsigned char *ok, int threads)
{
    const int K = vgo_num_params(model);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (long i = 0; i < n; i++) {
        ok[i] = (unsigned char)vgo_project(model, params, X + 3 * i, uv + 2 * i);
        vgo_projection_jacobian(model, params, X + 3 * i, dPdX + 6 * i, dPdX + 6 * i + 3);
        vgo_intrinsic_jacobian(model, params, X + 3 * i, dPdintr + 2 * (size_t)K * i, dPdintr + 2 * (size_t)K * i + K);
    }
}

/* TransformationPrior / OdometryPrior: functors built once (init), then Evaluate over all blocks -- the timed part */
void vgo_transformation_prior_batch(int n, const double *stiffness, const double *xi_prior, const double *xi, double *r, double *J,
                                    int reps, int threads)
{
    vgo_transformation_prior *tp = (vgo_transformation_prior *)malloc(sizeof(*tp) * (size_t)n);
    for (int i = 0; i < n; i++) vgo_transformation_prior_init(tp + i, stiffness + 6 * i, xi_prior + 6 * i);
    (void)threads;
    for (int k = 0; k < reps; k++) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
        for (int i = 0; i < n; i++) vgo_transformation_prior_eval(tp + i, xi + 6 * i, r + 6 * i, J + 36 * (size_t)i);
    }
    free(tp);
}

void vgo_odometry_prior_batch(int n, double errV, double errW, double lambda, const double *o1, const double *o2, const double *x1,
                              const double *x2, double *r, double *J1, double *J2, int reps, int threads)
{
    vgo_odometry_prior *op = (vgo_odometry_prior *)malloc(sizeof(*op) * (size_t)n);
    for (int i = 0; i < n; i++) vgo_odometry_prior_init(op + i, errV, errW, lambda, o1 + 6 * i, o2 + 6 * i);
    (void)threads;
    for (int k = 0; k < reps; k++) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
        for (int i = 0; i < n; i++) vgo_odometry_prior_eval(op + i, x1 + 6 * i, x2 + 6 * i, r + 6 * i, J1 + 36 * (size_t)i, J2 + 36 * (size_t)i);
    }
    free(op);
}

void vgo_odometry_cost_batch(int n, double errV, double errW, double lambda, const int *dq_offset, const double *dq,
                             const double *intr_prior, const double *x1, const double *x2, const double *intr, double *r,
                             double *J1, double *J2, double *J3, int reps, int threads)
{
    vgo_odometry_prior *oc = (vgo_odometry_prior *)malloc(sizeof(*oc) * (size_t)n);
    for (int i = 0; i < n; i++)
        vgo_odometry_cost_init(oc + i, errV, errW, lambda, dq_offset[i + 1] - dq_offset[i], dq + 2 * (size_t)dq_offset[i], intr_prior);
    (void)threads;
    for (int k = 0; k < reps; k++) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
        for (int i = 0; i < n; i++)
            vgo_odometry_cost_eval(oc + i, dq_offset[i + 1] - dq_offset[i], dq + 2 * (size_t)dq_offset[i], x1 + 6 * i, x2 + 6 * i, intr,
                                   r + 6 * i, J1 + 36 * (size_t)i, J2 + 36 * (size_t)i, J3 + 18 * (size_t)i);
    }
    free(oc);
}
