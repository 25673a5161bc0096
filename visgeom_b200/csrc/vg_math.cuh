// vg_math.cuh -- fp64 device math of the reprojection path (sm_100a).
//
// Re-derived for the GPU, not transcribed: every per-corner quantity (rho, eta and
// their reciprocals) is computed once and shared by the projection, dP/dX and
// dP/dintr (the reference recomputes them in three virtual calls, eucm.h:45,
// 128-130,184-187), and the transform chain is accumulated with rotation matrices
// instead of quaternion round trips (transformation.h:80-110).  Mathematical
// definitions followed:
//   rotationMatrix      include/geometry/geometry_core.h:40-76
//   interOmegaRot       include/geometry/geometry_core.h:158-180
//   EUCM                include/projection/eucm.h:32-63,115-226
//   UCM                 include/projection/ucm.h:35-59,106-197
//   MEI                 include/projection/mei.h:31-66,121-285
#pragma once

#include <cuda_runtime.h>

#include <type_traits>

namespace vg {

// loops whose indices must be compile-time constants (register arrays: a loop the compiler declines to unroll would send
// the whole array to local memory -- which is what "#pragma unroll" on the 6x6 Cholesky below used to end in)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
#define VG_IDX(c) decltype(c)::value


constexpr int MODEL_EUCM = 0;
constexpr int MODEL_UCM = 1;
constexpr int MODEL_MEI = 2;
constexpr double DOUBLE_BIG = 1e15;  // include/std.h:71

// R = exp(hat(r)) and the SO(3) left Jacobian Jl (omega = Jl * rdot), row-major.
// Small-angle switch at 1e-5 exactly as geometry_core.h:44,162.
__device__ __forceinline__ void rodrigues_and_left_jacobian(const double r0, const double r1, const double r2,
                                                            double (&R)[9], double (&Jl)[9])
{
    const double th2 = r0 * r0 + r1 * r1 + r2 * r2;
    const double th = sqrt(th2);
    if (th < 1e-5) {
        R[0] = 1.0;  R[1] = -r2;  R[2] = r1;
        R[3] = r2;   R[4] = 1.0;  R[5] = -r0;
        R[6] = -r1;  R[7] = r0;   R[8] = 1.0;
        const double h0 = 0.5 * r0, h1 = 0.5 * r1, h2 = 0.5 * r2;
        Jl[0] = 1.0;  Jl[1] = -h2;  Jl[2] = h1;
        Jl[3] = h2;   Jl[4] = 1.0;  Jl[5] = -h0;
        Jl[6] = -h1;  Jl[7] = h0;   Jl[8] = 1.0;
        return;
    }
    const double inv = 1.0 / th;
    const double u0 = r0 * inv, u1 = r1 * inv, u2 = r2 * inv;
    // half-angle forms: sin(th) = 2 s c, 1 - cos(th) = 2 s^2 (no cancellation)
    double sh, ch;
    sincos(0.5 * th, &sh, &ch);
    const double s = 2.0 * sh * ch;
    const double v = 2.0 * sh * sh;        // 1 - cos(th)
    const double k1 = v * inv;             // (th/2) sinc^2(th/2)
    const double k2 = 1.0 - s * inv;       // 1 - sinc(th)
    // uhat^2 = u u^T - I for a unit axis
    const double u00 = u0 * u0 - 1.0, u11 = u1 * u1 - 1.0, u22 = u2 * u2 - 1.0;
    const double u01 = u0 * u1, u02 = u0 * u2, u12 = u1 * u2;
    R[0] = 1.0 + v * u00;       R[1] = v * u01 - s * u2;   R[2] = v * u02 + s * u1;
    R[3] = v * u01 + s * u2;    R[4] = 1.0 + v * u11;      R[5] = v * u12 - s * u0;
    R[6] = v * u02 - s * u1;    R[7] = v * u12 + s * u0;   R[8] = 1.0 + v * u22;
    Jl[0] = 1.0 + k2 * u00;     Jl[1] = k2 * u01 - k1 * u2;  Jl[2] = k2 * u02 + k1 * u1;
    Jl[3] = k2 * u01 + k1 * u2; Jl[4] = 1.0 + k2 * u11;      Jl[5] = k2 * u12 - k1 * u0;
    Jl[6] = k2 * u02 - k1 * u1; Jl[7] = k2 * u12 + k1 * u0;  Jl[8] = 1.0 + k2 * u22;
}

// Reciprocal / reciprocal square root for the per-corner path: hardware seed (MUFU.RCP64H /
// MUFU.RSQ64H, ~2^-22) refined by two Newton steps in fp64 -> <= 1-2 ulp.  No denormal /
// infinity slow paths: the operands here (eta, rho^2 of a point in front of the camera) are
// ordinary numbers, and results for rejected points are discarded.
__device__ __forceinline__ double fast_rcp(const double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ double fast_rsqrt(const double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);                 // 1 - x y^2
    y = fma(y * fma(0.375, e, 0.5), e, y);          // third-order step
    e = fma(-x * y, y, 1.0);
    return fma(0.5 * y, e, y);
}

__device__ __forceinline__ void mat3_mul(const double (&A)[9], const double (&B)[9], double (&C)[9])
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = fma(A[3 * i + 2], B[6 + j], fma(A[3 * i + 1], B[3 + j], A[3 * i] * B[j]));
}

// C = A * B^T
__device__ __forceinline__ void mat3_mul_bt(const double (&A)[9], const double (&B)[9], double (&C)[9])
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = fma(A[3 * i + 2], B[3 * j + 2], fma(A[3 * i + 1], B[3 * j + 1], A[3 * i] * B[3 * j]));
}

// ---------------------------------------------------------------------------
// Camera models.  eval() returns validity and fills
//   uv      projected point
//   Pu, Pv  rows of d(u,v)/dX            (projectionJacobian)
//   Ju, Jv  rows of d(u,v)/d intrinsics  (intrinsicJacobian)
// On failure the caller writes the 1e15 sentinel and zero Jacobians
// (calib_cost_functions.cpp:66-70, eucm.h:141-150,198-206).
// ---------------------------------------------------------------------------
template <int MODEL> struct Camera;

template <> struct Camera<MODEL_EUCM> {
    static constexpr int K = 6;
    struct Consts { double gamma, ab, hemi; bool hemi_test; };
    static constexpr int NCONST = 3;      // doubles of Consts a kernel may park in shared memory (save / load)
    __device__ __forceinline__ static Consts prepare(const double (&p)[K])
    {
        Consts c;
        c.gamma = 1.0 - p[0];
        c.ab = p[0] * p[1];
        c.hemi_test = p[0] > 0.5;                                  // eucm.h:49
        c.hemi = (p[0] - 1.0) / (p[0] + p[0] - 1.0);               // eucm.h:52
        return c;
    }
    __device__ __forceinline__ static void save(const Consts &c, double *d) { d[0] = c.gamma; d[1] = c.ab; d[2] = c.hemi; }
    __device__ __forceinline__ static Consts load(const double (&p)[K], const double *d)
    {
        Consts c;
        c.gamma = d[0]; c.ab = d[1]; c.hemi = d[2];
        c.hemi_test = p[0] > 0.5;
        return c;
    }
    __device__ __forceinline__ static bool eval(const double (&p)[K], const Consts &cc, const double x, const double y,
                                                const double z, double &u, double &v, double (&Pu)[3], double (&Pv)[3],
                                                double (&Ju)[K], double (&Jv)[K])
    {
        const double alpha = p[0], beta = p[1], fu = p[2], fv = p[3], u0 = p[4], v0 = p[5];
        const double gamma = cc.gamma;
        const double x2y2 = fma(x, x, y * y);
        const double rho2 = fma(beta, x2y2, z * z);
        const double ir = fast_rsqrt(rho2);
        const double rho = rho2 * ir;
        const double eta = fma(alpha, rho, gamma * z);
        bool ok = !(eta < 1e-3);                                   // eucm.h:46
        const double ie = fast_rcp(eta);
        if (cc.hemi_test && z * ie < cc.hemi) ok = false;          // eucm.h:49-54
        const double xn = x * ie, yn = y * ie;
        u = fma(fu, xn, u0);
        v = fma(fv, yn, v0);
        // dP/dX (eucm.h:152-163)
        const double k = ie * ie;
        const double ab = cc.ab * ir;
        const double fuk = fu * k, fvk = fv * k;
        const double jxy = ab * x * y;
        const double jz = fma(alpha * z, ir, gamma);
        Pu[0] = fuk * fma(-ab * x, x, eta);
        Pu[1] = -fuk * jxy;
        Pu[2] = -fuk * x * jz;
        Pv[0] = -fvk * jxy;
        Pv[1] = fvk * fma(-ab * y, y, eta);
        Pv[2] = -fvk * y * jz;
        // dP/dintr (eucm.h:208-222)
        const double da = (rho - z);
        const double db = 0.5 * alpha * x2y2 * ir;
        Ju[0] = -fuk * x * da;
        Ju[1] = -fuk * x * db;
        Ju[2] = xn;
        Ju[3] = 0.0;
        Ju[4] = 1.0;
        Ju[5] = 0.0;
        Jv[0] = -fvk * y * da;
        Jv[1] = -fvk * y * db;
        Jv[2] = 0.0;
        Jv[3] = yn;
        Jv[4] = 0.0;
        Jv[5] = 1.0;
        return ok;
    }
};

// normalised point m = (x,y)/(z + xi rho) and dm/dX, shared by UCM and MEI
__device__ __forceinline__ void unified_normalise(const double xi, const double x, const double y, const double z,
                                                  double &xn, double &yn, double &rho, double &d,
                                                  double (&mx)[3], double (&my)[3])
{
    const double rho2 = fma(x, x, fma(y, y, z * z));
    const double ir = fast_rsqrt(rho2);
    rho = rho2 * ir;
    const double den = fma(xi, rho, z);
    d = fast_rcp(den);
    const double d2 = d * d;
    xn = x * d;
    yn = y * d;
    const double xir = xi * ir;
    const double cxy = -xir * x * y * d2;
    const double cz = fma(xir, z, 1.0) * d2;
    mx[0] = fma(-xir * x, x, den) * d2;
    mx[1] = cxy;
    mx[2] = -x * cz;
    my[0] = cxy;
    my[1] = fma(-xir * y, y, den) * d2;
    my[2] = -y * cz;
}

template <> struct Camera<MODEL_UCM> {
    static constexpr int K = 5;
    struct Consts {};
    static constexpr int NCONST = 0;
    __device__ __forceinline__ static Consts prepare(const double (&)[K]) { return Consts(); }
    __device__ __forceinline__ static void save(const Consts &, double *) {}
    __device__ __forceinline__ static Consts load(const double (&)[K], const double *) { return Consts(); }
    __device__ __forceinline__ static bool eval(const double (&p)[K], const Consts &, const double x, const double y, const double z,
                                                double &u, double &v, double (&Pu)[3], double (&Pv)[3],
                                                double (&Ju)[K], double (&Jv)[K])
    {
        const double xi = p[0], fu = p[1], fv = p[2], u0 = p[3], v0 = p[4];
        double xn, yn, rho, d, mx[3], my[3];
        unified_normalise(xi, x, y, z, xn, yn, rho, d, mx, my);
        u = fma(fu, xn, u0);
        v = fma(fv, yn, v0);
#pragma unroll
        for (int i = 0; i < 3; i++) { Pu[i] = fu * mx[i]; Pv[i] = fv * my[i]; }
        const double dr = d * rho;
        Ju[0] = -fu * xn * dr;  Ju[1] = xn;  Ju[2] = 0.0;  Ju[3] = 1.0;  Ju[4] = 0.0;
        Jv[0] = -fv * yn * dr;  Jv[1] = 0.0; Jv[2] = yn;   Jv[3] = 0.0;  Jv[4] = 1.0;
        return true;                                               // ucm.h:37,58: no validity test
    }
};

template <> struct Camera<MODEL_MEI> {
    static constexpr int K = 10;
    struct Consts {};
    static constexpr int NCONST = 0;
    __device__ __forceinline__ static Consts prepare(const double (&)[K]) { return Consts(); }
    __device__ __forceinline__ static void save(const Consts &, double *) {}
    __device__ __forceinline__ static Consts load(const double (&)[K], const double *) { return Consts(); }
    __device__ __forceinline__ static bool eval(const double (&p)[K], const Consts &, const double x, const double y, const double z,
                                                double &u, double &v, double (&Pu)[3], double (&Pv)[3],
                                                double (&Ju)[K], double (&Jv)[K])
    {
        const double xi = p[0], k1 = p[1], k2 = p[2], k3 = p[3], k4 = p[4], k5 = p[5];
        const double fu = p[6], fv = p[7], u0 = p[8], v0 = p[9];
        double xn, yn, rho, d, mx[3], my[3];
        unified_normalise(xi, x, y, z, xn, yn, rho, d, mx, my);
        const double xx = xn * xn, yy = yn * yn, xy = xn * yn;
        const double r2 = xx + yy;
        const double r4 = r2 * r2, r6 = r4 * r2;
        const double D = fma(k3, r6, fma(k2, r4, fma(k1, r2, 1.0)));
        const double dD = fma(3.0 * k3, r4, fma(2.0 * k2, r2, k1));
        const double tx = fma(2.0, xx, r2), ty = fma(2.0, yy, r2);
        const double xd = fma(xn, D, fma(2.0 * k4, xy, k5 * tx));
        const double yd = fma(yn, D, fma(2.0 * k5, xy, k4 * ty));
        u = fma(fu, xd, u0);
        v = fma(fv, yd, v0);
        // distortion Jacobian d(xd,yd)/d(xn,yn), rows scaled by fu / fv (mei.h:178-186)
        const double off = fma(2.0 * xy, dD, 2.0 * fma(k4, xn, k5 * yn));
        const double a00 = fu * fma(2.0 * xx, dD, D + fma(2.0 * k4, yn, 6.0 * k5 * xn));
        const double a01 = fu * off;
        const double a10 = fv * off;
        const double a11 = fv * fma(2.0 * yy, dD, D + fma(2.0 * k5, xn, 6.0 * k4 * yn));
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Pu[i] = fma(a00, mx[i], a01 * my[i]);
            Pv[i] = fma(a10, mx[i], a11 * my[i]);
        }
        const double dr = d * rho;
        const double dxn = -xn * dr, dyn = -yn * dr;               // d(xn,yn)/dxi
        const double fux = fu * xn, fvy = fv * yn;
        Ju[0] = fma(a00, dxn, a01 * dyn);
        Ju[1] = fux * r2;  Ju[2] = fux * r4;  Ju[3] = fux * r6;
        Ju[4] = 2.0 * fu * xy;
        Ju[5] = fu * tx;
        Ju[6] = xd;  Ju[7] = 0.0;  Ju[8] = 1.0;  Ju[9] = 0.0;
        Jv[0] = fma(a10, dxn, a11 * dyn);
        Jv[1] = fvy * r2;  Jv[2] = fvy * r4;  Jv[3] = fvy * r6;
        Jv[4] = fv * ty;
        Jv[5] = 2.0 * fv * xy;
        Jv[6] = 0.0;  Jv[7] = yd;  Jv[8] = 0.0;  Jv[9] = 1.0;
        return true;                                               // mei.h:65: always true
    }
};

}  // namespace vg
