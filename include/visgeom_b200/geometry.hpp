// geometry.hpp -- host-side SE(3) value type with visgeom's Transformation<double> interface
// (include/geometry/transformation.h:32-212, quaternion.h:25-131, geometry_core.h:32-76,120-124).
//
// Used by the calibration front end only for what the reference also does once per image on the host:
// reading transforms from JSON, un-winding a chain for an initial guess, composing the final chain for
// the report.  The per-corner residual / Jacobian path never comes through here (it is CUDA only).
#pragma once

#include <array>
#include <cmath>
#include <ostream>
#include <vector>

namespace visgeom_b200 {

using Array6d = std::array<double, 6>;

struct Vector3d {
    double v[3];
    Vector3d() : v{0, 0, 0} {}
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    Vector3d operator+(const Vector3d &o) const { return {v[0] + o[0], v[1] + o[1], v[2] + o[2]}; }
    Vector3d operator-(const Vector3d &o) const { return {v[0] - o[0], v[1] - o[1], v[2] - o[2]}; }
    Vector3d operator-() const { return {-v[0], -v[1], -v[2]}; }
    Vector3d operator*(double s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
    double dot(const Vector3d &o) const { return v[0] * o[0] + v[1] * o[1] + v[2] * o[2]; }
    Vector3d cross(const Vector3d &o) const
    {
        return {v[1] * o[2] - v[2] * o[1], v[2] * o[0] - v[0] * o[2], v[0] * o[1] - v[1] * o[0]};
    }
    double norm() const { return std::sqrt(dot(*this)); }
    void normalize() { const double n = norm(); v[0] /= n; v[1] /= n; v[2] /= n; }
};

struct Vector2d {
    double v[2];
    Vector2d() : v{0, 0} {}
    Vector2d(double x, double y) : v{x, y} {}
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};

using Vector2dVec = std::vector<Vector2d>;       // include/eigen.h:80-81
using Vector3dVec = std::vector<Vector3d>;

struct Matrix3d {
    double m[9];   // row-major
    double &operator()(int i, int j) { return m[3 * i + j]; }
    double operator()(int i, int j) const { return m[3 * i + j]; }
    Vector3d operator*(const Vector3d &x) const
    {
        return {m[0] * x[0] + m[1] * x[1] + m[2] * x[2], m[3] * x[0] + m[4] * x[1] + m[5] * x[2],
                m[6] * x[0] + m[7] * x[1] + m[8] * x[2]};
    }
    static Matrix3d fromColumns(const Vector3d &a, const Vector3d &b, const Vector3d &c)
    {
        return Matrix3d{{a[0], b[0], c[0], a[1], b[1], c[1], a[2], b[2], c[2]}};
    }
};

// angle wrapped to (-pi, pi]  (geometry_core.h:32-38)
inline double normalizeAngle(double a)
{
    const double two_pi = 2 * M_PI;
    a = std::fmod(a, two_pi);
    if (a > M_PI) a -= two_pi;
    else if (a <= -M_PI) a += two_pi;
    return a;
}

// unit quaternion (x, y, z, w); thresholds as quaternion.h:31-50,84-98
struct Quaternion {
    double x, y, z, w;
    Quaternion(double x_, double y_, double z_, double w_) : x(x_), y(y_), z(z_), w(w_) {}
    explicit Quaternion(const Vector3d &r)
    {
        const double th = r.norm();
        if (th < 1e-6) { x = r[0] / 2; y = r[1] / 2; z = r[2] / 2; w = 1; }
        else {
            const double s = std::sin(th / 2) / th;
            x = r[0] * s; y = r[1] * s; z = r[2] * s; w = std::cos(th / 2);
        }
    }
    // rotation matrix -> quaternion (quaternion.h:52-59; singular near a half turn, like the reference)
    explicit Quaternion(const Matrix3d &R)
    {
        w = std::sqrt(1 + R(0, 0) + R(1, 1) + R(2, 2)) / 2;
        const double k = 1 / (4 * w);
        x = (R(2, 1) - R(1, 2)) * k; y = (R(0, 2) - R(2, 0)) * k; z = (R(1, 0) - R(0, 1)) * k;
    }
    Quaternion inv() const { return {-x, -y, -z, w}; }
    Quaternion operator*(const Quaternion &q) const
    {
        return {w * q.x + x * q.w + y * q.z - z * q.y, w * q.y - x * q.z + y * q.w + z * q.x,
                w * q.z + x * q.y - y * q.x + z * q.w, w * q.w - x * q.x - y * q.y - z * q.z};
    }
    Vector3d rotate(const Vector3d &p) const
    {
        const Vector3d u(x, y, z);
        const Vector3d t = u.cross(p) * 2.0;
        return p + t * w + u.cross(t);
    }
    Vector3d toRotationVector() const
    {
        const double s = std::sqrt(x * x + y * y + z * z);
        if (s < 1e-5) return {2 * x, 2 * y, 2 * z};
        const double th = normalizeAngle(2 * std::atan2(s, w));
        return {x / s * th, y / s * th, z / s * th};
    }
};

inline Vector3d rotationVector(const Matrix3d &R) { return Quaternion(R).toRotationVector(); }

// R = exp(hat(r)); small-angle switch at 1e-5 as geometry_core.h:40-76
inline Matrix3d rotationMatrix(const Vector3d &r)
{
    const double th = r.norm();
    if (th < 1e-5) return Matrix3d{{1, -r[2], r[1], r[2], 1, -r[0], -r[1], r[0], 1}};
    const double ux = r[0] / th, uy = r[1] / th, uz = r[2] / th, s = std::sin(th), c1 = 1 - std::cos(th);
    return Matrix3d{{1 + c1 * (ux * ux - 1), c1 * ux * uy - s * uz, c1 * ux * uz + s * uy,
                     c1 * ux * uy + s * uz, 1 + c1 * (uy * uy - 1), c1 * uy * uz - s * ux,
                     c1 * ux * uz - s * uy, c1 * uy * uz + s * ux, 1 + c1 * (uz * uz - 1)}};
}

// translation + rotation vector; array / ABI order is [tx ty tz rx ry rz]
class Transformation {
public:
    Transformation() {}
    Transformation(const Vector3d &trans, const Vector3d &rot) : mtrans(trans), mrot(rot) {}
    Transformation(const Vector3d &trans, const Quaternion &q) : mtrans(trans), mrot(q.toRotationVector()) {}
    Transformation(const Vector3d &trans, const Matrix3d &R) : mtrans(trans), mrot(rotationVector(R)) {}
    explicit Transformation(const double *d) : mtrans(d[0], d[1], d[2]), mrot(d[3], d[4], d[5]) {}
    Transformation(double x, double y, double z, double rx, double ry, double rz) : mtrans(x, y, z), mrot(rx, ry, rz) {}
    Transformation(double x, double y, double z, double qx, double qy, double qz, double qw)
        : mtrans(x, y, z), mrot(Quaternion(qx, qy, qz, qw).toRotationVector()) {}

    // this o other
    Transformation compose(const Transformation &o) const
    {
        const Quaternion q1(mrot), q2(o.mrot);
        return Transformation(q1.rotate(o.mtrans) + mtrans, (q1 * q2).toRotationVector());
    }
    // this^-1 o other
    Transformation inverseCompose(const Transformation &o) const
    {
        const Quaternion q1i = Quaternion(mrot).inv(), q2(o.mrot);
        return Transformation(q1i.rotate(o.mtrans - mtrans), (q1i * q2).toRotationVector());
    }
    // this o other^-1
    Transformation composeInverse(const Transformation &o) const
    {
        const Quaternion q = Quaternion(mrot) * Quaternion(o.mrot).inv();
        return Transformation(mtrans - q.rotate(o.mtrans), q.toRotationVector());
    }
    Transformation inverse() const { return Transformation(-(rotMatInv() * mtrans), -mrot); }

    const Vector3d &trans() const { return mtrans; }
    const Vector3d &rot() const { return mrot; }
    Matrix3d rotMat() const { return rotationMatrix(mrot); }
    Matrix3d rotMatInv() const { return rotationMatrix(-mrot); }
    void transform(const Vector3d &src, Vector3d &dst) const { dst = rotMat() * src + mtrans; }
    Array6d toArray() const { return {mtrans[0], mtrans[1], mtrans[2], mrot[0], mrot[1], mrot[2]}; }
    void toArray(double *d) const { const Array6d a = toArray(); for (int i = 0; i < 6; i++) d[i] = a[i]; }

    friend std::ostream &operator<<(std::ostream &os, const Transformation &t)
    {
        return os << t.mtrans[0] << " " << t.mtrans[1] << " " << t.mtrans[2] << " " << t.mrot[0] << " " << t.mrot[1]
                  << " " << t.mrot[2];
    }

private:
    Vector3d mtrans, mrot;
};

using Transf = Transformation;

}  // namespace visgeom_b200
