// camera.hpp -- host-side camera objects with visgeom's ICamera interface
// (include/projection/generic_camera.h:29-123, eucm.h, ucm.h, mei.h): parameter storage, bounds, clone,
// reconstructPoint (closed form on the host: the calibration front end back-projects four corners per image,
// unified_calibration.cpp:1066-1084) and the projection side -- projectPoint, projectionJacobian,
// intrinsicJacobian, projectPointCloud, reconstructPointCloud.
//
// Projection and its Jacobians are NOT computed on the host: these methods call the CUDA kernels through the C ABI
// (vg_project_points / vg_reconstruct_points, visgeom_b200.h) and return false -- as the reference does for a point
// it cannot project -- when the call fails (vg_last_error() says why; without a GPU it always does: there is no CPU
// path).  The *PointCloud forms are the fast path (one launch for all points); the single-point virtuals are the same
// call with n = 1, kept for source compatibility with visgeom callers.
#pragma once

#include <cmath>
#include <vector>

#include "../visgeom_b200.h"
#include "geometry.hpp"

namespace visgeom_b200 {

class ICamera {
public:
    int width = 0, height = 0;

    ICamera(int model, const double *p) : mmodel(model), params(p, p + vg_model_num_params(model)) {}
    virtual ~ICamera() {}
    virtual bool reconstructPoint(const Vector2d &src, Vector3d &dst) const = 0;
    virtual ICamera *clone() const = 0;

    // generic_camera.h:39
    virtual bool projectPoint(const Vector3d &src, Vector2d &dst) const
    {
        unsigned char ok = 0;
        double uv[2] = {dst[0], dst[1]};
        if (vg_project_points(mmodel, params.data(), 1, src.v, uv, nullptr, nullptr, &ok) != VG_OK) return false;
        if (ok) dst = Vector2d(uv[0], uv[1]);
        return ok != 0;
    }
    // generic_camera.h:46: dudx, dvdx are 3 doubles each ("MEMORY IS SUPPOSED TO BE ALLOCATED")
    virtual bool projectionJacobian(const Vector3d &src, double *dudx, double *dvdx) const
    {
        unsigned char ok = 0;
        double J[6];
        if (vg_project_points(mmodel, params.data(), 1, src.v, nullptr, J, nullptr, &ok) != VG_OK) return false;
        for (int i = 0; i < 3; i++) { dudx[i] = J[i]; dvdx[i] = J[3 + i]; }
        return ok != 0;
    }
    // generic_camera.h:50: numParams() doubles each
    virtual bool intrinsicJacobian(const Vector3d &src, double *dudalpha, double *dvdalpha) const
    {
        unsigned char ok = 0;
        const int K = numParams();
        std::vector<double> J(2 * (size_t)K);
        if (vg_project_points(mmodel, params.data(), 1, src.v, nullptr, nullptr, J.data(), &ok) != VG_OK) return false;
        for (int i = 0; i < K; i++) { dudalpha[i] = J[i]; dvdalpha[i] = J[K + i]; }
        return ok != 0;
    }
    // generic_camera.h:90-113
    bool projectPointCloud(const Vector3dVec &src, Vector2dVec &dst) const
    {
        std::vector<bool> mask;
        return projectPointCloud(src, dst, mask);
    }
    bool projectPointCloud(const Vector3dVec &src, Vector2dVec &dst, std::vector<bool> &maskVec) const
    {
        static_assert(sizeof(Vector3d) == 3 * sizeof(double) && sizeof(Vector2d) == 2 * sizeof(double), "packed point types");
        const size_t n = src.size();
        dst.resize(n);
        maskVec.assign(n, false);
        if (n == 0) return true;
        std::vector<unsigned char> ok(n, 0);
        if (vg_project_points(mmodel, params.data(), (long long)n, src[0].v, dst[0].v, nullptr, nullptr, ok.data()) != VG_OK)
            return false;
        bool res = true;
        for (size_t i = 0; i < n; i++) { maskVec[i] = ok[i] != 0; res = res && maskVec[i]; }
        return res;
    }
    // generic_camera.h:64-88, batched on the GPU as well (single points stay on the host: reconstructPoint)
    bool reconstructPointCloud(const Vector2dVec &src, Vector3dVec &dst) const
    {
        std::vector<bool> mask;
        return reconstructPointCloud(src, dst, mask);
    }
    bool reconstructPointCloud(const Vector2dVec &src, Vector3dVec &dst, std::vector<bool> &maskVec) const
    {
        const size_t n = src.size();
        dst.resize(n);
        maskVec.assign(n, false);
        if (n == 0) return true;
        std::vector<unsigned char> ok(n, 0);
        if (vg_reconstruct_points(mmodel, params.data(), (long long)n, src[0].v, dst[0].v, ok.data()) != VG_OK) return false;
        bool res = true;
        for (size_t i = 0; i < n; i++) { maskVec[i] = ok[i] != 0; res = res && maskVec[i]; }
        return res;
    }
    virtual double getCenterU() { return width / 2; }       // generic_camera.h:41-43
    virtual double getCenterV() { return height / 2; }

    void setParameters(const double *p) { params.assign(p, p + params.size()); }
    const double *getParams() const { return params.data(); }
    int numParams() const { return (int)params.size(); }
    double lowerBound(int idx) const { double lo, hi; vg_model_bounds(mmodel, idx, &lo, &hi); return lo; }
    double upperBound(int idx) const { double lo, hi; vg_model_bounds(mmodel, idx, &lo, &hi); return hi; }
    int model() const { return mmodel; }      // VG_MODEL_* of the CUDA kernels

protected:
    int mmodel;
    std::vector<double> params;
};

// [alpha, beta, fu, fv, u0, v0]; back-projection eucm.h:85-106
class EnhancedCamera : public ICamera {
public:
    explicit EnhancedCamera(const double *p) : ICamera(VG_MODEL_EUCM, p) {}
    bool reconstructPoint(const Vector2d &src, Vector3d &dst) const override
    {
        const double alpha = params[0], beta = params[1], gamma = 1 - alpha;
        const double xn = (src[0] - params[4]) / params[2], yn = (src[1] - params[5]) / params[3];
        const double u2 = xn * xn + yn * yn;
        const double det = 1 - (alpha - gamma) * beta * u2;
        if (det < 0) return false;
        dst = Vector3d(xn, yn, (1 - u2 * alpha * alpha * beta) / (gamma + alpha * std::sqrt(det)));
        return true;
    }
    ICamera *clone() const override { return new EnhancedCamera(*this); }
};

// shared by UCM [xi, fu, fv, u0, v0] (ucm.h:81-103) and MEI [xi, k1..k5, fu, fv, u0, v0] (mei.h:90-112: the
// reference ignores the distortion terms there)
inline Vector3d unifiedBackProject(double xi, double xn, double yn)
{
    const double u2 = xn * xn + yn * yn;
    const double g = std::sqrt(1 + u2 * (1 - xi * xi));
    const double en = -g - xi * u2, ed = xi * xi * u2 - 1;
    return Vector3d(xn, yn, ed / (ed + xi * en));
}

class UnifiedCamera : public ICamera {
public:
    explicit UnifiedCamera(const double *p) : ICamera(VG_MODEL_UCM, p) {}
    bool reconstructPoint(const Vector2d &src, Vector3d &dst) const override
    {
        dst = unifiedBackProject(params[0], (src[0] - params[3]) / params[1], (src[1] - params[4]) / params[2]);
        return true;
    }
    ICamera *clone() const override { return new UnifiedCamera(*this); }
};

class MeiCamera : public ICamera {
public:
    explicit MeiCamera(const double *p) : ICamera(VG_MODEL_MEI, p) {}
    bool reconstructPoint(const Vector2d &src, Vector3d &dst) const override
    {
        dst = unifiedBackProject(params[0], (src[0] - params[8]) / params[6], (src[1] - params[9]) / params[7]);
        return true;
    }
    ICamera *clone() const override { return new MeiCamera(*this); }
};

}  // namespace visgeom_b200
