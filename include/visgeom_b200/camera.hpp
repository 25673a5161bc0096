// camera.hpp -- host-side camera objects with visgeom's ICamera interface
// (include/projection/generic_camera.h:29-123, eucm.h, ucm.h, mei.h), as far as the calibration front end needs
// it: parameter storage, bounds, clone and reconstructPoint (back-projection of the four outer corners for the
// initial pose, unified_calibration.cpp:1066-1084).
//
// projectPoint / projectionJacobian / intrinsicJacobian are deliberately NOT implemented on the host: in this
// engine they exist only inside the CUDA kernels (visgeom_b200.h: vg_eval_chain*, vg_problem_*); a camera
// object identifies its model to them through model().
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
