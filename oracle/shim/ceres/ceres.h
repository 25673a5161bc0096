// Stand-in for the few Ceres declarations visgeom's include/ceres.h and
// include/calibration/calib_cost_functions.h name.  TEST INFRASTRUCTURE ONLY (see oracle/shim/Eigen/Eigen).
// The CostFunction / FirstOrderFunction interfaces are functional; the LM solver itself is restated in
// oracle/oracle_lm.c, the line-search GradientProblemSolver in gradient_solver.h.
#ifndef VISGEOM_ORACLE_CERES_SHIM
#define VISGEOM_ORACLE_CERES_SHIM
#include <vector>
namespace ceres {
class CostFunction {
public:
    CostFunction() : num_residuals_(0) {}
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
    const std::vector<int> &parameter_block_sizes() const { return parameter_block_sizes_; }
    int num_residuals() const { return num_residuals_; }
protected:
    std::vector<int> *mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
    void set_num_residuals(int n) { num_residuals_ = n; }
private:
    std::vector<int> parameter_block_sizes_;
    int num_residuals_;
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {
public:
    SizedCostFunction()
    {
        set_num_residuals(kNumResiduals);
        const int sizes[] = {Ns...};
        for (int s : sizes) mutable_parameter_block_sizes()->push_back(s);
    }
};
template <typename F, int S = 4> class DynamicAutoDiffCostFunction;
class FirstOrderFunction {
public:
    virtual ~FirstOrderFunction() {}
    virtual bool Evaluate(const double *parameters, double *cost, double *gradient) const = 0;
    virtual int NumParameters() const = 0;
};
class LossFunction { public: virtual ~LossFunction() {} };
class SoftLOneLoss;
class CauchyLoss;
class Problem;
class Solver;
void Solve();
}  // namespace ceres
#include "gradient_solver.h"   // BiCubicInterpolator, GradientProblem, GradientProblemSolver (restated, see there)
#endif
