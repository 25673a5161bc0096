// Stand-in for ceres::BiCubicInterpolator, ceres::GradientProblem and ceres::GradientProblemSolver (TEST INFRASTRUCTURE
// ONLY, see ceres.h next to this file).  Ceres is a THIRD-PARTY dependency that is not under /root/reference
// (README.md: "ceres-solver", unpinned, not vendored); its published algorithms are restated here so that the
// reference's SubpixelCorner / CornerDetector::improveCorners (src/calibration/corner_detector.cpp:47-100, 162-198)
// compile and run where they lie:
//   * BiCubicInterpolator<Grid>::Evaluate: Catmull-Rom cubic convolution (Keys 1981) over the 4 x 4 neighbourhood, rows
//     first, then down the column, with both first derivatives (ceres/cubic_interpolation.h, CubicHermiteSpline);
//   * GradientProblemSolver with its default options: L-BFGS directions (rank 20, H0 = I, pairs with s.y <= 1e-14
//     skipped), Wolfe line search = bracketing phase + zoom phase with the minimiser of the cubic through two
//     (value, slope) samples (c1 1e-4, c2 0.9, expansion 10, 20 trial steps, min step 1e-9), first trial step
//     min(1, 1 / |g|_inf) then min(1, 2 (f_k - f_{k-1}) / g_k.d_k), termination on |g|_inf <= 1e-10,
//     |df| <= 1e-6 |f|, |dx| <= 1e-8 (|x| + 1e-8), 50 iterations (ceres/line_search_minimizer.cc, line_search.cc,
//     polynomial.cc, low_rank_inverse_hessian.cc).
// PARITY UNPINNED against Ceres itself for the minimiser's path (Ceres cannot be built here; it solves the interpolation
// through Eigen's FullPivLU where this file uses the closed form): results are held to a tolerance, see DESIGN.md.
#ifndef VISGEOM_ORACLE_CERES_GRADIENT_SOLVER
#define VISGEOM_ORACLE_CERES_GRADIENT_SOLVER
#include <algorithm>
#include <cmath>
#include <memory>
#include <string>
#include <vector>
namespace ceres {

inline void CubicHermiteSpline1(double p0, double p1, double p2, double p3, double x, double *f, double *dfdx)
{
    const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
    const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
    const double c = 0.5 * (-p0 + p2);
    const double d = p1;
    if (f) *f = d + x * (c + x * (b + x * a));
    if (dfdx) *dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

template <typename Grid> class BiCubicInterpolator {
public:
    explicit BiCubicInterpolator(const Grid &grid) : grid_(grid) {}
    void Evaluate(double r, double c, double *f, double *dfdr, double *dfdc) const
    {
        const int row = (int)std::floor(r), col = (int)std::floor(c);
        double fr[4], dfr[4];
        for (int k = 0; k < 4; k++) {
            double p0, p1, p2, p3;
            grid_.GetValue(row - 1 + k, col - 1, &p0);
            grid_.GetValue(row - 1 + k, col, &p1);
            grid_.GetValue(row - 1 + k, col + 1, &p2);
            grid_.GetValue(row - 1 + k, col + 2, &p3);
            CubicHermiteSpline1(p0, p1, p2, p3, c - col, &fr[k], &dfr[k]);
        }
        CubicHermiteSpline1(fr[0], fr[1], fr[2], fr[3], r - row, f, dfdr);
        if (dfdc) CubicHermiteSpline1(dfr[0], dfr[1], dfr[2], dfr[3], r - row, dfdc, nullptr);
    }
private:
    const Grid &grid_;
};

enum LoggingType { SILENT, PER_MINIMIZER_ITERATION };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE };

class GradientProblem {
public:
    explicit GradientProblem(FirstOrderFunction *f) : function_(f) {}
    int NumParameters() const { return function_->NumParameters(); }
    bool Evaluate(const double *x, double *cost, double *gradient) const { return function_->Evaluate(x, cost, gradient); }
private:
    std::unique_ptr<FirstOrderFunction> function_;       // Ceres takes ownership too
};

class GradientProblemSolver {
public:
    struct Options {
        int max_num_iterations = 50;
        int max_lbfgs_rank = 20;
        double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
        double min_line_search_step_size = 1e-9;
        double line_search_sufficient_function_decrease = 1e-4, line_search_sufficient_curvature_decrease = 0.9;
        double max_line_search_step_contraction = 1e-3, min_line_search_step_contraction = 0.6;
        double max_line_search_step_expansion = 10.0;
        int max_num_line_search_step_size_iterations = 20;
        int max_num_line_search_direction_restarts = 5;
        LoggingType logging_type = PER_MINIMIZER_ITERATION;
        bool minimizer_progress_to_stdout = false;
    };
    struct Summary {
        TerminationType termination_type = FAILURE;
        double initial_cost = 0, final_cost = 0;
        int num_iterations = 0, num_cost_evaluations = 0;
        std::string FullReport() const { return std::string(); }
    };
};

namespace gs_detail {
struct Sample { double x, f, g; };       // step, value, slope along the direction

// minimiser over [lo, hi] of the cubic through two (value, slope) samples: the interval's midpoint, its two ends and
// the real parts of the derivative's roots are the candidates (polynomial.cc: MinimizePolynomial)
inline double cubic_min(const Sample &a, const Sample &b, double lo, double hi)
{
    // p(x) = c3 x^3 + c2 x^2 + c1 x + c0 with p(a.x) = a.f, p'(a.x) = a.g, p(b.x) = b.f, p'(b.x) = b.g
    const double h = b.x - a.x;
    const double d = (b.f - a.f) / h;
    const double k3 = (a.g + b.g - 2.0 * d) / (h * h);          // in t = x - a.x: a.f + a.g t + k2 t^2 + k3 t^3
    const double k2 = (3.0 * d - 2.0 * a.g - b.g) / h;
    auto P = [&](double x) { const double t = x - a.x; return a.f + t * (a.g + t * (k2 + t * k3)); };
    double best_x = 0.5 * (lo + hi), best = P(best_x);
    auto take = [&](double x) { const double v = P(x); if (v < best) { best = v; best_x = x; } };
    take(lo);
    take(hi);
    // p'(t) = 3 k3 t^2 + 2 k2 t + a.g
    const double qa = 3.0 * k3, qb = 2.0 * k2, qc = a.g;
    double roots[2];
    int nr = 0;
    if (qa != 0.0) {
        const double D = qb * qb - 4.0 * qa * qc, sD = std::sqrt(std::fabs(D));
        if (D >= 0.0) {
            if (qb >= 0.0) { roots[0] = (-qb - sD) / (2.0 * qa); roots[1] = (2.0 * qc) / (-qb - sD); }
            else { roots[0] = (2.0 * qc) / (-qb + sD); roots[1] = (-qb + sD) / (2.0 * qa); }
        } else roots[0] = roots[1] = -qb / (2.0 * qa);
        nr = 2;
    } else if (qb != 0.0) { roots[0] = -qc / qb; nr = 1; }
    for (int i = 0; i < nr; i++) {
        const double x = roots[i] + a.x;
        if (!(x >= lo && x <= hi)) continue;
        take(x);
    }
    return best_x;
}
}  // namespace gs_detail

inline void Solve(const GradientProblemSolver::Options &opt, const GradientProblem &problem, double *x,
                  GradientProblemSolver::Summary *summary)
{
    using gs_detail::Sample;
    const int n = problem.NumParameters();
    std::vector<double> g(n), xn(n), gn(n), dir(n), trial(n), gt(n);
    std::vector<std::vector<double> > S, Y;
    std::vector<double> SY;
    double f = 0, f_prev = 0;
    int evals = 0;
    problem.Evaluate(x, &f, g.data()); evals++;
    summary->initial_cost = f;
    auto max_norm = [&](const std::vector<double> &v) { double m = 0; for (double e : v) m = std::max(m, std::fabs(e)); return m; };
    auto dot = [&](const std::vector<double> &a, const std::vector<double> &b) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; };
    summary->termination_type = NO_CONVERGENCE;
    int it = 0, restarts = 0;
    if (max_norm(g) <= opt.gradient_tolerance) summary->termination_type = CONVERGENCE;
    else
    for (;;) {
        if (it >= opt.max_num_iterations) break;
        it++;
        // L-BFGS two-loop recursion, H0 = I
        bool steepest = S.empty();
        for (int i = 0; i < n; i++) dir[i] = g[i];
        if (!steepest) {
            const int m = (int)S.size();
            std::vector<double> alpha(m);
            for (int i = m - 1; i >= 0; i--) {
                alpha[i] = dot(S[i], dir) / SY[i];
                for (int k = 0; k < n; k++) dir[k] -= alpha[i] * Y[i][k];
            }
            for (int i = 0; i < m; i++) {
                const double beta = dot(Y[i], dir) / SY[i];
                for (int k = 0; k < n; k++) dir[k] += S[i][k] * (alpha[i] - beta);
            }
        }
        for (int i = 0; i < n; i++) dir[i] = -dir[i];
        double slope = dot(g, dir);
        if (!steepest && slope >= 0.0) {                 // not a descent direction: restart from steepest descent
            if (++restarts > opt.max_num_line_search_direction_restarts) { summary->termination_type = FAILURE; break; }
            S.clear(); Y.clear(); SY.clear();
            for (int i = 0; i < n; i++) dir[i] = -g[i];
            slope = dot(g, dir);
            steepest = true;
        }
        const double step0 = (it == 1 || steepest) ? std::min(1.0, 1.0 / max_norm(g)) : std::min(1.0, 2.0 * (f - f_prev) / slope);
        if (!(step0 > 0.0)) { summary->termination_type = FAILURE; break; }
        // ---- Wolfe line search along dir ----
        const double dmax = max_norm(dir);
        auto phi = [&](double a, std::vector<double> &grad_out) {
            for (int i = 0; i < n; i++) trial[i] = x[i] + a * dir[i];
            Sample s; s.x = a;
            problem.Evaluate(trial.data(), &s.f, grad_out.data()); evals++;
            s.g = dot(grad_out, dir);
            return s;
        };
        const Sample init = {0.0, f, slope};
        Sample prev = init, cur = phi(step0, gt), lo = init, hi = init, sol = init;
        bool zoom = false, ok = true;
        int ls_it = 0;
        for (;;) {                                       // bracketing phase
            ls_it++;
            if (cur.f > init.f + opt.line_search_sufficient_function_decrease * init.g * cur.x || (prev.x > 0.0 && cur.f > prev.f)) {
                zoom = true; lo = prev; hi = cur; break;
            }
            if (std::fabs(cur.g) <= -opt.line_search_sufficient_curvature_decrease * init.g) { lo = hi = cur; break; }
            if (cur.g >= 0.0) { zoom = true; lo = cur; hi = prev; break; }
            if (std::fabs(cur.x - prev.x) * dmax < opt.min_line_search_step_size) { ok = false; break; }
            if (ls_it >= opt.max_num_line_search_step_size_iterations) { lo = cur.f < lo.f ? cur : lo; break; }
            const double a = gs_detail::cubic_min(prev, cur, cur.x, cur.x * opt.max_line_search_step_expansion);
            if (a * dmax < opt.min_line_search_step_size) { ok = false; break; }
            prev = cur;
            cur = phi(a, gt);
        }
        if (!ok) { summary->termination_type = FAILURE; break; }
        Sample best = lo;
        if (zoom) {
            if (lo.f > hi.f) std::swap(lo, hi);
            bool have_sol = false;
            for (;;) {                                   // zoom phase
                if (ls_it >= opt.max_num_line_search_step_size_iterations) break;
                if (std::fabs(hi.x - lo.x) * dmax < opt.min_line_search_step_size) break;
                ls_it++;
                const Sample &lb = lo.x < hi.x ? lo : hi, &ub = lo.x < hi.x ? hi : lo;
                const double a = gs_detail::cubic_min(lb, ub, lb.x, ub.x);
                sol = phi(a, gt);
                have_sol = true;
                if (sol.f > init.f + opt.line_search_sufficient_function_decrease * init.g * sol.x || sol.f >= lo.f) { hi = sol; continue; }
                if (std::fabs(sol.g) <= -opt.line_search_sufficient_curvature_decrease * init.g) break;
                if (sol.g * (hi.x - lo.x) >= 0.0) hi = lo;
                lo = sol;
            }
            best = (!have_sol || sol.f > lo.f) ? lo : sol;
        }
        if (!(best.x > 0.0)) { summary->termination_type = FAILURE; break; }
        // ---- take the step ----
        for (int i = 0; i < n; i++) xn[i] = x[i] + best.x * dir[i];
        double fn;
        problem.Evaluate(xn.data(), &fn, gn.data()); evals++;
        std::vector<double> s(n), y(n);
        double step_norm = 0, x_norm = 0;
        for (int i = 0; i < n; i++) { s[i] = best.x * dir[i]; y[i] = gn[i] - g[i]; step_norm += s[i] * s[i]; x_norm += xn[i] * xn[i]; }
        step_norm = std::sqrt(step_norm); x_norm = std::sqrt(x_norm);
        const double sy = dot(s, y);
        if (sy > 1e-14) {
            if ((int)S.size() == opt.max_lbfgs_rank) { S.erase(S.begin()); Y.erase(Y.begin()); SY.erase(SY.begin()); }
            S.push_back(s); Y.push_back(y); SY.push_back(sy);
        }
        f_prev = f; f = fn;
        for (int i = 0; i < n; i++) { x[i] = xn[i]; g[i] = gn[i]; }
        if (max_norm(g) <= opt.gradient_tolerance) { summary->termination_type = CONVERGENCE; break; }
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { summary->termination_type = CONVERGENCE; break; }
        if (std::fabs(f_prev - f) <= opt.function_tolerance * std::fabs(f_prev)) { summary->termination_type = CONVERGENCE; break; }
    }
    summary->final_cost = f;
    summary->num_iterations = it;
    summary->num_cost_evaluations = evals;
}
}  // namespace ceres
#endif
