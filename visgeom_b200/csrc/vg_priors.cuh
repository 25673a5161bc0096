// vg_priors.cuh -- the 6-residual blocks that share the global problem with the reprojection blocks
// (SURVEY.md 8f-3): TransformationPrior (calib_cost_functions.h:83-108, calib_cost_functions.cpp:215-228)
// and OdometryPrior (calib_cost_functions.h:64-81, calib_cost_functions.cpp:119-213), plus the pose-block
// elimination for sequences whose consecutive elements an odometry block couples (the pose part of J^T J is
// then block tridiagonal instead of block diagonal).
//
// Everything numeric -- the constructors' weight matrices included -- runs on the GPU; the host only builds the
// index tables.  Records (doubles):
//   TransformationPrior constants  TP_CONST = [xi_prior 6 | A 36 row-major | R 9 | A^T A 21 packed lower]
//   TransformationPrior outputs    TP_OUT   = [g = A^T r 6 | cost]
//   OdometryPrior constants        OP_CONST = [zeta_prior 6 | A 36 row-major]
//   OdometryPrior outputs          OP_OUT   = [H11 21 | H22 21 | O = J2^T J1 36 row-major | g1 6 | g2 6 | cost]
#pragma once
#include "vg_solver_kernels.cuh"

namespace vg {

constexpr int TP_CONST = 72, TP_OUT = 7, OP_CONST = 42, OP_OUT = 91;
constexpr int TP_OFF_A = 6, TP_OFF_R = 42, TP_OFF_AtA = 51;
constexpr int OP_OFF_H11 = 0, OP_OFF_H22 = 21, OP_OFF_O = 42, OP_OFF_G1 = 78, OP_OFF_G2 = 84, OP_OFF_COST = 90;

// what a pose gathers besides its images' blocks
constexpr int EXTRA_TP = 0, EXTRA_OP_FIRST = 1, EXTRA_OP_SECOND = 2;

struct PriorTables {
    int n_tp, n_op;
    const double *tp_const, *op_const;
    const double *const *tp_xi;          // per prior: the element's 6 doubles in the parameter set being evaluated
    const double *const *op_xi;          // per edge: element i (element i+1 follows it in memory)
    double *tp_out, *op_out;             // outputs of that parameter set
    // shared-block contributions (priors on free global transforms), added by one rank only
    int n_tp_shared;
    const int *tp_shared_rec, *tp_shared_off;
};

// pose elimination with coupled / constant elements (vg_solver_kernels.cuh has the independent-pose kernels)
struct ChainTables {
    int n_seg;
    const int *seg_start, *seg_len;      // contiguous ranges of the global pose list
    const unsigned char *mask;           // per pose: 1 = handled by the chain kernels
    const unsigned char *fixed;          // per pose: constant element
    const int *extra_start, *extra_kind, *extra_rec;   // CSR over poses: prior records a pose gathers
    const int *prev_edge;                // per pose: odometry edge from its predecessor, -1 if none
    const double *tp_const, *tp_out, *op_out;
    double *off;                         // n_pose x 36: factor block L(i, i-1), row-major
    double *w;                           // n_pose x 6: right-hand sides of the back substitution
};

cudaError_t launch_prior_setup(int n_tp, const double *tp_in /* n_tp x [stiffness 6 | xi_prior 6] */, double *tp_const,
                               int n_op, const double *op_in /* n_op x [errV errW lambda | odom1 6 | odom2 6] */,
                               double *op_const, SolverLaunch sl);
// residuals, J^T J / J^T r pieces and costs of every prior at one parameter set; then cost and the shared-block
// pieces are added to red[A, g, cost] (add_shared = 0 on ranks other than the first)
cudaError_t launch_prior_eval(const PriorTables &t, int Ks, double *red, int add_shared, double *partial, SolverLaunch sl);

cudaError_t launch_chain_factor(const DatasetDesc *d_desc, int Ks, const int *pose_start, const int *contrib_ds,
                                const int *contrib_img, double *scale, LmConsts lm, double *ws, const ChainTables &c,
                                double *seg_gmax, int *fail_flag, SolverLaunch sl);
cudaError_t launch_chain_backsub(int Ks, const double *const *seq_cur, double *const *seq_cand, const int *pose_seq,
                                 const int *pose_local, const double *ws, const ChainTables &c, double *seg_partial,
                                 SolverLaunch sl);

}  // namespace vg
