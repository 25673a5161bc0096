// vg_solver_kernels.cuh -- device side of the Levenberg-Marquardt step that replaces
// the normal-equation build / elimination Ceres performs inside ceres::Solve
// (unified_calibration.cpp:53; structure described in SURVEY.md 3.2).
//
// The Hessian of the calibration problem is an arrowhead: one shared block (all
// free intrinsics + free global transforms, Ks columns) plus one independent 6x6
// block per free sequence pose.  Per evaluation the fused kernel (vg_eval.cu) leaves
// one packed [J r]^T [J r] block per image; the kernels here
//   (the fold of the evaluation kernel's per-CTA block sums into A (Ks x Ks), g_a and the cost is fused into
//    that kernel's tail, vg_eval_impl.cuh: fused_reduce)
//   pose_factor       : per pose gather C (6x6), E (Ks x 6), b; damp, Cholesky,
//                       Z = L^-1 E^T, z = L^-1 b
//   gram_reduce       : S_red = sum Z^T Z, v_red = sum Z^T z  (the Schur complement terms)
//   pose_backsub      : delta_p = -L^-T (z + Z delta_a), candidate poses, model-decrease
//                       and step-norm partial sums
#pragma once
#include "vg_eval.cuh"

#include <cuda_runtime.h>

#include <vector>

namespace vg {

constexpr int MAX_W = 41;          // 10 intrinsics + 5*6 + residual column
constexpr int MAX_SHARED_LOCAL = 34;

// column kinds of a dataset's local layout
constexpr int COL_CONST = 0, COL_SHARED = 1, COL_POSE = 2, COL_RESID = 3;

struct DatasetDesc {
    const double *H;               // n_img x ne (the set being reduced)
    const int *seq_index;          // nullable
    int n_img, ne, W;
    int pose_col;                  // first local column of the sequence element, -1 if that element is constant
    int pose_base;                 // index of pose 0 of its sequence in the global pose list
    int n_sl;                      // number of local columns that map to shared parameters
    int sl_col[MAX_SHARED_LOCAL];  // their local column
    int sl_idx[MAX_SHARED_LOCAL];  // their global shared index
    int kind[MAX_W];
    int idx[MAX_W];
};

struct LmConsts {
    double radius, min_diag, max_diag;
    int init_scale;                // 1: (re)compute the Jacobi scaling from this evaluation
    int jacobi_scaling;
};

// ws layout (doubles) per pose p, SoA over poses would not coalesce better for one thread
// per pose, so the scratch is pose-major: [L (21) | lambda (6) | z (6) | Z (6 x Ks)]
__host__ __device__ inline int pose_ws_stride(int Ks) { return 21 + 6 + 6 + 6 * Ks; }

// Reduction buffer layout (doubles); every entry is a sum over the local images /
// poses, laid out as the two segments that are exchanged across GPUs:
//   segment E (fresh after every evaluation + back-substitution):
//     A (Ks*Ks, row-major symmetric) | g_a (Ks) | cost (1/2 sum r^2) | model partial,
//     step^2, x^2 over poses (3)
//   segment S (fresh after every pose_schur):
//     S_red (Ks*Ks) | v_red (Ks) | failed factorisations (count) | max |g_pose|, one
//     slot per rank (only the own slot is non-zero, so a SUM all-reduce gathers them)
__host__ __device__ inline int red_off_A(int) { return 0; }
__host__ __device__ inline int red_off_g(int Ks) { return Ks * Ks; }
__host__ __device__ inline int red_off_cost(int Ks) { return Ks * Ks + Ks; }
__host__ __device__ inline int red_off_model(int Ks) { return Ks * Ks + Ks + 1; }
__host__ __device__ inline int red_segE_size(int Ks) { return Ks * Ks + Ks + 4; }
__host__ __device__ inline int red_off_S(int Ks) { return red_segE_size(Ks); }
__host__ __device__ inline int red_off_v(int Ks) { return red_off_S(Ks) + Ks * Ks; }
__host__ __device__ inline int red_off_fail(int Ks) { return red_off_v(Ks) + Ks; }
__host__ __device__ inline int red_off_gmax(int Ks) { return red_off_fail(Ks) + 1; }
__host__ __device__ inline int red_segS_size(int Ks, int nranks) { return Ks * Ks + Ks + 1 + nranks; }
__host__ __device__ inline int red_size(int Ks, int nranks) { return red_segE_size(Ks) + red_segS_size(Ks, nranks); }

// after the two segments, per parameter set: what reduced_solve leaves for the host about the step TOWARDS this set
//   [max |projected gradient| at the current point, failed pose factorisations, reduced Cholesky ok, g_a . delta_a,
//    delta_a^T A delta_a, |step|^2 and |x|^2 over the shared parameters, spare] then a copy of the candidate slab
constexpr int SOLVE_OUT = 8;

// what the reduced solve reads and leaves (reduced_solve_kernel, or pose_factor's fused tail on one rank)
struct SolveArgs {
    int slab_n, nranks;
    const double *red_cur;          // the current set's reduction buffer: A, g_a (segment E), S_red, v_red, max |g| (segment S)
    double *red_cand;               // the candidate set's buffer: scalars + candidate slab go after its two segments
    const double *slab_cur;
    double *slab_cand, *delta_a;
    const int *sh_off;              // slab position of shared parameter j
    const double *sh_lo, *sh_hi;    // its box bounds
    double *scale_a;                // Jacobi scaling of the shared block (written when lm.init_scale)
};

struct SolverLaunch {
    cudaStream_t stream;
    unsigned long long *launches;
};

// Tables of the shared-block reduction fused into the evaluation kernel (vg_eval.cuh: FinOut / FinSrc).
// rows[ds] = rows of dataset ds' sums region (1: the kernel leaves one reduced row per dataset)
void build_finalize_tables(const DatasetDesc *h_desc, const int *rows, int n_ds, int Ks, std::vector<int> &offsets,
                           std::vector<FinOut> &outs, std::vector<FinSrc> &srcs, size_t *sum_doubles);

// per-pose factorisation + Schur terms -> ws, red[S,v], red[gmax]
cudaError_t launch_pose_schur(const DatasetDesc *d_desc, int n_pose, int Ks,
                              const int *pose_start, const int *contrib_ds, const int *contrib_img,
                              double *scale, LmConsts lm, double *ws, double *partial, size_t partial_doubles,
                              double *red, int *fail_flag, int rank, int nranks, SolverLaunch sl,
                              const unsigned char *chain_mask = nullptr, int n_seg = 0,
                              const SolveArgs *fused = nullptr, unsigned int *ticket = nullptr);

size_t pose_scratch(int n_pose, int Ks, int n_seg = 0);                  // doubles
int pose_factor_blocks(int n_pose);                                      // blocks of pose_factor (slots of max |g|)
int pose_backsub_blocks(int n_pose);                                     // blocks of pose_backsub (rows of 3 sums)

// candidate poses and model / norm partial sums -> red[model..]
// pose_ptr_cur/cand: per pose-list entry the device address of its 6 doubles is
// base_cur[seq] + 6*i, expressed through pose_seq (sequence id) and pose_local (index).
cudaError_t launch_pose_backsub(int n_pose, int Ks, const double *delta_a,
                                const double *const *seq_cur, double *const *seq_cand,
                                const int *pose_seq, const int *pose_local,
                                const double *ws, double *partial, size_t partial_doubles, double *red,
                                SolverLaunch sl, const unsigned char *chain_mask = nullptr, double *chain_w = nullptr,
                                unsigned int *ticket = nullptr);   // ticket: the last block also sums the rows into red
// reduced system of the current set -> delta_a, candidate slab, host scalars (red_cand tail)
cudaError_t launch_reduced_solve(int Ks, const SolveArgs &sa, LmConsts lm, SolverLaunch sl);
// buf[0..count) <- its sum over the ranks, through the peers' inboxes (one block; vg_peer.cuh)
cudaError_t launch_peer_exchange(double *buf, int count, const PeerCtx &pc, SolverLaunch sl);
// the collection half of an exchange an evaluation kernel has posted; *done <- its number
// post: buf still holds this rank's own block of that exchange (nobody posted it yet): post it, then collect
cudaError_t launch_peer_collect(double *buf, int count, const PeerCtx &pc, unsigned long long *done, bool post, SolverLaunch sl);
cudaError_t launch_finalize_backsub(int Ks, int n_rows, const double *partial, double *red, SolverLaunch sl);

// ---- the plain structure's step in two launches (vg_solver_fast.cu) ------------------------------------------
// one rank (or several over peer memory, every rank with this structure), ONE dataset whose image i is the only block of free pose i (identity image -> element map, every element of
// the one free sequence observed and free), no prior / odometry blocks, 1 <= Ks <= FAST_MAX_KS
constexpr int FAST_MAX_KS = 12;
constexpr int FAST_GROUP = 20;       // blocks of fast_factor whose rows one of them folds
// host-mapped words of the polled step: [reduced solve's scalars 0..7 | 16..19: cost and the three model / norm sums, as
// the candidate's evaluation left them after its exchange | 24: its flag | candidate slab from FAST_HOST_SLAB]
constexpr int FAST_HOST_EVAL = 16, FAST_HOST_FLAG = 24, FAST_HOST_SLAB = 32;
// the LM loop's state on the device (vg_lm_dev.cuh): what the step's kernels read instead of host-supplied values, and what
// the factorisation copies over when the previous candidate was accepted (set C -> set A)
struct FastLm {
    LmState *st;                     // null: the host drives the loop (radius etc. from LmConsts)
    const double *H_alt;             // the other H buffer (LmState::hcur == 1: factorise this one)
    double *pose_a; const double *pose_c;     // n_pose x 6
    double *slab_a; const double *slab_c; int slab_n;
    double *red_a; const double *red_c; int red_n;      // segment E of the reduction buffers
    LmRecord *ring;                  // host-mapped ring of records (LM_RING entries) and, at the end of the solve, the final
    double *final_out;               // shared-parameter slab: published by one extra block of the factorisation kernel
};
// the shared parameters' slab positions and box bounds, by value (what SolveArgs::sh_off / sh_lo / sh_hi point to)
struct FastShared {
    int off[FAST_MAX_KS];
    double lo[FAST_MAX_KS], hi[FAST_MAX_KS];
};
struct FastDesc {
    const double *H;                 // n_pose x ne packed blocks (the set being factorised)
    int ne, W, pose_col, n_sl;
    int sl_col[FAST_MAX_KS], sl_idx[FAST_MAX_KS];
};
size_t fast_scratch(int n_pose, int Ks);      // doubles
int fast_factor_blocks(int n_pose);
int fast_backsub_blocks(int n_pose);
int fast_groups(int n_pose);                  // tickets needed: fast_groups + 1
// factorisation + Schur terms, then reduced solve + back-substitution: leaves exactly what launch_pose_schur (fused
// tail) + launch_pose_backsub (fused finalize) leave
const double *fast_partial_rows(const double *scratch, int n_pose, int Ks);     // rows of [model decrease, |step|^2, |x|^2]
cudaError_t launch_fast_step(const FastDesc &d, int n_pose, int Ks, double *scale, LmConsts lm, double *ws, double *scratch,
                             unsigned int *tickets, int *fail_flag, const SolveArgs &sa, const double *seq_cur,
                             double *seq_cand, bool backsub, SolverLaunch sl, cudaEvent_t between,
                             double *host_out, const PeerCtx *peer, const FastLm *flm, const FastShared &fs);
// host_out: host-mapped doubles, see FAST_HOST_SLAB; peer: several ranks, see fast_exchange; flm: the loop's state lives
// on the device (seq_cur / sa's "current" members = set A, the candidate ones = set C)

}  // namespace vg
