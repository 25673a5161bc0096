// vg_eval.cuh -- launch interface of the fused reprojection kernels.
#pragma once

#include <cuda_runtime.h>

#include "vg_peer.cuh"
#include "vg_lm_dev.cuh"

namespace vg {

// the library-wide launch counter (vg_launch_count) may be bumped from several host threads at once
#ifndef VG_COUNT_LAUNCH_DEFINED
#define VG_COUNT_LAUNCH_DEFINED
inline void count_launch(unsigned long long *c) { __atomic_fetch_add(c, 1ull, __ATOMIC_RELAXED); }
#endif

constexpr int MAX_CHAIN = 5;
constexpr int EVAL_REDUCE_GROUP = 24;     // rows per first-level group of the fused reduction (592 CTAs: 25 groups)

// Tables of the fused shared-block reduction (built on the host once per problem): every entry of the reduced
// system (A_ij with i <= j, g_i, cost) lists the (dataset, packed entry) sums that feed it.
struct FinSrc { int off, ne, nb, e; };                       // offset of the dataset's sums, row stride, rows, packed entry
struct FinOut { int dst0, dst1; double scale; int src_begin, src_end; };   // red[] targets (dst1 = mirror or -1)
// The same assembly seen from the one dataset that feeds it alone: where entry e of its sums goes (dst0 < 0: nowhere)
struct EMapEntry { int dst0, dst1; double scale; };

// One dataset = one camera + one board + n_img images sharing a transform chain.
// All pointers are device memory.  Output pointers may be null.
struct EvalArgs {
    const double *intr;             // K
    const double *board;            // P x 3
    const double *obs;              // n_img x 2P
    const double *xi[MAX_CHAIN];    // 6 (global) or n_seq x 6 (sequence)
    int xi_stride[MAX_CHAIN];       // 0 global, 6 sequence
    int inverse[MAX_CHAIN];         // TRANSFORM_INVERSE flags
    const int *seq_index;           // nullable: image -> sequence element
    double *r;                      // n_img x 2P
    double *Ja;                     // n_img x 2P x K
    double *Je[MAX_CHAIN];          // n_img x 2P x 6
    double *H;                      // n_img x NE packed normal-equation blocks
    double *cta_partial;            // nullable: gridDim.x x NE, each CTA's sum of its images' blocks
    // Fused reduction of cta_partial inside the same launch (nullable: the caller reduces cta_partial itself).
    // The last CTA of every group of EVAL_REDUCE_GROUP rows sums its group into lvl1, the last of those sums lvl1
    // into ds_sum (both in index order: deterministic); counters return to zero.  If red is set (last dataset of
    // an evaluation) the same CTA then assembles the reduced system from all datasets' sums (fin_base).
    unsigned int *tickets;          // 1 + ceil(rows / EVAL_REDUCE_GROUP) counters, zero between launches
    double *lvl1;                   // ceil(rows / EVAL_REDUCE_GROUP) x NE
    double *ds_sum;                 // NE
    const FinOut *fin_outs;
    const FinSrc *fin_srcs;
    int n_fin_out;
    const double *fin_base;         // all datasets' ds_sum regions
    const EMapEntry *emap;          // nullable: NE entries, set when this dataset is the only source of every entry of red
    double *red;
    // nullable: host-mapped words the CTA that assembled (and exchanged) red also leaves red[host_index .. + host_count)
    // and, after them, host_seq in, for a host that polls instead of synchronising the stream (vg_problem_solve)
    double *host_value;
    unsigned long long *host_flag;
    unsigned long long host_seq;
    int host_index, host_count;
    // Robust loss on the block of an image (Ceres applies a LossFunction to the squared norm s of the whole residual
    // block): 0 = none, else b = a^2 of SoftLOneLoss(a), rho(s) = 2 b (sqrt(1 + s / b) - 1).  rho'' < 0, so Ceres'
    // corrector scales residuals and Jacobians by sqrt(rho'): the packed block leaves as rho' [J r]^T [J r] with
    // rho(s) in its last entry; r / J outputs stay raw (what Evaluate returns).
    double loss_b;
    // Multi-GPU (vg_peer.cuh): the CTA that assembles the reduced system also sums its first peer_count doubles
    // across the ranks over peer memory, in the same launch.  peer.n <= 1: nothing to exchange.
    // peer_deferred: the tail only posts this rank's block; the sum is formed later -- by the head of this problem's
    // next launch (collect / collect_buf: the exchange it still owes, and the buffer its sum belongs in) or by a
    // one-block kernel when something needs it sooner.  collect_done: device word, the last exchange collected.
    PeerCtx peer;
    int peer_count;
    int peer_deferred;                  // 1: the tail posts; 2: nobody posts here -- the collecting head (collect_post) does
    int collect_post;                   // the head posts collect_buf (this rank's block of that exchange) before it collects
    PeerCtx collect;
    double *collect_buf;
    unsigned long long *collect_done;
    int n_img;
    int P;
    int late_wait;                      // set by the launcher: the main loop does not wait for the launch ahead (vg_eval_impl.cuh)
    int wait_at_head;                   // set by the caller: a kernel ahead writes this launch's inputs and releases its dependents early
    // The LM loop's decision on the device (vg_lm_dev.cuh).  lm_mode 1: this is the solve's first evaluation -- the
    // thread that finishes red[] records the cost; 2: a candidate's evaluation -- every CTA leaves at once when the solve
    // is over (or only the gradient test is due), the packed blocks go to H or H_alt (LmState::hcur), and the finishing
    // thread decides.  lm_so: the reduced solve's scalars about this step (SOLVE_OUT); red + host_index: cost and the
    // three pose sums.
    int lm_mode;
    int lm_partial_rows;            // mode 2: rows of lm_partial = [model decrease, |step|^2, |x|^2] per block of the
    const double *lm_partial;       // back-substitution; the last CTA adds them up at its head -> red[host_index + 1 .. + 3]
    LmState *lm;
    const LmState *lm_init;         // mode 1: the loop's initial state in host-mapped memory; block 0 copies it to *lm under
                                    // the kernel's own work (no upload in front of the solve's first launch)
    const double *lm_so;
    double *H_alt;
};

// returns cudaError_t of the launch (cudaSuccess, or cudaErrorInvalidValue for an
// unsupported shape); *launches is incremented by the number of kernels launched
cudaError_t launch_eval(int model, int chain_len, const EvalArgs &args, cudaStream_t stream,
                        unsigned long long *launches, int *grid_out = nullptr);

// the grid (number of persistent CTAs = rows of cta_partial) launch_eval will use for this shape
cudaError_t eval_grid_size(int model, int chain_len, int n_img, int P, int *grid_out);

// shared memory the kernel would need for this shape (bytes); <0 if unsupported
long long eval_smem_bytes(int model, int chain_len, int P, int *images_per_cta, int *threads);

}  // namespace vg
