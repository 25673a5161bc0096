// vg_lm_dev.cuh -- the Levenberg-Marquardt loop's control state, kept on the device.
//
// vg_problem_solve's accept / reject decision, trust-region update and termination tests (the outer loop of
// ceres::Solve, unified_calibration.cpp:39-53; Ceres' documented trust-region defaults) need a handful of scalars per
// iteration.  With the host taking that decision every iteration paid a device -> host -> device round trip plus three
// launch latencies on an idle GPU (22 us of an 81 us iteration at 10 000 images).  Here the thread that finishes
// the candidate's evaluation (the last CTA of reproj_eval_kernel's reduction tail, after the cross-rank exchange) takes
// the decision itself and leaves it in LmState; the kernels of the next iteration -- which the host has queued
// already, one iteration ahead -- read radius / current set / "done" from there.  The host only watches a ring of
// records in host-mapped memory (progress output, summary, and when to stop queueing); launches queued past the
// end of the solve publish the last record and return.
//
// Buffers: the solve starts with the current point in parameter set A (the problem's `cur`) and always evaluates
// candidates in set C (the other one): poses, shared-parameter slab and reduction buffer of an accepted candidate are
// copied C -> A by the next iteration's factorisation kernel (`migrate`; half a megabyte), the packed blocks H are not
// copied: evaluations alternate between the two H buffers (`hcur`).
#pragma once

namespace vg {

constexpr int LM_RING = 16;                 // records the host may lag behind (it queues one iteration ahead)

struct LmOptions {
    double gradient_tolerance, function_tolerance, parameter_tolerance, min_relative_decrease, min_radius, max_radius;
    int max_num_iterations, max_consecutive_invalid;
};

struct LmRecord {                           // host-mapped; seq is written last (release)
    unsigned long long seq;
    int iter;                               // iterations counted so far
    int done;                               // 0: running; else 1 + termination code (vg_solve_summary)
    int accepted, valid;                    // this pass' step
    int num_successful, num_unsuccessful;
    int migrate, hcur;                      // at done: where the final point lives (see above)
    unsigned long long epoch;               // next peer exchange number
    double cost, radius;                    // of the current point / for the next pass, after this pass' decision
    double prev_cost, prev_radius, new_cost, rho, step_norm, gmax;     // what the decision was taken on
};

// Device memory, one per problem.  The deciding thread leaves its record in `rec`; the NEXT launch of the chain (the
// factorisation kernel of the pass queued behind, which starts as the evaluation ends) copies it to the host ring, so
// that the evaluation kernel's end does not wait for writes that cross PCIe.
struct LmState {
    int done, limits, migrate, hcur;        // (one 16-byte load)
    int init_scale, iter, invalid_run, pad;
    double radius, decrease_factor, cost;
    int num_successful, num_unsuccessful;
    unsigned long long records;             // decided so far (+ the solve's base)
    unsigned long long published;           // ... of which the host ring has
    unsigned long long epoch;               // next peer exchange number (several ranks)
    LmOptions opt;
    LmRecord rec;                           // the latest decision
#ifdef VG_LM_STAMPS
    // developer build: %globaltimer at the start (min over blocks) and end (max) of the three kernels of each pass
    unsigned long long stamp[64][6];
    unsigned long long phase[2][10];        // block 10's way through fast_factor / fast_backsub, latest pass
    long long clk[16];                      // clock64 inside the reduced solve (block 10, warp 0)
#endif
};

#ifdef VG_LM_STAMPS
#define VG_LM_STAMP(st, k, is_end)                                                                                      \
    if ((st) && threadIdx.x == 0) {                                                                                     \
        unsigned long long t_;                                                                                          \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                          \
        unsigned long long *w_ = &(st)->stamp[((st)->records) & 63][2 * (k) + (is_end)];                                \
        if (is_end) atomicMax(w_, t_); else atomicMin(w_, t_);                                                          \
    }
#define VG_LM_PHASE(st, k, i)                                                                                           \
    if ((st) && threadIdx.x == 0 && blockIdx.x == 10) {                                                                 \
        unsigned long long t_;                                                                                          \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                          \
        (st)->phase[k][i] = t_;                                                                                         \
    }                                                                                                                   \
    __syncwarp();
#define VG_LM_CLK(st, i)                                                                                                \
    if ((st) && threadIdx.x == 0 && blockIdx.x == 10) (st)->clk[i] = clock64();                                         \
    __syncwarp();
#else
#define VG_LM_CLK(st, i)
#define VG_LM_STAMP(st, k, is_end)
#define VG_LM_PHASE(st, k, i)
#endif

}  // namespace vg
