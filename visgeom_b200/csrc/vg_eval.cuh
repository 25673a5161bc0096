// vg_eval.cuh -- launch interface of the fused reprojection kernels.
#pragma once

#include <cuda_runtime.h>

namespace vg {

constexpr int MAX_CHAIN = 5;

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
    int n_img;
    int P;
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
