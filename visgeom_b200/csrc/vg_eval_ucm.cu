// vg_eval_ucm.cu -- instantiates the fused kernel (vg_eval_impl.cuh) for the UCM camera model,
// chain lengths 1..5.  One translation unit per model keeps the build parallel.
#include "vg_eval_impl.cuh"

namespace vg {

cudaError_t launch_model_ucm(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n, int *grid, bool query)
{
    return launch_model<MODEL_UCM>(L, a, s, n, grid, query);
}

long long smem_for_ucm(int L, int G, int P, int PCG) { return smem_for<MODEL_UCM>(L, G, P, PCG); }

}  // namespace vg
