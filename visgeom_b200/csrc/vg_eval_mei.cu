// vg_eval_mei.cu -- instantiates the fused kernel (vg_eval_impl.cuh) for the MEI camera model,
// chain lengths 1..5.  One translation unit per model keeps the build parallel.
#include "vg_eval_impl.cuh"

namespace vg {

cudaError_t launch_model_mei(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n, int *grid, bool query)
{
    return launch_model<MODEL_MEI>(L, a, s, n, grid, query);
}

long long smem_for_mei(int L, int G, int P, int PCG) { return smem_for<MODEL_MEI>(L, G, P, PCG); }

}  // namespace vg
