// vg_eval_eucm.cu -- instantiates the fused kernel (vg_eval_impl.cuh) for the EUCM camera model,
// chain lengths 1..5.  One translation unit per model keeps the build parallel.
#include "vg_eval_impl.cuh"

namespace vg {

cudaError_t launch_model_eucm(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n, int *grid, bool query)
{
    return launch_model<MODEL_EUCM>(L, a, s, n, grid, query);
}

long long smem_for_eucm(int L, int G, int P, int PCG) { return smem_for<MODEL_EUCM>(L, G, P, PCG); }

}  // namespace vg

#ifdef VG_PHASE_CLOCKS
// developer build only: phase counters of the EUCM instantiations
extern "C" int vg_debug_phase_clocks(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out, vg::g_phase_clocks, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(vg::g_phase_clocks, z, sizeof(z)); }
    return e == cudaSuccess ? 0 : -2;
}
#endif
