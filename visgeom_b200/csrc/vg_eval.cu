// vg_eval.cu -- launch planning and model dispatch of the fused reprojection kernels
// (the kernels themselves: vg_eval_impl.cuh, instantiated per model in vg_eval_{eucm,ucm,mei}.cu).
#include "vg_eval.cuh"

#include <cstdlib>


namespace vg {

constexpr int MODEL_EUCM = 0, MODEL_UCM = 1, MODEL_MEI = 2;
struct LaunchPlan { int G, threads, PCG; long long smem; };

cudaError_t launch_model_eucm(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n, int *grid, bool query);
cudaError_t launch_model_ucm(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n, int *grid, bool query);
cudaError_t launch_model_mei(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n, int *grid, bool query);
long long smem_for_eucm(int L, int G, int P, int PCG);
long long smem_for_ucm(int L, int G, int P, int PCG);
long long smem_for_mei(int L, int G, int P, int PCG);

static long long smem_any(int model, int L, int G, int P, int PCG)
{
    switch (model) {
    case MODEL_EUCM: return smem_for_eucm(L, G, P, PCG);
    case MODEL_UCM: return smem_for_ucm(L, G, P, PCG);
    case MODEL_MEI: return smem_for_mei(L, G, P, PCG);
    default: return -1;
    }
}

// G images per group (G*P <= 256 so that one corner pass covers a group; at most 8: one
// normal-equation warp per image), PCG groups per pose prologue (about 4.5 KB of pose records).
bool plan_eval(int model, int L, int P, LaunchPlan *pl)
{
    if (P < 1 || L < 1 || L > MAX_CHAIN) return false;
    const long long LIMIT = 227 * 1024;
    const int pose_doubles = ((12 + 21 * L) + 1) & ~1;
    for (int G = 4; G >= 1; G >>= 1) {
        if (G > 1 && G * P > 224) continue;
        int t = ((G * P + 31) / 32) * 32;
        if (t < 96) t = 96;
        if (t > 224) t = 224;   // __launch_bounds__ of the kernel
        int pcg = 5632 / (pose_doubles * 8 * G);
        if (pcg * G > 32) pcg = 32 / G;     // one warp restages a batch of poses (G <= 4)
        static const int pcg_env = [] { const char *e = getenv("VG_PCG"); return e ? atoi(e) : 0; }();   // developer knob
        if (pcg_env > 0 && pcg_env * G <= 32) pcg = pcg_env;
        if (pcg < 1) pcg = 1;
        long long b = smem_any(model, L, G, P, pcg);
        if (b < 0) return false;
        if (b > LIMIT) { pcg = 1; b = smem_any(model, L, G, P, pcg); }
        if (b > LIMIT) continue;
        pl->G = G; pl->threads = t; pl->PCG = pcg; pl->smem = b;
        return true;
    }
    return false;
}

long long eval_smem_bytes(int model, int chain_len, int P, int *images_per_cta, int *threads)
{
    LaunchPlan pl;
    if (!plan_eval(model, chain_len, P, &pl)) return -1;
    if (images_per_cta) *images_per_cta = pl.G;
    if (threads) *threads = pl.threads;
    return pl.smem;
}

static cudaError_t launch_any(int model, int chain_len, const EvalArgs &args, cudaStream_t stream,
                              unsigned long long *launches, int *grid, bool query)
{
    switch (model) {
    case MODEL_EUCM: return launch_model_eucm(chain_len, args, stream, launches, grid, query);
    case MODEL_UCM: return launch_model_ucm(chain_len, args, stream, launches, grid, query);
    case MODEL_MEI: return launch_model_mei(chain_len, args, stream, launches, grid, query);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_eval(int model, int chain_len, const EvalArgs &args, cudaStream_t stream,
                        unsigned long long *launches, int *grid_out)
{
    return launch_any(model, chain_len, args, stream, launches, grid_out, false);
}

cudaError_t eval_grid_size(int model, int chain_len, int n_img, int P, int *grid_out)
{
    EvalArgs a{};
    a.n_img = n_img; a.P = P;
    return launch_any(model, chain_len, a, nullptr, nullptr, grid_out, true);
}

}  // namespace vg
