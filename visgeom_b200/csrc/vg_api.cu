// vg_api.cu -- C ABI, inner boundary: batched GenericProjectionJac::Evaluate
// (include/visgeom_b200.h).  There is no CPU path here: without a CUDA device
// every compute entry point returns VG_ERR_CUDA.
#include "vg_common.h"
#include "vg_eval.cuh"

#include <cstdint>
#include <cstring>
#include <vector>

namespace vg {

static thread_local std::string g_error;
static unsigned long long g_launches = 0;

void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg) { g_error = msg; return code; }
int fail_cuda(cudaError_t e, const char *where)
{
    g_error = std::string("CUDA error in ") + where + ": " + cudaGetErrorString(e);
    cudaGetLastError();   // clear the sticky-free error state
    return VG_ERR_CUDA;
}
unsigned long long &launch_counter() { return g_launches; }

void release_host_ctx();     // vg_host.cu: the calling thread's streams, pinned slots and device workspace

static int model_K(int model)
{
    switch (model) {
    case VG_MODEL_EUCM: return 6;    // eucm.h:64
    case VG_MODEL_UCM: return 5;     // ucm.h:61
    case VG_MODEL_MEI: return 10;    // mei.h:69
    default: return -1;
    }
}

}  // namespace vg

using namespace vg;

extern "C" {

const char *vg_last_error(void) { return g_error.c_str(); }
int vg_version(void) { return 100; }

int vg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int vg_model_num_params(int model)
{
    const int K = model_K(model);
    return K < 0 ? fail(VG_ERR_INVALID, "invalid camera model name") : K;   // unified_calibration.cpp:177
}

int vg_model_bounds(int model, int idx, double *lower, double *upper)
{
    const int K = model_K(model);
    if (K < 0 || idx < 0 || idx >= K) return fail(VG_ERR_INVALID, "vg_model_bounds: bad model or index");
    double lo, hi;
    if (model == VG_MODEL_EUCM) {            // eucm.h:228-246
        lo = idx == 0 ? 0.0 : (idx == 1 ? 0.1 : 1.0);
        hi = idx == 0 ? 1.0 : (idx == 1 ? 10.0 : 1e5);
    } else if (model == VG_MODEL_UCM) {      // ucm.h:199-215
        lo = idx == 0 ? 0.0 : 1.0;
        hi = idx == 0 ? 3.0 : 1e5;
    } else {                                 // mei.h:287-313
        lo = idx == 0 ? 0.0 : (idx <= 5 ? -10.0 : 1.0);
        hi = idx == 0 ? 3.0 : (idx <= 5 ? 10.0 : 1e5);
    }
    if (lower) *lower = lo;
    if (upper) *upper = hi;
    return VG_OK;
}

int vg_hessian_entries(int model, int chain_len)
{
    const int K = model_K(model);
    if (K < 0 || chain_len < 0 || chain_len > VG_MAX_CHAIN) return fail(VG_ERR_INVALID, "vg_hessian_entries: bad arguments");
    const int W = K + 6 * chain_len + 1;
    return W * (W + 1) / 2;
}

unsigned long long vg_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

void vg_release_workspace(void) { release_host_ctx(); }

static int check_eval_args(int model, int n_img, int P, int chain_len, const void *intr, const void *board,
                           const void *obs, const int *status, const int *is_global, const void *xi)
{
    if (model_K(model) < 0) return fail(VG_ERR_INVALID, "invalid camera model name");
    if (chain_len < 1) return fail(VG_ERR_INVALID, "empty transform chain");
    if (chain_len > VG_MAX_CHAIN)
        return fail(VG_ERR_INVALID, "the transform chain is too long (5 transforms at max are supproted)");  // :567
    if (n_img < 0 || P < 1) return fail(VG_ERR_INVALID, "n_img < 0 or P < 1");
    if (!intr || !board || !status || !is_global || !xi || (n_img > 0 && !obs))
        return fail(VG_ERR_INVALID, "null input pointer");
    return VG_OK;
}

int vg_eval_chain_dev(int model, const double *intr, int n_img, int P,
                      const double *board, const double *obs,
                      int chain_len, const int *status, const int *is_global,
                      const double *const *xi, const int *seq_index,
                      double *r, double *J_intr, double *const *J_xi, double *H, void *stream)
{
    int rc = check_eval_args(model, n_img, P, chain_len, intr, board, obs, status, is_global, xi);
    if (rc) return rc;
    EvalArgs a;
    memset(&a, 0, sizeof a);
    a.intr = intr; a.board = board; a.obs = obs; a.seq_index = seq_index;
    a.r = r; a.Ja = J_intr; a.H = H; a.n_img = n_img; a.P = P;
    auto misaligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
    if (misaligned(obs) || misaligned(r) || misaligned(J_intr))
        return fail(VG_ERR_INVALID, "device buffers must be 16-byte aligned");
    for (int e = 0; e < chain_len; e++) {
        if (!xi[e]) return fail(VG_ERR_INVALID, "null transform pointer");
        a.xi[e] = xi[e];
        a.xi_stride[e] = is_global[e] ? 0 : 6;
        a.inverse[e] = status[e] == VG_TRANSFORM_INVERSE;
        a.Je[e] = J_xi ? J_xi[e] : nullptr;
        if (misaligned(a.Je[e])) return fail(VG_ERR_INVALID, "device buffers must be 16-byte aligned");
    }
    if (eval_smem_bytes(model, chain_len, P, nullptr, nullptr) < 0)
        return fail(VG_ERR_UNSUPPORTED, "board has too many points for one CTA's shared memory");
    cudaError_t e = launch_eval(model, chain_len, a, static_cast<cudaStream_t>(stream), &g_launches);
    if (e != cudaSuccess) return fail_cuda(e, "reproj_eval_kernel launch");
    return VG_OK;
}

}  // extern "C"
