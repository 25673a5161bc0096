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

static char *g_ws = nullptr;
static size_t g_ws_bytes = 0;
static int g_ws_dev = -1;

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

unsigned long long vg_launch_count(void) { return g_launches; }

void vg_release_workspace(void)
{
    if (g_ws) { cudaFree(g_ws); g_ws = nullptr; g_ws_bytes = 0; g_ws_dev = -1; }
}

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

int vg_eval_chain(int model, const double *intr, int n_img, int P,
                  const double *board, const double *obs,
                  int chain_len, const int *status, const int *is_global,
                  const double *const *xi,
                  double *r, double *J_intr, double *const *J_xi, double *H)
{
    int rc = check_eval_args(model, n_img, P, chain_len, intr, board, obs, status, is_global, xi);
    if (rc) return rc;
    if (vg_device_count() < 1) return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    const int K = model_K(model);
    const size_t rows = (size_t)n_img * 2 * P;
    const int ne = vg_hessian_entries(model, chain_len);
    // one device slab: inputs then outputs, every piece 256-byte aligned
    struct Piece { size_t off, bytes; };
    size_t total = 0;
    auto reserve = [&](size_t bytes) { Piece p{total, bytes}; total += (bytes + 255) & ~size_t(255); return p; };
    Piece p_intr = reserve(K * 8), p_board = reserve((size_t)P * 24), p_obs = reserve(rows * 8);
    Piece p_xi[VG_MAX_CHAIN], p_je[VG_MAX_CHAIN];
    for (int e = 0; e < chain_len; e++) p_xi[e] = reserve(is_global[e] ? 48 : (size_t)n_img * 48);
    Piece p_r = reserve(r ? rows * 8 : 0), p_ja = reserve(J_intr ? rows * K * 8 : 0);
    for (int e = 0; e < chain_len; e++) p_je[e] = reserve((J_xi && J_xi[e]) ? rows * 48 : 0);
    Piece p_h = reserve(H ? (size_t)n_img * ne * 8 : 0);
    // grow-only device workspace kept between calls (released by vg_release_workspace)
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    if (g_ws && (g_ws_dev != dev || g_ws_bytes < total)) { cudaFree(g_ws); g_ws = nullptr; g_ws_bytes = 0; }
    if (!g_ws) {
        VG_CUDA(cudaMalloc(&g_ws, total ? total : 256));
        g_ws_bytes = total ? total : 256;
        g_ws_dev = dev;
    }
    char *d = g_ws;
    cudaStream_t st = nullptr;
    auto cleanup = [&](int code) { return code; };
#define VG_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cleanup(fail_cuda(e__, #call)); } while (0)
    VG_TRY(cudaMemcpyAsync(d + p_intr.off, intr, p_intr.bytes, cudaMemcpyHostToDevice, st));
    VG_TRY(cudaMemcpyAsync(d + p_board.off, board, p_board.bytes, cudaMemcpyHostToDevice, st));
    if (rows) VG_TRY(cudaMemcpyAsync(d + p_obs.off, obs, p_obs.bytes, cudaMemcpyHostToDevice, st));
    const double *dxi[VG_MAX_CHAIN];
    double *dje[VG_MAX_CHAIN];
    for (int e = 0; e < chain_len; e++) {
        if (p_xi[e].bytes) VG_TRY(cudaMemcpyAsync(d + p_xi[e].off, xi[e], p_xi[e].bytes, cudaMemcpyHostToDevice, st));
        dxi[e] = reinterpret_cast<const double *>(d + p_xi[e].off);
        dje[e] = p_je[e].bytes ? reinterpret_cast<double *>(d + p_je[e].off) : nullptr;
    }
    rc = vg_eval_chain_dev(model, reinterpret_cast<const double *>(d + p_intr.off), n_img, P,
                           reinterpret_cast<const double *>(d + p_board.off),
                           reinterpret_cast<const double *>(d + p_obs.off), chain_len, status, is_global, dxi,
                           nullptr, p_r.bytes ? reinterpret_cast<double *>(d + p_r.off) : nullptr,
                           p_ja.bytes ? reinterpret_cast<double *>(d + p_ja.off) : nullptr, dje,
                           p_h.bytes ? reinterpret_cast<double *>(d + p_h.off) : nullptr, st);
    if (rc) return cleanup(rc);
    if (p_r.bytes) VG_TRY(cudaMemcpyAsync(r, d + p_r.off, p_r.bytes, cudaMemcpyDeviceToHost, st));
    if (p_ja.bytes) VG_TRY(cudaMemcpyAsync(J_intr, d + p_ja.off, p_ja.bytes, cudaMemcpyDeviceToHost, st));
    for (int e = 0; e < chain_len; e++)
        if (p_je[e].bytes) VG_TRY(cudaMemcpyAsync(J_xi[e], d + p_je[e].off, p_je[e].bytes, cudaMemcpyDeviceToHost, st));
    if (p_h.bytes) VG_TRY(cudaMemcpyAsync(H, d + p_h.off, p_h.bytes, cudaMemcpyDeviceToHost, st));
    VG_TRY(cudaStreamSynchronize(st));
#undef VG_TRY
    return cleanup(VG_OK);
}

}  // extern "C"
