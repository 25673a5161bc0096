// vg_host.cu -- vg_eval_chain: the inner boundary with HOST buffers, i.e. what GenericProjectionJac::Evaluate is to
// Ceres (calib_cost_functions.h:53-54: caller-owned parameter, residual and Jacobian arrays), batched over the images
// of a dataset.  The kernel needs ~30 us per 10 000 images; the call is bound by moving ~12 KB of Jacobians per image
// back over PCIe into the caller's pageable memory, so it is organised around that:
//   * the images are cut into chunks (a few MB of outputs); chunk c's upload + kernel + download run on one of two
//     non-blocking streams while the host copies chunk c-1 out of its pinned staging slot into the caller's arrays --
//     DMA at PCIe speed into pinned memory, never a pageable cudaMemcpy (which the driver stages and serialises);
//   * that last hop is a plain memcpy at host memory speed: for large calls a few helper threads (VG_HOST_COPY_THREADS,
//     default min(8, cores / 2), alive for the duration of the call) share it with the calling thread;
//   * an output array that is page-locked (the caller pinned it once with vg_host_register: Ceres keeps its Jacobian
//     arrays for the whole solve) is written by the DMA itself, without the staging hop and the memcpy;
//   * all state -- streams, events, pinned slots, device workspace -- belongs to the CALLING THREAD (thread_local):
//     Ceres may evaluate residual blocks from num_threads threads at once.
// No CPU path: without a CUDA device the call fails with VG_ERR_CUDA.
#include "vg_common.h"
#include "vg_eval.cuh"

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace vg {
namespace {

constexpr size_t ALIGN = 256;
size_t up(size_t b) { return (b + ALIGN - 1) & ~(ALIGN - 1); }

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    char *h_in = nullptr, *h_out = nullptr;      // pinned
    char *d_in = nullptr, *d_out = nullptr;
};

struct HostCtx {
    int dev = -1;
    size_t in_cap = 0, out_cap = 0, fixed_cap = 0;
    char *d_fixed = nullptr, *h_fixed = nullptr;  // intrinsics, board, global transforms
    Slot slot[2];

    void release()
    {
        if (dev < 0) return;
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(dev);
        for (Slot &s : slot) {
            if (s.st) cudaStreamSynchronize(s.st);
            if (s.h_in) cudaFreeHost(s.h_in);
            if (s.h_out) cudaFreeHost(s.h_out);
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.done) cudaEventDestroy(s.done);
            if (s.st) cudaStreamDestroy(s.st);
            s = Slot();
        }
        if (d_fixed) cudaFree(d_fixed);
        if (h_fixed) cudaFreeHost(h_fixed);
        d_fixed = h_fixed = nullptr;
        in_cap = out_cap = fixed_cap = 0;
        cudaSetDevice(cur);
        dev = -1;
    }
    ~HostCtx() { release(); }

    cudaError_t ensure(int device, size_t fixed, size_t in, size_t out)
    {
        if (dev >= 0 && dev != device) release();
        dev = device;
        cudaError_t e = cudaSuccess;
        for (Slot &s : slot) {
            if (!s.st && (e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if (!s.done && (e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if (fixed > fixed_cap) {
            if (d_fixed) cudaFree(d_fixed);
            if (h_fixed) cudaFreeHost(h_fixed);
            d_fixed = h_fixed = nullptr; fixed_cap = 0;
            if ((e = cudaMalloc(&d_fixed, fixed)) != cudaSuccess) return e;
            if ((e = cudaMallocHost(&h_fixed, fixed)) != cudaSuccess) return e;
            fixed_cap = fixed;
        }
        if (in > in_cap) {
            for (Slot &s : slot) {
                if (s.st) cudaStreamSynchronize(s.st);
                if (s.h_in) cudaFreeHost(s.h_in);
                if (s.d_in) cudaFree(s.d_in);
                s.h_in = s.d_in = nullptr;
            }
            in_cap = 0;
            for (Slot &s : slot) {
                if ((e = cudaMallocHost(&s.h_in, in)) != cudaSuccess) return e;
                if ((e = cudaMalloc(&s.d_in, in)) != cudaSuccess) return e;
            }
            in_cap = in;
        }
        if (out > out_cap) {
            for (Slot &s : slot) {
                if (s.st) cudaStreamSynchronize(s.st);
                if (s.h_out) cudaFreeHost(s.h_out);
                if (s.d_out) cudaFree(s.d_out);
                s.h_out = s.d_out = nullptr;
            }
            out_cap = 0;
            for (Slot &s : slot) {
                if ((e = cudaMallocHost(&s.h_out, out)) != cudaSuccess) return e;
                if ((e = cudaMalloc(&s.d_out, out)) != cudaSuccess) return e;
            }
            out_cap = out;
        }
        return cudaSuccess;
    }
};

thread_local std::unique_ptr<HostCtx> t_ctx;

// one contiguous piece of a chunk's outputs: pinned staging offset -> the caller's array
// direct: the caller's array is page-locked (vg_host_register, or allocated pinned): the DMA writes it, no staging hop
struct Piece { size_t off; char *user; size_t per_img; bool direct; };

bool page_locked(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// copy-out of a chunk, shared between the calling thread (part 0) and the helpers (parts 1 .. T-1)
void copy_share(const std::vector<Piece> &pieces, const char *h_out, size_t img0, size_t m, int part, int parts)
{
    for (const Piece &p : pieces) {
        if (p.direct) continue;
        const size_t bytes = m * p.per_img;
        const size_t lo = bytes * part / parts & ~size_t(63), hi = part + 1 == parts ? bytes : (bytes * (part + 1) / parts & ~size_t(63));
        if (hi > lo) memcpy(p.user + img0 * p.per_img + lo, h_out + p.off + lo, hi - lo);
    }
}

struct Team {
    std::atomic<int> ready{0};            // chunks whose outputs sit complete in their pinned slot
    std::atomic<int> abort{0};
    std::vector<std::atomic<int>> done;   // per chunk: helpers that have copied their share
    explicit Team(int chunks) : done(chunks) { for (auto &d : done) d.store(0); }
};

}  // namespace

void release_host_ctx() { t_ctx.reset(); }

}  // namespace vg

using namespace vg;

extern "C" int vg_host_register(void *p, size_t bytes)
{
    if (!p || bytes == 0) return fail(VG_ERR_INVALID, "vg_host_register: null pointer or empty range");
    if (vg_device_count() < 1) return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_host_register");
}

extern "C" int vg_host_unregister(void *p)
{
    if (!p) return fail(VG_ERR_INVALID, "vg_host_unregister: null pointer");
    const cudaError_t e = cudaHostUnregister(p);
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_host_unregister");
}

extern "C" int vg_eval_chain(int model, const double *intr, int n_img, int P,
                             const double *board, const double *obs,
                             int chain_len, const int *status, const int *is_global,
                             const double *const *xi,
                             double *r, double *J_intr, double *const *J_xi, double *H)
{
    const int K = vg_model_num_params(model);
    if (K < 0) return K;
    if (chain_len < 1) return fail(VG_ERR_INVALID, "empty transform chain");
    if (chain_len > VG_MAX_CHAIN)
        return fail(VG_ERR_INVALID, "the transform chain is too long (5 transforms at max are supproted)");   // :567
    if (n_img < 0 || P < 1) return fail(VG_ERR_INVALID, "n_img < 0 or P < 1");
    if (!intr || !board || !status || !is_global || !xi || (n_img > 0 && !obs)) return fail(VG_ERR_INVALID, "null input pointer");
    for (int e = 0; e < chain_len; e++)
        if (!xi[e]) return fail(VG_ERR_INVALID, "null transform pointer");
    if (vg_device_count() < 1) return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    if (eval_smem_bytes(model, chain_len, P, nullptr, nullptr) < 0)
        return fail(VG_ERR_UNSUPPORTED, "board has too many points for one CTA's shared memory");
    if (n_img == 0) return VG_OK;
    const int ne = vg_hessian_entries(model, chain_len);
    const size_t row_bytes = (size_t)2 * P * 8;

    // per-image bytes in / out, chunk size (multiple of 4 images: the kernel's group), piece offsets inside a slot
    size_t in_per = row_bytes, out_per = 0;
    for (int e = 0; e < chain_len; e++)
        if (!is_global[e]) in_per += 48;
    std::vector<Piece> pieces;
    if (r) out_per += row_bytes;
    if (J_intr) out_per += row_bytes * K;
    for (int e = 0; e < chain_len; e++)
        if (J_xi && J_xi[e]) out_per += row_bytes * 6;
    if (H) out_per += (size_t)ne * 8;
    static const size_t chunk_bytes = [] { const char *v = getenv("VG_HOST_CHUNK_MB"); return (size_t)(v ? atoi(v) : 8) << 20; }();
    size_t CH = chunk_bytes / std::max<size_t>(std::max(out_per, in_per), 1);
    CH = std::max<size_t>(64, CH & ~size_t(3));
    if (CH > (size_t)n_img) CH = (size_t)n_img;
    const int chunks = (int)(((size_t)n_img + CH - 1) / CH);
    size_t off = 0;
    auto add_piece = [&](double *user, size_t per) {
        pieces.push_back(Piece{off, reinterpret_cast<char *>(user), per, page_locked(user)});
        off += up(CH * per);
    };
    if (r) add_piece(r, row_bytes);
    if (J_intr) add_piece(J_intr, row_bytes * K);
    for (int e = 0; e < chain_len; e++)
        if (J_xi && J_xi[e]) add_piece(J_xi[e], row_bytes * 6);
    if (H) add_piece(H, (size_t)ne * 8);
    const size_t out_cap = std::max<size_t>(off, ALIGN);
    // inputs of a chunk: observations, then one block per sequence element
    size_t in_off[VG_MAX_CHAIN + 1];
    size_t ioff = up(CH * row_bytes);
    in_off[0] = 0;
    for (int e = 0; e < chain_len; e++) { in_off[1 + e] = ioff; if (!is_global[e]) ioff += up(CH * 48); }
    const size_t in_cap = ioff;
    // fixed inputs: intrinsics, board, global transforms
    const size_t f_board = up(8 * 16), f_glob = f_board + up((size_t)P * 24), fixed_cap = f_glob + up(48 * VG_MAX_CHAIN);

    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    if (!t_ctx) t_ctx.reset(new HostCtx());
    HostCtx &cx = *t_ctx;
    {
        const cudaError_t e = cx.ensure(dev, fixed_cap, in_cap, out_cap);
        if (e != cudaSuccess) { cx.release(); return fail_cuda(e, "vg_eval_chain workspace"); }
    }
    memcpy(cx.h_fixed, intr, 8 * K);
    memcpy(cx.h_fixed + f_board, board, (size_t)P * 24);
    for (int e = 0; e < chain_len; e++)
        if (is_global[e]) memcpy(cx.h_fixed + f_glob + 48 * e, xi[e], 48);
    cudaStream_t s0 = cx.slot[0].st;
    VG_CUDA(cudaMemcpyAsync(cx.d_fixed, cx.h_fixed, fixed_cap, cudaMemcpyHostToDevice, s0));
    VG_CUDA(cudaEventRecord(cx.slot[0].done, s0));
    VG_CUDA(cudaStreamWaitEvent(cx.slot[1].st, cx.slot[0].done, 0));

    // helper threads for the copy-out of large calls
    static const int max_helpers = [] {
        const char *v = getenv("VG_HOST_COPY_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        int t = v ? atoi(v) : std::min(8, std::max(1, hw / 2));      // host memcpy moves ~9 GB/s per core; PCIe brings ~50
        if (hw > 0 && t > hw) t = hw;
        return t < 1 ? 1 : t;
    }();
    size_t staged_per = 0;
    for (const Piece &p : pieces)
        if (!p.direct) staged_per += p.per_img;
    const size_t total_out = (size_t)n_img * staged_per;           // what still goes through the pinned slots
    const int parts = total_out >= ((size_t)8 << 20) ? max_helpers : 1;
    Team team(chunks);
    std::vector<std::thread> helpers;
    auto chunk_m = [&](int c) { return std::min(CH, (size_t)n_img - (size_t)c * CH); };
    for (int t = 1; t < parts; t++)
        helpers.emplace_back([&, t] {
            for (int c = 0; c < chunks; c++) {
                for (unsigned spins = 0; team.ready.load(std::memory_order_acquire) <= c; spins++) {
                    if (team.abort.load(std::memory_order_relaxed)) return;
                    if (spins > 64) std::this_thread::yield();
                }
                copy_share(pieces, cx.slot[c & 1].h_out, (size_t)c * CH, chunk_m(c), t, parts);
                team.done[c].fetch_add(1, std::memory_order_release);
            }
        });
    auto finish = [&](int code) {
        team.abort.store(1);
        for (auto &h : helpers) h.join();
        return code;
    };
    auto consume = [&](int c) -> int {
        Slot &s = cx.slot[c & 1];
        const cudaError_t e = cudaEventSynchronize(s.done);
        if (e != cudaSuccess) return fail_cuda(e, "vg_eval_chain: chunk");
        team.ready.store(c + 1, std::memory_order_release);
        copy_share(pieces, s.h_out, (size_t)c * CH, chunk_m(c), 0, parts);
        for (unsigned spins = 0; team.done[c].load(std::memory_order_acquire) < parts - 1; spins++)
            if (spins > 64) std::this_thread::yield();
        return VG_OK;
    };

    for (int c = 0; c < chunks; c++) {
        Slot &s = cx.slot[c & 1];
        const size_t i0 = (size_t)c * CH, m = chunk_m(c);
        // (the slot's previous tenant, chunk c - 2, was consumed before chunk c - 1 was issued... its copy-out is over)
        memcpy(s.h_in, obs + i0 * 2 * P, m * row_bytes);
        for (int e = 0; e < chain_len; e++)
            if (!is_global[e]) memcpy(s.h_in + in_off[1 + e], xi[e] + i0 * 6, m * 48);
        cudaError_t ce = cudaMemcpyAsync(s.d_in, s.h_in, in_cap, cudaMemcpyHostToDevice, s.st);
        if (ce != cudaSuccess) return finish(fail_cuda(ce, "vg_eval_chain: upload"));
        EvalArgs a;
        memset(&a, 0, sizeof a);
        a.intr = reinterpret_cast<const double *>(cx.d_fixed);
        a.board = reinterpret_cast<const double *>(cx.d_fixed + f_board);
        a.obs = reinterpret_cast<const double *>(s.d_in);
        a.n_img = (int)m; a.P = P;
        size_t pi = 0;
        if (r) a.r = reinterpret_cast<double *>(s.d_out + pieces[pi++].off);
        if (J_intr) a.Ja = reinterpret_cast<double *>(s.d_out + pieces[pi++].off);
        for (int e = 0; e < chain_len; e++) {
            a.xi[e] = is_global[e] ? reinterpret_cast<const double *>(cx.d_fixed + f_glob + 48 * e)
                                   : reinterpret_cast<const double *>(s.d_in + in_off[1 + e]);
            a.xi_stride[e] = is_global[e] ? 0 : 6;
            a.inverse[e] = status[e] == VG_TRANSFORM_INVERSE;
            if (J_xi && J_xi[e]) a.Je[e] = reinterpret_cast<double *>(s.d_out + pieces[pi++].off);
        }
        if (H) a.H = reinterpret_cast<double *>(s.d_out + pieces[pi++].off);
        ce = launch_eval(model, chain_len, a, s.st, &launch_counter());
        if (ce != cudaSuccess) return finish(fail_cuda(ce, "reproj_eval_kernel launch"));
        for (const Piece &p : pieces) {
            ce = cudaMemcpyAsync(p.direct ? p.user + i0 * p.per_img : s.h_out + p.off, s.d_out + p.off, m * p.per_img,
                                 cudaMemcpyDeviceToHost, s.st);
            if (ce != cudaSuccess) return finish(fail_cuda(ce, "vg_eval_chain: download"));
        }
        ce = cudaEventRecord(s.done, s.st);
        if (ce != cudaSuccess) return finish(fail_cuda(ce, "vg_eval_chain: event"));
        if (c >= 1) {
            const int rc = consume(c - 1);
            if (rc) return finish(rc);
        }
    }
    {
        const int rc = consume(chunks - 1);
        if (rc) return finish(rc);
    }
    for (auto &h : helpers) h.join();
    return VG_OK;
}
