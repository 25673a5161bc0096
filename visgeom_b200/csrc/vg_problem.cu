// vg_problem.cu -- C ABI, outer boundary: the problem GenericCameraCalibration
// assembles (unified_calibration.cpp:514-630) and the Levenberg-Marquardt solve that
// stands in for ceres::Solve (:39-53).  The loop follows Ceres' documented
// trust-region defaults: Jacobi scaling fixed at iteration 0, LM
// diagonal clamp(diag(J^T J),1e-6,1e32)/radius, step accepted when the relative
// decrease exceeds 1e-3, radius /= max(1/3, 1-(2 rho-1)^3) on accept, radius /= nu
// with nu doubling on reject, box bounds by projection.  Everything that touches an
// image or a pose runs on the GPU; the host only solves the Ks x Ks reduced system
// (Ks <= a few dozen) and takes the accept / reject decision.
#include "vg_common.h"
#include "vg_eval.cuh"
#include "vg_solver_kernels.cuh"
#include "vg_priors.cuh"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace vg;

namespace {

constexpr int CAM_STRIDE = 16;   // doubles per camera in the shared-parameter slab

struct Cam {
    int model, K, constant, shared_off;
    double params[VG_MAX_INTRINSICS], lo[VG_MAX_INTRINSICS], hi[VG_MAX_INTRINSICS];
};

struct Tr {
    int is_global, constant, n;
    int shared_off, pose_off;      // free global / free sequence, else -1
    int glob_slot;                 // index among global transforms (slab position)
    int seq_slot;                  // index among sequence transforms
    std::vector<double> host;      // n x 6
    std::vector<unsigned char> fixed;   // per element: constant ("anchor", unified_calibration.cpp:803-806)
    double *dev[2];                // sequences: device poses (set 0 / 1); constants alias one buffer
};

// TransformationPrior block (unified_calibration.cpp:808-829) and the OdometryPrior blocks of one "odometry"
// dataset (:742-807)
struct TPrior { int tr, index; double stiffness[6], xi_prior[6]; };
struct Odom { int tr; double errV, errW, lambda; std::vector<double> odom; };

struct Ds {
    int cam, P, n_img, L, D, W, ne;
    int tr[VG_MAX_CHAIN], status[VG_MAX_CHAIN];
    int seq_tr;                    // transform id of the chain's sequence element
    double loss_a;                 // SoftLOneLoss(a) on every block of the dataset; 0: none (NULL loss)
    std::vector<int> seq_index;
    bool identity_index;
    double *d_board, *d_obs;
    int *d_seq_index;
    double *d_H[2];
    double *d_r, *d_Ja, *d_Je[VG_MAX_CHAIN];   // only when Jacobians are materialised
};

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int chol(std::vector<double> &M, int n)
{
    for (int j = 0; j < n; j++) {
        double s = M[j * n + j];
        for (int k = 0; k < j; k++) s -= M[j * n + k] * M[j * n + k];
        if (!(s > 0.0)) return -1;
        s = std::sqrt(s);
        M[j * n + j] = s;
        for (int i = j + 1; i < n; i++) {
            double t = M[i * n + j];
            for (int k = 0; k < j; k++) t -= M[i * n + k] * M[j * n + k];
            M[i * n + j] = t / s;
        }
    }
    return 0;
}

void chol_solve(const std::vector<double> &Lm, int n, double *x)
{
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= Lm[i * n + k] * x[k];
        x[i] = s / Lm[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= Lm[k * n + i] * x[k];
        x[i] = s / Lm[i * n + i];
    }
}

double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace

struct vg_problem {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    std::vector<Cam> cams;
    std::vector<Tr> trs;
    std::vector<Ds> dss;
    std::vector<TPrior> tps;
    std::vector<Odom> odoms;
    int n_glob = 0, n_seq = 0;
    bool materialize = false;

    // prepared state
    bool prepared = false;
    // Evaluation launches normally do not wait for the launch ahead of them before their main loop (vg_eval_impl.cuh);
    // that rests on the launch ahead not writing what an evaluation reads.  Inside vg_problem_solve the kernels ahead do
    // (the back-substitution writes the candidate's poses) and release their dependents early: there, and for the first
    // evaluation after a solve, every evaluation waits at its head.
    bool in_solve = false, after_solve = false;
    int Ks = 0, n_pose = 0, cur = 0;
    int rank = 0, nranks = 1;
    vg_allreduce_fn allreduce = nullptr;
    void *allreduce_ctx = nullptr;
    double *d_slab[2] = {nullptr, nullptr};   // [cams | globals], per parameter set
    size_t slab_doubles = 0;
    std::vector<double> h_slab;
    DatasetDesc *d_desc[2] = {nullptr, nullptr};
    std::vector<DatasetDesc> h_desc[2];
    int *d_pose_start = nullptr, *d_contrib_ds = nullptr, *d_contrib_img = nullptr;
    int *d_pose_seq = nullptr, *d_pose_local = nullptr, *d_fail = nullptr;
    std::vector<int> h_acc_tab;                 // offset of each dataset's cta_partial region
    FinOut *d_fin_out = nullptr;
    FinSrc *d_fin_src = nullptr;
    EMapEntry *d_emap = nullptr;                // direct assembly from the last dataset's sums (vg_eval.cuh), or null
    int n_fin_out = 0;
    double **d_seq_ptr[2] = {nullptr, nullptr};
    double *d_scale = nullptr, *d_ws = nullptr, *d_partial = nullptr, *d_delta = nullptr;
    // per parameter set: [segment E | segment S | reduced_solve's scalars | candidate slab] (vg_solver_kernels.cuh)
    double *d_redbuf[2] = {nullptr, nullptr};
    int red_doubles = 0;
    // shared parameter j: its slab position and box bounds; Jacobi scaling of the shared block
    int *d_sh_off = nullptr;
    // peer-memory exchange (vg_peer.cuh): own inbox, every rank's inbox as mapped here, exchanges done so far
    unsigned long long *d_inbox = nullptr;
    unsigned long long **d_peer_ptrs = nullptr;
    std::vector<void *> peer_opened;
    bool peers = false, exchanged_in_kernel = false;
    // one process per rank (IPC handles): a deferred exchange is posted by the head of the problem's NEXT launch (or by the
    // fetch), which every rank issues as well; ranks of one process (connect_local, possibly driven from one thread
    // that fetches them one after the other) post in the tail of the launch itself
    bool post_at_head = false;
    unsigned long long epoch = 0;
    // deferred exchange (vg_problem_evaluate_async): posted by an evaluation kernel, not collected yet
    bool pending = false;
    unsigned long long pending_epoch = 0;
    int pending_set = 0;
    unsigned long long *d_collect_done = nullptr;
    unsigned long long *h_peer_fail = nullptr;  // host-mapped: the number of an exchange whose collect gave up (vg_peer.cuh)
    long long peer_spin_limit = 0;
    PeerCtx peer_ctx(unsigned long long e) const { return PeerCtx{d_peer_ptrs, rank, nranks, e, h_peer_fail, peer_spin_limit}; }
    PeerCtx next_peer_ctx() { return peer_ctx(++epoch); }
    unsigned int *d_solver_tickets = nullptr;   // "last block done" counters of pose_factor / pose_backsub
    // the plain structure's two-launch step (vg_solver_fast.cu): dataset / sequence it applies to, or -1
    int fast_ds = -1, fast_tr = -1;
    double *d_fast_scratch = nullptr;
    unsigned int *d_fast_tickets = nullptr;
    // host-mapped words the step's kernels leave their scalars in, and the flag the host polls (no copy, no stream sync)
    double *d_agree = nullptr;                  // several ranks: do all of them have the plain structure? (one exchange per solve)
    double *h_poll = nullptr;
    unsigned long long poll_seq = 0;            // != 0: the next evaluation's last launch posts cost + flag
    unsigned long long poll_counter = 0;
    // the LM loop's control state on the device (vg_lm_dev.cuh): state, its pinned upload staging, the host-mapped ring of
    // records the host watches and the host-mapped copy of the final shared-parameter slab
    LmState *d_lm = nullptr, *h_lm_stage = nullptr;
    LmRecord *h_lm_ring = nullptr;
    double *h_lm_final = nullptr;
    unsigned long long lm_seq_base = 0;
    int lm_eval_mode = 0, lm_set_a = 0;         // != 0: the next evaluate_set is part of that loop (EvalArgs::lm_mode)
    double *d_sh_lo = nullptr, *d_sh_hi = nullptr, *d_scale_a = nullptr;
    bool bounds_on_device = false;              // d_sh_lo / d_sh_hi hold what h_up holds
    FastShared fast_shared;                     // the same and the slab positions, by value, for the fast step's kernels (Ks <= FAST_MAX_KS)
    std::vector<int> h_sh_off;
    double *d_cta_partial = nullptr;
    // fused reduction (vg_eval.cuh): per dataset its tickets / level-1 rows, all datasets' sums, table offsets
    unsigned int *d_tickets = nullptr;
    double *d_lvl1 = nullptr, *d_ds_sum = nullptr;
    std::vector<int> h_ticket_off, h_lvl1_off, h_sum_off;
    int last_ds = -1;                           // the last dataset with images: its launch assembles the reduced system
    size_t partial_doubles = 0, cta_partial_doubles = 0;
    double *h_red = nullptr;                  // pinned
    double *h_up = nullptr;                   // pinned upload staging (the shared parameters' bounds at the start of a solve)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double eval_ms = 0;
    int n_eval = 0;
    // prior blocks and coupled / constant sequence elements (vg_priors.cuh); all empty in a plain grid problem
    int n_tp = 0, n_op = 0, n_tp_shared = 0, n_seg = 0;
    double *d_tp_const = nullptr, *d_op_const = nullptr, *d_tp_out[2] = {nullptr, nullptr}, *d_op_out[2] = {nullptr, nullptr};
    const double **d_tp_xi[2] = {nullptr, nullptr}, **d_op_xi[2] = {nullptr, nullptr};
    int *d_tp_shared_rec = nullptr, *d_tp_shared_off = nullptr;
    int *d_seg_start = nullptr, *d_seg_len = nullptr, *d_extra_start = nullptr, *d_extra_kind = nullptr, *d_extra_rec = nullptr,
        *d_prev_edge = nullptr;
    unsigned char *d_mask = nullptr, *d_fixed = nullptr;
    double *d_chain_off = nullptr, *d_chain_w = nullptr;

    PriorTables prior_tables(int s) const
    {
        PriorTables t;
        t.n_tp = n_tp; t.n_op = n_op; t.tp_const = d_tp_const; t.op_const = d_op_const;
        t.tp_xi = d_tp_xi[s]; t.op_xi = d_op_xi[s]; t.tp_out = d_tp_out[s]; t.op_out = d_op_out[s];
        t.n_tp_shared = n_tp_shared; t.tp_shared_rec = d_tp_shared_rec; t.tp_shared_off = d_tp_shared_off;
        return t;
    }
    ChainTables chain_tables(int s) const
    {
        ChainTables c;
        c.n_seg = n_seg; c.seg_start = d_seg_start; c.seg_len = d_seg_len; c.mask = d_mask; c.fixed = d_fixed;
        c.extra_start = d_extra_start; c.extra_kind = d_extra_kind; c.extra_rec = d_extra_rec; c.prev_edge = d_prev_edge;
        c.tp_const = d_tp_const; c.tp_out = d_tp_out[s]; c.op_out = d_op_out[s]; c.off = d_chain_off; c.w = d_chain_w;
        return c;
    }

    double *cam_ptr(int set, int cam) const { return d_slab[set] + (size_t)cam * CAM_STRIDE; }
    double *glob_ptr(int set, int slot) const { return d_slab[set] + cams.size() * CAM_STRIDE + (size_t)slot * 6; }
};

namespace {

void free_prepared(vg_problem *p)
{
    auto F = [](auto *&ptr) { if (ptr) { cudaFree(ptr); ptr = nullptr; } };
    if (p->prepared) {
        // the device holds the current sequence poses (a solve leaves them there): bring the host copies up to date
        // before the buffers that say which set is current go away
        for (Tr &t : p->trs)
            if (!t.is_global && !t.constant)
                cudaMemcpyAsync(t.host.data(), t.dev[p->cur], sizeof(double) * 6 * t.n, cudaMemcpyDeviceToHost, p->stream);
        cudaStreamSynchronize(p->stream);
    }
    for (int s = 0; s < 2; s++) { F(p->d_slab[s]); F(p->d_desc[s]); F(p->d_seq_ptr[s]); }
    F(p->d_pose_start); F(p->d_contrib_ds); F(p->d_contrib_img); F(p->d_pose_seq); F(p->d_pose_local);
    F(p->d_fail); F(p->d_fin_out); F(p->d_fin_src); F(p->d_emap); F(p->d_cta_partial); F(p->d_tickets); F(p->d_lvl1); F(p->d_ds_sum); F(p->d_scale); F(p->d_ws); F(p->d_partial); F(p->d_redbuf[0]); F(p->d_redbuf[1]); F(p->d_delta);
    F(p->d_solver_tickets); F(p->d_fast_scratch); F(p->d_fast_tickets); F(p->d_agree); F(p->d_sh_off); F(p->d_sh_lo); F(p->d_sh_hi); F(p->d_scale_a);
    for (int s = 0; s < 2; s++) { F(p->d_tp_out[s]); F(p->d_op_out[s]); F(p->d_tp_xi[s]); F(p->d_op_xi[s]); }
    F(p->d_tp_const); F(p->d_op_const); F(p->d_tp_shared_rec); F(p->d_tp_shared_off); F(p->d_seg_start); F(p->d_seg_len);
    F(p->d_extra_start); F(p->d_extra_kind); F(p->d_extra_rec); F(p->d_prev_edge); F(p->d_mask); F(p->d_fixed);
    F(p->d_chain_off); F(p->d_chain_w);
    p->n_tp = p->n_op = p->n_tp_shared = p->n_seg = 0;
    p->pending = false;                 // an exchange posted but never asked for: its words are simply overwritten
    if (p->h_red) { cudaFreeHost(p->h_red); p->h_red = nullptr; }
    if (p->h_up) { cudaFreeHost(p->h_up); p->h_up = nullptr; }
    if (p->h_poll) { cudaFreeHost(p->h_poll); p->h_poll = nullptr; }
    if (p->h_lm_stage) { cudaFreeHost(p->h_lm_stage); p->h_lm_stage = nullptr; }
    if (p->h_lm_ring) { cudaFreeHost(p->h_lm_ring); p->h_lm_ring = nullptr; }
    if (p->h_lm_final) { cudaFreeHost(p->h_lm_final); p->h_lm_final = nullptr; }
    F(p->d_lm);
    for (auto &d : p->dss)
        for (int s = 0; s < 2; s++) F(d.d_H[s]);
    p->prepared = false;
}

void fill_slab(const vg_problem *p, double *dst)
{
    memset(dst, 0, p->slab_doubles * sizeof(double));
    for (size_t c = 0; c < p->cams.size(); c++)
        memcpy(dst + c * CAM_STRIDE, p->cams[c].params, sizeof(double) * p->cams[c].K);
    for (const Tr &t : p->trs)
        if (t.is_global) memcpy(dst + p->cams.size() * CAM_STRIDE + (size_t)t.glob_slot * 6, t.host.data(), 48);
}

// prior records, the poses that gather them, and the segments the chain kernels walk
int prepare_priors(vg_problem *p)
{
    const int NP = p->n_pose;
    if (p->tps.empty() && p->odoms.empty()) {
        bool any_fixed = false;
        for (const Tr &t : p->trs)
            if (t.pose_off >= 0)
                for (unsigned char f : t.fixed) any_fixed = any_fixed || f;
        if (!any_fixed) return VG_OK;
    }
    std::vector<double> tp_in, op_in;
    std::vector<const double *> tp_xi[2], op_xi[2];
    std::vector<int> sh_rec, sh_off, prev_edge(NP ? NP : 1, -1);
    std::vector<std::vector<std::pair<int, int>>> extras(NP);
    std::vector<unsigned char> mask(NP ? NP : 1, 0), fixed(NP ? NP : 1, 0), linked_next(NP ? NP : 1, 0);
    for (const TPrior &tp : p->tps) {
        const Tr &t = p->trs[tp.tr];
        if (t.is_global && p->rank != 0) continue;       // shared blocks are counted once across ranks
        const int rec = (int)tp_xi[0].size();
        tp_in.insert(tp_in.end(), tp.stiffness, tp.stiffness + 6);
        tp_in.insert(tp_in.end(), tp.xi_prior, tp.xi_prior + 6);
        for (int s = 0; s < 2; s++)
            tp_xi[s].push_back(t.is_global ? p->glob_ptr(s, t.glob_slot) : t.dev[s] + (size_t)6 * tp.index);
        if (t.is_global) { if (t.shared_off >= 0) { sh_rec.push_back(rec); sh_off.push_back(t.shared_off); } }
        else if (t.pose_off >= 0) extras[t.pose_off + tp.index].push_back({EXTRA_TP, rec});
    }
    for (const Odom &od : p->odoms) {
        const Tr &t = p->trs[od.tr];
        for (int i = 0; i + 1 < t.n; i++) {
            const int e = (int)op_xi[0].size();
            op_in.push_back(od.errV); op_in.push_back(od.errW); op_in.push_back(od.lambda);
            op_in.insert(op_in.end(), od.odom.begin() + (size_t)6 * i, od.odom.begin() + (size_t)6 * (i + 2));
            for (int s = 0; s < 2; s++) op_xi[s].push_back(t.dev[s] + (size_t)6 * i);
            if (t.pose_off >= 0) {
                const int q = t.pose_off + i;
                extras[q].push_back({EXTRA_OP_FIRST, e});
                extras[q + 1].push_back({EXTRA_OP_SECOND, e});
                prev_edge[q + 1] = e;
                linked_next[q] = 1;
            }
        }
    }
    p->n_tp = (int)tp_xi[0].size(); p->n_op = (int)op_xi[0].size(); p->n_tp_shared = (int)sh_rec.size();
    for (const Tr &t : p->trs)
        if (t.pose_off >= 0)
            for (int i = 0; i < t.n; i++) fixed[t.pose_off + i] = t.fixed[i];
    std::vector<int> extra_start(NP + 1, 0), extra_kind, extra_rec, seg_start, seg_len;
    for (int q = 0; q < NP; q++) {
        extra_start[q] = (int)extra_kind.size();
        for (auto &x : extras[q]) { extra_kind.push_back(x.first); extra_rec.push_back(x.second); }
        mask[q] = (!extras[q].empty() || fixed[q]) ? 1 : 0;
    }
    extra_start[NP] = (int)extra_kind.size();
    for (int q = 0; q < NP; q++) {
        if (!mask[q] || prev_edge[q] >= 0) continue;     // not a segment head
        int len = 1;
        while (linked_next[q + len - 1]) len++;
        seg_start.push_back(q); seg_len.push_back(len);
    }
    p->n_seg = (int)seg_start.size();

    auto up_d = [&](double *&dptr, const std::vector<double> &v, size_t min_n) -> cudaError_t {
        cudaError_t e = cudaMalloc(&dptr, sizeof(double) * (v.size() > min_n ? v.size() : min_n));
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(dptr, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice);
    };
    auto up_i = [&](int *&dptr, const std::vector<int> &v) -> cudaError_t {
        cudaError_t e = cudaMalloc(&dptr, sizeof(int) * (v.size() ? v.size() : 1));
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(dptr, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice);
    };
    auto up_b = [&](unsigned char *&dptr, const std::vector<unsigned char> &v) -> cudaError_t {
        cudaError_t e = cudaMalloc(&dptr, v.size() ? v.size() : 1);
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(dptr, v.data(), v.size(), cudaMemcpyHostToDevice);
    };
    double *d_tp_in = nullptr, *d_op_in = nullptr;
    VG_CUDA(up_d(d_tp_in, tp_in, 1)); VG_CUDA(up_d(d_op_in, op_in, 1));
    VG_CUDA(cudaMalloc(&p->d_tp_const, sizeof(double) * TP_CONST * (size_t)(p->n_tp + 1)));
    VG_CUDA(cudaMalloc(&p->d_op_const, sizeof(double) * OP_CONST * (size_t)(p->n_op + 1)));
    for (int s = 0; s < 2; s++) {
        VG_CUDA(cudaMalloc(&p->d_tp_out[s], sizeof(double) * TP_OUT * (size_t)(p->n_tp + 1)));
        VG_CUDA(cudaMalloc(&p->d_op_out[s], sizeof(double) * OP_OUT * (size_t)(p->n_op + 1)));
        VG_CUDA(cudaMalloc(&p->d_tp_xi[s], sizeof(double *) * (size_t)(p->n_tp + 1)));
        VG_CUDA(cudaMalloc(&p->d_op_xi[s], sizeof(double *) * (size_t)(p->n_op + 1)));
        if (p->n_tp) VG_CUDA(cudaMemcpy(p->d_tp_xi[s], tp_xi[s].data(), sizeof(double *) * p->n_tp, cudaMemcpyHostToDevice));
        if (p->n_op) VG_CUDA(cudaMemcpy(p->d_op_xi[s], op_xi[s].data(), sizeof(double *) * p->n_op, cudaMemcpyHostToDevice));
    }
    VG_CUDA(up_i(p->d_tp_shared_rec, sh_rec)); VG_CUDA(up_i(p->d_tp_shared_off, sh_off));
    VG_CUDA(up_i(p->d_seg_start, seg_start)); VG_CUDA(up_i(p->d_seg_len, seg_len));
    VG_CUDA(up_i(p->d_extra_start, extra_start)); VG_CUDA(up_i(p->d_extra_kind, extra_kind)); VG_CUDA(up_i(p->d_extra_rec, extra_rec));
    VG_CUDA(up_i(p->d_prev_edge, prev_edge));
    VG_CUDA(up_b(p->d_mask, mask)); VG_CUDA(up_b(p->d_fixed, fixed));
    VG_CUDA(cudaMalloc(&p->d_chain_off, sizeof(double) * 36 * (size_t)(NP ? NP : 1)));
    VG_CUDA(cudaMalloc(&p->d_chain_w, sizeof(double) * 6 * (size_t)(NP ? NP : 1)));
    VG_CUDA(cudaMemset(p->d_chain_off, 0, sizeof(double) * 36 * (size_t)(NP ? NP : 1)));
    SolverLaunch sl{p->stream, &launch_counter()};
    cudaError_t e = launch_prior_setup(p->n_tp, d_tp_in, p->d_tp_const, p->n_op, d_op_in, p->d_op_const, sl);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(d_tp_in); cudaFree(d_op_in);
    if (e != cudaSuccess) return fail_cuda(e, "prior setup");
    return VG_OK;
}

// (re)build everything that depends on the problem structure
int prepare(vg_problem *p)
{
    if (p->prepared) return VG_OK;
    VG_CUDA(cudaSetDevice(p->device));
    // parameter layout: free cameras then free globals are the shared block
    int off = 0, po = 0;
    for (Cam &c : p->cams) { c.shared_off = c.constant ? -1 : off; if (!c.constant) off += c.K; }
    for (Tr &t : p->trs) {
        t.shared_off = t.pose_off = -1;
        if (t.constant) continue;
        if (t.is_global) { t.shared_off = off; off += 6; }
        else { t.pose_off = po; po += t.n; }
    }
    p->Ks = off;
    p->n_pose = po;
    const int Ks = p->Ks, NP = p->n_pose;
    // limits of the shared block (the reference has none; these are the engine's): the reduced solve keeps
    // 2 Ks^2 + 9 Ks doubles in 40 KB of shared memory, a rank's slot of the peer exchange holds PEER_SLOT_DOUBLES
    if (sizeof(double) * (2 * (size_t)Ks * Ks + 9 * (size_t)Ks + 1) > 40 * 1024)
        return fail(VG_ERR_UNSUPPORTED, "too many free shared parameters (cameras + global transforms): at most 48 are supported");
    if (red_size(Ks, p->nranks) > PEER_SLOT_DOUBLES)
        return fail(VG_ERR_UNSUPPORTED, "the shared block does not fit a peer-exchange slot");

    p->slab_doubles = p->cams.size() * CAM_STRIDE + (size_t)p->n_glob * 6 + 2;
    p->h_slab.assign(p->slab_doubles, 0.0);
    for (int s = 0; s < 2; s++) VG_CUDA(cudaMalloc(&p->d_slab[s], p->slab_doubles * sizeof(double)));
    VG_CUDA(cudaMallocHost(&p->h_up, (p->slab_doubles + Ks + 2) * sizeof(double)));
    p->bounds_on_device = false;
    VG_CUDA(cudaMalloc(&p->d_delta, (Ks + 2) * sizeof(double)));

    // dataset descriptors (one per parameter set: they differ in the H pointer)
    for (int s = 0; s < 2; s++) p->h_desc[s].assign(p->dss.size(), DatasetDesc());
    std::vector<std::vector<std::pair<int, int>>> contrib(NP);
    for (size_t k = 0; k < p->dss.size(); k++) {
        Ds &d = p->dss[k];
        const Cam &c = p->cams[d.cam];
        DatasetDesc dd;
        memset(&dd, 0, sizeof dd);
        dd.n_img = d.n_img; dd.ne = d.ne; dd.W = d.W;
        dd.seq_index = d.d_seq_index;
        dd.pose_col = -1; dd.pose_base = -1; dd.n_sl = 0;
        for (int a = 0; a < c.K; a++) {
            dd.kind[a] = c.shared_off >= 0 ? COL_SHARED : COL_CONST;
            dd.idx[a] = c.shared_off + a;
        }
        for (int e = 0; e < d.L; e++) {
            const Tr &t = p->trs[d.tr[e]];
            for (int q = 0; q < 6; q++) {
                const int a = c.K + 6 * e + q;
                if (t.is_global) { dd.kind[a] = t.shared_off >= 0 ? COL_SHARED : COL_CONST; dd.idx[a] = t.shared_off + q; }
                else { dd.kind[a] = t.pose_off >= 0 ? COL_POSE : COL_CONST; dd.idx[a] = q; }
            }
            if (!t.is_global && t.pose_off >= 0) { dd.pose_col = c.K + 6 * e; dd.pose_base = t.pose_off; }
        }
        dd.kind[d.D] = COL_RESID; dd.idx[d.D] = 0;
        for (int a = 0; a < d.D; a++)
            if (dd.kind[a] == COL_SHARED) { dd.sl_col[dd.n_sl] = a; dd.sl_idx[dd.n_sl] = dd.idx[a]; dd.n_sl++; }
        for (int s = 0; s < 2; s++) {
            if (!d.d_H[s]) VG_CUDA(cudaMalloc(&d.d_H[s], sizeof(double) * (size_t)d.ne * (d.n_img ? d.n_img : 1)));
            dd.H = d.d_H[s];
            p->h_desc[s][k] = dd;
        }
        if (dd.pose_base >= 0)
            for (int i = 0; i < d.n_img; i++) contrib[dd.pose_base + d.seq_index[i]].push_back({(int)k, i});
    }
    for (int s = 0; s < 2; s++) {
        VG_CUDA(cudaMalloc(&p->d_desc[s], sizeof(DatasetDesc) * (p->dss.size() ? p->dss.size() : 1)));
        if (!p->dss.empty())
            VG_CUDA(cudaMemcpy(p->d_desc[s], p->h_desc[s].data(), sizeof(DatasetDesc) * p->dss.size(), cudaMemcpyHostToDevice));
    }
    // pose list: CSR of (dataset, image) contributions, owning sequence and local index
    std::vector<int> pose_start(NP + 1, 0), cds, cimg, pseq(NP ? NP : 1, 0), ploc(NP ? NP : 1, 0);
    for (int q = 0; q < NP; q++) {
        pose_start[q] = (int)cds.size();
        for (auto &pr : contrib[q]) { cds.push_back(pr.first); cimg.push_back(pr.second); }
    }
    pose_start[NP] = (int)cds.size();
    std::vector<double *> seq_ptr[2];
    for (const Tr &t : p->trs) {
        if (t.is_global) continue;
        for (int s = 0; s < 2; s++) seq_ptr[s].push_back(t.dev[s]);
        if (t.pose_off >= 0)
            for (int i = 0; i < t.n; i++) { pseq[t.pose_off + i] = t.seq_slot; ploc[t.pose_off + i] = i; }
    }
    auto upload_i = [&](int *&dptr, const std::vector<int> &v) -> cudaError_t {
        cudaError_t e = cudaMalloc(&dptr, sizeof(int) * (v.size() ? v.size() : 1));
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(dptr, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice);
    };
    VG_CUDA(upload_i(p->d_pose_start, pose_start));
    VG_CUDA(upload_i(p->d_contrib_ds, cds));
    VG_CUDA(upload_i(p->d_contrib_img, cimg));
    VG_CUDA(upload_i(p->d_pose_seq, pseq));
    VG_CUDA(upload_i(p->d_pose_local, ploc));
    for (int s = 0; s < 2; s++) {
        VG_CUDA(cudaMalloc(&p->d_seq_ptr[s], sizeof(double *) * (seq_ptr[s].size() ? seq_ptr[s].size() : 1)));
        if (!seq_ptr[s].empty())
            VG_CUDA(cudaMemcpy(p->d_seq_ptr[s], seq_ptr[s].data(), sizeof(double *) * seq_ptr[s].size(), cudaMemcpyHostToDevice));
    }
    std::vector<int> grids(p->dss.size() + 1, 0);
    for (size_t k = 0; k < p->dss.size(); k++)
        VG_CUDA(eval_grid_size(p->cams[p->dss[k].cam].model, p->dss[k].L, p->dss[k].n_img, p->dss[k].P, &grids[k]));
    // cta_partial regions (one row per persistent CTA), tickets and level-1 rows of every dataset
    const int n_ds = (int)p->dss.size();
    p->h_acc_tab.assign(n_ds + 1, 0); p->h_ticket_off.assign(n_ds + 1, 0); p->h_lvl1_off.assign(n_ds + 1, 0);
    p->last_ds = -1;
    for (int k = 0; k < n_ds; k++) {
        const int ne = p->h_desc[0][k].ne;
        const int ngrp = (grids[k] + EVAL_REDUCE_GROUP - 1) / EVAL_REDUCE_GROUP;
        p->h_acc_tab[k + 1] = p->h_acc_tab[k] + grids[k] * ne;
        p->h_ticket_off[k + 1] = p->h_ticket_off[k] + 1 + ngrp;
        p->h_lvl1_off[k + 1] = p->h_lvl1_off[k] + ngrp * ne;
        if (p->dss[k].n_img > 0) p->last_ds = k;
    }
    p->cta_partial_doubles = (size_t)p->h_acc_tab[n_ds] + 8;
    // the reduced system is assembled from one row of sums per dataset
    std::vector<int> one_row(n_ds + 1, 1);
    std::vector<FinOut> fin_out;
    std::vector<FinSrc> fin_src;
    size_t sum_doubles = 0;
    build_finalize_tables(p->h_desc[0].data(), one_row.data(), n_ds, Ks, p->h_sum_off, fin_out, fin_src, &sum_doubles);
    VG_CUDA(cudaMalloc(&p->d_tickets, sizeof(unsigned int) * (p->h_ticket_off[n_ds] + 1)));
    VG_CUDA(cudaMemset(p->d_tickets, 0, sizeof(unsigned int) * (p->h_ticket_off[n_ds] + 1)));
    VG_CUDA(cudaMalloc(&p->d_lvl1, sizeof(double) * (p->h_lvl1_off[n_ds] + 1)));
    VG_CUDA(cudaMalloc(&p->d_ds_sum, sizeof(double) * sum_doubles));
    VG_CUDA(cudaMemset(p->d_ds_sum, 0, sizeof(double) * sum_doubles));
    p->n_fin_out = (int)fin_out.size();
    VG_CUDA(cudaMalloc(&p->d_fin_out, sizeof(FinOut) * (fin_out.size() + 1)));
    VG_CUDA(cudaMalloc(&p->d_fin_src, sizeof(FinSrc) * (fin_src.size() + 1)));
    VG_CUDA(cudaMemcpy(p->d_fin_out, fin_out.data(), sizeof(FinOut) * fin_out.size(), cudaMemcpyHostToDevice));
    if (!fin_src.empty())
        VG_CUDA(cudaMemcpy(p->d_fin_src, fin_src.data(), sizeof(FinSrc) * fin_src.size(), cudaMemcpyHostToDevice));
    // when every entry of the reduced system has exactly one source and all of them lie in the last dataset's sums,
    // that dataset's launch writes them straight from its reduction
    if (p->last_ds >= 0) {
        const int ne = p->h_desc[0][p->last_ds].ne;
        std::vector<EMapEntry> emap(ne, EMapEntry{-1, -1, 0.0});
        bool ok = true;
        for (const FinOut &f : fin_out) {
            if (f.src_end - f.src_begin != 1) { ok = false; break; }
            const FinSrc &sr = fin_src[f.src_begin];
            if (sr.off != p->h_sum_off[p->last_ds] || sr.e < 0 || sr.e >= ne || emap[sr.e].dst0 >= 0) { ok = false; break; }
            emap[sr.e] = EMapEntry{f.dst0, f.dst1, f.scale};
        }
        if (ok) {
            VG_CUDA(cudaMalloc(&p->d_emap, sizeof(EMapEntry) * ne));
            VG_CUDA(cudaMemcpy(p->d_emap, emap.data(), sizeof(EMapEntry) * ne, cudaMemcpyHostToDevice));
        }
    }
    VG_CUDA(cudaMalloc(&p->d_cta_partial, sizeof(double) * p->cta_partial_doubles));
    VG_CUDA(cudaMalloc(&p->d_fail, sizeof(int)));
    VG_CUDA(cudaMemset(p->d_fail, 0, sizeof(int)));
    VG_CUDA(cudaMalloc(&p->d_scale, sizeof(double) * 6 * (size_t)(NP ? NP : 1)));
    VG_CUDA(cudaMalloc(&p->d_ws, sizeof(double) * (size_t)pose_ws_stride(Ks) * (NP ? NP : 1)));
    {
        int rc = prepare_priors(p);
        if (rc) return rc;
    }
    p->partial_doubles = pose_scratch(NP, Ks, p->n_seg) + 3 * (size_t)pose_backsub_blocks(NP) + 64;
    VG_CUDA(cudaMalloc(&p->d_partial, sizeof(double) * p->partial_doubles));
    const int rs = p->red_doubles = red_size(Ks, p->nranks) + SOLVE_OUT + (int)p->slab_doubles;
    for (int s = 0; s < 2; s++) {
        VG_CUDA(cudaMalloc(&p->d_redbuf[s], sizeof(double) * rs));
        VG_CUDA(cudaMemset(p->d_redbuf[s], 0, sizeof(double) * rs));
    }
    VG_CUDA(cudaMallocHost(&p->h_red, sizeof(double) * rs));
    // where each shared parameter lives in the slab (cameras first, then free globals: the order of shared_off)
    p->h_sh_off.assign(Ks ? Ks : 1, 0);
    for (size_t ci = 0; ci < p->cams.size(); ci++)
        if (p->cams[ci].shared_off >= 0)
            for (int k = 0; k < p->cams[ci].K; k++) p->h_sh_off[p->cams[ci].shared_off + k] = (int)(ci * CAM_STRIDE) + k;
    for (const Tr &t : p->trs)
        if (t.shared_off >= 0)
            for (int k = 0; k < 6; k++) p->h_sh_off[t.shared_off + k] = (int)(p->cams.size() * CAM_STRIDE) + t.glob_slot * 6 + k;
    VG_CUDA(cudaMalloc(&p->d_solver_tickets, 2 * sizeof(unsigned int)));
    VG_CUDA(cudaMemset(p->d_solver_tickets, 0, 2 * sizeof(unsigned int)));
    VG_CUDA(cudaMalloc(&p->d_sh_off, sizeof(int) * (Ks ? Ks : 1)));
    VG_CUDA(cudaMemcpy(p->d_sh_off, p->h_sh_off.data(), sizeof(int) * (Ks ? Ks : 1), cudaMemcpyHostToDevice));
    VG_CUDA(cudaMalloc(&p->d_sh_lo, sizeof(double) * (Ks ? Ks : 1)));
    VG_CUDA(cudaMalloc(&p->d_sh_hi, sizeof(double) * (Ks ? Ks : 1)));
    VG_CUDA(cudaMalloc(&p->d_scale_a, sizeof(double) * (Ks ? Ks : 1)));
    // the plain structure (see vg_solver_kernels.cuh): its step runs in two launches
    p->fast_ds = p->fast_tr = -1;
    if ((p->nranks == 1 || p->peers) && p->n_seg == 0 && p->n_tp + p->n_op == 0 && Ks >= 1 && Ks <= FAST_MAX_KS && NP > 0) {
        int n_free_seq = 0, tr_id = -1, n_ds_img = 0, ds_id = -1;
        for (size_t i = 0; i < p->trs.size(); i++)
            if (p->trs[i].pose_off >= 0) { n_free_seq++; tr_id = (int)i; }
        for (size_t k = 0; k < p->dss.size(); k++)
            if (p->dss[k].n_img > 0) { n_ds_img++; ds_id = (int)k; }
        if (n_free_seq == 1 && n_ds_img == 1) {
            const Ds &d = p->dss[ds_id];
            const Tr &t = p->trs[tr_id];
            bool ok = d.seq_tr == tr_id && d.identity_index && d.n_img == t.n && t.n == NP && t.pose_off == 0 &&
                      p->h_desc[0][ds_id].pose_col >= 0 && p->h_desc[0][ds_id].n_sl <= FAST_MAX_KS;
            for (unsigned char f : t.fixed) ok = ok && !f;
            if (ok) { p->fast_ds = ds_id; p->fast_tr = tr_id; }
        }
    }
    if (p->peers && p->nranks > 1) VG_CUDA(cudaMalloc(&p->d_agree, 2 * sizeof(double)));
    if (p->fast_ds >= 0) {
        VG_CUDA(cudaMalloc(&p->d_fast_scratch, sizeof(double) * fast_scratch(NP, Ks)));
        VG_CUDA(cudaMalloc(&p->d_fast_tickets, sizeof(unsigned int) * (fast_groups(NP) + 2)));
        VG_CUDA(cudaMemset(p->d_fast_tickets, 0, sizeof(unsigned int) * (fast_groups(NP) + 2)));
        VG_CUDA(cudaHostAlloc(&p->h_poll, sizeof(double) * (FAST_HOST_SLAB + p->slab_doubles + 2), cudaHostAllocMapped));
        memset(p->h_poll, 0, sizeof(double) * (FAST_HOST_SLAB + p->slab_doubles + 2));
        VG_CUDA(cudaMalloc(&p->d_lm, sizeof(LmState)));
        VG_CUDA(cudaHostAlloc(&p->h_lm_stage, sizeof(LmState), cudaHostAllocMapped));
        VG_CUDA(cudaHostAlloc(&p->h_lm_ring, sizeof(LmRecord) * LM_RING, cudaHostAllocMapped));
        memset(p->h_lm_ring, 0, sizeof(LmRecord) * LM_RING);
        VG_CUDA(cudaHostAlloc(&p->h_lm_final, sizeof(double) * (p->slab_doubles + 2), cudaHostAllocMapped));
    }
    p->cur = 0;
    // both parameter sets start from the host values
    fill_slab(p, p->h_slab.data());
    for (int s = 0; s < 2; s++)
        VG_CUDA(cudaMemcpy(p->d_slab[s], p->h_slab.data(), p->slab_doubles * sizeof(double), cudaMemcpyHostToDevice));
    for (Tr &t : p->trs)
        if (!t.is_global)
            for (int s = 0; s < 2; s++)
                if (s == 0 || t.dev[1] != t.dev[0])
                    VG_CUDA(cudaMemcpy(t.dev[s], t.host.data(), sizeof(double) * 6 * t.n, cudaMemcpyHostToDevice));
    p->prepared = true;
    return VG_OK;
}

// after a synchronisation with the device: did a collect give up on a rank that never posted?
int check_peer_fail(vg_problem *p)
{
    if (!p->h_peer_fail) return VG_OK;
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(p->h_peer_fail);
    if (!e) return VG_OK;
    *p->h_peer_fail = 0;
    return fail(VG_ERR_PEER, "peer exchange " + std::to_string(e) + " timed out on rank " + std::to_string(p->rank) +
                                 ": a rank never posted its block (it died, or issued a different sequence of evaluations)");
}

// fused residual + Jacobian + normal-equation kernels of every dataset at parameter set s,
// then the shared-block reduction -> segment E (A, g_a, cost) of that set's reduction buffer
// the sum of an exchange that an evaluation kernel posted and nothing has collected yet
int flush_pending(vg_problem *p)
{
    if (!p->pending) return VG_OK;
    p->pending = false;
    SolverLaunch sl{p->stream, &launch_counter()};
    const PeerCtx pc = p->peer_ctx(p->pending_epoch);
    cudaError_t e = launch_peer_collect(p->d_redbuf[p->pending_set], red_segE_size(p->Ks), pc, p->d_collect_done, p->post_at_head, sl);
    if (e != cudaSuccess) return fail_cuda(e, "peer collect");
    return VG_OK;
}

// deferred: several GPUs over peer memory -- the kernel only posts its block, the sum is formed by this problem's
// next launch (or by flush_pending when something needs it sooner), so the NVLink round trip is off the critical path
int evaluate_set(vg_problem *p, int s, bool timed, bool deferred = false)
{
    deferred = deferred && p->peers && p->nranks > 1 && p->n_tp + p->n_op == 0 && p->last_ds >= 0;
    if (!deferred) {
        int rc = flush_pending(p);
        if (rc) return rc;
    }
    if (timed) VG_CUDA(cudaEventRecord(p->ev0, p->stream));
    for (Ds &d : p->dss) {
        EvalArgs a;
        memset(&a, 0, sizeof a);
        a.intr = p->cam_ptr(s, d.cam);
        a.board = d.d_board; a.obs = d.d_obs;
        a.seq_index = d.identity_index ? nullptr : d.d_seq_index;
        for (int e = 0; e < d.L; e++) {
            const Tr &t = p->trs[d.tr[e]];
            a.xi[e] = t.is_global ? p->glob_ptr(s, t.glob_slot) : t.dev[s];
            a.xi_stride[e] = t.is_global ? 0 : 6;
            a.inverse[e] = d.status[e] == VG_TRANSFORM_INVERSE;
            a.Je[e] = p->materialize ? d.d_Je[e] : nullptr;
        }
        a.r = p->materialize ? d.d_r : nullptr;
        a.Ja = p->materialize ? d.d_Ja : nullptr;
        a.H = d.d_H[s];
        const int k = (int)(&d - p->dss.data());
        a.cta_partial = p->d_cta_partial + p->h_acc_tab[k];
        a.tickets = p->d_tickets + p->h_ticket_off[k];
        a.lvl1 = p->d_lvl1 + p->h_lvl1_off[k];
        a.ds_sum = p->d_ds_sum + p->h_sum_off[k];
        if (k == p->last_ds) {      // the shared-block reduction -> segment E is the tail of this launch
            a.fin_outs = p->d_fin_out; a.fin_srcs = p->d_fin_src; a.n_fin_out = p->n_fin_out;
            a.fin_base = p->d_ds_sum; a.red = p->d_redbuf[s]; a.emap = p->d_emap;
            if (p->poll_seq) {
                a.host_value = p->h_poll + FAST_HOST_EVAL;
                a.host_flag = reinterpret_cast<unsigned long long *>(p->h_poll + FAST_HOST_FLAG);
                a.host_seq = p->poll_seq; a.host_index = red_off_cost(p->Ks); a.host_count = 4;   // cost, model, step^2, x^2
                p->poll_seq = 0;
            }
            if (p->lm_eval_mode) {
                // part of the LM loop that runs on the device (solve_on_device): s is set A (mode 1) or set C (mode 2)
                const int sa_ = p->lm_set_a, sc_ = sa_ ^ 1;
                a.lm_mode = p->lm_eval_mode; a.lm = p->d_lm; a.lm_init = p->h_lm_stage;
                a.lm_so = p->d_redbuf[sc_] + red_size(p->Ks, p->nranks);
                a.H_alt = d.d_H[sa_];
                a.host_index = red_off_cost(p->Ks);
                a.lm_partial = fast_partial_rows(p->d_fast_scratch, p->n_pose, p->Ks);
                a.lm_partial_rows = fast_backsub_blocks(p->n_pose);
            }
            if (p->peers && p->nranks > 1 && p->n_tp + p->n_op == 0) {
                a.peer_count = red_segE_size(p->Ks);
                if (deferred) {
                    a.peer_deferred = p->post_at_head ? 2 : 1;
                    if (p->pending) {
                        a.collect_post = p->post_at_head ? 1 : 0;
                        a.collect = p->peer_ctx(p->pending_epoch);
                        a.collect_buf = p->d_redbuf[p->pending_set];
                        a.collect_done = p->d_collect_done;
                    }
                    a.peer = p->next_peer_ctx();
                    p->pending = true; p->pending_epoch = a.peer.epoch; p->pending_set = s;
                } else {
                    a.peer = p->next_peer_ctx();
                    p->exchanged_in_kernel = true;
                }
            }
        }
        a.n_img = d.n_img; a.P = d.P;
        a.wait_at_head = p->in_solve || p->after_solve;
        a.loss_b = d.loss_a * d.loss_a;
        cudaError_t e = launch_eval(p->cams[d.cam].model, d.L, a, p->stream, &launch_counter());
        if (e != cudaSuccess) return fail_cuda(e, "reproj_eval_kernel launch");
    }
    p->after_solve = false;
    // no dataset with images on this rank (fewer images than ranks, or a problem of prior blocks only): nothing has
    // written segment E, and what it holds is the cross-rank SUM of this set's previous evaluation -- it must not be
    // contributed again
    if (p->last_ds < 0) VG_CUDA(cudaMemsetAsync(p->d_redbuf[s], 0, sizeof(double) * red_off_model(p->Ks), p->stream));
    if (p->n_tp + p->n_op > 0) {
        // the 6-residual blocks: their normal-equation pieces, then cost and shared-block terms on top of the reduced system
        SolverLaunch sl{p->stream, &launch_counter()};
        cudaError_t e = launch_prior_eval(p->prior_tables(s), p->Ks, p->d_redbuf[s], 1, nullptr, sl);
        if (e != cudaSuccess) return fail_cuda(e, "prior_eval launch");
    }
    if (timed) VG_CUDA(cudaEventRecord(p->ev1, p->stream));
    p->n_eval++;
    return VG_OK;
}

// sum a segment of a set's reduction buffer across ranks (if any)
int exchange_segment(vg_problem *p, int s, int off, int count)
{
    if (count > 0 && p->peers && p->nranks > 1) {
        int rc = flush_pending(p);
        if (rc) return rc;
        // peer memory: segment E has usually been exchanged by the evaluation kernel itself
        const bool done = off == 0 && p->exchanged_in_kernel;
        p->exchanged_in_kernel = false;
        if (done) return VG_OK;
        SolverLaunch sl{p->stream, &launch_counter()};
        cudaError_t e = launch_peer_exchange(p->d_redbuf[s] + off, count, p->next_peer_ctx(), sl);
        if (e != cudaSuccess) return fail_cuda(e, "peer exchange");
        return VG_OK;
    }
    if (count > 0 && p->allreduce && p->nranks > 1) {
        if (p->allreduce(p->allreduce_ctx, p->d_redbuf[s] + off, count, p->stream) != 0)
            return fail(VG_ERR_CUDA, "all-reduce callback failed");
    }
    return VG_OK;
}

// exchange a segment across ranks and fetch `fetch` doubles from its start (>= count: what follows the segment
// is local to the rank)
int fetch_segment(vg_problem *p, int s, int off, int count, int fetch = 0)
{
    int rc = exchange_segment(p, s, off, count);
    if (rc) return rc;
    if (fetch < count) fetch = count;
    VG_CUDA(cudaMemcpyAsync(p->h_red + off, p->d_redbuf[s] + off, sizeof(double) * fetch, cudaMemcpyDeviceToHost, p->stream));
    VG_CUDA(cudaStreamSynchronize(p->stream));
    return check_peer_fail(p);
}

int ensure_materialized(vg_problem *p)
{
    for (Ds &d : p->dss) {
        const size_t rows = (size_t)(d.n_img ? d.n_img : 1) * 2 * d.P;
        const int K = p->cams[d.cam].K;
        if (!d.d_r) VG_CUDA(cudaMalloc(&d.d_r, rows * 8));
        if (!d.d_Ja) VG_CUDA(cudaMalloc(&d.d_Ja, rows * K * 8));
        for (int e = 0; e < d.L; e++)
            if (!d.d_Je[e]) VG_CUDA(cudaMalloc(&d.d_Je[e], rows * 48));
    }
    return VG_OK;
}

}  // namespace

extern "C" {

void vg_solve_options_default(vg_solve_options *o)
{
    o->max_num_iterations = 1000;       // unified_calibration.cpp:46
    o->function_tolerance = 1e-15;      // :47
    o->gradient_tolerance = 1e-15;      // :48
    o->parameter_tolerance = 1e-15;     // :49
    o->initial_radius = 1e4;
    o->max_radius = 1e16;
    o->min_radius = 1e-32;
    o->min_relative_decrease = 1e-3;
    o->min_lm_diagonal = 1e-6;
    o->max_lm_diagonal = 1e32;
    o->jacobi_scaling = 1;
    o->max_consecutive_invalid = 5;
    o->verbose = 0;
    o->reserved = 0;
}

vg_problem *vg_problem_create(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
        cudaGetLastError();
        set_error("no CUDA device: this engine has no CPU path");
        return nullptr;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) { set_error("cudaGetDevice failed"); return nullptr; }
    if (device >= n) { set_error("device index out of range"); return nullptr; }
    vg_problem *p = new vg_problem();
    p->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&p->ev0) != cudaSuccess || cudaEventCreate(&p->ev1) != cudaSuccess) {
        set_error("CUDA stream / event creation failed");
        delete p;
        return nullptr;
    }
    p->stream = p->own_stream;
    return p;
}

void vg_problem_destroy(vg_problem *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    free_prepared(p);
    for (Tr &t : p->trs) {
        if (!t.is_global) {
            if (t.dev[1] && t.dev[1] != t.dev[0]) cudaFree(t.dev[1]);
            if (t.dev[0]) cudaFree(t.dev[0]);
        }
    }
    for (Ds &d : p->dss) {
        cudaFree(d.d_board); cudaFree(d.d_obs); cudaFree(d.d_seq_index);
        cudaFree(d.d_r); cudaFree(d.d_Ja);
        for (int e = 0; e < VG_MAX_CHAIN; e++) cudaFree(d.d_Je[e]);
    }
    for (void *q : p->peer_opened) cudaIpcCloseMemHandle(q);
    cudaFree(p->d_peer_ptrs); cudaFree(p->d_inbox); cudaFree(p->d_collect_done);
    if (p->h_peer_fail) cudaFreeHost(p->h_peer_fail);
    cudaEventDestroy(p->ev0); cudaEventDestroy(p->ev1);
    cudaStreamDestroy(p->own_stream);
    // a caller's stream (vg_problem_set_stream) may be gone by now: what the calls above made of that must not surface
    // as the error of the next launch
    cudaGetLastError();
    delete p;
}

int vg_problem_add_camera(vg_problem *p, int model, const double *value, int constant)
{
    if (!p || !value) return fail(VG_ERR_INVALID, "null argument");
    const int K = vg_model_num_params(model);
    if (K < 0) return K;
    Cam c;
    memset(&c, 0, sizeof c);
    c.model = model; c.K = K; c.constant = constant ? 1 : 0; c.shared_off = -1;
    for (int i = 0; i < K; i++) {
        c.params[i] = value[i];
        vg_model_bounds(model, i, &c.lo[i], &c.hi[i]);
    }
    cudaSetDevice(p->device);
    free_prepared(p);
    p->cams.push_back(c);
    return (int)p->cams.size() - 1;
}

int vg_problem_set_bounds(vg_problem *p, int camera, int idx, double lower, double upper)
{
    if (!p || camera < 0 || camera >= (int)p->cams.size() || idx < 0 || idx >= p->cams[camera].K)
        return fail(VG_ERR_INVALID, "vg_problem_set_bounds: bad camera or index");
    p->cams[camera].lo[idx] = lower;
    p->cams[camera].hi[idx] = upper;
    return VG_OK;
}

int vg_problem_add_transform(vg_problem *p, int is_global, int constant, int n, const double *values)
{
    if (!p || !values || n < 1 || (is_global && n != 1)) return fail(VG_ERR_INVALID, "vg_problem_add_transform: bad arguments");
    VG_CUDA(cudaSetDevice(p->device));
    free_prepared(p);
    Tr t;
    t.is_global = is_global ? 1 : 0; t.constant = constant ? 1 : 0; t.n = n;
    t.shared_off = t.pose_off = -1;
    t.glob_slot = t.seq_slot = -1;
    t.host.assign(values, values + (size_t)6 * n);
    t.fixed.assign(n, 0);
    t.dev[0] = t.dev[1] = nullptr;
    if (t.is_global) t.glob_slot = p->n_glob++;
    else {
        t.seq_slot = p->n_seq++;
        VG_CUDA(cudaMalloc(&t.dev[0], sizeof(double) * 6 * n));
        if (t.constant) t.dev[1] = t.dev[0];
        else VG_CUDA(cudaMalloc(&t.dev[1], sizeof(double) * 6 * n));
    }
    p->trs.push_back(t);
    return (int)p->trs.size() - 1;
}

int vg_problem_add_dataset(vg_problem *p, int camera, int P, const double *board,
                           int n_img, const double *obs, const int *seq_index,
                           int chain_len, const int *transform_ids, const int *status)
{
    if (!p || !board || !transform_ids || !status || (n_img > 0 && !obs)) return fail(VG_ERR_INVALID, "null argument");
    if (camera < 0 || camera >= (int)p->cams.size()) return fail(VG_ERR_INVALID, "unknown camera");
    if (chain_len < 1) return fail(VG_ERR_INVALID, "empty transform chain");
    if (chain_len > VG_MAX_CHAIN)
        return fail(VG_ERR_INVALID, "the transform chain is too long (5 transforms at max are supproted)");   // :567
    if (P < 1 || n_img < 0) return fail(VG_ERR_INVALID, "P < 1 or n_img < 0");
    int nseq = 0, seq_tr = -1;
    for (int e = 0; e < chain_len; e++) {
        if (transform_ids[e] < 0 || transform_ids[e] >= (int)p->trs.size()) return fail(VG_ERR_INVALID, "unknown transform");
        for (int f = 0; f < e; f++)
            if (transform_ids[f] == transform_ids[e]) return fail(VG_ERR_INVALID, "a transform appears twice in one chain");
        if (!p->trs[transform_ids[e]].is_global) { nseq++; seq_tr = transform_ids[e]; }
    }
    if (nseq != 1) return fail(VG_ERR_INVALID, "not one sequences in a transform chain");   // :228
    if (eval_smem_bytes(p->cams[camera].model, chain_len, P, nullptr, nullptr) < 0)
        return fail(VG_ERR_UNSUPPORTED, "board has too many points for one CTA's shared memory");
    Ds d;
    memset(static_cast<void *>(&d), 0, offsetof(Ds, seq_index));
    d.d_board = d.d_obs = nullptr; d.d_seq_index = nullptr;
    d.d_H[0] = d.d_H[1] = nullptr; d.d_r = d.d_Ja = nullptr;
    for (int e = 0; e < VG_MAX_CHAIN; e++) d.d_Je[e] = nullptr;
    d.cam = camera; d.P = P; d.n_img = n_img; d.L = chain_len;
    d.D = p->cams[camera].K + 6 * chain_len; d.W = d.D + 1; d.ne = d.W * (d.W + 1) / 2;
    d.seq_tr = seq_tr;
    d.identity_index = true;
    d.seq_index.resize(n_img);
    for (int i = 0; i < n_img; i++) {
        const int s = seq_index ? seq_index[i] : i;
        if (s < 0 || s >= p->trs[seq_tr].n) return fail(VG_ERR_INVALID, "seq_index out of range");
        if (s != i) d.identity_index = false;
        d.seq_index[i] = s;
    }
    for (int e = 0; e < chain_len; e++) { d.tr[e] = transform_ids[e]; d.status[e] = status[e]; }
    VG_CUDA(cudaSetDevice(p->device));
    free_prepared(p);
    const size_t obs_bytes = sizeof(double) * 2 * (size_t)P * (n_img ? n_img : 1);
    VG_CUDA(cudaMalloc(&d.d_board, sizeof(double) * 3 * P));
    VG_CUDA(cudaMalloc(&d.d_obs, obs_bytes));
    VG_CUDA(cudaMalloc(&d.d_seq_index, sizeof(int) * (n_img ? n_img : 1)));
    VG_CUDA(cudaMemcpy(d.d_board, board, sizeof(double) * 3 * P, cudaMemcpyHostToDevice));
    if (n_img) {
        VG_CUDA(cudaMemcpy(d.d_obs, obs, sizeof(double) * 2 * (size_t)P * n_img, cudaMemcpyHostToDevice));
        VG_CUDA(cudaMemcpy(d.d_seq_index, d.seq_index.data(), sizeof(int) * n_img, cudaMemcpyHostToDevice));
    }
    p->dss.push_back(d);
    return (int)p->dss.size() - 1;
}

int vg_problem_add_transformation_prior(vg_problem *p, int transform, int index, const double *stiffness, const double *xi_prior)
{
    if (!p || !stiffness) return fail(VG_ERR_INVALID, "null argument");
    if (transform < 0 || transform >= (int)p->trs.size()) return fail(VG_ERR_INVALID, "unknown transform");
    Tr &t = p->trs[transform];
    if (index < 0 || index >= t.n) return fail(VG_ERR_INVALID, "vg_problem_add_transformation_prior: element index out of range");
    if (!xi_prior && !t.is_global && p->prepared) {
        // the prior defaults to the element's current value: fetch the sequence if a solve has moved it
        std::vector<double> tmp((size_t)6 * t.n);
        int rc = vg_problem_get_transform(p, transform, tmp.data());
        if (rc) return rc;
    }
    cudaSetDevice(p->device);
    free_prepared(p);
    TPrior tp;
    tp.tr = transform; tp.index = index;
    memcpy(tp.stiffness, stiffness, 48);
    memcpy(tp.xi_prior, xi_prior ? xi_prior : t.host.data() + (size_t)6 * index, 48);
    p->tps.push_back(tp);
    return (int)p->tps.size() - 1;
}

int vg_problem_add_odometry(vg_problem *p, int transform, double errV, double errW, double lambda, int n, const double *odom)
{
    if (!p || !odom) return fail(VG_ERR_INVALID, "null argument");
    if (transform < 0 || transform >= (int)p->trs.size()) return fail(VG_ERR_INVALID, "unknown transform");
    const Tr &t = p->trs[transform];
    if (t.is_global) return fail(VG_ERR_INVALID, "the transform is global. Odometry must be a sequence");   // :752-755
    if (n != t.n) return fail(VG_ERR_INVALID, "vg_problem_add_odometry: one odometry reading per sequence element expected");
    if (!(lambda > 0.0)) return fail(VG_ERR_INVALID, "vg_problem_add_odometry: lambda must be positive");
    for (const Odom &o : p->odoms)
        if (o.tr == transform) return fail(VG_ERR_UNSUPPORTED, "the sequence already has odometry blocks");
    cudaSetDevice(p->device);
    free_prepared(p);
    Odom od;
    od.tr = transform; od.errV = errV; od.errW = errW; od.lambda = lambda;
    od.odom.assign(odom, odom + (size_t)6 * n);
    p->odoms.push_back(od);
    return n - 1;
}

int vg_problem_set_loss(vg_problem *p, int dataset, double a)
{
    if (!p || dataset < 0 || dataset >= (int)p->dss.size()) return fail(VG_ERR_INVALID, "bad dataset");
    if (!(a >= 0.0)) return fail(VG_ERR_INVALID, "vg_problem_set_loss: the scale must be positive (0: no loss)");
    p->dss[dataset].loss_a = a;        // a launch parameter: nothing prepared depends on it
    return VG_OK;
}

int vg_problem_set_pose_constant(vg_problem *p, int transform, int index, int constant)
{
    if (!p || transform < 0 || transform >= (int)p->trs.size()) return fail(VG_ERR_INVALID, "unknown transform");
    Tr &t = p->trs[transform];
    if (t.is_global) return fail(VG_ERR_INVALID, "vg_problem_set_pose_constant: not a sequence transform");
    if (index < 0 || index >= t.n) return fail(VG_ERR_INVALID, "vg_problem_set_pose_constant: element index out of range");
    if (p->prepared) {
        std::vector<double> tmp((size_t)6 * t.n);
        int rc = vg_problem_get_transform(p, transform, tmp.data());    // keep what a solve has reached
        if (rc) return rc;
    }
    cudaSetDevice(p->device);
    free_prepared(p);
    t.fixed[index] = constant ? 1 : 0;
    return VG_OK;
}

static int finish_peer_connect(vg_problem *p, int rank, int nranks, const std::vector<unsigned long long *> &ptrs)
{
    VG_CUDA(cudaMalloc(&p->d_collect_done, sizeof(unsigned long long)));
    VG_CUDA(cudaMemset(p->d_collect_done, 0, sizeof(unsigned long long)));
    VG_CUDA(cudaMalloc(&p->d_peer_ptrs, sizeof(unsigned long long *) * nranks));
    VG_CUDA(cudaMemcpy(p->d_peer_ptrs, ptrs.data(), sizeof(unsigned long long *) * nranks, cudaMemcpyHostToDevice));
    VG_CUDA(cudaHostAlloc(&p->h_peer_fail, sizeof(unsigned long long), cudaHostAllocMapped));
    *p->h_peer_fail = 0;
    p->rank = rank; p->nranks = nranks; p->peers = true; p->epoch = 0;
    return VG_OK;
}

int vg_problem_peer_export(vg_problem *p, void *ipc_handle_out)
{
    if (!p || !ipc_handle_out) return fail(VG_ERR_INVALID, "null argument");
    VG_CUDA(cudaSetDevice(p->device));
    if (!p->d_inbox) {
        VG_CUDA(cudaMalloc(&p->d_inbox, peer_inbox_bytes()));
        VG_CUDA(cudaMemset(p->d_inbox, 0, peer_inbox_bytes()));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == VG_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    VG_CUDA(cudaIpcGetMemHandle(&h, p->d_inbox));
    memcpy(ipc_handle_out, &h, sizeof h);
    return VG_OK;
}

int vg_problem_peer_connect(vg_problem *p, int rank, int nranks, const void *ipc_handles)
{
    if (!p || !ipc_handles || nranks < 1 || nranks > PEER_MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(VG_ERR_INVALID, "vg_problem_peer_connect: bad arguments");
    if (!p->d_inbox) return fail(VG_ERR_INVALID, "vg_problem_peer_connect: call vg_problem_peer_export first");
    if (p->peers) return fail(VG_ERR_INVALID, "vg_problem_peer_connect: already connected");
    VG_CUDA(cudaSetDevice(p->device));
    free_prepared(p);
    std::vector<unsigned long long *> ptrs(nranks, nullptr);
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { ptrs[r] = p->d_inbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(ipc_handles) + (size_t)r * sizeof h, sizeof h);
        void *q = nullptr;
        VG_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
        p->peer_opened.push_back(q);
        ptrs[r] = static_cast<unsigned long long *>(q);
    }
    p->post_at_head = getenv("VG_PEER_POST_TAIL") == nullptr;      // (developer knob: the tail posts, as between local ranks)
    return finish_peer_connect(p, rank, nranks, ptrs);
}

int vg_problem_peer_inbox(vg_problem *p, void **inbox_out, int *device_out)
{
    if (!p || !inbox_out) return fail(VG_ERR_INVALID, "null argument");
    VG_CUDA(cudaSetDevice(p->device));
    if (!p->d_inbox) {
        VG_CUDA(cudaMalloc(&p->d_inbox, peer_inbox_bytes()));
        VG_CUDA(cudaMemset(p->d_inbox, 0, peer_inbox_bytes()));
    }
    *inbox_out = p->d_inbox;
    if (device_out) *device_out = p->device;
    return VG_OK;
}

int vg_problem_peer_connect_local(vg_problem *p, int rank, int nranks, void *const *inboxes, const int *devices)
{
    if (!p || !inboxes || nranks < 1 || nranks > PEER_MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(VG_ERR_INVALID, "vg_problem_peer_connect_local: bad arguments");
    if (!p->d_inbox || inboxes[rank] != p->d_inbox) return fail(VG_ERR_INVALID, "vg_problem_peer_connect_local: inboxes[rank] is not this problem's inbox");
    if (p->peers) return fail(VG_ERR_INVALID, "vg_problem_peer_connect_local: already connected");
    VG_CUDA(cudaSetDevice(p->device));
    free_prepared(p);
    std::vector<unsigned long long *> ptrs(nranks, nullptr);
    for (int r = 0; r < nranks; r++) {
        if (!inboxes[r]) return fail(VG_ERR_INVALID, "vg_problem_peer_connect_local: null inbox");
        if (devices && devices[r] != p->device) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail_cuda(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        ptrs[r] = static_cast<unsigned long long *>(inboxes[r]);
    }
    return finish_peer_connect(p, rank, nranks, ptrs);
}

int vg_problem_set_peer_timeout(vg_problem *p, long long polls)
{
    if (!p || polls < 0) return fail(VG_ERR_INVALID, "vg_problem_set_peer_timeout: bad arguments");
    p->peer_spin_limit = polls;
    return VG_OK;
}

int vg_problem_set_allreduce(vg_problem *p, vg_allreduce_fn fn, void *ctx, int rank, int nranks)
{
    if (!p || nranks < 1 || rank < 0 || rank >= nranks) return fail(VG_ERR_INVALID, "vg_problem_set_allreduce: bad arguments");
    cudaSetDevice(p->device);
    free_prepared(p);
    if (p->peers) return fail(VG_ERR_INVALID, "vg_problem_set_allreduce: the problem exchanges over peer memory already");
    p->allreduce = fn; p->allreduce_ctx = ctx; p->rank = rank; p->nranks = fn ? nranks : 1;
    if (!fn) p->rank = 0;
    return VG_OK;
}

int vg_problem_materialize_jacobians(vg_problem *p, int enable)
{
    if (!p) return fail(VG_ERR_INVALID, "null problem");
    VG_CUDA(cudaSetDevice(p->device));
    p->materialize = enable != 0;
    if (p->materialize) return ensure_materialized(p);
    return VG_OK;
}

void *vg_problem_stream(vg_problem *p) { return p ? p->stream : nullptr; }

int vg_problem_set_stream(vg_problem *p, void *stream)
{
    if (!p) return fail(VG_ERR_INVALID, "null problem");
    VG_CUDA(cudaSetDevice(p->device));
    VG_CUDA(cudaStreamSynchronize(p->stream));
    p->stream = stream ? static_cast<cudaStream_t>(stream) : p->own_stream;
    return VG_OK;
}

int vg_problem_evaluate_async(vg_problem *p)
{
    if (!p) return fail(VG_ERR_INVALID, "null problem");
    int rc = prepare(p);
    if (rc) return rc;
    VG_CUDA(cudaSetDevice(p->device));
    const bool deferred = p->peers && p->nranks > 1 && p->n_tp + p->n_op == 0 && p->last_ds >= 0;
    rc = evaluate_set(p, p->cur, false, deferred);
    if (rc) return rc;
    if (deferred) return VG_OK;          // posted; vg_problem_fetch_reduced (or the next evaluation) forms the sum
    return exchange_segment(p, p->cur, 0, red_segE_size(p->Ks));
}

int vg_problem_fetch_reduced(vg_problem *p, double *cost, double *reduced)
{
    if (!p || !p->prepared) return fail(VG_ERR_INVALID, "nothing evaluated yet");
    VG_CUDA(cudaSetDevice(p->device));
    {
        int rc = flush_pending(p);
        if (rc) return rc;
    }
    const int Ks = p->Ks, n = red_off_model(Ks);
    VG_CUDA(cudaMemcpyAsync(p->h_red, p->d_redbuf[p->cur], sizeof(double) * n, cudaMemcpyDeviceToHost, p->stream));
    VG_CUDA(cudaStreamSynchronize(p->stream));
    {
        int rc = check_peer_fail(p);
        if (rc) return rc;
    }
    if (cost) *cost = p->h_red[red_off_cost(Ks)];
    if (reduced) memcpy(reduced, p->h_red, sizeof(double) * (Ks * Ks + Ks));
    return VG_OK;
}

int vg_problem_device_buffer(vg_problem *p, int dataset, int which, void **ptr, size_t *bytes)
{
    if (!p || dataset < 0 || dataset >= (int)p->dss.size() || !ptr) return fail(VG_ERR_INVALID, "bad dataset");
    int rc = prepare(p);
    if (rc) return rc;
    Ds &d = p->dss[dataset];
    const size_t rows = (size_t)d.n_img * 2 * d.P;
    void *q = nullptr; size_t b = 0;
    if (which == -1) { q = d.d_obs; b = rows * 8; }
    else if (which == -2) { q = d.d_H[p->cur]; b = (size_t)d.n_img * d.ne * 8; }
    else if (which == 0) { q = d.d_r; b = rows * 8; }
    else if (which == 1) { q = d.d_Ja; b = rows * p->cams[d.cam].K * 8; }
    else if (which >= 2 && which < 2 + d.L) { q = d.d_Je[which - 2]; b = rows * 48; }
    else return fail(VG_ERR_INVALID, "bad buffer selector");
    *ptr = q;
    if (bytes) *bytes = b;
    return VG_OK;
}

int vg_problem_num_shared(vg_problem *p)
{
    if (!p) return fail(VG_ERR_INVALID, "null problem");
    int rc = prepare(p);
    return rc ? rc : p->Ks;
}

int vg_problem_get_camera(vg_problem *p, int camera, double *out)
{
    if (!p || !out || camera < 0 || camera >= (int)p->cams.size()) return fail(VG_ERR_INVALID, "bad camera");
    memcpy(out, p->cams[camera].params, sizeof(double) * p->cams[camera].K);
    return VG_OK;
}

int vg_problem_set_camera(vg_problem *p, int camera, const double *value)
{
    if (!p || !value || camera < 0 || camera >= (int)p->cams.size()) return fail(VG_ERR_INVALID, "bad camera");
    memcpy(p->cams[camera].params, value, sizeof(double) * p->cams[camera].K);
    if (p->prepared) {
        VG_CUDA(cudaSetDevice(p->device));
        // source is pageable memory owned by the handle: the runtime stages it before returning
        VG_CUDA(cudaMemcpyAsync(p->cam_ptr(p->cur, camera), p->cams[camera].params, sizeof(double) * p->cams[camera].K,
                                cudaMemcpyHostToDevice, p->stream));
    }
    return VG_OK;
}

int vg_problem_get_transform(vg_problem *p, int transform, double *out)
{
    if (!p || !out || transform < 0 || transform >= (int)p->trs.size()) return fail(VG_ERR_INVALID, "bad transform");
    Tr &t = p->trs[transform];
    if (!t.is_global && p->prepared) {
        VG_CUDA(cudaSetDevice(p->device));
        VG_CUDA(cudaMemcpyAsync(t.host.data(), t.dev[p->cur], sizeof(double) * 6 * t.n, cudaMemcpyDeviceToHost, p->stream));
        VG_CUDA(cudaStreamSynchronize(p->stream));
    }
    memcpy(out, t.host.data(), sizeof(double) * 6 * t.n);
    return VG_OK;
}

int vg_problem_set_transform(vg_problem *p, int transform, const double *values)
{
    if (!p || !values || transform < 0 || transform >= (int)p->trs.size()) return fail(VG_ERR_INVALID, "bad transform");
    Tr &t = p->trs[transform];
    const bool same = (values == t.host.data());
    if (!same) memcpy(t.host.data(), values, sizeof(double) * 6 * t.n);
    if (p->prepared) {
        VG_CUDA(cudaSetDevice(p->device));
        if (t.is_global)
            VG_CUDA(cudaMemcpyAsync(p->glob_ptr(p->cur, t.glob_slot), values, 48, cudaMemcpyHostToDevice, p->stream));
        else
            VG_CUDA(cudaMemcpyAsync(t.dev[p->cur], values, sizeof(double) * 6 * t.n, cudaMemcpyHostToDevice, p->stream));
        if (!same) VG_CUDA(cudaStreamSynchronize(p->stream));
    }
    return VG_OK;
}

int vg_problem_update_observations(vg_problem *p, int dataset, const double *obs)
{
    if (!p || !obs || dataset < 0 || dataset >= (int)p->dss.size()) return fail(VG_ERR_INVALID, "bad dataset");
    Ds &d = p->dss[dataset];
    VG_CUDA(cudaSetDevice(p->device));
    VG_CUDA(cudaMemcpyAsync(d.d_obs, obs, sizeof(double) * 2 * (size_t)d.P * d.n_img, cudaMemcpyHostToDevice, p->stream));
    return VG_OK;
}

int vg_problem_update_poses(vg_problem *p, int transform, const double *values)
{
    if (!p || !values || transform < 0 || transform >= (int)p->trs.size()) return fail(VG_ERR_INVALID, "bad transform");
    Tr &t = p->trs[transform];
    if (t.is_global) return fail(VG_ERR_INVALID, "vg_problem_update_poses: not a sequence transform");
    int rc = prepare(p);
    if (rc) return rc;
    VG_CUDA(cudaSetDevice(p->device));
    VG_CUDA(cudaMemcpyAsync(t.dev[p->cur], values, sizeof(double) * 6 * t.n, cudaMemcpyHostToDevice, p->stream));
    return VG_OK;
}

int vg_problem_evaluate(vg_problem *p, double *cost, double *reduced)
{
    if (!p) return fail(VG_ERR_INVALID, "null problem");
    int rc = prepare(p);
    if (rc) return rc;
    VG_CUDA(cudaSetDevice(p->device));
    rc = evaluate_set(p, p->cur, false);
    if (rc) return rc;
    const int Ks = p->Ks;
    rc = fetch_segment(p, p->cur, 0, red_segE_size(Ks));
    if (rc) return rc;
    if (cost) *cost = p->h_red[red_off_cost(Ks)];
    if (reduced) memcpy(reduced, p->h_red, sizeof(double) * (Ks * Ks + Ks));
    return VG_OK;
}

int vg_problem_residuals(vg_problem *p, int dataset, double *r)
{
    if (!p || !r || dataset < 0 || dataset >= (int)p->dss.size()) return fail(VG_ERR_INVALID, "bad dataset");
    int rc = prepare(p);
    if (rc) return rc;
    VG_CUDA(cudaSetDevice(p->device));
    Ds &d = p->dss[dataset];
    const size_t bytes = sizeof(double) * 2 * (size_t)d.P * (d.n_img ? d.n_img : 1);
    double *dr = nullptr;
    VG_CUDA(cudaMalloc(&dr, bytes));
    EvalArgs a;
    memset(&a, 0, sizeof a);
    a.intr = p->cam_ptr(p->cur, d.cam);
    a.board = d.d_board; a.obs = d.d_obs;
    a.seq_index = d.identity_index ? nullptr : d.d_seq_index;
    for (int e = 0; e < d.L; e++) {
        const Tr &t = p->trs[d.tr[e]];
        a.xi[e] = t.is_global ? p->glob_ptr(p->cur, t.glob_slot) : t.dev[p->cur];
        a.xi_stride[e] = t.is_global ? 0 : 6;
        a.inverse[e] = d.status[e] == VG_TRANSFORM_INVERSE;
    }
    a.r = dr; a.n_img = d.n_img; a.P = d.P;
    cudaError_t e = launch_eval(p->cams[d.cam].model, d.L, a, p->stream, &launch_counter());
    if (e == cudaSuccess && d.n_img) e = cudaMemcpyAsync(r, dr, sizeof(double) * 2 * (size_t)d.P * d.n_img, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(dr);
    if (e != cudaSuccess) return fail_cuda(e, "vg_problem_residuals");
    return VG_OK;
}

// The LM loop of the plain structure with its control state on the device (vg_lm_dev.cuh): the host queues the passes
// (factorisation + Schur terms, reduced solve + back-substitution, the candidate's evaluation whose last thread takes the
// decision), one pass ahead of the last record it has seen, and never waits for the stream.
static int solve_on_device(vg_problem *p, const vg_solve_options &o, vg_solve_summary *sum, const double t_start)
{
    const int Ks = p->Ks, NP = p->n_pose, A = p->cur, C = A ^ 1;
    SolverLaunch sl{p->stream, &launch_counter()};
    const bool multi = p->peers && p->nranks > 1;
    const unsigned long long base = p->lm_seq_base;
    LmState &s0 = *p->h_lm_stage;
    memset(&s0, 0, sizeof s0);
    s0.limits = (0 >= o.max_num_iterations || o.initial_radius < o.min_radius) ? 1 : 0;
    s0.init_scale = 1;
    s0.records = base; s0.published = base;
    s0.epoch = p->epoch + 2;                 // (the first evaluation's exchange takes p->epoch + 1)
    s0.radius = o.initial_radius; s0.decrease_factor = 2.0;
    s0.opt = LmOptions{o.gradient_tolerance, o.function_tolerance, o.parameter_tolerance, o.min_relative_decrease, o.min_radius,
                       o.max_radius, o.max_num_iterations, o.max_consecutive_invalid};
#ifdef VG_LM_STAMPS
    for (auto &row : s0.stamp) for (int k = 0; k < 6; k++) row[k] = (k & 1) ? 0ull : ~0ull;
#endif
    // (no upload: block 0 of the first evaluation fetches *h_lm_stage itself, EvalArgs::lm_init)
    std::atomic_thread_fence(std::memory_order_release);
    // (several ranks: the first evaluation's exchange sums segment E in place, the three pose sums of set A included --
    // never read, but they must not pile up from solve to solve)
    if (multi) VG_CUDA(cudaMemsetAsync(p->d_redbuf[A] + red_off_model(Ks), 0, 3 * sizeof(double), p->stream));
    p->lm_set_a = A;
    p->lm_eval_mode = 1;
    int rc = evaluate_set(p, A, false);
    p->lm_eval_mode = 0;
    if (rc) return rc;

    const SolveArgs sa{(int)p->slab_doubles, p->nranks, p->d_redbuf[A], p->d_redbuf[C], p->d_slab[A], p->d_slab[C], p->d_delta,
                       p->d_sh_off, p->d_sh_lo, p->d_sh_hi, p->d_scale_a};
    const DatasetDesc &hd = p->h_desc[A][p->fast_ds];
    FastDesc fd;
    memset(&fd, 0, sizeof fd);
    fd.H = hd.H; fd.ne = hd.ne; fd.W = hd.W; fd.pose_col = hd.pose_col; fd.n_sl = hd.n_sl;
    for (int q = 0; q < hd.n_sl; q++) { fd.sl_col[q] = hd.sl_col[q]; fd.sl_idx[q] = hd.sl_idx[q]; }
    const Tr &ft = p->trs[p->fast_tr];
    Ds &fds = p->dss[p->fast_ds];
    const FastLm flm{p->d_lm, fds.d_H[C], ft.dev[A], ft.dev[C], p->d_slab[A], p->d_slab[C], (int)p->slab_doubles,
                     p->d_redbuf[A], p->d_redbuf[C], red_segE_size(Ks), p->h_lm_ring, p->h_lm_final};
    const LmConsts lm{o.initial_radius, o.min_lm_diagonal, o.max_lm_diagonal, 1, o.jacobi_scaling};   // (radius, init_scale: LmState's)
    const long long max_passes = (long long)o.max_num_iterations + 1;     // every pass ends the solve or counts an iteration
    long long queued = 0;
    auto queue_pass = [&]() -> int {
        PeerCtx pcx;
        const PeerCtx *pcp = nullptr;
        if (multi) { pcx = p->peer_ctx(0); pcp = &pcx; }          // (exchange numbers: LmState::epoch)
        cudaError_t ce = launch_fast_step(fd, NP, Ks, p->d_scale, lm, p->d_ws, p->d_fast_scratch, p->d_fast_tickets, p->d_fail, sa,
                                          ft.dev[A], ft.dev[C], true, sl, nullptr, nullptr, pcp, &flm, p->fast_shared);
        if (ce != cudaSuccess) return fail_cuda(ce, "fast LM step");
        p->lm_eval_mode = 2;
        const int r = evaluate_set(p, C, false);
        p->lm_eval_mode = 0;
        queued++;
        return r;
    };
    static const bool tl = getenv("VG_LM_TIMELINE") != nullptr;      // developer knob: when the host saw and did what
    std::vector<double> t_seen, t_queued;
    if (tl) fprintf(stderr, "[vg lm] setup + first evaluation queued at %.1f us\n", (now_s() - t_start) * 1e6);
    rc = queue_pass();
    if (rc) return rc;
    if (tl) fprintf(stderr, "[vg lm] pass 1 queued at %.1f us\n", (now_s() - t_start) * 1e6);

    LmRecord rec;
    memset(&rec, 0, sizeof rec);
    for (long long n = 0;; n++) {
        // record n: the first evaluation (n = 0), then the decision of pass n
        const unsigned long long want = base + (unsigned long long)n + 1;
        volatile unsigned long long *seq = &p->h_lm_ring[(base + (unsigned long long)n) % LM_RING].seq;
        for (unsigned long long spins = 1; *seq != want; spins++) {
            if ((spins & 0x3FFF) == 0) {
                const cudaError_t qe = cudaStreamQuery(p->stream);
                if (qe == cudaSuccess) {
                    if (*seq == want) break;
                    return fail(VG_ERR_CUDA, "LM loop on the device: the stream drained without a decision");
                }
                if (qe != cudaErrorNotReady) return fail_cuda(qe, "LM step");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        memcpy(&rec, const_cast<const LmRecord *>(&p->h_lm_ring[(base + (unsigned long long)n) % LM_RING]), sizeof rec);
        if (n == 0) {
            sum->initial_cost = rec.cost;
        } else if (o.verbose && rec.valid && !(rec.done == 1 + 2)) {
            printf("%4d  cost %.12e  new %.12e  rho %.3e  radius %.3e  |step| %.3e\n", rec.iter, rec.prev_cost, rec.new_cost,
                   rec.rho, rec.prev_radius, rec.step_norm);
        }
        if (tl) t_seen.push_back((now_s() - t_start) * 1e6);
        if (rec.done) break;
        // pass n + 1 is queued (and possibly running); queue pass n + 2 behind it
        // (a pass queued past the end of the solve only publishes the last record; max_passes + 1 passes cannot all run)
        if (queued > max_passes + 1) return fail(VG_ERR_CUDA, "LM loop on the device: no termination within the iteration limit");
        rc = queue_pass();
        if (rc) return rc;
        if (tl) t_queued.push_back((now_s() - t_start) * 1e6);
    }
    if (tl)
        for (size_t i = 0; i < t_seen.size(); i++)
            fprintf(stderr, "[vg lm] record %zu seen at %.1f us, next pass queued by %.1f us\n", i, t_seen[i],
                    i < t_queued.size() ? t_queued[i] : 0.0);
#ifdef VG_LM_STAMPS
    {
        cudaStreamSynchronize(p->stream);
        cudaMemcpy(&s0, p->d_lm, sizeof s0, cudaMemcpyDeviceToHost);
        const char *nm[3] = {"factor", "backsub", "eval"};
        unsigned long long prev_end = 0;
        for (long long r_ = 1; r_ <= rec.iter && r_ < 64; r_++) {
            const unsigned long long *w = s0.stamp[(base + r_) & 63];
            fprintf(stderr, "[vg lm] pass %lld:", r_);
            for (int k = 0; k < 3; k++) {
                fprintf(stderr, "  gap %.1f %s %.1f", prev_end ? (double)(long long)(w[2 * k] - prev_end) * 1e-3 : 0.0, nm[k],
                        (double)(long long)(w[2 * k + 1] - w[2 * k]) * 1e-3);
                prev_end = w[2 * k + 1];
            }
            fprintf(stderr, "  (us)\n");
        }
        fprintf(stderr, "[vg lm] clock64 in the reduced solve:");
        for (int i = 1; i < 14; i++) fprintf(stderr, " %lld", s0.clk[i] - s0.clk[i - 1]);
        fprintf(stderr, "\n");
        for (int k = 0; k < 2; k++) {
            fprintf(stderr, "[vg lm] %s, block 10, since the kernel's first block started:", nm[k]);
            const unsigned long long t0 = s0.stamp[(base + rec.iter) & 63][2 * k];
            for (int i = 0; i < 10; i++) fprintf(stderr, " %.1f", (double)(long long)(s0.phase[k][i] - t0) * 1e-3);
            fprintf(stderr, "  end %.1f (us)\n", (double)(long long)(s0.stamp[(base + rec.iter) & 63][2 * k + 1] - t0) * 1e-3);
        }
    }
#endif
    p->lm_seq_base = base + LM_RING * ((unsigned long long)queued / LM_RING + 2);
    p->exchanged_in_kernel = false;
    if (multi) p->epoch = rec.epoch - 1;
    // where the final point lives: set C if the last pass accepted its candidate (nothing has copied it over), else A;
    // its packed blocks in the buffer the evaluations alternated to
    const int fin = rec.migrate ? C : A, hset = rec.hcur ? C : A;
    if (hset != fin)
        VG_CUDA(cudaMemcpyAsync(fds.d_H[fin], fds.d_H[hset], sizeof(double) * (size_t)fds.n_img * fds.ne, cudaMemcpyDeviceToDevice, p->stream));
    p->cur = fin;
    {
        const size_t glob0 = p->cams.size() * CAM_STRIDE;
        const double *cs = p->h_lm_final;
        if (rec.num_successful > 0) {
            for (size_t ci = 0; ci < p->cams.size(); ci++)
                memcpy(p->cams[ci].params, cs + ci * CAM_STRIDE, sizeof(double) * p->cams[ci].K);
            for (Tr &t : p->trs)
                if (t.is_global) memcpy(t.host.data(), cs + glob0 + (size_t)t.glob_slot * 6, 48);
        }
    }
    rc = check_peer_fail(p);
    if (rc) return rc;
    sum->iterations = rec.iter;
    sum->final_cost = rec.cost;
    sum->termination = rec.done - 1;
    sum->num_successful = rec.num_successful;
    sum->num_unsuccessful = rec.num_unsuccessful;
    sum->seconds_total = now_s() - t_start;
    sum->seconds_evaluate = 0.0;             // (no events between the launches of this loop)
    sum->num_evaluations = rec.iter + 1;
    return VG_OK;
}

int vg_problem_solve(vg_problem *p, const vg_solve_options *opt, vg_solve_summary *sum)
{
    if (!p || !sum) return fail(VG_ERR_INVALID, "null argument");
    vg_solve_options o;
    if (opt) o = *opt; else vg_solve_options_default(&o);
    const double t_start = now_s();
    int rc = prepare(p);
    if (rc) return rc;
    struct SolveScope {
        vg_problem *p;
        explicit SolveScope(vg_problem *q) : p(q) { p->in_solve = true; }
        ~SolveScope() { p->in_solve = false; p->after_solve = true; }
    } solve_scope(p);
    VG_CUDA(cudaSetDevice(p->device));
    memset(sum, 0, sizeof *sum);
    p->eval_ms = 0; p->n_eval = 0;
    const int Ks = p->Ks, NP = p->n_pose;
    const int segE = red_segE_size(Ks), offS = red_off_S(Ks), segS = red_segS_size(Ks, p->nranks);
    const int off_out = red_size(Ks, p->nranks), off_slab = off_out + SOLVE_OUT;
    SolverLaunch sl{p->stream, &launch_counter()};
    const size_t glob0 = p->cams.size() * CAM_STRIDE;

    auto collect_eval_time = [&]() {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p->ev0, p->ev1) == cudaSuccess) p->eval_ms += ms;
    };

    // box bounds of the shared parameters (cameras: the model's or vg_problem_set_bounds'; transforms: none)
    {
        std::vector<double> lo(Ks ? Ks : 1, -1e300), hi(Ks ? Ks : 1, 1e300);
        for (const Cam &c : p->cams)
            if (c.shared_off >= 0)
                for (int k = 0; k < c.K; k++) { lo[c.shared_off + k] = c.lo[k]; hi[c.shared_off + k] = c.hi[k]; }
        // h_up is pinned and at least 2 Ks doubles long; nothing else uses it while a solve runs; it keeps what the device
        // has (prepare() clears bounds_on_device)
        if (Ks <= FAST_MAX_KS) {
            memset(&p->fast_shared, 0, sizeof p->fast_shared);
            for (int j = 0; j < Ks; j++) { p->fast_shared.off[j] = p->h_sh_off[j]; p->fast_shared.lo[j] = lo[j]; p->fast_shared.hi[j] = hi[j]; }
        }
        const bool same = p->bounds_on_device && memcmp(p->h_up, lo.data(), sizeof(double) * Ks) == 0 &&
                          memcmp(p->h_up + Ks, hi.data(), sizeof(double) * Ks) == 0;
        if (!same) {
            memcpy(p->h_up, lo.data(), sizeof(double) * Ks);
            memcpy(p->h_up + Ks, hi.data(), sizeof(double) * Ks);
        }
        if (Ks && !same) {
            p->bounds_on_device = true;
            VG_CUDA(cudaMemcpyAsync(p->d_sh_lo, p->h_up, sizeof(double) * Ks, cudaMemcpyHostToDevice, p->stream));
            VG_CUDA(cudaMemcpyAsync(p->d_sh_hi, p->h_up + Ks, sizeof(double) * Ks, cudaMemcpyHostToDevice, p->stream));
        }
    }

    // several ranks: the two-launch step only if EVERY rank's share has the plain structure (the ranks must issue the same
    // sequence of exchanges): one exchange of a flag per solve
    bool fast_everywhere = p->fast_ds >= 0;
    if (p->peers && p->nranks > 1) {
        const double mine = p->fast_ds >= 0 ? 1.0 : 0.0;
        VG_CUDA(cudaMemcpyAsync(p->d_agree, &mine, sizeof(double), cudaMemcpyHostToDevice, p->stream));
        VG_CUDA(cudaStreamSynchronize(p->stream));
        cudaError_t ae = launch_peer_exchange(p->d_agree, 1, p->next_peer_ctx(), sl);
        if (ae != cudaSuccess) return fail_cuda(ae, "peer exchange");
        double all = 0.0;
        VG_CUDA(cudaMemcpyAsync(&all, p->d_agree, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        VG_CUDA(cudaStreamSynchronize(p->stream));
        rc = check_peer_fail(p);
        if (rc) return rc;
        fast_everywhere = all == (double)p->nranks;
    }
    // the plain structure: the whole loop with its decisions on the device (developer knobs: VG_LM_HOSTLOOP -- the host
    // decides every iteration, as for every other structure; VG_LM_NOFAST -- the general kernels; VG_LM_TRACE)
    if (fast_everywhere && !getenv("VG_LM_HOSTLOOP") && !getenv("VG_LM_NOFAST") && !getenv("VG_LM_TRACE") && !getenv("VG_LM_BLIND"))
        return solve_on_device(p, o, sum, t_start);
    // iteration 0: evaluate at the starting point
    for (int s_ = 0; s_ < 2; s_++)
        VG_CUDA(cudaMemsetAsync(p->d_redbuf[s_] + red_off_model(Ks), 0, 3 * sizeof(double), p->stream));
    rc = evaluate_set(p, p->cur, true);
    if (rc) return rc;
    rc = fetch_segment(p, p->cur, 0, segE);
    if (rc) return rc;
    collect_eval_time();
    double cost = p->h_red[red_off_cost(Ks)];
    sum->initial_cost = cost;

    double radius = o.initial_radius, decrease_factor = 2.0;
    int invalid_run = 0, iter = 0;
    bool init_scale = true;
    sum->termination = 3;
    static const bool trace = getenv("VG_LM_TRACE") != nullptr;      // developer knob: host timeline of the loop
    cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (trace) for (auto &e : tev) cudaEventCreate(&e);
    auto mark = [&](int i) { if (trace) cudaEventRecord(tev[i], p->stream); };
    double t_iter = now_s();
    if (trace) fprintf(stderr, "[vg lm] start + first evaluation: %.1f us\n", (t_iter - t_start) * 1e6);

    for (;;) {
        // One host synchronisation per iteration: the per-pose damped factorisation + Schur terms at the current
        // radius, the reduced solve, the back-substitution and the candidate's evaluation are queued back to back;
        // what the decisions below need comes back in one copy.  When the iteration or radius limit is already
        // reached only the gradient test is still due (Ceres tests it first), so the step itself is not queued.
        const bool limits = iter >= o.max_num_iterations || radius < o.min_radius;
        const int cand = p->cur ^ 1;
        unsigned long long polled = 0;          // != 0: this iteration's results arrive through p->h_poll
        LmConsts lm{radius, o.min_lm_diagonal, o.max_lm_diagonal, init_scale ? 1 : 0, o.jacobi_scaling};
        cudaError_t ce = cudaSuccess;
        mark(0);
        static const bool nofast = getenv("VG_LM_NOFAST") != nullptr;   // developer knob: the general kernels everywhere
        // developer knob VG_LM_BLIND=M: the third iteration's launches are queued M times over without waiting for anything
        // in between (the same step, recomputed): device time of an iteration with the host out of the loop
        static const int blind = getenv("VG_LM_BLIND") ? atoi(getenv("VG_LM_BLIND")) : 0;
        const int reps = (blind > 1 && iter == 2 && fast_everywhere) ? blind : 1;
        cudaEvent_t bev[2] = {nullptr, nullptr};
        if (reps > 1) { cudaEventCreate(&bev[0]); cudaEventCreate(&bev[1]); cudaEventRecord(bev[0], p->stream); }
        for (int rep = 0; rep < reps; rep++)
        if (fast_everywhere && !nofast) {
            // the plain structure: factorisation + Schur terms, then reduced solve + back-substitution, two launches
            const SolveArgs sa{(int)p->slab_doubles, p->nranks, p->d_redbuf[p->cur], p->d_redbuf[cand], p->d_slab[p->cur],
                               p->d_slab[cand], p->d_delta, p->d_sh_off, p->d_sh_lo, p->d_sh_hi, p->d_scale_a};
            const DatasetDesc &hd = p->h_desc[p->cur][p->fast_ds];
            FastDesc fd;
            memset(&fd, 0, sizeof fd);
            fd.H = hd.H; fd.ne = hd.ne; fd.W = hd.W; fd.pose_col = hd.pose_col; fd.n_sl = hd.n_sl;
            for (int q = 0; q < hd.n_sl; q++) { fd.sl_col[q] = hd.sl_col[q]; fd.sl_idx[q] = hd.sl_idx[q]; }
            const Tr &ft = p->trs[p->fast_tr];
            PeerCtx pcx;
            const PeerCtx *pcp = nullptr;
            if (p->peers && p->nranks > 1) { pcx = p->next_peer_ctx(); pcp = &pcx; }
            ce = launch_fast_step(fd, NP, Ks, p->d_scale, lm, p->d_ws, p->d_fast_scratch, p->d_fast_tickets, p->d_fail, sa,
                                  ft.dev[p->cur], ft.dev[cand], !limits, sl, trace ? tev[1] : nullptr, p->h_poll, pcp, nullptr,
                                  p->fast_shared);
            if (ce != cudaSuccess) return fail_cuda(ce, "fast LM step");
            init_scale = false;
            mark(2); mark(3);
            if (!limits) {
                static const bool nopoll = getenv("VG_LM_NOPOLL") != nullptr;     // developer knob: copy + stream sync instead
                if (!nopoll && !trace && rep == reps - 1) { p->poll_seq = ++p->poll_counter; polled = p->poll_seq; }
                rc = evaluate_set(p, cand, !polled && reps == 1);
                if (rc) return rc;
            }
            mark(4);
            if (reps > 1 && rep == reps - 2) cudaEventRecord(bev[1], p->stream);
        } else {
        if (p->n_seg > 0)      // coupled / constant elements first: their rows of ws, max |g| per segment
            ce = launch_chain_factor(p->d_desc[p->cur], Ks, p->d_pose_start, p->d_contrib_ds, p->d_contrib_img, p->d_scale, lm,
                                     p->d_ws, p->chain_tables(p->cur), p->d_partial + pose_factor_blocks(NP), p->d_fail, sl);
        // reduced system  (A + D_a - S_red) delta_a = -(g_a - v_red), candidate shared parameters Pi(x + delta_a): on one
        // rank in the tail of the pose factorisation kernel, else in its own launch after the exchange of S_red, v_red
        const SolveArgs sa{(int)p->slab_doubles, p->nranks, p->d_redbuf[p->cur], p->d_redbuf[cand], p->d_slab[p->cur],
                           p->d_slab[cand], p->d_delta, p->d_sh_off, p->d_sh_lo, p->d_sh_hi, p->d_scale_a};
        static const bool nofuse = getenv("VG_LM_NOFUSE") != nullptr;   // developer knob: the tail kernels on their own
        const bool fuse = p->nranks == 1 && NP > 0 && !nofuse;
        if (ce == cudaSuccess)
            ce = launch_pose_schur(p->d_desc[p->cur], NP, Ks, p->d_pose_start, p->d_contrib_ds, p->d_contrib_img,
                                   p->d_scale, lm, p->d_ws, p->d_partial, p->partial_doubles, p->d_redbuf[p->cur],
                                   p->d_fail, p->rank, p->nranks, sl, p->d_mask, p->n_seg, fuse ? &sa : nullptr,
                                   fuse ? p->d_solver_tickets : nullptr);
        if (ce != cudaSuccess) return fail_cuda(ce, "pose_schur");
        mark(1);
        if (!fuse) {
            rc = exchange_segment(p, p->cur, offS, segS);
            if (rc) return rc;
            ce = launch_reduced_solve(Ks, sa, lm, sl);
            if (ce != cudaSuccess) return fail_cuda(ce, "reduced_solve");
        }
        init_scale = false;
        mark(2);
        if (!limits) {
            // candidate poses + model-decrease / norm partial sums (summed by the kernel's last block unless chain
            // segments add rows of their own), then the evaluation there
            const bool fuse_b = NP > 0 && p->n_seg == 0;
            ce = launch_pose_backsub(NP, Ks, p->d_delta, p->d_seq_ptr[p->cur], p->d_seq_ptr[cand], p->d_pose_seq,
                                     p->d_pose_local, p->d_ws, p->d_partial, p->partial_doubles, p->d_redbuf[cand], sl, p->d_mask,
                                     p->d_chain_w, fuse_b ? p->d_solver_tickets + 1 : nullptr);
            const int nb_b = pose_backsub_blocks(NP);
            if (ce == cudaSuccess && p->n_seg > 0)
                ce = launch_chain_backsub(Ks, p->d_seq_ptr[p->cur], p->d_seq_ptr[cand], p->d_pose_seq, p->d_pose_local, p->d_ws,
                                          p->chain_tables(p->cur), p->d_partial + 3 * (size_t)nb_b, sl);
            if (ce == cudaSuccess && !fuse_b) ce = launch_finalize_backsub(Ks, nb_b + p->n_seg, p->d_partial, p->d_redbuf[cand], sl);
            if (ce != cudaSuccess) return fail_cuda(ce, "pose_backsub");
            mark(3);
            rc = evaluate_set(p, cand, true);
            if (rc) return rc;
            mark(4);
        }
        }
        if (reps > 1) {
            cudaEventSynchronize(bev[1]);
            float ms = 0;
            cudaEventElapsedTime(&ms, bev[0], bev[1]);
            fprintf(stderr, "[vg lm] blind: %d iterations queued back to back: %.2f us each on the device\n", reps - 1,
                    ms * 1e3 / (reps - 1));
            cudaEventDestroy(bev[0]); cudaEventDestroy(bev[1]);
        }
        const double t_queued = trace ? now_s() : 0.0;
        if (polled) {
            // the kernels wrote what the decisions below read straight into host-mapped memory; the evaluation's last
            // CTA raised the flag after its cost (a finished stream without the flag cannot happen: copy as a last resort)
            volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(p->h_poll + FAST_HOST_FLAG);
            bool seen = false;
            for (unsigned long long spins = 1; !(seen = *flag == polled); spins++) {
                if ((spins & 0xFFF) == 0) {
                    const cudaError_t qe = cudaStreamQuery(p->stream);
                    if (qe == cudaSuccess) { seen = *flag == polled; break; }
                    if (qe != cudaErrorNotReady) return fail_cuda(qe, "LM step");
                }
            }
            p->exchanged_in_kernel = false;      // (segment E was summed across the ranks by the evaluation kernel itself)
            if (seen) {
                std::atomic_thread_fence(std::memory_order_acquire);
                const double *hp = p->h_poll;
                for (int i = 0; i < SOLVE_OUT; i++) p->h_red[off_out + i] = hp[i];
                for (int i = 0; i < 4; i++) p->h_red[red_off_cost(Ks) + i] = hp[FAST_HOST_EVAL + i];
                for (size_t i = 0; i < p->slab_doubles; i++) p->h_red[off_slab + i] = hp[FAST_HOST_SLAB + i];
            } else {
                polled = 0;
            }
        }
        if (!polled) {
            rc = fetch_segment(p, cand, 0, limits ? 0 : segE, p->red_doubles);
            if (rc) return rc;
        }
        if (trace) {
            const double t_now = now_s();
            fprintf(stderr, "[vg lm] iteration %d: queued in %.1f us, waited %.1f us\n", iter + 1, (t_queued - t_iter) * 1e6,
                    (t_now - t_queued) * 1e6);
            t_iter = t_now;
            if (!limits) {
                float a = 0, b = 0, c = 0, d = 0;
                cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[1], tev[2]);
                cudaEventElapsedTime(&c, tev[2], tev[3]); cudaEventElapsedTime(&d, tev[3], tev[4]);
                fprintf(stderr, "[vg lm]   device: pose factor + Schur terms %.1f us, reduced solve %.1f us, back-substitution %.1f us, "
                        "evaluation %.1f us\n", a * 1e3, b * 1e3, c * 1e3, d * 1e3);
            }
        }
        const double *so = p->h_red + off_out;
        // gradient tolerance: max-norm of the projected gradient at the current point
        if (so[0] <= o.gradient_tolerance) { sum->termination = 1; break; }
        if (iter >= o.max_num_iterations) { sum->termination = 3; break; }
        if (radius < o.min_radius) { sum->termination = 4; break; }
        iter++;
        if (!polled) collect_eval_time();      // (a polled iteration records no events: seconds_evaluate covers the others)

        bool ok = so[1] == 0.0 && so[2] != 0.0;      // every pose block and the reduced system factorised
        double model_change = 0, step2 = 0, x2 = 0;
        const double new_cost = p->h_red[red_off_cost(Ks)];
        if (ok) {
            model_change = -(so[3] + 0.5 * so[4] + p->h_red[red_off_model(Ks)]);
            step2 = so[5] + p->h_red[red_off_model(Ks) + 1];
            x2 = so[6] + p->h_red[red_off_model(Ks) + 2];
            if (!(model_change > 0.0)) ok = false;
        }
        if (!ok) {
            // invalid step (linear solve failed or the model predicts no decrease)
            invalid_run++;
            sum->num_unsuccessful++;
            if (invalid_run >= o.max_consecutive_invalid) { sum->termination = 5; break; }
            radius /= decrease_factor; decrease_factor *= 2.0;
            continue;
        }
        invalid_run = 0;
        if (std::sqrt(step2) <= o.parameter_tolerance * (std::sqrt(x2) + o.parameter_tolerance)) {
            sum->termination = 2;   // candidate discarded, as Ceres stops before taking the step
            break;
        }
        const double rho = (cost - new_cost) / model_change;
        if (o.verbose)
            printf("%4d  cost %.12e  new %.12e  rho %.3e  radius %.3e  |step| %.3e\n", iter, cost, new_cost, rho,
                   radius, std::sqrt(step2));
        if (rho > o.min_relative_decrease) {
            const double cost_change = cost - new_cost, old_cost = cost;
            p->cur = cand;
            const double *cs = p->h_red + off_slab;     // the candidate slab reduced_solve left
            for (size_t ci = 0; ci < p->cams.size(); ci++)
                memcpy(p->cams[ci].params, cs + ci * CAM_STRIDE, sizeof(double) * p->cams[ci].K);
            for (Tr &t : p->trs)
                if (t.is_global) memcpy(t.host.data(), cs + glob0 + (size_t)t.glob_slot * 6, 48);
            cost = new_cost;
            sum->num_successful++;
            radius = radius / std::fmax(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3.0));
            if (radius > o.max_radius) radius = o.max_radius;
            decrease_factor = 2.0;
            if (std::fabs(cost_change) <= o.function_tolerance * old_cost) { sum->termination = 0; break; }
        } else {
            sum->num_unsuccessful++;
            radius /= decrease_factor; decrease_factor *= 2.0;
        }
    }
    // Nothing to copy back: the last fetch synchronised the stream; the candidate set's slab and poses are rewritten
    // by every step before they are used; sequence poses stay on the device until vg_problem_get_transform asks for
    // them (or the problem's structure changes: free_prepared).
    sum->iterations = iter;
    sum->final_cost = cost;
    sum->seconds_total = now_s() - t_start;
    if (trace) fprintf(stderr, "[vg lm] epilogue: %.1f us\n", (now_s() - t_iter) * 1e6);
    sum->seconds_evaluate = p->eval_ms * 1e-3;
    sum->num_evaluations = p->n_eval;
    return VG_OK;
}

}  // extern "C"
